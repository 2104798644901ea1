// Backward of the fused link scorer under the nll read-out, exploiting that d out is one-hot per row.
//
//   out = act(Z @ W0^T + b0),  Z[p] = h_i[src[p]] * h_j[dst[p]]                     (LLP.py:105-115)
//   loss = -(1/P) sum_p out[p, t_p]                                                  (F.nll_loss, LLP.py:235)
//
// d out[p, c] = -g/P for c == t_p and 0 elsewhere, hence G = d out * act'(out) has ONE non-zero per row,
// g_p = -g/P * act'(out[p, t_p]), and the two GEMMs of the dense backward collapse to
//   dZ[p, :]      = g_p * W0[t_p, :]                       (a scaled row of the weight, no contraction left)
//   dh_i[src[p]] += dZ[p] * h_j[dst[p]],   dh_j[dst[p]] += dZ[p] * h_i[src[p]]
//   dW0[c, :]     = sum_{p : t_p == c} g_p * Z[p, :],      db0[c] = sum_{p : t_p == c} g_p
// -- O(P C) gather / scatter work bounded by the same atomics traffic as the tensor-core version's epilogue, instead of
// 4 P C Hd flops.  Same result as msha_score_mlp_nll_bwd (which multiplies the zeros out on the tensor cores); pairs whose
// activation is flat at t_p (relu inactive) are skipped altogether.
//
// Mapping: a warp owns `chunk` consecutive positions of `order` (pairs sorted by label, so label runs are long: the
// dW0 / db0 partial sums live in registers and are flushed with vector atomics when the label changes and at the end of
// the chunk).  Indices, labels and g_p of 32 pairs are fetched by one coalesced batch of loads (lane l <- pair l) and
// broadcast with shuffles; each lane then owns 4 NV channels of the 128-bit row gathers.
#include "common.cuh"

namespace {

__device__ __forceinline__ float act_grad_out(float y, int act, float slope) {   // as dense_kernels.cu: from the output
    switch (act) {
        case 1: return y > 0.f ? 1.f : y + 1.f;                  // elu
        case 2: return y > 0.f ? 1.f : 0.f;                      // relu
        case 3: return y > 0.5f ? y * (1.f - y) : 0.f;           // sigmoid(relu(x))
        case 4: return y > 0.f ? 1.f : slope;                    // leaky relu
        case 5: return y * (1.f - y);                            // sigmoid
        default: return 1.f;
    }
}

template <int NV>
__device__ __forceinline__ void flush_label(float4 (&acc)[NV], float& gsum, int label, int lane, int C, int C4,
                                            float* __restrict__ dW0, float* __restrict__ db0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c4 = lane + 32 * v;
        if (c4 < C4) atomicAdd(reinterpret_cast<float4*>(dW0 + (int64_t)label * C + 4 * c4), acc[v]);
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (lane == 0) atomicAdd(db0 + label, gsum);
    gsum = 0.f;
}

template <int NV>
__device__ __forceinline__ void flush_rows(float4 (&acc)[NV], float* __restrict__ row, int lane, int C4) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c4 = lane + 32 * v;
        if (c4 < C4) atomicAdd(reinterpret_cast<float4*>(row + 4 * c4), acc[v]);
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int NV>
__global__ void __launch_bounds__(256)
score_nll_sparse_bwd_kernel(const uint32_t* __restrict__ order, const int64_t* __restrict__ target,
                            const float* __restrict__ gout, const float* __restrict__ out, int64_t ldo,
                            const float* __restrict__ hi, const float* __restrict__ hj, const int64_t* __restrict__ src,
                            const int64_t* __restrict__ dst, int64_t P, int C, int Hd, const float* __restrict__ W0, int act,
                            float slope, int chunk, float* __restrict__ dhi, float* __restrict__ dhj,
                            float* __restrict__ dW0, float* __restrict__ db0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t q0 = warp * chunk;
    if (q0 >= P) return;
    const int64_t q1 = q0 + chunk < P ? q0 + chunk : P;
    const int C4 = C >> 2;
    const float scale = -(*gout) / (float)P;
    float4 acc[NV], w0[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = w0[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = -1;
    float gsum = 0.f;
    // positives arrive in CSR order (runs of equal src, and the label sort is stable): their dh_i contributions are summed
    // in registers and scattered once per run
    float4 run_hi[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) run_hi[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t run_a = -1;
    for (int64_t qb = q0; qb < q1; qb += 32) {
        // batch phase: lane l fetches everything scalar about pair qb + l
        const int64_t q = qb + lane;
        int64_t a = 0, b = 0;
        int t = -1;
        float g = 0.f;
        if (q < q1) {
            const int64_t p = order ? (int64_t)order[q] : q;
            const int64_t tt = target[p];
            if (tt >= 0 && tt < Hd) {                     // out-of-range labels contribute nothing (as the dense path)
                t = (int)tt;
                g = scale * act_grad_out(out[p * ldo + tt], act, slope);
                a = src ? src[p] : p;
                b = dst ? dst[p] : p;
            }
        }
        const int n = (int)(q1 - qb < 32 ? q1 - qb : 32);
        for (int j = 0; j < n; ++j) {
            const float gj = __shfl_sync(FULL_MASK, g, j);
            if (gj == 0.f) continue;                      // flat activation (or bad label): no gradient anywhere
            const int tj = __shfl_sync(FULL_MASK, t, j);
            const int64_t aj = __shfl_sync(FULL_MASK, a, j), bj = __shfl_sync(FULL_MASK, b, j);
            if (tj != cur) {
                if (cur >= 0) flush_label<NV>(acc, gsum, cur, lane, C, C4, dW0, db0);
                cur = tj;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int c4 = lane + 32 * v;
                    if (c4 < C4) w0[v] = ldg4(W0 + (int64_t)tj * C + 4 * c4);
                }
            }
            gsum += gj;
            if (aj != run_a) {
                if (run_a >= 0) flush_rows<NV>(run_hi, dhi + run_a * C, lane, C4);
                run_a = aj;
            }
            const float* xi_row = hi + aj * C;
            const float* xj_row = hj + bj * C;
            float4 xi[NV], xj[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c4 = lane + 32 * v;
                if (c4 < C4) {
                    xi[v] = ldg4(xi_row + 4 * c4);
                    xj[v] = ldg4(xj_row + 4 * c4);
                }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c4 = lane + 32 * v;
                if (c4 < C4) {
                    const float4 dz = make_float4(gj * w0[v].x, gj * w0[v].y, gj * w0[v].z, gj * w0[v].w);
                    run_hi[v].x = fmaf(dz.x, xj[v].x, run_hi[v].x);
                    run_hi[v].y = fmaf(dz.y, xj[v].y, run_hi[v].y);
                    run_hi[v].z = fmaf(dz.z, xj[v].z, run_hi[v].z);
                    run_hi[v].w = fmaf(dz.w, xj[v].w, run_hi[v].w);
                    atomicAdd(reinterpret_cast<float4*>(dhj + bj * C + 4 * c4),
                              make_float4(dz.x * xi[v].x, dz.y * xi[v].y, dz.z * xi[v].z, dz.w * xi[v].w));
                    acc[v].x = fmaf(gj, xi[v].x * xj[v].x, acc[v].x);
                    acc[v].y = fmaf(gj, xi[v].y * xj[v].y, acc[v].y);
                    acc[v].z = fmaf(gj, xi[v].z * xj[v].z, acc[v].z);
                    acc[v].w = fmaf(gj, xi[v].w * xj[v].w, acc[v].w);
                }
            }
        }
    }
    if (cur >= 0) flush_label<NV>(acc, gsum, cur, lane, C, C4, dW0, db0);
    if (run_a >= 0) flush_rows<NV>(run_hi, dhi + run_a * C, lane, C4);
}

// keys for the pair sort: key = label (clamped into [0, Hd] so that bad labels sort last), optionally refined by the
// source row (key = label << sbits | src) so that equal sources become neighbours inside a label; value = pair index
__global__ void label_keys_kernel(const int64_t* __restrict__ target, const int64_t* __restrict__ src, int64_t P, int Hd,
                                  int64_t n_src, int sbits, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int64_t t = target[p];
    uint64_t k = (t >= 0 && t < Hd) ? (uint64_t)t : (uint64_t)Hd;
    if (src) {
        int64_t a = src[p];
        a = a < 0 ? 0 : (a >= n_src ? n_src - 1 : a);
        k = (k << sbits) | (uint64_t)a;
    }
    keys[p] = k;
    vals[p] = (uint32_t)p;
}

}  // namespace

// order: uint32[P] = pair indices stably sorted by label (msha_score_nll_label_order), or NULL for the identity
// (labels already in runs).  dhi / dhj are accumulated into (caller zeroes them); dW0 / db0 are overwritten.
// Requirements: C % 4 == 0, C <= 1024, 16-byte aligned hi / hj / W0 / dhi / dhj / dW0.
MSHA_API int msha_score_mlp_nll_bwd_sparse(const uint32_t* order, const int64_t* target, const float* gout,
                                           const float* out, int64_t ldo, const float* hi_tab, const float* hj_tab,
                                           const int64_t* src, const int64_t* dst, int64_t P, int64_t C, const float* W0,
                                           int64_t Hd, int act, float slope, float* dhi, float* dhj, float* dW0, float* db0,
                                           void* stream) {
    MSHA_REQUIRE(target && gout && out && hi_tab && hj_tab && W0 && dhi && dhj && dW0 && db0, "score_mlp_nll_bwd_sparse: NULL argument");
    MSHA_REQUIRE(P >= 0 && P < ((int64_t)1 << 32) && C >= 4 && (C & 3) == 0 && C <= 1024 && Hd >= 1 && Hd < ((int64_t)1 << 31) &&
                     ldo >= Hd,
                 "score_mlp_nll_bwd_sparse: bad shape (need C %% 4 == 0, C <= 1024, ldo >= Hd)");
    MSHA_REQUIRE(((((uintptr_t)hi_tab) | ((uintptr_t)hj_tab) | ((uintptr_t)W0) | ((uintptr_t)dhi) | ((uintptr_t)dhj) |
                   ((uintptr_t)dW0)) & 15) == 0,
                 "score_mlp_nll_bwd_sparse: tables and gradients must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    MSHA_CUDA(cudaMemsetAsync(db0, 0, (size_t)Hd * sizeof(float), st));
    MSHA_CUDA(cudaMemsetAsync(dW0, 0, (size_t)Hd * C * sizeof(float), st));
    if (P == 0) return 0;
    // chunk: long enough that the per-chunk flush (C atomics onto a handful of dW0 rows) is rare, short enough for
    // >= 8 warps per SM of parallelism
    int64_t chunk = 256;
    while (chunk > 32 && msha_cdiv(P, chunk) < (int64_t)MSHA_NUM_SMS * 8) chunk >>= 1;
    const int64_t warps = msha_cdiv(P, chunk);
    const unsigned grid = (unsigned)msha_cdiv(warps * 32, 256);
    const int c4 = (int)(C >> 2);
#define GO(NVv)                                                                                                            \
    score_nll_sparse_bwd_kernel<NVv><<<grid, 256, 0, st>>>(order, target, gout, out, ldo, hi_tab, hj_tab, src, dst, P, (int)C, \
                                                           (int)Hd, W0, act, slope, (int)chunk, dhi, dhj, dW0, db0)
    if (c4 <= 32) GO(1);
    else if (c4 <= 64) GO(2);
    else if (c4 <= 128) GO(4);
    else GO(8);
#undef GO
    MSHA_LAUNCH_OK();
    return 0;
}

// Stable order of the pairs for the kernel above: by label, and inside a label by source row when `src` is given (then the
// dh_i contributions of a source are summed in registers and scattered once per run, and its h_i row stays in L1).  Fills
// keys / order and sorts them with the library's LSD radix sort (one 8-bit pass per 8 key bits).  keys, keys_tmp:
// uint64[P]; order, order_tmp: uint32[P]; ws from msha_radix_sort_workspace_bytes(P).  The result ends up in `order`.
MSHA_API int msha_radix_sort_u64(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, int64_t n,
                                 int begin_bit, int end_bit, void* ws, size_t ws_bytes, void* stream);   // graph_build.cu
MSHA_API int msha_score_nll_label_order(const int64_t* target, const int64_t* src, int64_t P, int64_t Hd, int64_t n_src,
                                        uint64_t* keys, uint64_t* keys_tmp, uint32_t* order, uint32_t* order_tmp, void* ws,
                                        size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(P >= 0 && P < ((int64_t)1 << 32) && Hd >= 1 && Hd < ((int64_t)1 << 31), "score_nll_label_order: bad shape");
    MSHA_REQUIRE(src == nullptr || (n_src >= 1 && n_src < ((int64_t)1 << 31)), "score_nll_label_order: bad n_src");
    if (P == 0) return 0;
    int lbits = 1;
    while (((int64_t)1 << lbits) <= Hd) ++lbits;          // labels lie in [0, Hd]
    int sbits = 0;
    if (src)
        while (((int64_t)1 << sbits) < n_src) ++sbits;
    label_keys_kernel<<<(unsigned)msha_cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(target, src, P, (int)Hd, n_src, sbits, keys,
                                                                                  order);
    MSHA_LAUNCH_OK();
    const int end_bit = ((lbits + sbits + 7) / 8) * 8;
    return msha_radix_sort_u64(keys, keys_tmp, order, order_tmp, P, 0, end_bit, ws, ws_bytes, stream);
}
