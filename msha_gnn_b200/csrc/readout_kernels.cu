// Read-out kernels of the callers either side of the hot path:
//   * KD_cosine(s, t) = 1 - mean_p cos(s[is[p]], t[it[p]])  fused with the row gathers h[source_index]   (LLP.py:34-35,236)
//   * MSELoss(a, b), mean over all elements                                                               (LLP.py:221,237)
//   * GraphSAGE's `adj[source_index] * x`: batch rows of a CSR adjacency times a dense (B, M) block       (SGAE.py:53)
//   * attention export: per row (CSR) or per column (CSC + perm) maximum of the per-edge attention, the slot of
//     its first occurrence and the number of ties -- what Explainer.py:25-30 extracts from the dense dumps
// All HBM-bound streaming / gather passes: a warp per row, coalesced (128-bit where the row allows) loads,
// deterministic two-stage fp64 sums for the scalar losses.
#include "common.cuh"

constexpr int LOSS_BLOCKS = 592;   // 4 CTAs per SM

__device__ __forceinline__ void block_partial(double acc, double* __restrict__ partial) {
    __shared__ double sm[32];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w];
        partial[blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// KD_cosine.  torch.cosine_similarity divides each vector by max(||.||, eps) (eps = 1e-8) and sums the products.
//   forward : cos[p] kept for the backward; loss = 1 - (1/P) sum_p cos[p]
//   backward: g = -gout/P;  ds[is[p]] += g * (t / (a b) - cos s / (a |s|)),  a = max(|s|, eps), b = max(|t|, eps)
//             (the clamp is applied outside autograd in torch, so the norm's own gradient s/|s| is kept; 0 for s = 0)
// ---------------------------------------------------------------------------------------------
template <bool VEC>
__device__ __forceinline__ void row_dots(const float* __restrict__ sp, const float* __restrict__ tp, int C, int lane,
                                         float& dot, float& ss, float& tt) {
    dot = ss = tt = 0.f;
    if (VEC) {
        for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
            const float4 x = ldg4(sp + 4 * c4), y = ldg4(tp + 4 * c4);
            dot = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, dot))));
            ss = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, ss))));
            tt = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, tt))));
        }
    } else {
        for (int c = lane; c < C; c += 32) {
            const float x = __ldg(sp + c), y = __ldg(tp + c);
            dot = fmaf(x, y, dot);
            ss = fmaf(x, x, ss);
            tt = fmaf(y, y, tt);
        }
    }
    dot = warp_sum(dot);
    ss = warp_sum(ss);
    tt = warp_sum(tt);
}

template <bool VEC>
__global__ void kd_cosine_fwd_kernel(const float* __restrict__ s, const float* __restrict__ t,
                                     const int64_t* __restrict__ idx_s, const int64_t* __restrict__ idx_t, int64_t P,
                                     int C, int64_t n_s, int64_t n_t, float eps, float* __restrict__ cosv,
                                     double* __restrict__ partial, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    for (int64_t p = warp; p < P; p += nwarps) {
        const int64_t a = idx_s ? idx_s[p] : p, b = idx_t ? idx_t[p] : p;
        float c = 0.f;
        if (a < 0 || a >= n_s || b < 0 || b >= n_t) {
            if (lane == 0) atomicOr(status, 1);
        } else {
            float dot, ss, tt;
            row_dots<VEC>(s + a * C, t + b * C, C, lane, dot, ss, tt);
            c = dot / (fmaxf(sqrtf(ss), eps) * fmaxf(sqrtf(tt), eps));
        }
        if (lane == 0) {
            cosv[p] = c;
            acc += (double)c;
        }
    }
    block_partial(acc, partial);
}
__global__ void kd_cosine_fwd_stage2(const double* __restrict__ partial, int n, int64_t P, float* __restrict__ loss) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
    s = warp_sum_d(s);
    if (threadIdx.x == 0) *loss = (float)(1.0 - s / (double)P);
}

template <bool VEC>
__global__ void kd_cosine_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t,
                                     const int64_t* __restrict__ idx_s, const int64_t* __restrict__ idx_t, int64_t P,
                                     int C, int64_t n_s, int64_t n_t, float eps, const float* __restrict__ cosv,
                                     const float* __restrict__ gout, float* __restrict__ ds, float* __restrict__ dt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float g = -(*gout) / (float)P;
    for (int64_t p = warp; p < P; p += nwarps) {
        const int64_t a = idx_s ? idx_s[p] : p, b = idx_t ? idx_t[p] : p;
        if (a < 0 || a >= n_s || b < 0 || b >= n_t) continue;
        const float* sp = s + a * C;
        const float* tp = t + b * C;
        float dot, ss, tt;
        row_dots<VEC>(sp, tp, C, lane, dot, ss, tt);
        const float ns = sqrtf(ss), nt = sqrtf(tt);
        const float na = fmaxf(ns, eps), nb = fmaxf(nt, eps);
        const float c = cosv[p];
        const float cross = g / (na * nb);                              // coefficient of the other vector
        const float self_s = ns > 0.f ? g * c / (na * ns) : 0.f;        // coefficient of the vector itself
        const float self_t = nt > 0.f ? g * c / (nb * nt) : 0.f;
        for (int k = lane; k < C; k += 32) {                            // second touch of the rows: L1/L2 hits
            const float x = __ldg(sp + k), y = __ldg(tp + k);
            if (ds) atomicAdd(ds + a * C + k, cross * y - self_s * x);
            if (dt) atomicAdd(dt + b * C + k, cross * x - self_t * y);
        }
    }
}

MSHA_API size_t msha_loss_workspace_bytes(void) { return LOSS_BLOCKS * sizeof(double); }

static inline bool rows_vec4(const float* a, const float* b, int64_t C) {
    return (C & 3) == 0 && ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0;
}

// cos: float[P]; loss: float[1]; status: int32[1] (bit 0: an index was out of range -- that pair contributes cos = 0)
MSHA_API int msha_kd_cosine_fwd(const float* s, const float* t, const int64_t* idx_s, const int64_t* idx_t, int64_t P,
                                int64_t C, int64_t n_s, int64_t n_t, float eps, float* cosv, float* loss,
                                int32_t* status, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(P >= 1 && C >= 1 && C < ((int64_t)1 << 31) && n_s >= 1 && n_t >= 1, "kd_cosine: bad shape");
    MSHA_REQUIRE((idx_s || n_s >= P) && (idx_t || n_t >= P), "kd_cosine: fewer rows than pairs without an index");
    MSHA_REQUIRE(ws_bytes >= msha_loss_workspace_bytes(), "kd_cosine: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    MSHA_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    const int64_t want = msha_cdiv(P * 32, 256);
    const int nb = (int)(want < LOSS_BLOCKS ? want : LOSS_BLOCKS);
    if (rows_vec4(s, t, C))
        kd_cosine_fwd_kernel<true><<<nb, 256, 0, st>>>(s, t, idx_s, idx_t, P, (int)C, n_s, n_t, eps, cosv, (double*)ws, status);
    else
        kd_cosine_fwd_kernel<false><<<nb, 256, 0, st>>>(s, t, idx_s, idx_t, P, (int)C, n_s, n_t, eps, cosv, (double*)ws, status);
    MSHA_LAUNCH_OK();
    kd_cosine_fwd_stage2<<<1, 32, 0, st>>>((const double*)ws, nb, P, loss);
    MSHA_LAUNCH_OK();
    return 0;
}
// gout: float[1] device.  ds / dt (either may be NULL: LLP.py:35 detaches the teacher) are accumulated into.
MSHA_API int msha_kd_cosine_bwd(const float* s, const float* t, const int64_t* idx_s, const int64_t* idx_t, int64_t P,
                                int64_t C, int64_t n_s, int64_t n_t, float eps, const float* cosv, const float* gout,
                                float* ds, float* dt, void* stream) {
    MSHA_REQUIRE(P >= 1 && C >= 1 && C < ((int64_t)1 << 31) && n_s >= 1 && n_t >= 1, "kd_cosine_bwd: bad shape");
    if (!ds && !dt) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t want = msha_cdiv(P * 32, 256);
    const int nb = (int)(want < MSHA_NUM_SMS * 8 ? want : MSHA_NUM_SMS * 8);
    if (rows_vec4(s, t, C))
        kd_cosine_bwd_kernel<true><<<nb, 256, 0, st>>>(s, t, idx_s, idx_t, P, (int)C, n_s, n_t, eps, cosv, gout, ds, dt);
    else
        kd_cosine_bwd_kernel<false><<<nb, 256, 0, st>>>(s, t, idx_s, idx_t, P, (int)C, n_s, n_t, eps, cosv, gout, ds, dt);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// MSELoss (mean): loss = (1/n) sum (a - b)^2;  da = 2 (a - b) g / n,  db = -da
// ---------------------------------------------------------------------------------------------
__global__ void mse_fwd_stage1(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                               double* __restrict__ partial) {
    double acc = 0.0;
    float run = 0.f;                   // short fp32 runs folded into the fp64 accumulator
    int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        run = fmaf(d, d, run);
        if (++cnt == 16) { acc += (double)run; run = 0.f; cnt = 0; }
    }
    block_partial(acc + (double)run, partial);
}
__global__ void mse_fwd_stage2(const double* __restrict__ partial, int nparts, int64_t n, float* __restrict__ loss) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) s += partial[i];
    s = warp_sum_d(s);
    if (threadIdx.x == 0) *loss = (float)(s / (double)n);
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                               const float* __restrict__ gout, float* __restrict__ da, float* __restrict__ db) {
    const float g = 2.f * (*gout) / (float)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = g * (a[i] - b[i]);
        if (da) da[i] = v;
        if (db) db[i] = -v;
    }
}
MSHA_API int msha_mse_loss_fwd(const float* a, const float* b, int64_t n, float* loss, void* ws, size_t ws_bytes,
                               void* stream) {
    MSHA_REQUIRE(n >= 1, "mse_loss: empty input");
    MSHA_REQUIRE(ws_bytes >= msha_loss_workspace_bytes(), "mse_loss: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t want = msha_cdiv(n, 256);
    const int nb = (int)(want < LOSS_BLOCKS ? want : LOSS_BLOCKS);
    mse_fwd_stage1<<<nb, 256, 0, st>>>(a, b, n, (double*)ws);
    MSHA_LAUNCH_OK();
    mse_fwd_stage2<<<1, 32, 0, st>>>((const double*)ws, nb, n, loss);
    MSHA_LAUNCH_OK();
    return 0;
}
// da / db are overwritten (either may be NULL: LLP.py:236 detaches the teacher's scores)
MSHA_API int msha_mse_loss_bwd(const float* a, const float* b, int64_t n, const float* gout, float* da, float* db,
                               void* stream) {
    MSHA_REQUIRE(n >= 1, "mse_loss_bwd: empty input");
    if (!da && !db) return 0;
    const int64_t want = msha_cdiv(n, 256);
    const int nb = (int)(want < MSHA_NUM_SMS * 8 ? want : MSHA_NUM_SMS * 8);
    mse_bwd_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(a, b, n, gout, da, db);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// GraphSAGE: out[b, :] = adj[src[b], :] * x[b, :]  with adj given as CSR (+ values)           (SGAE.py:53)
// The map is diagonal, so the same call with x = d out is its backward.
// ---------------------------------------------------------------------------------------------
__global__ void csr_rows_mul_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                    const float* __restrict__ val, const int64_t* __restrict__ src, int64_t B,
                                    int64_t n_rows, int M, const float* __restrict__ x, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < B; b += nwarps) {
        float* o = out + b * M;
        for (int j = lane; j < M; j += 32) o[j] = 0.f;
        __syncwarp();                                   // the zeros are ordered before the patches below
        const int64_t r = src ? src[b] : b;
        if (r < 0 || r >= n_rows) continue;
        const int e0 = rowptr[r], e1 = rowptr[r + 1];
        for (int e = e0 + lane; e < e1; e += 32) {
            const int j = col[e];
            o[j] = (val ? val[e] : 1.f) * x[b * M + j];
        }
        __syncwarp();
    }
}
MSHA_API int msha_csr_rows_mul(const int32_t* rowptr, const int32_t* col, const float* val, const int64_t* src, int64_t B,
                               int64_t n_rows, int64_t M, const float* x, float* out, void* stream) {
    MSHA_REQUIRE(B >= 0 && n_rows >= 0 && M >= 1 && M < ((int64_t)1 << 31), "csr_rows_mul: bad shape");
    if (B == 0) return 0;
    const int64_t want = msha_cdiv(B * 32, 256);
    const int nb = (int)(want < MSHA_NUM_SMS * 8 ? want : MSHA_NUM_SMS * 8);
    csr_rows_mul_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, src, B, n_rows, (int)M, x, out);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Attention export (Explainer.py:25-30: argwhere(row == max(row)) over the dense Coeff dumps of train.py:284-321).
// Item i owns slots [ptr[i], ptr[i+1]); slot e reads w[(perm ? perm[e] : e) * H + head] (head < 0: mean over heads).
// Every positive attention sits on an edge, so the sparse maximum equals the dense row maximum.
//   vmax[i]  = maximum (0 for an empty item)      first[i] = smallest slot attaining it (-1 for an empty item)
//   ties[i]  = number of slots attaining it
// ---------------------------------------------------------------------------------------------
__global__ void segment_argmax_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm,
                                      const float* __restrict__ w, int H, int head, int64_t n_items,
                                      float* __restrict__ vmax, int32_t* __restrict__ first, int32_t* __restrict__ ties) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv_h = 1.f / (float)H;
    for (int64_t i = warp; i < n_items; i += nwarps) {
        const int e0 = ptr[i], e1 = ptr[i + 1];
        float m = -INFINITY;
        int f = 0x7fffffff, cnt = 0;
        for (int e = e0 + lane; e < e1; e += 32) {
            const int64_t slot = perm ? perm[e] : e;
            float v;
            if (head >= 0) {
                v = w[slot * H + head];
            } else {
                v = 0.f;
                for (int h = 0; h < H; ++h) v += w[slot * H + h];
                v *= inv_h;
            }
            if (v > m) { m = v; f = e; cnt = 1; }
            else if (v == m) { ++cnt; }                 // slots ascend within a lane: f stays the smallest
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(FULL_MASK, m, o);
            const int f2 = __shfl_xor_sync(FULL_MASK, f, o);
            const int c2 = __shfl_xor_sync(FULL_MASK, cnt, o);
            if (m2 > m) { m = m2; f = f2; cnt = c2; }
            else if (m2 == m) { f = f2 < f ? f2 : f; cnt += c2; }
        }
        if (lane == 0) {
            const bool empty = e1 <= e0;
            vmax[i] = empty ? 0.f : m;
            first[i] = empty ? -1 : f;
            ties[i] = empty ? 0 : cnt;
        }
    }
}
MSHA_API int msha_segment_argmax(const int32_t* ptr, const int32_t* perm, const float* w, int H, int head,
                                 int64_t n_items, float* vmax, int32_t* first, int32_t* ties, void* stream) {
    MSHA_REQUIRE(n_items >= 0 && H >= 1 && head < H, "segment_argmax: bad arguments");
    if (n_items == 0) return 0;
    const int64_t want = msha_cdiv(n_items * 32, 256);
    const int nb = (int)(want < MSHA_NUM_SMS * 8 ? want : MSHA_NUM_SMS * 8);
    segment_argmax_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(ptr, perm, w, H, head, n_items, vmax, first, ties);
    MSHA_LAUNCH_OK();
    return 0;
}
