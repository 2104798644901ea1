// K-5: intra-scale (city / province) block of OursLayer (Ours.py:71-90,99) without the (B,N) / (N,N) dense
// tensors.  The logit of a batch row is a row constant, so att3[b, n] = coef3[b] for every member n of the
// batch row's group: the (B,N)@(B,d') products collapse into segment broadcasts / segment sums.
//
//   rowsum_exp        T[b,h] = sum_{j<M} exp(alpha_drop[src_b, j, h])        third term of SUM_county, Ours.py:86
//                     (non-edges contribute e^0 = 1; duplicates in src handled row by row)
//   group_scatter_add out[n] += coef[b,h] * drop * feat[b,h,:]  for n in list(src_b)   att3.t() @ h2_  Ours.py:99
//   group_gather_sum  G[b]   = sum_{n in list(src_b)} drop * dout[n]                    its backward
//
// list(r) = col[rowptr[k] .. rowptr[k+1]) with k = row_map ? row_map[r] : r  (generic CSR of a dense (N,N)
// adjacency, or group-member lists when the adjacency is a group-equality block structure).
#include "common.cuh"

MSHA_DEFINE_DROP_EPOCH_HOOK(intra)

struct DropI { uint32_t thr; float inv_keep; uint64_t seed; uint32_t stream; };
static DropI make_drop_i(float p, uint64_t seed, uint32_t stream) {
    DropI d; d.thr = 0; d.inv_keep = 1.f; d.seed = seed; d.stream = stream;
    if (p > 0.f) {
        double t = (double)p * 4294967296.0;
        d.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        if (d.thr == 0) d.thr = 1;
        d.inv_keep = 1.f / (1.f - p);
    }
    return d;
}

// one warp per batch entry
__global__ void rowsum_exp_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ alpha, int H,
                                  const int64_t* __restrict__ src, int64_t B, int n_cols, float* __restrict__ T,
                                  DropI drop) {
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int64_t r = src[b];
    const int beg = rowptr[r], end = rowptr[r + 1];
    for (int h = 0; h < H; ++h) {
        float s = 0.f;
        for (int e = beg + lane; e < end; e += 32) {
            float a = alpha[(int64_t)e * H + h];
            if (drop.thr) a *= dropout_scale(drop.seed, drop.stream, (uint64_t)e * H + h, drop.thr, drop.inv_keep);
            s += expf(a);
        }
        s = warp_sum(s);
        if (lane == 0) T[b * H + h] = s + (float)(n_cols - (end - beg));
    }
}
// dalpha[e,h] += dT[b,h] * exp(alpha*k) * k      (dalpha zero-initialised by the caller)
__global__ void rowsum_exp_bwd_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ alpha, int H,
                                      const int64_t* __restrict__ src, int64_t B, const float* __restrict__ dT,
                                      float* __restrict__ dalpha, DropI drop) {
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int64_t r = src[b];
    const int beg = rowptr[r], end = rowptr[r + 1];
    for (int h = 0; h < H; ++h) {
        const float g = dT[b * H + h];
        for (int e = beg + lane; e < end; e += 32) {
            float k = 1.f;
            if (drop.thr) k = dropout_scale(drop.seed, drop.stream, (uint64_t)e * H + h, drop.thr, drop.inv_keep);
            const float a = alpha[(int64_t)e * H + h] * k;
            atomicAdd(&dalpha[(int64_t)e * H + h], g * expf(a) * k);
        }
    }
}

MSHA_API int msha_rowsum_exp(const int32_t* rowptr, const float* alpha, int H, const int64_t* src, int64_t B,
                             int64_t n_cols, float* T, float drop_p, uint64_t drop_seed, void* stream) {
    MSHA_REQUIRE(H >= 1 && B >= 0 && n_cols >= 0, "rowsum_exp: bad shape");
    if (B == 0) return 0;
    rowsum_exp_kernel<<<(unsigned)msha_cdiv(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        rowptr, alpha, H, src, B, (int)n_cols, T, make_drop_i(drop_p, drop_seed, 2u));
    MSHA_LAUNCH_OK();
    return 0;
}
MSHA_API int msha_rowsum_exp_bwd(const int32_t* rowptr, const float* alpha, int H, const int64_t* src, int64_t B,
                                 const float* dT, float* dalpha, float drop_p, uint64_t drop_seed, void* stream) {
    MSHA_REQUIRE(H >= 1 && B >= 0, "rowsum_exp_bwd: bad shape");
    if (B == 0) return 0;
    rowsum_exp_bwd_kernel<<<(unsigned)msha_cdiv(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        rowptr, alpha, H, src, B, dT, dalpha, make_drop_i(drop_p, drop_seed, 2u));
    MSHA_LAUNCH_OK();
    return 0;
}

// grid = (B, chunks); each warp of a block handles a strided share of the member list.
// dropout element index: ((h*B + b) * n_nodes + n)  -- the dense (B,N) attention3 of head h (Ours.py:88).
__global__ void group_scatter_add_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                         const int32_t* __restrict__ row_map, const int64_t* __restrict__ src, int64_t B,
                                         int64_t n_nodes, const float* __restrict__ coef, const float* __restrict__ feat,
                                         int H, int D, float* __restrict__ out, DropI drop) {
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int64_t r = src[b];
    const int64_t k = row_map ? row_map[r] : r;
    const int beg = rowptr[k], end = rowptr[k + 1];
    const int C = H * D;
    for (int m = beg + blockIdx.y * nwarp + warp; m < end; m += gridDim.y * nwarp) {
        int n = col[m];
        n = n < 0 ? ~n : n;
        for (int c = lane; c < C; c += 32) {
            const int h = c / D;
            float w = coef[(int64_t)b * H + h];
            if (drop.thr)
                w *= dropout_scale(drop.seed, drop.stream, ((uint64_t)h * B + b) * n_nodes + n, drop.thr, drop.inv_keep);
            atomicAdd(&out[(int64_t)n * C + c], w * feat[(int64_t)b * C + c]);
        }
    }
}
// G[b,c] = sum_n drop * dout[n,c]   (atomics across the chunks of one batch entry; G zero-initialised)
__global__ void group_gather_sum_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                        const int32_t* __restrict__ row_map, const int64_t* __restrict__ src, int64_t B,
                                        int64_t n_nodes, const float* __restrict__ dout, int H, int D,
                                        float* __restrict__ G, DropI drop) {
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int64_t r = src[b];
    const int64_t k = row_map ? row_map[r] : r;
    const int beg = rowptr[k], end = rowptr[k + 1];
    const int C = H * D;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const int h = c < C ? c / D : 0;
        float acc = 0.f;
        for (int m = beg + blockIdx.y * nwarp + warp; m < end; m += gridDim.y * nwarp) {
            int n = col[m];
            n = n < 0 ? ~n : n;
            if (c < C) {
                float w = 1.f;
                if (drop.thr)
                    w = dropout_scale(drop.seed, drop.stream, ((uint64_t)h * B + b) * n_nodes + n, drop.thr, drop.inv_keep);
                acc = fmaf(w, dout[(int64_t)n * C + c], acc);
            }
        }
        if (c < C && acc != 0.f) atomicAdd(&G[(int64_t)b * C + c], acc);
    }
}

// out [n_nodes, H*D] must hold the running value (accumulated with atomics).
MSHA_API int msha_group_scatter_add(const int32_t* rowptr, const int32_t* col, const int32_t* row_map,
                                    const int64_t* src, int64_t B, int64_t n_nodes, const float* coef, const float* feat,
                                    int H, int D, float* out, float drop_p, uint64_t drop_seed, uint32_t drop_stream,
                                    void* stream) {
    MSHA_REQUIRE(H >= 1 && D >= 1 && B >= 0 && B < ((int64_t)1 << 31), "group_scatter_add: bad shape");
    if (B == 0) return 0;
    dim3 grid((unsigned)B, 16);
    group_scatter_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, col, row_map, src, B, n_nodes, coef, feat, H, D,
                                                                   out, make_drop_i(drop_p, drop_seed, drop_stream));
    MSHA_LAUNCH_OK();
    return 0;
}
// G [B, H*D] must be zero-initialised.
MSHA_API int msha_group_gather_sum(const int32_t* rowptr, const int32_t* col, const int32_t* row_map, const int64_t* src,
                                   int64_t B, int64_t n_nodes, const float* dout, int H, int D, float* G, float drop_p,
                                   uint64_t drop_seed, uint32_t drop_stream, void* stream) {
    MSHA_REQUIRE(H >= 1 && D >= 1 && B >= 0 && B < ((int64_t)1 << 31), "group_gather_sum: bad shape");
    if (B == 0) return 0;
    dim3 grid((unsigned)B, 16);
    group_gather_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, col, row_map, src, B, n_nodes, dout, H, D, G,
                                                                  make_drop_i(drop_p, drop_seed, drop_stream));
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Group-sum form of the same block (no dropout on the intra attention): every member of a group receives the SAME row
//     IntraNC[n] = G3[city(n)] + G4[province(n)],   G3[g] = sum_{b : city(src_b) = g} coef3[b] * h2[src_b]       Ours.py:99
// (attention3[b, n] = coef3[b] for every member n of the batch row's city, Ours.py:71-75,87).  O(B + N) work instead of
// O(B x |group|), and the only thing ranks of a partitioned graph have to exchange is the (n_groups, C) table.
//   group_rows_sum   G[gid[i]] += x[i]            (forward of the table, backward of the broadcast)
//   group_rows_add   out[i] (+)= G[gid[i]]        (forward of the broadcast, backward of the table)
// ---------------------------------------------------------------------------------------------
constexpr int GRS_THREADS = 256;

// few groups (table fits in shared memory): per-CTA partial table in shared memory, one atomic flush per CTA
__global__ void __launch_bounds__(GRS_THREADS)
group_rows_sum_smem_kernel(const float* __restrict__ x, const int64_t* __restrict__ gid, int64_t n, int C, int n_groups,
                           float* __restrict__ G) {
    extern __shared__ float sm[];
    const int tot = n_groups * C;
    for (int i = threadIdx.x; i < tot; i += GRS_THREADS) sm[i] = 0.f;
    __syncthreads();
    const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t g = gid[r];
        if (g < 0 || g >= n_groups) continue;
        for (int c = threadIdx.x; c < C; c += GRS_THREADS) atomicAdd(&sm[g * C + c], x[r * C + c]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tot; i += GRS_THREADS) {
        const float v = sm[i];
        if (v != 0.f) atomicAdd(&G[i], v);
    }
}
// many groups: one warp per row, 128-bit vector atomics straight into the (L2-resident) table
__global__ void __launch_bounds__(GRS_THREADS)
group_rows_sum_kernel(const float* __restrict__ x, const int64_t* __restrict__ gid, int64_t n, int C4, int n_groups,
                      float* __restrict__ G) {
    const int64_t row = ((int64_t)blockIdx.x * GRS_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int64_t g = gid[row];
    if (g < 0 || g >= n_groups) return;
    const float4* __restrict__ xr = reinterpret_cast<const float4*>(x) + row * C4;
    float4* Gr = reinterpret_cast<float4*>(G) + g * C4;
    for (int c = lane; c < C4; c += 32) atomicAdd(Gr + c, __ldg(xr + c));
}
__global__ void __launch_bounds__(GRS_THREADS)
group_rows_add_kernel(float* __restrict__ out, const float* __restrict__ G3, const int64_t* __restrict__ gid3,
                      const float* __restrict__ G4, const int64_t* __restrict__ gid4, int64_t n, int C4, int accumulate) {
    const int64_t idx = (int64_t)blockIdx.x * GRS_THREADS + threadIdx.x;
    if (idx >= n * C4) return;
    const int64_t row = idx / C4;
    const int c = (int)(idx - row * C4);
    float4 v = __ldg(reinterpret_cast<const float4*>(G3) + gid3[row] * C4 + c);
    if (G4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(G4) + gid4[row] * C4 + c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    float4* o = reinterpret_cast<float4*>(out) + idx;
    if (accumulate) { const float4 p = *o; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
    *o = v;
}

// G: float[n_groups, C], zeroed here.  gid: int64[n] group of every row (rows with an id outside [0, n_groups) are skipped).
MSHA_API int msha_group_rows_sum(const float* x, const int64_t* gid, int64_t n, int64_t C, int64_t n_groups, float* G,
                                 void* stream) {
    MSHA_REQUIRE(n >= 0 && C >= 4 && C % 4 == 0 && n_groups >= 1 && G != nullptr, "group_rows_sum: bad shape (C % 4 == 0)");
    cudaStream_t st = (cudaStream_t)stream;
    MSHA_CUDA(cudaMemsetAsync(G, 0, (size_t)n_groups * C * sizeof(float), st));
    if (n == 0) return 0;
    const size_t tab = (size_t)n_groups * C * sizeof(float);
    if (tab <= 40 * 1024) {
        int64_t ctas = msha_cdiv(n, 64);
        if (ctas > 4 * MSHA_NUM_SMS) ctas = 4 * MSHA_NUM_SMS;
        group_rows_sum_smem_kernel<<<(unsigned)ctas, GRS_THREADS, tab, st>>>(x, gid, n, (int)C, (int)n_groups, G);
    } else {
        group_rows_sum_kernel<<<(unsigned)msha_cdiv(n * 32, GRS_THREADS), GRS_THREADS, 0, st>>>(x, gid, n, (int)(C / 4),
                                                                                               (int)n_groups, G);
    }
    MSHA_LAUNCH_OK();
    return 0;
}

// out[i] (+)= G3[gid3[i]] (+ G4[gid4[i]]); ids must lie inside their tables.
MSHA_API int msha_group_rows_add(float* out, const float* G3, const int64_t* gid3, const float* G4, const int64_t* gid4,
                                 int64_t n, int64_t C, int accumulate, void* stream) {
    MSHA_REQUIRE(n >= 0 && C >= 4 && C % 4 == 0 && out && G3 && gid3, "group_rows_add: bad arguments (C % 4 == 0)");
    MSHA_REQUIRE((G4 == nullptr) == (gid4 == nullptr), "group_rows_add: G4 and gid4 go together");
    if (n == 0) return 0;
    group_rows_add_kernel<<<(unsigned)msha_cdiv(n * (C / 4), GRS_THREADS), GRS_THREADS, 0, (cudaStream_t)stream>>>(
        out, G3, gid3, G4, gid4, n, (int)(C / 4), accumulate);
    MSHA_LAUNCH_OK();
    return 0;
}
