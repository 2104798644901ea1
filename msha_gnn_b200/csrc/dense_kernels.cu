// Dense / node-level kernels of the hot path (fp32):
//   * SIMT fp32 GEMM with bias + activation epilogue (validation path and odd shapes; the tcgen05 3xTF32
//     GEMM in gemm_tcgen05.cu is the production path for the aligned shapes)
//   * a-1 GraphAttentionLayer epilogue:  out = elu(att * h), att = mask/deg           (GAT.py:29-35)
//   * activations / their backward from the saved output
//   * BatchNorm1d over the node axis fused with LeakyReLU, training + eval            (Ours.py:100-101)
//   * row log_softmax (M small) fused with a leading ELU                              (Ours.py:166-167)
//   * pair gather * Hadamard / scatter-add for the link scorer                        (LLP.py:105)
#include "common.cuh"

MSHA_DEFINE_DROP_EPOCH_HOOK(dense)

enum { ACT_NONE = 0, ACT_ELU = 1, ACT_RELU = 2, ACT_SIGMOID_RELU = 3, ACT_LRELU = 4, ACT_SIGMOID = 5 };

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
    switch (act) {
        case ACT_ELU: return x > 0.f ? x : expm1f(x);
        case ACT_RELU: return fmaxf(x, 0.f);
        case ACT_SIGMOID_RELU: return 1.f / (1.f + expf(-fmaxf(x, 0.f)));
        case ACT_LRELU: return x > 0.f ? x : x * slope;
        case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        default: return x;
    }
}
// derivative of the activation expressed through its *output* y
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
    switch (act) {
        case ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
        case ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case ACT_SIGMOID_RELU: return y > 0.5f ? y * (1.f - y) : 0.f;   // sigmoid(relu(x)) == 0.5 <=> x <= 0
        case ACT_LRELU: return y > 0.f ? 1.f : slope;
        case ACT_SIGMOID: return y * (1.f - y);
        default: return 1.f;
    }
}

// ---------------------------------------------------------------------------------------------
// SIMT GEMM   C[M,N] = act( opA(A)[M,K] * opB(B)[K,N] + bias[N] ) (+ beta*C before act if beta != 0)
// row-major; transA: A is stored [K,M]; transB: B is stored [N,K].
// ---------------------------------------------------------------------------------------------
constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_THREADS = 256;

__global__ void __launch_bounds__(GM_THREADS)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Cmat, int M, int N,
                int K, int64_t lda, int64_t ldb, int64_t ldc, int transA, int transB, const float* __restrict__ bias,
                float beta, int act, float slope) {
    __shared__ float As[GM_BK][GM_BM + 4];
    __shared__ float Bs[GM_BK][GM_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GM_BK) {
        // A tile
        for (int idx = tid; idx < GM_BM * GM_BK; idx += GM_THREADS) {
            int m, k;
            if (transA) { k = idx / GM_BM; m = idx % GM_BM; } else { m = idx / GM_BK; k = idx % GM_BK; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < K) v = transA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
            As[k][m] = v;
        }
        for (int idx = tid; idx < GM_BN * GM_BK; idx += GM_THREADS) {
            int n, k;
            if (transB) { n = idx / GM_BK; k = idx % GM_BK; } else { k = idx / GM_BN; n = idx % GM_BN; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < K) v = transB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GM_BK; ++k) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; a[4 + i] = As[k][64 + ty * 4 + i]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) { b[j] = Bs[k][tx * 4 + j]; b[4 + j] = Bs[k][64 + tx * 4 + j]; }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[gn];
            float* cp = Cmat + (int64_t)gm * ldc + gn;
            if (beta != 0.f) v = fmaf(beta, *cp, v);
            *cp = apply_act(v, act, slope);
        }
    }
}

MSHA_API int msha_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                           int64_t ldb, int64_t ldc, int transA, int transB, const float* bias, float beta, int act,
                           float slope, void* stream) {
    MSHA_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative size");
    MSHA_REQUIRE(M < ((int64_t)1 << 31) && N < ((int64_t)1 << 31) && K < ((int64_t)1 << 31), "gemm: size overflow");
    if (M == 0 || N == 0) return 0;
    dim3 grid((unsigned)msha_cdiv(N, GM_BN), (unsigned)msha_cdiv(M, GM_BM));
    MSHA_REQUIRE(grid.y <= 65535u * 1024u, "gemm: M too large");
    if (grid.y > 65535) {   // fold very tall problems through repeated launches
        const int64_t rows_per = (int64_t)65535 * GM_BM;
        for (int64_t r = 0; r < M; r += rows_per) {
            const int64_t mm = M - r < rows_per ? M - r : rows_per;
            const float* Ar = transA ? A + r : A + r * lda;
            int rc = msha_gemm_f32(Ar, B, C + r * ldc, mm, N, K, lda, ldb, ldc, transA, transB, bias, beta, act, slope, stream);
            if (rc) return rc;
        }
        return 0;
    }
    gemm_f32_kernel<<<grid, GM_THREADS, 0, (cudaStream_t)stream>>>(A, B, C, (int)M, (int)N, (int)K, lda, ldb, ldc, transA,
                                                                  transB, bias, beta, act, slope);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// a-1 GraphAttentionLayer epilogue (GAT.py:29-35).  Attention is uniform over the row's neighbours
// (1/deg; rows without neighbours carry M masked edges in the attention CSR -> 1/M everywhere).
// Dropout acts on the dense (N,M) attention matrix: element index i*M + j  (GAT.py:32).
// ---------------------------------------------------------------------------------------------
struct DropArgsD { uint32_t thr; float inv_keep; uint64_t seed; uint32_t stream; };
static DropArgsD make_drop_d(float p, uint64_t seed, uint32_t stream) {
    DropArgsD d; d.thr = 0; d.inv_keep = 1.f; d.seed = seed; d.stream = stream;
    if (p > 0.f) {
        double t = (double)p * 4294967296.0;
        d.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        if (d.thr == 0) d.thr = 1;
        d.inv_keep = 1.f / (1.f - p);
    }
    return d;
}

// h / out are [n_rows, H, M] (H heads share the mask; head batching of GAT.py:55).
// mode 0: out = elu(att*h) ; mode 1 (backward): out = dout * elu'(y) * att   (h := dout, y := saved output)
__global__ void gal_rows_kernel(const float* __restrict__ h, const float* __restrict__ y,
                                const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                                int H, int M, float* __restrict__ out, int mode, DropArgsD drop) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const int HM = H * M;
    float* o = out + row * HM;
    for (int j = lane; j < HM; j += 32) o[j] = 0.f;
    __syncwarp();
    const int beg = rowptr[row], end = rowptr[row + 1];
    if (end == beg) return;
    const int deg = end - beg;
    const float w = 1.f / (float)deg;
    // lanes tile the (head, edge) pairs of the row: a 2015-shaped row has ~2.3 edges, so an edge-per-lane loop would keep
    // 2-3 lanes busy with H serial Philox draws and dependent loads each
    for (int t = lane; t < deg * H; t += 32) {
        const int hh = t / deg;
        const int c = col[beg + (t - hh * deg)];
        const int j = c < 0 ? ~c : c;
        float a = w;
        if (drop.thr)
            a *= dropout_scale(drop.seed, drop.stream, ((uint64_t)hh * n_rows + row) * M + j, drop.thr, drop.inv_keep);
        const int64_t idx = row * HM + hh * M + j;
        const float x = h[idx];
        if (mode == 0) {
            const float v = a * x;
            o[hh * M + j] = v > 0.f ? v : expm1f(v);
        } else {
            const float yy = y[idx];
            o[hh * M + j] = x * (yy > 0.f ? 1.f : yy + 1.f) * a;
        }
    }
}

MSHA_API int msha_gal_fwd(const float* h, const int32_t* rowptr, const int32_t* col, int64_t n_rows, int H, int64_t M,
                          float* out, float drop_p, uint64_t drop_seed, void* stream) {
    MSHA_REQUIRE(n_rows >= 0 && H >= 1 && M >= 1 && H * M < ((int64_t)1 << 31), "gal_fwd: bad shape");
    if (n_rows == 0) return 0;
    gal_rows_kernel<<<(unsigned)msha_cdiv(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        h, nullptr, rowptr, col, n_rows, H, (int)M, out, 0, make_drop_d(drop_p, drop_seed, 3u));
    MSHA_LAUNCH_OK();
    return 0;
}

MSHA_API int msha_gal_bwd(const float* dout, const float* y, const int32_t* rowptr, const int32_t* col, int64_t n_rows,
                          int H, int64_t M, float* dh, float drop_p, uint64_t drop_seed, void* stream) {
    MSHA_REQUIRE(n_rows >= 0 && H >= 1 && M >= 1 && H * M < ((int64_t)1 << 31), "gal_bwd: bad shape");
    if (n_rows == 0) return 0;
    gal_rows_kernel<<<(unsigned)msha_cdiv(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        dout, y, rowptr, col, n_rows, H, (int)M, dh, 1, make_drop_d(drop_p, drop_seed, 3u));
    MSHA_LAUNCH_OK();
    return 0;
}

// element-wise dropout y = x * keep/(1-p) (feature dropout, Ours.py:161-162,165); its own backward.
__global__ void dropout_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, DropArgsD drop) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = x[i] * dropout_scale(drop.seed, drop.stream, (uint64_t)i, drop.thr, drop.inv_keep);
}
// same masks, one Philox block (4 words) per thread and 128-bit accesses: element 4b + k <- word k of block b
__global__ void dropout_apply_kernel4(const float* x, float* y, int64_t n, DropArgsD drop) {
    const uint64_t seed = drop_seed_eff(drop.seed);
    const int64_t nb = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += stride) {
        const Philox4 r = philox4x32_10((uint32_t)b, (uint32_t)((uint64_t)b >> 32), drop.stream, 0u, (uint32_t)seed,
                                        (uint32_t)(seed >> 32));
        float4 v = reinterpret_cast<const float4*>(x)[b];
        v.x *= r.x >= drop.thr ? drop.inv_keep : 0.f;
        v.y *= r.y >= drop.thr ? drop.inv_keep : 0.f;
        v.z *= r.z >= drop.thr ? drop.inv_keep : 0.f;
        v.w *= r.w >= drop.thr ? drop.inv_keep : 0.f;
        reinterpret_cast<float4*>(y)[b] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const int64_t i = (nb << 2) + threadIdx.x;
        y[i] = x[i] * dropout_scale(drop.seed, drop.stream, (uint64_t)i, drop.thr, drop.inv_keep);
    }
}
MSHA_API int msha_dropout_apply(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t stream_id,
                                void* stream) {
    MSHA_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0,1)");
    if (n <= 0) return 0;
    if (p == 0.f) {
        if (x != y) MSHA_CUDA(cudaMemcpyAsync(y, x, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return 0;
    }
    const bool vec = ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 && n >= 4;
    int64_t g = msha_cdiv(vec ? (n >> 2) : n, 256);
    if (g > (int64_t)MSHA_NUM_SMS * 16) g = (int64_t)MSHA_NUM_SMS * 16;
    if (vec)
        dropout_apply_kernel4<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, y, n, make_drop_d(p, seed, stream_id));
    else
        dropout_apply_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, y, n, make_drop_d(p, seed, stream_id));
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// element-wise activation forward / backward-from-output
// ---------------------------------------------------------------------------------------------
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = apply_act(x[i], act, slope);
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                               int64_t n, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dx[i] = dy[i] * act_grad_from_out(y[i], act, slope);
}
static unsigned ew_grid(int64_t n) {
    int64_t g = msha_cdiv(n, 256);
    const int64_t cap = (int64_t)MSHA_NUM_SMS * 16;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}
MSHA_API int msha_act_fwd(const float* x, float* y, int64_t n, int act, float slope, void* stream) {
    if (n <= 0) return 0;
    act_fwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, act, slope);
    MSHA_LAUNCH_OK();
    return 0;
}
MSHA_API int msha_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, float slope, void* stream) {
    if (n <= 0) return 0;
    act_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, act, slope);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d over rows (node axis) + LeakyReLU                      Ours.py:50-52,100-101
// ---------------------------------------------------------------------------------------------
constexpr int BN_BLOCKS = 296;
// partial[b][0][c] = sum x, partial[b][1][c] = sum x*(z ? z : x)
__global__ void bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ z, int64_t n, int C,
                                  double* __restrict__ partial) {
    const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * rows_per;
    const int64_t r1 = r0 + rows_per < n ? r0 + rows_per : n;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int64_t r = r0; r < r1; ++r) {
            const float v = x[r * C + c];
            const float w = z ? z[r * C + c] : v;
            s1 += (double)v;
            s2 += (double)v * (double)w;
        }
        partial[((int64_t)blockIdx.x * 2 + 0) * C + c] = s1;
        partial[((int64_t)blockIdx.x * 2 + 1) * C + c] = s2;
    }
}
// Second stage of the two column sums: 32 columns per CTA of 8 warps, warp w adds every 8th partial in a fixed order and
// warp 0 combines the 8 (one thread per column walking all ~300 partials was a 30 us chain of dependent fp64 adds).
// Returns true in the threads of warp 0 that own a column; s1 / s2 are valid there.
__device__ __forceinline__ bool bn_column_sums(const double* __restrict__ partial, int nblocks, int C, int& c, double& s1,
                                               double& s2) {
    __shared__ double sm[2][8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    c = blockIdx.x * 32 + lane;
    double a1 = 0.0, a2 = 0.0;
    if (c < C)
        for (int b = w; b < nblocks; b += 8) {
            a1 += partial[((int64_t)b * 2 + 0) * C + c];
            a2 += partial[((int64_t)b * 2 + 1) * C + c];
        }
    sm[0][w][lane] = a1;
    sm[1][w][lane] = a2;
    __syncthreads();
    s1 = s2 = 0.0;
    if (w != 0 || c >= C) return false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s1 += sm[0][k][lane];
        s2 += sm[1][k][lane];
    }
    return true;
}

// training: mean/invstd from batch stats; updates running stats (unbiased var) -- eval: from running stats
// launch: cdiv(C, 32) CTAs of 256 threads
__global__ void bn_finalize_kernel(const double* __restrict__ partial, int nblocks, int64_t n, int C, int training,
                                   float momentum, float eps, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
    int c;
    double s1, s2;
    if (!bn_column_sums(partial, training ? nblocks : 0, C, c, s1, s2)) return;
    if (training) {
        const double mean = s1 / (double)n;
        double var = s2 / (double)n - mean * mean;
        if (var < 0.0) var = 0.0;
        save_mean[c] = (float)mean;
        save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) {
            const double var_u = n > 1 ? var * (double)n / (double)(n - 1) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * var_u);
        }
    } else {
        save_mean[c] = running_mean[c];
        save_invstd[c] = 1.f / sqrtf(running_var[c] + eps);
    }
}
__global__ void bn_apply_kernel(const float* __restrict__ x, int64_t n, int C, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float slope, float* __restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tot = n * C;
    for (; i < tot; i += stride) {
        const int c = (int)(i % C);
        float v = (x[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
        y[i] = v > 0.f ? v : v * slope;
    }
}

MSHA_API size_t msha_bn_workspace_bytes(int C) { return (size_t)BN_BLOCKS * 2 * C * sizeof(double); }

// y = lrelu(bn(x)).  save_mean/save_invstd: float[C] outputs used by the backward.
MSHA_API int msha_bn_lrelu_fwd(const float* x, int64_t n, int C, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, int training, float momentum, float eps,
                               float slope, float* y, float* save_mean, float* save_invstd, void* ws, size_t ws_bytes,
                               void* stream) {
    MSHA_REQUIRE(n >= 1 && C >= 1, "bn: bad shape");
    MSHA_REQUIRE(training || (running_mean && running_var), "bn: eval mode needs running stats");
    MSHA_REQUIRE(ws_bytes >= msha_bn_workspace_bytes(C), "bn: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (int)(n < BN_BLOCKS ? n : BN_BLOCKS);
    if (training) {
        bn_partial_kernel<<<nb, 256, 0, st>>>(x, nullptr, n, C, (double*)ws);
        MSHA_LAUNCH_OK();
    }
    bn_finalize_kernel<<<(unsigned)msha_cdiv(C, 32), 256, 0, st>>>((const double*)ws, nb, n, C, training, momentum, eps,
                                                                  running_mean, running_var, save_mean, save_invstd);
    MSHA_LAUNCH_OK();
    bn_apply_kernel<<<ew_grid(n * C), 256, 0, st>>>(x, n, C, save_mean, save_invstd, gamma, beta, slope, y);
    MSHA_LAUNCH_OK();
    return 0;
}

// g = dy * lrelu'(y) ; xhat = (x-mean)*invstd
__global__ void bn_bwd_prep_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                   int64_t n, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                                   float slope, float* __restrict__ g, float* __restrict__ xhat) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tot = n * C;
    for (; i < tot; i += stride) {
        const int c = (int)(i % C);
        g[i] = dy[i] * (y[i] > 0.f ? 1.f : slope);
        xhat[i] = (x[i] - mean[c]) * invstd[c];
    }
}
__global__ void bn_bwd_reduce_kernel(const double* __restrict__ partial, int nblocks, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
    int c;
    double s1, s2;
    if (!bn_column_sums(partial, nblocks, C, c, s1, s2)) return;
    dbeta[c] = (float)s1;
    dgamma[c] = (float)s2;
}
__global__ void bn_bwd_apply_kernel(float* __restrict__ g_io, const float* __restrict__ xhat, int64_t n, int C,
                                    const float* __restrict__ gamma, const float* __restrict__ invstd,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, int training) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tot = n * C;
    const float inv_n = 1.f / (float)n;
    for (; i < tot; i += stride) {
        const int c = (int)(i % C);
        float g = g_io[i];
        if (training) g = g - dbeta[c] * inv_n - xhat[i] * dgamma[c] * inv_n;
        g_io[i] = g * gamma[c] * invstd[c];
    }
}

// dx may alias nothing else; xhat is scratch [n*C].  dgamma/dbeta: float[C].
MSHA_API int msha_bn_lrelu_bwd(const float* dy, const float* y, const float* x, int64_t n, int C, const float* gamma,
                               const float* save_mean, const float* save_invstd, int training, float slope, float* dx,
                               float* xhat, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n >= 1 && C >= 1, "bn_bwd: bad shape");
    MSHA_REQUIRE(ws_bytes >= msha_bn_workspace_bytes(C), "bn_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (int)(n < BN_BLOCKS ? n : BN_BLOCKS);
    bn_bwd_prep_kernel<<<ew_grid(n * C), 256, 0, st>>>(dy, y, x, n, C, save_mean, save_invstd, slope, dx, xhat);
    MSHA_LAUNCH_OK();
    bn_partial_kernel<<<nb, 256, 0, st>>>(dx, xhat, n, C, (double*)ws);
    MSHA_LAUNCH_OK();
    bn_bwd_reduce_kernel<<<(unsigned)msha_cdiv(C, 32), 256, 0, st>>>((const double*)ws, nb, C, dgamma, dbeta);
    MSHA_LAUNCH_OK();
    bn_bwd_apply_kernel<<<ew_grid(n * C), 256, 0, st>>>(dx, xhat, n, C, gamma, save_invstd, dgamma, dbeta, training);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---- the same BatchNorm over a node axis that is partitioned across GPUs ("ncclAllReduce of the
// tiny BN statistics"): the column sums leave the library between the statistics pass and the apply pass, the caller
// all-reduces the 2*C doubles over the ranks and passes the GLOBAL row count.  Same kernels as the fused entry points.
__global__ void bn_sums_kernel(const double* __restrict__ partial, int nblocks, int C, double* __restrict__ sums) {
    int c;
    double s1, s2;
    if (!bn_column_sums(partial, nblocks, C, c, s1, s2)) return;
    sums[c] = s1;
    sums[C + c] = s2;
}
__global__ void bn_finalize_sums_kernel(const double* __restrict__ sums, int64_t n, int C, int training, float momentum,
                                        float eps, float* __restrict__ running_mean, float* __restrict__ running_var,
                                        float* __restrict__ save_mean, float* __restrict__ save_invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (training) {
        const double mean = sums[c] / (double)n;
        double var = sums[C + c] / (double)n - mean * mean;
        if (var < 0.0) var = 0.0;
        save_mean[c] = (float)mean;
        save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) {
            const double var_u = n > 1 ? var * (double)n / (double)(n - 1) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * var_u);
        }
    } else {
        save_mean[c] = running_mean[c];
        save_invstd[c] = 1.f / sqrtf(running_var[c] + eps);
    }
}
__global__ void bn_bwd_apply_sums_kernel(float* __restrict__ g_io, const float* __restrict__ xhat, int64_t n, int C,
                                         const float* __restrict__ gamma, const float* __restrict__ invstd,
                                         const double* __restrict__ sums, double inv_n_total, int training) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tot = n * C;
    for (; i < tot; i += stride) {
        const int c = (int)(i % C);
        float g = g_io[i];
        if (training) g = g - (float)(sums[c] * inv_n_total) - xhat[i] * (float)(sums[C + c] * inv_n_total);
        g_io[i] = g * gamma[c] * invstd[c];
    }
}
__global__ void bn_sums_to_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                        float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)sums[c];
    dgamma[c] = (float)sums[C + c];
}

// sums: double[2*C] = (sum_rows x, sum_rows x^2) of the LOCAL rows (n may be 0: zeros)
MSHA_API int msha_bn_stats(const float* x, int64_t n, int C, double* sums, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n >= 0 && C >= 1 && sums != nullptr, "bn_stats: bad arguments");
    MSHA_REQUIRE(ws_bytes >= msha_bn_workspace_bytes(C), "bn_stats: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (int)(n < BN_BLOCKS ? n : BN_BLOCKS);
    if (nb > 0) {
        bn_partial_kernel<<<nb, 256, 0, st>>>(x, nullptr, n, C, (double*)ws);
        MSHA_LAUNCH_OK();
    }
    bn_sums_kernel<<<(unsigned)msha_cdiv(C, 32), 256, 0, st>>>((const double*)ws, nb, C, sums);
    MSHA_LAUNCH_OK();
    return 0;
}
// y = lrelu(bn(x)) of the n local rows with the statistics of all n_total rows (sums already reduced over the ranks)
MSHA_API int msha_bn_lrelu_apply(const float* x, int64_t n, int C, const double* sums, int64_t n_total, const float* gamma,
                                 const float* beta, float* running_mean, float* running_var, int training, float momentum,
                                 float eps, float slope, float* y, float* save_mean, float* save_invstd, void* stream) {
    MSHA_REQUIRE(n >= 0 && C >= 1 && n_total >= 1, "bn_apply: bad shape");
    MSHA_REQUIRE(training || (running_mean && running_var), "bn_apply: eval mode needs running stats");
    MSHA_REQUIRE(!training || sums != nullptr, "bn_apply: training mode needs the column sums");
    cudaStream_t st = (cudaStream_t)stream;
    bn_finalize_sums_kernel<<<(unsigned)msha_cdiv(C, 128), 128, 0, st>>>(sums, n_total, C, training, momentum, eps, running_mean,
                                                                        running_var, save_mean, save_invstd);
    MSHA_LAUNCH_OK();
    if (n > 0) {
        bn_apply_kernel<<<ew_grid(n * C), 256, 0, st>>>(x, n, C, save_mean, save_invstd, gamma, beta, slope, y);
        MSHA_LAUNCH_OK();
    }
    return 0;
}
// backward, first half: g = dy * lrelu'(y) -> dx (scratch), xhat, and sums = (sum g, sum g*xhat) of the local rows
MSHA_API int msha_bn_lrelu_bwd_stats(const float* dy, const float* y, const float* x, int64_t n, int C,
                                     const float* save_mean, const float* save_invstd, float slope, float* dx, float* xhat,
                                     double* sums, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n >= 0 && C >= 1 && sums != nullptr, "bn_bwd_stats: bad arguments");
    MSHA_REQUIRE(ws_bytes >= msha_bn_workspace_bytes(C), "bn_bwd_stats: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (int)(n < BN_BLOCKS ? n : BN_BLOCKS);
    if (n > 0) {
        bn_bwd_prep_kernel<<<ew_grid(n * C), 256, 0, st>>>(dy, y, x, n, C, save_mean, save_invstd, slope, dx, xhat);
        MSHA_LAUNCH_OK();
        bn_partial_kernel<<<nb, 256, 0, st>>>(dx, xhat, n, C, (double*)ws);
        MSHA_LAUNCH_OK();
    }
    bn_sums_kernel<<<(unsigned)msha_cdiv(C, 32), 256, 0, st>>>((const double*)ws, nb, C, sums);
    MSHA_LAUNCH_OK();
    return 0;
}
// backward, second half with the reduced sums: dx in place, dgamma / dbeta (the global ones, identical on every rank)
MSHA_API int msha_bn_lrelu_bwd_apply(float* dx, const float* xhat, int64_t n, int C, const float* gamma,
                                     const float* save_invstd, const double* sums, int64_t n_total, int training,
                                     float* dgamma, float* dbeta, void* stream) {
    MSHA_REQUIRE(n >= 0 && C >= 1 && n_total >= 1 && sums != nullptr, "bn_bwd_apply: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bn_sums_to_grads_kernel<<<(unsigned)msha_cdiv(C, 128), 128, 0, st>>>(sums, C, dgamma, dbeta);
    MSHA_LAUNCH_OK();
    if (n > 0) {
        bn_bwd_apply_sums_kernel<<<ew_grid(n * C), 256, 0, st>>>(dx, xhat, n, C, gamma, save_invstd, sums, 1.0 / (double)n_total,
                                                                 training);
        MSHA_LAUNCH_OK();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// row log_softmax with optional leading ELU (Ours.py:166-167: log_softmax(elu(x)))
// ---------------------------------------------------------------------------------------------
__global__ void logsoftmax_fwd_kernel(const float* __restrict__ x, int64_t n, int M, int pre_elu, float* __restrict__ y) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* xr = x + row * M;
    float mx = -INFINITY;
    for (int j = lane; j < M; j += 32) {
        float v = xr[j];
        if (pre_elu) v = v > 0.f ? v : expm1f(v);
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < M; j += 32) {
        float v = xr[j];
        if (pre_elu) v = v > 0.f ? v : expm1f(v);
        s += expf(v - mx);
    }
    s = warp_sum(s);
    const float lse = mx + logf(s);
    for (int j = lane; j < M; j += 32) {
        float v = xr[j];
        if (pre_elu) v = v > 0.f ? v : expm1f(v);
        y[row * M + j] = v - lse;
    }
}
// dx = (dy - exp(y)*sum(dy)) * (pre_elu ? elu'(x) : 1)
__global__ void logsoftmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                      const float* __restrict__ x, int64_t n, int M, int pre_elu,
                                      float* __restrict__ dx) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float s = 0.f;
    for (int j = lane; j < M; j += 32) s += dy[row * M + j];
    s = warp_sum(s);
    for (int j = lane; j < M; j += 32) {
        float g = dy[row * M + j] - expf(y[row * M + j]) * s;
        if (pre_elu) {
            const float xv = x[row * M + j];
            g *= xv > 0.f ? 1.f : expf(xv);
        }
        dx[row * M + j] = g;
    }
}
MSHA_API int msha_log_softmax_fwd(const float* x, int64_t n, int64_t M, int pre_elu, float* y, void* stream) {
    MSHA_REQUIRE(n >= 0 && M >= 1 && M < ((int64_t)1 << 31), "log_softmax: bad shape");
    if (n == 0) return 0;
    logsoftmax_fwd_kernel<<<(unsigned)msha_cdiv(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, n, (int)M, pre_elu, y);
    MSHA_LAUNCH_OK();
    return 0;
}
MSHA_API int msha_log_softmax_bwd(const float* dy, const float* y, const float* x, int64_t n, int64_t M, int pre_elu,
                                  float* dx, void* stream) {
    MSHA_REQUIRE(n >= 0 && M >= 1 && M < ((int64_t)1 << 31), "log_softmax_bwd: bad shape");
    if (n == 0) return 0;
    logsoftmax_bwd_kernel<<<(unsigned)msha_cdiv(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, x, n, (int)M, pre_elu, dx);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// link scorer helpers (LLP.py:105): z[p] = h_i[src[p]] * h_j[dst[p]] and its scatter-add backward
// src/dst may be NULL (identity: row p).
// ---------------------------------------------------------------------------------------------
__global__ void pair_gather_mul_kernel(const float* __restrict__ hi, const float* __restrict__ hj,
                                       const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t P, int C,
                                       float* __restrict__ z) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const int64_t a = src ? src[p] : p, b = dst ? dst[p] : p;
    if ((C & 3) == 0) {
        for (int c = lane * 4; c < C; c += 128) {
            float4 x = ldg4(hi + a * C + c), y = ldg4(hj + b * C + c);
            *reinterpret_cast<float4*>(z + p * C + c) = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
        }
    } else {
        for (int c = lane; c < C; c += 32) z[p * C + c] = hi[a * C + c] * hj[b * C + c];
    }
}
__global__ void pair_scatter_mul_kernel(const float* __restrict__ dz, const float* __restrict__ hi,
                                        const float* __restrict__ hj, const int64_t* __restrict__ src,
                                        const int64_t* __restrict__ dst, int64_t P, int C, float* __restrict__ dhi,
                                        float* __restrict__ dhj) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const int64_t a = src ? src[p] : p, b = dst ? dst[p] : p;
    for (int c = lane; c < C; c += 32) {
        const float g = dz[p * C + c];
        atomicAdd(dhi + a * C + c, g * hj[b * C + c]);
        atomicAdd(dhj + b * C + c, g * hi[a * C + c]);
    }
}
MSHA_API int msha_pair_gather_mul(const float* hi, const float* hj, const int64_t* src, const int64_t* dst, int64_t P,
                                  int64_t C, float* z, void* stream) {
    MSHA_REQUIRE(P >= 0 && C >= 1 && C < ((int64_t)1 << 31), "pair_gather_mul: bad shape");
    if (P == 0) return 0;
    pair_gather_mul_kernel<<<(unsigned)msha_cdiv(P * 32, 256), 256, 0, (cudaStream_t)stream>>>(hi, hj, src, dst, P, (int)C, z);
    MSHA_LAUNCH_OK();
    return 0;
}
// dhi/dhj must be zero-initialised (or hold the running gradient): the kernel accumulates with atomics.
MSHA_API int msha_pair_scatter_mul_add(const float* dz, const float* hi, const float* hj, const int64_t* src,
                                       const int64_t* dst, int64_t P, int64_t C, float* dhi, float* dhj, void* stream) {
    MSHA_REQUIRE(P >= 0 && C >= 1 && C < ((int64_t)1 << 31), "pair_scatter_mul_add: bad shape");
    if (P == 0) return 0;
    pair_scatter_mul_kernel<<<(unsigned)msha_cdiv(P * 32, 256), 256, 0, (cudaStream_t)stream>>>(dz, hi, hj, src, dst, P, (int)C,
                                                                                           dhi, dhj);
    MSHA_LAUNCH_OK();
    return 0;
}

// 'inner' predictor (LLP.py:112-113) and the bilinear pair read-out elu(u_i . v_j) (Ours.py:108-109 entries)
// out[p] = act( sum_c hi[src[p],c]*hj[dst[p],c] )
__global__ void pair_dot_kernel(const float* __restrict__ hi, const float* __restrict__ hj, const int64_t* __restrict__ src,
                                const int64_t* __restrict__ dst, int64_t P, int C, int act, float* __restrict__ out) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const int64_t a = src ? src[p] : p, b = dst ? dst[p] : p;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(hi[a * C + c], hj[b * C + c], s);
    s = warp_sum(s);
    if (lane == 0) out[p] = apply_act(s, act, 0.f);
}
// dhi[src[p]] += g[p]*hj[dst[p]] etc. with g = dout*act'(out)
__global__ void pair_dot_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ outp,
                                    const float* __restrict__ hi, const float* __restrict__ hj,
                                    const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t P, int C,
                                    int act, float* __restrict__ dhi, float* __restrict__ dhj) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const int64_t a = src ? src[p] : p, b = dst ? dst[p] : p;
    const float g = dout[p] * act_grad_from_out(outp[p], act, 0.f);
    for (int c = lane; c < C; c += 32) {
        atomicAdd(dhi + a * C + c, g * hj[b * C + c]);
        atomicAdd(dhj + b * C + c, g * hi[a * C + c]);
    }
}
MSHA_API int msha_pair_dot(const float* hi, const float* hj, const int64_t* src, const int64_t* dst, int64_t P, int64_t C,
                           int act, float* out, void* stream) {
    MSHA_REQUIRE(P >= 0 && C >= 1 && C < ((int64_t)1 << 31), "pair_dot: bad shape");
    if (P == 0) return 0;
    pair_dot_kernel<<<(unsigned)msha_cdiv(P * 32, 256), 256, 0, (cudaStream_t)stream>>>(hi, hj, src, dst, P, (int)C, act, out);
    MSHA_LAUNCH_OK();
    return 0;
}
MSHA_API int msha_pair_dot_bwd(const float* dout, const float* out, const float* hi, const float* hj, const int64_t* src,
                               const int64_t* dst, int64_t P, int64_t C, int act, float* dhi, float* dhj, void* stream) {
    MSHA_REQUIRE(P >= 0 && C >= 1 && C < ((int64_t)1 << 31), "pair_dot_bwd: bad shape");
    if (P == 0) return 0;
    pair_dot_bwd_kernel<<<(unsigned)msha_cdiv(P * 32, 256), 256, 0, (cudaStream_t)stream>>>(dout, out, hi, hj, src, dst, P, (int)C,
                                                                                        act, dhi, dhj);
    MSHA_LAUNCH_OK();
    return 0;
}

// Philox uniform negative pairs (oracle: negative_sample): pair p <- words 2p, 2p+1 of stream 1.
__global__ void negative_sample_kernel(uint64_t seed, int64_t P, uint32_t n_src, uint32_t n_dst, int64_t* __restrict__ src,
                                       int64_t* __restrict__ dst) {
    // one thread per Philox block = 2 pairs
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 2 >= P) return;
    seed = drop_seed_eff(seed);        // folds in the device-side epoch: a replayed CUDA graph draws fresh negatives (epoch 0: unchanged)
    Philox4 r = philox4x32_10((uint32_t)b, (uint32_t)((uint64_t)b >> 32), 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    src[2 * b] = (int64_t)(((uint64_t)r.x * n_src) >> 32);
    dst[2 * b] = (int64_t)(((uint64_t)r.y * n_dst) >> 32);
    if (2 * b + 1 < P) {
        src[2 * b + 1] = (int64_t)(((uint64_t)r.z * n_src) >> 32);
        dst[2 * b + 1] = (int64_t)(((uint64_t)r.w * n_dst) >> 32);
    }
}
MSHA_API int msha_negative_sample(uint64_t seed, int64_t P, int64_t n_src, int64_t n_dst, int64_t* src, int64_t* dst,
                                  void* stream) {
    MSHA_REQUIRE(P >= 0 && n_src >= 1 && n_dst >= 1 && n_src < ((int64_t)1 << 32) && n_dst < ((int64_t)1 << 32),
                 "negative_sample: bad arguments");
    if (P == 0) return 0;
    negative_sample_kernel<<<(unsigned)msha_cdiv(msha_cdiv(P, 2), 256), 256, 0, (cudaStream_t)stream>>>(
        seed, P, (uint32_t)n_src, (uint32_t)n_dst, src, dst);
    MSHA_LAUNCH_OK();
    return 0;
}

// dropout keep-mask stream as bytes (tests; oracle: dropout_keep_mask)
__global__ void dropout_mask_kernel(uint64_t seed, uint32_t stream_id, int64_t n, uint32_t thr, uint8_t* __restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keep[i] = philox_word(drop_seed_eff(seed), stream_id, (uint64_t)i) >= thr ? 1 : 0;
}
MSHA_API int msha_dropout_mask(uint64_t seed, uint32_t stream_id, int64_t n, float p, uint8_t* keep, void* stream) {
    if (n <= 0) return 0;
    DropArgsD d = make_drop_d(p, seed, stream_id);
    dropout_mask_kernel<<<(unsigned)msha_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, n, d.thr, keep);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// a-8 loss read-out: F.nll_loss(logp, target) with mean reduction (train.py:229, LLP.py:235)
//   forward : loss = -(1/P) * sum_p logp[p, target[p]]          (two-stage deterministic sum)
//   backward: dlogp = 0 except dlogp[p, target[p]] = -g/P       (one streaming pass writes the dense gradient)
// ---------------------------------------------------------------------------------------------
constexpr int NLL_BLOCKS = 592;
__global__ void nll_fwd_stage1(const float* __restrict__ logp, const int64_t* __restrict__ target, int64_t P, int C,
                               double* __restrict__ partial, int32_t* __restrict__ status) {
    __shared__ double sm[8];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // four pairs per iteration: the picked scores are one scattered 32 B sector each, so the labels and then the scores of
    // four pairs are requested together instead of one dependent load pair at a time
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += 4 * stride) {
        int64_t t[4];
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = p + k * stride < P ? target[p + k * stride] : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t pk = p + k * stride;
            const bool in_range = t[k] >= 0 && t[k] < C;
            v[k] = (pk < P && in_range) ? logp[pk * C + t[k]] : 0.f;
            if (pk < P && !in_range) atomicOr(status, 1);
        }
        acc += ((double)v[0] + (double)v[1]) + ((double)v[2] + (double)v[3]);
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[w];
        partial[blockIdx.x] = s;
    }
}
__global__ void nll_fwd_stage2(const double* __restrict__ partial, int n, int64_t P, float* __restrict__ loss) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
    s = warp_sum_d(s);
    if (threadIdx.x == 0) *loss = (float)(-s / (double)P);
}
__global__ void nll_bwd_kernel(const int64_t* __restrict__ target, const float* __restrict__ gout, int64_t P, int C,
                               float* __restrict__ dlogp) {
    // one warp per row: stream zeros (write-once data, streaming stores), patch the picked column
    const float g = -(*gout) / (float)P;
    const int lane = threadIdx.x & 31, c4n = C >> 2;        // C % 4 == 0 fast path handled by the caller
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < P; p += nwarps) {
        const int t = (int)__ldg(target + p);
        float4* d4 = reinterpret_cast<float4*>(dlogp + p * C);
        for (int c4 = lane; c4 < c4n; c4 += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((t >> 2) == c4) {
                const int r = t & 3;
                if (r == 0) v.x = g; else if (r == 1) v.y = g; else if (r == 2) v.z = g; else v.w = g;
            }
            __stcs(d4 + c4, v);
        }
    }
}
__global__ void nll_bwd_scalar_kernel(const int64_t* __restrict__ target, const float* __restrict__ gout, int64_t P, int C,
                                      float* __restrict__ dlogp) {
    const float g = -(*gout) / (float)P;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        dlogp[i] = ((int64_t)(i - p * C) == target[p]) ? g : 0.f;
    }
}
MSHA_API size_t msha_nll_workspace_bytes(void) { return NLL_BLOCKS * sizeof(double); }
// loss: float[1] device; status: int32[1] device (bit0: a target was out of range)
MSHA_API int msha_nll_loss_fwd(const float* logp, const int64_t* target, int64_t P, int64_t C, float* loss, int32_t* status,
                               void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(P >= 1 && C >= 1 && C < ((int64_t)1 << 31), "nll_loss: bad shape");
    MSHA_REQUIRE(ws_bytes >= msha_nll_workspace_bytes(), "nll_loss: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    MSHA_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    int nb = (int)(msha_cdiv(P, 256) < NLL_BLOCKS ? msha_cdiv(P, 256) : NLL_BLOCKS);
    nll_fwd_stage1<<<nb, 256, 0, st>>>(logp, target, P, (int)C, (double*)ws, status);
    MSHA_LAUNCH_OK();
    nll_fwd_stage2<<<1, 32, 0, st>>>((const double*)ws, nb, P, loss);
    MSHA_LAUNCH_OK();
    return 0;
}
// gout: float[1] device (upstream gradient of the scalar loss)
MSHA_API int msha_nll_loss_bwd(const int64_t* target, const float* gout, int64_t P, int64_t C, float* dlogp, void* stream) {
    MSHA_REQUIRE(P >= 1 && C >= 1 && C < ((int64_t)1 << 31), "nll_loss_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = MSHA_NUM_SMS * 16;
    if ((C & 3) == 0 && (((uintptr_t)dlogp) & 15) == 0)
        nll_bwd_kernel<<<grid, 256, 0, st>>>(target, gout, P, (int)C, dlogp);
    else
        nll_bwd_scalar_kernel<<<grid, 256, 0, st>>>(target, gout, P, (int)C, dlogp);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Linear-layer backward prologue: g = dy * act'(y) written once, and db = column sums of g in the same pass
// (bias gradient of lin(x), LLP.py:108) -- one streaming read of dy,y and one write of g.
// ---------------------------------------------------------------------------------------------
constexpr int ABC_BLOCKS = 592;
__global__ void act_bwd_colsum_stage1(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ g,
                                      int64_t n, int C, int act, float slope, double* __restrict__ partial) {
    const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * rows_per;
    const int64_t r1 = r0 + rows_per < n ? r0 + rows_per : n;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double acc = 0.0;
        float run = 0.f;                       // short fp32 runs folded into the fp64 accumulator
        int cnt = 0;
        for (int64_t r = r0; r < r1; ++r) {
            const float v = dy[r * C + c] * act_grad_from_out(y[r * C + c], act, slope);
            g[r * C + c] = v;
            run += v;
            if (++cnt == 64) { acc += (double)run; run = 0.f; cnt = 0; }
        }
        partial[(int64_t)blockIdx.x * C + c] = acc + (double)run;
    }
}
__global__ void colsum_stage2(const double* __restrict__ partial, int nblocks, int C, float* __restrict__ out) {
    // 32 columns per CTA, 8 warps each summing every 8th partial, fixed order (see colreduce_stage2 in gat_kernels.cu)
    __shared__ double sm[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    double acc = 0.0;
    if (c < C)
        for (int b = w; b < nblocks; b += 8) acc += partial[(int64_t)b * C + c];
    sm[w][lane] = acc;
    __syncthreads();
    if (w == 0 && c < C) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][lane];
        out[c] = (float)t;
    }
}
MSHA_API size_t msha_act_bwd_colsum_workspace_bytes(int C) { return (size_t)ABC_BLOCKS * C * sizeof(double); }
MSHA_API int msha_act_bwd_colsum(const float* dy, const float* y, float* g, int64_t n, int C, int act, float slope,
                                 float* colsum, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n >= 1 && C >= 1, "act_bwd_colsum: bad shape");
    MSHA_REQUIRE(ws_bytes >= msha_act_bwd_colsum_workspace_bytes(C), "act_bwd_colsum: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = (int)(n < ABC_BLOCKS ? n : ABC_BLOCKS);
    act_bwd_colsum_stage1<<<nb, C >= 256 ? 256 : (C >= 128 ? 128 : 64), 0, st>>>(dy, y, g, n, C, act, slope, (double*)ws);
    MSHA_LAUNCH_OK();
    colsum_stage2<<<(unsigned)msha_cdiv(C, 32), 256, 0, st>>>((const double*)ws, nb, C, colsum);
    MSHA_LAUNCH_OK();
    return 0;
}
