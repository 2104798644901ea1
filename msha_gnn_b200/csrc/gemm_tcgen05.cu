// K-7: fp32-accurate GEMM on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) -- the dense contractions of the hot path:
// Wh = X @ W (GAT.py:21, Ours.py:57-58), the LinkPredictor Linear layers (LLP.py:107-108) and their backward.
//
// fp32 accuracy (north star: rel. err <= 1e-4) via the 3xTF32 split done in shared memory by converter warps:
//     x = hi + lo,  hi = x & 0xffffe000 (exact in tf32),  lo = x - hi
//     A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi            (error ~ 2^-21 relative)
//
// One persistent CTA per SM, 384 threads, warp-specialised:
//     warp 0   TMA producer          (one elected lane)         raw fp32 tiles -> smem
//     warp 1   MMA issuer            (one elected lane)         3 x tcgen05.mma per 8-wide k-step
//     warp 2   TMEM allocator
//     warps 4-7   converters         split raw -> hi (in place) / lo, fence.proxy.async
//     warps 8-15  epilogue           tcgen05.ld -> bias/activation -> smem transpose -> coalesced global stores
//                                    (or atomicAdd for split-K)
// Pipelines: smem ring (full_raw -> full_cvt -> empty), TMEM double buffer (tmem_full / tmem_empty).
//
// Operand layouts (row-major in global memory):
//     K-major : stored [MN, K] (K contiguous)  -> one TMA box {16 K, rows},  UMMA K-major SWIZZLE_64B
//     MN-major: stored [K, MN] (MN contiguous) -> boxes {32 MN, 16 K},       UMMA MN-major SWIZZLE_128B_BASE32B
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 16;                  // 16 fp32 = 64 B rows (SWIZZLE_64B); fine-grained stages keep 4+ in flight
constexpr int UMMA_K = 8;                    // tf32
constexpr int NUM_THREADS = 512;
constexpr int CVT_THREADS = 128;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4; // per epilogue warp: 32x32 fp32 transpose buffer

__device__ __forceinline__ float epi_act(float x, int act, float slope) {
    switch (act) {
        case 1: return x > 0.f ? x : expm1f(x);
        case 2: return fmaxf(x, 0.f);
        case 3: return 1.f / (1.f + expf(-fmaxf(x, 0.f)));
        case 4: return x > 0.f ? x : x * slope;
        case 5: return 1.f / (1.f + expf(-x));
        default: return x;
    }
}

template <int BLOCK_N>
struct Cfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;          // 8 KB
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);    // hi + lo of both operands
    static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;   // two accumulator stages
    static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = EPI_OFF + EPI_WARPS * EPI_STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 /*align slack*/ + 512 /*barriers*/;
};

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   float* __restrict__ Cmat, int M, int N, int K, int64_t ldc, const float* __restrict__ bias, int act,
                   float slope, int splits, int atomic_out) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + C::BAR_OFF);
    uint64_t* full_raw = bars;                       // [STAGES]
    uint64_t* full_cvt = bars + C::STAGES;           // [STAGES]
    uint64_t* empty = bars + 2 * C::STAGES;          // [STAGES]
    uint64_t* tmem_full = bars + 3 * C::STAGES;      // [2]
    uint64_t* tmem_empty = bars + 3 * C::STAGES + 2; // [2]
    uint32_t* tmem_ptr = (uint32_t*)(bars + 3 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(smem_u32(&full_raw[s]), 1);
            mbar_init(smem_u32(&full_cvt[s]), CVT_THREADS);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tmem_full[a]), 1);
            mbar_init(smem_u32(&tmem_empty[a]), EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int n_work = m_tiles * n_tiles * splits;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int split = w % splits, t = w / splits;
                const int n0 = (t % n_tiles) * BLOCK_N, m0 = (t / n_tiles) * BLOCK_M;
                const int kb0 = (int)((int64_t)split * total_kb / splits), kb1 = (int)((int64_t)(split + 1) * total_kb / splits);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    uint8_t* st = smem + stage * C::STAGE_BYTES;
                    const uint32_t a_dst = smem_u32(st), b_dst = smem_u32(st + 2 * C::A_BYTES);
                    const uint32_t bar = smem_u32(&full_raw[stage]);
                    mbar_arrive_expect_tx(bar, C::A_BYTES + C::B_BYTES);
                    const int k0 = kb * BLOCK_K;
                    if (A_MN) {
#pragma unroll
                        for (int i = 0; i < BLOCK_M / 32; ++i) tma_load_2d(a_dst + i * (128 * BLOCK_K), &tmA, bar, m0 + 32 * i, k0);
                    } else {
                        tma_load_2d(a_dst, &tmA, bar, k0, m0);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int i = 0; i < BLOCK_N / 32; ++i) tma_load_2d(b_dst + i * (128 * BLOCK_K), &tmB, bar, n0 + 32 * i, k0);
                    } else {
                        tma_load_2d(b_dst, &tmB, bar, k0, n0);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, A_MN, B_MN);
            constexpr uint32_t A_LBO = A_MN ? 128 * BLOCK_K : 16, A_SBO = 512, A_KSTEP = A_MN ? 1024 : UMMA_K * 4;
            constexpr uint32_t B_LBO = B_MN ? 128 * BLOCK_K : 16, B_SBO = 512, B_KSTEP = B_MN ? 1024 : UMMA_K * 4;
            constexpr uint32_t A_LT = A_MN ? 1 : 4, B_LT = B_MN ? 1 : 4;
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int split = w % splits;
                const int kb0 = (int)((int64_t)split * total_kb / splits), kb1 = (int)((int64_t)(split + 1) * total_kb / splits);
                mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&full_cvt[stage]), phase);
                    tcgen05_fence_after();
                    uint8_t* st = smem + stage * C::STAGE_BYTES;
                    const uint32_t a_hi = smem_u32(st), a_lo = a_hi + C::A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * C::A_BYTES, b_lo = b_hi + C::B_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t dah = make_smem_desc(a_hi + k * A_KSTEP, A_LBO, A_SBO, A_LT);
                        const uint64_t dal = make_smem_desc(a_lo + k * A_KSTEP, A_LBO, A_SBO, A_LT);
                        const uint64_t dbh = make_smem_desc(b_hi + k * B_KSTEP, B_LBO, B_SBO, B_LT);
                        const uint64_t dbl = make_smem_desc(b_lo + k * B_KSTEP, B_LBO, B_SBO, B_LT);
                        umma_tf32(tmem_d, dal, dbh, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                    }
                    tcgen05_commit(smem_u32(&empty[stage]));          // frees the smem slot when the MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(smem_u32(&tmem_full[acc]));            // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // =============================== converters: raw -> hi / lo ===============================
        const int tid = threadIdx.x - 128;
        uint32_t stage = 0, phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int split = w % splits;
            const int kb0 = (int)((int64_t)split * total_kb / splits), kb1 = (int)((int64_t)(split + 1) * total_kb / splits);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&full_raw[stage]), phase);
                uint8_t* st = smem + stage * C::STAGE_BYTES;
                const uint32_t a_hi = smem_u32(st), b_hi = a_hi + 2 * C::A_BYTES;
                split_tile_inplace(a_hi, a_hi + C::A_BYTES, C::A_BYTES / 16, tid, CVT_THREADS);
                split_tile_inplace(b_hi, b_hi + C::B_BYTES, C::B_BYTES / 16, tid, CVT_THREADS);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
                mbar_arrive(smem_u32(&full_cvt[stage]));
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // =============================== epilogue ===============================
        // warp -> TMEM lane quarter q (rows 32q..32q+31 of the tile) and column half `hf`
        const int q = warp & 3, hf = (warp - 8) >> 2;
        constexpr int COLS_PER_WARP = BLOCK_N / 2;
        const uint32_t stage_s = smem_u32(smem + C::EPI_OFF + (warp - 8) * EPI_STAGE_BYTES);
        const bool vec_ok = ((ldc & 3) == 0) && ((((uintptr_t)Cmat) & 15) == 0);
        uint32_t acc = 0, acc_phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int t = w / splits;
            const int n0 = (t % n_tiles) * BLOCK_N, m0 = (t / n_tiles) * BLOCK_M;
            mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
            tcgen05_fence_after();
            const int row0 = m0 + q * 32;
#pragma unroll 1
            for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
                const int c0 = hf * COLS_PER_WARP + cc;
                const int nb = n0 + c0;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + c0, v);
                if (nb >= N) continue;                         // warp-uniform
                if (atomic_out) {                              // split-K partial sums
                    const int row = row0 + lane;
                    if (row < M) {
                        float* crow = Cmat + (int64_t)row * ldc;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < N) atomicAdd(crow + nb + j, __uint_as_float(v[j]));
                    }
                    continue;
                }
                float bj = 0.f;
                if (bias != nullptr && nb + lane < N) bj = __ldg(bias + nb + lane);
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bj, j);
                switch (act) {                                 // hoisted: one branch per 32 values
                    case 1:
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = x[j] > 0.f ? x[j] : expm1f(x[j]);
                        break;
                    case 2:
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                        break;
                    case 3:
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = __fdividef(1.f, 1.f + __expf(-fmaxf(x[j], 0.f)));
                        break;
                    case 4:
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = x[j] > 0.f ? x[j] : x[j] * slope;
                        break;
                    case 5:
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = __fdividef(1.f, 1.f + __expf(-x[j]));
                        break;
                    default: break;
                }
                // transpose through shared memory (16-byte chunks XOR-swizzled by row) -> coalesced 128 B row stores
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    sts128(stage_s + (uint32_t)(lane * 32 + ((g ^ (lane & 7)) << 2)) * 4u,
                           make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]));
                __syncwarp();
                const int rs = lane >> 3, cg = lane & 7;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = 4 * k + rs;
                    const float4 o = lds128(stage_s + (uint32_t)(r * 32 + ((cg ^ (r & 7)) << 2)) * 4u);
                    const int row = row0 + r, col = nb + 4 * cg;
                    if (row < M) {
                        float* cp = Cmat + (int64_t)row * ldc + col;
                        if (vec_ok && col + 4 <= N) {
                            *reinterpret_cast<float4*>(cp) = o;
                        } else {
                            if (col + 0 < N) cp[0] = o.x;
                            if (col + 1 < N) cp[1] = o.y;
                            if (col + 2 < N) cp[2] = o.z;
                            if (col + 3 < N) cp[3] = o.w;
                        }
                    }
                }
                __syncwarp();
            }
            tcgen05_fence_before();
            mbar_arrive(smem_u32(&tmem_empty[acc]));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN, bool AM, bool BM>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, float* C, int M, int N, int K, int64_t ldc,
                  const float* bias, int act, float slope, int splits, int atomic_out, cudaStream_t st) {
    auto kern = gemm_tf32x3_kernel<BN, AM, BM>;
    static MshaPerDeviceOnce attr_set;
    if (attr_set.need()) {
        MSHA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
        attr_set.mark();
    }
    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BN - 1) / BN;
    int64_t n_work = (int64_t)m_tiles * n_tiles * splits;
    int grid = (int)(n_work < MSHA_NUM_SMS ? n_work : MSHA_NUM_SMS);
    kern<<<grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, st>>>(ta, tb, C, M, N, K, ldc, bias, act, slope, splits, atomic_out);
    MSHA_LAUNCH_OK();
    return 0;
}

}  // namespace

// 0 if the operands satisfy the TMA constraints of msha_gemm_tf32x3 (16-byte aligned base, ld % 4 == 0)
MSHA_API int msha_gemm_tf32x3_supported(const float* A, const float* B, int64_t M, int64_t N, int64_t K, int64_t lda,
                                        int64_t ldb) {
    if (M <= 0 || N <= 0 || K <= 0) return -1;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 3) || (ldb & 3)) return -1;
    if (M >= ((int64_t)1 << 31) || N >= ((int64_t)1 << 31) || K >= ((int64_t)1 << 31)) return -1;
    return 0;
}

// C[M,N] = act( opA(A) * opB(B) + bias )   (splits == 1)
// C[M,N] += opA(A) * opB(B)                (splits  > 1: split-K partial sums are atomically added; the caller zeroes C;
//                                           bias / act must be none)
// transA == 0: A stored [M,K] (K-major); transA == 1: A stored [K,M].  transB == 0: B stored [K,N]; transB == 1: [N,K].
MSHA_API int msha_gemm_tf32x3(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                              int64_t ldb, int64_t ldc, int transA, int transB, const float* bias, int act, float slope,
                              int splits, void* stream) {
    MSHA_REQUIRE(msha_gemm_tf32x3_supported(A, B, M, N, K, lda, ldb) == 0,
                 "gemm_tf32x3: operands must be 16-byte aligned with ld %% 4 == 0 and positive sizes");
    MSHA_REQUIRE(splits >= 1, "gemm_tf32x3: splits >= 1");
    MSHA_REQUIRE(splits == 1 || (bias == nullptr && act == 0), "gemm_tf32x3: split-K excludes bias/activation");
    const int total_kb = (int)((K + BLOCK_K - 1) / BLOCK_K);
    if (splits > total_kb) splits = total_kb;
    const bool a_mn = transA != 0;      // A stored [K, M]  -> MN-major
    const bool b_mn = transB == 0;      // B stored [K, N]  -> MN-major
    // N tile: the widest that N allows, narrowed while that shortens the schedule -- a persistent CTA per SM runs
    // ceil(work / 148) rounds of tiles whose cost grows with BLOCK_N (+ a fixed part: pipeline fill, epilogue drain).
    // 4 267 x 256 (DDI layer): 34 tiles of 256 leave 114 SMs idle, 136 tiles of 64 do not.
    int BN = N > 128 ? 256 : (N > 64 ? 128 : 64);
    {
        const int64_t m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
        int64_t best = -1;
        for (int bn = BN; bn >= 64; bn >>= 1) {
            const int64_t work = m_tiles * ((N + bn - 1) / bn) * splits;
            const int64_t cost = ((work + MSHA_NUM_SMS - 1) / MSHA_NUM_SMS) * (bn + 32);
            if (best < 0 || cost < best) { best = cost; BN = bn; }
        }
    }
    CUtensorMap ta, tb;
    int rc;
    if (a_mn) rc = tc_make_map(&ta, A, M, K, lda, 32, BLOCK_K, true); else rc = tc_make_map(&ta, A, K, M, lda, BLOCK_K, BLOCK_M, false);
    if (rc) return rc;
    if (b_mn) rc = tc_make_map(&tb, B, N, K, ldb, 32, BLOCK_K, true); else rc = tc_make_map(&tb, B, K, N, ldb, BLOCK_K, BN, false);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int atomic_out = splits > 1 ? 1 : 0;
#define GO(BNv, AMv, BMv) return launch<BNv, AMv, BMv>(ta, tb, C, (int)M, (int)N, (int)K, ldc, bias, act, slope, splits, atomic_out, st)
#define GO_N(AMv, BMv) if (BN == 256) { GO(256, AMv, BMv); } else if (BN == 128) { GO(128, AMv, BMv); } else { GO(64, AMv, BMv); }
    if (!a_mn && !b_mn) { GO_N(false, false) }
    else if (!a_mn && b_mn) { GO_N(false, true) }
    else if (a_mn && !b_mn) { GO_N(true, false) }
    else { GO_N(true, true) }
#undef GO_N
#undef GO
}
