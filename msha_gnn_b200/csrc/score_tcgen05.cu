// K-9: fused link-scorer kernels on the tensor cores (LinkPredictor 'mlp' with one hidden Linear, LLP.py:104-115):
//
//   score_fwd     out[p,:] = act( (h_i[src[p]] * h_j[dst[p]]) @ W0^T + b0 )       -- LLP.py:105-110,115
//                 pair-gather + Hadamard + 3xTF32 split are done by producer warps straight into the swizzled
//                 shared-memory A tile (no Z = x_i*x_j tensor in HBM); W0 (pre-split hi/lo) arrives by TMA;
//                 tcgen05.mma accumulates in TMEM; epilogue adds bias, applies relu+sigmoid, stores coalesced.
//   score_bwd_dz  G = dOut * act'(out) (written once, + bias gradient), dZ = G @ W0 on the tensor cores, and the
//                 epilogue scatters  dh_i[src] += dZ*h_j[dst],  dh_j[dst] += dZ*h_i[src]  with 128-bit atomics.
//   score_bwd_dw  dW0 = G^T @ Z with Z regenerated from the gathers (MN-major operands, whole K range per CTA in TMEM).
//
// Same warp-specialised skeleton as gemm_tcgen05.cu (TMA thread, MMA thread, TMEM allocator; mbarrier rings;
// double-buffered TMEM accumulators).  score_fwd and score_bwd_dz run as CTA pairs (tcgen05 cta_group::2: 256-row
// tiles, half of the W0 slice staged per CTA) with 4 producer warps and 8 epilogue warps per CTA; score_bwd_dw keeps the
// whole 256 x 256 accumulator of one CTA in TMEM.  What bounds them (shared-memory bandwidth of the three MMAs per
// k-step, L2 fill of the W0 slices, epilogue latency) is measured in profiles/r01_tensor_kernel_ablations.txt.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 16;
constexpr int UMMA_K = 8;
constexpr int NUM_THREADS = 512;
constexpr int PROD_THREADS = 256;             // warps 8-15 of the dW kernel (Z gather producers)
constexpr int CVT_THREADS = 128;              // warps 4-7 of the dW kernel (TMA tile converters)
constexpr int FWD_PROD_WARPS = 4;              // forward kernel: warps 4-7 produce the A tile (pair gather, 4 rows per thread),
constexpr int FWD_EPI_WARPS = 8;               //                 warps 8-15 run the store epilogue (4 epilogue warps: 3.9 ms)
constexpr int FWD_THREADS = (4 + FWD_PROD_WARPS + FWD_EPI_WARPS) * 32;
constexpr int MAX_EPI_WARPS = 8;
#ifndef MSHA_DZ_PROD_WARPS
#define MSHA_DZ_PROD_WARPS 4
#endif
constexpr int DZ_PROD_WARPS = MSHA_DZ_PROD_WARPS;    // dZ kernel: warps 4.. turn the TMA-staged dOut / out tiles into G_hi / G_lo,
constexpr int DZ_PROD_THREADS = DZ_PROD_WARPS * 32;
constexpr int DZ_EPI_WARPS = 12 - DZ_PROD_WARPS;     //            the remaining warps scatter (TMEM lane quarter = warp % 4)
static_assert(DZ_EPI_WARPS == 4 || DZ_EPI_WARPS == 8, "epilogue warps come in groups of four");
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BLOCK_N, int CTAS>
struct SCfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;          // 8 KB per hi / lo tile (this CTA's 128 rows)
    static constexpr int B_ROWS = BLOCK_N / CTAS;                  // a CTA of a pair stages half of the B tile
    static constexpr int B_BYTES = B_ROWS * BLOCK_K * 4;
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES < 8 ? (192 * 1024) / STAGE_BYTES : 8;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = EPI_OFF + MAX_EPI_WARPS * EPI_STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 512;
};

// Activation of one 16-byte chunk (the branch on `act` is warp-uniform).
__device__ __forceinline__ float fast_sigmoid(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return r;
}
__device__ __forceinline__ float act1(float x, int act, float slope) {
    switch (act) {
        case 1: return x > 0.f ? x : expm1f(x);
        case 2: return fmaxf(x, 0.f);
        case 3: return fast_sigmoid(fmaxf(x, 0.f));
        case 4: return x > 0.f ? x : x * slope;
        case 5: return fast_sigmoid(x);
        default: return x;
    }
}
__device__ __forceinline__ void act4(float4& o, int act, float slope) {
    switch (act) {
        case 3:                                                   // the scorer's relu -> sigmoid
            o.x = fast_sigmoid(fmaxf(o.x, 0.f)); o.y = fast_sigmoid(fmaxf(o.y, 0.f));
            o.z = fast_sigmoid(fmaxf(o.z, 0.f)); o.w = fast_sigmoid(fmaxf(o.w, 0.f));
            break;
        case 0: break;
        default:
            o.x = act1(o.x, act, slope); o.y = act1(o.y, act, slope);
            o.z = act1(o.z, act, slope); o.w = act1(o.w, act, slope);
            break;
    }
}

// ------------------------------------------------------------------------------------------------
// Shared skeleton of the forward and dZ kernels.  CTAS = 2 runs them as CTA pairs (2-CTA cluster, tcgen05 cta_group::2):
// the pair computes a 256-row tile, each CTA keeps its own 128 A rows / accumulator rows and only HALF of the B tile
// (W0 slices come from L2 for every tile, so this halves the dominant shared-memory fill traffic per SM).  The even CTA
// issues the MMAs and owns the full_a / full_b / tmem_empty barriers; empty / tmem_full are signalled in both CTAs by a
// multicast commit.
// ------------------------------------------------------------------------------------------------
template <int CTAS>
__device__ __forceinline__ void pair_sync() {
    if constexpr (CTAS == 2) cluster_sync_all();
    else __syncthreads();
}
// one arrival per warp on a barrier of the MMA-issuing CTA (callers order their own writes first)
template <int CTAS>
__device__ __forceinline__ void warp_arrive_leader(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_remote(bar, 0);
        else mbar_arrive(bar);
    }
}
template <int CTAS>
__device__ __forceinline__ void wait_leader_bar(uint32_t bar, uint32_t parity) {
    if constexpr (CTAS == 2) mbar_wait_cluster(bar, parity);
    else mbar_wait(bar, parity);
}
template <int CTAS>
__device__ __forceinline__ void load_b_slice(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if constexpr (CTAS == 2) tma_load_2d_pair(dst, map, bar, c0, c1);
    else tma_load_2d(dst, map, bar, c0, c1);
}
// the three MMAs of the 3xTF32 product for one k-block, then release the stage
template <int BLOCK_N, int CTAS, class S>
__device__ __forceinline__ void issue_kblock(uint8_t* st, uint32_t tmem_d, bool first, uint32_t empty_bar) {
    constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M * CTAS, BLOCK_N, false, false);
    const uint32_t a_hi = smem_u32(st), a_lo = a_hi + S::A_BYTES;
    const uint32_t b_hi = a_hi + 2 * S::A_BYTES, b_lo = b_hi + S::B_BYTES;
#pragma unroll
    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
        const uint64_t dah = make_smem_desc(a_hi + k * 32, 16, 512, 4);
        const uint64_t dal = make_smem_desc(a_lo + k * 32, 16, 512, 4);
        const uint64_t dbh = make_smem_desc(b_hi + k * 32, 16, 512, 4);
        const uint64_t dbl = make_smem_desc(b_lo + k * 32, 16, 512, 4);
        const uint32_t accum = (!first || k > 0) ? 1u : 0u;
        if constexpr (CTAS == 2) {
            umma_tf32_pair(tmem_d, dal, dbh, idesc, accum);
            umma_tf32_pair(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32_pair(tmem_d, dah, dbh, idesc, 1u);
        } else {
            umma_tf32(tmem_d, dal, dbh, idesc, accum);
            umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
        }
    }
    if constexpr (CTAS == 2) tcgen05_commit_pair(empty_bar);
    else tcgen05_commit(empty_bar);
}

// ------------------------------------------------------------------------------------------------
// fused forward
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int CTAS>
__global__ void __launch_bounds__(FWD_THREADS, 1)
score_fwd_kernel(const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                 const float* __restrict__ hi_tab, const float* __restrict__ hj_tab, const int64_t* __restrict__ src,
                 const int64_t* __restrict__ dst, int64_t P, int C, int N, const float* __restrict__ bias, int act,
                 float slope, float* __restrict__ out, int64_t ldo) {
    using S = SCfg<BLOCK_N, CTAS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* full_a = bars;
    uint64_t* full_b = bars + S::STAGES;
    uint64_t* empty = bars + 2 * S::STAGES;
    uint64_t* tmem_full = bars + 4 * S::STAGES;
    uint64_t* tmem_empty = bars + 4 * S::STAGES + 2;
    uint32_t* tmem_ptr = (uint32_t*)(bars + 4 * S::STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if constexpr (CTAS == 2) rank = cluster_ctarank();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBl) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(smem_u32(&full_a[s]), CTAS * FWD_PROD_WARPS);
            mbar_init(smem_u32(&full_b[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tmem_full[a]), 1);
            mbar_init(smem_u32(&tmem_empty[a]), CTAS * FWD_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc<CTAS>(smem_u32(tmem_ptr), (uint32_t)S::TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    pair_sync<CTAS>();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int64_t m_tiles = (P + BLOCK_M - 1) / BLOCK_M;
    const int64_t pair_tiles = (m_tiles + CTAS - 1) / CTAS;
    const int64_t tp0 = blockIdx.x / CTAS, tp_step = gridDim.x / CTAS;
    const int total_kb = (C + BLOCK_K - 1) / BLOCK_K;

    if (warp == 0) {
        // ---------------- TMA producer for this CTA's slice of the (pre-split) weights ----------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    uint8_t* st = smem + stage * S::STAGE_BYTES;
                    const uint32_t b_hi = smem_u32(st + 2 * S::A_BYTES), b_lo = b_hi + S::B_BYTES;
                    const uint32_t bar = smem_u32(&full_b[stage]);
                    if (rank == 0) mbar_arrive_expect_tx(bar, CTAS * 2 * S::B_BYTES);
                    load_b_slice<CTAS>(b_hi, &tmBh, bar, kb * BLOCK_K, (int)rank * S::B_ROWS);
                    load_b_slice<CTAS>(b_lo, &tmBl, bar, kb * BLOCK_K, (int)rank * S::B_ROWS);
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (even CTA of a pair) ----------------
        if (lane == 0 && rank == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
                wait_leader_bar<CTAS>(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(smem_u32(&full_b[stage]), phase);
                    wait_leader_bar<CTAS>(smem_u32(&full_a[stage]), phase);
                    tcgen05_fence_after();
                    issue_kblock<BLOCK_N, CTAS, S>(smem + stage * S::STAGE_BYTES, tmem_d, kb == 0, smem_u32(&empty[stage]));
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
                if constexpr (CTAS == 2) tcgen05_commit_pair(smem_u32(&tmem_full[acc]));
                else tcgen05_commit(smem_u32(&tmem_full[acc]));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 4 + FWD_PROD_WARPS) {
        // ---------------- A producers: gather h_i[src], h_j[dst], multiply, split, store swizzled ----------------
        constexpr int RPT = 4, RSTEP = 32;                       // rows per thread, row stride between them
        const int tid = threadIdx.x - 128;
        const int c = tid & 3, rbase = tid >> 2;                 // 16-byte chunk c of rows rbase + 32*i
        uint32_t stage = 0, phase = 0;
        for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
            const int64_t m0 = (tp * CTAS + rank) * BLOCK_M;
            const float* pa[RPT];
            const float* pb[RPT];
            bool valid[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int64_t p = m0 + rbase + RSTEP * i;
                valid[i] = p < P;
                const int64_t si = valid[i] ? (src ? src[p] : p) : 0;
                const int64_t dj = valid[i] ? (dst ? dst[p] : p) : 0;
                pa[i] = hi_tab + si * C + c * 4;
                pb[i] = hj_tab + dj * C + c * 4;
            }
            // gathers are software-pipelined one k-block ahead (L2 latency ~ one stage of MMA work)
            float4 a[RPT], b[RPT], an[RPT], bn[RPT];
            auto load = [&](int kb, float4 (&x)[RPT], float4 (&y)[RPT]) {
                const int k = kb * BLOCK_K;
                const bool kvalid = (kb < total_kb) && (k + c * 4 < C);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    if (valid[i] && kvalid) {
                        x[i] = ldg4(pa[i] + k);
                        y[i] = ldg4(pb[i] + k);
                    } else {
                        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        y[i] = x[i];
                    }
                }
            };
            load(0, a, b);
            for (int kb = 0; kb < total_kb; ++kb) {
                load(kb + 1, an, bn);
                mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                uint8_t* st = smem + stage * S::STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    float4 h, l;
                    split_tf32(a[i].x * b[i].x, h.x, l.x);
                    split_tf32(a[i].y * b[i].y, h.y, l.y);
                    split_tf32(a[i].z * b[i].z, h.z, l.z);
                    split_tf32(a[i].w * b[i].w, h.w, l.w);
                    const uint32_t off = sw64_offset(rbase + RSTEP * i, c);
                    sts128(smem_u32(st) + off, h);
                    sts128(smem_u32(st) + S::A_BYTES + off, l);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                warp_arrive_leader<CTAS>(smem_u32(&full_a[stage]), lane);
                if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
                for (int i = 0; i < RPT; ++i) { a[i] = an[i]; b[i] = bn[i]; }
            }
        }
    } else if (warp >= 4 + FWD_PROD_WARPS) {
        // ---------------- epilogue ----------------
        const int ew = warp - (4 + FWD_PROD_WARPS);
        const int q = warp & 3, hf = ew >> 2;
        constexpr int COLS_PER_WARP = BLOCK_N / (FWD_EPI_WARPS / 4);
        const uint32_t stage_s = smem_u32(smem + S::EPI_OFF + ew * EPI_STAGE_BYTES);
        const bool vec_ok = ((ldo & 3) == 0) && ((((uintptr_t)out) & 15) == 0);
        uint32_t acc = 0, acc_phase = 0;
        const int rs = lane >> 3, cg = lane & 7;                 // after the transpose: rows 4k + rs, 16-byte column chunk cg
        constexpr int NCH = COLS_PER_WARP / 32;
        float4 bias_r[NCH];                                      // bias of this thread's 4 columns in each chunk of the warp
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
            const int col = hf * COLS_PER_WARP + 32 * u + 4 * cg;
            float bv[4] = {0.f, 0.f, 0.f, 0.f};
            if (bias != nullptr) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (col + e < N) bv[e] = __ldg(bias + col + e);
            }
            bias_r[u] = make_float4(bv[0], bv[1], bv[2], bv[3]);
        }
        for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
            const int64_t row0 = (tp * CTAS + rank) * BLOCK_M + q * 32;
            mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
            tcgen05_fence_after();
#pragma unroll
            for (int u = 0; u < NCH; ++u) {
                const int nb = hf * COLS_PER_WARP + 32 * u;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + nb, v);
                if (nb >= N) continue;
                // transpose through shared memory (16-byte chunks XOR-swizzled by row); bias + activation run on the
                // transposed side, where a thread's four columns (and so its bias) are the same for every row
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    sts128(stage_s + (uint32_t)(lane * 32 + ((g ^ (lane & 7)) << 2)) * 4u,
                           make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                       __uint_as_float(v[4 * g + 3])));
                __syncwarp();
                const float4 bj = bias_r[u];
                const int col = nb + 4 * cg;
                float4 o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = 4 * k + rs;
                    o[k] = lds128(stage_s + (uint32_t)(r * 32 + ((cg ^ (r & 7)) << 2)) * 4u);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    o[k].x += bj.x; o[k].y += bj.y; o[k].z += bj.z; o[k].w += bj.w;
                    act4(o[k], act, slope);
                    const int64_t row = row0 + 4 * k + rs;
                    if (row < P) {
                        float* cp = out + row * ldo + col;
                        if (vec_ok && col + 4 <= N) {
                            *reinterpret_cast<float4*>(cp) = o[k];
                        } else {
                            if (col + 0 < N) cp[0] = o[k].x;
                            if (col + 1 < N) cp[1] = o[k].y;
                            if (col + 2 < N) cp[2] = o[k].z;
                            if (col + 3 < N) cp[3] = o[k].w;
                        }
                    }
                }
                __syncwarp();
            }
            tcgen05_fence_before();
            warp_arrive_leader<CTAS>(smem_u32(&tmem_empty[acc]), lane);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    pair_sync<CTAS>();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<CTAS>(tmem_base, (uint32_t)S::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// fused backward, part 1:  G = dOut * act'(out)  (+ bias gradient),  dZ = G @ W0,  scatter into dh_i / dh_j
//   A operand: dOut / out tiles arrive by TMA and are turned into G_hi / G_lo in place (K-major, K = hidden)
//   B operand: W0^T [C, Hd] pre-split hi/lo by TMA (K-major), one slice per CTA of the pair
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_bwd_dz_kernel(const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                    const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmY, int act, float slope,
                    float* __restrict__ G, float* __restrict__ db, const float* __restrict__ hi_tab,
                    const float* __restrict__ hj_tab, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                    int64_t P, int K /*hidden*/, int N /*C*/, float* __restrict__ dhi, float* __restrict__ dhj,
                    const int64_t* __restrict__ nll_target, const float* __restrict__ nll_gout) {
    // nll_target != NULL: dOut is not read -- it is -gout/P in column target[p] of row p and 0 elsewhere (nll read-out)
    using S = SCfg<BLOCK_N, CTAS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* full_a = bars;
    uint64_t* full_b = bars + S::STAGES;
    uint64_t* empty = bars + 2 * S::STAGES;
    uint64_t* full_raw = bars + 3 * S::STAGES;               // dOut / out tiles of this CTA have landed (local)
    uint64_t* tmem_full = bars + 4 * S::STAGES;
    uint64_t* tmem_empty = bars + 4 * S::STAGES + 2;
    uint32_t* tmem_ptr = (uint32_t*)(bars + 4 * S::STAGES + 4);
    __shared__ float db_sm[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if constexpr (CTAS == 2) rank = cluster_ctarank();

    if (threadIdx.x < 256) db_sm[threadIdx.x] = 0.f;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmBl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmD) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmY) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S::STAGES; ++s) {
            mbar_init(smem_u32(&full_a[s]), CTAS * DZ_PROD_WARPS);
            mbar_init(smem_u32(&full_b[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
            mbar_init(smem_u32(&full_raw[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tmem_full[a]), 1);
            mbar_init(smem_u32(&tmem_empty[a]), CTAS * DZ_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc<CTAS>(smem_u32(tmem_ptr), (uint32_t)S::TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    pair_sync<CTAS>();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int64_t m_tiles = (P + BLOCK_M - 1) / BLOCK_M;
    const int64_t pair_tiles = (m_tiles + CTAS - 1) / CTAS;
    const int64_t tp0 = blockIdx.x / CTAS, tp_step = gridDim.x / CTAS;
    const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
                const int64_t m0 = (tp * CTAS + rank) * BLOCK_M;
                const int row_c = (int)(m0 < P ? m0 : P);                // a tile past the end reads zeros
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                    uint8_t* st = smem + stage * S::STAGE_BYTES;
                    const uint32_t a_d = smem_u32(st), a_y = a_d + S::A_BYTES;
                    const uint32_t b_hi = smem_u32(st + 2 * S::A_BYTES), b_lo = b_hi + S::B_BYTES;
                    const uint32_t raw = smem_u32(&full_raw[stage]);
                    mbar_arrive_expect_tx(raw, (nll_target ? 1 : 2) * S::A_BYTES);
                    if (!nll_target) tma_load_2d(a_d, &tmD, raw, kb * BLOCK_K, row_c);   // dOut tile -> becomes G_hi in place
                    tma_load_2d(a_y, &tmY, raw, kb * BLOCK_K, row_c);           // out  tile -> becomes G_lo in place
                    const uint32_t bar = smem_u32(&full_b[stage]);
                    if (rank == 0) mbar_arrive_expect_tx(bar, CTAS * 2 * S::B_BYTES);
                    load_b_slice<CTAS>(b_hi, &tmBh, bar, kb * BLOCK_K, (int)rank * S::B_ROWS);
                    load_b_slice<CTAS>(b_lo, &tmBl, bar, kb * BLOCK_K, (int)rank * S::B_ROWS);
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
                wait_leader_bar<CTAS>(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(smem_u32(&full_b[stage]), phase);
                    wait_leader_bar<CTAS>(smem_u32(&full_a[stage]), phase);
                    tcgen05_fence_after();
                    issue_kblock<BLOCK_N, CTAS, S>(smem + stage * S::STAGE_BYTES, tmem_d, kb == 0, smem_u32(&empty[stage]));
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
                if constexpr (CTAS == 2) tcgen05_commit_pair(smem_u32(&tmem_full[acc]));
                else tcgen05_commit(smem_u32(&tmem_full[acc]));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 4 + DZ_PROD_WARPS) {
        // ---------------- producers: the TMA-staged dOut / out tiles become G_hi / G_lo in place; G also goes to HBM
        // (for the dW kernel) and its column sums (bias gradient) accumulate in registers ----
        constexpr int RSTEP = DZ_PROD_THREADS / 4, RPT = BLOCK_M / RSTEP;
        const int tid = threadIdx.x - 128;
        const int c = tid & 3, rbase = tid >> 2;                 // 16-byte chunk c of rows rbase + RSTEP * i
        uint32_t stage = 0, phase = 0;
        // act'(x) from the output y, branch free:  y > thr ? a0 + y*(a1 + a2*y) : b0 + b1*y
        float thr = -INFINITY, a0 = 1.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f;
        switch (act) {
            case 1: thr = 0.f; b0 = 1.f; b1 = 1.f; break;                       // ELU: 1 | y + 1
            case 2: thr = 0.f; break;                                            // ReLU: 1 | 0
            case 3: thr = 0.5f; a0 = 0.f; a1 = 1.f; a2 = -1.f; break;            // sigmoid(relu): y(1-y) | 0
            case 4: thr = 0.f; b0 = slope; break;                                // LeakyReLU: 1 | slope
            case 5: a0 = 0.f; a1 = 1.f; a2 = -1.f; break;                        // sigmoid: y(1-y)
            default: break;
        }
        auto dact = [&](float yv) { return yv > thr ? fmaf(yv, fmaf(a2, yv, a1), a0) : fmaf(b1, yv, b0); };
        const float nll_g = nll_target ? -__ldg(nll_gout) / (float)P : 0.f;
        float4 csum[16];                                         // bias-gradient partials per k-block (K <= 256)
#pragma unroll
        for (int j = 0; j < 16; ++j) csum[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
            const int64_t m0 = (tp * CTAS + rank) * BLOCK_M;
            uint32_t off[RPT];
            bool pv[RPT];
            int lab[RPT];                                        // nll read-out: the one column of the row that carries a gradient
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                off[i] = sw64_offset(rbase + RSTEP * i, c);
                pv[i] = m0 + rbase + RSTEP * i < P;
                lab[i] = -1;
                if (nll_target && pv[i]) {
                    const int64_t tg = nll_target[m0 + rbase + RSTEP * i];
                    lab[i] = (tg >= 0 && tg < K) ? (int)tg : -1;
                }
            }
#pragma unroll
            for (int kb = 0; kb < 16; ++kb) {
                if (kb < total_kb) {
                    const int k = kb * BLOCK_K + c * 4;
                    const bool kvalid = k < K;
                    mbar_wait(smem_u32(&full_raw[stage]), phase);        // dOut / out tiles of this stage have landed
                    uint8_t* st = smem + stage * S::STAGE_BYTES;
                    float4 d[RPT], y[RPT];
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        if (nll_target) {
                            const int dk = lab[i] - k;               // 0..3: the labelled column lies in this 16-byte chunk
                            d[i] = make_float4(dk == 0 ? nll_g : 0.f, dk == 1 ? nll_g : 0.f, dk == 2 ? nll_g : 0.f,
                                               dk == 3 ? nll_g : 0.f);
                        } else {
                            d[i] = lds128(smem_u32(st) + off[i]);
                        }
                        y[i] = lds128(smem_u32(st) + S::A_BYTES + off[i]);
                    }
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {                      // rows beyond P / columns beyond K arrive as zeros
                        float4 g, h, l;
                        g.x = d[i].x * dact(y[i].x);
                        g.y = d[i].y * dact(y[i].y);
                        g.z = d[i].z * dact(y[i].z);
                        g.w = d[i].w * dact(y[i].w);
                        csum[kb].x += g.x; csum[kb].y += g.y; csum[kb].z += g.z; csum[kb].w += g.w;
#ifndef MSHA_ABL_NO_GSTORE
                        if (pv[i] && kvalid) *reinterpret_cast<float4*>(G + (m0 + rbase + RSTEP * i) * K + k) = g;
#endif
                        split_tf32(g.x, h.x, l.x);
                        split_tf32(g.y, h.y, l.y);
                        split_tf32(g.z, h.z, l.z);
                        split_tf32(g.w, h.w, l.w);
                        sts128(smem_u32(st) + off[i], h);                               // in place: G_hi over the dOut tile
                        sts128(smem_u32(st) + S::A_BYTES + off[i], l);                  //           G_lo over the out tile
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    warp_arrive_leader<CTAS>(smem_u32(&full_a[stage]), lane);
                    if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        // bias gradient: reduce the register partials over the lanes sharing chunk c, then over warps / CTAs
#pragma unroll
        for (int kb = 0; kb < 16; ++kb) {
            if (kb < total_kb) {
                float4 v = csum[kb];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
                    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
                    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
                }
                const int k = kb * BLOCK_K + c * 4;
                if (lane < 4 && k < K) {
                    atomicAdd(&db_sm[k + 0], v.x);
                    atomicAdd(&db_sm[k + 1], v.y);
                    atomicAdd(&db_sm[k + 2], v.z);
                    atomicAdd(&db_sm[k + 3], v.w);
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(DZ_PROD_THREADS) : "memory");
        if (db != nullptr)
            for (int k = tid; k < K; k += DZ_PROD_THREADS) atomicAdd(db + k, db_sm[k]);
    } else if (warp >= 4 + DZ_PROD_WARPS) {
        // ---------------- epilogue: dZ tile -> dh_i[src] += dZ * h_j[dst],  dh_j[dst] += dZ * h_i[src] ----------------
        // The gathers of h_i / h_j do not depend on the accumulator, so they run one batch (4 row groups x 2 tables =
        // 8 x 16 B per lane) ahead of the multiply + atomics, ping-ponging between two register buffers.
        const int ew = warp - (4 + DZ_PROD_WARPS);
        const int q = warp & 3, hf = ew >> 2;
        constexpr int COLS_PER_WARP = BLOCK_N / (DZ_EPI_WARPS / 4);
        const uint32_t stage_s = smem_u32(smem + S::EPI_OFF + ew * EPI_STAGE_BYTES);
        const int rs = lane >> 3, cg = lane & 7;
        uint32_t acc = 0, acc_phase = 0;
        for (int64_t tp = tp0; tp < pair_tiles; tp += tp_step) {
            const int64_t row0 = (tp * CTAS + rank) * BLOCK_M + q * 32;
            const int64_t prow = row0 + lane;
            int s_l = 0, d_l = 0;
            if (prow < P) {
                s_l = (int)(src ? src[prow] : prow);
                d_l = (int)(dst ? dst[prow] : prow);
            }
            auto gather = [&](int nb, int kh, float4 (&xi)[4], float4 (&xj)[4]) {
                const int col = nb + 4 * cg;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int r = 4 * (4 * kh + kk) + rs;
                    const int si = __shfl_sync(0xffffffffu, s_l, r), dj = __shfl_sync(0xffffffffu, d_l, r);
                    if ((row0 + r < P) && (col < N)) {                 // N % 4 == 0
#ifndef MSHA_ABL_NO_GATHER
                        xj[kk] = ldg4(hj_tab + (int64_t)dj * N + col);
                        xi[kk] = ldg4(hi_tab + (int64_t)si * N + col);
#else
                        xj[kk] = make_float4(1.f, 1.f, 1.f, (float)dj); xi[kk] = make_float4(1.f, 1.f, 1.f, (float)si);
#endif
                    }
                }
            };
            auto scatter = [&](int nb, int kh, const float4 (&xi)[4], const float4 (&xj)[4]) {
                const int col = nb + 4 * cg;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int r = 4 * (4 * kh + kk) + rs;
                    const int si = __shfl_sync(0xffffffffu, s_l, r), dj = __shfl_sync(0xffffffffu, d_l, r);
                    if ((row0 + r < P) && (col < N)) {
                        const float4 dz = lds128(stage_s + (uint32_t)(r * 32 + ((cg ^ (r & 7)) << 2)) * 4u);
#ifndef MSHA_ABL_NO_ATOMICS
                        atomicAdd(reinterpret_cast<float4*>(dhi + (int64_t)si * N + col),
                                  make_float4(dz.x * xj[kk].x, dz.y * xj[kk].y, dz.z * xj[kk].z, dz.w * xj[kk].w));
                        atomicAdd(reinterpret_cast<float4*>(dhj + (int64_t)dj * N + col),
                                  make_float4(dz.x * xi[kk].x, dz.y * xi[kk].y, dz.z * xi[kk].z, dz.w * xi[kk].w));
#else
                        if (dz.x * xj[kk].x + dz.y * xi[kk].w == 123.456f) dhi[0] = 1.f;
#endif
                    }
                }
            };
            float4 xi0[4], xj0[4], xi1[4], xj1[4];
            gather(hf * COLS_PER_WARP, 0, xi0, xj0);             // in flight while the accumulator finishes
            mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
            tcgen05_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
                const int nb = hf * COLS_PER_WARP + cc;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + nb, v);
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    sts128(stage_s + (uint32_t)(lane * 32 + ((g ^ (lane & 7)) << 2)) * 4u,
                           make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                       __uint_as_float(v[4 * g + 3])));
                __syncwarp();
                gather(nb, 1, xi1, xj1);
                scatter(nb, 0, xi0, xj0);
                if (cc + 32 < COLS_PER_WARP) gather(nb + 32, 0, xi0, xj0);
                scatter(nb, 1, xi1, xj1);
                __syncwarp();
            }
            tcgen05_fence_before();
            warp_arrive_leader<CTAS>(smem_u32(&tmem_empty[acc]), lane);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    pair_sync<CTAS>();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<CTAS>(tmem_base, (uint32_t)S::TMEM_COLS);
    }
}


// ------------------------------------------------------------------------------------------------
// fused backward, part 2:  dW0[Hd, C] = G^T @ Z,  Z[p,:] = h_i[src[p]] * h_j[dst[p]] regenerated by gather warps.
//   A operand: G [K = pairs, M = Hd]  MN-major, TMA (128B-atom-32B swizzle) + converter warps (hi / lo)
//   B operand: Z [K = pairs, N = C]   MN-major, written hi / lo by the gather warps in the same swizzle
//   A CTA (CTAS = 1) or CTA pair (CTAS = 2) owns a contiguous range of pairs and keeps the whole 256 x 256 accumulator
//   in TMEM; partial results are added to dW0 with atomics at the end.  In a pair each CTA stages 128 of the Hd rows
//   of G and 128 of the C columns of Z -- half the shared-memory traffic per SM for the same tensor work.
// ------------------------------------------------------------------------------------------------
template <int CTAS>
struct WCfg {
    static constexpr int ROWS = 256 / CTAS;                         // Hd rows of G / C columns of Z staged per CTA
    static constexpr int BK = 16 * CTAS;                            // pairs per stage (a pair needs 32 to amortise its hand-offs)
    static constexpr int CHUNK_BYTES = BK * 128;                    // one 32-wide MN chunk: BK K-rows of 128 B
    static constexpr int TILE_BYTES = ROWS * BK * 4;                // 16 KB per hi / lo tile
    static constexpr int STAGE_BYTES = 4 * TILE_BYTES;              // A_hi, A_lo, B_hi, B_lo
    static constexpr int STAGES = 3;
    static constexpr int TMEM_COLS = 512 / CTAS;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 512;
};

template <int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_bwd_dw_kernel(const __grid_constant__ CUtensorMap tmG, const float* __restrict__ hi_tab,
                    const float* __restrict__ hj_tab, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                    int64_t P, int Hd, int C, float* __restrict__ dW) {
    using W = WCfg<CTAS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + W::BAR_OFF);
    uint64_t* full_raw = bars;
    uint64_t* full_cvt = bars + W::STAGES;
    uint64_t* full_b = bars + 2 * W::STAGES;
    uint64_t* empty = bars + 3 * W::STAGES;
    uint64_t* acc_full = bars + 4 * W::STAGES;
    uint32_t* tmem_ptr = (uint32_t*)(bars + 4 * W::STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if constexpr (CTAS == 2) rank = cluster_ctarank();

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmG) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < W::STAGES; ++s) {
            mbar_init(smem_u32(&full_raw[s]), 1);
            mbar_init(smem_u32(&full_cvt[s]), CTAS * (CVT_THREADS / 32));
            mbar_init(smem_u32(&full_b[s]), CTAS * (PROD_THREADS / 32));
            mbar_init(smem_u32(&empty[s]), 1);
        }
        mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc<CTAS>(smem_u32(tmem_ptr), (uint32_t)W::TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    pair_sync<CTAS>();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // contiguous range of BK-pair chunks owned by this CTA (pair)
    const int64_t n_chunks = (P + W::BK - 1) / W::BK;
    const int64_t n_owner = gridDim.x / CTAS;
    const int64_t per = (n_chunks + n_owner - 1) / n_owner;
    const int64_t c0 = (int64_t)(blockIdx.x / CTAS) * per;
    const int64_t c1 = c0 + per < n_chunks ? c0 + per : n_chunks;
    const int m_tiles = CTAS == 2 ? 1 : (Hd + 127) / 128;
    const int a_chunks = m_tiles * 4;                     // 32-wide MN chunks of G actually loaded
    const int row_base = (int)rank * W::ROWS;             // first Hd row (A) / C column (B) staged by this CTA

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t ch = c0; ch < c1; ++ch) {
                mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
                const uint32_t a_dst = smem_u32(smem + stage * W::STAGE_BYTES);
                const uint32_t bar = smem_u32(&full_raw[stage]);
                mbar_arrive_expect_tx(bar, (uint32_t)a_chunks * W::CHUNK_BYTES);
                for (int i = 0; i < a_chunks; ++i)
                    tma_load_2d(a_dst + i * W::CHUNK_BYTES, &tmG, bar, row_base + 32 * i, (int)(ch * W::BK));
                if (++stage == W::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(128 * CTAS, 256, true, true);
            uint32_t stage = 0, phase = 0;
            for (int64_t ch = c0; ch < c1; ++ch) {
                wait_leader_bar<CTAS>(smem_u32(&full_cvt[stage]), phase);
                wait_leader_bar<CTAS>(smem_u32(&full_b[stage]), phase);
                tcgen05_fence_after();
                uint8_t* st = smem + stage * W::STAGE_BYTES;
                const uint32_t a_hi = smem_u32(st), a_lo = a_hi + W::TILE_BYTES;
                const uint32_t b_hi = a_hi + 2 * W::TILE_BYTES, b_lo = b_hi + W::TILE_BYTES;
                for (int mt = 0; mt < m_tiles; ++mt) {
                    const uint32_t tmem_d = tmem_base + mt * 256;
#pragma unroll
                    for (int k = 0; k < W::BK / UMMA_K; ++k) {
                        const uint64_t dah = make_smem_desc(a_hi + mt * 4 * W::CHUNK_BYTES + k * 1024, W::CHUNK_BYTES, 512, 1);
                        const uint64_t dal = make_smem_desc(a_lo + mt * 4 * W::CHUNK_BYTES + k * 1024, W::CHUNK_BYTES, 512, 1);
                        const uint64_t dbh = make_smem_desc(b_hi + k * 1024, W::CHUNK_BYTES, 512, 1);
                        const uint64_t dbl = make_smem_desc(b_lo + k * 1024, W::CHUNK_BYTES, 512, 1);
                        const uint32_t accum = (ch > c0 || k > 0) ? 1u : 0u;
                        if constexpr (CTAS == 2) {
                            umma_tf32_pair(tmem_d, dal, dbh, idesc, accum);
                            umma_tf32_pair(tmem_d, dah, dbl, idesc, 1u);
                            umma_tf32_pair(tmem_d, dah, dbh, idesc, 1u);
                        } else {
                            umma_tf32(tmem_d, dal, dbh, idesc, accum);
                            umma_tf32(tmem_d, dah, dbl, idesc, 1u);
                            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
                        }
                    }
                }
                if constexpr (CTAS == 2) tcgen05_commit_pair(smem_u32(&empty[stage]));
                else tcgen05_commit(smem_u32(&empty[stage]));
                if (++stage == W::STAGES) { stage = 0; phase ^= 1; }
            }
            if constexpr (CTAS == 2) tcgen05_commit_pair(smem_u32(acc_full));
            else tcgen05_commit(smem_u32(acc_full));
        }
    } else if (warp >= 4 && warp < 8) {
        // ---------------- converters for the TMA-loaded G tile ----------------
        const int tid = threadIdx.x - 128;
        uint32_t stage = 0, phase = 0;
        for (int64_t ch = c0; ch < c1; ++ch) {
            mbar_wait(smem_u32(&full_raw[stage]), phase);
            uint8_t* st = smem + stage * W::STAGE_BYTES;
            split_tile_inplace(smem_u32(st), smem_u32(st) + W::TILE_BYTES, a_chunks * W::CHUNK_BYTES / 16, tid, CVT_THREADS);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            warp_arrive_leader<CTAS>(smem_u32(&full_cvt[stage]), lane);
            if (++stage == W::STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 8) {
        // ---------------- Z producers (8 warps): BK pairs x this CTA's channels per stage, MN-major 128B-atom-32B swizzle ----
        constexpr int TPP = PROD_THREADS / W::BK;                // threads per pair (16 / 8)
        constexpr int NI = W::ROWS / (4 * TPP);                  // 16-byte chunks per thread and table (4)
        const int tid = threadIdx.x - 256;
        const int kk = tid / TPP, j = tid % TPP;                 // pair kk of the chunk; 16-byte column j (+TPP per step)
        uint32_t stage = 0, phase = 0;
        float4 za[NI], zb[NI], na[NI], nb4[NI];
        auto load = [&](int64_t ch, float4 (&x)[NI], float4 (&y)[NI]) {
            const int64_t p = ch * W::BK + kk;
            const bool pvalid = (ch < c1) && (p < P);
            const int64_t si = pvalid ? (src ? src[p] : p) : 0;
            const int64_t dj = pvalid ? (dst ? dst[p] : p) : 0;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int col = row_base + j * 4 + 4 * TPP * i;
                if (pvalid && col < C) {
                    x[i] = ldg4(hi_tab + si * C + col);
                    y[i] = ldg4(hj_tab + dj * C + col);
                } else {
                    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    y[i] = x[i];
                }
            }
        };
        if (c0 < c1) load(c0, za, zb);
        for (int64_t ch = c0; ch < c1; ++ch) {
            load(ch + 1, na, nb4);
            mbar_wait(smem_u32(&empty[stage]), phase ^ 1);
            const uint32_t st = smem_u32(smem + stage * W::STAGE_BYTES + 2 * W::TILE_BYTES);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                float4 h, l;
                split_tf32(za[i].x * zb[i].x, h.x, l.x);
                split_tf32(za[i].y * zb[i].y, h.y, l.y);
                split_tf32(za[i].z * zb[i].z, h.z, l.z);
                split_tf32(za[i].w * zb[i].w, h.w, l.w);
                const int col = j * 4 + 4 * TPP * i;             // local channel of this 16-byte chunk
                const uint32_t off = (col >> 5) * W::CHUNK_BYTES + sw128b32_offset(kk, (col & 31) >> 2);
                sts128(st + off, h);
                sts128(st + W::TILE_BYTES + off, l);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            warp_arrive_leader<CTAS>(smem_u32(&full_b[stage]), lane);
            if (++stage == W::STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
            for (int i = 0; i < NI; ++i) { za[i] = na[i]; zb[i] = nb4[i]; }
        }
    }
    // ---------------- epilogue (warps 8-15): TMEM -> atomicAdd into dW0 ----------------
    if (warp >= 8 && c1 > c0) {
        const int q = warp & 3, hf = (warp - 8) >> 2;
        mbar_wait(smem_u32(acc_full), 0);
        tcgen05_fence_after();
        for (int mt = 0; mt < m_tiles; ++mt) {
            const int row = (CTAS == 2 ? (int)rank * 128 : mt * 128) + q * 32 + lane;
#pragma unroll 1
            for (int cc = 0; cc < 128; cc += 32) {
                const int nb = hf * 128 + cc;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mt * 256 + nb, v);
                if (row < Hd && nb < C) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < C) atomicAdd(dW + (int64_t)row * C + nb + j, __uint_as_float(v[j]));
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    pair_sync<CTAS>();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<CTAS>(tmem_base, (uint32_t)W::TMEM_COLS);
    }
}

__global__ void split_weights_t_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int Hd,
                                       int C) {
    // w [Hd, C] -> hi/lo [C, Hd] (transposed)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Hd * C) return;
    const int c = (int)(i / Hd), h = (int)(i % Hd);
    float hh, ll;
    split_tf32(w[(int64_t)h * C + c], hh, ll);
    hi[i] = hh;
    lo[i] = ll;
}

__global__ void split_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h, l;
    split_tf32(w[i], h, l);
    hi[i] = h;
    lo[i] = l;
}

#ifndef MSHA_SCORE_CTAS
#define MSHA_SCORE_CTAS 2
#endif
constexpr int SCORE_CTAS = MSHA_SCORE_CTAS;   // 2: forward / dZ kernels run as CTA pairs (cta_group::2); 1: single CTAs

// persistent launch: one CTA (or CTA pair) per SM, tiles strided over the grid
template <class Kern, class... Args>
int launch_tiles(Kern kern, int threads, int smem_bytes, int64_t m_tiles, cudaStream_t st, Args... args) {
    const int64_t pair_tiles = (m_tiles + SCORE_CTAS - 1) / SCORE_CTAS;
    const int max_pairs = MSHA_NUM_SMS / SCORE_CTAS;
    const int grid = (int)(pair_tiles < max_pairs ? pair_tiles : max_pairs) * SCORE_CTAS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SCORE_CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MSHA_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    MSHA_LAUNCH_OK();
    return 0;
}

template <int BN>
int launch_fwd(const CUtensorMap& tbh, const CUtensorMap& tbl, const float* hi_tab, const float* hj_tab, const int64_t* src,
               const int64_t* dst, int64_t P, int C, int N, const float* bias, int act, float slope, float* out, int64_t ldo,
               cudaStream_t st) {
    using S = SCfg<BN, SCORE_CTAS>;
    auto kern = score_fwd_kernel<BN, SCORE_CTAS>;
    static MshaPerDeviceOnce attr_set;
    if (attr_set.need()) {
        MSHA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
        attr_set.mark();
    }
    return launch_tiles(kern, FWD_THREADS, S::SMEM_BYTES, (P + BLOCK_M - 1) / BLOCK_M, st, tbh, tbl, hi_tab, hj_tab, src, dst, P, C, N, bias,
                        act, slope, out, ldo);
}

}  // namespace

// 0 when the fused scorer kernels apply: C % 4 == 0, hidden <= 256, 16-byte aligned tables
MSHA_API int msha_score_mlp_supported(const float* hi_tab, const float* hj_tab, const float* W0, int64_t C, int64_t Hd) {
    if (C < 4 || (C & 3) || Hd < 1 || Hd > 256 || C >= ((int64_t)1 << 20)) return -1;
    if (((uintptr_t)hi_tab & 15) || ((uintptr_t)hj_tab & 15) || ((uintptr_t)W0 & 15)) return -1;
    return 0;
}

// workspace: hi/lo copies of W0 [Hd, C] (and of W0^T for the backward)
MSHA_API size_t msha_score_mlp_workspace_bytes(int64_t C, int64_t Hd) { return (size_t)4 * C * Hd * sizeof(float) + 1024; }

// out[p, :Hd] = act( (hi_tab[src[p]] * hj_tab[dst[p]]) @ W0^T + b0 );  W0 is [Hd, C] (nn.Linear layout); src/dst may be
// NULL (identity).  Replaces x_i*x_j -> lin -> relu -> sigmoid, LLP.py:105-115.
MSHA_API int msha_score_mlp_fwd(const float* hi_tab, const float* hj_tab, const int64_t* src, const int64_t* dst, int64_t P,
                                int64_t C, const float* W0, const float* b0, int64_t Hd, int act, float slope, float* out,
                                int64_t ldo, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(msha_score_mlp_supported(hi_tab, hj_tab, W0, C, Hd) == 0, "score_mlp_fwd: unsupported shape/alignment");
    MSHA_REQUIRE(ws_bytes >= msha_score_mlp_workspace_bytes(C, Hd), "score_mlp_fwd: workspace too small");
    MSHA_REQUIRE(P >= 0 && ldo >= Hd, "score_mlp_fwd: bad shape");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    float* w_hi = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    float* w_lo = w_hi + C * Hd;
    split_weights_kernel<<<(unsigned)msha_cdiv(C * Hd, 256), 256, 0, st>>>(W0, w_hi, w_lo, C * Hd);
    MSHA_LAUNCH_OK();
    const int BN = Hd > 128 ? 256 : (Hd > 64 ? 128 : 64);
    CUtensorMap tbh, tbl;
    int rc = tc_make_map(&tbh, w_hi, C, Hd, C, BLOCK_K, BN / SCORE_CTAS, false);   // one box per CTA of a pair
    if (rc) return rc;
    rc = tc_make_map(&tbl, w_lo, C, Hd, C, BLOCK_K, BN / SCORE_CTAS, false);
    if (rc) return rc;
    if (BN == 256) return launch_fwd<256>(tbh, tbl, hi_tab, hj_tab, src, dst, P, (int)C, (int)Hd, b0, act, slope, out, ldo, st);
    if (BN == 128) return launch_fwd<128>(tbh, tbl, hi_tab, hj_tab, src, dst, P, (int)C, (int)Hd, b0, act, slope, out, ldo, st);
    return launch_fwd<64>(tbh, tbl, hi_tab, hj_tab, src, dst, P, (int)C, (int)Hd, b0, act, slope, out, ldo, st);
}

namespace {
template <int BN>
int launch_dz(const CUtensorMap& tbh, const CUtensorMap& tbl, const float* dout, const float* outp, int act, float slope,
              float* G, float* db, const float* hi_tab, const float* hj_tab, const int64_t* src, const int64_t* dst, int64_t P,
              int Hd, int C, float* dhi, float* dhj, const int64_t* nll_target, const float* nll_gout, cudaStream_t st) {
    CUtensorMap td, ty;                                   // dOut / out as K-major [P, Hd] operands, boxes {16, 128}
    int rcm = tc_make_map(&td, dout ? dout : outp, Hd, P, Hd, BLOCK_K, BLOCK_M, false);   // unused in the nll variant
    if (rcm) return rcm;
    rcm = tc_make_map(&ty, outp, Hd, P, Hd, BLOCK_K, BLOCK_M, false);
    if (rcm) return rcm;
    using S = SCfg<BN, SCORE_CTAS>;
    auto kern = score_bwd_dz_kernel<BN, SCORE_CTAS>;
    static MshaPerDeviceOnce attr_set;
    if (attr_set.need()) {
        MSHA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
        attr_set.mark();
    }
    return launch_tiles(kern, NUM_THREADS, S::SMEM_BYTES, (P + BLOCK_M - 1) / BLOCK_M, st, tbh, tbl, td, ty, act, slope, G, db, hi_tab,
                        hj_tab, src, dst, P, Hd, C, dhi, dhj, nll_target, nll_gout);
}
}  // namespace

// Backward of msha_score_mlp_fwd.  dout/out: [P, Hd] contiguous.  G: [P, Hd] scratch (receives dOut*act'(out)).
// dhi/dhj [n, C] must hold the running gradients (atomically accumulated); dW0 [Hd, C] and db0 [Hd] are overwritten.
// Needs Hd % 4 == 0 in addition to msha_score_mlp_supported.
static int score_mlp_bwd_impl(const float* dout, const int64_t* nll_target, const float* nll_gout, const float* out,
                              const float* hi_tab, const float* hj_tab, const int64_t* src, const int64_t* dst, int64_t P,
                              int64_t C, const float* W0, int64_t Hd, int act, float slope, float* G, float* dhi, float* dhj,
                              float* dW0, float* db0, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(msha_score_mlp_supported(hi_tab, hj_tab, W0, C, Hd) == 0 && (Hd & 3) == 0 && C <= 256,
                 "score_mlp_bwd: unsupported shape/alignment (need C <= 256, Hd <= 256, both multiples of 4)");
    MSHA_REQUIRE(ws_bytes >= msha_score_mlp_workspace_bytes(C, Hd), "score_mlp_bwd: workspace too small");
    MSHA_REQUIRE(((uintptr_t)dout & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)G & 15) == 0,
                 "score_mlp_bwd: dout/out/G must be 16-byte aligned");
    MSHA_REQUIRE((dout != nullptr) != (nll_target != nullptr), "score_mlp_bwd: exactly one of dout / nll target");
    MSHA_REQUIRE(nll_target == nullptr || nll_gout != nullptr, "score_mlp_nll_bwd: gout required");
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    float* wt_hi = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255) + 2 * C * Hd;
    float* wt_lo = wt_hi + C * Hd;
    split_weights_t_kernel<<<(unsigned)msha_cdiv(C * Hd, 256), 256, 0, st>>>(W0, wt_hi, wt_lo, (int)Hd, (int)C);
    MSHA_LAUNCH_OK();
    MSHA_CUDA(cudaMemsetAsync(db0, 0, (size_t)Hd * sizeof(float), st));
    MSHA_CUDA(cudaMemsetAsync(dW0, 0, (size_t)Hd * C * sizeof(float), st));
    // ---- part 1: G, db0, dZ = G @ W0, scatter
    const int BN = C > 128 ? 256 : (C > 64 ? 128 : 64);
    CUtensorMap tbh, tbl, tg;
    int rc = tc_make_map(&tbh, wt_hi, Hd, C, Hd, BLOCK_K, BN / SCORE_CTAS, false);     // B operand: [N = C rows, K = Hd] K-major
    if (rc) return rc;
    rc = tc_make_map(&tbl, wt_lo, Hd, C, Hd, BLOCK_K, BN / SCORE_CTAS, false);
    if (rc) return rc;
    if (BN == 256) rc = launch_dz<256>(tbh, tbl, dout, out, act, slope, G, db0, hi_tab, hj_tab, src, dst, P, (int)Hd, (int)C, dhi, dhj, nll_target, nll_gout, st);
    else if (BN == 128) rc = launch_dz<128>(tbh, tbl, dout, out, act, slope, G, db0, hi_tab, hj_tab, src, dst, P, (int)Hd, (int)C, dhi, dhj, nll_target, nll_gout, st);
    else rc = launch_dz<64>(tbh, tbl, dout, out, act, slope, G, db0, hi_tab, hj_tab, src, dst, P, (int)Hd, (int)C, dhi, dhj, nll_target, nll_gout, st);
    if (rc) return rc;
    // ---- part 2: dW0 = G^T @ Z
    const bool dw_pairs = SCORE_CTAS == 2 && Hd > 128;   // CTA pairs split the Hd rows of G: pointless for Hd <= 128
    const int dw_bk = dw_pairs ? WCfg<2>::BK : WCfg<1>::BK;
    rc = tc_make_map(&tg, G, Hd, P, Hd, 32, dw_bk, true);                 // A operand: G stored [K = P, M = Hd] MN-major
    if (rc) return rc;
    const int64_t n_chunks = (P + dw_bk - 1) / dw_bk;
    if (dw_pairs) {
        auto kern = score_bwd_dw_kernel<2>;
        static MshaPerDeviceOnce attr_set2;
        if (attr_set2.need()) {
            MSHA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<2>::SMEM_BYTES));
            attr_set2.mark();
        }
        const int max_pairs = MSHA_NUM_SMS / 2;
        const int n_pairs = (int)(n_chunks < max_pairs ? n_chunks : max_pairs);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_pairs * 2);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = (size_t)WCfg<2>::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MSHA_CUDA(cudaLaunchKernelEx(&cfg, kern, tg, hi_tab, hj_tab, src, dst, P, (int)Hd, (int)C, dW0));
        MSHA_LAUNCH_OK();
        return 0;
    }
    static MshaPerDeviceOnce attr_set;
    if (attr_set.need()) {
        MSHA_CUDA(cudaFuncSetAttribute(score_bwd_dw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<1>::SMEM_BYTES));
        attr_set.mark();
    }
    const int grid = (int)(n_chunks < MSHA_NUM_SMS ? n_chunks : MSHA_NUM_SMS);
    score_bwd_dw_kernel<1><<<grid, NUM_THREADS, WCfg<1>::SMEM_BYTES, st>>>(tg, hi_tab, hj_tab, src, dst, P, (int)Hd, (int)C, dW0);
    MSHA_LAUNCH_OK();
    return 0;
}

MSHA_API int msha_score_mlp_bwd(const float* dout, const float* out, const float* hi_tab, const float* hj_tab,
                                const int64_t* src, const int64_t* dst, int64_t P, int64_t C, const float* W0, int64_t Hd,
                                int act, float slope, float* G, float* dhi, float* dhj, float* dW0, float* db0, void* ws,
                                size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(dout != nullptr, "score_mlp_bwd: dout is NULL");
    return score_mlp_bwd_impl(dout, nullptr, nullptr, out, hi_tab, hj_tab, src, dst, P, C, W0, Hd, act, slope, G, dhi, dhj, dW0,
                              db0, ws, ws_bytes, stream);
}

// Backward of the scorer with the nll read-out folded in (loss = -mean_p out[p, target[p]], LLP.py:235): dOut is generated
// by the producer warps, never stored.
MSHA_API int msha_score_mlp_nll_bwd(const int64_t* target, const float* gout, const float* out, const float* hi_tab,
                                    const float* hj_tab, const int64_t* src, const int64_t* dst, int64_t P, int64_t C,
                                    const float* W0, int64_t Hd, int act, float slope, float* G, float* dhi, float* dhj,
                                    float* dW0, float* db0, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(target != nullptr && gout != nullptr, "score_mlp_nll_bwd: target / gout is NULL");
    return score_mlp_bwd_impl(nullptr, target, gout, out, hi_tab, hj_tab, src, dst, P, C, W0, Hd, act, slope, G, dhi, dhj, dW0,
                              db0, ws, ws_bytes, stream);
}
