// K-2/K-3/K-4: edge kernels of the graph-attention hot path (all fp32, int32 CSR).
//
//   gat_fwd        per destination row: SDDMM logit  e_ij = lrelu(s_nbr[j] + s_self[i])   (Ours.py:64-65)
//                  + masked row softmax (Ours.py:66-68) + dropout (Ours.py:69)
//                  + weighted aggregation out_i = sum_j alpha_ij feat[j]                    (Ours.py:98)
//                  for H heads at once (feat is [n, H, D]); optional fused ELU.
//   gat_bwd_rows   per row: d alpha = dOut_i . feat_j (+ dT_j . fT_i), softmax/LeakyReLU backward,
//                  d s_self, writes d logit per edge.
//   spmm_csc       per column (CSC + perm): out_j = sum_i w_ij feat[i]  (alpha.T @ h2, Ours.py:100, and the
//                  d feat pass of the backward), optional per-column sums of a second edge array.
//
// Mapping: one warp per row/column; lanes tile the C = H*D channels in VW-wide vectors (VW = 4 -> 128-bit
// loads), VPL vectors per lane.  Per-edge weights of the current 32-edge chunk are staged in shared memory.
// Masked edges (col < 0, j = ~col) are the reference's isolated-row semantics: logit = -9e15, no gradient.
#include "common.cuh"
#include <stdlib.h>

MSHA_DEFINE_DROP_EPOCH_HOOK(gat)
#include "msha_b200.h"

#define NEG_MASK_F (-9e15f)

constexpr int GAT_WARPS = 8;
constexpr int GAT_THREADS = GAT_WARPS * 32;

template <int VW> struct VecT;
template <> struct VecT<4> {
    static __device__ __forceinline__ void load(const float* p, float (&r)[4]) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&r)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
    }
};
template <> struct VecT<1> {
    static __device__ __forceinline__ void load(const float* p, float (&r)[1]) { r[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float* p, const float (&r)[1]) { *p = r[0]; }
};

__device__ __forceinline__ float lrelu(float x, float slope) { return x > 0.f ? x : x * slope; }
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

// Hub rows / columns (power-law graphs): entries beyond `seg_limit` per row are processed as independent segments
// of at most seg_limit edges (one warp each) and merged afterwards, so no warp walks a 100k-edge row alone.
struct HubArgs {
    int seg_limit;                 // 0: disabled
    int n_segs;
    const int32_t* seg_item;       // row (column) id of the segment
    const int32_t* seg_beg;
    const int32_t* seg_end;
    int n_hub;
    const int32_t* hub_ids;        // the split rows (columns)
    const int32_t* hub_seg_ptr;    // [n_hub + 1] into the segment arrays
};

struct DropArgs {
    uint32_t thr;       // keep iff philox word >= thr ; 0 -> dropout inactive
    float inv_keep;     // 1/(1-p)
    uint64_t seed;
    uint32_t stream;
};

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// PIPE (column-block pass of the partitioned forward, dist.py): the row's edge range is [rowptr[row], rowend[row]) --
// the slots whose neighbour lives in one owner block --, the softmax statistics come in as the row's log-sum-exp over
// ALL its edges (gat_stats_kernel), alpha = exp(logit - lse) is final, and the aggregate is added to `out`
// (accumulate != 0: read-modify-write; hub segments: atomics) so that the blocks can be processed as they arrive.
template <int VW, int VPL, bool PIPE = false, int WPC = GAT_WARPS, int MINB = 1, int UNR = 4>
__global__ void __launch_bounds__(WPC * 32, MINB)
gat_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rowend,
               const int32_t* __restrict__ col, int n_rows,
               const float* __restrict__ s_nbr, const float* __restrict__ s_self,
               const float* __restrict__ feat, int H, int D, float slope,
               const float* __restrict__ alpha_in, float* __restrict__ alpha_out,
               float* __restrict__ out, int act, float* __restrict__ lse_out, DropArgs drop, HubArgs hub,
               float* __restrict__ part, const float* __restrict__ lse_in = nullptr, int accumulate = 0) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * WPC + warp;
    // the hub segments (the longest work items, up to seg_limit edges each) come FIRST in the grid: dispatched last they
    // were the tail of every launch
    int row, beg, end, seg = -1;
    if (w >= hub.n_segs) {
        row = w - hub.n_segs;
        if (row >= n_rows) return;
        beg = rowptr[row]; end = rowend[row];
        if (hub.seg_limit && end - beg > hub.seg_limit) return;         // handled segment-wise
        if (PIPE && accumulate && beg == end) return;                   // nothing to add from this block
    } else {
        seg = w;
        row = hub.seg_item[seg]; beg = hub.seg_beg[seg]; end = hub.seg_end[seg];
    }
    const bool part_mode = seg >= 0;       // partial (un-normalised) result of one segment; merged by gat_fwd_merge_kernel
    float* sm_w = smem + warp * 32 * H;
    const int C = H * D;
    const int nvec = C / VW;

    int head_of[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        int v = lane + 32 * k;
        head_of[k] = (v < nvec) ? (v * VW) / D : 0;
    }

    // ---- softmax statistics
    // H a power of two: lanes are (edge slot es, head hh) pairs -> coalesced s_nbr / alpha accesses, one online
    // max/sum pass and log2(32/H) shuffle steps per row.  Other H: per-head loops, lane h keeps (max, sum) of head h.
    const bool hp2 = (H & (H - 1)) == 0;
    const int hh = lane & (H - 1), es = hp2 ? lane / H : 0, EPI = hp2 ? 32 / H : 1;
    float m_stat = 0.f, l_stat = 1.f;
    if (PIPE) {
        // hp2 only (checked on the host).  The row's (max, sum) over ALL its edges: a = exp(lg - m) / l is the very
        // expression of the fused kernel -- a log-sum-exp would bias sum_j alpha_ij by ~1e-6 (one rounding of m + log l
        // shared by the whole row), which the softmax backward turns into a spurious d s_self = r_i (1 - sum alpha).
        m_stat = lse_in[(int64_t)row * 2 * H + hh];
        l_stat = lse_in[(int64_t)row * 2 * H + H + hh];
    } else if (alpha_in == nullptr && hp2) {
        const float ss = s_self[(int64_t)row * H + hh];
        float m = -INFINITY, l = 0.f;
        for (int e = beg + es; e < end; e += EPI) {
            const int c = col[e];
            const float lg = c < 0 ? NEG_MASK_F : lrelu(__ldg(s_nbr + (int64_t)c * H + hh) + ss, slope);
            if (lg > m) { l = l * expf(m - lg) + 1.f; m = lg; } else { l += expf(lg - m); }
        }
        for (int o = H; o < 32; o <<= 1) {
            const float m2 = __shfl_xor_sync(FULL_MASK, m, o), l2 = __shfl_xor_sync(FULL_MASK, l, o);
            const float M = fmaxf(m, m2);
            l = (m == -INFINITY ? 0.f : l * expf(m - M)) + (m2 == -INFINITY ? 0.f : l2 * expf(m2 - M));
            m = M;
        }
        m_stat = m; l_stat = l;                     // every lane: statistics of its head hh
        if (!part_mode && lse_out != nullptr && lane < H)
            lse_out[(int64_t)row * H + lane] = (end > beg) ? m_stat + logf(l_stat) : -INFINITY;
    } else if (alpha_in == nullptr) {
        for (int h = 0; h < H; ++h) {
            const float ss = s_self[(int64_t)row * H + h];
            float mx = -INFINITY;
            for (int e = beg + lane; e < end; e += 32) {
                int c = col[e];
                float lg = c < 0 ? NEG_MASK_F : lrelu(__ldg(s_nbr + (int64_t)c * H + h) + ss, slope);
                mx = fmaxf(mx, lg);
            }
            mx = warp_max(mx);
            float sum = 0.f;
            for (int e = beg + lane; e < end; e += 32) {
                int c = col[e];
                float lg = c < 0 ? NEG_MASK_F : lrelu(__ldg(s_nbr + (int64_t)c * H + h) + ss, slope);
                sum += expf(lg - mx);
            }
            sum = warp_sum(sum);
            if (lane == h) { m_stat = mx; l_stat = sum; }
        }
        // log-sum-exp of the row's logits (HGANE's joint normaliser needs the un-normalised sums, HGANE.py:61-62)
        if (!part_mode && lse_out != nullptr && lane < H)
            lse_out[(int64_t)row * H + lane] = (end > beg) ? m_stat + logf(l_stat) : -INFINITY;
    }
    int* sm_j = reinterpret_cast<int*>(smem + WPC * 32 * H) + warp * 32;     // neighbour ids of the current chunk

    float acc[VPL][VW];
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int q = 0; q < VW; ++q) acc[k][q] = 0.f;

    for (int e0 = beg; e0 < end; e0 += 32) {
        int j = 0;
        if (hp2) {
            // (edge slot, head) lanes: H sub-iterations cover the 32 edges of the chunk
            for (int sub = 0; sub < H; ++sub) {
                const int t = sub * EPI + es;
                const int e = e0 + t;
                float a = 0.f;
                int jj = 0;
                if (e < end) {
                    const int c = col[e];
                    const bool masked = c < 0;
                    jj = masked ? ~c : c;
                    if (alpha_in != nullptr) {
                        a = alpha_in[(int64_t)e * H + hh];
                    } else {
                        const float lg = masked ? NEG_MASK_F
                                                : lrelu(__ldg(s_nbr + (int64_t)jj * H + hh) + s_self[(int64_t)row * H + hh], slope);
                        a = (part_mode && !PIPE) ? expf(lg - m_stat) : expf(lg - m_stat) / l_stat;
                        if (alpha_out) alpha_out[(int64_t)e * H + hh] = a;            // coalesced
                    }
                    if (drop.thr) a *= dropout_scale(drop.seed, drop.stream, (uint64_t)e * H + hh, drop.thr, drop.inv_keep);
                }
                sm_w[t * H + hh] = a;
                if (hh == 0) sm_j[t] = jj;
            }
        } else {
            const int e = e0 + lane;
            const bool valid = e < end;
            const int c = valid ? col[e] : 0;
            const bool masked = c < 0;
            j = masked ? ~c : c;
            sm_j[lane] = j;
            for (int h = 0; h < H; ++h) {
                const float mh = __shfl_sync(FULL_MASK, m_stat, h);
                const float lh = __shfl_sync(FULL_MASK, l_stat, h);
                float a = 0.f;
                if (valid) {
                    if (alpha_in != nullptr) {
                        a = alpha_in[(int64_t)e * H + h];
                    } else {
                        float lg = masked ? NEG_MASK_F
                                          : lrelu(__ldg(s_nbr + (int64_t)j * H + h) + s_self[(int64_t)row * H + h], slope);
                        a = part_mode ? expf(lg - mh) : expf(lg - mh) / lh;      // hub segments: normalised at merge time
                        if (alpha_out) alpha_out[(int64_t)e * H + h] = a;
                    }
                    if (drop.thr) a *= dropout_scale(drop.seed, drop.stream, (uint64_t)e * H + h, drop.thr, drop.inv_keep);
                }
                sm_w[lane * H + h] = a;
            }
        }
        __syncwarp();
        const int cnt = min(32, end - e0);
#pragma unroll UNR
        for (int t = 0; t < cnt; ++t) {
            const int jt = sm_j[t];
            const float* f = feat + (int64_t)jt * C;
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int v = lane + 32 * k;
                if (v < nvec) {
                    float x[VW];
                    VecT<VW>::load(f + v * VW, x);
                    const float w = sm_w[t * H + head_of[k]];
#pragma unroll
                    for (int q = 0; q < VW; ++q) acc[k][q] = fmaf(w, x[q], acc[k][q]);
                }
            }
        }
        __syncwarp();
    }
    if (PIPE) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
                float* o = out + (int64_t)row * C + v * VW;
                if (part_mode) {                       // hub segment: rows zeroed (or carried over) by the host wrapper
#pragma unroll
                    for (int q = 0; q < VW; ++q) atomicAdd(o + q, acc[k][q]);
                } else {
                    if (accumulate) {
                        float x[VW];
                        VecT<VW>::load(o, x);
#pragma unroll
                        for (int q = 0; q < VW; ++q) acc[k][q] += x[q];
                    }
                    VecT<VW>::store(o, acc[k]);
                }
            }
        }
        return;
    }
    if (part_mode) {
        float* P = part + (int64_t)seg * (2 * H + C);
        if (lane < H) { P[lane] = m_stat; P[H + lane] = l_stat; }
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
#pragma unroll
                for (int q = 0; q < VW; ++q) P[2 * H + v * VW + q] = acc[k][q];
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        if (v < nvec) {
            if (act == 1) {
#pragma unroll
                for (int q = 0; q < VW; ++q) acc[k][q] = elu1(acc[k][q]);
            }
            VecT<VW>::store(out + (int64_t)row * C + v * VW, acc[k]);
        }
    }
}

// Merge of the hub-row segments: online-softmax combination (m, l, acc) -> out, lse, and the final scaling of the
// un-normalised alpha the segments stored.  plain_sum: weights were given (no softmax) -> plain sum of partials.
__global__ void gat_fwd_merge_kernel(HubArgs hub, float* __restrict__ part, int H, int D, float* __restrict__ out,
                                     int act, float* __restrict__ lse_out, int plain_sum) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * GAT_WARPS + warp;
    if (i >= hub.n_hub) return;
    const int C = H * D;
    const int row = hub.hub_ids[i];
    const int s0 = hub.hub_seg_ptr[i], s1 = hub.hub_seg_ptr[i + 1];
    // pass A: per-head combination weights  sc_s = exp(m_s - M) / L, stored in the segment record's `l` slot
    if (lane < H) {
        if (plain_sum) {
            for (int s = s0; s < s1; ++s) part[(int64_t)s * (2 * H + C) + H + lane] = 1.f;
        } else {
            float M = -INFINITY, L = 0.f;
            for (int s = s0; s < s1; ++s) M = fmaxf(M, part[(int64_t)s * (2 * H + C) + lane]);
            for (int s = s0; s < s1; ++s) {
                const float* P = part + (int64_t)s * (2 * H + C);
                L += P[H + lane] * expf(P[lane] - M);
            }
            for (int s = s0; s < s1; ++s) {
                float* P = part + (int64_t)s * (2 * H + C);
                P[H + lane] = expf(P[lane] - M) / L;
            }
            if (lse_out) lse_out[(int64_t)row * H + lane] = M + logf(L);
        }
    }
    __syncwarp();
    // pass B: weighted sum of the partial accumulators
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        if (c < C) {
            const int h = c / D;
            float acc = 0.f;
            for (int s = s0; s < s1; ++s) {
                const float* P = part + (int64_t)s * (2 * H + C);
                acc = fmaf(P[2 * H + c], P[H + h], acc);
            }
            out[(int64_t)row * C + c] = act == 1 ? elu1(acc) : acc;
        }
    }
}

// final scaling of the un-normalised alpha stored by the hub segments: one warp per segment
__global__ void gat_fwd_rescale_kernel(HubArgs hub, const float* __restrict__ part, int H, int D, float* __restrict__ alpha) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * GAT_WARPS + warp;
    if (s >= hub.n_segs) return;
    const float* P = part + (int64_t)s * (2 * H + H * D);
    const int64_t b = (int64_t)hub.seg_beg[s] * H, e = (int64_t)hub.seg_end[s] * H;
    if ((H & (H - 1)) == 0 && H <= 32) {
        const float sc = P[H + (lane & (H - 1))];          // 32 % H == 0: a lane always hits the same head
        for (int64_t idx = b + lane; idx < e; idx += 32) alpha[idx] *= sc;
    } else {
        for (int64_t idx = b + lane; idx < e; idx += 32) alpha[idx] *= P[H + (int)(idx % H)];
    }
}

// ---------------------------------------------------------------------------------------------
// backward, row pass
// ---------------------------------------------------------------------------------------------
// EB: edges whose feature gathers are in flight together; MINB: resident CTAs per SM the register budget targets.
// <4, 2> suits dense rows (long gather streams), <2, 3> sparse power-law rows (latency hidden by occupancy).
// LPH_T > 0 (VW == 4, no dT/fT term): lanes per head known at compile time -- the per-edge dot products of EB edges are
// reduced together by a transposing butterfly (EB/2 + EB/4 + ... shuffles for EB edges instead of log2(LPH) each) and
// every (edge, head) sum is stored once, by the lane that ends up owning it.
template <int VW, int VPL, int EB, int MINB, int LPH_T = 0, int WPC = GAT_WARPS>
__global__ void __launch_bounds__(WPC * 32, MINB)
gat_bwd_rows_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n_rows,
                    const float* __restrict__ s_nbr, const float* __restrict__ s_self, float slope,
                    const float* __restrict__ alpha, const float* __restrict__ feat,
                    const float* __restrict__ dout, const float* __restrict__ outp, int act,
                    float* __restrict__ dz_out,
                    const float* __restrict__ dT, const float* __restrict__ fT,
                    const float* __restrict__ dalpha_extra, const float* __restrict__ dlse,
                    int H, int D, float* __restrict__ dlogit, float* __restrict__ ds_self, DropArgs drop, HubArgs hub,
                    int mode, float* __restrict__ r_buf) {
    // mode 0: one warp per row (hub rows skipped); mode 1 / 2: hub segments, phase 1 (d alpha, partial r -> r_buf) and
    // phase 2 (softmax / LeakyReLU backward with the complete r, partial d s_self -> atomics)
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * WPC + warp;
    int row, beg, end;
    if (mode == 0) {
        if (w >= n_rows) return;
        row = w; beg = rowptr[row]; end = rowptr[row + 1];
        if (hub.seg_limit && end - beg > hub.seg_limit) return;
    } else {
        if (w >= hub.n_segs) return;
        row = hub.seg_item[w]; beg = hub.seg_beg[w]; end = hub.seg_end[w];
    }
    float* sm_d = smem + warp * (32 * H + 2 * H);
    float* sm_r = sm_d + 32 * H;
    float* sm_ds = sm_r + H;
    const int C = H * D;
    const int nvec = C / VW;
    const int LPH = D / VW;                        // lanes (vectors) per head
    const bool pow2 = (LPH & (LPH - 1)) == 0;
    const bool softmax_mode = (s_nbr != nullptr);
    const bool hp2 = (H & (H - 1)) == 0;              // (edge slot, head) lane mapping for the per-edge scalar work
    const int hh = lane & (H - 1), es = hp2 ? lane / H : 0, EPI = hp2 ? 32 / H : 1;
    float r_lane = 0.f;

    float dz[VPL][VW], ft[VPL][VW];
    int head_of[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        head_of[k] = (v < nvec) ? (v * VW) / D : 0;
#pragma unroll
        for (int q = 0; q < VW; ++q) { dz[k][q] = 0.f; ft[k][q] = 0.f; }
        if (v < nvec) {
            VecT<VW>::load(dout + (int64_t)row * C + v * VW, dz[k]);
            if (act == 1) {
                float o[VW];
                VecT<VW>::load(outp + (int64_t)row * C + v * VW, o);
#pragma unroll
                for (int q = 0; q < VW; ++q) dz[k][q] = o[q] > 0.f ? dz[k][q] : dz[k][q] * (o[q] + 1.f);
            }
            if (dz_out && (mode == 0 || (mode == 1 && beg == rowptr[row])))
                VecT<VW>::store(dz_out + (int64_t)row * C + v * VW, dz[k]);
            if (fT) VecT<VW>::load(fT + (int64_t)row * C + v * VW, ft[k]);
        }
    }
    for (int h = lane; h < 2 * H; h += 32) sm_r[h] = 0.f;   // sm_r and sm_ds are contiguous
    __syncwarp();

    // ---- phase 1: d alpha per edge (-> dlogit as scratch), r_h = sum alpha * dalpha
    for (int e0 = beg; mode != 2 && e0 < end; e0 += 32) {
        const int e = e0 + lane;
        const bool valid = e < end;
        const int c = valid ? col[e] : 0;
        const int j = c < 0 ? ~c : c;
        const int cnt = min(32, end - e0);
        if constexpr (LPH_T > 0) {
            static_assert(LPH_T == 0 || VW == 4, "fast path is for 128-bit vectors");
            constexpr int KPH = LPH_T > 32 ? LPH_T / 32 : 1;        // vectors of one lane that belong to the same head
            constexpr int NG = VPL / KPH;                            // heads a lane contributes to
            constexpr int G = LPH_T > 32 ? 32 : LPH_T;               // lanes to reduce across
            constexpr int LOGG = G == 32 ? 5 : (G == 16 ? 4 : (G == 8 ? 3 : (G == 4 ? 2 : (G == 2 ? 1 : 0))));
            constexpr int LOGE = EB == 8 ? 3 : (EB == 4 ? 2 : (EB == 2 ? 1 : 0));
            constexpr int T = LOGG < LOGE ? LOGG : LOGE;              // transposing steps
            static_assert(VPL % KPH == 0 && (1 << LOGG) == G && (1 << LOGE) == EB, "fast path layout");
            int estart = 0;                                          // first edge (of a batch) whose sum this lane ends up with
#pragma unroll
            for (int sidx = 0; sidx < T; ++sidx)
                if (lane & (G >> (sidx + 1))) estart += EB >> (sidx + 1);
            const bool writer = (lane & ((G >> T) - 1)) == 0;
            for (int t0 = 0; t0 < cnt; t0 += EB) {
                float x[EB][VPL][VW];
#pragma unroll
                for (int u = 0; u < EB; ++u) {
                    const int jt = __shfl_sync(FULL_MASK, j, (t0 + u) & 31);
                    const float* f = feat + (int64_t)jt * C;
#pragma unroll
                    for (int k = 0; k < VPL; ++k) {
                        if (t0 + u < cnt) {
                            VecT<VW>::load(f + (lane + 32 * k) * VW, x[u][k]);
                        } else {
#pragma unroll
                            for (int q = 0; q < VW; ++q) x[u][k][q] = 0.f;
                        }
                    }
                }
                float pr[NG][EB];
#pragma unroll
                for (int g = 0; g < NG; ++g)
#pragma unroll
                    for (int u = 0; u < EB; ++u) {
                        float acc = 0.f;
#pragma unroll
                        for (int kk = 0; kk < KPH; ++kk)
#pragma unroll
                            for (int q = 0; q < VW; ++q) acc = fmaf(x[u][g * KPH + kk][q], dz[g * KPH + kk][q], acc);
                        pr[g][u] = acc;
                    }
                int n = EB;
#pragma unroll
                for (int sidx = 0; sidx < T; ++sidx) {
                    const int o = G >> (sidx + 1);
                    const bool up = (lane & o) != 0;
                    n >>= 1;
#pragma unroll
                    for (int g = 0; g < NG; ++g)
#pragma unroll
                        for (int i = 0; i < (EB >> (sidx + 1)); ++i) {
                            const float send = up ? pr[g][i] : pr[g][i + n];
                            const float keep = up ? pr[g][i + n] : pr[g][i];
                            pr[g][i] = keep + __shfl_xor_sync(FULL_MASK, send, o);
                        }
                }
#pragma unroll
                for (int sidx = T; sidx < LOGG; ++sidx) {
                    const int o = G >> (sidx + 1);
#pragma unroll
                    for (int g = 0; g < NG; ++g)
#pragma unroll
                        for (int i = 0; i < (EB >> T); ++i) pr[g][i] += __shfl_xor_sync(FULL_MASK, pr[g][i], o);
                }
                if (writer) {
#pragma unroll
                    for (int g = 0; g < NG; ++g)
#pragma unroll
                        for (int i = 0; i < (EB >> T); ++i)
                            sm_d[(t0 + estart + i) * H + (lane + 32 * g * KPH) / LPH_T] = pr[g][i];
                }
            }
        } else {
        for (int q = lane; q < 32 * H; q += 32) sm_d[q] = 0.f;
        __syncwarp();
        for (int t0 = 0; t0 < cnt; t0 += EB) {
            float x[EB][VPL][VW], y[EB][VPL][VW];
#pragma unroll
            for (int u = 0; u < EB; ++u) {
                const int jt = __shfl_sync(FULL_MASK, j, (t0 + u) & 31);
                const bool on = t0 + u < cnt;
                const float* f = feat + (int64_t)jt * C;
                const float* g = dT ? dT + (int64_t)jt * C : nullptr;
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    const int v = lane + 32 * k;
#pragma unroll
                    for (int q = 0; q < VW; ++q) { x[u][k][q] = 0.f; y[u][k][q] = 0.f; }
                    if (on && v < nvec) {
                        VecT<VW>::load(f + v * VW, x[u][k]);
                        if (g) VecT<VW>::load(g + v * VW, y[u][k]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < EB; ++u) {
                const int t = t0 + u;
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    const int v = lane + 32 * k;
                    float p = 0.f;
#pragma unroll
                    for (int q = 0; q < VW; ++q) p = fmaf(x[u][k][q], dz[k][q], p);
                    if (dT) {
#pragma unroll
                        for (int q = 0; q < VW; ++q) p = fmaf(y[u][k][q], ft[k][q], p);
                    }
                    if (pow2) {
                        if (LPH >= 32) {
                            p = warp_sum(p);
                            if (lane == 0 && (32 * k) < nvec && t < cnt) sm_d[t * H + head_of[k]] += p;
                        } else {
                            for (int o = LPH >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(FULL_MASK, p, o);
                            if ((lane & (LPH - 1)) == 0 && v < nvec && t < cnt) sm_d[t * H + head_of[k]] += p;
                        }
                    } else if (v < nvec && t < cnt) {
                        atomicAdd(&sm_d[t * H + head_of[k]], p);
                    }
                }
            }
        }
        }
        __syncwarp();
        if (hp2) {
            // (edge slot, head) lanes: coalesced alpha / dlogit accesses, r accumulated per lane
            for (int sub = 0; sub < H; ++sub) {
                const int t = sub * EPI + es;
                const int ee = e0 + t;
                if (ee < end) {
                    float da = sm_d[t * H + hh];
                    if (drop.thr) da *= dropout_scale(drop.seed, drop.stream, (uint64_t)ee * H + hh, drop.thr, drop.inv_keep);
                    if (dalpha_extra) da += dalpha_extra[(int64_t)ee * H + hh];   // grad w.r.t. the pre-dropout alpha
                    if (softmax_mode) r_lane = fmaf(alpha[(int64_t)ee * H + hh], da, r_lane);
                    dlogit[(int64_t)ee * H + hh] = da;
                }
            }
        } else {
            for (int h = 0; h < H; ++h) {
                float da = 0.f, a = 0.f;
                if (valid) {
                    da = sm_d[lane * H + h];
                    if (drop.thr) da *= dropout_scale(drop.seed, drop.stream, (uint64_t)e * H + h, drop.thr, drop.inv_keep);
                    if (dalpha_extra) da += dalpha_extra[(int64_t)e * H + h];   // grad w.r.t. the pre-dropout alpha
                    if (softmax_mode) a = alpha[(int64_t)e * H + h];
                    dlogit[(int64_t)e * H + h] = da;
                }
                if (softmax_mode) {
                    float s = warp_sum(a * da);
                    if (lane == 0) sm_r[h] += s;
                }
            }
        }
        __syncwarp();
    }
    if (hp2 && softmax_mode && mode != 2) {           // fold the per-lane partials: every lane gets r of its head
        for (int o = H; o < 32; o <<= 1) r_lane += __shfl_xor_sync(FULL_MASK, r_lane, o);
        if (lane < H) sm_r[lane] = r_lane;
        __syncwarp();
    }
    if (!softmax_mode) return;
    if (mode == 1) {                                   // hub segment: publish the partial r and stop
        for (int h = lane; h < H; h += 32) atomicAdd(&r_buf[(int64_t)row * H + h], sm_r[h]);
        return;
    }
    if (mode == 2) {
        for (int h = lane; h < H; h += 32) sm_r[h] = r_buf[(int64_t)row * H + h];
        __syncwarp();
    }
    // d lse_i / d e_ij = alpha_ij : fold the upstream gradient of the row's log-sum-exp into r_h
    if (dlse != nullptr) {
        __syncwarp();
        for (int h = lane; h < H; h += 32) sm_r[h] -= dlse[(int64_t)row * H + h];
        __syncwarp();
    }

    // ---- phase 2: softmax + LeakyReLU backward
    if (hp2) {
        const float r = sm_r[hh];
        const float ss = s_self[(int64_t)row * H + hh];
        float ds = 0.f;
        for (int e0 = beg + es; e0 < end; e0 += 4 * EPI) {      // 4 independent edges per lane in flight
            int cc[4];
            float av[4], dav[4], sv[4];
            bool on[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * EPI;
                on[u] = e < end;
                cc[u] = on[u] ? col[e] : 0;
                av[u] = on[u] ? alpha[(int64_t)e * H + hh] : 0.f;
                dav[u] = on[u] ? dlogit[(int64_t)e * H + hh] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = cc[u] < 0 ? ~cc[u] : cc[u];
                sv[u] = on[u] ? __ldg(s_nbr + (int64_t)j * H + hh) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (on[u]) {
                    const float de = av[u] * (dav[u] - r);
                    const float pre = sv[u] + ss;
                    const float dl = cc[u] < 0 ? 0.f : (pre > 0.f ? de : de * slope);
                    dlogit[(int64_t)(e0 + u * EPI) * H + hh] = dl;
                    ds += dl;
                }
            }
        }
        for (int o = H; o < 32; o <<= 1) ds += __shfl_xor_sync(FULL_MASK, ds, o);
        if (lane < H) sm_ds[lane] = ds;
    } else {
        for (int e0 = beg; e0 < end; e0 += 32) {
            const int e = e0 + lane;
            const bool valid = e < end;
            const int c = valid ? col[e] : 0;
            const bool masked = c < 0;
            const int j = masked ? ~c : c;
            for (int h = 0; h < H; ++h) {
                float dl = 0.f;
                if (valid) {
                    const float a = alpha[(int64_t)e * H + h];
                    const float da = dlogit[(int64_t)e * H + h];
                    const float de = a * (da - sm_r[h]);
                    const float pre = __ldg(s_nbr + (int64_t)j * H + h) + s_self[(int64_t)row * H + h];
                    dl = masked ? 0.f : (pre > 0.f ? de : de * slope);
                    dlogit[(int64_t)e * H + h] = dl;
                }
                float s = warp_sum(dl);
                if (lane == 0) sm_ds[h] += s;
            }
        }
    }
    __syncwarp();
    if (ds_self) {
        if (mode == 2) {
            for (int h = lane; h < H; h += 32) atomicAdd(&ds_self[(int64_t)row * H + h], sm_ds[h]);
        } else {
            for (int h = lane; h < H; h += 32) ds_self[(int64_t)row * H + h] = sm_ds[h];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// column pass (CSC + perm)
// ---------------------------------------------------------------------------------------------
template <int VW, int VPL, int WPC = GAT_WARPS, int MINB = 1, int UNR = 4>
__global__ void __launch_bounds__(WPC * 32, MINB)
spmm_csc_kernel(const int32_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
                const int32_t* __restrict__ perm, int n_cols,
                const float* __restrict__ w, const float* __restrict__ feat, int H, int D,
                float* __restrict__ out, int accumulate,
                const float* __restrict__ esum_in, float* __restrict__ esum_out, DropArgs drop, HubArgs hub) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w0 = blockIdx.x * WPC + warp;
    int cidx, beg, end;
    bool part_mode = false;
    if (w0 >= hub.n_segs) {                  // hub segments first in the grid (longest items: not the launch's tail)
        cidx = w0 - hub.n_segs;
        if (cidx >= n_cols) return;
        beg = colptr[cidx]; end = colptr[cidx + 1];
        if (hub.seg_limit && end - beg > hub.seg_limit) return;
    } else {
        const int seg = w0;
        cidx = hub.seg_item[seg]; beg = hub.seg_beg[seg]; end = hub.seg_end[seg];
        part_mode = true;                   // partial sums of one segment: atomically added (rows zeroed beforehand)
    }
    float* sm_w = smem + warp * (32 * H + H);
    float* sm_s = sm_w + 32 * H;
    const int C = H * D;
    const int nvec = C / VW;

    int head_of[VPL];
    float acc[VPL][VW];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        head_of[k] = (v < nvec) ? (v * VW) / D : 0;
#pragma unroll
        for (int q = 0; q < VW; ++q) acc[k][q] = 0.f;
    }
    for (int h = lane; h < H; h += 32) sm_s[h] = 0.f;
    __syncwarp();

    for (int k0 = beg; k0 < end; k0 += 32) {
        const int kk = k0 + lane;
        const bool valid = kk < end;
        const int i = valid ? rowidx[kk] : 0;
        const int e = valid ? perm[kk] : 0;
        for (int h = 0; h < H; ++h) {
            float a = 0.f, s = 0.f;
            if (valid) {
                a = w ? w[(int64_t)e * H + h] : 0.f;
                if (drop.thr) a *= dropout_scale(drop.seed, drop.stream, (uint64_t)e * H + h, drop.thr, drop.inv_keep);
                if (esum_in) s = esum_in[(int64_t)e * H + h];
            }
            sm_w[lane * H + h] = a;
            if (esum_in) {
                s = warp_sum(s);
                if (lane == 0) sm_s[h] += s;
            }
        }
        __syncwarp();
        if (w) {
            const int cnt = min(32, end - k0);
#pragma unroll UNR
            for (int t = 0; t < cnt; ++t) {
                const int it = __shfl_sync(FULL_MASK, i, t);
                const float* f = feat + (int64_t)it * C;
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    const int v = lane + 32 * k;
                    if (v < nvec) {
                        float x[VW];
                        VecT<VW>::load(f + v * VW, x);
                        const float ww = sm_w[t * H + head_of[k]];
#pragma unroll
                        for (int q = 0; q < VW; ++q) acc[k][q] = fmaf(ww, x[q], acc[k][q]);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (w) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int v = lane + 32 * k;
            if (v < nvec) {
                float* o = out + (int64_t)cidx * C + v * VW;
                if (part_mode) {
#pragma unroll
                    for (int q = 0; q < VW; ++q) atomicAdd(o + q, acc[k][q]);
                    continue;
                }
                if (accumulate) {
#pragma unroll
                    for (int q = 0; q < VW; ++q) acc[k][q] += o[q];
                }
                VecT<VW>::store(o, acc[k]);
            }
        }
    }
    if (esum_out) {
        if (part_mode) {
            for (int h = lane; h < H; h += 32) atomicAdd(&esum_out[(int64_t)cidx * H + h], sm_s[h]);
        } else {
            for (int h = lane; h < H; h += 32) esum_out[(int64_t)cidx * H + h] = sm_s[h];
        }
    }
}

// zero the rows `ids` of a [n, C] matrix (hub rows / columns that are accumulated with atomics)
__global__ void zero_rows_kernel(float* __restrict__ x, const int32_t* __restrict__ ids, int n_ids, int C) {
    const int i = blockIdx.x;
    if (i >= n_ids) return;
    float* r = x + (int64_t)ids[i] * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) r[c] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
static int pick_layout(int H, int D, int* vw, int* vpl) {
    const int C = H * D;
    if (D % 4 == 0) {
        int nvec = C / 4;
        *vw = 4;
        *vpl = nvec <= 32 ? 1 : (nvec <= 64 ? 2 : (nvec <= 128 ? 4 : 0));
        if (*vpl) return 0;
    }
    *vw = 1;
    *vpl = C <= 32 ? 1 : (C <= 64 ? 2 : (C <= 128 ? 4 : (C <= 256 ? 8 : 0)));
    return *vpl ? 0 : -1;
}

static DropArgs make_drop(float p, uint64_t seed, uint32_t stream) {
    DropArgs d;
    d.thr = 0; d.inv_keep = 1.f; d.seed = seed; d.stream = stream;
    if (p > 0.f) {
        double t = (double)p * 4294967296.0;
        d.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        if (d.thr == 0) d.thr = 1;                      // p tiny but > 0
        d.inv_keep = 1.f / (1.f - p);
    }
    return d;
}

#define DISPATCH_LAYOUT(VWv, VPLv, CALL)                        \
    if (VWv == 4 && VPLv == 1) { CALL(4, 1); }                  \
    else if (VWv == 4 && VPLv == 2) { CALL(4, 2); }             \
    else if (VWv == 4 && VPLv == 4) { CALL(4, 4); }             \
    else if (VWv == 1 && VPLv == 1) { CALL(1, 1); }             \
    else if (VWv == 1 && VPLv == 2) { CALL(1, 2); }             \
    else if (VWv == 1 && VPLv == 4) { CALL(1, 4); }             \
    else { CALL(1, 8); }

// Warps per CTA of the three edge kernels on the 256-channel layout.  A CTA lives as long as its longest row: on power-law
// graphs 8 consecutive rows per CTA leave most warps of a resident CTA finished and idle behind one long row (ncu, round
// 1: 17 % of the warp slots active in the row pass).  Fewer warps per CTA free the slots as rows finish.
// Measured on the 100 M-edge R-MAT graph (profiles/r02_gat_variant_sweep.txt): 8 -> 1 warps per CTA together with the
// register budget of 32 resident warps per SM takes the row pass from 38.3 to 15.4 ms, the column pass from 26.3 to 12.9 ms
// and the forward from 24.4 to 19.6 ms per layer.  MSHA_GAT_WPC = 1 (default) | 2 | 8, MSHA_GAT_ROWS_EB = 4 (default) | 8 | 2,
// MSHA_GAT_FWD_VAR / MSHA_GAT_CSC_VAR = 3 (default: 2 edges unrolled, 32 CTAs / SM) | 0 | 1 | 2 -- tuning knobs, read once.
static int gat_wpc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MSHA_GAT_WPC");
        v = e ? atoi(e) : 1;
        if (v != 8 && v != 2) v = 1;
    }
    return v;
}
static int gat_env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int gat_fwd_var() { static int v = -1; if (v < 0) v = gat_env_int("MSHA_GAT_FWD_VAR", 3); return v; }
static int gat_csc_var() { static int v = -1; if (v < 0) v = gat_env_int("MSHA_GAT_CSC_VAR", 3); return v; }
static int gat_rows_eb() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MSHA_GAT_ROWS_EB");
        v = e ? atoi(e) : 4;
        if (v != 8 && v != 2) v = 4;
    }
    return v;
}

// host view of the hub description (include/msha_b200.h: msha_hub_t)
struct msha_hub_host {
    int32_t seg_limit, n_segs;
    const int32_t* seg_item;
    const int32_t* seg_beg;
    const int32_t* seg_end;
    int32_t n_hub, pad_;
    const int32_t* hub_ids;
    const int32_t* hub_seg_ptr;
};
static HubArgs make_hub(const void* hp) {
    HubArgs h;
    h.seg_limit = 0; h.n_segs = 0; h.seg_item = h.seg_beg = h.seg_end = nullptr; h.n_hub = 0; h.hub_ids = h.hub_seg_ptr = nullptr;
    if (hp) {
        const msha_hub_host* s = (const msha_hub_host*)hp;
        if (s->n_hub > 0 && s->n_segs > 0) {
            h.seg_limit = s->seg_limit; h.n_segs = s->n_segs; h.seg_item = s->seg_item; h.seg_beg = s->seg_beg;
            h.seg_end = s->seg_end; h.n_hub = s->n_hub; h.hub_ids = s->hub_ids; h.hub_seg_ptr = s->hub_seg_ptr;
        }
    }
    return h;
}

// floats of scratch msha_gat_fwd needs for a hub description with n_segs segments
MSHA_API size_t msha_gat_fwd_hub_scratch_floats(int64_t n_segs, int H, int D) { return (size_t)n_segs * (2 * H + H * D); }

// alpha_in == NULL: compute softmax attention from (s_nbr, s_self) and optionally store it in alpha_out.
// alpha_in != NULL: plain weighted SpMM with the given per-edge, per-head weights.
// hub (nullable): rows with more than hub->seg_limit edges are processed as segments; hub_scratch holds their partials.
MSHA_API int msha_gat_fwd(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* s_nbr,
                          const float* s_self, const float* feat, int H, int D, float slope, const float* alpha_in,
                          float* alpha_out, float* out, int act, float* lse_out, float drop_p, uint64_t drop_seed,
                          const msha_hub_t* hub_p, float* hub_scratch, void* stream) {
    MSHA_REQUIRE(H >= 1 && H <= 32 && D >= 1, "gat_fwd: need 1 <= H <= 32, D >= 1");
    MSHA_REQUIRE(alpha_in != nullptr || (s_nbr != nullptr && s_self != nullptr), "gat_fwd: scores or alpha required");
    MSHA_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31), "gat_fwd: bad n_rows");
    int vw, vpl;
    MSHA_REQUIRE(pick_layout(H, D, &vw, &vpl) == 0, "gat_fwd: unsupported channel count H*D=%d", H * D);
    if (n_rows == 0) return 0;
    HubArgs hub = make_hub(hub_p);
    MSHA_REQUIRE(hub.n_segs == 0 || hub_scratch != nullptr, "gat_fwd: hub rows need scratch");
    MSHA_REQUIRE(hub.n_segs == 0 || alpha_in != nullptr || alpha_out != nullptr, "gat_fwd: hub rows need alpha_out");
    DropArgs drop = make_drop(drop_p, drop_seed, 2u);
    const int wpc = (vw == 4 && vpl == 2) ? gat_wpc() : GAT_WARPS;
    const unsigned grid = (unsigned)msha_cdiv(n_rows + hub.n_segs, wpc);
    const size_t smem = (size_t)wpc * (32 * H + 32) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(A, B)                                                                                             \
    gat_fwd_kernel<A, B><<<grid, GAT_THREADS, smem, st>>>(rowptr, rowptr + 1, col, (int)n_rows, s_nbr, s_self, feat, H, D, \
                                                          slope, alpha_in, alpha_out, out, act, lse_out, drop, hub, hub_scratch)
#define CALLWU(W, MB, U)                                                                                                       \
    gat_fwd_kernel<4, 2, false, W, MB, U><<<grid, W * 32, smem, st>>>(rowptr, rowptr + 1, col, (int)n_rows, s_nbr, s_self, feat, \
                                                                   H, D, slope, alpha_in, alpha_out, out, act, lse_out, drop, \
                                                                   hub, hub_scratch)
#define CALLW(W, MB) CALLWU(W, MB, 4)
    if (wpc == 1 && gat_fwd_var() == 1) { CALLWU(1, 24, 4); }
    else if (wpc == 1 && gat_fwd_var() == 2) { CALLWU(1, 16, 8); }
    else if (wpc == 1 && gat_fwd_var() == 3) { CALLWU(1, 32, 2); }
    else if (wpc == 1) { CALLW(1, 16); }
    else if (wpc == 2) { CALLW(2, 8); }
    else { DISPATCH_LAYOUT(vw, vpl, CALL) }
#undef CALLW
#undef CALLWU
#undef CALL
    MSHA_LAUNCH_OK();
    if (hub.n_hub > 0) {
        gat_fwd_merge_kernel<<<(unsigned)msha_cdiv(hub.n_hub, GAT_WARPS), GAT_THREADS, 0, st>>>(
            hub, hub_scratch, H, D, out, act, lse_out, alpha_in != nullptr ? 1 : 0);
        MSHA_LAUNCH_OK();
        if (alpha_in == nullptr && alpha_out != nullptr) {
            gat_fwd_rescale_kernel<<<(unsigned)msha_cdiv(hub.n_segs, GAT_WARPS), GAT_THREADS, 0, st>>>(hub, hub_scratch, H, D,
                                                                                                   alpha_out);
            MSHA_LAUNCH_OK();
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// partitioned forward (dist.py): softmax statistics first, then one aggregation pass per owner block of columns
// ---------------------------------------------------------------------------------------------
// stats[row] = (max_j e_ij [H], sum_j exp(e_ij - max) [H]) over the whole row (needs only the H scores per node, which
// are exchanged before the features).  H a power of two <= 32: lanes are (edge slot, head) pairs.  Hub rows: one warp per segment writes a partial
// (m, l) pair, gat_stats_merge_kernel folds them.
__global__ void __launch_bounds__(GAT_THREADS)
gat_stats_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n_rows,
                 const float* __restrict__ s_nbr, const float* __restrict__ s_self, int H, float slope,
                 float* __restrict__ lse_out, HubArgs hub, float* __restrict__ part) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * GAT_WARPS + warp;
    int row, beg, end, seg = -1;
    if (w < n_rows) {
        row = w; beg = rowptr[row]; end = rowptr[row + 1];
        if (hub.seg_limit && end - beg > hub.seg_limit) return;
    } else {
        seg = w - n_rows;
        if (seg >= hub.n_segs) return;
        row = hub.seg_item[seg]; beg = hub.seg_beg[seg]; end = hub.seg_end[seg];
    }
    const int hh = lane & (H - 1), es = lane / H, EPI = 32 / H;
    const float ss = s_self[(int64_t)row * H + hh];
    float m = -INFINITY, l = 0.f;
    for (int e0 = beg + es; e0 < end; e0 += 4 * EPI) {          // 4 independent score gathers per lane in flight
        float lg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * EPI;
            lg[u] = -INFINITY;
            if (e < end) {
                const int c = col[e];
                lg[u] = c < 0 ? NEG_MASK_F : lrelu(__ldg(s_nbr + (int64_t)c * H + hh) + ss, slope);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (lg[u] == -INFINITY) continue;
            if (lg[u] > m) { l = l * expf(m - lg[u]) + 1.f; m = lg[u]; } else { l += expf(lg[u] - m); }
        }
    }
    for (int o = H; o < 32; o <<= 1) {
        const float m2 = __shfl_xor_sync(FULL_MASK, m, o), l2 = __shfl_xor_sync(FULL_MASK, l, o);
        const float M = fmaxf(m, m2);
        l = (m == -INFINITY ? 0.f : l * expf(m - M)) + (m2 == -INFINITY ? 0.f : l2 * expf(m2 - M));
        m = M;
    }
    if (lane < H) {
        float* o = seg >= 0 ? part + (int64_t)seg * 2 * H : lse_out + (int64_t)row * 2 * H;
        o[lane] = m;
        o[H + lane] = l;
    }
}

__global__ void gat_stats_merge_kernel(HubArgs hub, const float* __restrict__ part, int H, float* __restrict__ lse_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hub.n_hub * H) return;
    const int i = idx / H, h = idx - i * H;
    const int s0 = hub.hub_seg_ptr[i], s1 = hub.hub_seg_ptr[i + 1];
    float M = -INFINITY, L = 0.f;
    for (int s = s0; s < s1; ++s) M = fmaxf(M, part[(int64_t)s * 2 * H + h]);
    for (int s = s0; s < s1; ++s) {
        const float m = part[(int64_t)s * 2 * H + h];
        if (m != -INFINITY) L += part[(int64_t)s * 2 * H + H + h] * expf(m - M);
    }
    lse_out[(int64_t)hub.hub_ids[i] * 2 * H + h] = M;
    lse_out[(int64_t)hub.hub_ids[i] * 2 * H + H + h] = L;
}

// lse: float[n_rows, 2, H] = per row the H maxima then the H sums.  hub_scratch: float[2 * H * hub->n_segs] when the graph
// has hub rows.  H must be a power of two <= 32.
MSHA_API int msha_gat_softmax_stats(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* s_nbr,
                                    const float* s_self, int H, float slope, float* lse, const msha_hub_t* hub_p,
                                    float* hub_scratch, void* stream) {
    MSHA_REQUIRE(H >= 1 && H <= 32 && (H & (H - 1)) == 0, "gat_softmax_stats: H must be a power of two <= 32");
    MSHA_REQUIRE(s_nbr != nullptr && s_self != nullptr && lse != nullptr, "gat_softmax_stats: null argument");
    MSHA_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31), "gat_softmax_stats: bad n_rows");
    if (n_rows == 0) return 0;
    HubArgs hub = make_hub(hub_p);
    MSHA_REQUIRE(hub.n_segs == 0 || hub_scratch != nullptr, "gat_softmax_stats: hub rows need scratch");
    cudaStream_t st = (cudaStream_t)stream;
    gat_stats_kernel<<<(unsigned)msha_cdiv(n_rows + hub.n_segs, GAT_WARPS), GAT_THREADS, 0, st>>>(
        rowptr, col, (int)n_rows, s_nbr, s_self, H, slope, lse, hub, hub_scratch);
    MSHA_LAUNCH_OK();
    if (hub.n_hub > 0) {
        gat_stats_merge_kernel<<<(unsigned)msha_cdiv((int64_t)hub.n_hub * H, 128), 128, 0, st>>>(hub, hub_scratch, H, lse);
        MSHA_LAUNCH_OK();
    }
    return 0;
}

// One owner block of the partitioned forward: edges [rowbeg[i], rowend[i]) of every row i, alpha = exp(logit - lse)
// written to alpha_out, out[i] (+)= sum alpha_ij feat[j].  accumulate == 0: out is overwritten (first block).
// hub: segment description of THIS block's ranges (rows whose range exceeds seg_limit); their sums arrive by atomics.
MSHA_API int msha_gat_fwd_block(const int32_t* rowbeg, const int32_t* rowend, const int32_t* col, int64_t n_rows,
                                const float* s_nbr, const float* s_self, const float* lse, const float* feat, int H,
                                int D, float slope, float* alpha_out, float* out, int accumulate, float drop_p,
                                uint64_t drop_seed, const msha_hub_t* hub_p, void* stream) {
    MSHA_REQUIRE(H >= 1 && H <= 32 && (H & (H - 1)) == 0 && D >= 4 && D % 4 == 0,
                 "gat_fwd_block: H must be a power of two <= 32 and D a multiple of 4");
    MSHA_REQUIRE(rowbeg && rowend && col && s_nbr && s_self && lse && feat && out, "gat_fwd_block: null argument");
    MSHA_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31), "gat_fwd_block: bad n_rows");
    int vw, vpl;
    MSHA_REQUIRE(pick_layout(H, D, &vw, &vpl) == 0 && vw == 4, "gat_fwd_block: unsupported channel count H*D=%d", H * D);
    if (n_rows == 0) return 0;
    HubArgs hub = make_hub(hub_p);
    DropArgs drop = make_drop(drop_p, drop_seed, 2u);
    cudaStream_t st = (cudaStream_t)stream;
    if (hub.n_hub > 0 && !accumulate) {
        zero_rows_kernel<<<hub.n_hub, 128, 0, st>>>(out, hub.hub_ids, hub.n_hub, H * D);
        MSHA_LAUNCH_OK();
    }
    const int wpc = vpl == 2 ? gat_wpc() : GAT_WARPS;
    const unsigned grid = (unsigned)msha_cdiv(n_rows + hub.n_segs, wpc);
    const size_t smem = (size_t)wpc * (32 * H + 32) * sizeof(float);
#define CALLPW(W, MB)                                                                                                      \
    gat_fwd_kernel<4, 2, true, W, MB, (W == 1 ? 2 : 4)><<<grid, W * 32, smem, st>>>(rowbeg, rowend, col, (int)n_rows, s_nbr, s_self, feat,    \
                                                                  H, D, slope, nullptr, alpha_out, out, 0, nullptr, drop,   \
                                                                  hub, nullptr, lse, accumulate)
#define CALLP(B)                                                                                                        \
    gat_fwd_kernel<4, B, true><<<grid, GAT_THREADS, smem, st>>>(rowbeg, rowend, col, (int)n_rows, s_nbr, s_self, feat, \
                                                                H, D, slope, nullptr, alpha_out, out, 0, nullptr, drop, \
                                                                hub, nullptr, lse, accumulate)
    if (vpl == 2 && wpc == 1) { CALLPW(1, 32); }
    else if (vpl == 2 && wpc == 2) { CALLPW(2, 8); }
    else if (vpl == 1) { CALLP(1); } else if (vpl == 2) { CALLP(2); } else { CALLP(4); }
#undef CALLP
#undef CALLPW
    MSHA_LAUNCH_OK();
    return 0;
}

// Row pass of the backward.  s_nbr == NULL -> no softmax: dlogit receives d(weights).
// hub (nullable) + r_buf (float[n_rows * H] scratch) handle rows split into segments.
MSHA_API int msha_gat_bwd_rows(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* s_nbr,
                               const float* s_self, float slope, const float* alpha, const float* feat,
                               const float* dout, const float* out, int act, float* dz_out, const float* dT,
                               const float* fT, const float* dalpha_extra, const float* dlse, int H, int D,
                               float* dlogit, float* ds_self, float drop_p, uint64_t drop_seed,
                               const msha_hub_t* hub_p, float* r_buf, int avg_degree_hint, void* stream) {
    MSHA_REQUIRE(H >= 1 && H <= 32 && D >= 1, "gat_bwd_rows: need 1 <= H <= 32, D >= 1");
    MSHA_REQUIRE((dT == nullptr) == (fT == nullptr), "gat_bwd_rows: dT and fT go together");
    MSHA_REQUIRE(act == 0 || out != nullptr, "gat_bwd_rows: activated output needed for ELU backward");
    MSHA_REQUIRE(s_nbr == nullptr || (s_self != nullptr && alpha != nullptr), "gat_bwd_rows: softmax mode needs s_self, alpha");
    int vw, vpl;
    MSHA_REQUIRE(pick_layout(H, D, &vw, &vpl) == 0, "gat_bwd_rows: unsupported channel count H*D=%d", H * D);
    if (n_rows == 0) return 0;
    HubArgs hub = make_hub(hub_p);
    MSHA_REQUIRE(hub.n_segs == 0 || r_buf != nullptr, "gat_bwd_rows: hub rows need r_buf");
    DropArgs drop = make_drop(drop_p, drop_seed, 2u);
    // 256 channels in 128-bit vectors, no second (dT, fT) term: compile-time lanes-per-head variants
    const int lph = (vw == 4 && vpl == 2 && dT == nullptr && H * D == 256) ? D / 4 : 0;
    const int wpc = (lph == 8 || lph == 64) ? gat_wpc() : GAT_WARPS;
    const int eb = gat_rows_eb();
    const size_t smem = (size_t)wpc * (32 * H + 2 * H) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (hub.n_segs > 0) {
        MSHA_CUDA(cudaMemsetAsync(r_buf, 0, (size_t)n_rows * H * sizeof(float), st));
        if (ds_self) {
            zero_rows_kernel<<<hub.n_hub, 32, 0, st>>>(ds_self, hub.hub_ids, hub.n_hub, H);
            MSHA_LAUNCH_OK();
        }
    }
    const bool dense_rows = avg_degree_hint >= 64;
#define CALL(A, B)                                                                                                       \
    if (dense_rows)                                                                                                      \
        gat_bwd_rows_kernel<A, B, 4, 2><<<grid, GAT_THREADS, smem, st>>>(rowptr, col, (int)n_rows, s_nbr, s_self, slope, \
                                                                         alpha, feat, dout, out, act, dz_out, dT, fT,    \
                                                                         dalpha_extra, dlse, H, D, dlogit, ds_self,      \
                                                                         drop, hub, mode, r_buf);                        \
    else                                                                                                                 \
        gat_bwd_rows_kernel<A, B, 2, 3><<<grid, GAT_THREADS, smem, st>>>(rowptr, col, (int)n_rows, s_nbr, s_self, slope, \
                                                                         alpha, feat, dout, out, act, dz_out, dT, fT,    \
                                                                         dalpha_extra, dlse, H, D, dlogit, ds_self,      \
                                                                         drop, hub, mode, r_buf)
#define CALL_FASTW(LPH, EBv, MB, W)                                                                                       \
    gat_bwd_rows_kernel<4, 2, EBv, MB, LPH, W><<<grid, W * 32, smem, st>>>(rowptr, col, (int)n_rows, s_nbr, s_self, slope, \
                                                                          alpha, feat, dout, out, act, dz_out, dT, fT,    \
                                                                          dalpha_extra, dlse, H, D, dlogit, ds_self,      \
                                                                          drop, hub, mode, r_buf)
#define CALL_FAST(LPH)                                                          \
    if (wpc == 1 && eb == 2) { CALL_FASTW(LPH, 2, 32, 1); }                     \
    else if (wpc == 1 && eb == 4) { CALL_FASTW(LPH, 4, 24, 1); }                \
    else if (wpc == 1) { CALL_FASTW(LPH, 8, 16, 1); }                           \
    else if (wpc == 2 && eb == 4) { CALL_FASTW(LPH, 4, 12, 2); }                \
    else if (wpc == 2) { CALL_FASTW(LPH, 8, 8, 2); }                            \
    else if (eb == 4) { CALL_FASTW(LPH, 4, 3, 8); }                             \
    else { CALL_FASTW(LPH, 8, 2, 8); }
#define LAUNCH_ROWS()                                   \
    if (lph == 8) { CALL_FAST(8) }                      \
    else if (lph == 64) { CALL_FAST(64) }               \
    else { DISPATCH_LAYOUT(vw, vpl, CALL) }
    unsigned grid = (unsigned)msha_cdiv(n_rows, wpc);
    int mode = 0;
    LAUNCH_ROWS()
    MSHA_LAUNCH_OK();
    if (hub.n_segs > 0) {
        grid = (unsigned)msha_cdiv(hub.n_segs, wpc);
        mode = 1;
        LAUNCH_ROWS()
        MSHA_LAUNCH_OK();
        if (s_nbr != nullptr) {
            mode = 2;
            LAUNCH_ROWS()
            MSHA_LAUNCH_OK();
        }
    }
#undef LAUNCH_ROWS
#undef CALL_FAST
#undef CALL_FASTW
#undef CALL
    return 0;
}

// out[j] (+)= sum_{i in col j} w[perm]*feat[i];  esum_out[j,h] = sum esum_in[perm,h].   w may be NULL (sums only).
// hub (nullable): columns with more than hub->seg_limit entries are processed as segments with atomic accumulation.
MSHA_API int msha_spmm_csc(const int32_t* colptr, const int32_t* rowidx, const int32_t* perm, int64_t n_cols,
                           const float* w, const float* feat, int H, int D, float* out, int accumulate,
                           const float* esum_in, float* esum_out, float drop_p, uint64_t drop_seed,
                           const msha_hub_t* hub_p, void* stream) {
    MSHA_REQUIRE(H >= 1 && H <= 32 && D >= 1, "spmm_csc: need 1 <= H <= 32, D >= 1");
    MSHA_REQUIRE(w == nullptr || (feat != nullptr && out != nullptr), "spmm_csc: feat/out required with weights");
    MSHA_REQUIRE((esum_in == nullptr) == (esum_out == nullptr), "spmm_csc: esum_in/esum_out go together");
    int vw, vpl;
    MSHA_REQUIRE(pick_layout(H, D, &vw, &vpl) == 0, "spmm_csc: unsupported channel count H*D=%d", H * D);
    if (n_cols == 0) return 0;
    HubArgs hub = make_hub(hub_p);
    DropArgs drop = make_drop(drop_p, drop_seed, 2u);
    const int wpc = (vw == 4 && vpl == 2) ? gat_wpc() : GAT_WARPS;
    const unsigned grid = (unsigned)msha_cdiv(n_cols + hub.n_segs, wpc);
    const size_t smem = (size_t)wpc * (32 * H + H) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (hub.n_hub > 0) {
        if (w != nullptr && !accumulate) {
            zero_rows_kernel<<<hub.n_hub, 128, 0, st>>>(out, hub.hub_ids, hub.n_hub, H * D);
            MSHA_LAUNCH_OK();
        }
        if (esum_out != nullptr) {
            zero_rows_kernel<<<hub.n_hub, 32, 0, st>>>(esum_out, hub.hub_ids, hub.n_hub, H);
            MSHA_LAUNCH_OK();
        }
    }
#define CALL(A, B)                                                                                             \
    spmm_csc_kernel<A, B><<<grid, GAT_THREADS, smem, st>>>(colptr, rowidx, perm, (int)n_cols, w, feat, H, D,   \
                                                           out, accumulate, esum_in, esum_out, drop, hub)
#define CALLWU(W, MB, U)                                                                                           \
    spmm_csc_kernel<4, 2, W, MB, U><<<grid, W * 32, smem, st>>>(colptr, rowidx, perm, (int)n_cols, w, feat, H, D,  \
                                                                out, accumulate, esum_in, esum_out, drop, hub)
#define CALLW(W, MB) CALLWU(W, MB, 4)
    if (wpc == 1 && gat_csc_var() == 1) { CALLWU(1, 24, 4); }
    else if (wpc == 1 && gat_csc_var() == 2) { CALLWU(1, 16, 8); }
    else if (wpc == 1 && gat_csc_var() == 3) { CALLWU(1, 32, 2); }
    else if (wpc == 1) { CALLW(1, 16); }
    else if (wpc == 2) { CALLW(2, 8); }
    else { DISPATCH_LAYOUT(vw, vpl, CALL) }
#undef CALLW
#undef CALLWU
#undef CALL
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// node-level helpers
// ---------------------------------------------------------------------------------------------
// s1[n,h] = sum_d feat[n,h,d]*a1[h,d] ; s2 likewise (a2/s2 optional)
__global__ void node_scores_kernel(const float* __restrict__ feat, int64_t n, int H, int D,
                                   const float* __restrict__ a1, float* __restrict__ s1,
                                   const float* __restrict__ a2, float* __restrict__ s2) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* f = feat + row * (int64_t)H * D;
    for (int h = 0; h < H; ++h) {
        float p1 = 0.f, p2 = 0.f;
        for (int d = lane; d < D; d += 32) {
            float x = f[h * D + d];
            p1 = fmaf(x, a1[h * D + d], p1);
            if (a2) p2 = fmaf(x, a2[h * D + d], p2);
        }
        p1 = warp_sum(p1);
        if (a2) p2 = warp_sum(p2);
        if (lane == 0) {
            s1[row * H + h] = p1;
            if (a2) s2[row * H + h] = p2;
        }
    }
}

MSHA_API int msha_node_scores(const float* feat, int64_t n, int H, int D, const float* a1, float* s1,
                              const float* a2, float* s2, void* stream) {
    MSHA_REQUIRE(H >= 1 && D >= 1 && n >= 0, "node_scores: bad shape");
    if (n == 0) return 0;
    node_scores_kernel<<<(unsigned)msha_cdiv(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(feat, n, H, D, a1, s1, a2, s2);
    MSHA_LAUNCH_OK();
    return 0;
}

// out[n,h,d] (+)= s1[n,h]*a1[h,d] + s2[n,h]*a2[h,d]
__global__ void node_outer_kernel(float* __restrict__ out, int64_t n, int H, int D, const float* __restrict__ s1,
                                  const float* __restrict__ a1, const float* __restrict__ s2,
                                  const float* __restrict__ a2, int accumulate) {
    const int C = H * D;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * C) return;
    const int64_t row = idx / C;
    const int ch = (int)(idx - row * C);
    const int h = ch / D;
    float v = s1[row * H + h] * a1[ch];
    if (s2) v = fmaf(s2[row * H + h], a2[ch], v);
    out[idx] = accumulate ? out[idx] + v : v;
}

MSHA_API int msha_node_outer_add(float* out, int64_t n, int H, int D, const float* s1, const float* a1,
                                 const float* s2, const float* a2, int accumulate, void* stream) {
    MSHA_REQUIRE(H >= 1 && D >= 1 && n >= 0, "node_outer_add: bad shape");
    MSHA_REQUIRE((s2 == nullptr) == (a2 == nullptr), "node_outer_add: s2/a2 go together");
    if (n == 0) return 0;
    node_outer_kernel<<<(unsigned)msha_cdiv(n * H * D, 256), 256, 0, (cudaStream_t)stream>>>(out, n, H, D, s1, a1, s2, a2,
                                                                                          accumulate);
    MSHA_LAUNCH_OK();
    return 0;
}

// Column reduction over nodes:  out[c] = sum_n x[n,c] * (y ? y[n,c] : 1) * (s ? s[n, c/D] : 1)
// two stages through `partial` ([COLRED_BLOCKS, C] doubles) for a deterministic result.
constexpr int COLRED_BLOCKS = 296;
__global__ void colreduce_stage1(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ s,
                                 int64_t n, int C, int D, double* __restrict__ partial) {
    const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * rows_per;
    const int64_t r1 = r0 + rows_per < n ? r0 + rows_per : n;
    const int H = C / D;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double acc = 0.0;
        const int h = c / D;
        for (int64_t r = r0; r < r1; ++r) {
            float v = x[r * C + c];
            if (y) v *= y[r * C + c];
            if (s) v *= s[r * H + h];
            acc += (double)v;
        }
        partial[(int64_t)blockIdx.x * C + c] = acc;
    }
}
__global__ void colreduce_stage2(const double* __restrict__ partial, int nblocks, int C, float* __restrict__ out,
                                 double* __restrict__ out_d) {
    // 32 columns per CTA, 8 warps each summing every 8th partial (a single thread per column made this a chain of
    // ~300 dependent fp64 adds: 32 us for a 600 KB reduction); fixed order -> still deterministic
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
        if (out) out[c] = (float)t;
        if (out_d) out_d[c] = t;
    }
}

MSHA_API size_t msha_colreduce_workspace_bytes(int C) { return (size_t)COLRED_BLOCKS * C * sizeof(double); }

MSHA_API int msha_colreduce(const float* x, const float* y, const float* s, int64_t n, int C, int D, float* out,
                            void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(C >= 1 && D >= 1 && C % D == 0 && n >= 0, "colreduce: bad shape");
    MSHA_REQUIRE(ws_bytes >= msha_colreduce_workspace_bytes(C), "colreduce: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (int)(n < COLRED_BLOCKS ? (n > 0 ? n : 1) : COLRED_BLOCKS);
    colreduce_stage1<<<nb, 256, 0, st>>>(x, y, s, n, C, D, (double*)ws);
    MSHA_LAUNCH_OK();
    colreduce_stage2<<<(unsigned)msha_cdiv(C, 32), 256, 0, st>>>((const double*)ws, nb, C, out, nullptr);
    MSHA_LAUNCH_OK();
    return 0;
}
