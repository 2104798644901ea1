// Peer-memory primitives of the node-partitioned path: every rank maps the exchange buffers of
// its peers (NVLink 5 / NVSwitch peer access) and the halo exchange is done by this library's own kernels and by
// copy-engine pulls ordered with the flags below -- no collective library call sits between the attention kernels.
//
//   flags     uint32[n_channels][world] per rank, in peer-mapped memory.  flags[ch][q] on rank r is written only by
//             rank q (msha_peer_signal) and read only by rank r (msha_peer_wait); values are monotonically increasing
//             sequence numbers, so nothing is ever reset.
//   signal    release at system scope: everything this rank's earlier work on the stream wrote (kernel boundary) is
//             visible to a peer that has observed the flag.
//   wait      one thread per awaited peer spins with ld.acquire.sys; a bounded spin (timeout in ns of %globaltimer)
//             sets a status word and returns instead of hanging the GPU when a peer died (PeerGroup.check raises).
//   pull      gathers row blocks (or listed halo rows) out of the peers' buffers with 128-bit P2P loads.
//   sum       out = sum over a table of pointers (local staging slots or peer buffers) in table order: the
//             reduce-scatter of d Wh / d s_nbr; fixed order -> run-to-run deterministic.
//
// The reference has no distributed code (train.py:18 is single-device); this is new functionality behind the kernels
// of gat_kernels.cu.
#include "common.cuh"
#include "msha_b200.h"
#include <stdlib.h>

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// mode 0: st.release.sys (the PTX-model way; its system-scope fence was measured at up to 0.4 ms under copy-engine traffic)
// mode 1: fence.acq_rel.gpu + st.relaxed.sys -- the data the flag publishes was written by EARLIER KERNELS of this stream
//         (complete, device-visible at the kernel boundary) and is read by the peers through this GPU's L2, the point of
//         coherence for peer accesses; only the flag itself crosses NVLink
// mode 2: st.relaxed.sys alone
__global__ void peer_signal_kernel(const uint64_t* __restrict__ flag_tab, int world, int rank, int channel,
                                   uint32_t peer_mask, const uint32_t* __restrict__ epoch, uint32_t value, int mode) {
    const int q = threadIdx.x;
    if (q >= world || !((peer_mask >> q) & 1u)) return;
    const uint32_t v = value + (epoch ? *epoch : 0u);
    uint32_t* f = reinterpret_cast<uint32_t*>(flag_tab[q]) + (int64_t)channel * world + rank;
    if (mode == 0) {
        st_release_sys(f, v);
    } else {
        if (mode == 1) __threadfence();
        st_relaxed_sys(f, v);
    }
}

__global__ void peer_wait_kernel(const uint32_t* __restrict__ flags, int world, int channel, uint32_t peer_mask,
                                 const uint32_t* __restrict__ epoch, uint32_t value, uint64_t timeout_ns,
                                 int32_t* __restrict__ status) {
    const int q = threadIdx.x;
    if (q >= world || !((peer_mask >> q) & 1u)) return;
    const uint32_t want = value + (epoch ? *epoch : 0u);
    const uint32_t* f = flags + (int64_t)channel * world + q;
    const uint64_t t0 = globaltimer_ns();
    // sequence numbers compare as signed differences (wrap-safe)
    while ((int32_t)(ld_acquire_sys(f) - want) < 0) {
        __nanosleep(64);
        if (timeout_ns && globaltimer_ns() - t0 > timeout_ns) {
            if (status) atomicExch(status, 0x100 | q);      // a peer never arrived: give up instead of hanging the GPU;
            return;                                          // the host raises when it reads the status word
        }
    }
}

// flag_tab: device array uint64[world], entry q = address of rank q's flag array as mapped in THIS process.
MSHA_API int msha_peer_signal(const uint64_t* flag_tab, int world, int rank, int channel, uint32_t peer_mask,
                              const uint32_t* epoch, uint32_t value, void* stream) {
    MSHA_REQUIRE(flag_tab != nullptr && world >= 1 && world <= 32 && rank >= 0 && rank < world && channel >= 0,
                 "peer_signal: bad arguments");
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("MSHA_PEER_SIGNAL_MODE");
        mode = e ? atoi(e) : 1;
        if (mode < 0 || mode > 2) mode = 1;
    }
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flag_tab, world, rank, channel, peer_mask, epoch, value, mode);
    MSHA_LAUNCH_OK();
    return 0;
}

MSHA_API int msha_peer_wait(const uint32_t* flags, int world, int channel, uint32_t peer_mask, const uint32_t* epoch,
                            uint32_t value, uint64_t timeout_ns, int32_t* status, void* stream) {
    MSHA_REQUIRE(flags != nullptr && world >= 1 && world <= 32 && channel >= 0, "peer_wait: bad arguments");
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, world, channel, peer_mask, epoch, value, timeout_ns, status);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// pull: dst[blk q] = peer q's buffer[blk q] for every q != rank (blocks of `block_bytes` at q * block_bytes), or only
// the listed rows of every block (halo lists: row_ptr[q] .. row_ptr[q+1] index row_ids, ids are rows of the gathered
// buffer).  128-bit loads over NVLink, many in flight per thread.
// ---------------------------------------------------------------------------------------------
constexpr int PULL_THREADS = 256;
constexpr int PULL_UNROLL = 4;

__global__ void __launch_bounds__(PULL_THREADS)
peer_pull_blocks_kernel(uint4* __restrict__ dst, const uint64_t* __restrict__ src_tab, int world, int rank,
                        int64_t block_stride_vec, int64_t block_vec) {
    // grid.y enumerates the W-1 remote blocks starting after the own one (staggered: no two ranks start on the same peer)
    const int q = (rank + 1 + blockIdx.y) % world;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(src_tab[q]) + (int64_t)q * block_stride_vec;
    uint4* __restrict__ d = dst + (int64_t)q * block_stride_vec;
    const int64_t stride = (int64_t)gridDim.x * PULL_THREADS;
    int64_t i = (int64_t)blockIdx.x * PULL_THREADS + threadIdx.x;
    for (; i + (PULL_UNROLL - 1) * stride < block_vec; i += PULL_UNROLL * stride) {
        uint4 v[PULL_UNROLL];
#pragma unroll
        for (int u = 0; u < PULL_UNROLL; ++u) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < PULL_UNROLL; ++u) d[i + u * stride] = v[u];
    }
    for (; i < block_vec; i += stride) d[i] = src[i];
}

__global__ void __launch_bounds__(PULL_THREADS)
peer_pull_rows_kernel(float* __restrict__ dst, const uint64_t* __restrict__ src_tab, int world, int rank,
                      const int32_t* __restrict__ row_ids, const int64_t* __restrict__ row_ptr, int row_vec) {
    const int q = (rank + 1 + blockIdx.y) % world;
    const float4* __restrict__ src = reinterpret_cast<const float4*>(src_tab[q]);
    float4* __restrict__ d = reinterpret_cast<float4*>(dst);
    const int64_t r0 = row_ptr[q], r1 = row_ptr[q + 1];
    const int64_t total = (r1 - r0) * row_vec;
    const int64_t stride = (int64_t)gridDim.x * PULL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * PULL_THREADS + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / row_vec;
        const int v = (int)(i - r * row_vec);
        const int64_t row = row_ids[r0 + r];
        d[row * row_vec + v] = src[row * row_vec + v];
    }
}

// dst / every src_tab[q]: a [world * block] buffer of `block_bytes`-sized rank blocks (the gathered layout); the own
// block is left alone.  The first `nbytes` of every remote block are copied (nbytes < block_bytes: padded blocks).
MSHA_API int msha_peer_pull_blocks(void* dst, const uint64_t* src_tab, int world, int rank, int64_t block_bytes,
                                   int64_t nbytes, int max_ctas, void* stream) {
    MSHA_REQUIRE(dst != nullptr && src_tab != nullptr && world >= 1 && rank >= 0 && rank < world, "peer_pull_blocks: bad arguments");
    MSHA_REQUIRE(block_bytes % 16 == 0 && nbytes % 16 == 0 && ((uintptr_t)dst & 15) == 0,
                 "peer_pull_blocks: 16-byte granularity required");
    MSHA_REQUIRE(nbytes >= 0 && nbytes <= block_bytes, "peer_pull_blocks: range outside the block");
    if (world == 1 || nbytes == 0) return 0;
    const int64_t nvec = nbytes / 16;
    int64_t ctas = msha_cdiv(nvec, (int64_t)PULL_THREADS * PULL_UNROLL);
    const int64_t cap = max_ctas > 0 ? max_ctas : 4 * MSHA_NUM_SMS;
    const int64_t per_peer = ctas < 1 ? 1 : (ctas > cap / (world - 1) + 1 ? cap / (world - 1) + 1 : ctas);
    dim3 grid((unsigned)per_peer, (unsigned)(world - 1));
    peer_pull_blocks_kernel<<<grid, PULL_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint4*>(dst), src_tab, world, rank, block_bytes / 16, nvec);
    MSHA_LAUNCH_OK();
    return 0;
}

MSHA_API int msha_peer_pull_rows(float* dst, const uint64_t* src_tab, int world, int rank, const int32_t* row_ids,
                                 const int64_t* row_ptr, int64_t max_rows_per_peer, int64_t C, int max_ctas, void* stream) {
    MSHA_REQUIRE(dst != nullptr && src_tab != nullptr && row_ids != nullptr && row_ptr != nullptr, "peer_pull_rows: bad arguments");
    MSHA_REQUIRE(C % 4 == 0 && C > 0 && ((uintptr_t)dst & 15) == 0, "peer_pull_rows: rows must be whole 128-bit vectors");
    if (world == 1 || max_rows_per_peer == 0) return 0;
    const int64_t total = max_rows_per_peer * (C / 4);
    const int64_t cap = max_ctas > 0 ? max_ctas : 4 * MSHA_NUM_SMS;
    int64_t per_peer = msha_cdiv(total, PULL_THREADS * 4);
    if (per_peer > cap / (world - 1) + 1) per_peer = cap / (world - 1) + 1;
    dim3 grid((unsigned)per_peer, (unsigned)(world - 1));
    peer_pull_rows_kernel<<<grid, PULL_THREADS, 0, (cudaStream_t)stream>>>(dst, src_tab, world, rank, row_ids, row_ptr, (int)(C / 4));
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// sum: out[i] = sum_k src_tab[k][offset + i], k in table order (the reduce-scatter's reduction; the table mixes local
// staging slots filled by copy-engine pulls and, for small messages, the peers' buffers themselves)
// ---------------------------------------------------------------------------------------------
constexpr int SUM_MAX_SRC = 32;
struct SumTab { const float4* p[SUM_MAX_SRC]; };

template <int NSRC>
__global__ void __launch_bounds__(256)
peer_sum_kernel(float4* __restrict__ out, SumTab tab, int n_src, int64_t nvec) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
        if (NSRC > 0) {
            float4 v[NSRC > 0 ? NSRC : 1];
#pragma unroll
            for (int k = 0; k < NSRC; ++k) v[k] = tab.p[k][i];
            float4 a = v[0];
#pragma unroll
            for (int k = 1; k < NSRC; ++k) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
            out[i] = a;
        } else {
            float4 a = tab.p[0][i];
            for (int k = 1; k < n_src; ++k) {
                const float4 b = tab.p[k][i];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            out[i] = a;
        }
    }
}

// src_ptrs: HOST array of n_src device addresses (each already offset to the first element to sum)
MSHA_API int msha_peer_sum(float* out, const uint64_t* src_ptrs, int n_src, int64_t n, int max_ctas, void* stream) {
    MSHA_REQUIRE(out != nullptr && src_ptrs != nullptr && n_src >= 1 && n_src <= SUM_MAX_SRC, "peer_sum: bad arguments");
    MSHA_REQUIRE(n % 4 == 0 && ((uintptr_t)out & 15) == 0, "peer_sum: 128-bit granularity required");
    if (n == 0) return 0;
    SumTab tab;
    for (int k = 0; k < n_src; ++k) {
        MSHA_REQUIRE((src_ptrs[k] & 15) == 0, "peer_sum: source %d is not 16-byte aligned", k);
        tab.p[k] = reinterpret_cast<const float4*>(src_ptrs[k]);
    }
    const int64_t nvec = n / 4;
    int64_t ctas = msha_cdiv(nvec, 256);
    const int64_t cap = max_ctas > 0 ? max_ctas : 8 * MSHA_NUM_SMS;
    if (ctas > cap) ctas = cap;
    cudaStream_t st = (cudaStream_t)stream;
    float4* o = reinterpret_cast<float4*>(out);
    switch (n_src) {
        case 1: peer_sum_kernel<1><<<(unsigned)ctas, 256, 0, st>>>(o, tab, n_src, nvec); break;
        case 2: peer_sum_kernel<2><<<(unsigned)ctas, 256, 0, st>>>(o, tab, n_src, nvec); break;
        case 3: peer_sum_kernel<3><<<(unsigned)ctas, 256, 0, st>>>(o, tab, n_src, nvec); break;
        case 4: peer_sum_kernel<4><<<(unsigned)ctas, 256, 0, st>>>(o, tab, n_src, nvec); break;
        case 8: peer_sum_kernel<8><<<(unsigned)ctas, 256, 0, st>>>(o, tab, n_src, nvec); break;
        default: peer_sum_kernel<0><<<(unsigned)ctas, 256, 0, st>>>(o, tab, n_src, nvec); break;
    }
    MSHA_LAUNCH_OK();
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Fused exchanges for the launch-latency regime (small blocks): the flag waits and the completion signal live inside
// the data kernel, so one gather (or one reduce-scatter) of up to two buffers is ONE launch after the producer's
// msha_peer_signal -- instead of wait / pull / pull / signal launches and a separate reuse-guard wait.
//   * every CTA waits only for the peer whose block it copies (gather) / for all peers (sum);
//   * the last CTA to finish (device counter) additionally waits for `guard` flags -- the peers' completion flags of the
//     OPPOSITE direction's previous exchange, which must hold before this rank overwrites what they read; checking them
//     here, long after they were set, replaces a dedicated wait launch -- and then publishes `done` to every peer.
// ---------------------------------------------------------------------------------------------
struct PeerFlagOps {
    const uint32_t* flags;        // this rank's flag array
    const uint64_t* flag_tab;     // device table of every rank's flag array
    int world, rank;
    int wait_ch;  uint32_t wait_val;      // data ready        (wait_ch < 0: none)
    int guard_ch; uint32_t guard_val;     // reuse guard       (guard_ch < 0: none)
    int done_ch;  uint32_t done_val;      // completion signal (done_ch < 0: none)
    uint64_t timeout_ns;
    int32_t* status;
    unsigned int* counter;        // zero before the launch; left zero by it
};

__device__ __forceinline__ bool spin_flag(const uint32_t* f, uint32_t want, uint64_t timeout_ns, int32_t* status, int q) {
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(f) - want) < 0) {
        __nanosleep(32);
        if (timeout_ns && globaltimer_ns() - t0 > timeout_ns) {
            if (status) atomicExch(status, 0x100 | q);
            return false;
        }
    }
    return true;
}

// call by all threads of the CTA after its data work: the last CTA of the grid checks the guard and signals completion
__device__ __forceinline__ void peer_finish(const PeerFlagOps& fo, unsigned int total_ctas) {
    __shared__ unsigned int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(fo.counter, 1u) == total_ctas - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    const int q = threadIdx.x;
    if (q < fo.world && q != fo.rank) {
        if (fo.guard_ch >= 0) spin_flag(fo.flags + (int64_t)fo.guard_ch * fo.world + q, fo.guard_val, fo.timeout_ns, fo.status, q);
        if (fo.done_ch >= 0) {
            __threadfence();                              // the CTAs' stores (counted in through the atomic) before the flag
            st_relaxed_sys(reinterpret_cast<uint32_t*>(fo.flag_tab[q]) + (int64_t)fo.done_ch * fo.world + fo.rank, fo.done_val);
        }
    }
    if (threadIdx.x == 0) *fo.counter = 0u;
}

struct PullBufs { uint4* dst[2]; const uint64_t* tab[2]; int64_t stride_vec[2]; int64_t nvec[2]; };

__global__ void __launch_bounds__(PULL_THREADS)
peer_exchange_pull_kernel(PullBufs b, PeerFlagOps fo) {
    const int q = (fo.rank + 1 + blockIdx.y) % fo.world;
    const int k = blockIdx.z;
    __shared__ int s_ok;
    if (threadIdx.x == 0)
        s_ok = fo.wait_ch < 0 ? 1 : (int)spin_flag(fo.flags + (int64_t)fo.wait_ch * fo.world + q, fo.wait_val, fo.timeout_ns, fo.status, q);
    __syncthreads();
    if (s_ok) {
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(b.tab[k][q]) + (int64_t)q * b.stride_vec[k];
        uint4* __restrict__ d = b.dst[k] + (int64_t)q * b.stride_vec[k];
        const int64_t n = b.nvec[k];
        const int64_t stride = (int64_t)gridDim.x * PULL_THREADS;
        int64_t i = (int64_t)blockIdx.x * PULL_THREADS + threadIdx.x;
        for (; i + (PULL_UNROLL - 1) * stride < n; i += PULL_UNROLL * stride) {
            uint4 v[PULL_UNROLL];
#pragma unroll
            for (int u = 0; u < PULL_UNROLL; ++u) v[u] = src[i + u * stride];
#pragma unroll
            for (int u = 0; u < PULL_UNROLL; ++u) d[i + u * stride] = v[u];
        }
        for (; i < n; i += stride) d[i] = src[i];
    }
    peer_finish(fo, gridDim.x * gridDim.y * gridDim.z);
}

struct SumBufs { float4* out[2]; const uint64_t* tab[2]; int64_t offset_vec[2]; int64_t nvec[2]; };

__global__ void __launch_bounds__(256)
peer_exchange_sum_kernel(SumBufs b, PeerFlagOps fo) {
    const int k = blockIdx.y;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (fo.wait_ch >= 0 && threadIdx.x < fo.world && threadIdx.x != fo.rank) {
        if (!spin_flag(fo.flags + (int64_t)fo.wait_ch * fo.world + threadIdx.x, fo.wait_val, fo.timeout_ns, fo.status, threadIdx.x))
            s_ok = 0;
    }
    __syncthreads();
    if (s_ok) {
        const int64_t n = b.nvec[k];
        const int64_t stride = (int64_t)gridDim.x * 256;
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < fo.world; ++q) {                       // rank order: deterministic
                const float4 v = (reinterpret_cast<const float4*>(b.tab[k][q]) + b.offset_vec[k])[i];
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            b.out[k][i] = a;
        }
    }
    peer_finish(fo, gridDim.x * gridDim.y);
}

static int fill_flag_ops(PeerFlagOps& fo, const uint32_t* flags, const uint64_t* flag_tab, int world, int rank, int wait_ch,
                         uint32_t wait_val, int guard_ch, uint32_t guard_val, int done_ch, uint32_t done_val,
                         uint64_t timeout_ns, int32_t* status, uint32_t* counter) {
    MSHA_REQUIRE(flags && flag_tab && counter && world >= 1 && world <= 32 && rank >= 0 && rank < world, "peer_exchange: bad arguments");
    fo.flags = flags; fo.flag_tab = flag_tab; fo.world = world; fo.rank = rank;
    fo.wait_ch = wait_ch; fo.wait_val = wait_val; fo.guard_ch = guard_ch; fo.guard_val = guard_val;
    fo.done_ch = done_ch; fo.done_val = done_val; fo.timeout_ns = timeout_ns; fo.status = status; fo.counter = counter;
    return 0;
}

// Gather of n_bufs (1 or 2) buffers in the gathered layout [world][block_bytes[k]]: the first nbytes[k] of every remote
// block, each CTA after the owner's `wait` flag; `guard` / `done` as described above (channel < 0: skip).
MSHA_API int msha_peer_exchange_pull(int n_bufs, void* const* dst, const uint64_t* const* src_tab, const int64_t* block_bytes,
                                     const int64_t* nbytes, const uint32_t* flags, const uint64_t* flag_tab, int world,
                                     int rank, int wait_ch, uint32_t wait_val, int guard_ch, uint32_t guard_val, int done_ch,
                                     uint32_t done_val, uint64_t timeout_ns, int32_t* status, uint32_t* counter,
                                     int max_ctas, void* stream) {
    MSHA_REQUIRE(n_bufs == 1 || n_bufs == 2, "peer_exchange_pull: 1 or 2 buffers");
    PeerFlagOps fo;
    if (fill_flag_ops(fo, flags, flag_tab, world, rank, wait_ch, wait_val, guard_ch, guard_val, done_ch, done_val, timeout_ns,
                      status, counter)) return -1;
    if (world == 1) return 0;
    PullBufs b;
    int64_t max_vec = 0;
    for (int k = 0; k < 2; ++k) {
        const int kk = k < n_bufs ? k : 0;
        MSHA_REQUIRE(block_bytes[kk] % 16 == 0 && nbytes[kk] % 16 == 0 && nbytes[kk] <= block_bytes[kk] && ((uintptr_t)dst[kk] & 15) == 0,
                     "peer_exchange_pull: 16-byte granularity required");
        b.dst[k] = reinterpret_cast<uint4*>(dst[kk]); b.tab[k] = src_tab[kk];
        b.stride_vec[k] = block_bytes[kk] / 16; b.nvec[k] = nbytes[kk] / 16;
        if (b.nvec[k] > max_vec) max_vec = b.nvec[k];
    }
    const int64_t cap = max_ctas > 0 ? max_ctas : 4 * MSHA_NUM_SMS;
    int64_t per = msha_cdiv(max_vec, (int64_t)PULL_THREADS * PULL_UNROLL);
    const int64_t lim = cap / ((world - 1) * n_bufs) + 1;
    if (per > lim) per = lim;
    if (per < 1) per = 1;
    dim3 grid((unsigned)per, (unsigned)(world - 1), (unsigned)n_bufs);
    peer_exchange_pull_kernel<<<grid, PULL_THREADS, 0, (cudaStream_t)stream>>>(b, fo);
    MSHA_LAUNCH_OK();
    return 0;
}

// Reduce-scatter tail of n_bufs (1 or 2) buffers: out[k][i] = sum over ranks q (rank order) of rank q's buffer k at
// offset_bytes[k] + i -- read in place over NVLink after every peer's `wait` flag; `guard` / `done` as above.
MSHA_API int msha_peer_exchange_sum(int n_bufs, float* const* out, const uint64_t* const* src_tab, const int64_t* offset_bytes,
                                    const int64_t* n_floats, const uint32_t* flags, const uint64_t* flag_tab, int world,
                                    int rank, int wait_ch, uint32_t wait_val, int guard_ch, uint32_t guard_val, int done_ch,
                                    uint32_t done_val, uint64_t timeout_ns, int32_t* status, uint32_t* counter, int max_ctas,
                                    void* stream) {
    MSHA_REQUIRE(n_bufs == 1 || n_bufs == 2, "peer_exchange_sum: 1 or 2 buffers");
    PeerFlagOps fo;
    if (fill_flag_ops(fo, flags, flag_tab, world, rank, wait_ch, wait_val, guard_ch, guard_val, done_ch, done_val, timeout_ns,
                      status, counter)) return -1;
    SumBufs b;
    int64_t max_vec = 0;
    for (int k = 0; k < 2; ++k) {
        const int kk = k < n_bufs ? k : 0;
        MSHA_REQUIRE(offset_bytes[kk] % 16 == 0 && n_floats[kk] % 4 == 0 && ((uintptr_t)out[kk] & 15) == 0,
                     "peer_exchange_sum: 128-bit granularity required");
        b.out[k] = reinterpret_cast<float4*>(out[kk]); b.tab[k] = src_tab[kk];
        b.offset_vec[k] = offset_bytes[kk] / 16; b.nvec[k] = n_floats[kk] / 4;
        if (b.nvec[k] > max_vec) max_vec = b.nvec[k];
    }
    const int64_t cap = max_ctas > 0 ? max_ctas : 4 * MSHA_NUM_SMS;
    int64_t per = msha_cdiv(max_vec, 256);
    if (per > cap / n_bufs + 1) per = cap / n_bufs + 1;
    if (per < 1) per = 1;
    dim3 grid((unsigned)per, (unsigned)n_bufs);
    peer_exchange_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(b, fo);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// all-reduce (sum) of a SMALL fp64 vector: the BatchNorm column statistics of a partitioned node axis (2*C doubles,
// SURVEY.md section 8e) keep their fp64 accumulation across ranks.  One CTA; every rank reads every rank's slot in
// rank order (identical result everywhere) after its flag.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* __restrict__ out, const uint64_t* __restrict__ tab, int64_t offset_bytes, int n,
                          PeerFlagOps fo) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (fo.wait_ch >= 0 && threadIdx.x < fo.world && threadIdx.x != fo.rank) {
        if (!spin_flag(fo.flags + (int64_t)fo.wait_ch * fo.world + threadIdx.x, fo.wait_val, fo.timeout_ns, fo.status, threadIdx.x))
            s_ok = 0;
    }
    __syncthreads();
    if (!s_ok) return;
    for (int i = threadIdx.x; i < n; i += 256) {
        double a = 0.0;
        for (int q = 0; q < fo.world; ++q)
            a += reinterpret_cast<const double*>(tab[q] + offset_bytes)[i];
        out[i] = a;
    }
}

MSHA_API int msha_peer_allreduce_f64(double* out, const uint64_t* src_tab, int64_t offset_bytes, int64_t n,
                                     const uint32_t* flags, const uint64_t* flag_tab, int world, int rank, int wait_ch,
                                     uint32_t wait_val, uint64_t timeout_ns, int32_t* status, void* stream) {
    MSHA_REQUIRE(out && src_tab && n >= 0 && n < (1 << 24) && offset_bytes % 8 == 0, "peer_allreduce_f64: bad arguments");
    PeerFlagOps fo;
    uint32_t dummy = 0;
    if (fill_flag_ops(fo, flags, flag_tab, world, rank, wait_ch, wait_val, -1, 0, -1, 0, timeout_ns, status, &dummy)) return -1;
    fo.counter = nullptr;
    if (n == 0) return 0;
    peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(out, src_tab, offset_bytes, (int)n, fo);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Halo exchange: a rank needs only the feature rows its own edges reference (on the 100 M-edge R-MAT graph 38 % of the
// remote rows at 8 GPUs: a third of the nodes has no in-edge at all).  Columns are renumbered per rank to
// [own block | halo rows of peer 0 | halo rows of peer 1 | ...] (dist_p2p.HaloPlan), so
//   gather_rows        packs the listed rows of every peer's own block straight into this rank's halo segments
//                      (128-bit P2P loads, one launch for all peers and one row chunk);
//   scatter_add_rows   is its adjoint: the owner adds the peers' halo-segment gradients onto the listed rows of its block.
// seg arrays are DEVICE int64[world] (entry of the own rank ignored):
//   list_beg/list_end  range of `lists` (row ids inside the peer's block) handled by this launch, per peer
//   local_off          row of this rank's compact buffer where entry list_first[q] of peer q's segment lives
// ---------------------------------------------------------------------------------------------
template <typename VecT>
__global__ void __launch_bounds__(PULL_THREADS)
peer_gather_rows_kernel(VecT* __restrict__ dst, const uint64_t* __restrict__ src_tab, int world, int rank,
                        const int32_t* __restrict__ lists, const int64_t* __restrict__ list_beg,
                        const int64_t* __restrict__ list_end, const int64_t* __restrict__ list_first,
                        const int64_t* __restrict__ local_off, int row_vec) {
    const int q = (rank + 1 + blockIdx.y) % world;
    const VecT* __restrict__ src = reinterpret_cast<const VecT*>(src_tab[q]);
    const int64_t b = list_beg[q], e = list_end[q];
    const int64_t total = (e - b) * row_vec;
    const int64_t base = local_off[q] - list_first[q];
    const int64_t stride = (int64_t)gridDim.x * PULL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * PULL_THREADS + threadIdx.x; i < total; i += stride) {
        const int64_t k = b + i / row_vec;
        const int v = (int)(i % row_vec);
        dst[(base + k) * row_vec + v] = src[(int64_t)lists[k] * row_vec + v];
    }
}

__device__ __forceinline__ void vec_atomic_add(float4* p, float4 x) { atomicAdd(p, x); }
__device__ __forceinline__ void vec_atomic_add(float2* p, float2 x) { atomicAdd(p, x); }
__device__ __forceinline__ void vec_atomic_add(float* p, float x) { atomicAdd(p, x); }

template <typename VecT>
__global__ void __launch_bounds__(PULL_THREADS)
peer_scatter_add_rows_kernel(VecT* __restrict__ dst, const uint64_t* __restrict__ src_tab, int world, int rank,
                             const int32_t* __restrict__ lists, const int64_t* __restrict__ list_ptr,
                             const int64_t* __restrict__ remote_off, int row_vec) {
    const int q = (rank + 1 + blockIdx.y) % world;
    const VecT* __restrict__ src = reinterpret_cast<const VecT*>(src_tab[q]) + remote_off[q] * row_vec;
    const int64_t b = list_ptr[q], e = list_ptr[q + 1];
    const int64_t total = (e - b) * row_vec;
    const int64_t stride = (int64_t)gridDim.x * PULL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * PULL_THREADS + threadIdx.x; i < total; i += stride) {
        const int64_t k = i / row_vec;
        const int v = (int)(i % row_vec);
        vec_atomic_add(dst + (int64_t)lists[b + k] * row_vec + v, src[k * row_vec + v]);     // rows of different peers collide
    }
}

static unsigned rows_grid(int64_t max_rows, int row_vec, int world, int max_ctas) {
    const int64_t cap = max_ctas > 0 ? max_ctas : 4 * MSHA_NUM_SMS;
    int64_t per = msha_cdiv(max_rows * row_vec, (int64_t)PULL_THREADS * 4);
    const int64_t lim = cap / (world - 1) + 1;
    if (per > lim) per = lim;
    return (unsigned)(per < 1 ? 1 : per);
}

// dst: this rank's compact buffer [*, C]; src_tab: every rank's buffer (only rows < n_max, the own blocks, are read).
MSHA_API int msha_peer_gather_rows(float* dst, const uint64_t* src_tab, int world, int rank, const int32_t* lists,
                                   const int64_t* list_beg, const int64_t* list_end, const int64_t* list_first,
                                   const int64_t* local_off, int64_t max_rows_per_peer, int64_t C, int max_ctas,
                                   void* stream) {
    MSHA_REQUIRE(dst && src_tab && lists && list_beg && list_end && list_first && local_off, "peer_gather_rows: null argument");
    MSHA_REQUIRE(C > 0 && ((uintptr_t)dst & 15) == 0, "peer_gather_rows: bad row width / alignment");
    if (world == 1 || max_rows_per_peer <= 0) return 0;
    const int vw = C % 4 == 0 ? 4 : (C % 2 == 0 ? 2 : 1);           // widest vector that tiles a row (rows stay aligned to it)
    dim3 grid(rows_grid(max_rows_per_peer, (int)(C / vw), world, max_ctas), (unsigned)(world - 1));
    cudaStream_t st = (cudaStream_t)stream;
    if (vw == 4)
        peer_gather_rows_kernel<float4><<<grid, PULL_THREADS, 0, st>>>(reinterpret_cast<float4*>(dst), src_tab, world, rank, lists,
                                                                        list_beg, list_end, list_first, local_off, (int)(C / 4));
    else if (vw == 2)
        peer_gather_rows_kernel<float2><<<grid, PULL_THREADS, 0, st>>>(reinterpret_cast<float2*>(dst), src_tab, world, rank, lists,
                                                                        list_beg, list_end, list_first, local_off, (int)(C / 2));
    else
        peer_gather_rows_kernel<float><<<grid, PULL_THREADS, 0, st>>>(dst, src_tab, world, rank, lists, list_beg, list_end,
                                                                       list_first, local_off, (int)C);
    MSHA_LAUNCH_OK();
    return 0;
}

// dst: this rank's own block [n_max, C] (accumulated into); src_tab: every rank's compact gradient buffer; lists: for every
// peer q the rows of THIS rank's block that q references (list_ptr: int64[world + 1] into lists), in the order of q's halo
// segment, which starts at row remote_off[q] of q's buffer.
MSHA_API int msha_peer_scatter_add_rows(float* dst, const uint64_t* src_tab, int world, int rank, const int32_t* lists,
                                        const int64_t* list_ptr, const int64_t* remote_off, int64_t max_rows_per_peer,
                                        int64_t C, int max_ctas, void* stream) {
    MSHA_REQUIRE(dst && src_tab && lists && list_ptr && remote_off, "peer_scatter_add_rows: null argument");
    MSHA_REQUIRE(C > 0 && ((uintptr_t)dst & 15) == 0, "peer_scatter_add_rows: bad row width / alignment");
    if (world == 1 || max_rows_per_peer <= 0) return 0;
    const int vw = C % 4 == 0 ? 4 : (C % 2 == 0 ? 2 : 1);
    dim3 grid(rows_grid(max_rows_per_peer, (int)(C / vw), world, max_ctas), (unsigned)(world - 1));
    cudaStream_t st = (cudaStream_t)stream;
    if (vw == 4)
        peer_scatter_add_rows_kernel<float4><<<grid, PULL_THREADS, 0, st>>>(reinterpret_cast<float4*>(dst), src_tab, world, rank, lists,
                                                                             list_ptr, remote_off, (int)(C / 4));
    else if (vw == 2)
        peer_scatter_add_rows_kernel<float2><<<grid, PULL_THREADS, 0, st>>>(reinterpret_cast<float2*>(dst), src_tab, world, rank, lists,
                                                                             list_ptr, remote_off, (int)(C / 2));
    else
        peer_scatter_add_rows_kernel<float><<<grid, PULL_THREADS, 0, st>>>(dst, src_tab, world, rank, lists, list_ptr, remote_off, (int)C);
    MSHA_LAUNCH_OK();
    return 0;
}
