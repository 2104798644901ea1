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

__global__ void peer_signal_kernel(const uint64_t* __restrict__ flag_tab, int world, int rank, int channel,
                                   uint32_t peer_mask, const uint32_t* __restrict__ epoch, uint32_t value) {
    const int q = threadIdx.x;
    if (q >= world || !((peer_mask >> q) & 1u)) return;
    const uint32_t v = value + (epoch ? *epoch : 0u);
    __threadfence_system();
    uint32_t* f = reinterpret_cast<uint32_t*>(flag_tab[q]) + (int64_t)channel * world + rank;
    st_release_sys(f, v);
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
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flag_tab, world, rank, channel, peer_mask, epoch, value);
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
