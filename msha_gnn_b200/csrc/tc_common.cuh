// tcgen05 / TMEM / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr uint32_t SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();      // a dead pipeline must fault, never hang the GPU
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the even CTA of a 2-CTA cluster issues the MMAs for both, each CTA holds its
// own 128 accumulator rows in TMEM and half of the B tile in shared memory ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default (.release.cta) semantics on
// purpose: a cluster-scope release costs MEMBAR.ALL.GPU, which drains every load / store the thread still has in flight.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta)
        : "memory");
}
// wait on a local barrier whose arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();
    } while (!done);
}
// TMA load whose completion bytes are counted on the leader CTA's barrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// commit: one arrival on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int CTAS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    if constexpr (CTAS == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <int CTAS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    if constexpr (CTAS == 2)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), version 1.
// K-major fp32 tiles use SWIZZLE_64B (8 rows x 64 B atoms, SBO = 512).  MN-major 32-bit operands only exist in
// the SWIZZLE_128B_BASE32B layout (Swizzle<2,5,2>: 4 K-rows x 128 B atoms, SBO = 512 between K atoms, LBO between
// 32-element MN chunks) -- written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;  // 4 = SWIZZLE_64B (K-major), 1 = SWIZZLE_128B_BASE32B (MN-major tf32)
    return d;
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4)                    // c_format = F32
           | (2u << 7)                  // a_format = TF32
           | (2u << 10)                 // b_format = TF32
           | ((a_mn ? 1u : 0u) << 15)   // a_major
           | ((b_mn ? 1u : 0u) << 16)   // b_major
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// Tile addressing used by software producers (must match what TMA writes for the same swizzle mode):
//   K-major, 64-byte rows, SWIZZLE_64B  (Swizzle<2,4,3>): 16-byte chunk c of row r lands at chunk c ^ ((r >> 1) & 3)
__device__ __forceinline__ uint32_t sw64_offset(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
//   MN-major, 128-byte rows (32 fp32 of MN) x K rows, SWIZZLE_128B_BASE32B (Swizzle<2,5,2>): 32-byte unit u of K-row kk
//   lands at unit u ^ (kk & 3); c16 = 16-byte chunk (0..7) inside the 128-byte row
__device__ __forceinline__ uint32_t sw128b32_offset(int kk, int c16) {
    return (uint32_t)(kk * 128 + ((((c16 >> 1) ^ (kk & 3)) << 5) | ((c16 & 1) << 4)));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tc_get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D fp32 tensor map: inner dimension `inner` (contiguous), `outer` rows of stride ld floats.
// K-major operands use SWIZZLE_64B boxes {16, rows}; MN-major operands SWIZZLE_128B_ATOM_32B boxes {32, 16}.
inline int tc_make_map(CUtensorMap* map, const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                       int box_outer, bool mn_major) {
    EncodeTiledFn enc = tc_get_encode();
    if (!enc) { msha_set_error("cuTensorMapEncodeTiled entry point unavailable"); return -2; }
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { msha_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -2; }
    return 0;
}

// Explicit shared-state-space vector accesses.  Pointers carved out of the dynamic shared buffer are generic to the
// compiler, which then emits LD.E / ST.E (long-scoreboard latency, no reordering against global stores or atomics).
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// split x into hi (exact in tf32) and lo = x - hi
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// In-place hi/lo split of a TMA-staged fp32 tile: n16 16-byte chunks at `hi` become the tf32-exact parts, the
// remainders go to the same offsets at `lo`.  Loads are batched four deep ahead of the dependent stores.
__device__ __forceinline__ void split_tile_inplace(uint32_t hi, uint32_t lo, int n16, int tid, int nthreads) {
    int c = tid;
    for (; c + 3 * nthreads < n16; c += 4 * nthreads) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = lds128(hi + (uint32_t)(c + u * nthreads) * 16u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float4 h, l;
            split_tf32(x[u].x, h.x, l.x);
            split_tf32(x[u].y, h.y, l.y);
            split_tf32(x[u].z, h.z, l.z);
            split_tf32(x[u].w, h.w, l.w);
            sts128(hi + (uint32_t)(c + u * nthreads) * 16u, h);
            sts128(lo + (uint32_t)(c + u * nthreads) * 16u, l);
        }
    }
    for (; c < n16; c += nthreads) {
        const float4 x = lds128(hi + (uint32_t)c * 16u);
        float4 h, l;
        split_tf32(x.x, h.x, l.x);
        split_tf32(x.y, h.y, l.y);
        split_tf32(x.z, h.z, l.z);
        split_tf32(x.w, h.w, l.w);
        sts128(hi + (uint32_t)c * 16u, h);
        sts128(lo + (uint32_t)c * 16u, l);
    }
}

}  // namespace
