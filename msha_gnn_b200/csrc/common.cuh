// Shared helpers for the msha_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define MSHA_API extern "C" __attribute__((visibility("default")))

// thread-local last-error message (defined in api_common.cu)
void msha_set_error(const char* fmt, ...);

#define MSHA_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            msha_set_error(__VA_ARGS__);                          \
            return -1;                                            \
        }                                                         \
    } while (0)

#define MSHA_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            msha_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                               \
        }                                                                                 \
    } while (0)

extern unsigned long long g_msha_launches;   // kernels launched by this library (api_common.cu)

#define MSHA_LAUNCH_OK()                                                                   \
    do {                                                                                   \
        __atomic_fetch_add(&g_msha_launches, 1ull, __ATOMIC_RELAXED);                      \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            msha_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return (int)_e;                                                                \
        }                                                                                  \
    } while (0)

static inline int64_t msha_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t msha_align256(size_t x) { return (x + 255) & ~(size_t)255; }

// SM count of the CURRENT device (148 on B200), queried once per device: grid sizes follow the device a call runs on.
static inline int msha_num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}
#define MSHA_NUM_SMS msha_num_sms()

// Function attributes (cudaFuncSetAttribute: opt-in dynamic shared memory) are per DEVICE: a call site remembers, per
// device, that it has set them -- a process that drives several GPUs would otherwise launch with the 48 KB default on all
// but the first.  (A race between two host threads only repeats the harmless call.)
struct MshaPerDeviceOnce {
    unsigned long long done = 0ull;
    bool need() const {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        return !((done >> dev) & 1ull);
    }
    void mark() {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) done |= 1ull << dev;
    }
};

#ifdef __CUDACC__
constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Stream layout (oracle/msha_oracle.py:_philox_words):
// word i of stream s under seed k == lane (i & 3) of block i>>2 with counter
// (blk_lo, blk_hi, s, 0) and key (k_lo, k_hi).
// ---------------------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ uint32_t philox_word(uint64_t seed, uint32_t stream, uint64_t i) {
    uint64_t blk = i >> 2;
    Philox4 r = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u, (uint32_t)seed,
                              (uint32_t)(seed >> 32));
    uint32_t l = (uint32_t)(i & 3);
    return l == 0 ? r.x : (l == 1 ? r.y : (l == 2 ? r.z : r.w));
}

// ---------------------------------------------------------------------------------------------
// Dropout epoch (CUDA-graph replay).  A captured launch bakes its by-value seed, so a replayed graph would repeat its
// keep-masks.  Every dropout draw therefore keys Philox with seed + epoch * 2^64/phi, where `epoch` is a device word
// that msha_dropout_epoch_advance() bumps with a (capturable) one-thread kernel -- put it first in the captured step.
// The word is per translation unit (no relocatable device code): MSHA_DEFINE_DROP_EPOCH_HOOK(tu) defines the hook
// api_common.cu calls for each unit that draws dropout.  epoch == 0 (the default) leaves the stream unchanged.
// ---------------------------------------------------------------------------------------------
static __device__ unsigned long long g_drop_epoch __attribute__((unused)) = 0ull;

__device__ __forceinline__ uint64_t drop_seed_eff(uint64_t seed) {
    return seed + (uint64_t)g_drop_epoch * 0x9E3779B97F4A7C15ull;
}

// dropout multiplier of element i: 0 if dropped, 1/(1-p) if kept (keep iff word >= thr)
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t stream, uint64_t i, uint32_t thr,
                                               float inv_keep) {
    return philox_word(drop_seed_eff(seed), stream, i) >= thr ? inv_keep : 0.f;
}

#define MSHA_DEFINE_DROP_EPOCH_HOOK(tu)                                                             \
    __global__ void drop_epoch_kernel_##tu(unsigned long long v, int set) {                         \
        g_drop_epoch = set ? v : g_drop_epoch + v;                                                  \
    }                                                                                               \
    int msha_drop_epoch_hook_##tu(unsigned long long v, int set, cudaStream_t st) {                 \
        drop_epoch_kernel_##tu<<<1, 1, 0, st>>>(v, set);                                            \
        return (int)cudaGetLastError();                                                             \
    }
#endif
