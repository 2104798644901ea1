// Error reporting and library identity for the msha_b200 C-ABI.
#include "common.cuh"
#include <string.h>

static thread_local char g_last_error[512] = "";

void msha_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

MSHA_API const char* msha_last_error(void) { return g_last_error; }

MSHA_API int msha_abi_version(void) { return 1; }

unsigned long long g_msha_launches = 0;
// number of CUDA kernels this library has launched in the calling process (bench.py: gpu_launches)
MSHA_API uint64_t msha_launch_count(void) { return (uint64_t)__atomic_load_n(&g_msha_launches, __ATOMIC_RELAXED); }

// 0 when a device of compute capability 10.x is current; otherwise a negative code + message.
MSHA_API int msha_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { msha_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return -2; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { msha_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return -2; }
    if (p.major != 10) {
        msha_set_error("msha_b200 is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
        return -3;
    }
    return 0;
}

// Dropout epoch: see common.cuh.  Stream-ordered and capturable; one hook per translation unit that draws dropout.
int msha_drop_epoch_hook_gat(unsigned long long v, int set, cudaStream_t st);
int msha_drop_epoch_hook_dense(unsigned long long v, int set, cudaStream_t st);
int msha_drop_epoch_hook_intra(unsigned long long v, int set, cudaStream_t st);

static int drop_epoch_all(unsigned long long v, int set, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int (*hooks[3])(unsigned long long, int, cudaStream_t) = {msha_drop_epoch_hook_gat, msha_drop_epoch_hook_dense,
                                                               msha_drop_epoch_hook_intra};
    for (int i = 0; i < 3; ++i) {
        int e = hooks[i](v, set, st);
        if (e != 0) {
            msha_set_error("dropout epoch kernel launch -> %s", cudaGetErrorString((cudaError_t)e));
            return e;
        }
        __atomic_fetch_add(&g_msha_launches, 1ull, __ATOMIC_RELAXED);
    }
    return 0;
}
MSHA_API int msha_dropout_epoch_set(uint64_t epoch, void* stream) { return drop_epoch_all(epoch, 1, stream); }
MSHA_API int msha_dropout_epoch_advance(uint64_t by, void* stream) { return drop_epoch_all(by, 0, stream); }
