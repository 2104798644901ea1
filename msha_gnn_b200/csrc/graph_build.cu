// K-1: on-device CSR / CSC construction, bit-exact with the reference's dense-adjacency neighbour
// sets and ordering (row-major `(adj > 0).nonzero()`; GAT.py:30, dataset.py:279-296).
//
//   * exclusive scan (3-phase, recursive over block sums)
//   * stable LSD radix sort on 64-bit keys (8-bit digits, optional 32-bit payload)
//   * COO flow records -> coalesced CSR with multiplicities as fp32 values
//   * dense (N,M) float adjacency -> CSR by ballot compaction
//   * CSR -> CSC (+ perm) by a stable sort on the column index
//   * isolated-row augmentation (rows with no neighbour attend uniformly to all M columns)
//
// All memory is caller-owned; data dependent sizes are returned through device scalars.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// exclusive scan
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const int32_t* __restrict__ in, int64_t n_in, int32_t* __restrict__ out, int64_t n_out,
                 int32_t* __restrict__ tile_sums) {
    __shared__ int32_t warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int32_t local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        v[k] = (i < n_in) ? in[i] : 0;
        local += v[k];
    }
    // inclusive warp scan of per-thread totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int32_t warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += warp_tot[w];
    int32_t run = warp_off + incl - local;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        if (i < n_out) out[i] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1 && tile_sums) tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int32_t* __restrict__ out, int64_t n_out, const int32_t* __restrict__ tile_offs) {
    const int32_t off = tile_offs[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_THREADS) {
        int64_t i = base + k;
        if (i < n_out) out[i] += off;
    }
}

static size_t scan_ws_bytes(int64_t n) {
    size_t total = 0;
    while (n > SCAN_TILE) {
        int64_t tiles = msha_cdiv(n, SCAN_TILE);
        total += msha_align256((size_t)tiles * sizeof(int32_t));
        n = tiles;
    }
    return total + 256;
}

// out[i] = sum_{k<i} in[k] for i < n_out; in[k] is taken as 0 for k >= n_in.
static int scan_exclusive(const int32_t* in, int64_t n_in, int32_t* out, int64_t n_out, char* ws,
                          cudaStream_t st) {
    if (n_out <= 0) return 0;
    int64_t tiles = msha_cdiv(n_out, SCAN_TILE);
    if (tiles == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, st>>>(in, n_in, out, n_out, nullptr);
        MSHA_LAUNCH_OK();
        return 0;
    }
    int32_t* sums = reinterpret_cast<int32_t*>(ws);
    char* ws_next = ws + msha_align256((size_t)tiles * sizeof(int32_t));
    scan_tile_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n_in, out, n_out, sums);
    MSHA_LAUNCH_OK();
    int rc = scan_exclusive(sums, tiles, sums, tiles, ws_next, st);
    if (rc) return rc;
    scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(out, n_out, sums);
    MSHA_LAUNCH_OK();
    return 0;
}

MSHA_API size_t msha_scan_workspace_bytes(int64_t n) { return scan_ws_bytes(n); }

MSHA_API int msha_scan_exclusive_i32(const int32_t* in, int64_t n_in, int32_t* out, int64_t n_out, void* ws,
                                     size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n_in >= 0 && n_out >= 0, "scan: negative size");
    MSHA_REQUIRE(ws_bytes >= scan_ws_bytes(n_out), "scan: workspace too small");
    return scan_exclusive(in, n_in, out, n_out, (char*)ws, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;   // 4096 keys per block

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist,
               int nblocks) {
    __shared__ int32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];   // digit-major
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, int shift,
                  const int32_t* __restrict__ offs, int nblocks, uint64_t* __restrict__ keys_out,
                  uint32_t* __restrict__ vals_out) {
    __shared__ int32_t base[256];
    __shared__ int32_t wcnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = offs[(int64_t)threadIdx.x * nblocks + blockIdx.x];
    const int64_t tile0 = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        const int64_t i = tile0 + r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        uint64_t key = valid ? keys[i] : 0ull;
        // invalid lanes get a per-lane unique pseudo digit so they never match a valid one
        const unsigned digit = valid ? (unsigned)((key >> shift) & 255) : (256u + lane);
        const unsigned peers = __match_any_sync(FULL_MASK, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            int off = base[digit] + rank;
            for (int w = 0; w < warp; ++w) off += wcnt[w][digit];
            keys_out[off] = key;
            if (vals) vals_out[off] = vals[i];
        }
        __syncthreads();
        int add = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) add += wcnt[w][threadIdx.x];
        base[threadIdx.x] += add;
        // the zeroing at the top of the next round is ordered by the __syncthreads below it
        __syncthreads();
    }
}

static size_t rs_ws_bytes(int64_t n) {
    int64_t nblocks = msha_cdiv(n > 0 ? n : 1, RS_TILE);
    int64_t nh = nblocks * 256;
    return msha_align256((size_t)nh * sizeof(int32_t)) + scan_ws_bytes(nh);
}

// Sorts by the 8-bit digits at `shifts[0..npasses)` in that order (least significant first).
// Result always ends in keys/vals (a copy is made if the pass count is odd).
static int radix_sort(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, int64_t n,
                      const int* shifts, int npasses, char* ws, cudaStream_t st) {
    if (n <= 1 || npasses == 0) return 0;
    const int nblocks = (int)msha_cdiv(n, RS_TILE);
    const int64_t nh = (int64_t)nblocks * 256;
    int32_t* hist = reinterpret_cast<int32_t*>(ws);
    char* scan_ws = ws + msha_align256((size_t)nh * sizeof(int32_t));
    uint64_t* src = keys; uint64_t* dst = keys_tmp;
    uint32_t* vsrc = vals; uint32_t* vdst = vals_tmp;
    for (int p = 0; p < npasses; ++p) {
        rs_hist_kernel<<<nblocks, RS_THREADS, 0, st>>>(src, n, shifts[p], hist, nblocks);
        MSHA_LAUNCH_OK();
        int rc = scan_exclusive(hist, nh, hist, nh, scan_ws, st);
        if (rc) return rc;
        rs_scatter_kernel<<<nblocks, RS_THREADS, 0, st>>>(src, vsrc, n, shifts[p], hist, nblocks, dst, vdst);
        MSHA_LAUNCH_OK();
        uint64_t* t = src; src = dst; dst = t;
        uint32_t* vt = vsrc; vsrc = vdst; vdst = vt;
    }
    if (src != keys) {
        MSHA_CUDA(cudaMemcpyAsync(keys, src, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        if (vals) MSHA_CUDA(cudaMemcpyAsync(vals, vsrc, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

static int bits_for(int64_t n_values) {   // bits needed to represent 0..n_values-1
    int b = 0;
    while (b < 63 && ((int64_t)1 << b) < n_values) ++b;
    return b;
}

MSHA_API size_t msha_radix_sort_workspace_bytes(int64_t n) { return rs_ws_bytes(n); }

// Generic entry (used by tests): sort keys (and payload) by bits [begin_bit, end_bit).
MSHA_API int msha_radix_sort_u64(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp,
                                 int64_t n, int begin_bit, int end_bit, void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n >= 0 && begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit, "radix_sort: bad arguments");
    MSHA_REQUIRE((vals == nullptr) == (vals_tmp == nullptr), "radix_sort: vals/vals_tmp must both be set");
    MSHA_REQUIRE(ws_bytes >= rs_ws_bytes(n), "radix_sort: workspace too small");
    int shifts[8], np = 0;
    for (int b = begin_bit; b < end_bit; b += 8) shifts[np++] = b;
    return radix_sort(keys, keys_tmp, vals, vals_tmp, n, shifts, np, (char*)ws, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// COO -> CSR
// ---------------------------------------------------------------------------------------------
__global__ void coo_make_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n,
                                     int64_t n_rows, int64_t n_cols, uint64_t* __restrict__ keys,
                                     int32_t* __restrict__ status) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t s = src[i], d = dst[i];
    if (s < 0 || s >= n_rows || d < 0 || d >= n_cols) {
        atomicOr(status, 1);
        s = 0; d = 0;
    }
    keys[i] = ((uint64_t)s << 32) | (uint64_t)d;
}

__global__ void coo_flag_heads_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// uid[i] (exclusive scan of head flags) + flag -1 == slot of the run this record belongs to.
__global__ void coo_compact_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ excl, int64_t n,
                                   int32_t* __restrict__ col, float* __restrict__ val, int32_t* __restrict__ rowcnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys[i];
    const bool head = (i == 0) || (k != keys[i - 1]);
    const int32_t slot = excl[i] + (head ? 1 : 0) - 1;
    if (head) {
        col[slot] = (int32_t)(k & 0xffffffffu);
        atomicAdd(&rowcnt[(int32_t)(k >> 32)], 1);
    }
    atomicAdd(&val[slot], 1.0f);   // integer valued, < 2^24: exact and order independent
}

MSHA_API size_t msha_csr_from_coo_workspace_bytes(int64_t n, int64_t n_rows, int64_t n_cols) {
    (void)n_cols;
    size_t keys = msha_align256((size_t)(n > 0 ? n : 1) * sizeof(uint64_t));
    size_t flags = msha_align256((size_t)(n + 1) * sizeof(int32_t));
    size_t rows = msha_align256((size_t)(n_rows + 1) * sizeof(int32_t));
    size_t sc = scan_ws_bytes((n > n_rows ? n : n_rows) + 1);
    return 2 * keys + 2 * flags + rows + sc + rs_ws_bytes(n);
}

// src/dst: int64[n] device.  rowptr: int32[n_rows+1].  col: int32[n], val: float[n] (first nnz entries
// valid, nnz = rowptr[n_rows]).  status: int32[1] device, bit0 = an endpoint was out of range.
MSHA_API int msha_csr_from_coo(const int64_t* src, const int64_t* dst, int64_t n, int64_t n_rows, int64_t n_cols,
                               int32_t* rowptr, int32_t* col, float* val, int32_t* status, void* ws,
                               size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n >= 0 && n_rows >= 0 && n_cols >= 0, "csr_from_coo: negative size");
    MSHA_REQUIRE(n < ((int64_t)1 << 31) && n_rows < ((int64_t)1 << 31) && n_cols < ((int64_t)1 << 31),
                 "csr_from_coo: sizes must fit int32");
    MSHA_REQUIRE(ws_bytes >= msha_csr_from_coo_workspace_bytes(n, n_rows, n_cols), "csr_from_coo: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    MSHA_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    char* p = (char*)ws;
    const size_t keys_b = msha_align256((size_t)(n > 0 ? n : 1) * sizeof(uint64_t));
    const size_t flags_b = msha_align256((size_t)(n + 1) * sizeof(int32_t));
    const size_t rows_b = msha_align256((size_t)(n_rows + 1) * sizeof(int32_t));
    uint64_t* keys = (uint64_t*)p; p += keys_b;
    uint64_t* keys_tmp = (uint64_t*)p; p += keys_b;
    int32_t* flags = (int32_t*)p; p += flags_b;
    int32_t* excl = (int32_t*)p; p += flags_b;
    int32_t* rowcnt = (int32_t*)p; p += rows_b;
    char* scan_ws = p; p += scan_ws_bytes((n > n_rows ? n : n_rows) + 1);
    char* sort_ws = p;
    MSHA_CUDA(cudaMemsetAsync(rowcnt, 0, (size_t)(n_rows + 1) * sizeof(int32_t), st));
    if (n > 0) {
        const int T = 256;
        const unsigned G = (unsigned)msha_cdiv(n, T);
        MSHA_CUDA(cudaMemsetAsync(val, 0, (size_t)n * sizeof(float), st));
        coo_make_keys_kernel<<<G, T, 0, st>>>(src, dst, n, n_rows, n_cols, keys, status);
        MSHA_LAUNCH_OK();
        int shifts[8], np = 0;
        for (int b = 0; b < bits_for(n_cols); b += 8) shifts[np++] = b;
        for (int b = 0; b < bits_for(n_rows); b += 8) shifts[np++] = 32 + b;
        int rc = radix_sort(keys, keys_tmp, nullptr, nullptr, n, shifts, np, sort_ws, st);
        if (rc) return rc;
        coo_flag_heads_kernel<<<G, T, 0, st>>>(keys, n, flags);
        MSHA_LAUNCH_OK();
        rc = scan_exclusive(flags, n, excl, n, scan_ws, st);
        if (rc) return rc;
        coo_compact_kernel<<<G, T, 0, st>>>(keys, excl, n, col, val, rowcnt);
        MSHA_LAUNCH_OK();
    }
    return scan_exclusive(rowcnt, n_rows, rowptr, n_rows + 1, scan_ws, st);
}

// ---------------------------------------------------------------------------------------------
// dense (N,M) float adjacency -> CSR  (neighbour set == adj > 0, GAT.py:30)
// ---------------------------------------------------------------------------------------------
__global__ void dense_count_kernel(const float* __restrict__ adj, int64_t n_rows, int64_t n_cols, int64_t ld,
                                   int32_t* __restrict__ rowcnt) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const float* a = adj + row * ld;
    int cnt = 0;
    for (int64_t j0 = 0; j0 < n_cols; j0 += 32) {
        int64_t j = j0 + lane;
        bool nz = (j < n_cols) && (a[j] > 0.f);
        cnt += __popc(__ballot_sync(FULL_MASK, nz));
    }
    if (lane == 0) rowcnt[row] = cnt;
}

__global__ void dense_fill_kernel(const float* __restrict__ adj, int64_t n_rows, int64_t n_cols, int64_t ld,
                                  const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                  float* __restrict__ val) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const float* a = adj + row * ld;
    int32_t off = rowptr[row];
    for (int64_t j0 = 0; j0 < n_cols; j0 += 32) {
        int64_t j = j0 + lane;
        float v = (j < n_cols) ? a[j] : 0.f;
        bool nz = v > 0.f;
        unsigned b = __ballot_sync(FULL_MASK, nz);
        if (nz) {
            int32_t pos = off + __popc(b & ((1u << lane) - 1u));
            col[pos] = (int32_t)j;
            if (val) val[pos] = v;
        }
        off += __popc(b);
    }
}

MSHA_API size_t msha_csr_from_dense_workspace_bytes(int64_t n_rows) {
    return msha_align256((size_t)(n_rows + 1) * sizeof(int32_t)) + scan_ws_bytes(n_rows + 1);
}

// Phase 1: rowptr[n_rows+1] (nnz = rowptr[n_rows]); Phase 2 (after the caller sized col/val): fill.
MSHA_API int msha_csr_from_dense_rowptr(const float* adj, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t* rowptr,
                                        void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "csr_from_dense: bad shape");
    MSHA_REQUIRE(n_rows < ((int64_t)1 << 31) && n_cols < ((int64_t)1 << 31), "csr_from_dense: sizes must fit int32");
    MSHA_REQUIRE(ws_bytes >= msha_csr_from_dense_workspace_bytes(n_rows), "csr_from_dense: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* rowcnt = (int32_t*)ws;
    char* scan_ws = (char*)ws + msha_align256((size_t)(n_rows + 1) * sizeof(int32_t));
    if (n_rows > 0) {
        dense_count_kernel<<<(unsigned)msha_cdiv(n_rows * 32, 256), 256, 0, st>>>(adj, n_rows, n_cols, ld, rowcnt);
        MSHA_LAUNCH_OK();
    }
    return scan_exclusive(rowcnt, n_rows, rowptr, n_rows + 1, scan_ws, st);
}

MSHA_API int msha_csr_from_dense_fill(const float* adj, int64_t n_rows, int64_t n_cols, int64_t ld,
                                      const int32_t* rowptr, int32_t* col, float* val, void* stream) {
    MSHA_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "csr_from_dense: bad shape");
    if (n_rows == 0) return 0;
    dense_fill_kernel<<<(unsigned)msha_cdiv(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        adj, n_rows, n_cols, ld, rowptr, col, val);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// isolated rows: a row with no neighbour attends uniformly to all M columns (softmax of an
// all -9e15 row, GAT.py:29-31).  The attention CSR stores those as M "masked" edges, col = ~j.
// ---------------------------------------------------------------------------------------------
__global__ void aug_count_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t n_cols,
                                 int32_t* __restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    int32_t d = rowptr[i + 1] - rowptr[i];
    cnt[i] = d == 0 ? n_cols : d;
}

__global__ void aug_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                                int32_t n_cols, const int32_t* __restrict__ rowptr_aug, int32_t* __restrict__ col_aug) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const int32_t b = rowptr[row], d = rowptr[row + 1] - b, o = rowptr_aug[row];
    if (d == 0) {
        for (int32_t j = lane; j < n_cols; j += 32) col_aug[o + j] = ~j;
    } else {
        for (int32_t k = lane; k < d; k += 32) col_aug[o + k] = col[b + k];
    }
}

MSHA_API int msha_csr_augment_rowptr(const int32_t* rowptr, int64_t n_rows, int64_t n_cols, int32_t* rowptr_aug,
                                     void* ws, size_t ws_bytes, void* stream) {
    MSHA_REQUIRE(ws_bytes >= msha_csr_from_dense_workspace_bytes(n_rows), "csr_augment: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* cnt = (int32_t*)ws;
    char* scan_ws = (char*)ws + msha_align256((size_t)(n_rows + 1) * sizeof(int32_t));
    if (n_rows > 0) {
        aug_count_kernel<<<(unsigned)msha_cdiv(n_rows, 256), 256, 0, st>>>(rowptr, n_rows, (int32_t)n_cols, cnt);
        MSHA_LAUNCH_OK();
    }
    return scan_exclusive(cnt, n_rows, rowptr_aug, n_rows + 1, scan_ws, st);
}

MSHA_API int msha_csr_augment_fill(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols,
                                   const int32_t* rowptr_aug, int32_t* col_aug, void* stream) {
    if (n_rows == 0) return 0;
    aug_fill_kernel<<<(unsigned)msha_cdiv(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        rowptr, col, n_rows, (int32_t)n_cols, rowptr_aug, col_aug);
    MSHA_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// CSR -> CSC (+perm).  Stable sort of the CSR slots by column: rows stay ascending inside a column.
// ---------------------------------------------------------------------------------------------
__global__ void csc_keys_kernel(const int32_t* __restrict__ col, int64_t nnz, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ vals, int32_t* __restrict__ colcnt) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int32_t c = col[e];
    c = c < 0 ? ~c : c;
    keys[e] = (uint64_t)(uint32_t)c;
    vals[e] = (uint32_t)e;
    atomicAdd(&colcnt[c], 1);
}

__global__ void csc_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, const uint32_t* __restrict__ perm_u,
                                int64_t nnz, int32_t* __restrict__ rowidx, int32_t* __restrict__ perm) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int32_t e = (int32_t)perm_u[k];
    // largest r with rowptr[r] <= e
    int64_t lo = 0, hi = n_rows;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] <= e) lo = mid; else hi = mid;
    }
    rowidx[k] = (int32_t)lo;
    perm[k] = e;
}

MSHA_API size_t msha_csc_from_csr_workspace_bytes(int64_t nnz, int64_t n_cols) {
    size_t keys = msha_align256((size_t)(nnz > 0 ? nnz : 1) * sizeof(uint64_t));
    size_t vals = msha_align256((size_t)(nnz > 0 ? nnz : 1) * sizeof(uint32_t));
    size_t cols = msha_align256((size_t)(n_cols + 1) * sizeof(int32_t));
    return 2 * keys + 2 * vals + cols + scan_ws_bytes(n_cols + 1) + rs_ws_bytes(nnz);
}

MSHA_API int msha_csc_from_csr(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols, int64_t nnz,
                               int32_t* colptr, int32_t* rowidx, int32_t* perm, void* ws, size_t ws_bytes,
                               void* stream) {
    MSHA_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "csc_from_csr: negative size");
    MSHA_REQUIRE(ws_bytes >= msha_csc_from_csr_workspace_bytes(nnz, n_cols), "csc_from_csr: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* p = (char*)ws;
    const size_t keys_b = msha_align256((size_t)(nnz > 0 ? nnz : 1) * sizeof(uint64_t));
    const size_t vals_b = msha_align256((size_t)(nnz > 0 ? nnz : 1) * sizeof(uint32_t));
    const size_t cols_b = msha_align256((size_t)(n_cols + 1) * sizeof(int32_t));
    uint64_t* keys = (uint64_t*)p; p += keys_b;
    uint64_t* keys_tmp = (uint64_t*)p; p += keys_b;
    uint32_t* vals = (uint32_t*)p; p += vals_b;
    uint32_t* vals_tmp = (uint32_t*)p; p += vals_b;
    int32_t* colcnt = (int32_t*)p; p += cols_b;
    char* scan_ws = p; p += scan_ws_bytes(n_cols + 1);
    char* sort_ws = p;
    MSHA_CUDA(cudaMemsetAsync(colcnt, 0, (size_t)(n_cols + 1) * sizeof(int32_t), st));
    if (nnz > 0) {
        const unsigned G = (unsigned)msha_cdiv(nnz, 256);
        csc_keys_kernel<<<G, 256, 0, st>>>(col, nnz, keys, vals, colcnt);
        MSHA_LAUNCH_OK();
        int shifts[8], np = 0;
        for (int b = 0; b < bits_for(n_cols); b += 8) shifts[np++] = b;
        int rc = radix_sort(keys, keys_tmp, vals, vals_tmp, nnz, shifts, np, sort_ws, st);
        if (rc) return rc;
        csc_rows_kernel<<<G, 256, 0, st>>>(rowptr, n_rows, vals, nnz, rowidx, perm);
        MSHA_LAUNCH_OK();
    }
    return scan_exclusive(colcnt, n_cols, colptr, n_cols + 1, scan_ws, st);
}

// ---------------------------------------------------------------------------------------------
// value utilities for the GCN baseline: column-normalised CSR values  A[:,j] / colsum[j]
// (model.py:95-100: degrees^-0.5 applied twice, in fp32).
// ---------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const int32_t* __restrict__ col, const float* __restrict__ val, int64_t nnz,
                              float* __restrict__ colsum) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    atomicAdd(&colsum[col[e]], val[e]);
}
__global__ void colnorm_kernel(const int32_t* __restrict__ col, const float* __restrict__ val, int64_t nnz,
                               const float* __restrict__ colsum, float* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    float d = powf(colsum[col[e]], -0.5f);
    out[e] = (val[e] * d) * d;
}

// colsum: float[n_cols] scratch (zeroed here).
MSHA_API int msha_csr_normalize_columns(const int32_t* col, const float* val, int64_t nnz, int64_t n_cols,
                                        float* colsum, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MSHA_CUDA(cudaMemsetAsync(colsum, 0, (size_t)n_cols * sizeof(float), st));
    if (nnz == 0) return 0;
    const unsigned G = (unsigned)msha_cdiv(nnz, 256);
    colsum_kernel<<<G, 256, 0, st>>>(col, val, nnz, colsum);
    MSHA_LAUNCH_OK();
    colnorm_kernel<<<G, 256, 0, st>>>(col, val, nnz, colsum, out);
    MSHA_LAUNCH_OK();
    return 0;
}
