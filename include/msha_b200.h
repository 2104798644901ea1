/*
 * msha_b200 -- C-ABI of the B200-native (sm_100a) graph-attention + link-scoring hot path of MSHA-GNN.
 *
 * The reference (Sienna12321/MSHA--GNN) has no FFI: its boundary is the torch nn.Module surface
 * (GAT.py:7,20,39,53; Ours.py:30,54,145,160; Ablation.py; HGANE.py:12,37; LLP.py:87,104; model.py:16,34).
 * The Python package msha_gnn_b200 mirrors that surface and calls the entry points below through ctypes;
 * every entry cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated; fp32 data, int32 CSR indices, int64 at the pair/COO surface
 *   - the caller owns every buffer (outputs and workspaces); the library never allocates or frees device memory
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, re-entrant, and never synchronise
 *   - return value: 0 ok, < 0 invalid argument, > 0 cudaError_t; message via msha_last_error() (thread local)
 *   - feature tensors are row-major [rows, H, D] (C = H*D contiguous channels per row)
 *   - attention CSR: a column index c < 0 denotes a "masked" edge to column ~c (rows without neighbours attend
 *     uniformly to all M columns: softmax of an all -9e15 row, GAT.py:29-31)
 *   - dropout: Philox4x32-10, element i of stream s is kept iff word_i >= floor(p*2^32); p == 0 disables
 */
#ifndef MSHA_B200_H
#define MSHA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* activation codes */
#define MSHA_ACT_NONE 0
#define MSHA_ACT_ELU 1
#define MSHA_ACT_RELU 2
#define MSHA_ACT_SIGMOID_RELU 3 /* sigmoid(relu(x)): LinkPredictor, LLP.py:108-115 */
#define MSHA_ACT_LRELU 4
#define MSHA_ACT_SIGMOID 5

/* ---- library ---- */
const char* msha_last_error(void);
int msha_abi_version(void);
int msha_check_device(void); /* 0 iff the current device is sm_100 */
uint64_t msha_launch_count(void); /* kernels launched by this library so far (bench.py gpu_launches) */

/* ---- K-1 graph build: replaces dataset.py:279-296 (inter_adjacent), dataset.py:260-277 (intra_adjacent),
 *      the dense `adj > 0` masks of GAT.py:30 / Ours.py:67,81-82 and model.py:95-100 (column normalisation) ---- */
size_t msha_scan_workspace_bytes(int64_t n);
int msha_scan_exclusive_i32(const int32_t* in, int64_t n_in, int32_t* out, int64_t n_out, void* ws, size_t ws_bytes,
                            void* stream);
size_t msha_radix_sort_workspace_bytes(int64_t n);
int msha_radix_sort_u64(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, int64_t n,
                        int begin_bit, int end_bit, void* ws, size_t ws_bytes, void* stream);
size_t msha_csr_from_coo_workspace_bytes(int64_t n, int64_t n_rows, int64_t n_cols);
int msha_csr_from_coo(const int64_t* src, const int64_t* dst, int64_t n, int64_t n_rows, int64_t n_cols,
                      int32_t* rowptr, int32_t* col, float* val, int32_t* status, void* ws, size_t ws_bytes,
                      void* stream);
size_t msha_csr_from_dense_workspace_bytes(int64_t n_rows);
int msha_csr_from_dense_rowptr(const float* adj, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t* rowptr, void* ws,
                               size_t ws_bytes, void* stream);
int msha_csr_from_dense_fill(const float* adj, int64_t n_rows, int64_t n_cols, int64_t ld, const int32_t* rowptr,
                             int32_t* col, float* val, void* stream);
int msha_csr_augment_rowptr(const int32_t* rowptr, int64_t n_rows, int64_t n_cols, int32_t* rowptr_aug, void* ws,
                            size_t ws_bytes, void* stream);
int msha_csr_augment_fill(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols,
                          const int32_t* rowptr_aug, int32_t* col_aug, void* stream);
size_t msha_csc_from_csr_workspace_bytes(int64_t nnz, int64_t n_cols);
int msha_csc_from_csr(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int64_t n_cols, int64_t nnz,
                      int32_t* colptr, int32_t* rowidx, int32_t* perm, void* ws, size_t ws_bytes, void* stream);
int msha_csr_normalize_columns(const int32_t* col, const float* val, int64_t nnz, int64_t n_cols, float* colsum,
                               float* out, void* stream);

/* Hub rows / columns of power-law graphs: rows with more than seg_limit entries are split into segments of at most
 * seg_limit consecutive CSR (CSC) slots, processed by independent warps and merged.  All pointers are device arrays;
 * the struct itself lives in host memory.  Pass NULL when the graph has no such rows. */
typedef struct msha_hub {
    int32_t seg_limit, n_segs;
    const int32_t* seg_item; /* row (column) id per segment */
    const int32_t* seg_beg;  /* first slot of the segment */
    const int32_t* seg_end;  /* one past its last slot */
    int32_t n_hub, pad_;
    const int32_t* hub_ids;     /* the split rows (columns) */
    const int32_t* hub_seg_ptr; /* [n_hub + 1] offsets into the segment arrays */
} msha_hub_t;
size_t msha_gat_fwd_hub_scratch_floats(int64_t n_segs, int H, int D);

/* ---- K-2/K-3 fused logit + segmented softmax + aggregation: replaces Ours.py:64-69,98 / Ablation.py:265-271,274 /
 *      HGANE.py:46-47,66-68.  alpha_in != NULL -> plain weighted SpMM (K-3 only). ---- */
int msha_gat_fwd(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* s_nbr, const float* s_self,
                 const float* feat, int H, int D, float slope, const float* alpha_in, float* alpha_out, float* out,
                 int act, float* lse_out, float drop_p, uint64_t drop_seed, const msha_hub_t* hub, float* hub_scratch,
                 void* stream);
/* Partitioned forward (no reference counterpart, train.py:18 is single-device): the same arithmetic as
 * msha_gat_fwd split into (1) the row statistics of the logits -- lse[row] = (H maxima, H sums of exp(e - max)), float[n_rows,
 * 2, H] --, which need only the H scores per node, and (2) one aggregation pass per owner block of columns -- edges
 * [rowbeg[i], rowend[i]) of row i -- so that a block is consumed as soon as its feature rows have arrived from the owning
 * GPU.  alpha = exp(logit - max) / sum is final in every block (the fused kernel's own expression).
 * H a power of two <= 32, D % 4 == 0.  hub_scratch of the stats call: float[2 * H * hub->n_segs]. */
int msha_gat_softmax_stats(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* s_nbr,
                           const float* s_self, int H, float slope, float* lse, const msha_hub_t* hub,
                           float* hub_scratch, void* stream);
int msha_gat_fwd_block(const int32_t* rowbeg, const int32_t* rowend, const int32_t* col, int64_t n_rows,
                       const float* s_nbr, const float* s_self, const float* lse, const float* feat, int H, int D,
                       float slope, float* alpha_out, float* out, int accumulate, float drop_p, uint64_t drop_seed,
                       const msha_hub_t* hub, void* stream);
/* backward row pass (autograd of the lines above): d alpha, softmax and LeakyReLU backward, d s_self */
int msha_gat_bwd_rows(const int32_t* rowptr, const int32_t* col, int64_t n_rows, const float* s_nbr,
                      const float* s_self, float slope, const float* alpha, const float* feat, const float* dout,
                      const float* out, int act, float* dz_out, const float* dT, const float* fT,
                      const float* dalpha_extra, const float* dlse, int H, int D, float* dlogit, float* ds_self,
                      float drop_p, uint64_t drop_seed, const msha_hub_t* hub, float* r_buf, int avg_degree_hint,
                      void* stream);
/* ---- K-4 transposed SpMM over CSC: replaces `attention_inter.t() @ h2` Ours.py:100 and the d feat pass ---- */
int msha_spmm_csc(const int32_t* colptr, const int32_t* rowidx, const int32_t* perm, int64_t n_cols, const float* w,
                  const float* feat, int H, int D, float* out, int accumulate, const float* esum_in, float* esum_out,
                  float drop_p, uint64_t drop_seed, const msha_hub_t* hub, void* stream);
/* node-level pieces of the logit: s = Wh . a (Ours.py:64: matmul(inter_input, a)) and its backward */
int msha_node_scores(const float* feat, int64_t n, int H, int D, const float* a1, float* s1, const float* a2,
                     float* s2, void* stream);
int msha_node_outer_add(float* out, int64_t n, int H, int D, const float* s1, const float* a1, const float* s2,
                        const float* a2, int accumulate, void* stream);
size_t msha_colreduce_workspace_bytes(int C);
int msha_colreduce(const float* x, const float* y, const float* s, int64_t n, int C, int D, float* out, void* ws,
                   size_t ws_bytes, void* stream);

/* ---- K-7 feature transform Wh = X @ W (GAT.py:21, Ours.py:57-58) and K-8 u @ v.T + ELU (Ours.py:108-109) ---- */
int msha_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                  int64_t ldc, int transA, int transB, const float* bias, float beta, int act, float slope,
                  void* stream);

/* tcgen05 / TMEM / TMA path of the same contraction (3xTF32 split, fp32-grade accuracy). transA/transB as above.
 * splits > 1: split-K, partial sums atomically added into a caller-zeroed C (no bias / activation). */
int msha_gemm_tf32x3_supported(const float* A, const float* B, int64_t M, int64_t N, int64_t K, int64_t lda,
                               int64_t ldb);
int msha_gemm_tf32x3(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldb, int64_t ldc, int transA, int transB, const float* bias, int act, float slope,
                     int splits, void* stream);

/* ---- a-1 GraphAttentionLayer epilogue: replaces GAT.py:24-35 (uniform masked softmax, elu(att*h)) ---- */
int msha_gal_fwd(const float* h, const int32_t* rowptr, const int32_t* col, int64_t n_rows, int H, int64_t M,
                 float* out, float drop_p, uint64_t drop_seed, void* stream);
int msha_gal_bwd(const float* dout, const float* y, const int32_t* rowptr, const int32_t* col, int64_t n_rows, int H,
                 int64_t M, float* dh, float drop_p, uint64_t drop_seed, void* stream);
/* feature dropout F.dropout(x) (Ours.py:161-162,165; GAT.py:54,56); the same call is its backward */
int msha_dropout_apply(const float* x, float* y, int64_t n, float p, uint64_t seed, uint32_t stream_id, void* stream);

/* ---- K-5 intra-scale (city / province) block: replaces Ours.py:71-90,99 without the (B,N)/(N,N) dense tensors ---- */
int msha_rowsum_exp(const int32_t* rowptr, const float* alpha, int H, const int64_t* src, int64_t B, int64_t n_cols,
                    float* T, float drop_p, uint64_t drop_seed, void* stream);
int msha_rowsum_exp_bwd(const int32_t* rowptr, const float* alpha, int H, const int64_t* src, int64_t B,
                        const float* dT, float* dalpha, float drop_p, uint64_t drop_seed, void* stream);
int msha_group_scatter_add(const int32_t* rowptr, const int32_t* col, const int32_t* row_map, const int64_t* src,
                           int64_t B, int64_t n_nodes, const float* coef, const float* feat, int H, int D, float* out,
                           float drop_p, uint64_t drop_seed, uint32_t drop_stream, void* stream);
int msha_group_gather_sum(const int32_t* rowptr, const int32_t* col, const int32_t* row_map, const int64_t* src,
                          int64_t B, int64_t n_nodes, const float* dout, int H, int D, float* G, float drop_p,
                          uint64_t drop_seed, uint32_t drop_stream, void* stream);

/* Group-sum form of the intra-scale block (no dropout on the intra attention): attention3[b, n] = coef3[b] for every member
 * n of the batch row's city (Ours.py:71-75,87), so att3.t() @ h2_ (Ours.py:99) is the same row for all members of a group:
 * IntraNC[n] = G3[city(n)] + G4[province(n)] with G[g] = sum_{b in group g} coef[b] * h2[src_b].  O(B + N), and a partitioned
 * graph exchanges only the (n_groups, C) tables.  G is zeroed by the sum; the two calls are each other's backward. */
int msha_group_rows_sum(const float* x, const int64_t* gid, int64_t n, int64_t C, int64_t n_groups, float* G, void* stream);
int msha_group_rows_add(float* out, const float* G3, const int64_t* gid3, const float* G4, const int64_t* gid4, int64_t n,
                        int64_t C, int accumulate, void* stream);

/* ---- activations (F.elu / F.leaky_relu / relu / sigmoid call sites) ---- */
int msha_act_fwd(const float* x, float* y, int64_t n, int act, float slope, void* stream);
int msha_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, float slope, void* stream);

/* Linear backward prologue: g = dy*act'(y) and bias gradient colsum(g) in one pass (lin(x) backward, LLP.py:108) */
size_t msha_act_bwd_colsum_workspace_bytes(int C);
int msha_act_bwd_colsum(const float* dy, const float* y, float* g, int64_t n, int C, int act, float slope,
                        float* colsum, void* ws, size_t ws_bytes, void* stream);

/* ---- K-6 BatchNorm1d over the node axis + LeakyReLU: replaces leakyrelu(bn(.)) Ours.py:100-101 ---- */
size_t msha_bn_workspace_bytes(int C);
int msha_bn_lrelu_fwd(const float* x, int64_t n, int C, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, int training, float momentum, float eps, float slope, float* y,
                      float* save_mean, float* save_invstd, void* ws, size_t ws_bytes, void* stream);
int msha_bn_lrelu_bwd(const float* dy, const float* y, const float* x, int64_t n, int C, const float* gamma,
                      const float* save_mean, const float* save_invstd, int training, float slope, float* dx,
                      float* xhat, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);

/* The same BatchNorm when the node axis is partitioned over GPUs: statistics pass -> the caller
 * all-reduces sums (double[2*C]: column sums of x and x^2; backward: of g and g*xhat) over the ranks -> apply pass with the
 * global row count n_total.  dgamma / dbeta come out global (identical on every rank). */
int msha_bn_stats(const float* x, int64_t n, int C, double* sums, void* ws, size_t ws_bytes, void* stream);
int msha_bn_lrelu_apply(const float* x, int64_t n, int C, const double* sums, int64_t n_total, const float* gamma,
                        const float* beta, float* running_mean, float* running_var, int training, float momentum,
                        float eps, float slope, float* y, float* save_mean, float* save_invstd, void* stream);
int msha_bn_lrelu_bwd_stats(const float* dy, const float* y, const float* x, int64_t n, int C, const float* save_mean,
                            const float* save_invstd, float slope, float* dx, float* xhat, double* sums, void* ws,
                            size_t ws_bytes, void* stream);
int msha_bn_lrelu_bwd_apply(float* dx, const float* xhat, int64_t n, int C, const float* gamma, const float* save_invstd,
                            const double* sums, int64_t n_total, int training, float* dgamma, float* dbeta, void* stream);

/* ---- read-out: log_softmax(elu(x)) rows, Ours.py:166-167 / GAT.py:57-58 ---- */
int msha_log_softmax_fwd(const float* x, int64_t n, int64_t M, int pre_elu, float* y, void* stream);
int msha_log_softmax_bwd(const float* dy, const float* y, const float* x, int64_t n, int64_t M, int pre_elu, float* dx,
                         void* stream);

/* ---- a-8 loss read-out: F.nll_loss(logp, target), mean reduction (train.py:229, LLP.py:235) ---- */
size_t msha_nll_workspace_bytes(void);
int msha_nll_loss_fwd(const float* logp, const int64_t* target, int64_t P, int64_t C, float* loss, int32_t* status,
                      void* ws, size_t ws_bytes, void* stream);
int msha_nll_loss_bwd(const int64_t* target, const float* gout, int64_t P, int64_t C, float* dlogp, void* stream);

/* ---- K-9 link scorer: replaces LinkPredictor.forward LLP.py:104-115 (x_i*x_j, Linear, ReLU, sigmoid) ---- */
int msha_pair_gather_mul(const float* hi, const float* hj, const int64_t* src, const int64_t* dst, int64_t P, int64_t C,
                         float* z, void* stream);
int msha_pair_scatter_mul_add(const float* dz, const float* hi, const float* hj, const int64_t* src,
                              const int64_t* dst, int64_t P, int64_t C, float* dhi, float* dhj, void* stream);
int msha_pair_dot(const float* hi, const float* hj, const int64_t* src, const int64_t* dst, int64_t P, int64_t C,
                  int act, float* out, void* stream);
int msha_pair_dot_bwd(const float* dout, const float* out, const float* hi, const float* hj, const int64_t* src,
                      const int64_t* dst, int64_t P, int64_t C, int act, float* dhi, float* dhj, void* stream);
/* fused tensor-core scorer (one hidden Linear): pair gather * Hadamard -> 3xTF32 tcgen05 GEMM -> bias, relu, sigmoid.
 * W0 is [Hd, C] (nn.Linear layout).  Replaces LLP.py:105-115 without materialising x_i * x_j. */
int msha_score_mlp_supported(const float* hi_tab, const float* hj_tab, const float* W0, int64_t C, int64_t Hd);
size_t msha_score_mlp_workspace_bytes(int64_t C, int64_t Hd);
int msha_score_mlp_fwd(const float* hi_tab, const float* hj_tab, const int64_t* src, const int64_t* dst, int64_t P,
                       int64_t C, const float* W0, const float* b0, int64_t Hd, int act, float slope, float* out,
                       int64_t ldo, void* ws, size_t ws_bytes, void* stream);
/* backward of msha_score_mlp_fwd in two tensor-core kernels: (1) G = dOut*act'(out), bias gradient, dZ = G @ W0 and the
 * scatter dh_i[src] += dZ*h_j[dst], dh_j[dst] += dZ*h_i[src] in the epilogue; (2) dW0 = G^T @ Z with Z regenerated from
 * the gathers.  dhi/dhj are accumulated into (caller zeroes them); dW0, db0 are overwritten; G is [P, Hd] scratch. */
int msha_score_mlp_bwd(const float* dout, const float* out, const float* hi_tab, const float* hj_tab, const int64_t* src,
                       const int64_t* dst, int64_t P, int64_t C, const float* W0, int64_t Hd, int act, float slope,
                       float* G, float* dhi, float* dhj, float* dW0, float* db0, void* ws, size_t ws_bytes, void* stream);
/* Same backward with the F.nll_loss read-out (LLP.py:235: loss = -mean_p out[p, target[p]]) folded in: dOut is never
 * materialised -- the producers of kernel (1) generate dOut[p, c] = (c == target[p]) ? -gout[0] / P : 0 on the fly.
 * gout: float[1] on the device (upstream gradient of the scalar loss).  Out-of-range targets contribute nothing. */
int msha_score_mlp_nll_bwd(const int64_t* target, const float* gout, const float* out, const float* hi_tab,
                           const float* hj_tab, const int64_t* src, const int64_t* dst, int64_t P, int64_t C,
                           const float* W0, int64_t Hd, int act, float slope, float* G, float* dhi, float* dhj,
                           float* dW0, float* db0, void* ws, size_t ws_bytes, void* stream);
/* The same backward without any contraction: under the nll read-out G = dOut * act'(out) has one non-zero per row
 * (g_p at column target[p]), so dZ[p] = g_p * W0[target[p]], dW0[c] = sum_{p: target[p] == c} g_p * Z[p], db0[c] = sum g_p --
 * O(P C) gathers and vector atomics instead of 4 P C Hd flops; pairs whose activation is flat at the target are skipped.
 * order: uint32[P] pair indices stably sorted by label (msha_score_nll_label_order; NULL = identity, for labels that
 * already come in runs): label runs keep the dW0 / db0 partial sums in registers.  dhi / dhj accumulated into,
 * dW0 / db0 overwritten.  Needs C % 4 == 0, C <= 1024 and 16-byte aligned tables / gradients; any Hd. */
int msha_score_mlp_nll_bwd_sparse(const uint32_t* order, const int64_t* target, const float* gout, const float* out,
                                  int64_t ldo, const float* hi_tab, const float* hj_tab, const int64_t* src,
                                  const int64_t* dst, int64_t P, int64_t C, const float* W0, int64_t Hd, int act,
                                  float slope, float* dhi, float* dhj, float* dW0, float* db0, void* stream);
/* Stable pair order by label, and by source row inside a label when src != NULL (rows of hi_tab, n_src of them).
 * keys, keys_tmp: uint64[P]; order, order_tmp: uint32[P]; ws: msha_radix_sort_workspace_bytes(P) */
int msha_score_nll_label_order(const int64_t* target, const int64_t* src, int64_t P, int64_t Hd, int64_t n_src,
                               uint64_t* keys, uint64_t* keys_tmp, uint32_t* order, uint32_t* order_tmp, void* ws,
                               size_t ws_bytes, void* stream);
/* builder-defined sampler (the reference's --ns_rate flags are dead code, LLP.py:26-29); like the dropout draws its Philox key
 * folds in the device-side epoch (msha_dropout_epoch_*), so a replayed CUDA graph samples fresh pairs */
int msha_negative_sample(uint64_t seed, int64_t P, int64_t n_src, int64_t n_dst, int64_t* src, int64_t* dst,
                         void* stream);
int msha_dropout_mask(uint64_t seed, uint32_t stream_id, int64_t n, float p, uint8_t* keep, void* stream);
/* Dropout epoch for CUDA-graph replay: every dropout draw keys Philox with seed + epoch * 0x9E3779B97F4A7C15 (mod 2^64).
 * A captured launch bakes its by-value seed; put msha_dropout_epoch_advance(1, capture_stream) first in the captured step
 * and each replay draws fresh keep-masks (forward and backward of one replay still agree).  Stream-ordered device-side
 * state (the one piece of global state besides the descriptor caches); epoch 0 is the default and changes nothing. */
int msha_dropout_epoch_set(uint64_t epoch, void* stream);
int msha_dropout_epoch_advance(uint64_t by, void* stream);

/* ==== multi-GPU exchange over peer-mapped memory (the halo all-gather of boundary features and the
 * reduce-scatter of their gradients; the reference is single-device, train.py:18).  Buffers are allocated by the caller in
 * memory every rank of the node has mapped (NVLink peer access); *_tab arguments are DEVICE arrays uint64[world] holding,
 * per rank, the address of that rank's buffer as mapped in the calling process. ==== */
/* flags: uint32[n_channels][world] per rank.  signal: flags_of_peer_q[channel][rank] = value (+ *epoch) for every q in
 * peer_mask, released at system scope after all earlier work of the stream.  wait: until flags[channel][q] >= value
 * (+ *epoch) for every q in peer_mask; gives up after timeout_ns (0 = never) leaving 0x100|q in *status. */
int msha_peer_signal(const uint64_t* flag_tab, int world, int rank, int channel, uint32_t peer_mask,
                     const uint32_t* epoch, uint32_t value, void* stream);
int msha_peer_wait(const uint32_t* flags, int world, int channel, uint32_t peer_mask, const uint32_t* epoch,
                   uint32_t value, uint64_t timeout_ns, int32_t* status, void* stream);
/* gathered layout [world][block_bytes]: copy the first nbytes of every remote rank's block out of that rank's buffer */
int msha_peer_pull_blocks(void* dst, const uint64_t* src_tab, int world, int rank, int64_t block_bytes, int64_t nbytes,
                          int max_ctas, void* stream);
/* halo lists: rows row_ids[row_ptr[q] .. row_ptr[q+1]) (indices into the gathered [world * n_max, C] buffer) from rank q */
int msha_peer_pull_rows(float* dst, const uint64_t* src_tab, int world, int rank, const int32_t* row_ids,
                        const int64_t* row_ptr, int64_t max_rows_per_peer, int64_t C, int max_ctas, void* stream);
/* out[i] = sum_k src_ptrs[k][i] in table order; src_ptrs is a HOST array of n_src (<= 32) device addresses */
int msha_peer_sum(float* out, const uint64_t* src_ptrs, int n_src, int64_t n, int max_ctas, void* stream);

/* Fused exchanges for small blocks: flag waits and the completion signal inside the data kernel (one launch per gather /
 * reduce-scatter of up to two buffers).  wait: every CTA waits for the flag of the peer it reads (pull) or of all peers
 * (sum); guard: flags that must hold before this rank may overwrite what its peers read in the opposite direction (checked
 * by the last CTA); done: published to every peer by the last CTA.  A channel < 0 skips that part.  counter: one device
 * word, zero on entry, left zero.  dst / src_tab / ... are HOST arrays of n_bufs (1 or 2) entries; src_tab[k] is a device
 * table uint64[world]. */
int msha_peer_exchange_pull(int n_bufs, void* const* dst, const uint64_t* const* src_tab, const int64_t* block_bytes,
                            const int64_t* nbytes, const uint32_t* flags, const uint64_t* flag_tab, int world, int rank,
                            int wait_ch, uint32_t wait_val, int guard_ch, uint32_t guard_val, int done_ch,
                            uint32_t done_val, uint64_t timeout_ns, int32_t* status, uint32_t* counter, int max_ctas,
                            void* stream);
int msha_peer_exchange_sum(int n_bufs, float* const* out, const uint64_t* const* src_tab, const int64_t* offset_bytes,
                           const int64_t* n_floats, const uint32_t* flags, const uint64_t* flag_tab, int world, int rank,
                           int wait_ch, uint32_t wait_val, int guard_ch, uint32_t guard_val, int done_ch, uint32_t done_val,
                           uint64_t timeout_ns, int32_t* status, uint32_t* counter, int max_ctas, void* stream);

/* Halo exchange (the north-star's "halo all-gather of boundary features"): per rank the columns are renumbered to
 * [own block | halo rows of peer 0 | halo rows of peer 1 | ...]; gather_rows packs the listed rows of the peers' own blocks
 * into this rank's halo segments, scatter_add_rows (its adjoint) adds the peers' halo-segment gradients onto the listed rows
 * of this rank's block.  lists: int32 row ids inside the owner's block; the per-peer arrays are DEVICE int64[world]
 * (list_ptr: [world + 1]); see csrc/peer_kernels.cu. */
int msha_peer_gather_rows(float* dst, const uint64_t* src_tab, int world, int rank, const int32_t* lists,
                          const int64_t* list_beg, const int64_t* list_end, const int64_t* list_first,
                          const int64_t* local_off, int64_t max_rows_per_peer, int64_t C, int max_ctas, void* stream);
int msha_peer_scatter_add_rows(float* dst, const uint64_t* src_tab, int world, int rank, const int32_t* lists,
                               const int64_t* list_ptr, const int64_t* remote_off, int64_t max_rows_per_peer, int64_t C,
                               int max_ctas, void* stream);
/* all-reduce (sum, rank order) of a small fp64 vector living at offset_bytes of every rank's peer-mapped buffer -- the
 * BatchNorm column statistics of a partitioned node axis; waits for every peer's flag first */
int msha_peer_allreduce_f64(double* out, const uint64_t* src_tab, int64_t offset_bytes, int64_t n, const uint32_t* flags,
                            const uint64_t* flag_tab, int world, int rank, int wait_ch, uint32_t wait_val,
                            uint64_t timeout_ns, int32_t* status, void* stream);

/* ==== callers either side of the path ==== */

/* ---- 8f-1 LLP distillation read-outs.  KD_cosine(s, t) = 1 - mean_p cos(s[idx_s[p]], t[idx_t[p]]) replaces
 *      `1 - cosine_similarity(h[source_index], t_h[source_index].detach(), dim=-1).mean()` (LLP.py:34-35,236) with the
 *      row gathers fused (idx NULL -> row p); torch semantics: each norm clamped at eps = 1e-8.  cos: float[P] saved for
 *      the backward; loss / gout: float[1] on the device; ds / dt are accumulated into, either may be NULL. ---- */
size_t msha_loss_workspace_bytes(void);
int msha_kd_cosine_fwd(const float* s, const float* t, const int64_t* idx_s, const int64_t* idx_t, int64_t P, int64_t C,
                       int64_t n_s, int64_t n_t, float eps, float* cosv, float* loss, int32_t* status, void* ws,
                       size_t ws_bytes, void* stream);
int msha_kd_cosine_bwd(const float* s, const float* t, const int64_t* idx_s, const int64_t* idx_t, int64_t P, int64_t C,
                       int64_t n_s, int64_t n_t, float eps, const float* cosv, const float* gout, float* ds, float* dt,
                       void* stream);
/* torch.nn.MSELoss() (mean over all n elements), LLP.py:221,237; da / db are overwritten, either may be NULL */
int msha_mse_loss_fwd(const float* a, const float* b, int64_t n, float* loss, void* ws, size_t ws_bytes, void* stream);
int msha_mse_loss_bwd(const float* a, const float* b, int64_t n, const float* gout, float* da, float* db, void* stream);

/* ---- 8f-3 GraphSAGE baseline: out[b, :] = adj[src[b], :] * x[b, :] (SGAE.py:53) over the CSR of adj (val NULL -> 1);
 *      the map is diagonal: the same call with x = d out is the backward ---- */
int msha_csr_rows_mul(const int32_t* rowptr, const int32_t* col, const float* val, const int64_t* src, int64_t B,
                      int64_t n_rows, int64_t M, const float* x, float* out, void* stream);

/* ---- 8f-4 attention export: per item (CSR row, or CSC column through perm) the maximum of the per-edge attention
 *      w[slot * H + head] (head < 0: mean over heads), the smallest slot attaining it (-1: empty item) and the number of
 *      ties -- replaces argwhere(row == max(row)) over the dense (N, M) / (N, N) dumps, Explainer.py:25-30 ---- */
int msha_segment_argmax(const int32_t* ptr, const int32_t* perm, const float* w, int H, int head, int64_t n_items,
                        float* vmax, int32_t* first, int32_t* ties, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSHA_B200_H */
