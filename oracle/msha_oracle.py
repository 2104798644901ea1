"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the MSHA-GNN hot path (the parity oracle).

This file restates, in sparse (edge-list) form and in fp64 on the CPU, the arithmetic of the
reference layers.  Every function cites the reference ``file:line`` it follows.  It is *never*
imported by the product package ``msha_gnn_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

Pinning: the reference ships no tests / golden vectors, so the oracle is
pinned against outputs of the reference classes themselves, executed in the build container
by ``oracle/make_golden.py`` and committed under ``tests/golden/`` (``tests/test_oracle.py``
checks every function here against them; max abs err is ~1e-6 because the reference is fp32).
Exceptions that stay "parity unpinned": the Philox4x32-10 negative sampler / dropout stream
(no counterpart exists in the reference) -- it is pinned to the published
Random123 known-answer vectors instead.

Integer work (CSR/CSC build) is numpy; floating point work is torch-CPU (fp64 by default) so
that ``torch.autograd`` provides the gradient oracle for the same restated formulas.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NEG_MASK = -9e15  # GAT.py:29, Ours.py:66


# --------------------------------------------------------------------------------------
# Graph build (integer exact)                                   dataset.py:279-296, model.py:95-100
# --------------------------------------------------------------------------------------
def csr_from_dense(adj: np.ndarray):
    """Neighbour sets ``{j : adj[i,j] > 0}`` in row-major order (== ``(adj>0).nonzero()``).

    The layers only look at ``adj > 0`` (GAT.py:30, Ours.py:67).  Returns int32 rowptr/col and
    the fp32 values at those positions.
    """
    adj = np.asarray(adj)
    mask = adj > 0
    r, c = np.nonzero(mask)  # row-major: rows ascending, cols ascending within a row
    rowptr = np.zeros(adj.shape[0] + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=rowptr[1:])
    return rowptr.astype(np.int32), c.astype(np.int32), adj[r, c].astype(np.float32)


def csr_from_coo(src: np.ndarray, dst: np.ndarray, n_rows: int, n_cols: int):
    """COO flow records -> coalesced CSR; multiplicity becomes the value (dataset.py:286-288).

    Equivalent to ``inter[s, r] += 1`` per record followed by ``nonzero()``.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    if src.size and (src.min() < 0 or src.max() >= n_rows or dst.min() < 0 or dst.max() >= n_cols):
        raise IndexError("edge endpoint out of range")
    key = src * np.int64(n_cols) + dst
    uniq, cnt = np.unique(key, return_counts=True)
    r = (uniq // n_cols).astype(np.int64)
    c = (uniq % n_cols).astype(np.int32)
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    np.cumsum(rowptr, out=rowptr)
    return rowptr.astype(np.int32), c, cnt.astype(np.float32)


def csc_from_csr(rowptr: np.ndarray, col: np.ndarray, n_cols: int):
    """Transpose structure: colptr, row index and ``perm`` (CSC slot -> CSR slot), rows ascending
    within a column (stable counting sort by column)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    n_rows = rowptr.size - 1
    row_of = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(rowptr))
    perm = np.argsort(col, kind="stable")
    colptr = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(colptr, col + 1, 1)
    np.cumsum(colptr, out=colptr)
    return colptr.astype(np.int32), row_of[perm].astype(np.int32), perm.astype(np.int32)


def group_adjacency_dense(group_ids: np.ndarray) -> np.ndarray:
    """``A[i,j] = 1 iff group(i) == group(j)`` incl. the diagonal (dataset.py:267-275)."""
    g = np.asarray(group_ids)
    return (g[:, None] == g[None, :]).astype(np.float32)


def normalize_adjacency_dense(adj: torch.Tensor) -> torch.Tensor:
    """``A @ D^-1/2 @ D^-1/2`` with ``D = colsum`` (model.py:95-100) == ``A[:, j] / colsum[j]``;
    a zero column makes everything NaN through inf*0 in the dense ``mm`` -- reproduced."""
    deg = adj.sum(dim=0)
    d = deg.pow(-0.5)
    out = adj * (d * d)[None, :]
    if bool((deg == 0).any()):
        out = torch.full_like(adj, float("nan"))
    return out


def normalize_csr_values(val: np.ndarray, col: np.ndarray, n_cols: int) -> np.ndarray:
    """Column-normalised CSR values (sparse form of model.py:95-100), fp32 like the reference."""
    colsum = np.zeros(n_cols, dtype=np.float32)
    np.add.at(colsum, col, val.astype(np.float32))
    d = np.power(colsum, np.float32(-0.5), dtype=np.float32)
    return (val.astype(np.float32) * d[col]) * d[col]


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _t(x, dtype=torch.float64):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def _rows_of(rowptr) -> torch.Tensor:
    rp = torch.as_tensor(np.asarray(rowptr), dtype=torch.int64)
    return torch.repeat_interleave(torch.arange(rp.numel() - 1), rp[1:] - rp[:-1])


def attention_edges(rowptr, col, n_cols: int):
    """Edge list that the masked softmax effectively runs over.

    A row with no neighbour has every logit == -9e15, so ``softmax`` returns the uniform
    ``1/M`` over *all* M columns (GAT.py:29-31).  Such rows therefore get M explicit "masked"
    edges.  Returns (row, col, masked) int64/int64/bool tensors in row-major order.
    """
    rp = np.asarray(rowptr, dtype=np.int64)
    c = np.asarray(col, dtype=np.int64)
    deg = np.diff(rp)
    rows, cols, msk = [], [], []
    r_of = np.repeat(np.arange(deg.size), deg)
    iso = np.nonzero(deg == 0)[0]
    if iso.size == 0:
        return (torch.from_numpy(r_of), torch.from_numpy(c), torch.zeros(c.size, dtype=torch.bool))
    # merge, keeping row-major order
    r_iso = np.repeat(iso, n_cols)
    c_iso = np.tile(np.arange(n_cols, dtype=np.int64), iso.size)
    r_all = np.concatenate([r_of, r_iso])
    c_all = np.concatenate([c, c_iso])
    m_all = np.concatenate([np.zeros(c.size, bool), np.ones(c_iso.size, bool)])
    order = np.lexsort((c_all, r_all))
    return (torch.from_numpy(r_all[order]), torch.from_numpy(c_all[order]),
            torch.from_numpy(m_all[order]))


def segment_softmax(logits: torch.Tensor, seg: torch.Tensor, n_seg: int) -> torch.Tensor:
    """Row softmax over an edge list (F.softmax(dim=1) on the masked dense matrix)."""
    shape = (n_seg,) + tuple(logits.shape[1:])
    idx = seg.view(-1, *([1] * (logits.dim() - 1))).expand_as(logits)
    mx = torch.full(shape, -float("inf"), dtype=logits.dtype).scatter_reduce(
        0, idx, logits.detach(), reduce="amax", include_self=True)
    p = torch.exp(logits - mx[seg])
    s = torch.zeros(shape, dtype=logits.dtype).index_add(0, seg, p)
    return p / s[seg]


def batchnorm1d(x, weight, bias, running_mean=None, running_var=None, training=True, eps=1e-5):
    """nn.BatchNorm1d over the node axis (Ours.py:50-52,100-101): batch statistics with the
    biased variance in training mode, running statistics in eval mode."""
    if training:
        mean = x.mean(dim=0)
        var = x.var(dim=0, unbiased=False)
    else:
        mean, var = running_mean, running_var
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def batchnorm1d_running_update(x, running_mean, running_var, momentum=0.1):
    """Side effect of a training-mode BN call: running stats use the *unbiased* variance."""
    n = x.shape[0]
    mean = x.mean(dim=0)
    var_u = x.var(dim=0, unbiased=True) if n > 1 else x.var(dim=0, unbiased=False)
    return ((1 - momentum) * running_mean + momentum * mean,
            (1 - momentum) * running_var + momentum * var_u)


# --------------------------------------------------------------------------------------
# a-1  GraphAttentionLayer (degenerate)                                         GAT.py:20-35
# --------------------------------------------------------------------------------------
def graph_attention_layer(x, W, a, rowptr, col, att_mask=None):
    """``elu(att * h)`` with ``att[i,j] = 1/deg(i)`` on neighbours (uniform 1/M when deg==0).

    The logit ``lrelu(h_i.(a[:M]+a[M:]))`` is constant along j (both halves of the concat are
    ``repeat_h``, GAT.py:24-27), so the masked softmax is uniform over the row's neighbours.
    The logit is still carried through ``segment_softmax`` so that d/da (== 0) flows.
    ``att_mask`` optionally holds the dense (N,M) dropout multiplier (GAT.py:32).
    """
    x, W, a = _t(x), _t(W), _t(a)
    h = x @ W                                                      # GAT.py:21
    N, M = h.shape
    r, c, masked = attention_edges(rowptr, col, M)
    e_row = F.leaky_relu(h @ (a[:M, 0] + a[M:, 0]), 0.2)           # GAT.py:27
    logit = torch.where(masked, torch.full_like(e_row[r], NEG_MASK), e_row[r])
    att_e = segment_softmax(logit, r, N)                           # GAT.py:29-31
    att = torch.zeros(N, M, dtype=h.dtype).index_put((r, c), att_e)
    if att_mask is not None:
        att = att * _t(att_mask)
    return F.elu(att * h)                                          # GAT.py:34-35


def gat_model(features, layer_params, out_params, rowptr, col):
    """GAT.forward (GAT.py:53-58) in eval / p=0 mode: heads concat -> out_att -> elu -> log_softmax.
    ``layer_params`` = [(W,a)] per head, ``out_params`` = (W,a).  ELU is applied twice on the
    out_att branch (inside the layer GAT.py:35 and outside GAT.py:57)."""
    x = _t(features)
    hs = [graph_attention_layer(x, W, a, rowptr, col) for (W, a) in layer_params]
    x = torch.cat(hs, dim=1)
    x = F.elu(graph_attention_layer(x, out_params[0], out_params[1], rowptr, col))
    return F.log_softmax(x, dim=1)


# --------------------------------------------------------------------------------------
# a-3/a-4  inter-scale bipartite GAT attention          Ours.py:57-69, Ablation.py:262-271
# --------------------------------------------------------------------------------------
def inter_attention(h1, h2, a, rowptr, col):
    """alpha over the attention edge list: e12[i,j] = lrelu(a[:d'].h1[j] + a[d':].h2[i])
    (Ours.py:64-65), masked row softmax (Ours.py:66-68).  Returns (alpha_e, r, c, masked)."""
    d = h1.shape[1]
    M = h1.shape[0]
    N = h2.shape[0]
    r, c, masked = attention_edges(rowptr, col, M)
    s_nbr = h1 @ a[:d, 0]
    s_self = h2 @ a[d:, 0]
    e = F.leaky_relu(s_nbr[c] + s_self[r], 0.2)
    e = torch.where(masked, torch.full_like(e, NEG_MASK), e)
    return segment_softmax(e, r, N), r, c, masked


def dense_from_edges(vals, r, c, N, M):
    return torch.zeros(N, M, dtype=vals.dtype).index_put((r, c), vals)


def ours_layer3(S, R, p, rowptr, col, training=True, alpha_mask=None, return_alpha=False):
    """OursLayer3.forward (Ablation.py:260-277).  ``p`` is a dict of the layer's parameters /
    buffers under their state_dict names.  ``alpha_mask`` optional dense (N,M) dropout multiplier."""
    S, R = _t(S), _t(R)
    W1, W2, a = _t(p["W1"]), _t(p["W2"]), _t(p["a"])
    h1 = R @ W1                                                    # :262
    h2 = S @ W2                                                    # :263
    N, M = h2.shape[0], h1.shape[0]
    alpha, r, c, _ = inter_attention(h1, h2, a, rowptr, col)       # :266-270
    if alpha_mask is not None:
        alpha = alpha * _t(alpha_mask)[r, c]                       # :271
    v_in = torch.zeros(M, h2.shape[1], dtype=h2.dtype).index_add(0, c, alpha[:, None] * h2[r])
    u_in = torch.zeros(N, h1.shape[1], dtype=h1.dtype).index_add(0, r, alpha[:, None] * h1[c])
    v = F.leaky_relu(batchnorm1d(v_in, _t(p["bn1.weight"]), _t(p["bn1.bias"]),
                                 _t(p["bn1.running_mean"]), _t(p["bn1.running_var"]), training), 0.2)
    u = F.leaky_relu(batchnorm1d(u_in, _t(p["bn2.weight"]), _t(p["bn2.bias"]),
                                 _t(p["bn2.running_mean"]), _t(p["bn2.running_var"]), training), 0.2)
    out = F.elu(u @ v.t())                                         # :276-277
    if return_alpha:
        return out, dense_from_edges(alpha, r, c, N, M)
    return out


def _intra_scales(h2, p, src, city_mask, prov_mask, alpha_dense_rows, joint: bool):
    """Intra-scale attention of the batch rows (Ours.py:71-90 joint normaliser;
    Ablation.py:194-197 separate softmax when ``joint`` is False).

    ``city_mask`` / ``prov_mask``: bool (B,N) == ``city_adj[source_index] > 0``.
    ``alpha_dense_rows``: (B,M) == ``attention_inter[source_index]`` (after dropout).
    Logits are row constants: t3[b] = lrelu(h2[src_b].(a3[:d']+a3[d':])) (Ours.py:71-75).
    """
    d = h2.shape[1]
    a3, a4 = _t(p["a3"]), _t(p["a4"])
    h2b = h2[src]
    t3 = F.leaky_relu(h2b @ (a3[:d, 0] + a3[d:, 0]), 0.2)
    t4 = F.leaky_relu(h2b @ (a4[:d, 0] + a4[d:, 0]), 0.2)
    neg = torch.full((1,), NEG_MASK, dtype=h2.dtype)
    att3 = torch.where(city_mask, t3[:, None], neg)
    att4 = torch.where(prov_mask, t4[:, None], neg)
    if joint:
        # no max-subtraction in the reference (overflow -> inf/NaN reproduced)  Ours.py:84-89
        total = (torch.exp(att3).sum(1, keepdim=True) + torch.exp(att4).sum(1, keepdim=True)
                 + torch.exp(alpha_dense_rows).sum(1, keepdim=True))
        att3 = torch.exp(att3) / total
        att4 = torch.exp(att4) / total
    else:
        att3 = F.softmax(att3, dim=1)                              # Ablation.py:194
        att4 = F.softmax(att4, dim=1)                              # Ablation.py:196
    return att3, att4, h2b


def ours_layer(S, R, p, rowptr, col, city_adj, prov_adj, src, training=True, variant=1,
               return_coeffs=False):
    """OursLayer.forward (Ours.py:54-109 == Ablation.py:35-83) for ``variant=1`` and
    OursLayer2.forward (Ablation.py:165-205) for ``variant=2``; dropout inactive (p=0 / eval).

    ``city_adj`` / ``prov_adj``: dense (N,N) arrays, only ``> 0`` is consulted (Ours.py:81-82).
    Duplicate entries of ``src`` add twice in IntraNC (Ours.py:99).
    """
    S, R = _t(S), _t(R)
    W1, W2, a = _t(p["W1"]), _t(p["W2"]), _t(p["a"])
    h1 = R @ W1
    h2 = S @ W2
    N, M = h2.shape[0], h1.shape[0]
    src = torch.as_tensor(np.asarray(src), dtype=torch.int64)
    alpha, r, c, _ = inter_attention(h1, h2, a, rowptr, col)
    alpha_dense = dense_from_edges(alpha, r, c, N, M)
    cm = torch.as_tensor(np.asarray(city_adj))[src] > 0
    pm = torch.as_tensor(np.asarray(prov_adj))[src] > 0
    att3, att4, h2b = _intra_scales(h2, p, src, cm, pm, alpha_dense[src], joint=(variant == 1))
    inter_rc = alpha_dense @ h1                                    # Ours.py:98
    intra_nc = att3.t() @ h2b + att4.t() @ h2b                     # Ours.py:99
    v_in = alpha_dense.t() @ h2                                    # Ours.py:100
    v = F.leaky_relu(batchnorm1d(v_in, _t(p["bn1.weight"]), _t(p["bn1.bias"]),
                                 _t(p["bn1.running_mean"]), _t(p["bn1.running_var"]), training), 0.2)
    u = F.leaky_relu(batchnorm1d(inter_rc + intra_nc, _t(p["bn2.weight"]), _t(p["bn2.bias"]),
                                 _t(p["bn2.running_mean"]), _t(p["bn2.running_var"]), training), 0.2)
    out = F.elu(u @ v.t())                                         # Ours.py:108-109
    if return_coeffs:
        return out, alpha_dense, att3, att4
    return out


def msha_model(Sfeat, Rfeat, head_params, out_params, rowptr, col, city_adj=None, prov_adj=None,
               src=None, training=True, variant=3):
    """Ours / ablation2 / ablation3 .forward (Ours.py:160-167, Ablation.py:224-231,295-301) with
    dropout inactive: heads concat -> out_att (a-1 with F = M*H) -> elu -> log_softmax."""
    hs = []
    for p in head_params:
        if variant == 3:
            hs.append(ours_layer3(Sfeat, Rfeat, p, rowptr, col, training))
        else:
            hs.append(ours_layer(Sfeat, Rfeat, p, rowptr, col, city_adj, prov_adj, src, training,
                                 variant=variant))
    x = torch.cat(hs, dim=1)
    x = F.elu(graph_attention_layer(x, out_params[0], out_params[1], rowptr, col))
    return F.log_softmax(x, dim=1)


def ablation1_model(Sfeat, Rfeat, p, rowptr, col, city_adj, prov_adj, src, training=True):
    """ablation1.forward (Ablation.py:130-136): single OursLayer, elu again, log_softmax."""
    x = ours_layer(Sfeat, Rfeat, p, rowptr, col, city_adj, prov_adj, src, training, variant=1)
    return F.log_softmax(F.elu(x), dim=1)


# --------------------------------------------------------------------------------------
# generic multi-head GAT layer (SURVEY.md section 8a "generalisation rule")
# --------------------------------------------------------------------------------------
def gat_layer(x, W, a_nbr, a_self, rowptr, col, heads: int, concat=True, apply_elu=True,
              return_alpha=False):
    """a-3's inter-scale block with S = R (Ablation.py:262-271 + ``alpha @ h1`` :274), H heads:
    Wh = x@W (N,H*d'); e[i,j,h] = lrelu(a_nbr[h].Wh[j,h] + a_self[h].Wh[i,h]); row softmax over
    adj[i,:]>0; out[i,h] = sum_j alpha[i,j,h] Wh[j,h]; heads concatenated or averaged; ELU."""
    x, W, a_nbr, a_self = _t(x), _t(W), _t(a_nbr), _t(a_self)
    N = x.shape[0]
    Wh = (x @ W).view(N, heads, -1)
    r, c, masked = attention_edges(rowptr, col, N)
    s_nbr = (Wh * a_nbr[None]).sum(-1)
    s_self = (Wh * a_self[None]).sum(-1)
    e = F.leaky_relu(s_nbr[c] + s_self[r], 0.2)
    e = torch.where(masked[:, None], torch.full_like(e, NEG_MASK), e)
    alpha = segment_softmax(e, r, N)                               # (E,H)
    out = torch.zeros_like(Wh).index_add(0, r, alpha[:, :, None] * Wh[c])
    out = out.reshape(N, -1) if concat else out.mean(dim=1)
    if apply_elu:
        out = F.elu(out)
    return (out, alpha) if return_alpha else out


def gat_layer_rows(x_nodes, W, a_nbr, a_self, self_idx, nbr_ptr, nbr_idx, heads: int, apply_elu=True):
    """``gat_layer`` restricted to a sample of rows (parity at benchmark scale, SURVEY.md section 8c: "check the GPU
    result on sampled rows"): ``x_nodes`` (n_sub, F) holds the input features of every node the sampled rows touch,
    row k of the sample is node ``self_idx[k]`` and attends to ``nbr_idx[nbr_ptr[k]:nbr_ptr[k+1]]`` (indices into
    ``x_nodes``).  Same arithmetic as above (Ablation.py:262-271 + ``alpha @ h1`` :274); heads concatenated."""
    x, W, a_nbr, a_self = _t(x_nodes), _t(W), _t(a_nbr), _t(a_self)
    self_idx = torch.as_tensor(np.asarray(self_idx), dtype=torch.int64)
    nbr_ptr = np.asarray(nbr_ptr, dtype=np.int64)
    c = torch.as_tensor(np.asarray(nbr_idx), dtype=torch.int64)
    K = self_idx.numel()
    Wh = (x @ W).view(x.shape[0], heads, -1)
    r = torch.repeat_interleave(torch.arange(K), torch.as_tensor(np.diff(nbr_ptr)))
    s_nbr = (Wh * a_nbr[None]).sum(-1)
    s_self = (Wh[self_idx] * a_self[None]).sum(-1)
    e = F.leaky_relu(s_nbr[c] + s_self[r], 0.2)
    alpha = segment_softmax(e, r, K)
    out = torch.zeros((K,) + Wh.shape[1:], dtype=Wh.dtype).index_add(0, r, alpha[:, :, None] * Wh[c]).reshape(K, -1)
    return F.elu(out) if apply_elu else out


# --------------------------------------------------------------------------------------
# a-6  HGANE.GraphAttentionLayer                                                HGANE.py:37-76
# --------------------------------------------------------------------------------------
def hgane_layer(p, adj_inter, adj_intra, src, training=True):
    """HGANE.GraphAttentionLayer.forward, dropout inactive.  ``p``: state_dict-named tensors.
    No max-subtraction; a batch row without an inter neighbour gives 0/0 = NaN (HGANE.py:61-69)."""
    src = torch.as_tensor(np.asarray(src), dtype=torch.int64)
    A_intra = _t(adj_intra)[src[:, None], src]
    A_inter = _t(adj_inter)[src]
    E_r = _t(p["recipient_embedding"])
    E_s = _t(p["source_embedding"])[src]
    W1, W2 = _t(p["W1.weight"]), _t(p["W2.weight"])
    a12, a3 = _t(p["a12.weight"])[0], _t(p["a3.weight"])[0]
    h1 = E_r @ W1.t()
    h2 = E_s @ W2.t()
    d = h1.shape[1]
    e12 = F.leaky_relu((h1 @ a12[:d])[None, :] + (h2 @ a12[d:])[:, None], 0.2)   # :46-47
    e3 = F.leaky_relu((h2 @ a3[:d])[:, None] + (h2 @ a3[d:])[None, :], 0.2)      # :49-52
    neg = torch.full((1,), NEG_MASK, dtype=h1.dtype)
    att_inter = torch.where(A_inter > 0, e12, neg)
    att_intra = torch.where(A_intra > 0, e3, neg)
    sum_county = torch.exp(att_intra).sum(1, keepdim=True) + torch.exp(att_inter).sum(1, keepdim=True)
    att_intra = torch.exp(att_intra) / sum_county
    att_inter = torch.exp(att_inter) / torch.exp(att_inter).sum(1, keepdim=True)
    u_in = (att_inter @ E_r) @ W1.t() + (att_intra @ E_s) @ W2.t()
    v_in = (att_inter.t() @ E_s) @ W1.t()
    u = F.leaky_relu(batchnorm1d(u_in, _t(p["bn1.weight"]), _t(p["bn1.bias"]),
                                 _t(p["bn1.running_mean"]), _t(p["bn1.running_var"]), training), 0.2)
    v = F.leaky_relu(batchnorm1d(v_in, _t(p["bn2.weight"]), _t(p["bn2.bias"]),
                                 _t(p["bn2.running_mean"]), _t(p["bn2.running_var"]), training), 0.2)
    return F.elu(u @ v.t())


# --------------------------------------------------------------------------------------
# a-7  LinkPredictor                                                          LLP.py:104-115
# --------------------------------------------------------------------------------------
def link_predictor(x_i, x_j, weights, biases, predictor="mlp"):
    """``sigmoid(relu(lin(x_i*x_j)))`` over ``lins[:-1]`` -- the final Linear is commented out
    (LLP.py:111) so the output is (P, hidden); 'inner' -> sigmoid(sum_c x) (LLP.py:112-113).
    ``weights`` / ``biases`` list *all* lins (the last one is ignored, as in the reference)."""
    x = _t(x_i) * _t(x_j)
    if predictor == "mlp":
        for W, b in list(zip(weights, biases))[:-1]:
            x = F.relu(x @ _t(W).t() + _t(b))
    elif predictor == "inner":
        x = x.sum(dim=-1)
    return torch.sigmoid(x)


def pair_dot(u, v, src, dst):
    """Bilinear read-out for sampled pairs: ``elu(u_i . v_j)`` == entries of Ours.py:108-109."""
    u, v = _t(u), _t(v)
    return F.elu((u[src] * v[dst]).sum(-1))


# --------------------------------------------------------------------------------------
# model.GraphConvolution                                                      model.py:34-41
# --------------------------------------------------------------------------------------
def graph_convolution(x, weight, bias, rowptr, col, val, n_cols):
    """``adj.T @ (x @ W) + bias`` with a 0-dim bias (model.py:23,36-39); adj as value CSR."""
    x, weight = _t(x), _t(weight)
    support = x @ weight
    r = _rows_of(rowptr)
    c = torch.as_tensor(np.asarray(col), dtype=torch.int64)
    v = _t(val)
    out = torch.zeros(n_cols, support.shape[1], dtype=support.dtype).index_add(
        0, c, v[:, None] * support[r])
    return out + _t(bias) if bias is not None else out


# --------------------------------------------------------------------------------------
# LLP student + distillation step, GCN / GraphSAGE baselines, attention export
# --------------------------------------------------------------------------------------
def mlp(x, weights, biases):
    """``LLP.MLP.forward`` with ``norm_type='none'`` in eval / dropout 0 (LLP.py:75-84): relu on all but the last."""
    h = _t(x)
    for k, (W, b) in enumerate(zip(weights, biases)):
        h = h @ _t(W).t() + _t(b)
        if k != len(weights) - 1:
            h = F.relu(h)
    return h


def kd_cosine(s, t, idx_s=None, idx_t=None, eps=1e-8, detach_teacher=True):
    """``KD_cosine`` (LLP.py:34-35) = ``1 - cosine_similarity(s, t.detach(), dim=-1).mean()``; torch divides each
    vector by ``max(||.||, eps)`` (clamped outside autograd) before the dot product."""
    s, t = _t(s), _t(t)
    if detach_teacher:
        t = t.detach()
    if idx_s is not None:
        s = s[torch.as_tensor(np.asarray(idx_s), dtype=torch.int64)]
    if idx_t is not None:
        t = t[torch.as_tensor(np.asarray(idx_t), dtype=torch.int64)]
    ns = torch.linalg.vector_norm(s, 2, dim=-1, keepdim=True)
    nt = torch.linalg.vector_norm(t, 2, dim=-1, keepdim=True)
    ns = ns + (ns.detach().clamp_min(eps) - ns.detach())          # value clamped, gradient of the norm kept
    nt = nt + (nt.detach().clamp_min(eps) - nt.detach())
    return 1 - ((s / ns) * (t / nt)).sum(-1).mean()


def mse_loss(a, b):
    """``torch.nn.MSELoss()`` (LLP.py:221,237): mean over every element."""
    return ((_t(a) - _t(b)) ** 2).mean()


def llp_step_loss(features, student, predictor, teacher_heads, teacher_out, teacher_pred, rowptr, col, src, rec,
                  True_label=10.0, KD_f=0.1, KD_p=100.0):
    """Loss of one LLP step (LLP.py:230-237), dropout 0.  ``student`` / ``predictor`` / ``teacher_pred`` are
    ``(weights, biases)`` lists, ``teacher_heads`` a list of ``(W, a)``, ``teacher_out`` one ``(W, a)``."""
    src = torch.as_tensor(np.asarray(src), dtype=torch.int64)
    rec = torch.as_tensor(np.asarray(rec), dtype=torch.int64)
    feats = _t(features)
    h = mlp(feats, *student)
    x = torch.cat([graph_attention_layer(feats, W, a, rowptr, col) for W, a in teacher_heads], dim=1)   # LLP.py:165
    t_h = F.log_softmax(F.elu(graph_attention_layer(x, teacher_out[0], teacher_out[1], rowptr, col)), dim=1)
    output = link_predictor(h[src], h[rec], *predictor)
    label_loss = -(output[torch.arange(src.numel()), rec]).mean()                                      # LLP.py:235
    t_out = link_predictor(t_h[src], t_h[rec], *teacher_pred).detach()
    kd_f = kd_cosine(h[src], t_h[src])
    kd_p = mse_loss(output, t_out)
    return True_label * label_loss + KD_f * kd_f + KD_p * kd_p, dict(label_loss=label_loss, kd_f=kd_f, kd_p=kd_p,
                                                                     h=h, t_h=t_h, output=output, t_out=t_out)


def spmm_rows(x, rowptr, col, val, n_rows):
    """``adj @ x`` with adj as value CSR (GCN's second layer ``gc2(x, adj.t())``, model.py:37,61)."""
    x = _t(x)
    r = _rows_of(rowptr)
    c = torch.as_tensor(np.asarray(col), dtype=torch.int64)
    return torch.zeros(n_rows, x.shape[1], dtype=x.dtype).index_add(0, r, _t(val)[:, None] * x[c])


def gcn_model(features, p, rowptr, col, val, n_rows, n_cols):
    """``GCN.forward`` (model.py:58-64), dropout 0: relu(gc1(features, adj)) -> relu(gc2(., adj.t())) -> log_softmax."""
    x = F.relu(graph_convolution(features, p["gc1.weight"], p["gc1.bias"], rowptr, col, val, n_cols))
    x = F.relu(spmm_rows(x @ _t(p["gc2.weight"]), rowptr, col, val, n_rows) + _t(p["gc2.bias"]))
    return F.log_softmax(x, dim=1)


def graphsage_model(p, src, rowptr, col, val, n_cols):
    """``GraphSAGE.forward`` (SGAE.py:49-56): relu(linear1(S[src])) * adj[src] -> relu(linear2) -> log_softmax."""
    src = np.asarray(src, dtype=np.int64)
    x = _t(p["Sfeatures"])[torch.as_tensor(src)]
    x = F.relu(x @ _t(p["linear1.weight"]).t() + _t(p["linear1.bias"]))
    rowptr = np.asarray(rowptr, dtype=np.int64)
    rows = torch.zeros(src.size, n_cols, dtype=x.dtype)
    v = _t(val)
    for b, r in enumerate(src):                                      # adj[source_index], SGAE.py:53
        e0, e1 = int(rowptr[r]), int(rowptr[r + 1])
        rows[b, torch.as_tensor(np.asarray(col[e0:e1]), dtype=torch.int64)] = v[e0:e1]
    x = F.relu((rows * x) @ _t(p["linear2.weight"]).t() + _t(p["linear2.bias"]))
    return F.log_softmax(x, dim=1)


def explainer_argmax(dense: np.ndarray):
    """``[np.argwhere(row == np.max(row)).flatten().tolist() for row in Coeff]`` (Explainer.py:25-30)."""
    return [np.argwhere(row == np.max(row)).flatten().tolist() for row in np.asarray(dense)]


# --------------------------------------------------------------------------------------
# a-8  loss read-out                                                          train.py:229
# --------------------------------------------------------------------------------------
def nll_readout(logp, src, rec):
    logp = _t(logp)
    src = torch.as_tensor(np.asarray(src), dtype=torch.int64)
    rec = torch.as_tensor(np.asarray(rec), dtype=torch.int64)
    return -(logp[src, rec]).mean()


# --------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3")
# builder-defined counter-based stream for negative sampling and attention dropout.
# The reference has neither a seeded sampler nor reproducible dropout -> "parity unpinned";
# pinned to the Random123 known-answer vectors in tests/test_oracle.py.
# --------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: (...,4) uint32, key: (...,2) uint32 -> (...,4) uint32."""
    c = np.array(ctr, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(np.asarray(key, dtype=np.uint32), c.shape[:-1] + (2,)), copy=True)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c[..., 0].astype(np.uint64)
            p1 = _PHILOX_M1 * c[..., 2].astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c = np.stack([hi1 ^ c[..., 1] ^ k[..., 0], lo1, hi0 ^ c[..., 3] ^ k[..., 1], lo0], axis=-1)
            k = np.stack([k[..., 0] + _PHILOX_W0, k[..., 1] + _PHILOX_W1], axis=-1)
    return c


def _philox_words(seed: int, stream: int, n: int) -> np.ndarray:
    """Word ``i`` of the stream: block ``i//4`` with counter (blk_lo, blk_hi, stream, 0) and key
    (seed_lo, seed_hi); lane ``i%4``."""
    nblk = (n + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    ctr = np.stack([(blk & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                    (blk >> np.uint64(32)).astype(np.uint32),
                    np.full(nblk, stream, dtype=np.uint32),
                    np.zeros(nblk, dtype=np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key).reshape(-1)[:n]


def negative_sample(seed: int, n_pairs: int, n_src: int, n_dst: int):
    """Uniform negative pairs: pair p uses words 2p (source) and 2p+1 (destination) of stream 1,
    mapped with the multiply-shift ``(word * n) >> 32`` (no rejection step)."""
    w = _philox_words(seed, 1, 2 * n_pairs).astype(np.uint64)
    s = (w[0::2] * np.uint64(n_src)) >> np.uint64(32)
    d = (w[1::2] * np.uint64(n_dst)) >> np.uint64(32)
    return s.astype(np.int64), d.astype(np.int64)


DROP_EPOCH_STRIDE = 0x9E3779B97F4A7C15   # 2^64 / golden ratio: the per-replay key stride (csrc/common.cuh drop_seed_eff)


def dropout_keep_mask(seed: int, n: int, p: float, stream: int = 2, epoch: int = 0) -> np.ndarray:
    """Element i is kept iff word i of ``stream`` >= floor(p * 2^32); the Philox key is
    ``seed + epoch * DROP_EPOCH_STRIDE (mod 2^64)`` -- ``epoch`` is the CUDA-graph replay counter, 0 outside graphs."""
    thr = np.uint64(min(int(p * 4294967296.0), 0xFFFFFFFF))
    key = (int(seed) + int(epoch) * DROP_EPOCH_STRIDE) & 0xFFFFFFFFFFFFFFFF
    return _philox_words(key, stream, n).astype(np.uint64) >= thr
