"""Intra-scale (city / province) attention of OursLayer / OursLayer2 (Ours.py:71-90,99; Ablation.py:184-200)
without the dense (B, N) / (N, N) tensors.

The intra logits are row constants (both halves of the concat are the batch row's own embedding,
Ours.py:71-78), so ``attention3[b, n] = coef3[b]`` for every n in the batch row's city list:

    SUM[b]   = |city_b| e^{t3_b} + |prov_b| e^{t4_b} + sum_{j<M} exp(alpha_drop[src_b, j])     Ours.py:84-86
    coef3[b] = e^{t3_b} / SUM[b]                       (no max-subtraction -- overflow reproduced)
    IntraNC[n] = sum_{b : n in city_b} coef3[b] h2[src_b] + (same for provinces)              Ours.py:99

The (B, H)-sized coefficient algebra is plain tensor arithmetic; the edge / segment work runs in the
``intra_kernels.cu`` kernels.
"""
from __future__ import annotations

import weakref

import torch

from . import functional as Fn
from . import ops
from .graph import Graph, as_graph
from .ops import call, ptr, _stream

I32 = torch.int32
I64 = torch.int64


class GroupLists:
    """Member lists of a block-structured adjacency ``A[i,j] = 1 iff group(i) == group(j)``
    (dataset.py:267-275) -- N integers instead of an (N, N) matrix."""

    def __init__(self, group_ids: torch.Tensor):
        g = group_ids.detach().to(torch.int64).contiguous()
        if not g.is_cuda:
            raise RuntimeError("msha_b200 is CUDA-only: group ids must be a CUDA tensor")
        uniq, inv = torch.unique(g, return_inverse=True)
        order = torch.argsort(inv, stable=True)                 # members ascending inside a group
        counts = torch.bincount(inv, minlength=uniq.numel())
        rowptr = torch.zeros(uniq.numel() + 1, dtype=torch.int64, device=g.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        self.rowptr = rowptr.to(I32)
        self.col = order.to(I32)
        self.row_map = inv.to(I32)
        self.n_nodes = g.numel()
        self.counts_per_node = counts[inv].to(torch.float32)

    def lists(self):
        return self.rowptr, self.col, self.row_map

    def counts(self, src):
        return self.counts_per_node[src]


def _as_lists(adj, n_nodes):
    """dense (N,N) float adjacency / Graph / GroupLists / 1-D group-id vector -> (rowptr, col, row_map, counts_fn, nonempty);
    ``nonempty``: every node is known to have at least one list member (a group always contains the node itself)."""
    if isinstance(adj, GroupLists):
        return adj.rowptr, adj.col, adj.row_map, adj.counts, True
    if isinstance(adj, torch.Tensor) and adj.dim() == 1:
        gl = _group_cache(adj)
        return gl.rowptr, gl.col, gl.row_map, gl.counts, True
    g = as_graph(adj, n_rows=n_nodes, n_cols=n_nodes)
    deg = g.degrees.to(torch.float32)
    return g.rowptr, g.col, None, (lambda src: deg[src]), False


_gl_cache: dict = {}


def _group_cache(t):
    key = (t.data_ptr(), t._version, t.numel())
    hit = _gl_cache.get(key)
    if hit is not None and hit[0]() is t:          # the weakref guards against address reuse by a new tensor
        return hit[1]
    if len(_gl_cache) > 16:
        _gl_cache.clear()
    gl = GroupLists(t)
    _gl_cache[key] = (weakref.ref(t), gl)
    return gl


class _RowsumExp(torch.autograd.Function):
    """T[b,h] = sum_j exp(alpha_drop[src_b, j, h]) over all M columns (Ours.py:86)."""

    @staticmethod
    def forward(ctx, alpha, graph: Graph, src, p, seed):
        rp, _ = graph.attention_csr()
        B, H = src.numel(), alpha.shape[1]
        T = torch.empty((B, H), dtype=torch.float32, device=alpha.device)
        call("msha_rowsum_exp", ptr(rp, I32), ptr(alpha), H, ptr(src, I64), B, graph.n_cols, ptr(T), p, seed, _stream())
        ctx.graph, ctx.p, ctx.seed = graph, p, seed
        ctx.save_for_backward(alpha, src)
        return T

    @staticmethod
    def backward(ctx, dT):
        alpha, src = ctx.saved_tensors
        rp, _ = ctx.graph.attention_csr()
        dalpha = torch.zeros_like(alpha)
        call("msha_rowsum_exp_bwd", ptr(rp, I32), ptr(alpha), alpha.shape[1], ptr(src, I64), src.numel(),
             ptr(dT.contiguous()), ptr(dalpha), ctx.p, ctx.seed, _stream())
        return dalpha, None, None, None, None


class _GroupScatter(torch.autograd.Function):
    """out[n] = sum_{b : n in list(src_b)} coef[b,h] * drop * feat[b,h,:]   (attention3.t() @ h2_, Ours.py:99)."""

    @staticmethod
    def forward(ctx, coef, feat, lists, src, n_nodes, H, D, p, seed, stream_id):
        rowptr, col, row_map = lists
        coef, feat = coef.contiguous(), feat.contiguous()
        out = torch.zeros((n_nodes, H * D), dtype=torch.float32, device=feat.device)
        call("msha_group_scatter_add", ptr(rowptr, I32), ptr(col, I32), ptr(row_map, I32), ptr(src, I64), src.numel(),
             n_nodes, ptr(coef), ptr(feat), H, D, ptr(out), p, seed, stream_id, _stream())
        ctx.lists, ctx.n_nodes, ctx.H, ctx.D, ctx.p, ctx.seed, ctx.stream_id = lists, n_nodes, H, D, p, seed, stream_id
        ctx.save_for_backward(coef, feat, src)
        return out

    @staticmethod
    def backward(ctx, dout):
        coef, feat, src = ctx.saved_tensors
        rowptr, col, row_map = ctx.lists
        H, D = ctx.H, ctx.D
        B = src.numel()
        G = torch.zeros((B, H * D), dtype=torch.float32, device=feat.device)
        call("msha_group_gather_sum", ptr(rowptr, I32), ptr(col, I32), ptr(row_map, I32), ptr(src, I64), B, ctx.n_nodes,
             ptr(dout.contiguous()), H, D, ptr(G), ctx.p, ctx.seed, ctx.stream_id, _stream())
        G3 = G.view(B, H, D)
        dcoef = (G3 * feat.view(B, H, D)).sum(-1)
        dfeat = (G3 * coef.view(B, H, 1)).reshape(B, H * D)
        return dcoef, dfeat, None, None, None, None, None, None, None, None


def intra_scales(layers, graph: Graph, h2, alpha, city_adj, province_adj, source_index, training, joint=True,
                 want_coeffs=False):
    """Returns (IntraNC [N, H*d'], coeffs dict).  ``alpha``: pre-dropout attention over the attention CSR with
    the dropout seed of the producing attention block (read from its grad_fn)."""
    H = len(layers)
    d = layers[0].out_features
    p = float(layers[0].dropout) if training else 0.0
    N = graph.n_rows
    src = source_index.to(torch.int64).contiguous()
    if not src.is_cuda:
        raise RuntimeError("msha_b200 is CUDA-only: source_index must be a CUDA tensor")
    B = src.numel()
    a3 = torch.cat([(l.a3[:d, 0] + l.a3[d:, 0]).reshape(1, d) for l in layers], dim=0)     # Ours.py:74-75
    a4 = torch.cat([(l.a4[:d, 0] + l.a4[d:, 0]).reshape(1, d) for l in layers], dim=0)     # Ours.py:77-78
    h2b = h2.index_select(0, src)                                                           # h2_ Ours.py:62
    t3, t4 = Fn.node_scores(h2b, a3, a4, H, d)
    t3 = torch.nn.functional.leaky_relu(t3, 0.2)
    t4 = torch.nn.functional.leaky_relu(t4, 0.2)
    rp3, col3, map3, cnt3, full3 = _as_lists(city_adj, N)
    rp4, col4, map4, cnt4, full4 = _as_lists(province_adj, N)
    n3 = cnt3(src).view(B, 1)
    n4 = cnt4(src).view(B, 1)
    if joint:
        # (p, seed) of the dropout the producing attention block drew on alpha (Ours.py:69): attached to the tensor by
        # functional.attention_block; the grad_fn route is kept for tensors that went through other wrappers
        fn = alpha.grad_fn
        a_p, a_seed = getattr(alpha, "_msha_drop", None) or ((fn.p, fn.seed) if fn is not None and hasattr(fn, "seed") else (0.0, 0))
        if training and layers[0].dropout > 0 and a_p == 0.0:
            raise RuntimeError("intra_scales: the dropout stream of alpha is unknown (pass the tensor returned by attention_block)")
        T = _RowsumExp.apply(alpha, graph, src, a_p, a_seed)                                # Ours.py:86
        total = n3 * torch.exp(t3) + n4 * torch.exp(t4) + T                                 # Ours.py:84-86
        c3 = torch.exp(t3) / total                                                          # Ours.py:87
        c4 = torch.exp(t4) / total                                                          # Ours.py:89
    else:
        # dense / Graph adjacencies may hold empty rows: one host read to refuse them (group-id vectors cannot, and skip
        # the read -- which also keeps the step capturable into a CUDA graph)
        if not (full3 and full4) and (bool((n3 == 0).any()) or bool((n4 == 0).any())):
            raise RuntimeError("OursLayer2: a batch row without city/province neighbours is not supported")
        # softmax of a row-constant logit: uniform over the list, zero gradient w.r.t. a3/a4 (Ablation.py:194-197)
        c3 = (1.0 / n3).expand(B, H).contiguous()
        c4 = (1.0 / n4).expand(B, H).contiguous()
    seed3 = ops.next_seed() if p > 0 else 0
    seed4 = ops.next_seed() if p > 0 else 0
    out = (_GroupScatter.apply(c3, h2b, (rp3, col3, map3), src, N, H, d, p, seed3, 5)
           + _GroupScatter.apply(c4, h2b, (rp4, col4, map4), src, N, H, d, p, seed4, 6))    # Ours.py:99
    coeffs = {}
    if want_coeffs:
        # the reference's heads overwrite one another's export (Ours.py:92-96): the LAST head's coefficients survive
        coeffs["Coeff3"] = _dense_rows(c3.detach()[:, -1], rp3, col3, map3, src, N)
        coeffs["Coeff4"] = _dense_rows(c4.detach()[:, -1], rp4, col4, map4, src, N)
    return out, coeffs


def _dense_rows(coef, rowptr, col, row_map, src, N):
    """(B, N) dense attention rows for the explainer export (Ours.py:92-96)."""
    B = src.numel()
    out = torch.zeros((B, N), dtype=torch.float32, device=coef.device)
    rows = row_map.long()[src] if row_map is not None else src
    beg = rowptr.long()[rows]
    end = rowptr.long()[rows + 1]
    for b in range(B):
        out[b, col[int(beg[b]):int(end[b])].long()] = coef[b]
    return out
