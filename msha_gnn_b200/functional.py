"""torch.autograd.Function wrappers around the C-ABI kernels (forward + hand-written backward).

Each Function cites the reference arithmetic it replaces.  Tensors are fp32, contiguous, CUDA;
outputs and workspaces are allocated through PyTorch's caching allocator and handed to the
library as raw pointers on the current stream.
"""
from __future__ import annotations

import os
import weakref

import torch

from . import ops
from .graph import Graph
from .ops import (ACT_ELU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SIGMOID_RELU, LRELU_SLOPE, call, ptr,
                  workspace, _stream)

I32 = torch.int32


def _hub_scratch(hub, H, D, device):
    """Partial-result buffer of the hub-row segments (None when the graph has no hub rows)."""
    if not hub.n_segs:
        return None
    n = ops._lib.lib().msha_gat_fwd_hub_scratch_floats(hub.n_segs, H, D)
    return torch.empty(n, dtype=torch.float32, device=device)


def _c(t):
    """Contiguous and 16-byte aligned (views sliced out of a larger buffer may start at any 4-byte offset)."""
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


# ------------------------------------------------------------------------------------------------
# Linear: y = x @ W   (torch.mm(input, self.W) GAT.py:21, Ours.py:57-58)
# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, out=None, precomputed=False):
        x, W = _c(x), _c(W)
        ctx.save_for_backward(x, W)
        if out is None:
            return ops.gemm(x, W)
        if not precomputed:                  # precomputed: `out` already holds x @ W (issued early, row chunk by row chunk,
            ops.gemm(x, W, out=out)          # by the producer of x -- dist_p2p.py); only the graph node is created here
        ctx.mark_dirty(out)                  # caller-owned destination (a rank's own block of a peer-mapped buffer)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = _c(dy)
        dx = ops.gemm(dy, W, transB=True) if ctx.needs_input_grad[0] else None
        dW = ops.gemm(x, dy, transA=True) if ctx.needs_input_grad[1] else None
        return dx, dW, None, None


def linear(x, W, out=None, precomputed=False):
    return _Linear.apply(x, W, out, precomputed)


class _LinearBiasAct(torch.autograd.Function):
    """act(x @ W.T + b) with nn.Linear weight layout [out, in]  (lin(x); F.relu; sigmoid LLP.py:107-115)."""

    @staticmethod
    def forward(ctx, x, W, b, act):
        x, W = _c(x), _c(W)
        y = ops.gemm(x, W, transB=True, bias=b, act=act)
        ctx.act = act
        ctx.save_for_backward(x, W, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dy = _c(dy)
        db = None
        if ctx.needs_input_grad[2]:
            # g = dy * act'(y) and db = colsum(g) in one streaming pass
            n, C = y.shape
            g = torch.empty_like(y)
            db = torch.empty(C, dtype=torch.float32, device=y.device)
            lib = ops._lib.lib()
            ws = workspace(lib.msha_act_bwd_colsum_workspace_bytes(C), y.device)
            call("msha_act_bwd_colsum", ptr(dy), ptr(y), ptr(g), n, C, ctx.act, LRELU_SLOPE, ptr(db), ws.data_ptr(),
                 ws.numel(), _stream())
        else:
            g = ops.act_bwd(dy, y, ctx.act) if ctx.act != ACT_NONE else dy
        dx = ops.gemm(g, W) if ctx.needs_input_grad[0] else None
        dW = ops.gemm(g, x, transA=True) if ctx.needs_input_grad[1] else None
        return dx, dW, db, None


def linear_bias_act(x, W, b, act=ACT_NONE):
    return _LinearBiasAct.apply(x, W, b, act)


def colsum(x, y=None, s=None, D=None):
    """sum over rows of x (* y) (* s broadcast over D-wide heads) -> [C]."""
    n, C = x.shape
    out = torch.empty(C, dtype=torch.float32, device=x.device)
    lib = ops._lib.lib()
    ws = workspace(lib.msha_colreduce_workspace_bytes(C), x.device)
    call("msha_colreduce", ptr(x), ptr(y), ptr(s), n, C, D if D else C, ptr(out), ws.data_ptr(), ws.numel(), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# activations
# ------------------------------------------------------------------------------------------------
class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        y = ops.act_fwd(_c(x), act)
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return ops.act_bwd(_c(dy), y, ctx.act), None


def elu(x):
    return _Act.apply(x, ACT_ELU)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout_apply(_c(x), p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout_apply(_c(dy), ctx.p, ctx.seed), None, None


def dropout(x, p, training):
    """F.dropout(x, p, training) with the Philox stream of this library (Ours.py:161-162,165)."""
    if not training or p == 0.0:
        return x
    return _Dropout.apply(x, float(p), ops.next_seed())


# ------------------------------------------------------------------------------------------------
# node scores  s[n,h] = feat[n,h,:] . a[h,:]     (torch.matmul(inter_input, self.a), Ours.py:64)
# ------------------------------------------------------------------------------------------------
class _NodeScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, a1, a2, H, D, pre=None):
        feat, a1 = _c(feat), _c(a1)
        n = feat.shape[0]
        if a2 is not None:
            a2 = _c(a2)
        if pre is not None:                  # (s1, s2) already computed from exactly these operands (dist_p2p.py chunk hook)
            s1, s2 = pre
        else:
            s1 = torch.empty((n, H), dtype=torch.float32, device=feat.device)
            s2 = torch.empty((n, H), dtype=torch.float32, device=feat.device) if a2 is not None else None
            call("msha_node_scores", ptr(feat), n, H, D, ptr(a1), ptr(s1), ptr(a2), ptr(s2), _stream())
        ctx.H, ctx.D = H, D
        ctx.has2 = a2 is not None
        ctx.save_for_backward(feat, a1, a2 if a2 is not None else a1)
        if s2 is None:
            return s1
        return s1, s2

    @staticmethod
    def backward(ctx, ds1, ds2=None):
        feat, a1, a2 = ctx.saved_tensors
        H, D = ctx.H, ctx.D
        n = feat.shape[0]
        ds1 = _c(ds1)
        if ctx.has2:
            ds2 = _c(ds2)
        dfeat = da1 = da2 = None
        if ctx.needs_input_grad[0]:
            dfeat = torch.empty_like(feat)
            call("msha_node_outer_add", ptr(dfeat), n, H, D, ptr(ds1), ptr(a1), ptr(ds2) if ctx.has2 else None,
                 ptr(a2) if ctx.has2 else None, 0, _stream())
        if ctx.needs_input_grad[1]:
            da1 = colsum(feat, None, ds1, D).view(H, D)
        if ctx.has2 and ctx.needs_input_grad[2]:
            da2 = colsum(feat, None, ds2, D).view(H, D)
        return dfeat, da1, da2, None, None, None


def node_scores(feat, a1, a2=None, H=1, D=None, pre=None):
    """feat [n, H*D]; a1/a2 [H, D].  Returns s1 (and s2) of shape [n, H]."""
    D = D if D is not None else feat.shape[1] // H
    return _NodeScores.apply(feat, a1, a2, H, D, pre)


# ------------------------------------------------------------------------------------------------
# attention block: logits + masked row softmax + dropout + aggregation(s)
#   e_ij = lrelu(s_nbr[j] + s_self[i])                              Ours.py:64-65 / Ablation.py:265-266
#   alpha = softmax_j(where(adj>0, e, -9e15)); dropout              Ours.py:66-69
#   out_rows[i] = sum_j alpha_ij feat_nbr[j]        (alpha @ h1)     Ours.py:98
#   out_cols[j] = sum_i alpha_ij feat_self[i]       (alpha.T @ h2)   Ours.py:100   (optional)
# ------------------------------------------------------------------------------------------------
class _AttentionBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_nbr, s_self, feat_nbr, feat_self, graph: Graph, H, D, act, p, seed, want_cols, want_lse,
                grad_sink=None):
        s_nbr, s_self, feat_nbr = _c(s_nbr), _c(s_self), _c(feat_nbr)
        rp, col = graph.attention_csr()
        N, M = graph.n_rows, graph.n_cols
        E = col.numel()
        C = H * D
        assert s_nbr.shape == (M, H) and s_self.shape == (N, H) and feat_nbr.shape == (M, C), \
            (s_nbr.shape, s_self.shape, feat_nbr.shape, (N, M, H, D))
        dev = feat_nbr.device
        alpha = torch.empty((E, H), dtype=torch.float32, device=dev)
        out = torch.empty((N, C), dtype=torch.float32, device=dev)
        lse = torch.empty((N, H), dtype=torch.float32, device=dev) if want_lse else None
        hub = graph.hub_rows()
        scr = _hub_scratch(hub, H, D, dev)
        call("msha_gat_fwd", ptr(rp, I32), ptr(col, I32), N, ptr(s_nbr), ptr(s_self), ptr(feat_nbr), H, D, LRELU_SLOPE,
             None, ptr(alpha), ptr(out), act, ptr(lse), p, seed, hub.ptr, ptr(scr), _stream())
        out_cols = None
        if want_cols:
            feat_self = _c(feat_self)
            assert feat_self.shape == (N, C)
            colptr, rowidx, perm = graph.attention_csc()
            out_cols = torch.empty((M, C), dtype=torch.float32, device=dev)
            call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, ptr(alpha), ptr(feat_self), H, D,
                 ptr(out_cols), 0, None, None, p, seed, graph.hub_cols().ptr, _stream())
        ctx.graph, ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed = graph, H, D, act, p, seed
        ctx.want_cols, ctx.want_lse, ctx.grad_sink = want_cols, want_lse, grad_sink
        ctx.save_for_backward(s_nbr, s_self, feat_nbr, feat_self if want_cols else None, alpha,
                              out if act != ACT_NONE else None)
        ctx.set_materialize_grads(False)
        res = [out]
        if want_cols:
            res.append(out_cols)
        res.append(alpha)
        if want_lse:
            res.append(lse)
        return tuple(res)

    @staticmethod
    def backward(ctx, d_rows, *rest):
        s_nbr, s_self, feat_nbr, feat_self, alpha, out = ctx.saved_tensors
        graph, H, D, act, p, seed = ctx.graph, ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed
        rest = list(rest)
        d_cols = rest.pop(0) if ctx.want_cols else None
        d_alpha = rest.pop(0)                                # grad w.r.t. the pre-dropout alpha (may be None)
        d_lse = rest.pop(0) if ctx.want_lse else None
        rp, col = graph.attention_csr()
        colptr, rowidx, perm = graph.attention_csc()
        N, M = graph.n_rows, graph.n_cols
        C = H * D
        dev = alpha.device
        d_rows = torch.zeros((N, C), dtype=torch.float32, device=dev) if d_rows is None else _c(d_rows)
        if d_cols is not None:
            d_cols = _c(d_cols)
        dlogit = torch.empty_like(alpha)
        ds_self = torch.empty((N, H), dtype=torch.float32, device=dev)
        dz = torch.empty((N, C), dtype=torch.float32, device=dev) if act != ACT_NONE else None
        hub = graph.hub_rows()
        r_buf = torch.empty((N, H), dtype=torch.float32, device=dev) if hub.n_segs else None
        extra = _c(d_alpha) if d_alpha is not None else None
        dlse = _c(d_lse) if d_lse is not None else None
        # a sink may own the destinations of the two column-side gradients (peer-mapped buffers, dist.py)
        dfeat_nbr = getattr(ctx.grad_sink, "feat_grad", None)
        ds_nbr = getattr(ctx.grad_sink, "score_grad", None)
        if dfeat_nbr is None:
            dfeat_nbr = torch.empty((M, C), dtype=torch.float32, device=dev)
        if ds_nbr is None:
            ds_nbr = torch.empty((M, H), dtype=torch.float32, device=dev)
        assert dfeat_nbr.shape == (M, C) and ds_nbr.shape == (M, H)
        sink = ctx.grad_sink if (d_cols is None and hasattr(ctx.grad_sink, "start")) else None
        if sink is not None:
            # Partitioned graphs (dist.gat_encode): d feat_nbr is the big message of the backward.  Produce it first
            # -- it needs only alpha and dZ -- and hand it to the sink, which starts its reduce-scatter; the row pass
            # and the d s_nbr column sums then run while the collective is in flight.
            if dz is not None:
                call("msha_act_bwd", ptr(d_rows), ptr(out), ptr(dz), N * C, act, LRELU_SLOPE, _stream())
            dzz = dz if dz is not None else d_rows
            call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, ptr(alpha), ptr(dzz), H, D,
                 ptr(dfeat_nbr), 0, None, None, p, seed, graph.hub_cols().ptr, _stream())
            sink.start(dfeat_nbr)
            call("msha_gat_bwd_rows", ptr(rp, I32), ptr(col, I32), N, ptr(s_nbr), ptr(s_self), LRELU_SLOPE, ptr(alpha),
                 ptr(feat_nbr), ptr(dzz), None, ACT_NONE, None, None, None, ptr(extra), ptr(dlse), H, D, ptr(dlogit),
                 ptr(ds_self), p, seed, hub.ptr, ptr(r_buf), int(alpha.shape[0] // max(N, 1)), _stream())
            call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, None, None, H, D,
                 None, 0, ptr(dlogit), ptr(ds_nbr), p, seed, graph.hub_cols().ptr, _stream())
            return ds_nbr, ds_self, dfeat_nbr, None, None, None, None, None, None, None, None, None, None
        call("msha_gat_bwd_rows", ptr(rp, I32), ptr(col, I32), N, ptr(s_nbr), ptr(s_self), LRELU_SLOPE, ptr(alpha),
             ptr(feat_nbr), ptr(d_rows), ptr(out), act, ptr(dz), ptr(d_cols), ptr(feat_self) if d_cols is not None else None,
             ptr(extra), ptr(dlse), H, D, ptr(dlogit), ptr(ds_self), p, seed, hub.ptr, ptr(r_buf),
             int(alpha.shape[0] // max(N, 1)), _stream())
        dzz = dz if dz is not None else d_rows
        call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, ptr(alpha), ptr(dzz), H, D,
             ptr(dfeat_nbr), 0, ptr(dlogit), ptr(ds_nbr), p, seed, graph.hub_cols().ptr, _stream())
        dfeat_self = None
        if ctx.want_cols and d_cols is not None and ctx.needs_input_grad[3]:
            dfeat_self = torch.empty((N, C), dtype=torch.float32, device=dev)
            scr = _hub_scratch(hub, H, D, dev)
            call("msha_gat_fwd", ptr(rp, I32), ptr(col, I32), N, None, None, ptr(d_cols), H, D, LRELU_SLOPE, ptr(alpha),
                 None, ptr(dfeat_self), ACT_NONE, None, p, seed, hub.ptr, ptr(scr), _stream())
        return ds_nbr, ds_self, dfeat_nbr, dfeat_self, None, None, None, None, None, None, None, None, None


def attention_block(graph: Graph, s_nbr, s_self, feat_nbr, feat_self=None, heads=1, act=ACT_NONE, dropout_p=0.0,
                    training=True, want_cols=False, want_lse=False, grad_sink=None):
    """Returns (out_rows, [out_cols,] alpha [, lse]); alpha is [E_att, H], the pre-dropout attention over the
    attention CSR (differentiable: gradients flowing into it join the softmax backward); lse is the row-wise
    log-sum-exp of the logits ([N, H], -inf for rows without edges).  ``grad_sink`` (optional, ``.start(tensor)``) is
    handed d feat_nbr as soon as it exists in the backward (dist.py starts its reduce-scatter there)."""
    C = feat_nbr.shape[1]
    D = C // heads
    p = float(dropout_p) if training else 0.0
    seed = ops.next_seed() if p > 0 else 0
    res = _AttentionBlock.apply(s_nbr, s_self, feat_nbr, feat_self, graph, heads, D, act, p, seed, want_cols, want_lse,
                                grad_sink)
    # the dropout stream of this block travels with alpha: consumers that must re-draw the same mask (the SUM_county term
    # of the intra scales, Ours.py:86) read it from here -- also under no_grad, where alpha has no grad_fn
    res[2 if want_cols else 1]._msha_drop = (p, seed)
    return res


# ------------------------------------------------------------------------------------------------
# weighted SpMM with given (non-learned) edge weights: out[i] = sum_j w_ij feat[j]   (GraphConvolution
# model.py:37 uses the transposed form; see spmm_t)
# ------------------------------------------------------------------------------------------------
class _SpmmT(torch.autograd.Function):
    """out[j] = sum_i w[e(i,j)] * feat[i]  over the canonical CSR (adj.T @ support, model.py:37)."""

    @staticmethod
    def forward(ctx, feat, w, graph: Graph):
        feat, w = _c(feat), _c(w)
        colptr, rowidx, perm = graph.transpose_structure()
        M, C = graph.n_cols, feat.shape[1]
        out = torch.empty((M, C), dtype=torch.float32, device=feat.device)
        call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, ptr(w), ptr(feat), 1, C, ptr(out), 0,
             None, None, 0.0, 0, graph.hub_cols_plain().ptr, _stream())
        ctx.graph = graph
        ctx.save_for_backward(w)
        return out

    @staticmethod
    def backward(ctx, dout):
        (w,) = ctx.saved_tensors
        g = ctx.graph
        dout = _c(dout)
        C = dout.shape[1]
        dfeat = torch.empty((g.n_rows, C), dtype=torch.float32, device=dout.device)
        hub = g.hub_rows_plain()
        scr = _hub_scratch(hub, 1, C, dout.device)
        call("msha_gat_fwd", ptr(g.rowptr, I32), ptr(g.col, I32), g.n_rows, None, None, ptr(dout), 1, C, LRELU_SLOPE,
             ptr(w), None, ptr(dfeat), ACT_NONE, None, 0.0, 0, hub.ptr, ptr(scr), _stream())
        return dfeat, None, None


def spmm_t(graph: Graph, w, feat):
    return _SpmmT.apply(feat, w.view(-1, 1), graph)


# ------------------------------------------------------------------------------------------------
# a-1 GraphAttentionLayer epilogue: elu(softmax(mask(lrelu(rowconst))) * h)          GAT.py:24-35
# ------------------------------------------------------------------------------------------------
class _Gal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, a_dummy, graph: Graph, H, p, seed):
        h = _c(h)
        rp, col = graph.attention_csr()
        N, M = graph.n_rows, graph.n_cols
        assert h.shape == (N, H * M)
        out = torch.empty_like(h)
        call("msha_gal_fwd", ptr(h), ptr(rp, I32), ptr(col, I32), N, H, M, ptr(out), p, seed, _stream())
        ctx.graph, ctx.H, ctx.p, ctx.seed = graph, H, p, seed
        ctx.a_shape = None if a_dummy is None else a_dummy.shape
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        g = ctx.graph
        rp, col = g.attention_csr()
        dh = torch.empty_like(out)
        call("msha_gal_bwd", ptr(_c(dout)), ptr(out), ptr(rp, I32), ptr(col, I32), g.n_rows, ctx.H, g.n_cols, ptr(dh),
             ctx.p, ctx.seed, _stream())
        # d/da == 0: the logit is constant along the softmax axis
        da = None if ctx.a_shape is None else torch.zeros(ctx.a_shape, dtype=torch.float32, device=out.device)
        return dh, da, None, None, None, None


def gal(h, a, graph: Graph, heads=1, dropout_p=0.0, training=True):
    p = float(dropout_p) if training else 0.0
    seed = ops.next_seed() if p > 0 else 0
    return _Gal.apply(h, a, graph, heads, p, seed)


# ------------------------------------------------------------------------------------------------
# BatchNorm1d (node axis) + LeakyReLU                      self.leakyrelu(self.bn1(.)) Ours.py:100-101
# ------------------------------------------------------------------------------------------------
class _BnLrelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps):
        x = _c(x)
        n, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_bn_workspace_bytes(C), x.device)
        call("msha_bn_lrelu_fwd", ptr(x), n, C, ptr(_c(gamma)), ptr(_c(beta)), ptr(running_mean), ptr(running_var),
             int(training), float(momentum), float(eps), LRELU_SLOPE, ptr(y), ptr(mean), ptr(invstd), ws.data_ptr(),
             ws.numel(), _stream())
        ctx.training = training
        ctx.save_for_backward(x, y, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, invstd = ctx.saved_tensors
        n, C = x.shape
        dx = torch.empty_like(x)
        xhat = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_bn_workspace_bytes(C), x.device)
        call("msha_bn_lrelu_bwd", ptr(_c(dy)), ptr(y), ptr(x), n, C, ptr(_c(gamma)), ptr(mean), ptr(invstd),
             int(ctx.training), LRELU_SLOPE, ptr(dx), ptr(xhat), ptr(dgamma), ptr(dbeta), ws.data_ptr(), ws.numel(),
             _stream())
        return dx, dgamma, dbeta, None, None, None, None, None


def bn_lrelu(x, gamma, beta, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    if training and x.shape[0] < 2:
        raise ValueError("Expected more than 1 value per channel when training")   # nn.BatchNorm1d behaviour
    return _BnLrelu.apply(x, gamma, beta, running_mean, running_var, training, momentum, eps)


class _BnLreluDist(torch.autograd.Function):
    """``leakyrelu(bn(x))`` over a node axis that is partitioned across ranks (Ours.py:100-101 at multi-GPU scale, SURVEY.md
    section 8e): the 2*C column sums leave the library between the statistics and the apply pass, ``reduce`` (a callable
    summing a small tensor over the ranks and returning the result) is applied to them, ``n_total`` is the global row
    count.  Same kernels as ``_BnLrelu``; d gamma / d beta come out global, i.e. identical on every rank -- the caller
    must NOT sum them over ranks again (they are returned divided by ``grad_replicas`` so that a gradient all-reduce that
    sums every parameter's gradient over the ranks restores them)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, momentum, eps, n_total, reduce, grad_replicas):
        x = _c(x)
        n, C = x.shape
        dev = x.device
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=dev)
        invstd = torch.empty(C, dtype=torch.float32, device=dev)
        lib = ops._lib.lib()
        sums = None
        if training:
            ws = workspace(lib.msha_bn_workspace_bytes(C), dev)
            sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
            call("msha_bn_stats", ptr(x), n, C, sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
            sums = reduce(sums)
        call("msha_bn_lrelu_apply", ptr(x), n, C, sums.data_ptr() if sums is not None else None, int(n_total),
             ptr(_c(gamma)), ptr(_c(beta)), ptr(running_mean), ptr(running_var), int(training), float(momentum), float(eps),
             LRELU_SLOPE, ptr(y), ptr(mean), ptr(invstd), _stream())
        ctx.training, ctx.n_total, ctx.reduce, ctx.grad_replicas = training, int(n_total), reduce, grad_replicas
        ctx.save_for_backward(x, y, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, invstd = ctx.saved_tensors
        n, C = x.shape
        dev = x.device
        dx = torch.empty_like(x)
        xhat = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=dev)
        dbeta = torch.empty(C, dtype=torch.float32, device=dev)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_bn_workspace_bytes(C), dev)
        sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
        call("msha_bn_lrelu_bwd_stats", ptr(_c(dy)), ptr(y), ptr(x), n, C, ptr(mean), ptr(invstd), LRELU_SLOPE, ptr(dx),
             ptr(xhat), sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        sums = ctx.reduce(sums)
        call("msha_bn_lrelu_bwd_apply", ptr(dx), ptr(xhat), n, C, ptr(_c(gamma)), ptr(invstd), sums.data_ptr(), ctx.n_total,
             int(ctx.training), ptr(dgamma), ptr(dbeta), _stream())
        if ctx.grad_replicas != 1:
            dgamma, dbeta = dgamma / ctx.grad_replicas, dbeta / ctx.grad_replicas
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None


def bn_lrelu_dist(x, gamma, beta, running_mean, running_var, training, n_total, reduce, momentum=0.1, eps=1e-5,
                  grad_replicas=1):
    if training and n_total < 2:
        raise ValueError("Expected more than 1 value per channel when training")   # nn.BatchNorm1d behaviour
    return _BnLreluDist.apply(x, gamma, beta, running_mean, running_var, training, momentum, eps, n_total, reduce,
                              grad_replicas)


class _GroupRowsSum(torch.autograd.Function):
    """G[g] = sum_{i : gid[i] = g} x[i]  -- the per-group sums of the intra-scale block in its group-sum form
    (Ours.py:99: att3.t() @ h2_ is one row per group, see csrc/intra_kernels.cu)."""

    @staticmethod
    def forward(ctx, x, gid, n_groups):
        x = _c(x)
        G = torch.empty((n_groups, x.shape[1]), dtype=torch.float32, device=x.device)
        call("msha_group_rows_sum", ptr(x), ptr(gid, torch.int64), x.shape[0], x.shape[1], n_groups, ptr(G), _stream())
        ctx.save_for_backward(gid)
        return G

    @staticmethod
    def backward(ctx, dG):
        (gid,) = ctx.saved_tensors
        dG = _c(dG)
        dx = torch.empty((gid.numel(), dG.shape[1]), dtype=torch.float32, device=dG.device)
        call("msha_group_rows_add", ptr(dx), ptr(dG), ptr(gid, torch.int64), None, None, gid.numel(), dG.shape[1], 0, _stream())
        return dx, None, None


class _GroupRowsAdd(torch.autograd.Function):
    """out[i] = G3[gid3[i]] + G4[gid4[i]]  (IntraNC of every node from the two group tables)."""

    @staticmethod
    def forward(ctx, G3, gid3, G4, gid4):
        G3, G4 = _c(G3), _c(G4)
        n, C = gid3.numel(), G3.shape[1]
        out = torch.empty((n, C), dtype=torch.float32, device=G3.device)
        call("msha_group_rows_add", ptr(out), ptr(G3), ptr(gid3, torch.int64), ptr(G4), ptr(gid4, torch.int64), n, C, 0, _stream())
        ctx.shapes = (G3.shape[0], G4.shape[0])
        ctx.save_for_backward(gid3, gid4)
        return out

    @staticmethod
    def backward(ctx, dout):
        gid3, gid4 = ctx.saved_tensors
        dout = _c(dout)
        n, C = dout.shape
        d3 = torch.empty((ctx.shapes[0], C), dtype=torch.float32, device=dout.device)
        d4 = torch.empty((ctx.shapes[1], C), dtype=torch.float32, device=dout.device)
        call("msha_group_rows_sum", ptr(dout), ptr(gid3, torch.int64), n, C, ctx.shapes[0], ptr(d3), _stream())
        call("msha_group_rows_sum", ptr(dout), ptr(gid4, torch.int64), n, C, ctx.shapes[1], ptr(d4), _stream())
        return d3, None, d4, None


def group_rows_sum(x, gid, n_groups):
    return _GroupRowsSum.apply(x, gid.contiguous(), int(n_groups))


def group_rows_add(G3, gid3, G4, gid4):
    return _GroupRowsAdd.apply(G3, gid3.contiguous(), G4, gid4.contiguous())


# ------------------------------------------------------------------------------------------------
# score matrix: act(u @ v.T) per head                               F.elu(u @ v.t()) Ours.py:108-109
# ------------------------------------------------------------------------------------------------
class _MatmulNtAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, v, H, act):
        """u [N, H*d], v [M, H*d] -> out [N, H*M] with out[:, h*M:(h+1)*M] = act(u_h @ v_h.T)."""
        u, v = _c(u), _c(v)
        N, M = u.shape[0], v.shape[0]
        d = u.shape[1] // H
        out = torch.empty((N, H * M), dtype=torch.float32, device=u.device)
        for h in range(H):
            ops.gemm(u[:, h * d:(h + 1) * d], v[:, h * d:(h + 1) * d], transB=True, act=act, out=out[:, h * M:(h + 1) * M])
        ctx.H, ctx.act = H, act
        ctx.save_for_backward(u, v, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        u, v, out = ctx.saved_tensors
        H = ctx.H
        N, M = u.shape[0], v.shape[0]
        d = u.shape[1] // H
        g = ops.act_bwd(_c(dout), out, ctx.act) if ctx.act != ACT_NONE else _c(dout)
        du = torch.empty_like(u)
        dv = torch.empty_like(v)
        for h in range(H):
            gh = g[:, h * M:(h + 1) * M]
            ops.gemm(gh, v[:, h * d:(h + 1) * d], out=du[:, h * d:(h + 1) * d])
            ops.gemm(gh, u[:, h * d:(h + 1) * d], transA=True, out=dv[:, h * d:(h + 1) * d])
        return du, dv, None, None


def matmul_nt_act(u, v, heads=1, act=ACT_ELU):
    return _MatmulNtAct.apply(u, v, heads, act)


# ------------------------------------------------------------------------------------------------
# read-out: log_softmax(elu(x), dim=1)                                           Ours.py:166-167
# ------------------------------------------------------------------------------------------------
class _LogSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pre_elu):
        x = _c(x)
        y = torch.empty_like(x)
        call("msha_log_softmax_fwd", ptr(x), x.shape[0], x.shape[1], int(pre_elu), ptr(y), _stream())
        ctx.pre_elu = pre_elu
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dx = torch.empty_like(x)
        call("msha_log_softmax_bwd", ptr(_c(dy)), ptr(y), ptr(x), x.shape[0], x.shape[1], int(ctx.pre_elu), ptr(dx), _stream())
        return dx, None


def log_softmax(x, pre_elu=False):
    return _LogSoftmax.apply(x, pre_elu)


CHECK_TARGETS = os.environ.get("MSHA_CHECK_TARGETS", "0") != "0"


def _check_targets(status, n_classes):
    """The read-out kernels never synchronise: an out-of-range target contributes 0 and sets a status word, where
    F.nll_loss raises.  MSHA_CHECK_TARGETS=1 (debugging) reads the word back -- one host synchronisation per call."""
    if CHECK_TARGETS and int(status.item()) != 0:
        raise IndexError(f"nll_loss: a target lies outside [0, {n_classes})")


class _NllLoss(torch.autograd.Function):
    """F.nll_loss(logp, target) (mean): gather + deterministic sum; backward streams the dense gradient."""

    @staticmethod
    def forward(ctx, logp, target):
        logp = _c(logp)
        P, C = logp.shape
        loss = torch.empty((), dtype=torch.float32, device=logp.device)
        status = torch.empty(1, dtype=torch.int32, device=logp.device)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_nll_workspace_bytes(), logp.device)
        call("msha_nll_loss_fwd", ptr(logp), ptr(target, torch.int64), P, C, loss.data_ptr(), ptr(status, I32),
             ws.data_ptr(), ws.numel(), _stream())
        _check_targets(status, C)
        ctx.shape = (P, C)
        ctx.save_for_backward(target)
        return loss

    @staticmethod
    def backward(ctx, g):
        (target,) = ctx.saved_tensors
        P, C = ctx.shape
        d = torch.empty((P, C), dtype=torch.float32, device=target.device)
        call("msha_nll_loss_bwd", ptr(target, torch.int64), _c(g).data_ptr(), P, C, ptr(d), _stream())
        return d, None


def nll_loss(logp, target):
    """Mean negative log-likelihood read-out (train.py:229).  Targets must lie in [0, C)."""
    if target.dtype != torch.int64:
        target = target.long()
    return _NllLoss.apply(logp, target.contiguous())


# ------------------------------------------------------------------------------------------------
# link scorer pieces                                                        LLP.py:104-115
# ------------------------------------------------------------------------------------------------
class _PairMul(torch.autograd.Function):
    """z[p] = h_i[src[p]] * h_j[dst[p]]  (x_i * x_j after the caller's gather, LLP.py:105,233)."""

    @staticmethod
    def forward(ctx, hi, hj, src, dst):
        hi, hj = _c(hi), _c(hj)
        P = src.numel() if src is not None else hi.shape[0]
        C = hi.shape[1]
        z = torch.empty((P, C), dtype=torch.float32, device=hi.device)
        call("msha_pair_gather_mul", ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, C, ptr(z), _stream())
        ctx.save_for_backward(hi, hj, src, dst)
        return z

    @staticmethod
    def backward(ctx, dz):
        hi, hj, src, dst = ctx.saved_tensors
        P, C = dz.shape
        dhi = torch.zeros_like(hi)
        dhj = torch.zeros_like(hj)
        call("msha_pair_scatter_mul_add", ptr(_c(dz)), ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, C,
             ptr(dhi), ptr(dhj), _stream())
        return dhi, dhj, None, None


def pair_mul(hi, hj, src=None, dst=None):
    return _PairMul.apply(hi, hj, src, dst)



def _pair_grad_tables(ctx, hi, hj):
    """Destinations of d h_i / d h_j for the pair scatters.  One table scored against itself (``hi is hj``: link scoring
    over a single embedding matrix) gets ONE buffer -- both scatters are atomic adds -- instead of two tables that
    autograd would then have to add (three passes over the table); its gradient is returned for the first input only.
    ``ctx.grad_buffer`` (set in the forward from the table's ``_msha_grad_buffer`` attribute) lets the caller own that
    buffer: dist.py hands out the peer-mapped one its reduce-scatter reads."""
    same = hi.data_ptr() == hj.data_ptr() and hi.shape == hj.shape
    if same:
        buf = ctx.grad_buffer() if ctx.grad_buffer is not None else None
        dhi = buf.zero_() if buf is not None else torch.zeros_like(hi)
        return dhi, dhi, True
    return torch.zeros_like(hi), torch.zeros_like(hj), False

class _ScoreMLP(torch.autograd.Function):
    """Fused LinkPredictor with one hidden Linear: act((h_i[src] * h_j[dst]) @ W0.T + b0)  (LLP.py:105-115).
    Forward: one tensor-core kernel (gather, Hadamard, 3xTF32 split in the producer warps; no Z tensor in HBM)."""

    @staticmethod
    def forward(ctx, hi, hj, src, dst, W0, b0, act):
        hi, hj, W0 = _c(hi), _c(hj), _c(W0)
        P = src.numel() if src is not None else hi.shape[0]
        Hd, C = W0.shape
        out = torch.empty((P, Hd), dtype=torch.float32, device=hi.device)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_score_mlp_workspace_bytes(C, Hd), hi.device)
        call("msha_score_mlp_fwd", ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, C, ptr(W0),
             ptr(b0) if b0 is not None else None, Hd, act, LRELU_SLOPE, ptr(out), Hd, ws.data_ptr(), ws.numel(), _stream())
        ctx.act = act
        ctx.has_bias = b0 is not None
        ctx.grad_buffer = getattr(hi, "_msha_grad_buffer", None)
        ctx.save_for_backward(hi, hj, src, dst, W0, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        hi, hj, src, dst, W0, out = ctx.saved_tensors
        P, Hd = out.shape
        C = W0.shape[1]
        dout = _c(dout)
        g = torch.empty_like(out)
        lib = ops._lib.lib()
        if FUSED_SCORE_BWD and Hd % 4 == 0 and C <= 256:
            # two tensor-core kernels: (G, db, dZ, scatter) and (dW0 with Z regenerated from the gathers)
            dhi, dhj, same = _pair_grad_tables(ctx, hi, hj)
            dW = torch.empty_like(W0)
            db = torch.empty(Hd, dtype=torch.float32, device=out.device)
            ws = workspace(lib.msha_score_mlp_workspace_bytes(C, Hd), out.device)
            call("msha_score_mlp_bwd", ptr(dout), ptr(out), ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P,
                 C, ptr(W0), Hd, ctx.act, LRELU_SLOPE, ptr(g), ptr(dhi), ptr(dhj), ptr(dW), ptr(db), ws.data_ptr(), ws.numel(),
                 _stream())
            return dhi, (None if same else dhj), None, None, dW, (db if ctx.has_bias else None), None
        db = torch.empty(Hd, dtype=torch.float32, device=out.device)
        ws = workspace(lib.msha_act_bwd_colsum_workspace_bytes(Hd), out.device)
        call("msha_act_bwd_colsum", ptr(dout), ptr(out), ptr(g), P, Hd, ctx.act, LRELU_SLOPE, ptr(db), ws.data_ptr(),
             ws.numel(), _stream())
        z = torch.empty((P, C), dtype=torch.float32, device=out.device)          # Z regenerated, never saved
        call("msha_pair_gather_mul", ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, C, ptr(z), _stream())
        dW = ops.gemm(g, z, transA=True) if ctx.needs_input_grad[4] else None
        dhi = dhj = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dz = ops.gemm(g, W0, out=z)                                          # reuse the Z buffer
            dhi = torch.zeros_like(hi)
            dhj = torch.zeros_like(hj)
            call("msha_pair_scatter_mul_add", ptr(dz), ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, C,
                 ptr(dhi), ptr(dhj), _stream())
        return dhi, dhj, None, None, dW, (db if ctx.has_bias else None), None


FUSED_SCORE_BWD = True     # False: unfused backward (act', gather, two GEMMs, scatter) -- kept for validation


class _ScoreMLPNll(torch.autograd.Function):
    """Fused scorer + nll read-out: ``F.nll_loss(act((h_i[src]*h_j[dst]) @ W0.T + b0), target)`` (LLP.py:233-235).
    Forward = the scorer kernel + the read-out reduction; the backward never materialises d out -- the producer warps of
    the dZ kernel generate it from ``target`` (it is -g/P in one column per row)."""

    @staticmethod
    def forward(ctx, hi, hj, src, dst, W0, b0, target, act):
        hi, hj, W0 = _c(hi), _c(hj), _c(W0)
        P = src.numel() if src is not None else hi.shape[0]
        Hd, C = W0.shape
        dev = hi.device
        out = torch.empty((P, Hd), dtype=torch.float32, device=dev)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_score_mlp_workspace_bytes(C, Hd), dev)
        call("msha_score_mlp_fwd", ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, C, ptr(W0),
             ptr(b0) if b0 is not None else None, Hd, act, LRELU_SLOPE, ptr(out), Hd, ws.data_ptr(), ws.numel(), _stream())
        loss = torch.empty((), dtype=torch.float32, device=dev)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        ws2 = workspace(lib.msha_nll_workspace_bytes(), dev)
        call("msha_nll_loss_fwd", ptr(out), ptr(target, torch.int64), P, Hd, loss.data_ptr(), ptr(status, I32),
             ws2.data_ptr(), ws2.numel(), _stream())
        _check_targets(status, Hd)
        ctx.act, ctx.has_bias = act, b0 is not None
        ctx.grad_buffer = getattr(hi, "_msha_grad_buffer", None)
        ctx.order, ctx.order_owned = None, False
        if SPARSE_NLL_BWD and C % 4 == 0 and C <= 1024 and P < (1 << 31) and any(ctx.needs_input_grad):
            if NLL_ORDER_BY_SRC and src is not None:
                ctx.order = nll_label_order(target, Hd, src, hi.shape[0])   # per call: the sources change with the batch
                ctx.order_owned = True
            else:
                ctx.order = nll_label_order(target, Hd)   # here `target` is still the caller's tensor: the cache can hit
        ctx.save_for_backward(hi, hj, src, dst, W0, out, target)
        ctx.mark_non_differentiable(out)
        ctx.set_materialize_grads(False)      # otherwise autograd zero-fills a (P, Hd) gradient for `out` on every backward
        return loss, out

    @staticmethod
    def backward(ctx, gloss, _gout_unused):
        if gloss is None:
            return (None,) * 8
        hi, hj, src, dst, W0, out, target = ctx.saved_tensors
        P, Hd = out.shape
        C = W0.shape[1]
        lib = ops._lib.lib()
        dhi, dhj, same = _pair_grad_tables(ctx, hi, hj)
        dW = torch.empty_like(W0)
        db = torch.empty(Hd, dtype=torch.float32, device=out.device)
        gl = _c(gloss.reshape(1).float())
        if ctx.order is not None:
            # d out is one-hot per row: no GEMM left, only gathers and vector atomics (csrc/score_nll_sparse.cu)
            order = ctx.order
            if ctx.order_owned:
                ctx.order = None      # per-call buffer: the node outlives its backward while the caller holds the loss
                                      # (a second backward through a retained graph takes the tensor-core path below)
            call("msha_score_mlp_nll_bwd_sparse", ptr(order, I32), ptr(target, torch.int64), ptr(gl), ptr(out), Hd, ptr(hi),
                 ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, C, ptr(W0), Hd, ctx.act, LRELU_SLOPE, ptr(dhi),
                 ptr(dhj), ptr(dW), ptr(db), _stream())
            return dhi, (None if same else dhj), None, None, dW, (db if ctx.has_bias else None), None, None
        g = torch.empty_like(out)
        ws = workspace(lib.msha_score_mlp_workspace_bytes(C, Hd), out.device)
        call("msha_score_mlp_nll_bwd", ptr(target, torch.int64), ptr(gl), ptr(out), ptr(hi), ptr(hj), ptr(src, torch.int64),
             ptr(dst, torch.int64), P, C, ptr(W0), Hd, ctx.act, LRELU_SLOPE, ptr(g), ptr(dhi), ptr(dhj), ptr(dW), ptr(db),
             ws.data_ptr(), ws.numel(), _stream())
        return dhi, (None if same else dhj), None, None, dW, (db if ctx.has_bias else None), None, None


NLL_ORDER_BY_SRC = os.environ.get("MSHA_NLL_ORDER", "src") == "src"   # sort the pairs of a label by source row as well
SPARSE_NLL_BWD = os.environ.get("MSHA_NLL_BWD", "sparse") != "dense"   # "dense": the tensor-core backward (validation / comparison)


_order_cache: dict = {}


def nll_label_order(target, n_classes, src=None, n_src=0):
    """Pair indices stably sorted by label (uint32 bits in an int32 tensor) -- the traversal order of the sparse backward.
    With ``src`` the pairs of a label are further ordered by source row.  The label-only order is cached per label tensor
    (data pointer + version + weakref, like the adjacency cache): a training loop that scores a fixed edge set passes the
    same labels every step."""
    if src is not None:
        return _nll_label_order(target, n_classes, src, n_src)
    key = (target.data_ptr(), target._version, target.numel(), int(n_classes))
    hit = _order_cache.get(key)
    if hit is not None and hit[0]() is target:
        return hit[1]
    order = _nll_label_order(target, n_classes)
    if len(_order_cache) > 8:
        _order_cache.clear()
    _order_cache[key] = (weakref.ref(target), order)
    return order


def _nll_label_order(target, n_classes, src=None, n_src=0):
    P = target.numel()
    dev = target.device
    keys = torch.empty((2, P), dtype=torch.int64, device=dev)
    order = torch.empty((2, P), dtype=I32, device=dev)
    lib = ops._lib.lib()
    ws = workspace(lib.msha_radix_sort_workspace_bytes(P), dev)
    call("msha_score_nll_label_order", ptr(target, torch.int64), ptr(src, torch.int64), P, n_classes, n_src,
         keys[0].data_ptr(), keys[1].data_ptr(), order[0].data_ptr(), order[1].data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    return order[0]


def score_mlp_nll_supported(hi, hj, W0):
    Hd, C = W0.shape
    return FUSED_SCORE_BWD and score_mlp_supported(hi, hj, W0) and Hd % 4 == 0 and C <= 256


def score_mlp_nll(hi, hj, src, dst, W0, b0, target, act=ACT_SIGMOID_RELU):
    """-> (loss, scores): mean nll read-out of the fused scorer; ``scores`` is returned for metrics (not differentiable)."""
    if target.dtype != torch.int64:
        target = target.long()
    return _ScoreMLPNll.apply(hi, hj, src, dst, W0, b0, target.contiguous(), act)


def score_mlp_supported(hi, hj, W0):
    lib = ops._lib.lib()
    return (hi.is_cuda and hi.dtype == torch.float32 and hi.is_contiguous() and hj.is_contiguous() and W0.is_contiguous()
            and hi.shape[1] == W0.shape[1] == hj.shape[1]
            and lib.msha_score_mlp_supported(hi.data_ptr(), hj.data_ptr(), W0.data_ptr(), W0.shape[1], W0.shape[0]) == 0)


def score_mlp(hi, hj, src, dst, W0, b0, act=ACT_SIGMOID_RELU):
    return _ScoreMLP.apply(hi, hj, src, dst, W0, b0, act)


class _PairDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hi, hj, src, dst, act):
        hi, hj = _c(hi), _c(hj)
        P = src.numel() if src is not None else hi.shape[0]
        out = torch.empty(P, dtype=torch.float32, device=hi.device)
        call("msha_pair_dot", ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64), P, hi.shape[1], act, ptr(out),
             _stream())
        ctx.act = act
        ctx.save_for_backward(hi, hj, src, dst, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        hi, hj, src, dst, out = ctx.saved_tensors
        dhi = torch.zeros_like(hi)
        dhj = torch.zeros_like(hj)
        call("msha_pair_dot_bwd", ptr(_c(dout)), ptr(out), ptr(hi), ptr(hj), ptr(src, torch.int64), ptr(dst, torch.int64),
             out.numel(), hi.shape[1], ctx.act, ptr(dhi), ptr(dhj), _stream())
        return dhi, dhj, None, None, None


def pair_dot(hi, hj, src=None, dst=None, act=ACT_NONE):
    return _PairDot.apply(hi, hj, src, dst, act)


def negative_sample(seed: int, n_pairs: int, n_src: int, n_dst: int, device):
    """Uniform negative pairs from the library's Philox stream (bit-exact with oracle.negative_sample)."""
    src = torch.empty(n_pairs, dtype=torch.int64, device=device)
    dst = torch.empty(n_pairs, dtype=torch.int64, device=device)
    call("msha_negative_sample", seed, n_pairs, n_src, n_dst, ptr(src, torch.int64), ptr(dst, torch.int64), _stream())
    return src, dst


# ------------------------------------------------------------------------------------------------
# row SpMM over the canonical CSR with stored values: out[i] = sum_j w[e(i,j)] feat[j]
# (GCN's second layer `adj.t().transpose(0, 1) @ support` == adj @ support, model.py:37,60)
# ------------------------------------------------------------------------------------------------
class _Spmm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, w, graph: Graph):
        feat, w = _c(feat), _c(w)
        C = feat.shape[1]
        out = torch.empty((graph.n_rows, C), dtype=torch.float32, device=feat.device)
        hub = graph.hub_rows_plain()
        scr = _hub_scratch(hub, 1, C, feat.device)
        call("msha_gat_fwd", ptr(graph.rowptr, I32), ptr(graph.col, I32), graph.n_rows, None, None, ptr(feat), 1, C,
             LRELU_SLOPE, ptr(w), None, ptr(out), ACT_NONE, None, 0.0, 0, hub.ptr, ptr(scr), _stream())
        ctx.graph = graph
        ctx.save_for_backward(w)
        return out

    @staticmethod
    def backward(ctx, dout):
        (w,) = ctx.saved_tensors
        g = ctx.graph
        dout = _c(dout)
        colptr, rowidx, perm = g.transpose_structure()
        C = dout.shape[1]
        dfeat = torch.empty((g.n_cols, C), dtype=torch.float32, device=dout.device)
        call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), g.n_cols, ptr(w), ptr(dout), 1, C,
             ptr(dfeat), 0, None, None, 0.0, 0, g.hub_cols_plain().ptr, _stream())
        return dfeat, None, None


def spmm(graph: Graph, w, feat):
    return _Spmm.apply(feat, w.view(-1, 1), graph)


# ------------------------------------------------------------------------------------------------
# LLP distillation read-outs                                                   LLP.py:34-35,221,236-237
# ------------------------------------------------------------------------------------------------
class _KDCosine(torch.autograd.Function):
    """1 - mean_p cos(s[idx_s[p]], t[idx_t[p]]) with the row gathers fused (idx None -> row p)."""

    @staticmethod
    def forward(ctx, s, t, idx_s, idx_t, eps):
        s, t = _c(s), _c(t)
        if s.dim() != 2 or t.dim() != 2 or s.shape[1] != t.shape[1]:
            raise ValueError("kd_cosine: s and t must be (rows, C) with the same C")
        P = idx_s.numel() if idx_s is not None else (idx_t.numel() if idx_t is not None else s.shape[0])
        for idx, tab in ((idx_s, s), (idx_t, t)):
            if idx is not None and idx.numel() != P:
                raise ValueError("kd_cosine: index vectors must have the same length")
            if idx is None and tab.shape[0] != P:
                raise ValueError("kd_cosine: un-indexed operand must have one row per pair")
        cosv = torch.empty(P, dtype=torch.float32, device=s.device)
        loss = torch.empty((), dtype=torch.float32, device=s.device)
        status = torch.empty(1, dtype=torch.int32, device=s.device)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_loss_workspace_bytes(), s.device)
        call("msha_kd_cosine_fwd", ptr(s), ptr(t), ptr(idx_s, torch.int64), ptr(idx_t, torch.int64), P, s.shape[1],
             s.shape[0], t.shape[0], eps, ptr(cosv), loss.data_ptr(), ptr(status, I32), ws.data_ptr(), ws.numel(), _stream())
        ctx.eps = eps
        ctx.save_for_backward(s, t, idx_s, idx_t, cosv)
        return loss

    @staticmethod
    def backward(ctx, g):
        s, t, idx_s, idx_t, cosv = ctx.saved_tensors
        ds = torch.zeros_like(s) if ctx.needs_input_grad[0] else None
        dt = torch.zeros_like(t) if ctx.needs_input_grad[1] else None
        call("msha_kd_cosine_bwd", ptr(s), ptr(t), ptr(idx_s, torch.int64), ptr(idx_t, torch.int64), cosv.numel(), s.shape[1],
             s.shape[0], t.shape[0], ctx.eps, ptr(cosv), _c(g).data_ptr(), ptr(ds), ptr(dt), _stream())
        return ds, dt, None, None, None


def _index(idx):
    if idx is None:
        return None
    return (idx if idx.dtype == torch.int64 else idx.long()).contiguous()


def kd_cosine(s, t, idx_s=None, idx_t=None, eps=1e-8):
    """``KD_cosine`` (LLP.py:34-35): ``1 - cosine_similarity(s[idx_s], t[idx_t].detach(), dim=-1).mean()``; the teacher
    operand is detached like the reference's.  Indices must lie inside the tables: torch's indexing would trip a device
    assert, this kernel (which never synchronises) lets such a pair contribute cos = 0 and sets its status word."""
    return _KDCosine.apply(s, t.detach(), _index(idx_s), _index(idx_t), float(eps))


class _MseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        if a.shape != b.shape:
            raise ValueError(f"mse_loss: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
        a, b = _c(a), _c(b)
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        lib = ops._lib.lib()
        ws = workspace(lib.msha_loss_workspace_bytes(), a.device)
        call("msha_mse_loss_fwd", ptr(a), ptr(b), a.numel(), loss.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        call("msha_mse_loss_bwd", ptr(a), ptr(b), a.numel(), _c(g).data_ptr(), ptr(da), ptr(db), _stream())
        return da, db


def mse_loss(a, b):
    """``torch.nn.MSELoss()(a, b)`` (mean over every element, LLP.py:221,237)."""
    return _MseLoss.apply(a, b)


# ------------------------------------------------------------------------------------------------
# GraphSAGE baseline: adj[source_index] * x                                              SGAE.py:53
# ------------------------------------------------------------------------------------------------
class _CsrRowsMul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, graph: Graph, w, src):
        x = _c(x)
        B, M = x.shape
        if M != graph.n_cols:
            raise RuntimeError(f"adj[source_index] * x: x has {M} columns, the adjacency {graph.n_cols}")
        out = torch.empty_like(x)
        call("msha_csr_rows_mul", ptr(graph.rowptr, I32), ptr(graph.col, I32), ptr(w), ptr(src, torch.int64), B,
             graph.n_rows, M, ptr(x), ptr(out), _stream())
        ctx.graph = graph
        ctx.save_for_backward(w, src)
        return out

    @staticmethod
    def backward(ctx, dout):
        w, src = ctx.saved_tensors
        g = ctx.graph
        dout = _c(dout)
        dx = torch.empty_like(dout)
        call("msha_csr_rows_mul", ptr(g.rowptr, I32), ptr(g.col, I32), ptr(w), ptr(src, torch.int64), dout.shape[0],
             g.n_rows, dout.shape[1], ptr(dout), ptr(dx), _stream())
        return dx, None, None, None


def csr_rows_mul(graph: Graph, x, src=None, values=None):
    """``adj[src] * x`` for a (B, M) block ``x`` (rows of ``adj`` selected by ``src``; values default to the stored ones)."""
    w = graph.val if values is None else values
    return _CsrRowsMul.apply(x, graph, _c(w), _index(src))


# ------------------------------------------------------------------------------------------------
# attention export                                                   train.py:284-321, Explainer.py:25-30
# ------------------------------------------------------------------------------------------------
def segment_argmax(ptr_arr, w, heads=1, head=-1, perm=None):
    """Per item of a CSR / CSC pointer array: (max, smallest slot attaining it or -1, tie count) of the per-edge weights."""
    w = _c(w.detach())
    n = ptr_arr.numel() - 1
    vmax = torch.empty(n, dtype=torch.float32, device=w.device)
    first = torch.empty(n, dtype=I32, device=w.device)
    ties = torch.empty(n, dtype=I32, device=w.device)
    call("msha_segment_argmax", ptr(ptr_arr, I32), ptr(perm, I32), ptr(w), heads, head, n, ptr(vmax), ptr(first, I32),
         ptr(ties, I32), _stream())
    return vmax, first, ties
