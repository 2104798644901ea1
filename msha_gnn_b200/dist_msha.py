"""The MSHA layer (``OursLayer`` / ``OursLayer3``, Ours.py:54-109, Ablation.py:260-277) on a graph that is partitioned over
the GPUs of one node (BASELINE.json configs[4]: 10 M sources / 10 M recipients / 500 M flow edges at 8 GPUs), SURVEY.md
section 8e:

  * sources (rows) and recipients (columns) are both split into contiguous ranges; rank r owns the rows
    ``part_s`` gives it (their S features, h2, u) and the recipients ``part_r`` gives it (their R features, h1, v);
  * forward: gather of ``h1`` / ``s_nbr`` over the recipient partition (every rank's rows may attend to any recipient),
    local attention + ``alpha @ h1``; the transposed aggregation ``alpha.T @ h2`` (Ours.py:100) produces a partial
    (M, C) table per rank that is reduce-scattered to the recipients' owners;
  * BatchNorm1d over the node axis (Ours.py:100-101): all-reduce of the 2*C column sums (fp64), forward and backward;
  * intra scales (Ours.py:71-90,99) in their group-sum form: per-group sums of ``coef[b] * h2[src_b]`` over the local
    batch rows -> all-reduce of the (n_groups, C) tables -> every local row adds its city's and its province's row;
  * read-out at scale: the dense ``elu(u @ v.T)`` (N, M) matrix (Ours.py:108-109) does not exist at 10 M x 10 M -- pairs are
    scored as ``elu(u_i . v_j)`` (``functional.pair_dot``) against the gathered ``v``.

``Comm`` hides the transport: NCCL / gloo collectives (``TorchComm``) or the peer-memory path (``PeerComm``).
The reference has no distributed code (train.py:18 is single-device).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import functional as Fn
from .dist import Partition, all_gather_rows
from .graph import Graph
from .ops import ACT_ELU, ACT_NONE, call, _stream


# ------------------------------------------------------------------------------------------------
# transports
# ------------------------------------------------------------------------------------------------
class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x (same on every rank).  Every rank's loss depends on y, the total loss is the sum of the
    ranks' losses: d x = sum over ranks of d y."""

    @staticmethod
    def forward(ctx, x, comm):
        ctx.comm = comm
        return comm.all_reduce(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ctx.comm.all_reduce(g.contiguous()), None


class _ReduceScatterRowsTorch(torch.autograd.Function):
    """[world * n_max, C] per-rank contributions -> [n_local, C] sum over ranks of the own block; backward: all-gather."""

    @staticmethod
    def forward(ctx, x, part, group):
        ctx.part, ctx.group = part, group
        x = x.contiguous()
        if part.world == 1:
            return x[: part.n_local].clone()
        out = x.new_empty((part.n_max, x.shape[1]))
        if dist.get_backend(group) == "gloo":            # gloo has no reduce_scatter: all-reduce + slice (tests only)
            y = x.clone()
            dist.all_reduce(y, group=group)
            out.copy_(y[part.rank * part.n_max:(part.rank + 1) * part.n_max])
        else:
            dist.reduce_scatter_tensor(out, x, op=dist.ReduceOp.SUM, group=group)
        return out[: part.n_local]

    @staticmethod
    def backward(ctx, g):
        part, group = ctx.part, ctx.group
        C = g.shape[1]
        padded = g.new_zeros((part.n_max, C))
        padded[: part.n_local] = g
        if part.world == 1:
            return padded, None, None
        out = g.new_empty((part.n_padded, C))
        dist.all_gather_into_tensor(out, padded, group=group)
        return out, None, None


class TorchComm:
    """Collectives of torch.distributed (NCCL on GPUs; gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def all_reduce(self, t):
        if self.world > 1:
            t = t.clone()
            dist.all_reduce(t, group=self.group)
        return t

    def gather_rows(self, x, part, key):
        return all_gather_rows(x, part, self.group)

    def reduce_scatter_rows(self, x, part, key):
        return _ReduceScatterRowsTorch.apply(x, part, self.group)


class _ReduceScatterRowsPeer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p2p, ex):
        ctx.p2p, ctx.ex = p2p, ex
        dst = ex.grads[0].local
        if x.data_ptr() != dst.data_ptr():
            p2p.ensure_grad_guard(ex)
            dst.copy_(x)
        (out,) = p2p.reduce_scatter(ex)
        return out

    @staticmethod
    def backward(ctx, g):
        p2p, ex = ctx.p2p, ctx.ex
        p2p.begin_produce(ex)
        p2p.own_rows(ex, 0).copy_(g)
        p2p.publish(ex)
        p2p.pull(ex)
        ex.produced = False
        return p2p.all_rows(ex, 0), None, None


class PeerComm:
    """The peer-memory path (dist_p2p.P2P): one exchange per gathered / reduce-scattered tensor, keyed by ``key`` and the
    partition it runs over (sources and recipients have their own ``P2P``)."""

    def __init__(self, p2p_by_part, halo=None):
        import os
        self.halo = (os.environ.get("MSHA_MSHA_HALO", "0") != "0") if halo is None else bool(halo)
        self.by_part = p2p_by_part                     # {id(partition): P2P}
        self.p2p0 = next(iter(p2p_by_part.values()))
        self.world = self.p2p0.part.world
        self._ar = None

    def _p2p(self, part):
        return self.by_part[id(part)]

    def gather_rows(self, x, part, key):
        return self._p2p(part).gather(x, key)

    def reduce_scatter_rows(self, x, part, key):
        p2p = self._p2p(part)
        ex = p2p.exchange(key, (x.shape[1],))
        return _ReduceScatterRowsPeer.apply(x, p2p, ex)

    # ---- halo exchange over the column partition of a graph (dist_p2p.HaloPlan): only the referenced recipients move
    def halo_plan(self, part, graph):
        """Opt-in (``PeerComm(..., halo=True)`` / MSHA_MSHA_HALO=1): on the cfg-5 graph at 2 GPUs every recipient is
        referenced by every rank and the halo exchange measured 82.4 ms against 80.7 ms for whole blocks; at 8 GPUs (where
        a rank's 62.5 M flows face 10 M recipients) it is built and parity-tested but not yet measured."""
        from . import dist_p2p as dp
        return self._p2p(part).halo_plan(graph) if self.halo and dp.HALO and self.world > 1 else None

    def halo_gather(self, x, part, plan, key):
        from . import dist_p2p as dp
        return dp.halo_gather(self._p2p(part), plan, x, key)

    def halo_scatter_add(self, x, part, plan, key):
        from . import dist_p2p as dp
        return dp.halo_scatter_add(self._p2p(part), plan, x, key)

    def all_reduce(self, t, cap_bytes=8 << 20):
        """Sum of a small tensor (fp32 or fp64) over the ranks through a peer-mapped two-slot scratch buffer: a rank
        publishes its slot, then reads every rank's slot in rank order.  Two slots alternate: a peer that has published
        gather n has finished reading gather n - 1, so the slot written for n + 1 is free without a further handshake."""
        p2p = self.p2p0
        pg = p2p.pg
        if self.world == 1:
            return t
        if self._ar is None:
            self._ar = dict(buf=pg.alloc((2, cap_bytes), torch.uint8), ch=pg.new_channel(), seq=0, cap=cap_bytes,
                            counter=torch.zeros(1, dtype=torch.int32, device=pg.device))
        st = self._ar
        n = t.numel()
        nbytes = n * t.element_size()
        pad = (-nbytes) % 16
        if nbytes + pad > st["cap"]:
            raise RuntimeError("PeerComm.all_reduce is for small tensors")
        st["seq"] += 1
        slot = st["seq"] & 1
        st["buf"].local[slot, :nbytes].copy_(t.contiguous().view(-1).view(torch.uint8))
        if pad:
            st["buf"].local[slot, nbytes:nbytes + pad].zero_()
        pg.signal(st["ch"], st["seq"])
        import ctypes
        from . import peer as _peer
        if t.dtype == torch.float64:
            out = torch.empty(n, dtype=torch.float64, device=t.device)
            call("msha_peer_allreduce_f64", out.data_ptr(), st["buf"].tab.data_ptr(), slot * st["cap"], n,
                 pg.flags.local.data_ptr(), pg.flags.tab.data_ptr(), pg.world, pg.rank, st["ch"], st["seq"] & 0xFFFFFFFF,
                 _peer.TIMEOUT_NS, pg.status.data_ptr(), _stream())
            return out.view(t.shape)
        assert t.dtype == torch.float32
        n4 = (n + 3) // 4 * 4
        out = torch.empty(n4, dtype=torch.float32, device=t.device)
        P, L, U = ctypes.c_void_p * 1, ctypes.c_int64 * 1, ctypes.c_uint64 * 1
        call("msha_peer_exchange_sum", 1, P(out.data_ptr()), U(st["buf"].tab.data_ptr()), L(slot * st["cap"]), L(n4),
             pg.flags.local.data_ptr(), pg.flags.tab.data_ptr(), pg.world, pg.rank, st["ch"], st["seq"] & 0xFFFFFFFF, -1, 0, -1, 0,
             _peer.TIMEOUT_NS, pg.status.data_ptr(), st["counter"].data_ptr(), p2p.max_ctas, _stream())
        return out[:n].view(t.shape)


def all_reduce_sum(x, comm):
    return _AllReduceSum.apply(x, comm)


# ------------------------------------------------------------------------------------------------
# the partitioned MSHA layer
# ------------------------------------------------------------------------------------------------
def _split_a(a, d):
    return a[:d, 0].reshape(1, d), a[d:, 0].reshape(1, d)


def _bn_heads_dist(x, bns, training, n_total, comm):
    gamma = torch.cat([b.weight for b in bns]) if len(bns) > 1 else bns[0].weight
    beta = torch.cat([b.bias for b in bns]) if len(bns) > 1 else bns[0].bias
    rm = torch.cat([b.running_mean for b in bns]) if len(bns) > 1 else bns[0].running_mean
    rv = torch.cat([b.running_var for b in bns]) if len(bns) > 1 else bns[0].running_var
    y = Fn.bn_lrelu_dist(x, gamma, beta, rm, rv, training, n_total, comm.all_reduce, bns[0].momentum, bns[0].eps,
                         grad_replicas=comm.world)
    if training:
        d = bns[0].num_features
        with torch.no_grad():
            for i, b in enumerate(bns):
                if len(bns) > 1:
                    b.running_mean.copy_(rm[i * d:(i + 1) * d])
                    b.running_var.copy_(rv[i * d:(i + 1) * d])
                b.num_batches_tracked += 1
    return y


def ours_encode(layers, S_local, R_local, pgraph: Graph, part_s: Partition, part_r: Partition, comm, source_index=None,
                city_ids=None, province_ids=None, group_sizes=None, training=True):
    """Head-batched ``OursLayer{,3}.forward`` up to (u, v) on this rank's shard (``layers.msha_heads_forward`` is the
    single-GPU form).  ``pgraph``: local source rows x recipient columns in the padded gathered indexing of ``part_r``
    (``dist.partition_graph``).  Intra scales (variant 1): ``source_index`` = local row ids of this rank's share of the
    batch, ``city_ids`` / ``province_ids`` = int64 group id of every LOCAL source, ``group_sizes`` = (float tensors of the
    GLOBAL member counts per city, per province).  Dropout on the intra attention is not supported on this path.
    -> (u [n_s_local, H*d'], v [n_r_local, H*d'])."""
    H = len(layers)
    variant = layers[0].variant
    d = layers[0].out_features
    p = layers[0].dropout
    C = H * d
    W1 = torch.cat([l.W1 for l in layers], dim=1) if H > 1 else layers[0].W1
    W2 = torch.cat([l.W2 for l in layers], dim=1) if H > 1 else layers[0].W2
    h1 = Fn.linear(R_local, W1)                                                  # own recipients   Ours.py:57
    h2 = Fn.linear(S_local, W2)                                                  # own sources      Ours.py:58
    a_nbr = torch.cat([_split_a(l.a, d)[0] for l in layers], dim=0)
    a_self = torch.cat([_split_a(l.a, d)[1] for l in layers], dim=0)
    s_nbr = Fn.node_scores(h1, a_nbr, None, H, d)
    s_self = Fn.node_scores(h2, a_self, None, H, d)
    plan = comm.halo_plan(part_r, pgraph) if hasattr(comm, "halo_plan") else None
    if plan is not None:
        # halo exchange: only the recipients this rank's flows reference (compact column numbering, dist_p2p.HaloPlan)
        pgraph = plan.graph
        h1_g = comm.halo_gather(h1, part_r, plan, "msha.h1")
        s_nbr_g = comm.halo_gather(s_nbr, part_r, plan, "msha.s")
    else:
        h1_g = comm.gather_rows(h1, part_r, "msha.h1")                           # every recipient's h1 / score
        s_nbr_g = comm.gather_rows(s_nbr, part_r, "msha.s")
    u_in, v_part, alpha = Fn.attention_block(pgraph, s_nbr_g, s_self, h1_g, h2, heads=H, act=ACT_NONE, dropout_p=p,
                                             training=training, want_cols=True)   # Ours.py:65-69,98,100
    if plan is not None:
        v_in = comm.halo_scatter_add(v_part, part_r, plan, "msha.v")
    else:
        v_in = comm.reduce_scatter_rows(v_part, part_r, "msha.v")                # alpha.T @ h2 summed over the ranks' rows
    if variant == 1:
        if source_index is None or city_ids is None or province_ids is None or group_sizes is None:
            raise ValueError("OursLayer needs source_index, city_ids, province_ids and group_sizes")
        if training and p > 0:
            raise NotImplementedError("partitioned MSHA layer: dropout on the intra attention is not supported")
        src = source_index.to(torch.int64).contiguous()
        B = src.numel()
        n3_all, n4_all = group_sizes
        a3 = torch.cat([(l.a3[:d, 0] + l.a3[d:, 0]).reshape(1, d) for l in layers], dim=0)      # Ours.py:74-75
        a4 = torch.cat([(l.a4[:d, 0] + l.a4[d:, 0]).reshape(1, d) for l in layers], dim=0)      # Ours.py:77-78
        h2b = h2.index_select(0, src)
        t3, t4 = Fn.node_scores(h2b, a3, a4, H, d)
        t3 = torch.nn.functional.leaky_relu(t3, 0.2)
        t4 = torch.nn.functional.leaky_relu(t4, 0.2)
        g3, g4 = city_ids[src], province_ids[src]
        n3, n4 = n3_all[g3].view(B, 1), n4_all[g4].view(B, 1)
        from .intra import _RowsumExp
        T = _RowsumExp.apply(alpha, pgraph, src, 0.0, 0)                                        # Ours.py:86 (all M columns)
        T = T - float(pgraph.n_cols - part_r.n_nodes)         # the padded column space is wider than the M real recipients
        total = n3 * torch.exp(t3) + n4 * torch.exp(t4) + T                                     # Ours.py:84-86
        c3, c4 = torch.exp(t3) / total, torch.exp(t4) / total                                   # Ours.py:87,89
        h2b3 = h2b.view(B, H, d)
        G3 = Fn.group_rows_sum((c3.view(B, H, 1) * h2b3).reshape(B, C), g3, n3_all.numel())
        G4 = Fn.group_rows_sum((c4.view(B, H, 1) * h2b3).reshape(B, C), g4, n4_all.numel())
        G3, G4 = all_reduce_sum(G3, comm), all_reduce_sum(G4, comm)
        u_in = u_in + Fn.group_rows_add(G3, city_ids, G4, province_ids)                         # Ours.py:99,101
    elif variant != 3:
        raise NotImplementedError("partitioned MSHA layer: variants 1 (OursLayer) and 3 (OursLayer3)")
    v = _bn_heads_dist(v_in, [l.bn1 for l in layers], training, part_r.n_nodes, comm)            # Ours.py:100
    u = _bn_heads_dist(u_in, [l.bn2 for l in layers], training, part_s.n_nodes, comm)            # Ours.py:101
    return u, v


def score_uv_pairs(u_local, v_local, src_local, dst_global, part_r: Partition, comm, heads=1):
    """``elu(u_i . v_j)`` per head for pairs (local source row, global recipient id): the entries (i, j) of the reference's
    dense ``F.elu(u @ v.t())`` (Ours.py:108-109), head h in column h.  -> [P, heads]."""
    v_g = comm.gather_rows(v_local, part_r, "msha.vg")
    dst_p = part_r.to_padded(dst_global)
    C = u_local.shape[1]
    d = C // heads
    # head h of node i is row i * heads + h of the (n * heads, d') view: no per-head copies of the tables
    u2, v2 = u_local.contiguous().view(-1, d), v_g.contiguous().view(-1, d)
    src_h, dst_h = src_local.to(torch.int64) * heads, dst_p * heads
    cols = [Fn.pair_dot(u2, v2, src_h + h, dst_h + h, act=ACT_ELU) for h in range(heads)]
    return torch.stack(cols, dim=1)
