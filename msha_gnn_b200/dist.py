"""Multi-GPU execution of the hot path: contiguous destination-node ranges per rank.

Rank r owns rows ``[bounds[r], bounds[r+1])`` of the CSR, the matching rows of X / Wh / s_self, the outputs and
their gradients.  The one exchange per layer and direction is on the *column* side: neighbours of local rows live
anywhere, so the forward all-gathers the per-node tensors (``Wh``, ``s_nbr``) and the backward reduce-scatters
their gradients -- NCCL over NVLink 5 / NVSwitch (uniform all-to-all, so plain collectives; no ring ordering).
Score batches are split data-parallel; parameter gradients are all-reduced once per step.

Every rank pads its rows to ``n_max`` so the gathered buffer is ``[world * n_max, C]`` and global node ids are
remapped once, at partition time, to ``owner * n_max + (id - bounds[owner])`` -- no compaction copy after a gather.

The reference has no distributed code at all; this module is new functionality behind the
same kernels.  Host-side logic (bounds, id remap, collective autograd, gradient all-reduce) also runs on CPU
tensors under the gloo backend for the tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import functional as Fn
from .graph import Graph
from .ops import ACT_ELU, ACT_NONE


class Partition:
    """Contiguous node ranges.  ``bounds`` has world+1 entries; pass ``rowptr`` to balance edges instead of nodes."""

    def __init__(self, n_nodes: int, world: int, rank: int, bounds=None):
        self.n_nodes, self.world, self.rank = int(n_nodes), int(world), int(rank)
        if bounds is None:
            bounds = [(n_nodes * r) // world for r in range(world + 1)]
        self.bounds = [int(b) for b in bounds]
        assert len(self.bounds) == world + 1 and self.bounds[0] == 0 and self.bounds[-1] == n_nodes
        sizes = [self.bounds[r + 1] - self.bounds[r] for r in range(world)]
        # block stride of the gathered layout: a multiple of 4 rows, so that every block of an [n, H] or [n, C] fp32
        # buffer starts 16-byte aligned and holds whole 128-bit vectors (peer pulls / sums, csrc/peer_kernels.cu)
        self.n_max = (max(sizes) + 3) // 4 * 4 if sizes else 0
        self.sizes = sizes
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.hi - self.lo
        self.n_padded = self.n_max * world
        self._bt = {}

    @staticmethod
    def edge_balanced(rowptr_cpu: torch.Tensor, world: int, rank: int) -> "Partition":
        """Split points chosen on the cumulative edge count (power-law graphs).  Every rank keeps
        at least one row (a hub row heavier than total / world would otherwise produce equal cuts = empty ranks)."""
        n = rowptr_cpu.numel() - 1
        if n < world:
            raise ValueError(f"edge_balanced: {n} rows cannot be split over {world} ranks")
        total = int(rowptr_cpu[-1])
        targets = torch.tensor([(total * r) // world for r in range(1, world)], dtype=rowptr_cpu.dtype)
        cuts = torch.searchsorted(rowptr_cpu, targets).clamp_(0, n).tolist()
        for i in range(len(cuts)):                       # strictly increasing, room left for the ranks after
            lo = (cuts[i - 1] if i else 0) + 1
            cuts[i] = min(max(cuts[i], lo), n - (len(cuts) - i))
        return Partition(n, world, rank, [0] + cuts + [n])

    def _tables(self, device):
        """(inner bounds, block shift) on ``device``, uploaded once (a host -> device copy per call would synchronise)."""
        t = self._bt.get(device)
        if t is None:
            inner = torch.tensor(self.bounds[1:-1], dtype=torch.int64, device=device)
            shift = torch.tensor([r * self.n_max - self.bounds[r] for r in range(self.world)], dtype=torch.int64, device=device)
            t = self._bt[device] = (inner, shift)
        return t

    def owner_of(self, ids: torch.Tensor) -> torch.Tensor:
        return torch.bucketize(ids, self._tables(ids.device)[0], right=True)

    def to_padded(self, ids: torch.Tensor) -> torch.Tensor:
        """global node id -> index into the gathered [world * n_max, C] buffer."""
        ids = ids.to(torch.int64)
        inner, shift = self._tables(ids.device)
        return ids + shift[torch.bucketize(ids, inner, right=True)]

    def local_slice_of_padded(self):
        return slice(self.rank * self.n_max, self.rank * self.n_max + self.n_local)


class GradSink:
    """Early start of a reduce-scatter: the op that produces the gradient of a gathered tensor calls ``start(g)`` the
    moment ``g`` is complete; the collective then runs on NCCL's stream while that op finishes its other outputs, and
    ``_AllGatherRows.backward`` only waits for it."""

    def __init__(self, part: Partition, group=None):
        self.part, self.group, self.work, self.out, self.src = part, group, None, None, None

    def start(self, g: torch.Tensor):
        part = self.part
        if part.world == 1 or not g.is_cuda or dist.get_backend(self.group) == "gloo":
            return
        self.src = g                                      # keep the buffer alive until the collective has read it
        self.out = g.new_empty((part.n_max, g.shape[1]))
        self.work = dist.reduce_scatter_tensor(self.out, g, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def take(self, g: torch.Tensor):
        """Result of the started collective if it was started on exactly this gradient, else None."""
        if self.work is None or self.src is None or self.src.data_ptr() != g.data_ptr():
            return None
        self.work.wait()                                  # current stream waits for NCCL's stream
        out, self.work, self.src, self.out = self.out, None, None, None
        return out


class _AllGatherRows(torch.autograd.Function):
    """[n_local, C] -> [world * n_max, C]; backward: reduce-scatter (sum) of the gathered gradient."""

    @staticmethod
    def forward(ctx, x, part: Partition, group, sink=None):
        ctx.part, ctx.group, ctx.sink = part, group, sink
        C = x.shape[1]
        padded = x
        if part.n_local != part.n_max:
            padded = x.new_zeros((part.n_max, C))
            padded[: part.n_local] = x
        out = x.new_empty((part.n_padded, C))
        if part.world == 1:
            out.copy_(padded)
        else:
            dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        part, group = ctx.part, ctx.group
        g = g.contiguous()
        if part.world == 1:
            return g[: part.n_local], None, None, None
        early = ctx.sink.take(g) if ctx.sink is not None else None
        if early is not None:
            return early[: part.n_local], None, None, None
        out = g.new_empty((part.n_max, g.shape[1]))
        if dist.get_backend(group) == "gloo":            # gloo has no reduce_scatter: all-reduce + slice (tests only)
            dist.all_reduce(g, group=group)
            out.copy_(g[part.rank * part.n_max:(part.rank + 1) * part.n_max])
        else:
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=group)
        return out[: part.n_local], None, None, None


def all_gather_rows(x, part: Partition, group=None, sink: GradSink = None):
    return _AllGatherRows.apply(x, part, group, sink)


def allreduce_gradients(params, group=None, world=None):
    """Sum parameter gradients over ranks (each rank holds the contribution of its rows / pairs).  Every parameter of
    the list takes part on every rank -- a rank that produced no gradient for one (an empty shard) contributes zeros --
    so the flat message has the same length everywhere."""
    world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
    if world == 1:
        return
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])


def partition_graph(rows: torch.Tensor, cols: torch.Tensor, part: Partition, col_part: Partition = None) -> Graph:
    """Local CSR of this rank: ``rows`` (global ids in [lo, hi)) and ``cols`` (global ids) of the edges whose
    destination row the rank owns.  Columns are remapped to the padded gathered indexing of ``col_part`` (bipartite
    graphs: the recipients' partition; default: the rows' own partition)."""
    if rows.numel() and (int(rows.min()) < part.lo or int(rows.max()) >= part.hi):
        raise IndexError("partition_graph: a row id lies outside this rank's node range")
    col_part = part if col_part is None else col_part
    g = Graph.from_coo(rows.to(torch.int64) - part.lo, col_part.to_padded(cols), part.n_local, col_part.n_padded)
    # A row without neighbours: the reference's uniform attention over "all columns" (GAT.py:29-31) would here run over
    # the padded column space of the gathered layout and aggregate pad rows; on partitioned graphs such rows aggregate
    # nothing instead (add self loops for the reference's 1/N average).
    g.isolated = "zero"
    return g


# Early reduce-scatter of d Wh under the attention row pass (GradSink).  Off by default: measured on B200 x2 / x8 (R-MAT,
# profiles/README.md) the collective's CTAs and the row pass compete for the same SMs -- no gain with a host sync per
# step (24.5 vs 24.5 ms at N = 2, 70 vs 73 ms at N = 8) and a loss when the host runs ahead (29 vs 25 ms at N = 2).
OVERLAP_DEFAULT = os.environ.get("MSHA_DIST_OVERLAP", "0") != "0"


def gat_encode(convs, x_local, pgraph: Graph, part: Partition, group=None, training=True, overlap=None):
    """Partitioned forward of a stack of ``GATConv`` layers (same arithmetic as ``GATConv.forward``): per layer an
    all-gather of Wh and of s_nbr in the forward and the reduce-scatter of their gradients in the backward."""
    overlap = OVERLAP_DEFAULT if overlap is None else overlap
    h = x_local
    for conv in convs:
        H, D = conv.heads, conv.out_features
        Wh = Fn.linear(h, conv.W)                                              # local rows only
        s_nbr, s_self = Fn.node_scores(Wh, conv.a_nbr, conv.a_self, H, D)
        # two collectives (features, then the H scores per node) straight into the buffers the kernels read: packing
        # them into one message costs two extra passes over the gathered N x C matrix in each direction
        sink = GradSink(part, group) if overlap else None     # backward: reduce-scatter of d Wh_g under the row pass
        Wh_g = all_gather_rows(Wh, part, group, sink)
        s_nbr_g = all_gather_rows(s_nbr, part, group)
        fuse_elu = conv.activation == "elu" and conv.concat
        out, _ = Fn.attention_block(pgraph, s_nbr_g, s_self, Wh_g, heads=H, act=ACT_ELU if fuse_elu else ACT_NONE,
                                    dropout_p=conv.dropout, training=training, grad_sink=sink)
        if not conv.concat:
            out = out.view(out.shape[0], H, D).mean(dim=1)
            if conv.activation == "elu":
                out = Fn.elu(out)
        h = out
    return h


def score_pairs(predictor, h_local, src_global, dst_global, part: Partition, group=None, target=None,
                global_pairs=None, p2p=None, key="score_h"):
    """Data-parallel link scoring: this rank scores its own pair shard against the gathered embeddings; the backward
    reduce-scatters d h to the owning ranks.  With ``target`` the rank's share of the GLOBAL mean nll read-out is
    returned instead of the scores (fused scorer + loss backward): local mean * P_local / P_global, so that the sum
    over ranks -- which is what the summed gradients of ``allreduce_gradients`` and of the reduce-scatter amount to --
    is the mean over all pairs, exactly the single-GPU loss.  ``global_pairs``: total pair count over the ranks
    (all-reduced here when not given: every rank must then make this call).  ``p2p``: gather / reduce-scatter over the
    peer-memory path (dist_p2p.py) through the exchange named ``key``."""
    h_g = p2p.gather(h_local, key=key) if p2p is not None else all_gather_rows(h_local, part, group)
    src_p, dst_p = part.to_padded(src_global), part.to_padded(dst_global)
    if target is None:
        return predictor.forward_pairs(h_g, h_g, src_p, dst_p)
    p_local = int(src_global.numel())
    if part.world > 1 and global_pairs is None:
        t = torch.tensor([p_local], dtype=torch.int64, device=h_local.device)
        dist.all_reduce(t, group=group)
        global_pairs = int(t.item())
    loss = predictor.nll_loss_pairs(h_g, h_g, src_p, dst_p, target)
    if part.world > 1:
        loss = loss * (p_local / max(int(global_pairs), 1))
    return loss


# the peer-memory data path (NVLink peer mappings instead of NCCL collectives on the data path): dist_p2p.py
from .dist_p2p import P2P, gat_encode_p2p  # noqa: E402,F401
