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


def partition_graph(rows: torch.Tensor, cols: torch.Tensor, part: Partition) -> Graph:
    """Local CSR of this rank: ``rows`` (global ids in [lo, hi)) and ``cols`` (global ids) of the edges whose
    destination row the rank owns.  Columns are remapped to the padded gathered indexing."""
    if rows.numel() and (int(rows.min()) < part.lo or int(rows.max()) >= part.hi):
        raise IndexError("partition_graph: a row id lies outside this rank's node range")
    g = Graph.from_coo(rows.to(torch.int64) - part.lo, part.to_padded(cols), part.n_local, part.n_padded)
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
                global_pairs=None, p2p=None):
    """Data-parallel link scoring: this rank scores its own pair shard against the gathered embeddings; the backward
    reduce-scatters d h to the owning ranks.  With ``target`` the rank's share of the GLOBAL mean nll read-out is
    returned instead of the scores (fused scorer + loss backward): local mean * P_local / P_global, so that the sum
    over ranks -- which is what the summed gradients of ``allreduce_gradients`` and of the reduce-scatter amount to --
    is the mean over all pairs, exactly the single-GPU loss.  ``global_pairs``: total pair count over the ranks
    (all-reduced here when not given: every rank must then make this call)."""
    h_g = p2p.gather(h_local, key="score_h") if p2p is not None else all_gather_rows(h_local, part, group)
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


# =================================================================================================
# Peer-memory data path: the gather / reduce-scatter of the column-side tensors over NVLink peer mappings
# (msha_gnn_b200/peer.py, csrc/peer_kernels.cu) instead of NCCL collectives.
#
#   flat       (blocks below PIPELINE_MIN_BLOCK_BYTES: launch-latency regime) one SM kernel pulls every remote block;
#              the gradient reduce-scatter is one kernel summing the peers' buffers in place over NVLink.
#   pipelined  (large blocks: bandwidth regime) copy engines pull one owner block after the other while the attention
#              forward consumes the blocks that have arrived (msha_gat_fwd_block); in the backward the column pass
#              runs owner block by owner block, each finished block is pulled by its owner's copy engine while the
#              row pass runs, and a local sum finishes the reduce-scatter.  No SM is spent on the transfers.
# =================================================================================================
from . import peer as _peer
from .graph import Hub, default_seg_limit
from .ops import call, ptr, _stream, LRELU_SLOPE

I32 = torch.int32
PIPELINE_MIN_BLOCK_BYTES = int(os.environ.get("MSHA_PIPELINE_MIN_BYTES", str(16 << 20)))
PIPELINE_STAGES = int(os.environ.get("MSHA_PIPELINE_STAGES", "4"))


def _alias_rows(base: torch.Tensor, lo: int, n: int) -> torch.Tensor:
    """Rows [lo, lo + n) of a persistent 2-D buffer as a tensor that shares its storage but is NOT an autograd view of it:
    an op that writes its result there (``Fn.linear(..., out=)``, mark_dirty) must not rebase the buffer's history --
    the buffer outlives the step and would chain every step's graph."""
    return torch.empty(0, dtype=base.dtype, device=base.device).set_(
        base.untyped_storage(), base.storage_offset() + lo * base.stride(0), (n, base.shape[1]), base.stride())


class _Exchange:
    """Buffers and flag channels of one gathered tensor.  ``buf`` / ``grad``: [world * n_max, C] on every rank; a rank
    writes its own block of ``buf`` (forward) and all of ``grad`` (its contributions to everybody's rows, backward)."""

    def __init__(self, pg, part: Partition, C: int):
        self.C = int(C)
        self.buf = pg.alloc((part.n_padded, C))
        self.grad = pg.alloc((part.n_padded, C), zero=True)
        self.ch_ready, self.ch_done, self.ch_gready, self.ch_gdone = (pg.new_channel() for _ in range(4))
        self.seq = 0          # forward gathers so far (flag value of ch_ready / ch_done)
        self.gcount = 0       # backward reduce-scatters so far (ch_gready / ch_gdone)


class OwnerBlocks:
    """Per-graph launch structure of the pipelined path: for every row the CSR slot where each owner's columns start
    (columns are owner-major and sorted inside a row, so an owner block is a contiguous slot range), the forward stages
    (contiguous owner ranges in consumption order) with their hub-row segments, and the hub-column segments of every
    owner block of the CSC."""

    def __init__(self, graph: Graph, part: Partition, n_stages: int):
        W, r, n_max = part.world, part.rank, part.n_max
        rp, col = graph.attention_csr()
        N = graph.n_rows
        dev = graph.device
        deg = (rp[1:] - rp[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(N, device=dev), deg)
        key = rows * part.n_padded + col.long()
        q = (torch.arange(N, device=dev).view(-1, 1) * part.n_padded + torch.arange(W + 1, device=dev).view(1, -1) * n_max)
        self.blk = torch.searchsorted(key, q.reshape(-1)).view(N, W + 1).t().contiguous().to(I32)      # [W + 1, N]
        del rows, key, q
        self.seg_limit = default_seg_limit(graph.nnz)
        # stage 0: the own block (nothing to wait for); the remote owners in ring order r+1, r+2, ... split into groups
        remote = list(range(1, W))
        n_groups = max(1, min(n_stages - 1, len(remote))) if remote else 0
        groups, start = [], 0
        for g in range(n_groups):                       # later groups are larger: the first remote stage starts early
            size = (len(remote) - start) // (n_groups - g)
            groups.append(remote[start:start + size])
            start += size
        self.stage_offsets = [[0]] + groups
        self.fwd = []
        for offs in self.stage_offsets:
            owners = sorted((r + s) % W for s in offs)
            runs, a = [], 0
            while a < len(owners):                      # maximal runs of consecutive owner ids
                b = a
                while b + 1 < len(owners) and owners[b + 1] == owners[b] + 1:
                    b += 1
                runs.append((owners[a], owners[b]))
                a = b + 1
            launches = []
            for o1, o2 in runs:
                beg, end = self.blk[o1], self.blk[o2 + 1]
                launches.append((beg, end, Hub(beg=beg, end=end, seg_limit=self.seg_limit)))
            self.fwd.append(launches)
        colptr, _, _ = graph.attention_csc()
        self.col_hubs = [Hub(colptr[o * n_max:(o + 1) * n_max + 1], seg_limit=self.seg_limit) for o in range(W)]


class P2P:
    """One rank's state of the peer-memory data path: exchanges by key (created in first-use order, which must be the
    same on every rank -- it is, the ranks run the same model code), copy streams, staging for pulled gradient blocks."""

    def __init__(self, pg, part: Partition, n_copy_streams: int = 1):
        self.pg, self.part = pg, part
        self.ex = {}
        dev = pg.device
        self.copy_streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, n_copy_streams))]
        self.score_stream = torch.cuda.Stream(device=dev)
        self._staging = {}
        self._blocks = {}

    def exchange(self, key, C) -> _Exchange:
        ex = self.ex.get(key)
        if ex is None:
            ex = self.ex[key] = _Exchange(self.pg, self.part, C)
        assert ex.C == C, (key, ex.C, C)
        return ex

    def staging(self, C):
        t = self._staging.get(C)
        if t is None:
            t = self._staging[C] = torch.empty((self.part.world, self.part.n_max, C), dtype=torch.float32, device=self.pg.device)
        return t

    def owner_blocks(self, graph) -> OwnerBlocks:
        ob = self._blocks.get(id(graph))
        if ob is None or ob[0] is not graph:
            ob = self._blocks[id(graph)] = (graph, OwnerBlocks(graph, self.part, PIPELINE_STAGES))
        return ob[1]

    def pipelined(self, C) -> bool:
        return self.part.world > 1 and self.part.n_max * C * 4 >= PIPELINE_MIN_BLOCK_BYTES

    # ---- reuse guards: a rank may overwrite what its peers read only after their "done" flags
    def guard(self, ex):
        self.pg.wait(ex.ch_done, ex.seq)

    def guard_grad(self, ex):
        self.pg.wait(ex.ch_gdone, ex.gcount)

    def own_rows(self, ex):
        """This rank's block of the gathered buffer (destination of the producing kernel), safe to overwrite."""
        self.guard(ex)
        return _alias_rows(ex.buf.local, self.part.rank * self.part.n_max, self.part.n_local)

    def block_rows(self, q) -> slice:
        return slice(q * self.part.n_max, q * self.part.n_max + self.part.sizes[q])

    # ---- forward: gather
    def pull_all(self, ex, seq):
        """Own block of ``ex.buf`` is complete on this stream: publish it, fetch everybody else's."""
        pg, part = self.pg, self.part
        W, r = part.world, part.rank
        pg.signal(ex.ch_ready, seq)
        if part.n_max * ex.C * 4 < _peer.CE_MIN_BYTES:
            pg.wait(ex.ch_ready, seq)
            pg.pull_blocks_sm(ex.buf, part.n_max)
            pg.signal(ex.ch_done, seq)
            return
        main = torch.cuda.current_stream()
        ev0 = torch.cuda.Event()
        ev0.record(main)
        used = []
        for s in range(1, W):
            q = (r + s) % W
            cs = self.copy_streams[(s - 1) % len(self.copy_streams)]
            if cs not in used:
                cs.wait_event(ev0)
                used.append(cs)
            with torch.cuda.stream(cs):
                pg.wait(ex.ch_ready, seq, 1 << q)
                pg.pull_block(ex.buf, q, self.block_rows(q))
                pg.signal(ex.ch_done, seq, 1 << q)
        for cs in used:
            main.wait_stream(cs)

    def gather(self, x_local, key):
        """[n_local, C] -> [world * n_max, C] (autograd: the backward is the reduce-scatter of the gathered gradient)."""
        ex = self.exchange(key, x_local.shape[1])
        out = _P2PGather.apply(x_local, self, ex)
        out._msha_grad_buffer = lambda: self.grad_buffer(ex)
        return out

    def grad_buffer(self, ex):
        """The peer-mapped gradient buffer of an exchange, safe to overwrite (producers write d gathered here)."""
        self.guard_grad(ex)
        return ex.grad.local

    # ---- backward: reduce-scatter of ex.grad (complete on this stream) -> [n_local, C]
    def reduce_scatter(self, ex):
        pg, part = self.pg, self.part
        W, r, n_max, C = part.world, part.rank, part.n_max, ex.C
        ex.gcount += 1
        g = ex.gcount
        blk_bytes = n_max * C * 4
        out = torch.empty((n_max, C), dtype=torch.float32, device=pg.device)
        pg.signal(ex.ch_gready, g)
        if blk_bytes < _peer.CE_MIN_BYTES:
            pg.wait(ex.ch_gready, g)
            pg.sum_into(out, [ex.grad.addr[q] + r * blk_bytes for q in range(W)], n_max * C)
            pg.signal(ex.ch_gdone, g)
            return out[:part.n_local]
        stg = self.staging(C)
        main = torch.cuda.current_stream()
        ev0 = torch.cuda.Event()
        ev0.record(main)
        own = self.block_rows(r)
        used = []
        for s in range(1, W):
            q = (r - s) % W
            cs = self.copy_streams[(s - 1) % len(self.copy_streams)]
            if cs not in used:
                cs.wait_event(ev0)
                used.append(cs)
            with torch.cuda.stream(cs):
                pg.wait(ex.ch_gready, g, 1 << q)
                stg[q, :part.n_local].copy_(ex.grad.views[q][own], non_blocking=True)
                pg.signal(ex.ch_gdone, g, 1 << q)
        for cs in used:
            main.wait_stream(cs)
        addrs = [ex.grad.addr[r] + r * blk_bytes if q == r else stg[q].data_ptr() for q in range(W)]
        pg.sum_into(out, addrs, part.n_local * C)
        return out[:part.n_local]


class _P2PGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p2p: P2P, ex: _Exchange):
        part = p2p.part
        dst = ex.buf.local[part.local_slice_of_padded()]
        if x.data_ptr() != dst.data_ptr():               # otherwise the producer wrote straight into p2p.own_rows(ex)
            p2p.guard(ex)
            dst.copy_(x)
        ex.seq += 1
        p2p.pull_all(ex, ex.seq)
        ctx.p2p, ctx.ex = p2p, ex
        return _alias_rows(ex.buf.local, 0, part.n_padded)

    @staticmethod
    def backward(ctx, g):
        p2p, ex = ctx.p2p, ctx.ex
        if g.data_ptr() != ex.grad.local.data_ptr():     # otherwise the producer wrote straight into p2p.grad_buffer(ex)
            p2p.guard_grad(ex)
            ex.grad.local.copy_(g)
        return p2p.reduce_scatter(ex), None, None


class _BufferSink:
    """Hands the attention backward the peer-mapped destinations of d feat_nbr / d s_nbr (functional._AttentionBlock)."""

    def __init__(self, p2p, exW, exS):
        self.p2p, self.exW, self.exS = p2p, exW, exS

    @property
    def feat_grad(self):
        return self.p2p.grad_buffer(self.exW)

    @property
    def score_grad(self):
        return self.p2p.grad_buffer(self.exS)


class _P2PAttention(torch.autograd.Function):
    """Gather + attention of one GAT layer with the transfers hidden under the kernels (pipelined mode, see above).
    Same arithmetic as ``functional._AttentionBlock`` on the gathered tensors (Ours.py:64-69,98 / Ablation.py:262-274)."""

    @staticmethod
    def forward(ctx, Wh_own, s_nbr_own, s_self, p2p: P2P, exW, exS, graph, H, D, act, p, seed):
        pg, part = p2p.pg, p2p.part
        W, r, n_max = part.world, part.rank, part.n_max
        C = H * D
        ob = p2p.owner_blocks(graph)
        own = part.local_slice_of_padded()
        main = torch.cuda.current_stream()
        dev = Wh_own.device
        dstW = exW.buf.local[own]
        if Wh_own.data_ptr() != dstW.data_ptr():
            p2p.guard(exW)
            dstW.copy_(Wh_own)
        p2p.guard(exS)
        exS.buf.local[own].copy_(s_nbr_own)
        exW.seq += 1
        exS.seq = seq = exW.seq
        pg.signal(exW.ch_ready, seq)                      # one flag covers both own blocks
        ev0 = torch.cuda.Event()
        ev0.record(main)
        # the H scores per node of every peer (small): one SM pull on a side stream
        ss = p2p.score_stream
        ss.wait_event(ev0)
        with torch.cuda.stream(ss):
            pg.wait(exW.ch_ready, seq)
            pg.pull_blocks_sm(exS.buf, n_max)
            pg.signal(exS.ch_done, seq)
            ev_s = torch.cuda.Event()
            ev_s.record(ss)
        # the feature blocks: copy-engine pulls in consumption order
        stage_events = []
        for k, offs in enumerate(ob.stage_offsets[1:]):
            cs = p2p.copy_streams[k % len(p2p.copy_streams)]
            if k < len(p2p.copy_streams):
                cs.wait_event(ev0)
            with torch.cuda.stream(cs):
                for s in offs:
                    q = (r + s) % W
                    pg.wait(exW.ch_ready, seq, 1 << q)
                    pg.pull_block(exW.buf, q, p2p.block_rows(q))
                    pg.signal(exW.ch_done, seq, 1 << q)
                e = torch.cuda.Event()
                e.record(cs)
                stage_events.append(e)
        rp, col = graph.attention_csr()
        N, E = graph.n_rows, col.numel()
        s_nbr_g, Wh_g = exS.buf.local, exW.buf.local
        s_self = s_self.contiguous()
        lse = torch.empty((N, 2 * H), dtype=torch.float32, device=dev)      # per row: H maxima, H sums
        alpha = torch.empty((E, H), dtype=torch.float32, device=dev)
        out = torch.empty((N, C), dtype=torch.float32, device=dev)
        hub = graph.hub_rows()
        scr = torch.empty(2 * H * hub.n_segs, dtype=torch.float32, device=dev) if hub.n_segs else None
        main.wait_event(ev_s)
        call("msha_gat_softmax_stats", ptr(rp, I32), ptr(col, I32), N, ptr(s_nbr_g), ptr(s_self), H, LRELU_SLOPE, ptr(lse),
             hub.ptr, ptr(scr), _stream())
        first = True
        for k, launches in enumerate(ob.fwd):
            if k > 0:
                main.wait_event(stage_events[k - 1])
            for beg, end, bhub in launches:
                call("msha_gat_fwd_block", ptr(beg, I32), ptr(end, I32), ptr(col, I32), N, ptr(s_nbr_g), ptr(s_self), ptr(lse),
                     ptr(Wh_g), H, D, LRELU_SLOPE, ptr(alpha), ptr(out), 0 if first else 1, p, seed, bhub.ptr, _stream())
                first = False
        if act != ACT_NONE:
            call("msha_act_fwd", ptr(out), ptr(out), out.numel(), act, LRELU_SLOPE, _stream())
        ctx.p2p, ctx.exW, ctx.exS, ctx.graph = p2p, exW, exS, graph
        ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed = H, D, act, p, seed
        ctx.save_for_backward(s_nbr_g, s_self, Wh_g, alpha, out if act != ACT_NONE else None)
        return out

    @staticmethod
    def backward(ctx, d_out):
        s_nbr_g, s_self, Wh_g, alpha, out = ctx.saved_tensors
        p2p, exW, exS, graph = ctx.p2p, ctx.exW, ctx.exS, ctx.graph
        H, D, act, p, seed = ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed
        pg, part = p2p.pg, p2p.part
        W, r, n_max = part.world, part.rank, part.n_max
        C = H * D
        ob = p2p.owner_blocks(graph)
        rp, col = graph.attention_csr()
        colptr, rowidx, perm = graph.attention_csc()
        N = graph.n_rows
        dev = alpha.device
        main = torch.cuda.current_stream()
        d_out = d_out.contiguous()
        if act != ACT_NONE:
            dz = torch.empty_like(d_out)
            call("msha_act_bwd", ptr(d_out), ptr(out), ptr(dz), N * C, act, LRELU_SLOPE, _stream())
        else:
            dz = d_out
        dWh_g = p2p.grad_buffer(exW)                      # guards: the peers have pulled the previous contents
        ds_g = p2p.grad_buffer(exS)
        exW.gcount += 1
        exS.gcount = g = exW.gcount
        ev0 = torch.cuda.Event()
        ev0.record(main)
        # column pass (d Wh_j = sum_i alpha_ij dz_i) owner block by owner block, in the order the owners consume them
        for s in list(range(1, W)) + [0]:
            o = (r + s) % W
            lo = o * n_max
            call("msha_spmm_csc", colptr.data_ptr() + 4 * lo, ptr(rowidx, I32), ptr(perm, I32), n_max, ptr(alpha), ptr(dz), H, D,
                 dWh_g.data_ptr() + 4 * lo * C, 0, None, None, p, seed, ob.col_hubs[o].ptr, _stream())
            if s:
                pg.signal(exW.ch_gready, g, 1 << o)
        # meanwhile the copy engine fetches this rank's block from the peers, in the order they finish it
        stg = p2p.staging(C)
        own = p2p.block_rows(r)
        cs = p2p.copy_streams[0]
        cs.wait_event(ev0)
        with torch.cuda.stream(cs):
            for s in range(1, W):
                q = (r - s) % W
                pg.wait(exW.ch_gready, g, 1 << q)
                stg[q, :part.n_local].copy_(exW.grad.views[q][own], non_blocking=True)
                pg.signal(exW.ch_gdone, g, 1 << q)
            ev_p = torch.cuda.Event()
            ev_p.record(cs)
        # row pass: d alpha, softmax / LeakyReLU backward, d s_self
        E = alpha.shape[0]
        dlogit = torch.empty_like(alpha)
        ds_self = torch.empty((N, H), dtype=torch.float32, device=dev)
        hub = graph.hub_rows()
        r_buf = torch.empty((N, H), dtype=torch.float32, device=dev) if hub.n_segs else None
        call("msha_gat_bwd_rows", ptr(rp, I32), ptr(col, I32), N, ptr(s_nbr_g), ptr(s_self), LRELU_SLOPE, ptr(alpha),
             ptr(Wh_g), ptr(dz), None, ACT_NONE, None, None, None, None, None, H, D, ptr(dlogit), ptr(ds_self), p, seed,
             hub.ptr, ptr(r_buf), int(E // max(N, 1)), _stream())
        # d s_nbr: column sums of d logit into the peer-mapped buffer; summed over the peers' buffers in place (small)
        call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), part.n_padded, None, None, H, D, None, 0,
             ptr(dlogit), ptr(ds_g), p, seed, graph.hub_cols().ptr, _stream())
        pg.signal(exS.ch_gready, g)
        main.wait_event(ev_p)
        blk_bytes = n_max * C * 4
        dWh = torch.empty((part.n_local, C), dtype=torch.float32, device=dev)
        addrs = [exW.grad.addr[r] + r * blk_bytes if q == r else stg[q].data_ptr() for q in range(W)]
        pg.sum_into(dWh, addrs, part.n_local * C)
        pg.wait(exS.ch_gready, g)
        ds_nbr = torch.empty((n_max, H), dtype=torch.float32, device=dev)
        pg.sum_into(ds_nbr, [exS.grad.addr[q] + r * n_max * H * 4 for q in range(W)], n_max * H)
        pg.signal(exS.ch_gdone, g)
        return dWh, ds_nbr[:part.n_local], ds_self, None, None, None, None, None, None, None, None, None


def gat_encode_p2p(convs, x_local, pgraph: Graph, part: Partition, p2p: P2P, training=True):
    """``gat_encode`` over the peer-memory data path (same arithmetic, no NCCL call between the kernels)."""
    h = x_local
    pgraph.attention_csc()                                # host-synchronising one-off builds happen before any flag wait
    pgraph.hub_rows()
    pgraph.hub_cols()
    for l, conv in enumerate(convs):
        H, D = conv.heads, conv.out_features
        C = H * D
        if p2p.pipelined(C):
            p2p.owner_blocks(pgraph)
        exW = p2p.exchange(f"gat{l}.Wh", C)
        exS = p2p.exchange(f"gat{l}.s", H)
        Wh = Fn.linear(h, conv.W, out=p2p.own_rows(exW))                      # straight into this rank's block
        s_nbr, s_self = Fn.node_scores(Wh, conv.a_nbr, conv.a_self, H, D)
        fuse_elu = conv.activation == "elu" and conv.concat
        act = ACT_ELU if fuse_elu else ACT_NONE
        p = float(conv.dropout) if training else 0.0
        if p2p.pipelined(C) and (H & (H - 1)) == 0 and D % 4 == 0:
            from . import ops as _ops
            seed = _ops.next_seed() if p > 0 else 0
            out = _P2PAttention.apply(Wh, s_nbr, s_self, p2p, exW, exS, pgraph, H, D, act, p, seed)
        else:
            Wh_g = _P2PGather.apply(Wh, p2p, exW)
            s_nbr_g = _P2PGather.apply(s_nbr, p2p, exS)
            out, _ = Fn.attention_block(pgraph, s_nbr_g, s_self, Wh_g, heads=H, act=act, dropout_p=conv.dropout,
                                        training=training, grad_sink=_BufferSink(p2p, exW, exS))
        if not conv.concat:
            out = out.view(out.shape[0], H, D).mean(dim=1)
            if conv.activation == "elu":
                out = Fn.elu(out)
        h = out
    return h
