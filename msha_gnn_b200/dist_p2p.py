"""Peer-memory data path of the node-partitioned model: the gather of the column-side tensors
(``Wh``, ``s_nbr``, final embeddings) and the reduce-scatter of their gradients run over NVLink peer mappings with this
library's own kernels and copy-engine pulls (peer.py, csrc/peer_kernels.cu) -- no collective-library call sits between
the attention kernels.  The reference is single-device (train.py:18); the arithmetic is that of ``dist.gat_encode``.

Two regimes, chosen per exchange from the block size (rows per rank x row bytes):

  flat       (launch-latency regime, e.g. the DDI-shaped graph: 4 MB blocks)
             forward : producer kernels -> msha_peer_signal -> ONE msha_peer_exchange_pull (every CTA waits for the flag of
                       the peer it copies from, the last CTA publishes "done")
             backward: producer kernels -> msha_peer_signal -> ONE msha_peer_exchange_sum (sums the peers' gradient blocks in
                       place over NVLink)
             The reuse guards ("my peers have finished reading what I am about to overwrite") ride inside those kernels.

  pipelined  (bandwidth regime, e.g. the 100 M-edge graph: 280 MB blocks) -- transfers by the copy engines, hidden under
             compute by ROW CHUNKS of the producer:
             forward : a layer's attention kernel runs over K row chunks; after each chunk the NEXT layer's feature
                       transform of those rows is issued and the chunk is published, so the peers' copy engines pull chunk c
                       while chunk c + 1 is being computed.  The consumer is the unmodified fused attention kernel.
             backward: the column pass writes d Wh for every owner, the owners' copy engines pull their blocks while the row
                       pass runs, a local sum finishes the reduce-scatter.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import functional as Fn
from . import ops
from . import peer as _peer
from .graph import Graph, Hub, default_seg_limit
from .ops import ACT_ELU, ACT_NONE, LRELU_SLOPE, _stream, call, ptr

I32 = torch.int32
PIPELINE_MIN_BLOCK_BYTES = int(os.environ.get("MSHA_PIPELINE_MIN_BYTES", str(16 << 20)))
# row chunks of the pipelined regime; 0 = by exchange kind: 1 (sequential) with the halo exchange -- its volume is small and
# its SM copy kernels interfere with the attention kernels they would overlap (8 GPUs: 46.8 ms overlapped in 4 chunks) -- and 4
# with the whole-block copy-engine exchange
PIPELINE_CHUNKS = int(os.environ.get("MSHA_PIPELINE_CHUNKS", "0"))


def _alias_rows(base: torch.Tensor, lo: int, n: int) -> torch.Tensor:
    """Rows [lo, lo + n) of a persistent 2-D buffer as a tensor that shares its storage but is NOT an autograd view of it:
    an op that writes its result there (``Fn.linear(..., out=)``, mark_dirty) must not rebase the buffer's history --
    the buffer outlives the step and would chain every step's graph."""
    return torch.empty(0, dtype=base.dtype, device=base.device).set_(
        base.untyped_storage(), base.storage_offset() + lo * base.stride(0), (n, base.shape[1]), base.stride())


class Exchange:
    """Buffers and flag channels of one gathered tensor group (1 or 2 tensors of the same rows, e.g. [Wh | s_nbr]).
    ``bufs[k]`` / ``grads[k]``: [world * n_max, widths[k]] on every rank; a rank writes its own block of ``bufs`` (forward)
    and all of ``grads`` (its contributions to everybody's rows, backward)."""

    def __init__(self, pg, part, widths):
        self.widths = tuple(int(w) for w in widths)
        self.bufs = [pg.alloc((part.n_padded, w)) for w in self.widths]
        self.grads = [pg.alloc((part.n_padded, w), zero=True) for w in self.widths]
        (self.ch_ready, self.ch_done, self.ch_gready, self.ch_gdone, self.ch_gready2, self.ch_gdone2) = (
            pg.new_channel() for _ in range(6))
        self.counter = torch.zeros(1, dtype=I32, device=pg.device)
        self.seq = 0              # forward gathers produced so far
        self.gcount = 0           # backward reduce-scatters so far
        self.done_checked = 0     # highest seq whose "all peers pulled it" was verified on the main stream
        self.gdone_checked = 0    # same for the gradient buffers
        self.produced = False     # own blocks of seq are written and published (chunk hooks do that ahead of the consumer)
        self.uses_g2 = False
        self.ev_start = None
        self.s_self = None


class P2P:
    """One rank's state of the peer-memory data path: exchanges by key (created in first-use order, which is the same on
    every rank: the ranks run the same model code), copy stream, staging for pulled gradient blocks."""

    def __init__(self, pg, part, chunks: int = None):
        self.pg, self.part = pg, part
        self.ex = {}
        # high priority: its flag kernels are single CTAs that must get an SM slot while an attention kernel with 10^5 CTAs
        # is being dispatched on the main stream -- at equal priority they wait until that grid has drained (measured: the
        # copy chain then starts milliseconds late and the transfers stop overlapping)
        self.copy_stream = torch.cuda.Stream(device=pg.device, priority=-1)
        want = PIPELINE_CHUNKS if chunks is None else chunks
        self.chunks = want if want >= 1 else (1 if HALO else 4)
        # ranks emulated on ONE GPU share its SM slots: fused kernels whose CTAs spin on another rank's flag must not fill
        # the GPU, or the kernels that would set the flag cannot be scheduled (real ranks have a GPU each: no cap)
        self.max_ctas = 24 if getattr(pg.fabric, "emulated", False) else 0
        self._staging = {}
        self._row_hubs = {}
        self.trace = [] if os.environ.get("MSHA_P2P_TRACE") else None      # (label, event on the current stream)

    def mark(self, label):
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            self.trace.append((label, e))

    def dump_trace(self):
        """-> [(label, ms since the first mark)] of the marks recorded so far; clears them."""
        if not self.trace:
            return []
        torch.cuda.synchronize()
        t0 = self.trace[0][1]
        out = [(lab, round(t0.elapsed_time(e), 3)) for lab, e in self.trace]
        self.trace = []
        return out

    # ---------------------------------------------------------------------------------------------- bookkeeping
    def exchange(self, key, widths) -> Exchange:
        ex = self.ex.get(key)
        if ex is None:
            ex = self.ex[key] = Exchange(self.pg, self.part, widths)
        assert ex.widths == tuple(widths), (key, ex.widths, widths)
        return ex

    def staging(self, C):
        t = self._staging.get(C)
        if t is None:
            t = self._staging[C] = torch.empty((self.part.world, self.part.n_max, C), dtype=torch.float32, device=self.pg.device)
        return t

    def pipelined(self, C) -> bool:
        return self.part.world > 1 and self.part.n_max * C * 4 >= PIPELINE_MIN_BLOCK_BYTES

    def block_rows(self, q) -> slice:
        return slice(q * self.part.n_max, q * self.part.n_max + self.part.sizes[q])

    def chunk_rows(self, q, c):
        """Rows of chunk c (of ``self.chunks``) of rank q's block, as indices into the gathered buffer."""
        n, K = self.part.sizes[q], self.chunks
        return slice(q * self.part.n_max + (n * c) // K, q * self.part.n_max + (n * (c + 1)) // K)

    def local_chunk(self, c):
        n, K = self.part.n_local, self.chunks
        return (n * c) // K, (n * (c + 1)) // K

    def row_chunk_hub(self, graph: Graph, c):
        """Hub-row segments of the rows of local chunk c (row ids relative to the chunk)."""
        key = (id(graph), c)
        h = self._row_hubs.get(key)
        if h is None or h[0] is not graph:
            lo, hi = self.local_chunk(c)
            rp, _ = graph.attention_csr()
            h = self._row_hubs[key] = (graph, Hub(rp[lo:hi + 1], seg_limit=default_seg_limit(graph.nnz)))
        return h[1]

    def own_rows(self, ex, k=0):
        lo = 0 if getattr(ex, "plan", None) is not None else self.part.rank * self.part.n_max     # compact layout: own block first
        return _alias_rows(ex.bufs[k].local, lo, self.part.n_local)

    def halo_plan(self, graph):
        hp = self._row_hubs.get(("halo", id(graph)))
        if hp is None or hp[0] is not graph:
            hp = self._row_hubs[("halo", id(graph))] = (graph, HaloPlan(self, graph))
        return hp[1]

    def halo_exchange(self, key, plan, widths):
        ex = self.ex.get(key)
        if ex is None:
            ex = self.ex[key] = HaloExchange(self.pg, plan, widths)
        assert ex.widths == tuple(widths) and ex.plan is plan, key
        return ex

    def all_rows(self, ex, k=0):
        return _alias_rows(ex.bufs[k].local, 0, self.part.n_padded)

    def chunk_value(self, seq, c):
        """Flag value of "chunk c of gather number seq is published" (monotonic over the life of the exchange)."""
        return (seq - 1) * self.chunks + c + 1

    # ---------------------------------------------------------------------------------------------- reuse guards
    def ensure_fwd_guard(self, ex):
        """Before this rank overwrites its own blocks: every peer has pulled the previous contents."""
        if ex.done_checked < ex.seq:
            self.pg.wait(ex.ch_done, ex.seq)
            ex.done_checked = ex.seq

    def ensure_grad_guard(self, ex):
        if ex.gdone_checked < ex.gcount:
            self.pg.wait(ex.ch_gdone, ex.gcount)
            if ex.uses_g2:
                self.pg.wait(ex.ch_gdone2, ex.gcount)
            ex.gdone_checked = ex.gcount

    def begin_produce(self, ex):
        """Called before the first kernel that writes this rank's own blocks of a new gather."""
        self.ensure_fwd_guard(ex)
        ex.seq += 1
        ex.produced = False
        ex.ev_start = torch.cuda.Event()
        ex.ev_start.record(torch.cuda.current_stream())

    def publish(self, ex, c=None):
        """Own blocks (or their row chunk c) are complete on this stream."""
        last = self.chunks - 1
        self.pg.signal(ex.ch_ready, self.chunk_value(ex.seq, last if c is None else c))
        if c is None or c == last:
            ex.produced = True

    def grad_buffer(self, ex, k=0):
        """The peer-mapped gradient buffer k of an exchange, safe to overwrite (producers write d gathered here)."""
        self.ensure_grad_guard(ex)
        return ex.grads[k].local

    # ---------------------------------------------------------------------------------------------- forward: gather
    def _flag_args(self, wait, guard, done):
        pg = self.pg
        return (pg.flags.local.data_ptr(), pg.flags.tab.data_ptr(), pg.world, pg.rank, wait[0], wait[1] & 0xFFFFFFFF,
                guard[0], guard[1] & 0xFFFFFFFF, done[0], done[1] & 0xFFFFFFFF, _peer.TIMEOUT_NS, pg.status.data_ptr())

    def pull(self, ex):
        """Fetch every peer's blocks of gather ``ex.seq`` into the local buffers (own blocks published already)."""
        pg, part = self.pg, self.part
        W, r, n_max = part.world, part.rank, part.n_max
        nb = len(ex.bufs)
        if not self.pipelined(max(ex.widths)):
            P, L, U = ctypes.c_void_p * nb, ctypes.c_int64 * nb, ctypes.c_uint64 * nb
            call("msha_peer_exchange_pull", nb, P(*[b.local.data_ptr() for b in ex.bufs]), U(*[b.tab.data_ptr() for b in ex.bufs]),
                 L(*[n_max * w * 4 for w in ex.widths]), L(*[n_max * w * 4 for w in ex.widths]),
                 *self._flag_args((ex.ch_ready, self.chunk_value(ex.seq, self.chunks - 1)), (ex.ch_gdone, ex.gcount),
                                  (ex.ch_done, ex.seq)),
                 ex.counter.data_ptr(), self.max_ctas, _stream())
            if not ex.uses_g2:
                ex.gdone_checked = ex.gcount          # the kernel verified the peers' gdone flags
            return
        main = torch.cuda.current_stream()
        cs = self.copy_stream
        cs.wait_event(ex.ev_start)                    # the local copies of the remote blocks are free from here on
        big = [self.pipelined(w) for w in ex.widths]
        with torch.cuda.stream(cs):
            # one flag wait per chunk (the ranks run in lockstep), then that chunk of every peer back to back: a kernel
            # <-> copy-engine hand-over costs ~10 us, per peer and chunk that was 1 ms per exchange
            for c in range(self.chunks):
                pg.wait(ex.ch_ready, self.chunk_value(ex.seq, c))
                for s in range(1, W):
                    q = (r + s) % W
                    rows = self.chunk_rows(q, c)
                    for k, b in enumerate(ex.bufs):
                        if big[k]:
                            pg.pull_block(b, q, rows)
                self.mark(f"cs pulled chunk {c}")
            for k, b in enumerate(ex.bufs):           # the narrow tensors (H scores per node): one SM pull at the end
                if not big[k]:
                    pg.pull_blocks_sm(b, n_max)
            pg.signal(ex.ch_done, ex.seq)
            ev = torch.cuda.Event()
            ev.record(cs)
        main.wait_event(ev)
        self.mark("main: gather complete")

    def gather(self, x_local, key):
        """[n_local, C] -> [world * n_max, C] (autograd: the backward is the reduce-scatter of the gathered gradient)."""
        ex = self.exchange(key, (x_local.shape[1],))
        (out,) = _Gather.apply(self, ex, x_local)
        out._msha_grad_buffer = lambda: self.grad_buffer(ex)
        return out

    # ---------------------------------------------------------------------------------------------- backward: reduce-scatter
    def reduce_scatter(self, ex):
        """``ex.grads`` are complete on this stream -> list of [n_local, w] sums over the ranks of this rank's rows."""
        pg, part = self.pg, self.part
        W, r, n_max = part.world, part.rank, part.n_max
        nb = len(ex.grads)
        ex.gcount += 1
        g = ex.gcount
        outs = [torch.empty((n_max, w), dtype=torch.float32, device=pg.device) for w in ex.widths]
        pg.signal(ex.ch_gready, g)
        big = [self.pipelined(w) for w in ex.widths]
        small = [k for k in range(nb) if not big[k]]
        if any(big):
            main = torch.cuda.current_stream()
            ev0 = torch.cuda.Event()
            ev0.record(main)
            cs = self.copy_stream
            cs.wait_event(ev0)
            own = self.block_rows(r)
            with torch.cuda.stream(cs):
                pg.wait(ex.ch_gready, g)
                for s in range(1, W):
                    q = (r - s) % W
                    for k in range(nb):
                        if big[k]:
                            self.staging(ex.widths[k])[q, :part.n_local].copy_(ex.grads[k].views[q][own], non_blocking=True)
                pg.signal(ex.ch_gdone, g)
                ev = torch.cuda.Event()
                ev.record(cs)
        if small:
            self._sum_small(ex, small, outs, g, ex.ch_gready, ex.ch_gdone2 if any(big) else ex.ch_gdone)
            ex.uses_g2 = any(big)
        if any(big):
            main.wait_event(ev)
            for k in range(nb):
                if big[k]:
                    w = ex.widths[k]
                    stg = self.staging(w)
                    addrs = [ex.grads[k].addr[r] + r * n_max * w * 4 if q == r else stg[q].data_ptr() for q in range(W)]
                    pg.sum_into(outs[k], addrs, part.n_local * w)
        return [o[:part.n_local] for o in outs]

    def _sum_small(self, ex, ks, outs, g, wait_ch, done_ch):
        """One fused kernel: wait for the peers' gradient flags, sum their blocks of this rank's rows in place, verify the
        forward reuse guard, publish completion."""
        part = self.part
        r, n_max = part.rank, part.n_max
        nb = len(ks)
        P, L, U = ctypes.c_void_p * nb, ctypes.c_int64 * nb, ctypes.c_uint64 * nb
        call("msha_peer_exchange_sum", nb, P(*[outs[k].data_ptr() for k in ks]), U(*[ex.grads[k].tab.data_ptr() for k in ks]),
             L(*[r * n_max * ex.widths[k] * 4 for k in ks]), L(*[n_max * ex.widths[k] for k in ks]),
             *self._flag_args((wait_ch, g), (ex.ch_done, ex.seq), (done_ch, g)), ex.counter.data_ptr(), self.max_ctas, _stream())
        ex.done_checked = ex.seq                      # the kernel verified the peers' forward "done" flags


class _Gather(torch.autograd.Function):
    """Gather of 1 or 2 row-partitioned tensors; backward: reduce-scatter of the gathered gradients."""

    @staticmethod
    def forward(ctx, p2p: P2P, ex: Exchange, *xs):
        if not ex.produced:
            fresh = [k for k, x in enumerate(xs) if x.data_ptr() != p2p.own_rows(ex, k).data_ptr()]
            if len(fresh) == len(xs):                  # nobody has started this gather yet
                p2p.begin_produce(ex)
            for k in fresh:
                p2p.own_rows(ex, k).copy_(xs[k])
            p2p.publish(ex)
        p2p.pull(ex)
        ex.produced = False
        ctx.p2p, ctx.ex = p2p, ex
        return tuple(p2p.all_rows(ex, k) for k in range(len(xs)))

    @staticmethod
    def backward(ctx, *gs):
        p2p, ex = ctx.p2p, ctx.ex
        for k, g in enumerate(gs):
            dst = ex.grads[k].local
            if g is None:
                p2p.ensure_grad_guard(ex)
                dst.zero_()
            elif g.data_ptr() != dst.data_ptr():       # otherwise the producer wrote straight into p2p.grad_buffer(ex, k)
                p2p.ensure_grad_guard(ex)
                dst.copy_(g)
        return (None, None, *p2p.reduce_scatter(ex))


class _BufferSink:
    """Hands the attention backward the peer-mapped destinations of d feat_nbr / d s_nbr (functional._AttentionBlock)."""

    def __init__(self, p2p, ex):
        self.p2p, self.ex = p2p, ex

    @property
    def feat_grad(self):
        return self.p2p.grad_buffer(self.ex, 0)

    @property
    def score_grad(self):
        return self.p2p.grad_buffer(self.ex, 1)


class _AttentionRows(torch.autograd.Function):
    """Gather + attention of one GAT layer in the pipelined regime (see the module docstring).  Same arithmetic as
    ``functional._AttentionBlock`` on the gathered tensors (Ours.py:64-69,98 / Ablation.py:262-274)."""

    @staticmethod
    def forward(ctx, Wh_own, s_nbr_own, s_self, p2p: P2P, ex: Exchange, graph, H, D, act, p, seed, chunk_hook, out_buf):
        part = p2p.part
        C = H * D
        dev = Wh_own.device
        assert ex.produced and Wh_own.data_ptr() == p2p.own_rows(ex, 0).data_ptr()
        p2p.mark("fwd: layer start")
        p2p.pull(ex)
        ex.produced = False
        rp, col = graph.attention_csr()
        N, E = graph.n_rows, col.numel()
        s_nbr_g, Wh_g = p2p.all_rows(ex, 1), p2p.all_rows(ex, 0)
        s_self = s_self.contiguous()
        alpha = torch.empty((E, H), dtype=torch.float32, device=dev)
        out = out_buf if out_buf is not None else torch.empty((N, C), dtype=torch.float32, device=dev)
        lib = ops._lib.lib()
        for c in range(p2p.chunks):
            lo, hi = p2p.local_chunk(c)
            if hi > lo:
                hub = p2p.row_chunk_hub(graph, c)
                scr = None
                if hub.n_segs:
                    scr = torch.empty(lib.msha_gat_fwd_hub_scratch_floats(hub.n_segs, H, D), dtype=torch.float32, device=dev)
                call("msha_gat_fwd", rp.data_ptr() + 4 * lo, ptr(col, I32), hi - lo, ptr(s_nbr_g), s_self.data_ptr() + 4 * lo * H,
                     ptr(Wh_g), H, D, LRELU_SLOPE, None, ptr(alpha), out.data_ptr() + 4 * lo * C, act, None, p, seed, hub.ptr,
                     ptr(scr), _stream())
            if chunk_hook is not None:
                chunk_hook(c, lo, hi, out)
            p2p.mark(f"fwd: chunk {c} done")
        ctx.p2p, ctx.ex, ctx.graph = p2p, ex, graph
        ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed = H, D, act, p, seed
        ctx.save_for_backward(s_nbr_g, s_self, Wh_g, alpha, out if act != ACT_NONE else None)
        if out_buf is not None:
            ctx.mark_dirty(out_buf)
        return out

    @staticmethod
    def backward(ctx, d_out):
        s_nbr_g, s_self, Wh_g, alpha, out = ctx.saved_tensors
        p2p, ex, graph = ctx.p2p, ctx.ex, ctx.graph
        H, D, act, p, seed = ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed
        pg, part = p2p.pg, p2p.part
        W, r, n_max = part.world, part.rank, part.n_max
        C = H * D
        rp, col = graph.attention_csr()
        colptr, rowidx, perm = graph.attention_csc()
        N = graph.n_rows
        dev = alpha.device
        main = torch.cuda.current_stream()
        d_out = d_out.contiguous()
        p2p.mark("bwd: layer start")
        if act != ACT_NONE:
            dz = torch.empty_like(d_out)
            call("msha_act_bwd", ptr(d_out), ptr(out), ptr(dz), N * C, act, LRELU_SLOPE, _stream())
        else:
            dz = d_out
        dWh_g = p2p.grad_buffer(ex, 0)                    # guards: the peers have pulled the previous contents
        ds_g = p2p.grad_buffer(ex, 1)
        ex.gcount += 1
        g = ex.gcount
        ex.uses_g2 = True
        ev0 = torch.cuda.Event()
        ev0.record(main)
        # column pass: d Wh_j = sum_i alpha_ij dz_i for the columns of every owner
        call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), part.n_padded, ptr(alpha), ptr(dz), H, D,
             ptr(dWh_g), 0, None, None, p, seed, graph.hub_cols().ptr, _stream())
        pg.signal(ex.ch_gready, g)
        p2p.mark("bwd: column pass done")
        # the owners' copy engines fetch their blocks while the row pass runs
        stg = p2p.staging(C)
        own = p2p.block_rows(r)
        cs = p2p.copy_stream
        cs.wait_event(ev0)
        with torch.cuda.stream(cs):
            pg.wait(ex.ch_gready, g)
            for s in range(1, W):
                q = (r - s) % W
                stg[q, :part.n_local].copy_(ex.grads[0].views[q][own], non_blocking=True)
            pg.signal(ex.ch_gdone, g)
            p2p.mark("cs: grad blocks pulled")
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
        # d s_nbr: column sums of d logit; the peers' blocks of this rank's rows are summed in place (small)
        call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), part.n_padded, None, None, H, D, None, 0,
             ptr(dlogit), ptr(ds_g), p, seed, graph.hub_cols().ptr, _stream())
        pg.signal(ex.ch_gready2, g)
        p2p.mark("bwd: row pass + score sums done")
        ds_nbr = torch.empty((n_max, H), dtype=torch.float32, device=dev)
        p2p._sum_small(ex, [1], [None, ds_nbr], g, ex.ch_gready2, ex.ch_gdone2)
        p2p.mark("bwd: d s_nbr summed")
        main.wait_event(ev_p)
        p2p.mark("bwd: grad blocks here")
        dWh = torch.empty((part.n_local, C), dtype=torch.float32, device=dev)
        addrs = [ex.grads[0].addr[r] + r * n_max * C * 4 if q == r else stg[q].data_ptr() for q in range(W)]
        pg.sum_into(dWh, addrs, part.n_local * C)
        p2p.mark("bwd: d Wh summed")
        return (dWh, ds_nbr[:part.n_local], ds_self) + (None,) * 10


def gat_encode_p2p(convs, x_local, pgraph: Graph, part, p2p: P2P, training=True, score_key=None):
    """``dist.gat_encode`` over the peer-memory data path.  ``score_key``: name of the exchange through which the result
    will be gathered for link scoring (``dist.score_pairs(..., p2p=, key=)``): in the pipelined regime the last layer then
    writes its rows straight into that exchange and publishes them chunk by chunk."""
    h = x_local
    pgraph.attention_csc()                                # host-synchronising one-off builds happen before any flag wait
    pgraph.hub_rows()
    pgraph.hub_cols()
    convs = list(convs)
    chained = None                                        # exchange whose own blocks the previous layer's hook produced
    def layer_exchange(l_, conv_):
        H_, C_ = conv_.heads, conv_.heads * conv_.out_features
        big = p2p.pipelined(C_) and (H_ & (H_ - 1)) == 0 and conv_.out_features % 4 == 0
        if big and HALO and H_ % 4 == 0:
            return p2p.halo_exchange(("gat-halo", l_), p2p.halo_plan(pgraph), (C_, H_)), big
        return p2p.exchange(("gat", l_), (C_, H_)), big

    for l, conv in enumerate(convs):
        H, D = conv.heads, conv.out_features
        C = H * D
        ex, large = layer_exchange(l, conv)
        fuse_elu = conv.activation == "elu" and conv.concat
        act = ACT_ELU if fuse_elu else ACT_NONE
        p = float(conv.dropout) if training else 0.0
        if chained is ex:
            Wh = Fn.linear(h, conv.W, out=p2p.own_rows(ex, 0), precomputed=True)
            s_nbr, s_self = Fn.node_scores(Wh, conv.a_nbr, conv.a_self, H, D, pre=(p2p.own_rows(ex, 1), ex.s_self))
        else:
            p2p.begin_produce(ex)
            Wh = Fn.linear(h, conv.W, out=p2p.own_rows(ex, 0))        # straight into this rank's block
            s_nbr, s_self = Fn.node_scores(Wh, conv.a_nbr, conv.a_self, H, D)
            if large:
                p2p.own_rows(ex, 1).copy_(s_nbr.detach())
                p2p.publish(ex)
        chained = None
        if large:
            hook, out_buf = None, None
            nxt = convs[l + 1] if l + 1 < len(convs) else None
            if nxt is not None and fuse_elu and layer_exchange(l + 1, nxt)[1]:
                nex = layer_exchange(l + 1, nxt)[0]
                hook = _conv_hook(p2p, nex, nxt)
                chained = nex
            elif nxt is None and score_key is not None and conv.concat:
                exh = p2p.exchange(score_key, (C,))
                p2p.begin_produce(exh)
                out_buf = p2p.own_rows(exh, 0)
                hook = (lambda c, lo, hi, out, _e=exh: p2p.publish(_e, c))
            seed = ops.next_seed() if p > 0 else 0
            if isinstance(ex, HaloExchange):
                out = _AttentionHalo.apply(Wh, s_nbr, s_self, p2p, ex, H, D, act, p, seed, hook, out_buf)
            else:
                out = _AttentionRows.apply(Wh, s_nbr, s_self, p2p, ex, pgraph, H, D, act, p, seed, hook, out_buf)
        else:
            Wh_g, s_nbr_g = _Gather.apply(p2p, ex, Wh, s_nbr)
            out, _ = Fn.attention_block(pgraph, s_nbr_g, s_self, Wh_g, heads=H, act=act, dropout_p=conv.dropout,
                                        training=training, grad_sink=_BufferSink(p2p, ex))
        if not conv.concat:
            out = out.view(out.shape[0], H, D).mean(dim=1)
            if conv.activation == "elu":
                out = Fn.elu(out)
        h = out
    return h


def _conv_hook(p2p: P2P, nex: Exchange, conv):
    """After rows [lo, hi) of a layer's output exist: the NEXT layer's feature transform and node scores of those rows, into
    this rank's blocks of the next exchange, then the chunk is published -- the peers' copy engines pull it while the current
    layer's attention kernel works on the following rows.  (The autograd nodes of these two ops are created afterwards by
    ``Fn.linear(..., precomputed=True)`` / ``Fn.node_scores(..., pre=)`` without launching anything.)"""
    H, D = conv.heads, conv.out_features
    C = H * D

    def hook(c, lo, hi, out):
        if c == 0:
            p2p.begin_produce(nex)
            nex.s_self = torch.empty((p2p.part.n_local, H), dtype=torch.float32, device=out.device)
        if hi > lo:
            own0, own1 = p2p.own_rows(nex, 0), p2p.own_rows(nex, 1)
            ops.gemm(out[lo:hi], conv.W.detach(), out=own0[lo:hi])
            call("msha_node_scores", own0.data_ptr() + 4 * lo * C, hi - lo, H, D, ptr(conv.a_nbr.detach().contiguous()),
                 own1.data_ptr() + 4 * lo * H, ptr(conv.a_self.detach().contiguous()), nex.s_self.data_ptr() + 4 * lo * H, _stream())
        p2p.publish(nex, c)
    return hook


# =================================================================================================
# Halo exchange (the north-star's "NVLink halo all-gather of boundary features"): a rank fetches only the feature rows its
# own edges reference.  SURVEY.md section 8e expected the halo of a randomly permuted power-law graph to be "about
# everything"; on the 100 M-edge R-MAT graph it is 38 % of the remote rows at 8 GPUs (a third of the nodes has no in-edge
# at all, the median in-degree is 2) -- and the exchange, not the kernels, bounds the 8-GPU step.
# =================================================================================================
HALO = os.environ.get("MSHA_HALO", "1") != "0"


def halo_need(col_padded: torch.Tensor, world: int, rank: int, n_max: int):
    """Host-side logic of the halo exchange, step 1 (pure index arithmetic; runs on any device): from a rank's column ids in
    the gathered numbering ``owner * n_max + local`` -> (uniq, inv, owner, remote, cnt, need): the sorted distinct ids, the
    position of every edge's column in them, their owners, which of them are remote, how many rows of every peer's block the
    rank references (cnt[rank] = 0), and those rows (ids inside the owner's block, concatenated by owner, ascending)."""
    uniq, inv = torch.unique(col_padded.long(), return_inverse=True)       # sorted padded ids = sorted by (owner, local row)
    owner = torch.div(uniq, n_max, rounding_mode="floor")
    cnt = torch.bincount(owner, minlength=world)
    cnt[rank] = 0
    remote = owner != rank
    need = (uniq[remote] - owner[remote] * n_max).to(I32)
    return uniq, inv, owner, remote, [int(c) for c in cnt.tolist()], need


def halo_offsets(counts, q: int, n_max: int):
    """Step 2: rank q's compact numbering is [own block (n_max rows) | halo of peer 0 | halo of peer 1 | ...], every segment
    starting on a multiple of 4 rows.  ``counts[p]`` = rows q takes from p.  -> (segment start per peer, total rows)."""
    off, o = [], n_max
    for p_ in range(len(counts)):
        off.append(o)
        if p_ != q:
            o += (counts[p_] + 3) // 4 * 4
    return off, o


def halo_compact_ids(uniq, owner, remote, rank: int, n_max: int, hoff, need_ptr):
    """Step 3: compact id of every distinct column: own columns keep their local row, the k-th referenced row of peer p sits
    at ``hoff[p] + k``."""
    dev = uniq.device
    comp = torch.empty_like(uniq)
    comp[~remote] = uniq[~remote] - rank * n_max
    hoff_t = torch.tensor(hoff, dtype=torch.int64, device=dev)
    ptr_t = torch.tensor(need_ptr[:-1], dtype=torch.int64, device=dev)
    ro = owner[remote]
    comp[remote] = hoff_t[ro] + (torch.arange(int(remote.sum()), device=dev) - ptr_t[ro])
    return comp


class HaloPlan:
    """Per rank and graph: which rows of every peer's block this rank's edges reference (``need``), the compact column
    numbering ``[own block | halo of peer 0 | halo of peer 1 | ...]`` with the graph rebuilt in it, the same lists seen from
    the owner's side (``give``: what each peer takes from this rank's block, for the gradient's way back), and the row-chunk
    boundaries of the pipelined pulls.  Built once; two barriers over the fabric exchange the counts and the lists."""

    def __init__(self, p2p: P2P, pgraph: Graph):
        pg, part = p2p.pg, p2p.part
        W, r, n_max = part.world, part.rank, part.n_max
        dev = pgraph.device
        K = p2p.chunks
        rp, col = pgraph.attention_csr()
        uniq, inv, owner, remote, cnt_h, need = halo_need(col, W, r, n_max)
        # ---- counts of every rank (symmetric meta buffer), then the lists
        meta = pg.alloc((W, W), torch.int64)
        lists_sym = pg.alloc((part.n_padded,), I32)
        meta.local[r].copy_(torch.tensor(cnt_h, dtype=torch.int64, device=dev))
        lists_sym.local[: need.numel()].copy_(need)
        pg.barrier()
        all_cnt = torch.stack([meta.views[q][q] for q in range(W)]).cpu().tolist()      # all_cnt[q][p]: rank q needs from p

        def offsets(counts, q):
            return halo_offsets(counts, q, n_max)
        self.hoff, n_compact = offsets(all_cnt[r], r)
        self.n_compact = max(offsets(all_cnt[q], q)[1] for q in range(W))      # same buffer shape on every rank
        need_ptr = [0]
        for q in range(W):
            need_ptr.append(need_ptr[-1] + (cnt_h[q] if q != r else 0))
        self.need, self.need_ptr, self.cnt = need, need_ptr, cnt_h
        # ---- the graph in compact numbering
        comp = halo_compact_ids(uniq, owner, remote, r, n_max, self.hoff, need_ptr)
        self.graph = Graph(pgraph.rowptr, comp[inv].to(I32).contiguous(), pgraph.val, pgraph.n_rows, self.n_compact, isolated="zero")
        self.graph.col_padded = col                                    # the gathered-layout ids (bench.py's parity block)
        # the host-synchronising one-off builds of the compact graph happen here, not lazily in the first backward pass: a
        # backward node whose host thread blocks behind kernels that wait for a peer's flag stalls that rank, and ranks
        # emulated in ONE process share autograd's device thread -- the peer's signalling nodes would then never be enqueued
        self.graph.attention_csc()
        self.graph.hub_rows()
        self.graph.hub_cols()
        # ---- chunk boundaries of the pulls: chunk c of peer q = its rows [(n_q c) // K, (n_q (c+1)) // K)
        i64 = dict(dtype=torch.int64, device=dev)
        need_l = need.long()
        self.chunk_args = []
        for c in range(K):
            beg, end = [0] * W, [0] * W
            for q in range(W):
                if q == r:
                    continue
                seg = need_l[need_ptr[q]:need_ptr[q + 1]]
                lo_, hi_ = (part.sizes[q] * c) // K, (part.sizes[q] * (c + 1)) // K
                b, e = torch.searchsorted(seg, torch.tensor([lo_, hi_], **i64)).tolist()
                beg[q], end[q] = need_ptr[q] + b, need_ptr[q] + e
            self.chunk_args.append((torch.tensor(beg, **i64), torch.tensor(end, **i64),
                                    max(e_ - b_ for b_, e_ in zip(beg, end))))
        self.all_args = (torch.tensor(need_ptr[:-1], **i64), torch.tensor(need_ptr[1:], **i64), max(cnt_h + [0]))
        self.list_first = torch.tensor(need_ptr[:-1], **i64)
        self.local_off = torch.tensor(self.hoff, **i64)
        # ---- the owner's view: what every peer takes from this rank's block, and where it sits in the peer's buffer
        give, give_ptr, remote_off = [], [0], [0] * W
        for q in range(W):
            n_q = all_cnt[q][r] if q != r else 0
            if q != r:
                start = sum(all_cnt[q][p_] for p_ in range(r) if p_ != q)
                give.append(lists_sym.views[q][start:start + n_q].clone())
                remote_off[q] = offsets(all_cnt[q], q)[0][r]
            give_ptr.append(give_ptr[-1] + n_q)
        self.give = torch.cat(give) if give else torch.empty(0, dtype=I32, device=dev)
        self.give_ptr = torch.tensor(give_ptr, **i64)
        self.remote_off = torch.tensor(remote_off, **i64)
        self.max_give = max([give_ptr[q + 1] - give_ptr[q] for q in range(W)] + [0])
        pg.barrier()                                                   # nobody frees / reuses the list buffer under a reader
        torch.cuda.current_stream().synchronize()
        self.halo_rows = sum(cnt_h)
        self.full_rows = sum(part.sizes) - part.n_local


class HaloExchange:
    """Buffers of one gathered tensor group in the compact numbering: ``bufs[k]`` / ``grads[k]`` are [n_compact, w] on every
    rank; rows < n_max are the rank's own block (the only part peers read of ``bufs``), the rest its halo segments (the part
    peers read of ``grads``)."""

    def __init__(self, pg, plan: HaloPlan, widths):
        self.widths = tuple(int(w) for w in widths)
        self.plan = plan
        self.bufs = [pg.alloc((plan.n_compact, w)) for w in self.widths]
        self.grads = [pg.alloc((plan.n_compact, w), zero=True) for w in self.widths]
        (self.ch_ready, self.ch_done, self.ch_gready, self.ch_gdone, self.ch_gready2, self.ch_gdone2) = (
            pg.new_channel() for _ in range(6))
        self.seq = self.gcount = self.done_checked = self.gdone_checked = 0
        self.produced = False
        self.uses_g2 = len(self.widths) > 1
        self.ev_start = None
        self.s_self = None


def _halo_own_rows(p2p, ex, k=0):
    return _alias_rows(ex.bufs[k].local, 0, p2p.part.n_local)


class _AttentionHalo(torch.autograd.Function):
    """``_AttentionRows`` with the halo exchange: the pulls are ``msha_peer_gather_rows`` launches (listed rows of every
    peer, chunk by chunk, on the high-priority side stream), the gradient's way back is ``msha_peer_scatter_add_rows``."""

    @staticmethod
    def forward(ctx, Wh_own, s_nbr_own, s_self, p2p: P2P, ex: HaloExchange, H, D, act, p, seed, chunk_hook, out_buf):
        pg, part, plan = p2p.pg, p2p.part, ex.plan
        graph = plan.graph
        W, r = part.world, part.rank
        C = H * D
        dev = Wh_own.device
        assert ex.produced and Wh_own.data_ptr() == ex.bufs[0].local.data_ptr()
        p2p.mark("fwd: layer start")
        main = torch.cuda.current_stream()
        cs = p2p.copy_stream
        cs.wait_event(ex.ev_start)
        with torch.cuda.stream(cs):
            for c in range(p2p.chunks):
                beg, end, mx = plan.chunk_args[c]
                pg.wait(ex.ch_ready, p2p.chunk_value(ex.seq, c))
                for k, w in enumerate(ex.widths):
                    call("msha_peer_gather_rows", ex.bufs[k].local.data_ptr(), ex.bufs[k].tab.data_ptr(), W, r, ptr(plan.need, I32),
                         beg.data_ptr(), end.data_ptr(), plan.list_first.data_ptr(), plan.local_off.data_ptr(), mx, w,
                         p2p.max_ctas, _stream())
                p2p.mark(f"cs pulled chunk {c}")
            pg.signal(ex.ch_done, ex.seq)
            ev = torch.cuda.Event()
            ev.record(cs)
        main.wait_event(ev)
        p2p.mark("main: gather complete")
        ex.produced = False
        rp, col = graph.attention_csr()
        N, E = graph.n_rows, col.numel()
        Wh_g, s_nbr_g = ex.bufs[0].local, ex.bufs[1].local
        s_self = s_self.contiguous()
        alpha = torch.empty((E, H), dtype=torch.float32, device=dev)
        out = out_buf if out_buf is not None else torch.empty((N, C), dtype=torch.float32, device=dev)
        lib = ops._lib.lib()
        for c in range(p2p.chunks):
            lo, hi = p2p.local_chunk(c)
            if hi > lo:
                hub = p2p.row_chunk_hub(graph, c)
                scr = None
                if hub.n_segs:
                    scr = torch.empty(lib.msha_gat_fwd_hub_scratch_floats(hub.n_segs, H, D), dtype=torch.float32, device=dev)
                call("msha_gat_fwd", rp.data_ptr() + 4 * lo, ptr(col, I32), hi - lo, ptr(s_nbr_g), s_self.data_ptr() + 4 * lo * H,
                     ptr(Wh_g), H, D, LRELU_SLOPE, None, ptr(alpha), out.data_ptr() + 4 * lo * C, act, None, p, seed, hub.ptr,
                     ptr(scr), _stream())
            if chunk_hook is not None:
                chunk_hook(c, lo, hi, out)
            p2p.mark(f"fwd: chunk {c} done")
        ctx.p2p, ctx.ex = p2p, ex
        ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed = H, D, act, p, seed
        ctx.save_for_backward(s_nbr_g, s_self, Wh_g, alpha, out if act != ACT_NONE else None)
        if out_buf is not None:
            ctx.mark_dirty(out_buf)
        return out

    @staticmethod
    def backward(ctx, d_out):
        s_nbr_g, s_self, Wh_g, alpha, out = ctx.saved_tensors
        p2p, ex = ctx.p2p, ctx.ex
        plan = ex.plan
        graph = plan.graph
        H, D, act, p, seed = ctx.H, ctx.D, ctx.act, ctx.p, ctx.seed
        pg, part = p2p.pg, p2p.part
        W, r = part.world, part.rank
        C = H * D
        rp, col = graph.attention_csr()
        colptr, rowidx, perm = graph.attention_csc()
        N, M = graph.n_rows, graph.n_cols
        dev = alpha.device
        main = torch.cuda.current_stream()
        d_out = d_out.contiguous()
        p2p.mark("bwd: layer start")
        if act != ACT_NONE:
            dz = torch.empty_like(d_out)
            call("msha_act_bwd", ptr(d_out), ptr(out), ptr(dz), N * C, act, LRELU_SLOPE, _stream())
        else:
            dz = d_out
        dWh_g = p2p.grad_buffer(ex, 0)
        ds_g = p2p.grad_buffer(ex, 1)
        ex.gcount += 1
        g = ex.gcount
        E = alpha.shape[0]
        dlogit = torch.empty_like(alpha)
        ds_self = torch.empty((N, H), dtype=torch.float32, device=dev)
        hub = graph.hub_rows()
        r_buf = torch.empty((N, H), dtype=torch.float32, device=dev) if hub.n_segs else None

        def row_pass():
            call("msha_gat_bwd_rows", ptr(rp, I32), ptr(col, I32), N, ptr(s_nbr_g), ptr(s_self), LRELU_SLOPE, ptr(alpha),
                 ptr(Wh_g), ptr(dz), None, ACT_NONE, None, None, None, None, None, H, D, ptr(dlogit), ptr(ds_self), p, seed,
                 hub.ptr, ptr(r_buf), int(E // max(N, 1)), _stream())

        def add_rows(k, w):
            call("msha_peer_scatter_add_rows", ex.grads[k].local.data_ptr(), ex.grads[k].tab.data_ptr(), W, r, ptr(plan.give, I32),
                 plan.give_ptr.data_ptr(), plan.remote_off.data_ptr(), plan.max_give, w, p2p.max_ctas, _stream())

        if p2p.chunks == 1:
            # sequential: the halo is small (38 % of the remote rows at 8 GPUs), so the overlapped order below -- which has to
            # split the column pass in two and lets an SM copy kernel run beside the row pass -- costs more than it hides
            row_pass()
            call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, ptr(alpha), ptr(dz), H, D, ptr(dWh_g), 0,
                 ptr(dlogit), ptr(ds_g), p, seed, graph.hub_cols().ptr, _stream())
            pg.signal(ex.ch_gready, g)
            p2p.mark("bwd: row + column pass done")
            pg.wait(ex.ch_gready, g)
            add_rows(0, C)
            add_rows(1, H)
            pg.signal(ex.ch_gdone, g)
            pg.signal(ex.ch_gdone2, g)
            pg.wait(ex.ch_done, ex.seq)                       # forward reuse guard of this exchange
            ex.done_checked = ex.seq
            p2p.mark("bwd: grad rows added")
        else:
            ev0 = torch.cuda.Event()
            ev0.record(main)
            call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, ptr(alpha), ptr(dz), H, D, ptr(dWh_g), 0,
                 None, None, p, seed, graph.hub_cols().ptr, _stream())
            pg.signal(ex.ch_gready, g)
            p2p.mark("bwd: column pass done")
            ev1 = torch.cuda.Event()
            ev1.record(main)
            cs = p2p.copy_stream
            cs.wait_event(ev1)                                # this rank's own block of d Wh is complete: peers' rows add onto it
            with torch.cuda.stream(cs):
                pg.wait(ex.ch_gready, g)
                add_rows(0, C)
                pg.signal(ex.ch_gdone, g)
                p2p.mark("cs: grad rows added")
                ev_p = torch.cuda.Event()
                ev_p.record(cs)
            row_pass()
            call("msha_spmm_csc", ptr(colptr, I32), ptr(rowidx, I32), ptr(perm, I32), M, None, None, H, D, None, 0,
                 ptr(dlogit), ptr(ds_g), p, seed, graph.hub_cols().ptr, _stream())
            pg.signal(ex.ch_gready2, g)
            p2p.mark("bwd: row pass + score sums done")
            pg.wait(ex.ch_gready2, g)
            add_rows(1, H)
            pg.signal(ex.ch_gdone2, g)
            pg.wait(ex.ch_done, ex.seq)                       # forward reuse guard of this exchange
            ex.done_checked = ex.seq
            p2p.mark("bwd: d s_nbr summed")
            main.wait_event(ev_p)
            p2p.mark("bwd: grad rows here")
        # the sums live in this rank's own block of the gradient buffers; hand out copies (the buffers are reused next step)
        dWh = dWh_g[:part.n_local].clone()
        ds_nbr = ds_g[:part.n_local].clone()
        return (dWh, ds_nbr, ds_self) + (None,) * 9


def _halo_gather_all(p2p: P2P, ex: HaloExchange, k=0):
    """Own block of ``ex.bufs[k]`` is published: fetch every listed row of every peer into the halo segments."""
    pg, part, plan = p2p.pg, p2p.part, ex.plan
    beg, end, mx = plan.all_args
    pg.wait(ex.ch_ready, p2p.chunk_value(ex.seq, p2p.chunks - 1))
    call("msha_peer_gather_rows", ex.bufs[k].local.data_ptr(), ex.bufs[k].tab.data_ptr(), part.world, part.rank, ptr(plan.need, I32),
         beg.data_ptr(), end.data_ptr(), plan.list_first.data_ptr(), plan.local_off.data_ptr(), mx, ex.widths[k], p2p.max_ctas,
         _stream())
    pg.signal(ex.ch_done, ex.seq)


def _halo_add_all(p2p: P2P, ex: HaloExchange, k=0):
    """``ex.grads[k]`` (compact numbering) is complete on every rank's stream once its flag is seen: add the peers' halo
    segments onto the listed rows of this rank's own block."""
    pg, part, plan = p2p.pg, p2p.part, ex.plan
    ex.gcount += 1
    g = ex.gcount
    pg.signal(ex.ch_gready, g)
    pg.wait(ex.ch_gready, g)
    call("msha_peer_scatter_add_rows", ex.grads[k].local.data_ptr(), ex.grads[k].tab.data_ptr(), part.world, part.rank,
         ptr(plan.give, I32), plan.give_ptr.data_ptr(), plan.remote_off.data_ptr(), plan.max_give, ex.widths[k], p2p.max_ctas,
         _stream())
    pg.signal(ex.ch_gdone, g)


class _HaloGather(torch.autograd.Function):
    """[n_local, w] -> [n_compact, w]: own rows + the halo rows this rank's edges reference (compact numbering of the
    plan's graph); backward: the adjoint -- the peers' halo-segment gradients are added onto the owners' rows."""

    @staticmethod
    def forward(ctx, x, p2p: P2P, ex: HaloExchange):
        own = _alias_rows(ex.bufs[0].local, 0, p2p.part.n_local)
        p2p.begin_produce(ex)
        if x.data_ptr() != own.data_ptr():
            own.copy_(x)
        p2p.publish(ex)
        _halo_gather_all(p2p, ex)
        ex.produced = False
        ctx.p2p, ctx.ex = p2p, ex
        return _alias_rows(ex.bufs[0].local, 0, ex.plan.n_compact)

    @staticmethod
    def backward(ctx, g):
        p2p, ex = ctx.p2p, ctx.ex
        dst = ex.grads[0].local
        if g.data_ptr() != dst.data_ptr():
            p2p.ensure_grad_guard(ex)
            dst.copy_(g)
        _halo_add_all(p2p, ex)
        return dst[:p2p.part.n_local].clone(), None, None


class _HaloScatterAdd(torch.autograd.Function):
    """[n_compact, w] per-rank contributions in the compact numbering -> [n_local, w] sums over the ranks of the own rows
    (the reduce-scatter of ``alpha.T @ h2``, Ours.py:100, restricted to referenced recipients); backward: the halo gather."""

    @staticmethod
    def forward(ctx, x, p2p: P2P, ex: HaloExchange):
        dst = ex.grads[0].local
        if x.data_ptr() != dst.data_ptr():
            p2p.ensure_grad_guard(ex)
            dst.copy_(x)
        _halo_add_all(p2p, ex)
        ctx.p2p, ctx.ex = p2p, ex
        return dst[:p2p.part.n_local].clone()

    @staticmethod
    def backward(ctx, g):
        p2p, ex = ctx.p2p, ctx.ex
        own = _alias_rows(ex.bufs[0].local, 0, p2p.part.n_local)
        p2p.begin_produce(ex)
        own.copy_(g)
        p2p.publish(ex)
        _halo_gather_all(p2p, ex)
        ex.produced = False
        return _alias_rows(ex.bufs[0].local, 0, ex.plan.n_compact), None, None


def halo_gather(p2p: P2P, plan: HaloPlan, x_local, key):
    return _HaloGather.apply(x_local, p2p, p2p.halo_exchange(key, plan, (x_local.shape[1],)))


def halo_scatter_add(p2p: P2P, plan: HaloPlan, x_compact, key):
    return _HaloScatterAdd.apply(x_compact, p2p, p2p.halo_exchange(key, plan, (x_compact.shape[1],)))
