"""Peer-mapped exchange buffers and flags for the node-partitioned path.

One process per GPU; every rank allocates its exchange buffers in memory that all ranks of the node have mapped
(NVLink 5 / NVSwitch peer access; PyTorch's symmetric-memory allocator does the handle exchange -- plumbing) and the
data path is this library's own: ``msha_peer_signal`` / ``msha_peer_wait`` flags, copy-engine pulls of whole owner
blocks, ``msha_peer_pull_*`` / ``msha_peer_sum`` kernels (csrc/peer_kernels.cu).  The reference has no distributed code
(train.py:18 is single-device).

Two fabrics with one interface:
  * ``SymmFabric(group)``    -- real ranks: ``torch.distributed._symmetric_memory`` allocation + rendezvous.
  * ``LocalFabric(world)``   -- all ranks emulated inside ONE process on ONE GPU (a rank = a ``PeerGroup`` view + its own
                               CUDA stream); peers' buffers are ordinary device tensors.  The protocol (flags, pulls,
                               sums) is exactly the one real ranks run, which is what the single-GPU tests exercise.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .ops import call

MAX_CHANNELS = 512
TIMEOUT_NS = int(float(os.environ.get("MSHA_PEER_TIMEOUT_S", "60")) * 1e9)
# blocks up to this many bytes are pulled by one SM kernel (P2P loads, one launch for all peers); larger ones by the
# copy engines (no SM time, one cudaMemcpyAsync per peer)
CE_MIN_BYTES = int(os.environ.get("MSHA_PEER_CE_MIN_BYTES", str(4 << 20)))


def require_eager_module_loading():
    """Kernels of this path spin on flags that OTHER kernels set.  With CUDA's default lazy module loading the first
    launch of a kernel function loads it, which can need the context to drain -- behind a spinning wait kernel that is a
    deadlock (CUDA programming guide, "lazy loading": programs relying on concurrent kernels must preload).  The loading
    mode is read when the CUDA context is created, so it has to be in the environment before the first CUDA call."""
    if os.environ.get("CUDA_MODULE_LOADING", "").upper() != "EAGER" and os.environ.get("MSHA_PEER_ALLOW_LAZY") != "1":
        raise RuntimeError("msha_b200 peer path: set CUDA_MODULE_LOADING=EAGER in the environment before CUDA is "
                           "initialised (kernels that wait on one another must not be loaded lazily)")


class PeerTensor:
    """A buffer every rank holds with the same shape; ``views[q]`` is rank q's copy as addressable from this rank."""

    def __init__(self, views, rank):
        self.views, self.rank = views, rank
        self.local = views[rank]
        self.addr = [int(v.data_ptr()) for v in views]
        self.tab = torch.tensor(self.addr, dtype=torch.int64, device=self.local.device)      # device table for kernels


class PeerGroup:
    """One rank's handle on a fabric: allocation, flags, pulls.  All calls are enqueued on torch's current stream."""

    def __init__(self, fabric, rank: int, world: int, device):
        self.fabric, self.rank, self.world, self.device = fabric, int(rank), int(world), device
        self._n_alloc = 0
        self._n_chan = 1                                   # channel 0: barrier
        self._barrier_seq = 0
        self.all_mask = (1 << world) - 1
        self.others_mask = self.all_mask & ~(1 << rank)
        self.flags = self.alloc((MAX_CHANNELS * world,), torch.int32, zero=True)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.fabric.sync_after_setup(self)

    # -- allocation ------------------------------------------------------------------------------------------
    def alloc(self, shape, dtype=torch.float32, zero=False) -> PeerTensor:
        pt = self.fabric.alloc(self, self._n_alloc, tuple(int(s) for s in shape), dtype, zero)
        self._n_alloc += 1
        return pt

    def new_channel(self) -> int:
        """Channel ids are handed out in call order: every rank must create its exchanges in the same order."""
        c = self._n_chan
        self._n_chan += 1
        if c >= MAX_CHANNELS:
            raise RuntimeError("peer: out of flag channels")
        return c

    # -- flags -----------------------------------------------------------------------------------------------
    def signal(self, channel: int, value: int, mask=None):
        mask = self.others_mask if mask is None else mask
        if not mask:
            return
        call("msha_peer_signal", self.flags.tab.data_ptr(), self.world, self.rank, channel, mask, None,
             value & 0xFFFFFFFF, ops._stream())

    def wait(self, channel: int, value: int, mask=None):
        mask = self.others_mask if mask is None else mask
        if not mask:
            return
        call("msha_peer_wait", self.flags.local.data_ptr(), self.world, channel, mask, None, value & 0xFFFFFFFF,
             TIMEOUT_NS, self.status.data_ptr(), ops._stream())

    def check(self):
        """Raise if a flag wait of this rank ever timed out (reads one word: synchronises the host with the device)."""
        st = int(self.status.item())
        if st:
            raise RuntimeError(f"msha_b200 peer path: rank {self.rank} gave up waiting for rank {st & 0xFF} "
                               f"(no flag within {TIMEOUT_NS / 1e9:.0f} s)")

    def barrier(self):
        """Device-side barrier over the fabric (stream-ordered; the host does not block)."""
        self._barrier_seq += 1
        self.signal(0, self._barrier_seq)
        self.wait(0, self._barrier_seq)

    # -- data ------------------------------------------------------------------------------------------------
    def pull_block(self, pt: PeerTensor, q: int, rows: slice):
        """Copy-engine pull: rows ``rows`` of rank q's buffer -> the same rows of the local buffer."""
        pt.local[rows].copy_(pt.views[q][rows], non_blocking=True)

    def pull_blocks_sm(self, pt: PeerTensor, block_rows: int, rows_used: int = None, max_ctas: int = 0):
        """One kernel: every remote rank's block [q * block_rows, q * block_rows + rows_used) out of rank q's buffer."""
        row_bytes = pt.local[0].numel() * pt.local.element_size() if pt.local.dim() > 1 else pt.local.element_size()
        used = block_rows if rows_used is None else rows_used
        nbytes = (used * row_bytes + 15) // 16 * 16
        nbytes = min(nbytes, block_rows * row_bytes)
        call("msha_peer_pull_blocks", pt.local.data_ptr(), pt.tab.data_ptr(), self.world, self.rank, block_rows * row_bytes,
             nbytes, max_ctas, ops._stream())

    def sum_into(self, out: torch.Tensor, addrs, n: int, max_ctas: int = 0):
        """out[:n] = sum of the float arrays at the device addresses ``addrs`` (table order)."""
        import ctypes
        arr = (ctypes.c_uint64 * len(addrs))(*[int(a) for a in addrs])
        call("msha_peer_sum", out.data_ptr(), arr, len(addrs), int(n), max_ctas, ops._stream())


class LocalFabric:
    """All ranks in one process on one device (tests, single-GPU emulation)."""

    def __init__(self, world: int, device):
        require_eager_module_loading()
        import threading
        self.world, self.device = int(world), device
        self._allocs = []
        self._lock = threading.Lock()                      # the emulated ranks may run in one host thread each
        self._alloc_stream = torch.cuda.Stream(device=device)
        self.emulated = True
        self.groups = [None] * world
        self.streams = [torch.cuda.Stream(device=device) for _ in range(world)]
        for r in range(world):
            self.groups[r] = PeerGroup(self, r, world, device)

    def alloc(self, group, index, shape, dtype, zero):
        with self._lock:
            if index == len(self._allocs):
                # on a stream of its own: the calling rank's stream may hold kernels that wait for flags of ranks which are
                # themselves queueing for this lock -- synchronising THAT stream here would deadlock the emulation
                with torch.cuda.stream(self._alloc_stream):
                    mk = torch.zeros if zero else torch.empty
                    self._allocs.append([mk(shape, dtype=dtype, device=self.device) for _ in range(self.world)])
                self._alloc_stream.synchronize()               # zero-fill done before any rank's stream touches it
            views = self._allocs[index]
        if tuple(views[0].shape) != shape or views[0].dtype != dtype:
            raise RuntimeError("LocalFabric: ranks must allocate the same buffers in the same order")
        return PeerTensor(views, group.rank)

    def sync_after_setup(self, group):
        pass


class SymmFabric:
    """Real ranks over torch.distributed: symmetric-memory allocation, rendezvous over the process group."""

    def __init__(self, group=None, device=None):
        require_eager_module_loading()
        import torch.distributed as dist
        self.dist = dist
        self.emulated = False
        self.pg = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.pg)
        self.rank = dist.get_rank(self.pg)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._handles = []
        self.group = PeerGroup(self, self.rank, self.world, self.device)

    def alloc(self, group, index, shape, dtype, zero):
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(shape, dtype=dtype, device=self.device)
        if zero:
            t.zero_()
        hdl = symm.rendezvous(t, self.pg.group_name)
        views = [t if q == self.rank else hdl.get_buffer(q, shape, dtype) for q in range(self.world)]
        self._handles.append((t, hdl))
        return PeerTensor(views, self.rank)

    def sync_after_setup(self, group):
        # every rank's flags are zeroed before anybody signals
        torch.cuda.synchronize(self.device)
        self.dist.barrier(self.pg)
