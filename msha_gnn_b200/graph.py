"""Device-resident graph structure: canonical CSR (bit-exact with the reference's dense-adjacency
neighbour sets, row-major ``(adj > 0).nonzero()``), the attention CSR (isolated rows carry M masked
edges, GAT.py:29-31) and its CSC transpose with the CSR-slot permutation.

Replaces the dense ``(N, M)`` float adjacency consumed by every reference layer (GAT.py:20,30;
Ours.py:54,67; HGANE.py:38-39) and the O(N^2) Python builders ``dataset.py:260-296``.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import torch

from . import ops
from .ops import call, ptr, workspace, _stream

_L = ops._lib


# rows / columns with more entries are processed as segments of this many slots (hub handling); the default adapts to
# the graph so that small graphs with a few very long rows / columns still spread over the whole GPU
SEG_LIMIT = int(os.environ.get("MSHA_SEG_LIMIT", "0"))
SEG_LIMIT_MAX, SEG_LIMIT_MIN = 1024, 64


def default_seg_limit(nnz: int) -> int:
    """Segment length: ~16 warps of work per SM if the whole structure were hubs, a power of two in [64, 1024]."""
    if SEG_LIMIT:
        return SEG_LIMIT
    target = max(1, nnz // (148 * 16))
    lim = SEG_LIMIT_MIN
    while lim < target and lim < SEG_LIMIT_MAX:
        lim *= 2
    return lim


class _HubStruct(ctypes.Structure):
    """Mirror of ``msha_hub_t`` (include/msha_b200.h)."""
    _fields_ = [("seg_limit", ctypes.c_int32), ("n_segs", ctypes.c_int32), ("seg_item", ctypes.c_void_p),
                ("seg_beg", ctypes.c_void_p), ("seg_end", ctypes.c_void_p), ("n_hub", ctypes.c_int32),
                ("pad_", ctypes.c_int32), ("hub_ids", ctypes.c_void_p), ("hub_seg_ptr", ctypes.c_void_p)]


class Hub:
    """Segment decomposition of the rows (columns) with more than ``seg_limit`` entries; ``ptr`` is 0 when none."""

    def __init__(self, ptr_arr: torch.Tensor = None, seg_limit: int = None, beg: torch.Tensor = None,
                 end: torch.Tensor = None):
        """``ptr_arr``: a CSR / CSC pointer array (item i owns slots [ptr[i], ptr[i+1])); or explicit per-item slot
        ranges ``beg`` / ``end`` (the owner-block sub-ranges of the partitioned forward, dist.py)."""
        if ptr_arr is None:
            ptr_arr_b, ptr_arr_e = beg, end
            if seg_limit is None:
                seg_limit = default_seg_limit(int((end.long() - beg.long()).sum().item()) if beg.numel() else 0)
        else:
            ptr_arr_b, ptr_arr_e = ptr_arr[:-1], ptr_arr[1:]
            seg_limit = default_seg_limit(int(ptr_arr[-1])) if seg_limit is None else seg_limit
        deg = (ptr_arr_e - ptr_arr_b).long()
        ids = torch.nonzero(deg > seg_limit).flatten()
        ptr_arr = ptr_arr_b
        self.n_hub = int(ids.numel())
        self.n_segs = 0
        self.ptr = None
        if self.n_hub == 0:
            return
        nseg = (deg[ids] + seg_limit - 1) // seg_limit
        seg_ptr = torch.zeros(self.n_hub + 1, dtype=torch.int64, device=ptr_arr.device)
        seg_ptr[1:] = torch.cumsum(nseg, 0)
        self.n_segs = int(seg_ptr[-1].item())
        which = torch.repeat_interleave(torch.arange(self.n_hub, device=ptr_arr.device), nseg)
        local = torch.arange(self.n_segs, device=ptr_arr.device) - seg_ptr[:-1][which]
        item = ids[which]
        beg = ptr_arr.long()[item] + local * seg_limit
        end = torch.minimum(beg + seg_limit, ptr_arr_e.long()[item])
        self.tensors = [t.to(torch.int32).contiguous() for t in (item, beg, end, ids, seg_ptr)]
        self.struct = _HubStruct(seg_limit, self.n_segs, self.tensors[0].data_ptr(), self.tensors[1].data_ptr(),
                                 self.tensors[2].data_ptr(), self.n_hub, 0, self.tensors[3].data_ptr(),
                                 self.tensors[4].data_ptr())
        self.ptr = ctypes.addressof(self.struct)


class Graph:
    """CSR/CSC of a bipartite (rows = sources, cols = recipients) or square adjacency."""

    def __init__(self, rowptr, col, val, n_rows, n_cols, isolated="reference"):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.rowptr, self.col, self.val = rowptr, col, val          # int32 [N+1], int32 [nnz], fp32 [nnz]
        self.nnz = int(col.numel())
        self.device = rowptr.device
        self._att = None
        self._csc = None
        self._csc_plain = None
        self._deg = None
        self._hubs = {}
        self.isolated = isolated

    # ---------------------------------------------------------------- constructors
    @staticmethod
    def from_dense(adj: torch.Tensor) -> "Graph":
        """Neighbour set of row i == {j : adj[i,j] > 0} (GAT.py:30); values kept for GraphConvolution."""
        if adj.dim() != 2:
            raise ValueError("adjacency must be 2-D")
        if not adj.is_cuda:
            raise RuntimeError("msha_b200 is CUDA-only: move the adjacency to the GPU (train.py:213 does)")
        a = adj.detach()
        if a.dtype != torch.float32:
            a = a.float()
        if a.stride(1) != 1 or (a.shape[0] > 1 and a.stride(0) < a.shape[1]):
            a = a.contiguous()
        N, M = a.shape
        ld = a.stride(0) if N > 1 else M
        lib = _L.lib()
        rowptr = torch.empty(N + 1, dtype=torch.int32, device=a.device)
        ws = workspace(lib.msha_csr_from_dense_workspace_bytes(N), a.device)
        call("msha_csr_from_dense_rowptr", a.data_ptr(), N, M, ld, ptr(rowptr, torch.int32), ws.data_ptr(), ws.numel(), _stream())
        nnz = int(rowptr[-1].item())                       # the one host read of graph construction
        col = torch.empty(nnz, dtype=torch.int32, device=a.device)
        val = torch.empty(nnz, dtype=torch.float32, device=a.device)
        call("msha_csr_from_dense_fill", a.data_ptr(), N, M, ld, ptr(rowptr, torch.int32), ptr(col, torch.int32), ptr(val), _stream())
        return Graph(rowptr, col, val, N, M)

    @staticmethod
    def from_coo(src: torch.Tensor, dst: torch.Tensor, n_rows: int, n_cols: int) -> "Graph":
        """Flow records -> coalesced CSR; multiplicities become values (dataset.py:286-288)."""
        if not src.is_cuda or not dst.is_cuda:
            raise RuntimeError("msha_b200 is CUDA-only: edge endpoints must be CUDA tensors")
        src = src.detach().to(torch.int64).contiguous()
        dst = dst.detach().to(torch.int64).contiguous()
        if src.shape != dst.shape or src.dim() != 1:
            raise ValueError("src/dst must be 1-D tensors of equal length")
        n = src.numel()
        dev = src.device
        lib = _L.lib()
        rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        val = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = workspace(lib.msha_csr_from_coo_workspace_bytes(n, n_rows, n_cols), dev)
        call("msha_csr_from_coo", ptr(src, torch.int64), ptr(dst, torch.int64), n, n_rows, n_cols, ptr(rowptr, torch.int32),
             ptr(col, torch.int32), ptr(val), ptr(status, torch.int32), ws.data_ptr(), ws.numel(), _stream())
        st, nnz = torch.stack([status[0], rowptr[-1]]).tolist()
        if st & 1:
            raise IndexError("edge endpoint out of range")
        return Graph(rowptr, col[:nnz].clone() if nnz < n else col[:nnz], val[:nnz].clone() if nnz < n else val[:nnz],
                     n_rows, n_cols)

    @staticmethod
    def from_edge_index(edge_index: torch.Tensor, n_rows: int, n_cols: int | None = None) -> "Graph":
        """``edge_index`` int64 (2, E): row 0 = destination/row node i, row 1 = neighbour/column node j."""
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must have shape (2, E)")
        return Graph.from_coo(edge_index[0], edge_index[1], n_rows, n_rows if n_cols is None else n_cols)

    # ---------------------------------------------------------------- derived structures
    @property
    def degrees(self):
        if self._deg is None:
            self._deg = self.rowptr[1:] - self.rowptr[:-1]
        return self._deg

    def attention_csr(self):
        """(rowptr, col) the masked softmax runs over: rows without neighbours get all M columns as masked
        edges (col = ~j).  Identical to the canonical CSR when no row is isolated."""
        if self._att is None:
            n_iso = int((self.degrees == 0).sum().item()) if self.n_rows else 0
            if n_iso == 0 or self.isolated == "zero":
                self._att = (self.rowptr, self.col, 0)
            else:
                if n_iso * self.n_cols > (1 << 28):
                    raise RuntimeError(
                        f"{n_iso} isolated rows x {self.n_cols} columns: the reference's uniform 1/M attention over all "
                        "columns is not materialisable; add self loops or build the Graph with isolated='zero'")
                lib = _L.lib()
                dev = self.device
                rp = torch.empty(self.n_rows + 1, dtype=torch.int32, device=dev)
                ws = workspace(lib.msha_csr_from_dense_workspace_bytes(self.n_rows), dev)
                call("msha_csr_augment_rowptr", ptr(self.rowptr, torch.int32), self.n_rows, self.n_cols, ptr(rp, torch.int32),
                     ws.data_ptr(), ws.numel(), _stream())
                nnz = self.nnz + n_iso * self.n_cols
                c = torch.empty(nnz, dtype=torch.int32, device=dev)
                call("msha_csr_augment_fill", ptr(self.rowptr, torch.int32), ptr(self.col, torch.int32), self.n_rows,
                     self.n_cols, ptr(rp, torch.int32), ptr(c, torch.int32), _stream())
                self._att = (rp, c, n_iso)
        return self._att[0], self._att[1]

    @property
    def n_isolated(self):
        self.attention_csr()
        return self._att[2]

    def attention_csc(self):
        """(colptr, rowidx, perm) of the attention CSR; perm maps a CSC slot to its CSR slot."""
        if self._csc is None:
            rp, c = self.attention_csr()
            self._csc = _csc(rp, c, self.n_rows, self.n_cols)
        return self._csc

    def hub_rows(self) -> Hub:
        """Segments of the attention-CSR rows with more than SEG_LIMIT neighbours (power-law hubs)."""
        if "rows" not in self._hubs:
            self._hubs["rows"] = Hub(self.attention_csr()[0])
        return self._hubs["rows"]

    def hub_cols(self) -> Hub:
        if "cols" not in self._hubs:
            self._hubs["cols"] = Hub(self.attention_csc()[0])
        return self._hubs["cols"]

    def hub_rows_plain(self) -> Hub:
        if "rows_plain" not in self._hubs:
            self._hubs["rows_plain"] = Hub(self.rowptr)
        return self._hubs["rows_plain"]

    def hub_cols_plain(self) -> Hub:
        if "cols_plain" not in self._hubs:
            self._hubs["cols_plain"] = Hub(self.transpose_structure()[0])
        return self._hubs["cols_plain"]

    def transpose_structure(self):
        """CSC of the canonical CSR (no masked edges)."""
        if self._csc_plain is None:
            self._csc_plain = _csc(self.rowptr, self.col, self.n_rows, self.n_cols)
        return self._csc_plain

    def normalized_values(self):
        """Column-normalised values ``A[:, j] / colsum[j]`` (model.py:95-100) in CSR order."""
        out = torch.empty_like(self.val)
        colsum = torch.empty(self.n_cols, dtype=torch.float32, device=self.device)
        call("msha_csr_normalize_columns", ptr(self.col, torch.int32), ptr(self.val), self.nnz, self.n_cols, ptr(colsum),
             ptr(out), _stream())
        return out

    def edge_index(self):
        """(2, nnz) int64 in canonical order -- equals ``(adj > 0).nonzero().T``."""
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device), self.degrees.long())
        return torch.stack([rows, self.col.long()])


def _csc(rowptr, col, n_rows, n_cols):
    lib = _L.lib()
    dev = rowptr.device
    nnz = int(col.numel())
    colptr = torch.empty(n_cols + 1, dtype=torch.int32, device=dev)
    rowidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    perm = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    ws = workspace(lib.msha_csc_from_csr_workspace_bytes(nnz, n_cols), dev)
    call("msha_csc_from_csr", ptr(rowptr, torch.int32), ptr(col, torch.int32), n_rows, n_cols, nnz, ptr(colptr, torch.int32),
         ptr(rowidx, torch.int32), ptr(perm, torch.int32), ws.data_ptr(), ws.numel(), _stream())
    return colptr, rowidx[:nnz], perm[:nnz]


# ---------------------------------------------------------------------------------------------
# adjacency argument dispatch + cache (train.py:227 passes the same dense tensor every step)
# ---------------------------------------------------------------------------------------------
_cache: dict = {}


def as_graph(adj, n_rows=None, n_cols=None) -> Graph:
    """Accepts a Graph, a dense float (N, M) adjacency (reference behaviour) or an int64 (2, E) edge_index."""
    if isinstance(adj, Graph):
        return adj
    if not isinstance(adj, torch.Tensor):
        raise TypeError("adjacency must be a torch.Tensor or a msha_gnn_b200.Graph")
    if not adj.is_cuda:
        raise RuntimeError("msha_b200 is CUDA-only: the adjacency tensor lives on the CPU (no CPU fallback)")
    key = (adj.data_ptr(), adj._version, tuple(adj.shape), adj.dtype, n_rows, n_cols)
    hit = _cache.get(key)
    if hit is not None and hit[0]() is adj:
        return hit[1]
    if adj.dtype in (torch.int64, torch.int32) and adj.dim() == 2 and adj.shape[0] == 2:
        if n_rows is None:
            raise ValueError("edge_index adjacency needs n_rows")
        g = Graph.from_edge_index(adj.long(), n_rows, n_cols)
    elif adj.is_floating_point() and adj.dim() == 2:
        g = Graph.from_dense(adj)
    else:
        raise TypeError("adjacency must be a float (N, M) matrix or an int64 (2, E) edge_index")
    if len(_cache) > 64:
        _cache.clear()
    _cache[key] = (weakref.ref(adj), g)
    return g
