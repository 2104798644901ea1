"""Loader for the reference's on-disk dataset (``dataset.HigherDataset``, dataset.py:208-296) that feeds the
device graph build directly -- SURVEY.md section 8(f) row 2.

The reference loader builds three dense matrices on the host with Python loops: ``inter_adjacent`` (N, M) one
``+= 1`` per flow record (dataset.py:279-296) and ``intra_adjacent`` two (N, N) group-equality matrices by an
O(N^2) double loop (dataset.py:260-277; 39 179^2 floats = 6.1 GB each for the 2015 files).  Here the files are
parsed into three integer vectors -- COO flow records, city id per node, province id per node -- and

  * the inter-scale adjacency is coalesced on the device by K-1 (``Graph.from_coo``: 64-bit key radix sort +
    run-length count; multiplicities become the CSR values, exactly the dense ``+= 1`` counts),
  * the city / province adjacencies stay *group-id vectors*: the intra-scale kernels (``intra.py``) take the
    member lists of a block-structured adjacency, so the (N, N) matrices never exist.

``HigherDataset`` keeps the reference's Dataset surface (``__getitem__`` -> ``(source, recipient)``, ``__len__``,
``get_adjacent``, ``get_gdp``, ``get_count``), so ``train.py:180-195`` runs against it with only the import
changed; what ``get_adjacent`` returns is accepted by the drop-in models in place of the dense tensors.

File formats (``anonymous_data/``):
  ``Adjacent<year>.json`` (the reference code calls it ``indexMatch<year>.json``, dataset.py:219): ``source_index``
      ``{"<node>": [.., city, province]}`` in node order, ``recipient_index`` ``{name: column}``;
  ``GDP<year>.json``: ``{"GDP_embedding": {"<node>": float}}``;
  ``Flow<year>.csv`` (gb18030, one header line): ``source, recipient[, city, province]`` integers per record.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np
import torch

_INDEX_NAMES = ("indexMatch{year}.json", "Adjacent{year}.json")   # dataset.py:219 / the shipped file name


@dataclass
class FlowFiles:
    """Host-side content of one year's files (integers only; nothing dense)."""
    source: np.ndarray          # int64 [records]   Flow csv column 0   (dataset.py:229)
    recipient: np.ndarray       # int64 [records]   Flow csv column 1   (dataset.py:230)
    city: np.ndarray            # int64 [N]  city id of node i          (graph_dict values[-2], dataset.py:268)
    province: np.ndarray        # int64 [N]  province id of node i      (graph_dict values[-1], dataset.py:273)
    gdp: dict                   # {"<node>": float} in file order       (dataset.py:215)
    recipient_index: dict       # {name: column}                        (dataset.py:222)

    @property
    def n_sources(self) -> int:
        return int(self.city.shape[0])

    @property
    def n_recipients(self) -> int:
        return len(self.recipient_index)


def _find(root: str, patterns, year: str) -> str:
    for p in patterns:
        path = os.path.join(root, p.format(year=year))
        if os.path.isfile(path):
            return path
    raise FileNotFoundError(f"none of {[p.format(year=year) for p in patterns]} found under {root}")


def read_flow_files(root: str, year: str = "2015") -> FlowFiles:
    """Parse ``GDP<year>.json``, ``indexMatch|Adjacent<year>.json`` and ``Flow<year>.csv`` (dataset.py:215-232)."""
    year = str(year)
    with open(_find(root, ("GDP{year}.json",), year), "r", encoding="gbk") as f:
        gdp = json.load(f)["GDP_embedding"]
    with open(_find(root, _INDEX_NAMES, year), "r", encoding="gbk") as f:
        idx = json.load(f)
    graph_dict = idx["source_index"]
    recipient_index = idx["recipient_index"]
    n = len(graph_dict)
    # node i is the i-th entry of the dict (the reference enumerates .items(), dataset.py:266); the last two values are
    # (city, province): the shipped files hold [city, province], the code indexes values[1], values[2] of a 3-tuple
    groups = np.empty((n, 2), dtype=np.int64)
    for i, values in enumerate(graph_dict.values()):
        if len(values) < 2:
            raise ValueError(f"source_index entry {i} has {len(values)} values; need [.., city, province]")
        groups[i, 0] = values[-2]
        groups[i, 1] = values[-1]
    flow_path = _find(root, ("Flow{year}.csv",), year)
    rec = np.loadtxt(flow_path, delimiter=",", skiprows=1, usecols=(0, 1), dtype=np.int64, encoding="gb18030", ndmin=2)
    source = np.ascontiguousarray(rec[:, 0])
    recipient = np.ascontiguousarray(rec[:, 1])
    m = len(recipient_index)
    if source.size and (source.min() < 0 or source.max() >= n or recipient.min() < 0 or recipient.max() >= m):
        # the reference would raise IndexError inside inter_adjacent (dataset.py:287)
        raise IndexError(f"{flow_path}: record outside the ({n}, {m}) node tables")
    return FlowFiles(source, recipient, groups[:, 0].copy(), groups[:, 1].copy(), gdp, recipient_index)


class HigherDataset(torch.utils.data.Dataset):
    """Drop-in for ``dataset.HigherDataset`` (dataset.py:208-296) on top of :func:`read_flow_files`.

    ``root`` / ``year`` replace the reference's hard-coded directory and module-level ``year`` (dataset.py:13);
    both default from ``MSHA_DATA_ROOT`` / ``MSHA_DATA_YEAR``.  ``device`` is where ``get_adjacent`` builds the graph
    (CUDA only -- there is no host graph build)."""

    def __init__(self, root: str = None, year: str = None, device="cuda"):
        root = root or os.environ.get("MSHA_DATA_ROOT", "anonymous_data")
        year = str(year or os.environ.get("MSHA_DATA_YEAR", "2015"))
        self.files = read_flow_files(root, year)
        self.year = year
        self.device = torch.device(device)
        self.source = self.files.source.tolist()
        self.recipient = self.files.recipient.tolist()
        self.GDP = self.files.gdp
        self.count = len(self.source)
        self.N = self.files.n_sources
        self.M = self.files.n_recipients
        self._adjacent = None

    def __getitem__(self, index):
        return self.source[index], self.recipient[index]                      # dataset.py:238-241

    def __len__(self):
        return self.count

    def get_gdp(self):
        return self.GDP

    def get_count(self):
        return self.N, self.M                                                  # dataset.py:249-252

    def get_adjacent(self):
        """``(inter, city, province)`` like dataset.py:243-244 -- ``inter`` a device :class:`Graph` whose CSR values
        are the record counts, ``city`` / ``province`` int64 group-id vectors on the device."""
        if self._adjacent is None:
            from .graph import Graph
            if self.device.type != "cuda":
                raise RuntimeError("msha_b200 is CUDA-only: HigherDataset.get_adjacent builds the graph on the GPU")
            src = torch.from_numpy(self.files.source).to(self.device)
            dst = torch.from_numpy(self.files.recipient).to(self.device)
            inter = Graph.from_coo(src, dst, self.N, self.M)
            city = torch.from_numpy(self.files.city).to(self.device)
            province = torch.from_numpy(self.files.province).to(self.device)
            self._adjacent = (inter, city, province)
        return self._adjacent


def normalize_adjacency_matrix(adj):
    """``model.normalize_adjacency_matrix`` (model.py:95-100) for the objects ``get_adjacent`` returns.

    The attention layers only read ``adj > 0`` (GAT.py:30, Ours.py:67,81-82) and column normalisation keeps the sign
    pattern, so for a :class:`Graph` this returns a graph sharing the structure whose values are the column-normalised
    counts (what ``GraphConvolution`` consumes); a dense (N, M) tensor is compacted first; a group-id vector is returned
    unchanged.  Like the reference, a column without any record turns every value into NaN (inf * 0 in the dense ``mm``)."""
    from .graph import Graph
    if isinstance(adj, torch.Tensor) and adj.dim() == 1:
        return adj
    if isinstance(adj, torch.Tensor) and adj.dim() == 2:
        adj = Graph.from_dense(adj)
    if not isinstance(adj, Graph):
        raise TypeError("normalize_adjacency_matrix: expected a Graph, a group-id vector or a dense (N, M) tensor")
    vals = adj.normalized_values()
    empty = torch.bincount(adj.col.long(), minlength=adj.n_cols) == 0
    vals = torch.where(empty.any(), torch.full_like(vals, float("nan")), vals)      # inf * 0 of the dense mm
    g = Graph(adj.rowptr, adj.col, vals, adj.n_rows, adj.n_cols, adj.isolated)
    g._att, g._csc, g._csc_plain, g._deg, g._hubs = adj._att, adj._csc, adj._csc_plain, adj._deg, adj._hubs
    return g


def write_flow_files(root: str, year: str, source, recipient, city, province, gdp_values, recipient_names=None):
    """Write one year's files in the reference's format (synthetic datasets for tests, tools and the GPU box, where
    the reference's ``anonymous_data/`` does not exist)."""
    year = str(year)
    os.makedirs(root, exist_ok=True)
    n = len(city)
    m = int(max(recipient)) + 1 if recipient_names is None else len(recipient_names)
    names = recipient_names or [f"r{j}" for j in range(m)]
    with open(os.path.join(root, f"Adjacent{year}.json"), "w", encoding="gbk") as f:
        json.dump({"source_index": {str(i): [int(city[i]), int(province[i])] for i in range(n)},
                   "recipient_index": {names[j]: j for j in range(m)}}, f)
    with open(os.path.join(root, f"GDP{year}.json"), "w", encoding="gbk") as f:
        json.dump({"GDP_embedding": {str(i): float(gdp_values[i]) for i in range(n)}}, f)
    with open(os.path.join(root, f"Flow{year}.csv"), "w", encoding="gb18030") as f:
        f.write("source,recipient\n")
        for s, r in zip(source, recipient):
            f.write(f"{int(s)},{int(r)},{int(city[int(s)])},{int(province[int(s)])}\n")
