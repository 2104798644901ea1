"""Loader adapter: the host-side parse of the reference's on-disk files, pinned against the dense
matrices the reference's own ``HigherDataset.intra_adjacent`` / ``inter_adjacent`` produced (tests/golden/dataset.npz,
oracle/make_golden.py:case_dataset).  CPU only; the device graph build on top of it is in test_gpu_widen.py."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from msha_gnn_b200 import data as D
from oracle import msha_oracle as O

REAL = "/root/reference/anonymous_data"


def _write(tmp_path, g, year="2031"):
    D.write_flow_files(str(tmp_path), year, g["source"], g["recipient"], g["city"], g["province"], g["gdp"],
                       recipient_names=[f"R{j}" for j in range(int(g["n_recipients"]))])
    return str(tmp_path), year


def test_parse_matches_reference_dense_builders(tmp_path):
    g = load_golden("dataset")
    root, year = _write(tmp_path, g)
    ff = D.read_flow_files(root, year)
    N, M = g["inter"].shape
    assert (ff.n_sources, ff.n_recipients) == (N, M)
    assert np.array_equal(ff.source, g["source"]) and np.array_equal(ff.recipient, g["recipient"])
    # inter_adjacent (dataset.py:279-296): the coalesced COO is the dense count matrix, bit for bit
    rowptr, col, val = O.csr_from_coo(ff.source, ff.recipient, N, M)
    r2, c2, v2 = O.csr_from_dense(g["inter"])
    assert np.array_equal(rowptr, r2) and np.array_equal(col, c2) and np.array_equal(val, v2)
    assert rowptr[12] == rowptr[11]                      # the node without records is an empty row
    # intra_adjacent (dataset.py:260-277): group equality incl. the diagonal
    assert np.array_equal(O.group_adjacency_dense(ff.city), g["city_adj"])
    assert np.array_equal(O.group_adjacency_dense(ff.province), g["province_adj"])
    assert list(ff.gdp.keys()) == [str(i) for i in range(N)]
    np.testing.assert_allclose(list(ff.gdp.values()), g["gdp"])


def test_three_value_source_index_and_indexmatch_name(tmp_path):
    """dataset.py:219,268,273 reads ``indexMatch<year>.json`` with 3-tuples (values[1] = city, values[2] = province)."""
    g = load_golden("dataset")
    root, year = _write(tmp_path, g)
    adj_path = os.path.join(root, f"Adjacent{year}.json")
    with open(adj_path, encoding="gbk") as f:
        idx = json.load(f)
    idx["source_index"] = {k: [7, v[0], v[1]] for k, v in idx["source_index"].items()}
    os.remove(adj_path)
    with open(os.path.join(root, f"indexMatch{year}.json"), "w", encoding="gbk") as f:
        json.dump(idx, f)
    ff = D.read_flow_files(root, year)
    assert np.array_equal(ff.city, g["city"]) and np.array_equal(ff.province, g["province"])


def test_dataset_surface(tmp_path):
    g = load_golden("dataset")
    root, year = _write(tmp_path, g)
    ds = D.HigherDataset(root, year, device="cpu")
    assert len(ds) == g["source"].size
    assert ds[5] == (int(g["source"][5]), int(g["recipient"][5]))           # dataset.py:238-241
    assert ds.get_count() == g["inter"].shape
    assert ds.get_gdp() is ds.GDP and len(ds.GDP) == g["inter"].shape[0]
    loader = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)  # train.py:187-188
    s, r = next(iter(loader))
    assert s.dtype == torch.int64 and s.tolist() == g["source"][:64].tolist() and r.tolist() == g["recipient"][:64].tolist()
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ds.get_adjacent()                                                   # no host graph build


def test_errors(tmp_path):
    g = load_golden("dataset")
    root, year = _write(tmp_path, g)
    with pytest.raises(FileNotFoundError):
        D.read_flow_files(root, "1999")
    with open(os.path.join(root, f"Flow{year}.csv"), "a", encoding="gb18030") as f:
        f.write("9999,0,0,0\n")
    with pytest.raises(IndexError):
        D.read_flow_files(root, year)


@pytest.mark.skipif(not os.path.isdir(REAL), reason="reference data only exists in the build container")
def test_real_2015_files():
    """The shipped 2015 files: sizes of SURVEY.md section 8d cfg 1."""
    ff = D.read_flow_files(REAL, "2015")
    assert (ff.n_sources, ff.n_recipients, ff.source.size) == (39179, 32, 233887)
    rowptr, col, val = O.csr_from_coo(ff.source, ff.recipient, 39179, 32)
    assert col.size == 91283 and float(val.sum()) == 233887.0
    city_nnz = int((np.bincount(ff.city).astype(np.int64) ** 2).sum())
    prov_nnz = int((np.bincount(ff.province).astype(np.int64) ** 2).sum())
    assert abs(city_nnz - 8.93e6) < 0.01e6 and abs(prov_nnz - 83.3e6) < 0.1e6


def test_synthetic_flow_generator_is_2015_shaped():
    """bench.flow_graph (the generator behind the flow / yearly workloads and tools/train_flow.py --synthetic): every source has
    k distinct recipients (no duplicate edge before the repeated records are appended), nnz ~ 2.33 N like the real 2015 files
    (91 283 for N = 39 179), the requested record count, every recipient used."""
    import bench
    src, dst, city, prov = bench.flow_graph()
    assert src.size == 233887 and src.dtype == np.int64
    key = src * 32 + dst
    nnz = np.unique(key).size
    assert abs(nnz - 91283) < 0.02 * 91283
    k = np.bincount(np.unique(key) // 32, minlength=39179)
    assert k.min() >= 1 and k.max() <= 30 and abs(k.mean() - 2.33) < 0.05
    assert np.bincount(dst, minlength=32).min() > 0
    assert city.size == prov.size == 39179 and np.array_equal(prov, city % 25)
    s2, d2, _, _ = bench.flow_graph()
    assert np.array_equal(src, s2) and np.array_equal(dst, d2)          # deterministic
    # the files written from it parse back to the same records
    rowptr, col, val = O.csr_from_coo(src, dst, 39179, 32)
    assert col.size == nnz and float(val.sum()) == 233887.0
