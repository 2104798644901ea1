"""GPU parity of K-1 (graph build) and the Philox streams: bit-exact against the numpy oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import msha_gnn_b200 as mg
from msha_gnn_b200 import ops
from msha_gnn_b200.ops import call, ptr, workspace, _stream
from oracle import msha_oracle as O

DEV = "cuda:0"


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("n", [0, 1, 31, 2048, 2049, 100000, 5_000_001])
def test_exclusive_scan(n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 5, n).astype(np.int32)
    xin = torch.from_numpy(x).to(DEV)
    out = torch.empty(n + 1, dtype=torch.int32, device=DEV)
    lib = mg._lib.lib()
    ws = workspace(lib.msha_scan_workspace_bytes(n + 1), DEV)
    call("msha_scan_exclusive_i32", ptr(xin, torch.int32), n, ptr(out, torch.int32), n + 1, ws.data_ptr(), ws.numel(), _stream())
    want = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(x, out=want[1:])
    assert np.array_equal(_np(out).astype(np.int64), want)


@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 13), (4097, 40), (300000, 64), (2_000_000, 42)])
def test_radix_sort_stable(n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** min(bits, 62), n, dtype=np.int64).astype(np.uint64)
    if n > 10:
        keys[: n // 2] = keys[n // 2: n // 2 * 2]           # many duplicates -> stability matters
    vals = np.arange(n, dtype=np.uint32)
    k = torch.from_numpy(keys.view(np.int64)).to(DEV)
    kt = torch.empty_like(k)
    v = torch.from_numpy(vals.view(np.int32)).to(DEV)
    vt = torch.empty_like(v)
    lib = mg._lib.lib()
    ws = workspace(lib.msha_radix_sort_workspace_bytes(n), DEV)
    call("msha_radix_sort_u64", k.data_ptr(), kt.data_ptr(), v.data_ptr(), vt.data_ptr(), n, 0, bits, ws.data_ptr(),
         ws.numel(), _stream())
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(_np(k).view(np.uint64), keys[order])
    assert np.array_equal(_np(v).view(np.uint32), vals[order])


def _check_graph(g, rowptr, col, val, n_cols):
    assert np.array_equal(_np(g.rowptr), rowptr)
    assert np.array_equal(_np(g.col), col)
    assert np.array_equal(_np(g.val), val)
    colptr, rowidx, perm = g.transpose_structure()
    cp, ri, pm = O.csc_from_csr(rowptr, col, n_cols)
    assert np.array_equal(_np(colptr), cp) and np.array_equal(_np(rowidx), ri) and np.array_equal(_np(perm), pm)


@pytest.mark.parametrize("N,M,density", [(1, 1, 1.0), (37, 8, 0.3), (500, 32, 0.07), (300, 300, 0.2), (64, 1000, 0.01),
                                         (5, 70, 0.0)])
def test_csr_from_dense_bit_exact(N, M, density):
    rng = np.random.default_rng(N * 1000 + M)
    adj = ((rng.random((N, M)) < density) * rng.integers(1, 6, (N, M))).astype(np.float32)
    if N > 3:
        adj[2] = 0                      # isolated row
        adj[3, :] = -1.0                # negatives are not neighbours (adj > 0, GAT.py:30)
    g = mg.Graph.from_dense(torch.from_numpy(adj).to(DEV))
    rowptr, col, val = O.csr_from_dense(adj)
    _check_graph(g, rowptr, col, val, M)
    nz = torch.nonzero(torch.from_numpy(adj) > 0).numpy()          # the reference's canonical order
    assert np.array_equal(_np(g.edge_index()).T, nz)
    # non-contiguous (row-strided) input
    big = torch.zeros(N, M + 5, device=DEV)
    big[:, :M] = torch.from_numpy(adj).to(DEV)
    g2 = mg.Graph.from_dense(big[:, :M])
    assert np.array_equal(_np(g2.col), col)


@pytest.mark.parametrize("n,N,M", [(0, 5, 3), (1, 1, 1), (5000, 400, 32), (200000, 39179, 32), (300000, 5000, 5000),
                                   (100000, 3, 2)])
def test_csr_from_coo_bit_exact(n, N, M):
    rng = np.random.default_rng(n + N)
    s = rng.integers(0, N, n)
    d = rng.integers(0, M, n)
    if n > 100:
        s[:50] = s[50:100]; d[:50] = d[50:100]              # duplicate records -> multiplicities
        s[100:110] = N - 1; d[100:110] = M - 1             # maximum indices
    g = mg.Graph.from_coo(torch.from_numpy(s).to(DEV), torch.from_numpy(d).to(DEV), N, M)
    rowptr, col, val = O.csr_from_coo(s, d, N, M)
    _check_graph(g, rowptr, col, val, M)
    if n:
        coo = torch.sparse_coo_tensor(np.stack([s, d]), torch.ones(n), (N, M)).coalesce()
        assert np.array_equal(_np(g.edge_index()), coo.indices().numpy())
        assert np.array_equal(_np(g.val), coo.values().numpy())


def test_csr_from_coo_out_of_range():
    s = torch.tensor([0, 7], device=DEV)
    d = torch.tensor([0, 0], device=DEV)
    with pytest.raises(IndexError):
        mg.Graph.from_coo(s, d, 7, 3)


def test_hub_and_powerlaw_rows():
    rng = np.random.default_rng(5)
    N = 3000
    deg = np.minimum((rng.pareto(1.2, N) * 3 + 1).astype(np.int64), N)
    deg[7] = N                       # a full hub row
    s = np.repeat(np.arange(N), deg)
    d = np.concatenate([rng.choice(N, k, replace=False) for k in deg])
    g = mg.Graph.from_edge_index(torch.from_numpy(np.stack([s, d])).to(DEV), N)
    rowptr, col, val = O.csr_from_coo(s, d, N, N)
    _check_graph(g, rowptr, col, val, N)


def test_attention_csr_isolated_rows():
    rng = np.random.default_rng(9)
    adj = (rng.random((40, 6)) < 0.3).astype(np.float32)
    adj[[0, 17, 39]] = 0
    g = mg.Graph.from_dense(torch.from_numpy(adj).to(DEV))
    rp, col = g.attention_csr()
    rowptr, c, _ = O.csr_from_dense(adj)
    r, cc, masked = O.attention_edges(rowptr, c, 6)
    got_c = _np(col)
    dec = np.where(got_c < 0, ~got_c, got_c)
    assert g.n_isolated == int((adj.sum(1) == 0).sum())
    assert np.array_equal(dec, cc.numpy()) and np.array_equal(got_c < 0, masked.numpy())
    assert np.array_equal(np.repeat(np.arange(40), np.diff(_np(rp))), r.numpy())
    colptr, rowidx, perm = g.attention_csc()
    order = np.argsort(dec, kind="stable")
    assert np.array_equal(_np(perm), order) and np.array_equal(_np(rowidx), r.numpy()[order])


def test_normalized_values():
    rng = np.random.default_rng(11)
    adj = ((rng.random((50, 9)) < 0.4) * rng.integers(1, 4, (50, 9))).astype(np.float32)
    adj[:, 0] += 1
    g = mg.Graph.from_dense(torch.from_numpy(adj).to(DEV))
    rowptr, col, val = O.csr_from_dense(adj)
    want = O.normalize_csr_values(val, col, 9)
    np.testing.assert_allclose(_np(g.normalized_values()), want, rtol=2e-6)


def test_negative_sampler_and_dropout_stream_bit_exact():
    for seed, P, ns, nd in [(42, 1001, 4267, 4267), (2 ** 40 + 3, 64, 7, 3)]:
        s, d = mg.functional.negative_sample(seed, P, ns, nd, DEV)
        so, do = O.negative_sample(seed, P, ns, nd)
        assert np.array_equal(_np(s), so) and np.array_equal(_np(d), do)
    keep = torch.empty(10007, dtype=torch.uint8, device=DEV)
    call("msha_dropout_mask", 77, 2, 10007, 0.3, keep.data_ptr(), _stream())
    assert np.array_equal(_np(keep).astype(bool), O.dropout_keep_mask(77, 10007, 0.3, stream=2))


def test_graph_cache_reuses_structure():
    adj = (torch.rand(30, 5, device=DEV) < 0.4).float()
    g1 = mg.as_graph(adj)
    assert mg.as_graph(adj) is g1
    adj[0, 0] = 1 - adj[0, 0]           # in-place edit bumps _version -> rebuild
    assert mg.as_graph(adj) is not g1
