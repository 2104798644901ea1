"""Property tests (hypothesis) of the integer part of the oracle -- the CSR / CSC build that the device graph build has to
match bit for bit -- against scipy's coalescing COO -> CSR conversion and torch's sparse tensors.  CPU only."""
import numpy as np
import scipy.sparse as sp
import torch
from hypothesis import given, settings, strategies as st

from oracle import msha_oracle as O


@st.composite
def coo(draw):
    n_rows = draw(st.integers(1, 40))
    n_cols = draw(st.integers(1, 12))
    n = draw(st.integers(0, 300))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_rows, n), rng.integers(0, n_cols, n), n_rows, n_cols


@settings(max_examples=60, deadline=None)
@given(coo())
def test_csr_from_coo_equals_scipy_coalesce(c):
    src, dst, n_rows, n_cols = c
    rowptr, col, val = O.csr_from_coo(src, dst, n_rows, n_cols)
    m = sp.coo_matrix((np.ones(src.size, dtype=np.float32), (src, dst)), shape=(n_rows, n_cols)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    assert np.array_equal(rowptr, m.indptr) and np.array_equal(col, m.indices) and np.array_equal(val, m.data)
    assert rowptr.dtype == np.int32 and col.dtype == np.int32 and val.dtype == np.float32
    assert float(val.sum()) == float(src.size)                       # multiplicities, dataset.py:286-288


@settings(max_examples=60, deadline=None)
@given(coo())
def test_csc_is_the_transpose_and_perm_maps_slots(c):
    src, dst, n_rows, n_cols = c
    rowptr, col, val = O.csr_from_coo(src, dst, n_rows, n_cols)
    colptr, rowidx, perm = O.csc_from_csr(rowptr, col, n_cols)
    t = sp.csr_matrix((val, col, rowptr), shape=(n_rows, n_cols)).tocsc()
    t.sort_indices()
    assert np.array_equal(colptr, t.indptr) and np.array_equal(rowidx, t.indices)
    assert np.array_equal(val[perm], t.data)                         # perm: CSC slot -> CSR slot
    assert np.array_equal(np.sort(perm), np.arange(col.size))        # a permutation
    rows_of = np.repeat(np.arange(n_rows), np.diff(rowptr))
    assert np.array_equal(rows_of[perm], rowidx)
    for j in range(n_cols):                                          # rows ascend inside a column
        seg = rowidx[colptr[j]:colptr[j + 1]]
        assert np.all(np.diff(seg) > 0)


@settings(max_examples=40, deadline=None)
@given(coo())
def test_csr_from_dense_equals_torch_nonzero_and_attention_edges(c):
    src, dst, n_rows, n_cols = c
    dense = np.zeros((n_rows, n_cols), dtype=np.float32)
    np.add.at(dense, (src, dst), 1.0)
    rowptr, col, val = O.csr_from_dense(dense)
    nz = torch.nonzero(torch.from_numpy(dense) > 0)                  # the reference's mask `adj > 0` (GAT.py:30)
    assert np.array_equal(col, nz[:, 1].numpy()) and np.array_equal(np.repeat(np.arange(n_rows), np.diff(rowptr)), nz[:, 0].numpy())
    r, cc, masked = O.attention_edges(rowptr, col, n_cols)
    deg = np.diff(rowptr)
    n_iso = int((deg == 0).sum())
    assert r.numel() == col.size + n_iso * n_cols                    # rows without neighbours attend to all M columns
    assert int(masked.sum()) == n_iso * n_cols
    vals = O.normalize_csr_values(val, col, n_cols)
    colsum = dense.sum(axis=0)
    want = (dense / np.where(colsum > 0, colsum, 1.0))[dense > 0]
    np.testing.assert_allclose(vals, want, rtol=2e-6)


def test_gat_layer_rows_equals_gat_layer_on_the_sampled_rows():
    """The rows-restricted restatement bench.py's parity block uses == the whole-graph one (itself pinned to the reference's
    OursLayer3 lines by tests/test_oracle.py::test_generic_gat) on the rows it samples."""
    import torch
    from oracle import msha_oracle as O
    rng = np.random.default_rng(0)
    N, Fin, H, d = 60, 7, 4, 3
    adj = (rng.random((N, N)) < 0.15).astype(np.float32)
    adj[np.arange(N), np.arange(N)] = 1
    rowptr, col, _ = O.csr_from_dense(adj)
    x = torch.tensor(rng.standard_normal((N, Fin)))
    W = torch.tensor(rng.standard_normal((Fin, H * d)))
    an, as_ = torch.tensor(rng.standard_normal((H, d))), torch.tensor(rng.standard_normal((H, d)))
    full = O.gat_layer(x, W, an, as_, rowptr, col, H)
    S = np.array([3, 17, 17, 59, 0])
    nbrs = [col[rowptr[i]:rowptr[i + 1]] for i in S]
    nodes, inv = np.unique(np.concatenate(nbrs + [S]), return_inverse=True)
    nbr_ptr = np.concatenate([[0], np.cumsum([len(n) for n in nbrs])])
    nbr_idx, self_idx = inv[:nbr_ptr[-1]], inv[nbr_ptr[-1]:]
    got = O.gat_layer_rows(x[nodes], W, an, as_, self_idx, nbr_ptr, nbr_idx, H)
    assert torch.allclose(got, full[S], atol=1e-12)
