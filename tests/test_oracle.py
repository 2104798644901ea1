"""Pins oracle/msha_oracle.py (the CPU restatement) to the golden vectors produced by the
unmodified reference classes (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, params_of, sub_params, rel_err
from oracle import msha_oracle as O

TOL = 2e-5   # reference is fp32, oracle fp64: agreement is limited by the reference's rounding


def _grad(out, G, tensors):
    gs = torch.autograd.grad((out * torch.as_tensor(G, dtype=out.dtype)).sum(), tensors, allow_unused=True)
    return [torch.zeros_like(t) if g is None else g for g, t in zip(gs, tensors)]


def _leaf(x):
    return torch.tensor(np.asarray(x), dtype=torch.float64, requires_grad=True)


def test_csr_from_dense_matches_nonzero():
    rng = np.random.default_rng(0)
    adj = (rng.random((23, 9)) < 0.3) * rng.integers(1, 4, (23, 9))
    adj = adj.astype(np.float32)
    adj[4] = 0
    adj[7, 2] = -1.0          # only > 0 counts (GAT.py:30)
    rowptr, col, val = O.csr_from_dense(adj)
    ref = torch.nonzero(torch.from_numpy(adj) > 0)
    assert np.array_equal(col, ref[:, 1].numpy().astype(np.int32))
    rows = np.repeat(np.arange(23), np.diff(rowptr))
    assert np.array_equal(rows, ref[:, 0].numpy())
    pos = np.where(adj > 0, adj, 0)
    csr = torch.from_numpy(pos).to_sparse_csr()
    assert np.array_equal(rowptr, csr.crow_indices().numpy().astype(np.int32))
    assert np.array_equal(col, csr.col_indices().numpy().astype(np.int32))
    assert np.array_equal(val, csr.values().numpy())


def test_csr_from_coo_matches_dense_accumulate():
    rng = np.random.default_rng(1)
    n, N, M = 500, 40, 7
    s = rng.integers(0, N, n)
    d = rng.integers(0, M, n)
    dense = torch.zeros(N, M)
    for a, b in zip(s, d):                       # dataset.py:286-288
        dense[a][b] += 1
    rowptr, col, val = O.csr_from_coo(s, d, N, M)
    r2, c2, v2 = O.csr_from_dense(dense.numpy())
    assert np.array_equal(rowptr, r2) and np.array_equal(col, c2) and np.array_equal(val, v2)
    coo = torch.sparse_coo_tensor(np.stack([s, d]), torch.ones(n), (N, M)).coalesce()
    assert np.array_equal(col, coo.indices()[1].numpy().astype(np.int32))
    assert np.array_equal(val, coo.values().numpy())
    with pytest.raises(IndexError):
        O.csr_from_coo(np.array([N]), np.array([0]), N, M)
    rp, c, v = O.csr_from_coo(np.array([], dtype=np.int64), np.array([], dtype=np.int64), 3, 2)
    assert rp.tolist() == [0, 0, 0, 0] and c.size == 0


def test_csc_from_csr():
    rng = np.random.default_rng(2)
    adj = (rng.random((31, 11)) < 0.25).astype(np.float32)
    rowptr, col, val = O.csr_from_dense(adj)
    colptr, row, perm = O.csc_from_csr(rowptr, col, 11)
    rt, ct, _ = O.csr_from_dense(adj.T.copy())
    assert np.array_equal(colptr, rt) and np.array_equal(row, ct)
    rows = np.repeat(np.arange(31), np.diff(rowptr))
    assert np.array_equal(rows[perm], row) and np.array_equal(col[perm], np.repeat(np.arange(11), np.diff(colptr)))


def test_normalize_adjacency():
    g = load_golden("gcn")
    out = O.normalize_adjacency_dense(torch.from_numpy(g["adj"]).double())
    assert rel_err(out.numpy(), g["adj_norm"]) < TOL
    z = O.normalize_adjacency_dense(torch.from_numpy(g["adj_zero"]).double())
    assert np.isnan(g["adj_zero_norm"]).all() and torch.isnan(z).all()
    rowptr, col, val = O.csr_from_dense(g["adj"])
    vn = O.normalize_csr_values(val, col, g["adj"].shape[1])
    r = np.repeat(np.arange(g["adj"].shape[0]), np.diff(rowptr))
    assert rel_err(vn, g["adj_norm"][r, col]) < TOL


def test_graph_attention_layer():
    g = load_golden("gal")
    p = params_of(g)
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    x, W, a = _leaf(g["x"]), _leaf(p["W"]), _leaf(p["a"])
    out = O.graph_attention_layer(x, W, a, rowptr, col)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    gx, gW, ga = _grad(out, g["G"], [x, W, a])
    assert rel_err(gx.numpy(), g["gx"]) < TOL
    assert rel_err(gW.numpy(), g["gW"]) < TOL
    # d/da is identically zero in exact arithmetic; the reference shows round-off only
    assert np.max(np.abs(g["ga"])) < 1e-5 and np.max(np.abs(ga.numpy())) < 1e-12


def test_gat_model():
    g = load_golden("gat")
    p = params_of(g)
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    leaves = {k: _leaf(v) for k, v in p.items()}
    heads = [(leaves[f"attention_{i}.W"], leaves[f"attention_{i}.a"]) for i in range(2)]
    out = O.gat_model(leaves["features"], heads, (leaves["out_att.W"], leaves["out_att.a"]), rowptr, col)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    names = ["features", "attention_0.W", "attention_1.W", "out_att.W"]
    grads = _grad(out, g["G"], [leaves[n] for n in names])
    for n, gr in zip(names, grads):
        assert rel_err(gr.numpy(), g["g." + n]) < 5e-5, n


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_ours_layers(variant, mode):
    g = load_golden(f"ourslayer{variant}_{mode}")
    p = {k: _leaf(v) if v.dtype.kind == "f" else v for k, v in params_of(g).items()}
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    S, R = _leaf(g["S"]), _leaf(g["R"])
    training = mode == "train"
    if variant == 3:
        out = O.ours_layer3(S, R, p, rowptr, col, training)
    else:
        out = O.ours_layer(S, R, p, rowptr, col, O.group_adjacency_dense(g["city"]),
                           O.group_adjacency_dense(g["prov"]), g["src"], training, variant=variant)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    names = ["W1", "W2", "a", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"]
    if variant != 3:
        names += ["a3", "a4"]
    grads = _grad(out, g["G"], [S, R] + [p[n] for n in names])
    assert rel_err(grads[0].numpy(), g["gS"]) < 1e-4
    assert rel_err(grads[1].numpy(), g["gR"]) < 1e-4
    for n, gr in zip(names, grads[2:]):
        ref = g["g." + n]
        if np.max(np.abs(ref)) < 1e-5:           # identically-zero gradients (e.g. a3/a4 in variant 2)
            assert np.max(np.abs(gr.numpy())) < 1e-5, n
        else:
            assert rel_err(gr.numpy(), ref) < 1e-4, n


def test_bn_running_update():
    g = load_golden("ourslayer3_train")
    p = params_of(g)
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    S, R = O._t(g["S"]), O._t(g["R"])
    h1, h2 = R @ O._t(p["W1"]), S @ O._t(p["W2"])
    alpha, r, c, _ = O.inter_attention(h1, h2, O._t(p["a"]), rowptr, col)
    v_in = torch.zeros(h1.shape[0], h2.shape[1], dtype=torch.float64).index_add(0, c, alpha[:, None] * h2[r])
    rm, rv = O.batchnorm1d_running_update(v_in, O._t(p["bn1.running_mean"]), O._t(p["bn1.running_var"]))
    assert rel_err(rm.numpy(), g["after.bn1.running_mean"]) < TOL
    assert rel_err(rv.numpy(), g["after.bn1.running_var"]) < TOL


def test_ours_record_coefficients():
    g = load_golden("ours_record")
    p = params_of(g)
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    out, a12, a3, a4 = O.ours_layer(g["S"], g["R"], p, rowptr, col, O.group_adjacency_dense(g["city"]),
                                    O.group_adjacency_dense(g["prov"]), g["src"], training=False,
                                    return_coeffs=True)
    assert rel_err(out.numpy(), g["out"]) < TOL
    assert rel_err(a12.numpy(), g["coeff12"]) < TOL
    assert rel_err(a3.numpy(), g["coeff3"][g["src"]]) < TOL
    assert rel_err(a4.numpy(), g["coeff4"][g["src"]]) < TOL


@pytest.mark.parametrize("name,variant", [("ablation2", 2), ("ablation3", 3), ("ours", 1)])
def test_msha_models(name, variant):
    g = load_golden(name)
    p = {k: _leaf(v) if v.dtype.kind == "f" else v for k, v in params_of(g).items()}
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    heads = [sub_params(p, f"attention_{i}.") for i in range(2)]
    out = O.msha_model(p["Sfeatures"], p["Rfeatures"], heads, (p["out_att.W"], p["out_att.a"]), rowptr, col,
                       O.group_adjacency_dense(g["city"]), O.group_adjacency_dense(g["prov"]), g["src"],
                       training=True, variant=variant)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    loss = O.nll_readout(out, g["src"], g["rec"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    names = ["Sfeatures", "Rfeatures", "attention_0.W1", "attention_1.W2", "attention_0.a", "out_att.W"]
    grads = torch.autograd.grad(loss, [p[n] for n in names])
    for n, gr in zip(names, grads):
        assert rel_err(gr.numpy(), g["g." + n]) < 2e-4, n


def test_ablation1_model():
    g = load_golden("ablation1")
    p = {k: _leaf(v) if v.dtype.kind == "f" else v for k, v in params_of(g).items()}
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    out = O.ablation1_model(p["Sfeatures"], p["Rfeatures"], sub_params(p, "attention."), rowptr, col,
                            O.group_adjacency_dense(g["city"]), O.group_adjacency_dense(g["prov"]), g["src"])
    assert rel_err(out.detach().numpy(), g["out"]) < TOL


def test_hgane_layer():
    g = load_golden("hgane")
    p = {k: _leaf(v) if v.dtype.kind == "f" else v for k, v in params_of(g).items()}
    out = O.hgane_layer(p, g["adj_inter"], g["adj_intra"], g["src"], training=True)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    names = ["source_embedding", "recipient_embedding", "W1.weight", "W2.weight", "a12.weight", "a3.weight"]
    grads = _grad(out, g["G"], [p[n] for n in names])
    for n, gr in zip(names, grads):
        assert rel_err(gr.numpy(), g["g." + n]) < 2e-4, n


@pytest.mark.parametrize("tag,predictor,nl", [("mlp2", "mlp", 2), ("mlp3", "mlp", 3), ("inner", "inner", 2)])
def test_link_predictor(tag, predictor, nl):
    g = load_golden("linkpred_" + tag)
    p = params_of(g)
    h = _leaf(g["h"])
    Ws = [_leaf(p[f"lins.{i}.weight"]) for i in range(nl)]
    bs = [_leaf(p[f"lins.{i}.bias"]) for i in range(nl)]
    out = O.link_predictor(h[g["src"]], h[g["dst"]], Ws, bs, predictor)
    assert out.shape == g["out"].shape
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    grads = _grad(out, g["G"], [h] + Ws[:-1] + bs[:-1])
    assert rel_err(grads[0].numpy(), g["gh"]) < 1e-4
    if predictor == "mlp":
        for i in range(nl - 1):
            assert rel_err(grads[1 + i].numpy(), g[f"g.lins.{i}.weight"]) < 1e-4
            assert rel_err(grads[nl + i].numpy(), g[f"g.lins.{i}.bias"]) < 1e-4
    assert np.all(g[f"g.lins.{nl-1}.weight"] == 0)      # the last Linear never runs (LLP.py:111)


def test_graph_convolution():
    g = load_golden("gcn")
    p = params_of(g)
    rowptr, col, val = O.csr_from_dense(g["adj"])
    vn = O.normalize_csr_values(val, col, g["adj"].shape[1])
    x, w, b = _leaf(g["x"]), _leaf(p["weight"]), _leaf(p["bias"])
    out = O.graph_convolution(x, w, b, rowptr, col, vn, g["adj"].shape[1])
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    gx, gw, gb = _grad(out, g["G"], [x, w, b])
    assert rel_err(gx.numpy(), g["gx"]) < 1e-4 and rel_err(gw.numpy(), g["gw"]) < 1e-4
    assert rel_err(gb.numpy(), g["gb"]) < 1e-4


def test_generic_gat_layer():
    g = load_golden("generic_gat")
    H, Fin, d = g["W"].shape
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    W = np.concatenate([g["W"][h] for h in range(H)], axis=1)            # (F, H*d')
    a_nbr = np.stack([g["a"][h][:d, 0] for h in range(H)])
    a_self = np.stack([g["a"][h][d:, 0] for h in range(H)])
    out, alpha = O.gat_layer(g["X"], W, a_nbr, a_self, rowptr, col, H, concat=True, apply_elu=False,
                             return_alpha=True)
    N = g["X"].shape[0]
    r, c, _ = O.attention_edges(rowptr, col, N)
    for h in range(H):
        dense = O.dense_from_edges(alpha[:, h], r, c, N, N)
        assert rel_err(dense.numpy(), g["alpha"][h]) < TOL
        assert rel_err(out.view(N, H, d)[:, h].numpy(), g["agg"][h]) < TOL


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = O.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == want


def test_negative_sampler_and_dropout_stream():
    s, d = O.negative_sample(42, 1000, 37, 11)
    assert s.min() >= 0 and s.max() < 37 and d.min() >= 0 and d.max() < 11
    s2, d2 = O.negative_sample(42, 10, 37, 11)
    assert np.array_equal(s[:10], s2) and np.array_equal(d[:10], d2)      # counter based: prefix stable
    keep = O.dropout_keep_mask(7, 200000, 0.5)
    assert abs(keep.mean() - 0.5) < 0.01
    assert O.dropout_keep_mask(7, 100, 0.0).all()
