"""Pins oracle/msha_oracle.py (the CPU restatement) to the golden vectors produced by the
unmodified reference classes (oracle/make_golden.py).  CPU only."""
import os

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


@pytest.mark.parametrize("name,H", [("gat", 2), ("gat_h8", 8)])
def test_gat_model(name, H):
    g = load_golden(name)
    p = params_of(g)
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    leaves = {k: _leaf(v) for k, v in p.items()}
    heads = [(leaves[f"attention_{i}.W"], leaves[f"attention_{i}.a"]) for i in range(H)]
    out = O.gat_model(leaves["features"], heads, (leaves["out_att.W"], leaves["out_att.a"]), rowptr, col)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    names = ["features", "out_att.W"] + [f"attention_{i}.W" for i in range(H)]
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


# ------------------------------------------------------------------ SURVEY 8f rows (LLP student, baselines, export)
def _lins(p, prefix, names=("lins",)):
    ws, bs = [], []
    k = 0
    while f"{prefix}{names[0]}.{k}.weight" in p:
        ws.append(p[f"{prefix}{names[0]}.{k}.weight"])
        bs.append(p[f"{prefix}{names[0]}.{k}.bias"])
        k += 1
    return ws, bs


def test_llp_step_loss():
    g = load_golden("llp_step")
    p = params_of(g)
    rowptr, col, _ = O.csr_from_dense(g["adj_norm"])
    heads = [(p[f"teacher.attention_{k}.W"], p[f"teacher.attention_{k}.a"]) for k in range(2)]
    student = tuple([_leaf(w) for w in x] for x in _lins(p, "model.", ("layers",)))
    pred = tuple([_leaf(w) for w in x] for x in _lins(p, "predictor."))
    loss, parts = O.llp_step_loss(g["features"], student, pred, heads, (p["teacher.out_att.W"], p["teacher.out_att.a"]),
                                  _lins(p, "teacher_pred."), rowptr, col, g["src"], g["rec"])
    for k in ("label_loss", "kd_f", "kd_p", "h", "t_h", "output", "t_out"):
        assert rel_err(parts[k].detach().numpy(), g[k]) < TOL, k
    # the weighted sum cancels (10 * label_loss ~ -7, 100 * mse ~ +7): compare against the size of its terms
    scale = 10.0 * abs(float(g["label_loss"])) + 0.1 * abs(float(g["kd_f"])) + 100.0 * abs(float(g["kd_p"]))
    assert abs(float(loss) - float(g["loss"])) < TOL * scale
    loss.backward()
    for tag, (ws, bs), nm in (("model", student, "layers"), ("predictor", pred, "lins")):
        for k, (w, b) in enumerate(zip(ws, bs)):
            for leaf, kind in ((w, "weight"), (b, "bias")):
                ref = g[f"g.{tag}.{nm}.{k}.{kind}"]
                got = np.zeros_like(ref) if leaf.grad is None else leaf.grad.numpy()
                if np.abs(ref).max() < 1e-7:
                    assert np.abs(got).max() < 1e-7
                else:
                    assert rel_err(got, ref) < 5e-5, (tag, k, kind, rel_err(got, ref))


def test_mlp3():
    g = load_golden("mlp3")
    p = params_of(g)
    ws, bs = _lins(p, "", ("layers",))
    x = _leaf(g["x"])
    wl, bl = [_leaf(w) for w in ws], [_leaf(b) for b in bs]
    out = O.mlp(x, wl, bl)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    grads = _grad(out, g["G"], [x] + wl + bl)
    assert rel_err(grads[0].numpy(), g["gx"]) < TOL
    for k in range(3):
        assert rel_err(grads[1 + k].numpy(), g[f"g.layers.{k}.weight"]) < TOL
        assert rel_err(grads[4 + k].numpy(), g[f"g.layers.{k}.bias"]) < TOL


@pytest.mark.parametrize("tag", ["c12", "c7"])
def test_kd_losses(tag):
    g = load_golden("kd_losses_" + tag)
    s, t = _leaf(g["s"]), _leaf(g["t"])
    kd = O.kd_cosine(s, t, g["idx_s"], g["idx_t"])
    assert rel_err(kd.detach().numpy(), g["kd"]) < TOL
    (gs,) = torch.autograd.grad(kd, [s])
    zero_row = 3                                              # its gradient is t / (eps |t|): ~1e8, checked apart
    keep = np.arange(g["s"].shape[0]) != zero_row
    assert rel_err(gs.numpy()[keep], g["gs"][keep]) < TOL
    assert rel_err(gs.numpy()[zero_row], g["gs"][zero_row]) < TOL
    full = O.kd_cosine(s, t, g["idx_s"], g["idx_t"], detach_teacher=False)
    gs2, gt2 = torch.autograd.grad(full, [s, t])
    assert rel_err(gs2.numpy()[keep], g["gs_full"][keep]) < TOL
    keep_t = np.arange(g["t"].shape[0]) != 5
    assert rel_err(gt2.numpy()[keep_t], g["gt_full"][keep_t]) < TOL
    assert rel_err(gt2.numpy()[5], g["gt_full"][5]) < TOL
    a, b = _leaf(g["a"]), _leaf(g["b"])
    mse = O.mse_loss(a, b)
    assert rel_err(mse.detach().numpy(), g["mse"]) < TOL
    ga, gb = torch.autograd.grad(mse, [a, b])
    assert rel_err(ga.numpy(), g["ga"]) < TOL and rel_err(gb.numpy(), g["gb"]) < TOL


def test_gcn_model():
    g = load_golden("gcn_model")
    p = {k: _leaf(v) for k, v in params_of(g).items()}
    N, M = g["adj"].shape
    rowptr, col, _ = O.csr_from_dense(g["adj"])
    val = O.normalize_csr_values(O.csr_from_dense(g["adj"])[2], col, M)
    np.testing.assert_allclose(val, g["adj_norm"][g["adj"] > 0], rtol=1e-6)
    out = O.gcn_model(p["features"], p, rowptr, col, val, N, M)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    names = ["features", "gc1.weight", "gc1.bias", "gc2.weight", "gc2.bias"]
    grads = _grad(out, g["G"], [p[n] for n in names])
    for n, gr in zip(names, grads):
        assert rel_err(gr.numpy(), g["g." + n]) < TOL, n
    assert np.abs(g["g.gc3.weight"]).max() == 0               # allocated, never applied (model.py:62-63)


def test_graphsage_model():
    g = load_golden("graphsage")
    p = {k: _leaf(v) for k, v in params_of(g).items()}
    N, M = g["adj"].shape
    rowptr, col, v0 = O.csr_from_dense(g["adj"])
    val = O.normalize_csr_values(v0, col, M)
    out = O.graphsage_model(p, g["src"], rowptr, col, val, M)
    assert rel_err(out.detach().numpy(), g["out"]) < TOL
    names = list(p)
    grads = _grad(out, g["G"], [p[n] for n in names])
    for n, gr in zip(names, grads):
        assert rel_err(gr.numpy(), g["g." + n]) < TOL, n


def test_explainer_argmax():
    dense = np.array([[0.2, 0.5, 0.5, 0.0], [0.0, 0.0, 0.0, 0.0], [0.1, 0.0, 0.0, 0.9]], dtype=np.float32)
    assert O.explainer_argmax(dense) == [[1, 2], [0, 1, 2, 3], [3]]
    assert O.explainer_argmax(dense.T) == [[0], [0], [0], [2]]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_committed_goldens_regenerate_bit_for_bit(tmp_path, monkeypatch):
    """tests/golden/*.npz are exactly what oracle/make_golden.py produces from the unmodified reference classes today."""
    from oracle import make_golden as mgold
    committed = mgold.OUT
    monkeypatch.setattr(mgold, "OUT", str(tmp_path))
    mgold.main()
    names = sorted(f for f in os.listdir(committed) if f.endswith(".npz"))
    assert names == sorted(os.listdir(tmp_path)) and len(names) == 27
    for f in names:
        with np.load(os.path.join(committed, f)) as a, np.load(os.path.join(tmp_path, f)) as b:
            assert set(a.files) == set(b.files), f
            for k in a.files:
                assert np.array_equal(a[k], b[k], equal_nan=True), (f, k)
