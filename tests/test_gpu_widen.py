"""GPU parity of the callers either side of the hot path: the file loader feeding the device graph
build, the LLP student / distillation step, the GCN / GraphSAGE baselines and the attention export -- against the golden
vectors of the unmodified reference classes and the fp64 CPU oracle.  Tolerance: rel. err <= 1e-4 in fp32; integer
work bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import msha_gnn_b200 as mg
from msha_gnn_b200 import data as D
from msha_gnn_b200 import functional as Fn
from msha_gnn_b200 import llp as LLP
from conftest import load_golden, params_of, sub_params, rel_err
from oracle import msha_oracle as O

DEV = "cuda:0"
TOL = 1e-4


def _np(t):
    return t.detach().cpu().numpy()


def _t(a, grad=False):
    t = torch.tensor(np.asarray(a), dtype=torch.float32, device=DEV)
    return t.requires_grad_(True) if grad else t


def _i(a):
    return torch.tensor(np.asarray(a), dtype=torch.int64, device=DEV)


def _load(module, p, strict=True):
    module.load_state_dict({k: torch.tensor(v) for k, v in p.items()}, strict=strict)
    return module.to(DEV)


def _check_grads(named, g, prefix=""):
    for n, prm in named.items():
        ref = g["g." + prefix + n]
        if np.max(np.abs(ref)) < 1e-6:
            assert prm.grad is None or np.max(np.abs(_np(prm.grad))) < 1e-6, n
        else:
            assert prm.grad is not None, n
            assert rel_err(_np(prm.grad), ref) < TOL, (n, rel_err(_np(prm.grad), ref))


# ---------------------------------------------------------------------------------- 8f-2 loader -> device graph
def test_dataset_builds_reference_adjacency(tmp_path):
    g = load_golden("dataset")
    D.write_flow_files(str(tmp_path), "2031", g["source"], g["recipient"], g["city"], g["province"], g["gdp"],
                       recipient_names=[f"R{j}" for j in range(int(g["n_recipients"]))])
    ds = mg.HigherDataset(str(tmp_path), "2031", device=DEV)
    inter, city, prov = ds.get_adjacent()
    assert ds.get_adjacent()[0] is inter                                     # built once
    rowptr, col, val = O.csr_from_dense(g["inter"])                          # reference inter_adjacent (dense counts)
    assert np.array_equal(_np(inter.rowptr), rowptr) and np.array_equal(_np(inter.col), col)
    assert np.array_equal(_np(inter.val), val)
    assert (inter.n_rows, inter.n_cols) == g["inter"].shape == ds.get_count()
    assert np.array_equal(_np(city), g["city"]) and np.array_equal(_np(prov), g["province"])
    # the group-id vectors stand for the reference's (N, N) intra matrices
    gl = mg.GroupLists(city)
    rp, cl, rmap = (_np(x) for x in gl.lists())
    for i in (0, 11, 40):
        members = cl[rp[rmap[i]]:rp[rmap[i] + 1]]
        assert np.array_equal(np.sort(members), np.nonzero(g["city_adj"][i])[0])
    # model.normalize_adjacency_matrix: the reference result is all-NaN here (column M-1 has no record)
    norm = mg.normalize_adjacency_matrix(inter)
    assert torch.isnan(norm.val).all() and norm.col is inter.col


def test_normalize_adjacency_values():
    g = load_golden("gcn")
    graph = mg.normalize_adjacency_matrix(_t(g["adj"]))
    want = g["adj_norm"][g["adj"] > 0]
    np.testing.assert_allclose(_np(graph.val), want, rtol=1e-5)
    ids = _i([0, 1, 1])
    assert mg.normalize_adjacency_matrix(ids) is ids


@pytest.mark.parametrize("name,cls", [("ablation3", "ablation3"), ("ours", "Ours")])
def test_models_on_loader_output_match_reference(name, cls, tmp_path):
    """train.py:180-229 with the loader swapped in: files -> HigherDataset -> (Graph, city ids, province ids) -> model."""
    g = load_golden(name)
    p = params_of(g)
    N, M = g["adj"].shape
    r, c = np.nonzero(g["adj"] > 0)
    reps = np.maximum(g["adj"][r, c].astype(np.int64), 1)
    source, recipient = np.repeat(r, reps), np.repeat(c, reps)
    perm = np.random.default_rng(3).permutation(source.size)                 # records arrive in no particular order
    D.write_flow_files(str(tmp_path), "2032", source[perm], recipient[perm], g["city"], g["prov"], np.zeros(N),
                       recipient_names=[f"R{j}" for j in range(M)])
    ds = mg.HigherDataset(str(tmp_path), "2032", device=DEV)
    inter, city, prov = (mg.normalize_adjacency_matrix(a) for a in ds.get_adjacent())      # train.py:191-193
    Fin = p["Sfeatures"].shape[1]
    d = p["attention_0.W1"].shape[1]
    model = _load(getattr(mg, cls)(Fin, d, M, 2, 0.0, ds.get_gdp(), *ds.get_count()), p)
    model.train()
    src, rec = _i(g["src"]), _i(g["rec"])
    out = model(inter, city, prov, src)
    assert rel_err(_np(out), g["out"]) < TOL
    loss = torch.nn.functional.nll_loss(out[src], rec)
    assert abs(float(loss) - float(g["loss"])) < 1e-5


# ---------------------------------------------------------------------------------- 8f-1 LLP student + distillation
def test_mlp_golden():
    g = load_golden("mlp3")
    mlp = _load(mg.MLP(3, 12, 20, 7, 0.0), params_of(g))
    x = _t(g["x"], grad=True)
    out = mlp(x)
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    assert rel_err(_np(x.grad), g["gx"]) < TOL
    _check_grads(dict(mlp.named_parameters()), g)


@pytest.mark.parametrize("tag", ["c12", "c7"])
def test_kd_losses_golden(tag):
    g = load_golden("kd_losses_" + tag)
    s, t = _t(g["s"], grad=True), _t(g["t"], grad=True)
    i_s, i_t = _i(g["idx_s"]), _i(g["idx_t"])
    kd = Fn.kd_cosine(s, t, i_s, i_t)                                         # fused gathers, teacher detached
    assert rel_err(_np(kd), g["kd"]) < TOL
    kd.backward()
    zs, zt = 3, 5                                   # the all-zero rows: gradient ~ t / (eps |t|) ~ 1e8, compared apart
    keep_s = np.arange(g["s"].shape[0]) != zs
    keep_t = np.arange(g["t"].shape[0]) != zt
    assert rel_err(_np(s.grad)[keep_s], g["gs"][keep_s]) < TOL
    assert rel_err(_np(s.grad)[zs], g["gs"][zs]) < TOL
    assert t.grad is None
    # the reference call shape: rows gathered by the caller (LLP.py:236)
    s2 = _t(g["s"], grad=True)
    kd2 = mg.KD_cosine(s2[i_s], t[i_t])
    assert rel_err(_np(kd2), g["kd"]) < TOL
    kd2.backward()
    assert rel_err(_np(s2.grad)[keep_s], g["gs"][keep_s]) < TOL
    # both operands differentiable (the kernel's dt path)
    s3, t3 = _t(g["s"], grad=True), _t(g["t"], grad=True)
    full = Fn._KDCosine.apply(s3, t3, i_s, i_t, 1e-8)
    full.backward()
    assert rel_err(_np(s3.grad)[keep_s], g["gs_full"][keep_s]) < TOL
    assert rel_err(_np(t3.grad)[keep_t], g["gt_full"][keep_t]) < TOL
    assert rel_err(_np(t3.grad)[zt], g["gt_full"][zt]) < TOL
    a, b = _t(g["a"], grad=True), _t(g["b"], grad=True)
    mse = Fn.mse_loss(a, b)
    assert rel_err(_np(mse), g["mse"]) < TOL
    mse.backward()
    assert rel_err(_np(a.grad), g["ga"]) < TOL and rel_err(_np(b.grad), g["gb"]) < TOL
    with pytest.raises(ValueError):
        Fn.mse_loss(a, b[:-1])


def test_kd_losses_vs_oracle_at_scale():
    gen = torch.Generator().manual_seed(5)
    n_s, n_t, C, P = 5000, 3000, 256, 60_000
    s = torch.randn(n_s, C, generator=gen)
    t = torch.randn(n_t, C, generator=gen)
    i_s = torch.randint(0, n_s, (P,), generator=gen)
    i_t = torch.randint(0, n_t, (P,), generator=gen)
    sd = s.to(DEV).requires_grad_(True)
    kd = Fn.kd_cosine(sd, t.to(DEV), i_s.to(DEV), i_t.to(DEV))
    kd.backward()
    so = torch.tensor(s.numpy(), dtype=torch.float64, requires_grad=True)
    ref = O.kd_cosine(so, t.numpy(), i_s.numpy(), i_t.numpy())
    ref.backward()
    assert abs(float(kd) - float(ref)) < 1e-5
    assert rel_err(_np(sd.grad), so.grad.numpy()) < TOL
    n = 1_000_003
    a = torch.randn(n, generator=gen)
    b = torch.randn(n, generator=gen)
    ad = a.to(DEV).requires_grad_(True)
    mse = Fn.mse_loss(ad, b.to(DEV))
    mse.backward()
    assert rel_err(_np(mse), float(((a.double() - b.double()) ** 2).mean())) < 1e-6
    assert rel_err(_np(ad.grad), (2 * (a.double() - b.double()) / n).numpy()) < 1e-6


def test_llp_distill_step_golden():
    g = load_golden("llp_step")
    p = params_of(g)
    N, M = g["adj"].shape
    model = _load(LLP.MLP(2, M, M, M, 0.0), sub_params(p, "model."))
    predictor = _load(LLP.LinkPredictor("mlp", M, M, 1, 2, 0.0), sub_params(p, "predictor."))
    teacher = _load(LLP.GAT(n_features=M, n_classes=M, n_heads=2, dropout=0.0, gdp=None, N=N), sub_params(p, "teacher."))
    teacher_pred = _load(LLP.Teacher_LinkPredictor("mlp", M, M, 1, 2, 0.0), sub_params(p, "teacher_pred."))
    for m in (model, predictor, teacher, teacher_pred):
        m.train()
    feats = _t(g["features"])
    src, rec = _i(g["src"]), _i(g["rec"])
    # pieces first: student embedding, teacher embedding (LLP.GAT.forward(input, adj), LLP.py:163-168)
    assert rel_err(_np(model(feats)), g["h"]) < TOL
    assert rel_err(_np(teacher(feats, _t(g["adj_norm"]))), g["t_h"]) < TOL
    loss, parts = mg.llp_distill_loss(model, predictor, teacher, teacher_pred, feats, _t(g["adj_norm"]), src, rec)
    assert rel_err(_np(parts["label_loss"]), g["label_loss"]) < TOL
    assert rel_err(_np(parts["KD_cosine"]), g["kd_f"]) < TOL
    assert rel_err(_np(parts["mse_loss"]), g["kd_p"]) < TOL
    scale = 10.0 * abs(float(g["label_loss"])) + 0.1 * abs(float(g["kd_f"])) + 100.0 * abs(float(g["kd_p"]))
    assert abs(float(loss) - float(g["loss"])) < TOL * scale                 # the weighted terms cancel
    loss.backward()
    _check_grads(dict(model.named_parameters()), g, "model.")
    _check_grads(dict(predictor.named_parameters()), g, "predictor.")
    assert all(q.grad is None for q in teacher.parameters())                 # LLP.py:236-237 detach the teacher
    with pytest.raises(TypeError):
        teacher(_t(g["adj_norm"]))                                           # LLP.GAT has no features of its own


# ---------------------------------------------------------------------------------- 8f-3 baselines
@pytest.mark.parametrize("via", ["dense", "graph"])
def test_gcn_model_golden(via):
    g = load_golden("gcn_model")
    p = params_of(g)
    N, M = g["adj"].shape
    gdp = {str(i): 0.0 for i in range(N)}
    model = _load(mg.GCN(6, 5, M, 0.0, gdp, N), p)
    model.train()
    adj = _t(g["adj_norm"]) if via == "dense" else mg.normalize_adjacency_matrix(mg.Graph.from_dense(_t(g["adj"])))
    out = model(adj)
    assert out.shape == g["out"].shape and rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    named = {n: q for n, q in model.named_parameters() if not n.startswith("gc3")}
    _check_grads(named, g)
    assert model.gc3.weight.grad is None                                     # never applied (model.py:62-63)


def test_graphsage_golden():
    g = load_golden("graphsage")
    p = params_of(g)
    N, M = g["adj"].shape
    gdp = {str(i): float(v) for i, v in enumerate(g["gdp"])}
    model = _load(mg.GraphSAGE(7, M, 5, gdp, Scount=N), p)
    model.train()
    src = _i(g["src"])
    out = model(src, _t(g["adj_norm"]))
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    _check_grads(dict(model.named_parameters()), g)
    graph = mg.normalize_adjacency_matrix(mg.Graph.from_dense(_t(g["adj"])))
    assert rel_err(_np(model(src, graph)), g["out"]) < TOL


def test_csr_rows_mul_vs_dense():
    rng = np.random.default_rng(8)
    N, M, B = 300, 70, 1000
    adj = ((rng.random((N, M)) < 0.1) * rng.integers(1, 5, (N, M))).astype(np.float32)
    adj[17] = 0
    graph = mg.Graph.from_dense(_t(adj))
    src = rng.integers(0, N, B)
    src[3] = 17
    x = rng.standard_normal((B, M)).astype(np.float32)
    xd = _t(x, grad=True)
    out = Fn.csr_rows_mul(graph, xd, _i(src))
    assert np.array_equal(_np(out), adj[src] * x)                            # one product per element: bit-exact
    G = rng.standard_normal((B, M)).astype(np.float32)
    (out * _t(G)).sum().backward()
    assert np.array_equal(_np(xd.grad), adj[src] * G)
    out2 = Fn.csr_rows_mul(graph, _t(x[:N]), None)                           # no index: row b of adj
    assert np.array_equal(_np(out2), adj * x[:N])


# ---------------------------------------------------------------------------------- 8f-4 attention export
@pytest.mark.parametrize("head", [0, 2, -1])
def test_export_attention_matches_explainer(head):
    rng = np.random.default_rng(11)
    N, M, Fin, H, d = 120, 120, 16, 3, 8
    adj = (rng.random((N, M)) < 0.08).astype(np.float32)
    adj[5] = 0                                                               # isolated row: uniform 1/M over all columns
    adj[:, 9] = 0                                                            # a column only the isolated row reaches
    torch.manual_seed(4)
    conv = mg.GATConv(Fin, d, heads=H, concat=True).to(DEV)
    graph = mg.Graph.from_dense(_t(adj))
    _, alpha = conv(_t(rng.random((N, Fin))), graph, return_alpha=True)
    ex = mg.export_attention(graph, alpha, heads=H, head=head)
    dense = _np(mg.dense_attention(graph, alpha, H))                         # (H, N, M)
    if head >= 0:
        dense = dense[head]
    else:
        # same arithmetic as the kernel's head mean (sum in head order, times 1/H)
        acc = np.zeros_like(dense[0])
        for h in range(H):
            acc = acc + dense[h]
        dense = acc * np.float32(1.0 / H)
    rows_ref = O.explainer_argmax(dense)                                     # Explainer.py:25 interAttS
    cols_ref = O.explainer_argmax(dense.T)                                   # Explainer.py:26 interAttR
    assert _np(ex["row_argmax"]).tolist() == [r[0] for r in rows_ref]
    assert _np(ex["row_ties"]).tolist() == [len(r) for r in rows_ref]
    np.testing.assert_array_equal(_np(ex["row_max"]), dense.max(axis=1))
    assert int(ex["row_ties"][5]) == M
    col_arg, col_ties = _np(ex["col_argmax"]).tolist(), _np(ex["col_ties"]).tolist()
    for j in range(M):
        assert col_arg[j] == cols_ref[j][0] and col_ties[j] == len(cols_ref[j]), j
    np.testing.assert_array_equal(_np(ex["col_max"]), dense.max(axis=0))
    ei = _np(ex["edge_index"])
    got = np.zeros((N, M), dtype=np.float32)
    got[ei[0], ei[1]] = _np(ex["alpha"])
    np.testing.assert_allclose(got, dense, rtol=1e-6, atol=0)


def test_segment_argmax_empty_items():
    ptr = torch.tensor([0, 0, 3, 3, 5], dtype=torch.int32, device=DEV)
    w = _t([[0.5], [0.7], [0.7], [0.1], [0.1]])
    vmax, first, ties = Fn.segment_argmax(ptr, w, 1, 0)
    assert _np(vmax).tolist() == pytest.approx([0.0, 0.7, 0.0, 0.1])
    assert _np(first).tolist() == [-1, 1, -1, 3] and _np(ties).tolist() == [0, 2, 0, 2]
