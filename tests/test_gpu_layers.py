"""GPU parity of the layers against (i) the golden vectors produced by the unmodified reference classes
and (ii) the fp64 CPU oracle on larger random cases.  Tolerance (north star): rel. err <= 1e-4 in fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import msha_gnn_b200 as mg
from msha_gnn_b200 import functional as Fn
from conftest import load_golden, params_of, rel_err
from oracle import msha_oracle as O

DEV = "cuda:0"
TOL = 1e-4


def _np(t):
    return t.detach().cpu().numpy()


def _t(a, grad=False):
    t = torch.tensor(np.asarray(a), dtype=torch.float32, device=DEV)
    return t.requires_grad_(True) if grad else t


def _load(module, p):
    module.load_state_dict({k: torch.tensor(v) for k, v in p.items()})
    return module.to(DEV)


def _check_grads(named, g, names, tol=TOL):
    for n in names:
        got = named[n].grad
        ref = g["g." + n]
        if np.max(np.abs(ref)) < 1e-5:            # identically-zero gradient (unused parameter -> None is fine)
            assert got is None or np.max(np.abs(_np(got))) < 1e-5, n
        else:
            assert rel_err(_np(got), ref) < tol, (n, rel_err(_np(got), ref))


# ---------------------------------------------------------------------------------- a-1 / a-2
def test_graph_attention_layer_golden():
    g = load_golden("gal")
    layer = _load(mg.GraphAttentionLayer(12, 8, 0.0), params_of(g))
    x = _t(g["x"], grad=True)
    out = layer(x, _t(g["adj"]))
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    assert rel_err(_np(x.grad), g["gx"]) < TOL
    assert rel_err(_np(layer.W.grad), g["gW"]) < TOL
    assert layer.a.grad is not None and float(layer.a.grad.abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        layer(x, torch.ones(37, 9, device=DEV))          # adj.shape must equal (N, out_features)


@pytest.mark.parametrize("via,name,H", [("dense", "gat", 2), ("edge_index", "gat", 2), ("graph", "gat", 2),
                                        ("dense", "gat_h8", 8), ("graph", "gat_h8", 8)])
def test_gat_model_golden(via, name, H):
    """gat_h8: BASELINE.json configs[1]'s GAT(32, 32, 8 heads) -- golden from the unmodified GAT.py."""
    g = load_golden(name)
    N, M = g["adj"].shape
    gdp = {str(i): float(v) for i, v in enumerate(g["gdp"])}
    model = _load(mg.GAT(M, M, H, 0.0, gdp, N), params_of(g))
    model.train()
    adj = _t(g["adj"])
    if via == "edge_index":
        ei = torch.nonzero(adj > 0).t().contiguous()
        graph = mg.as_graph(ei, n_rows=N, n_cols=M)
        out = model(graph)
    elif via == "graph":
        out = model(mg.Graph.from_dense(adj))
    else:
        out = model(adj)
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    named = dict(model.named_parameters())
    _check_grads(named, g, ["features", "out_att.W"] + [f"attention_{i}.W" for i in range(H)])
    out2 = model(model.features, adj)                    # LLP.py:163 signature
    assert rel_err(_np(out2), g["out"]) < TOL


# ---------------------------------------------------------------------------------- a-3 / a-4
@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("intra", ["dense", "groups", "groups+hubs"])
def test_ours_layers_golden(variant, mode, intra, monkeypatch):
    if intra.endswith("+hubs"):         # hub-segment path on the bipartite graph (every recipient column is a hub)
        monkeypatch.setattr(mg.graph, "SEG_LIMIT", 3)
        intra = "groups"
    g = load_golden(f"ourslayer{variant}_{mode}")
    cls = {1: mg.OursLayer, 2: mg.OursLayer2, 3: mg.OursLayer3}[variant]
    layer = _load(cls(16, 8, 0.0), params_of(g))
    layer.train(mode == "train")
    S, R = _t(g["S"], grad=True), _t(g["R"], grad=True)
    if intra == "dense":
        city = _t(O.group_adjacency_dense(g["city"]))
        prov = _t(O.group_adjacency_dense(g["prov"]))
    else:
        city = torch.tensor(g["city"], device=DEV)
        prov = torch.tensor(g["prov"], device=DEV)
    src = torch.tensor(g["src"], device=DEV)
    out = layer(S, R, _t(g["adj"]), city, prov, src)
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    assert rel_err(_np(S.grad), g["gS"]) < TOL
    assert rel_err(_np(R.grad), g["gR"]) < TOL
    names = ["W1", "W2", "a", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"]
    if variant == 1:
        names += ["a3", "a4"]
    _check_grads(dict(layer.named_parameters()), g, names)
    if mode == "train":
        for k in ("bn1.running_mean", "bn1.running_var", "bn2.running_mean", "bn2.running_var"):
            assert rel_err(_np(layer.state_dict()[k]), g["after." + k]) < TOL, k
        assert int(layer.bn1.num_batches_tracked) == int(g["after.bn1.num_batches_tracked"])


def test_ours_record_coefficients_golden():
    g = load_golden("ours_record")
    layer = _load(mg.OursLayer(10, 4, 0.0), params_of(g)).eval()
    N = g["S"].shape[0]
    C3 = torch.zeros(N, N, device=DEV)
    C4 = torch.zeros(N, N, device=DEV)
    with torch.no_grad():
        out = layer(_t(g["S"]), _t(g["R"]), _t(g["adj"]), _t(O.group_adjacency_dense(g["city"])),
                    _t(O.group_adjacency_dense(g["prov"])), torch.tensor(g["src"], device=DEV), True, None, C3, C4)
    assert rel_err(_np(out), g["out"]) < TOL
    assert rel_err(_np(mg.last_attention["Coeff12"][0]), g["coeff12"]) < TOL
    assert rel_err(_np(C3), g["coeff3"]) < TOL and rel_err(_np(C4), g["coeff4"]) < TOL


@pytest.mark.parametrize("name,cls", [("ablation1", "ablation1"), ("ablation2", "ablation2"), ("ablation3", "ablation3"),
                                      ("ours", "Ours")])
def test_msha_models_golden(name, cls):
    g = load_golden(name)
    p = params_of(g)
    N, M = g["adj"].shape
    Fin = p["Sfeatures"].shape[1]
    d = (p.get("attention_0.W1", p.get("attention.W1"))).shape[1]
    gdp = {str(i): 0.0 for i in range(N)}
    model = _load(getattr(mg, cls)(Fin, d, M, 2, 0.0, gdp, N, M), p)
    model.train()
    src = torch.tensor(g["src"], device=DEV)
    rec = torch.tensor(g["rec"], device=DEV)
    out = model(_t(g["adj"]), _t(O.group_adjacency_dense(g["city"])), _t(O.group_adjacency_dense(g["prov"])), src)
    assert rel_err(_np(out), g["out"]) < TOL
    loss = torch.nn.functional.nll_loss(out[src], rec)              # train.py:229
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    names = [n for n, _ in model.named_parameters() if "bn3" not in n]
    _check_grads(dict(model.named_parameters()), g, names, tol=2e-4)


# ---------------------------------------------------------------------------------- a-6
def test_hgane_layer_golden():
    g = load_golden("hgane")
    p = params_of(g)
    Ns, Fin = p["source_embedding"].shape
    M = p["recipient_embedding"].shape[0]
    d = p["W1.weight"].shape[0]
    layer = _load(mg.HGANELayer(Fin, d, Ns, M, {str(i): 0.0 for i in range(Ns)}, dropout=0.0), p).train()
    out = layer(_t(g["adj_inter"]), _t(g["adj_intra"]), torch.tensor(g["src"], device=DEV))
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    names = ["source_embedding", "recipient_embedding", "W1.weight", "W2.weight", "a12.weight", "a3.weight",
             "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"]
    _check_grads(dict(layer.named_parameters()), g, names, tol=2e-4)
    # a batch row without inter neighbour -> NaN everywhere (0/0, HGANE.py:67-68)
    adj = g["adj_inter"].copy()
    adj[g["src"][0]] = 0
    with torch.no_grad():
        assert torch.isnan(layer(_t(adj), _t(g["adj_intra"]), torch.tensor(g["src"], device=DEV))).all()


# ---------------------------------------------------------------------------------- a-7
@pytest.mark.parametrize("tag,predictor,nl", [("mlp2", "mlp", 2), ("mlp3", "mlp", 3), ("inner", "inner", 2)])
@pytest.mark.parametrize("fused", [False, True, "tc"])
def test_link_predictor_golden(tag, predictor, nl, fused):
    g = load_golden("linkpred_" + tag)
    lp = _load(mg.LinkPredictor(predictor, 16, 24, 1, nl, 0.0), params_of(g))
    lp.fused = fused == "tc"                    # "tc": fused tcgen05 scorer kernel; True: pair-gather entry, unfused kernels
    h = _t(g["h"], grad=True)
    src, dst = torch.tensor(g["src"], device=DEV), torch.tensor(g["dst"], device=DEV)
    out = lp.forward_pairs(h, h, src, dst) if fused else lp(h[src], h[dst])
    assert out.shape == g["out"].shape
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    assert rel_err(_np(h.grad), g["gh"]) < TOL
    if predictor == "mlp":
        named = dict(lp.named_parameters())
        _check_grads(named, g, [f"lins.{i}.weight" for i in range(nl - 1)] + [f"lins.{i}.bias" for i in range(nl - 1)])
        assert named[f"lins.{nl-1}.weight"].grad is None      # the last Linear never runs (LLP.py:111)


def test_graph_convolution_golden():
    g = load_golden("gcn")
    gc = _load(mg.GraphConvolution(10, 5), params_of(g))
    graph = mg.Graph.from_dense(_t(g["adj"]))
    x = _t(g["x"], grad=True)
    out = gc(x, graph, values=graph.normalized_values())
    assert rel_err(_np(out), g["out"]) < TOL
    (out * _t(g["G"])).sum().backward()
    assert rel_err(_np(x.grad), g["gx"]) < TOL and rel_err(_np(gc.weight.grad), g["gw"]) < TOL
    assert rel_err(_np(gc.bias.grad), g["gb"]) < TOL
    out2 = gc(x, _t(g["adj_norm"]))                       # dense normalised adjacency, as train.py passes it
    assert rel_err(_np(out2), g["out"]) < TOL


# ---------------------------------------------------------------------------------- generic multi-head GAT
def test_generic_gat_golden():
    g = load_golden("generic_gat")
    H, Fin, d = g["W"].shape
    conv = mg.GATConv(Fin, d, heads=H, concat=True, activation=None).to(DEV)
    with torch.no_grad():
        conv.W.copy_(_t(np.concatenate([g["W"][h] for h in range(H)], axis=1)))
        conv.a_nbr.copy_(_t(np.stack([g["a"][h][:d, 0] for h in range(H)])))
        conv.a_self.copy_(_t(np.stack([g["a"][h][d:, 0] for h in range(H)])))
    graph = mg.Graph.from_dense(_t(g["adj"]))
    out, alpha = conv(_t(g["X"]), graph, return_alpha=True)
    N = g["X"].shape[0]
    dense = mg.dense_attention(graph, alpha, H)
    for h in range(H):
        assert rel_err(_np(dense[h]), g["alpha"][h]) < TOL
        assert rel_err(_np(out.view(N, H, d)[:, h]), g["agg"][h]) < TOL


@pytest.mark.parametrize("N,Fin,H,d,density,concat", [
    (257, 48, 8, 32, 0.08, True),       # C = 256: 128-bit path, 2 vectors / lane
    (300, 40, 4, 16, 0.10, True),       # C = 64
    (200, 33, 3, 6, 0.15, True),        # scalar path, H and D not powers of two
    (150, 20, 2, 64, 0.2, False),       # head mean, C = 128
    (180, 24, 1, 128, 0.1, True),       # one head spanning the whole warp
    (120, 16, 2, 12, 0.2, True),        # vector path, lanes-per-head = 3 (not a power of two)
])
@pytest.mark.parametrize("seg_limit", [None, 16])
def test_gat_conv_vs_oracle(N, Fin, H, d, density, concat, seg_limit, monkeypatch):
    if seg_limit:                       # force the hub-segment path: rows / columns above 16 entries are split and merged
        monkeypatch.setattr(mg.graph, "SEG_LIMIT", seg_limit)
    rng = np.random.default_rng(N)
    adj = (rng.random((N, N)) < density).astype(np.float32)
    adj[5] = 0                                            # isolated row -> uniform 1/N attention
    adj[9, :] = 1                                         # hub row
    x = rng.random((N, Fin)).astype(np.float32)
    torch.manual_seed(N)
    conv = mg.GATConv(Fin, d, heads=H, concat=concat).to(DEV)
    xg = _t(x, grad=True)
    out = conv(xg, mg.Graph.from_dense(_t(adj)))
    G = rng.standard_normal(out.shape).astype(np.float32)
    (out * _t(G)).sum().backward()
    rowptr, col, _ = O.csr_from_dense(adj)
    leaves = [torch.tensor(a, dtype=torch.float64, requires_grad=True)
              for a in (x, _np(conv.W), _np(conv.a_nbr), _np(conv.a_self))]
    ref = O.gat_layer(*leaves, rowptr, col, H, concat=concat)
    (ref * torch.tensor(G, dtype=torch.float64)).sum().backward()
    assert rel_err(_np(out), ref.detach().numpy()) < TOL
    for got, lf, name in zip((xg.grad, conv.W.grad, conv.a_nbr.grad, conv.a_self.grad), leaves, "x W a_nbr a_self".split()):
        assert rel_err(_np(got), lf.grad.numpy()) < TOL, name


def test_gat_conv_hub_row_at_the_production_segment_length(monkeypatch):
    """The segment length the large graphs run with (1 024, graph.SEG_LIMIT_MAX): one row and one column with 2 600 entries
    (3 segments each, the last one ragged), the cfg-3/4 layer shape (H = 8, d' = 32, C = 256), against the fp64 oracle
    (Ablation.py:262-274 arithmetic) -- output, input and parameter gradients."""
    monkeypatch.setattr(mg.graph, "SEG_LIMIT", 1024)
    rng = np.random.default_rng(1024)
    N, Fin, H, d = 2600, 32, 8, 32
    adj = (rng.random((N, N)) < 0.004).astype(np.float32)
    adj[np.arange(N), np.arange(N)] = 1
    adj[17, :] = 1                                        # hub row: 2 600 neighbours
    adj[:, 99] = 1                                        # hub column: referenced by every row
    x = rng.random((N, Fin)).astype(np.float32)
    torch.manual_seed(7)
    conv = mg.GATConv(Fin, d, heads=H, concat=True).to(DEV)
    xg = _t(x, grad=True)
    graph = mg.Graph.from_dense(_t(adj))
    assert graph.hub_rows().struct.seg_limit == 1024 and graph.hub_rows().n_segs == 3 and graph.hub_cols().n_segs == 3
    out = conv(xg, graph)
    G = rng.standard_normal(out.shape).astype(np.float32)
    (out * _t(G)).sum().backward()
    rowptr, col, _ = O.csr_from_dense(adj)
    leaves = [torch.tensor(a, dtype=torch.float64, requires_grad=True)
              for a in (x, _np(conv.W), _np(conv.a_nbr), _np(conv.a_self))]
    ref = O.gat_layer(*leaves, rowptr, col, H)
    (ref * torch.tensor(G, dtype=torch.float64)).sum().backward()
    assert rel_err(_np(out), ref.detach().numpy()) < TOL
    for got, lf, name in zip((xg.grad, conv.W.grad, conv.a_nbr.grad, conv.a_self.grad), leaves, "x W a_nbr a_self".split()):
        assert rel_err(_np(got), lf.grad.numpy()) < TOL, name


def test_attention_dropout_matches_injected_mask():
    """Training-mode attention dropout: the kernel's Philox mask, injected into the oracle, reproduces
    forward and gradients (the torch generator itself cannot be matched)."""
    rng = np.random.default_rng(3)
    N, M, Fin, d, p = 60, 7, 12, 8, 0.4
    adj = (rng.random((N, M)) < 0.4).astype(np.float32)
    adj[:, 0] = 1
    layer = mg.OursLayer3(Fin, d, p).to(DEV).train()
    S, R = _t(rng.random((N, Fin)), grad=True), _t(rng.random((M, Fin)), grad=True)
    torch.manual_seed(123)
    out = layer(S, R, _t(adj))
    # recover the mask through the public stream definition: element e*H+h of stream 2 under the block's seed
    blk = [f for f in _walk(out.grad_fn) if type(f).__name__ == "_AttentionBlockBackward"][0]
    rowptr, col, _ = O.csr_from_dense(adj)
    keep = O.dropout_keep_mask(blk.seed, col.size, p, stream=2)
    r = np.repeat(np.arange(N), np.diff(rowptr))
    mask = np.zeros((N, M)); mask[r, col] = keep / (1 - p)
    pp = {k: v.detach().cpu().numpy() for k, v in layer.state_dict().items()}
    So, Ro = (torch.tensor(_np(t), dtype=torch.float64, requires_grad=True) for t in (S, R))
    ref = O.ours_layer3(So, Ro, pp, rowptr, col, training=True, alpha_mask=mask)
    assert rel_err(_np(out), ref.detach().numpy()) < TOL
    out.sum().backward(); ref.sum().backward()
    assert rel_err(_np(S.grad), So.grad.numpy()) < TOL and rel_err(_np(R.grad), Ro.grad.numpy()) < TOL


def _walk(fn, seen=None):
    seen = seen if seen is not None else set()
    if fn is None or fn in seen:
        return
    seen.add(fn)
    yield fn
    for nf, _ in fn.next_functions:
        yield from _walk(nf, seen)


def test_nll_readout_train_step_matches_oracle_at_scale():
    """ablation3 step on a 2015-shaped synthetic graph (N ~ 4k, M = 32): loss and a parameter gradient."""
    rng = np.random.default_rng(7)
    N, M, Fin, d, H, B = 4000, 32, 128, 64, 2, 64
    k = rng.integers(1, 6, N)
    s = np.repeat(np.arange(N), k)
    dd = rng.integers(0, M, s.size)
    graph = mg.Graph.from_coo(torch.from_numpy(s).to(DEV), torch.from_numpy(dd).to(DEV), N, M)
    gdp = {str(i): float(v) for i, v in enumerate(rng.random(N))}
    torch.manual_seed(5)
    model = mg.ablation3(Fin, d, M, H, 0.0, gdp, N, M).to(DEV).train()
    src = torch.from_numpy(rng.integers(0, N, B)).to(DEV)
    rec = torch.from_numpy(rng.integers(0, M, B)).to(DEV)
    out = model(graph, None, None, src)
    loss = torch.nn.functional.nll_loss(out[src], rec)
    loss.backward()
    p = {k_: v.detach().cpu().numpy() for k_, v in model.state_dict().items()}
    # oracle with the *initial* running stats is irrelevant in training mode (batch statistics)
    leaves = {k_: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k_, v in p.items() if v.dtype.kind == "f"}
    from conftest import sub_params
    heads = [sub_params(leaves, f"attention_{i}.") for i in range(H)]
    rowptr, col, _ = O.csr_from_coo(s, dd, N, M)
    ref = O.msha_model(leaves["Sfeatures"], leaves["Rfeatures"], heads, (leaves["out_att.W"], leaves["out_att.a"]),
                       rowptr, col, training=True, variant=3)
    lref = O.nll_readout(ref, _np(src), _np(rec))
    lref.backward()
    assert rel_err(_np(out), ref.detach().numpy()) < TOL
    assert abs(float(loss) - float(lref)) < 1e-4 * abs(float(lref))
    for n in ("attention_0.W1", "attention_1.W2", "out_att.W", "Rfeatures"):
        assert rel_err(_np(dict(model.named_parameters())[n].grad), leaves[n].grad.numpy()) < 2e-4, n


def test_nll_readout_kernel_matches_torch():
    g = torch.Generator().manual_seed(3)
    for P, C in [(1000, 32), (777, 7), (50000, 256)]:
        logp = torch.log_softmax(torch.randn(P, C, generator=g), dim=1).to(DEV).requires_grad_(True)
        tgt = torch.randint(0, C, (P,), generator=g).to(DEV)
        loss = Fn.nll_loss(logp, tgt)
        (loss * 3.0).backward()
        ref_in = logp.detach().clone().requires_grad_(True)
        ref = torch.nn.functional.nll_loss(ref_in, tgt)
        (ref * 3.0).backward()
        assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
        assert torch.equal(logp.grad, ref_in.grad)


@pytest.mark.parametrize("fused_bwd", [True, False])
@pytest.mark.parametrize("P,C,Hd,Nn", [(1000, 256, 256, 300), (4097, 64, 128, 50), (129, 32, 40, 17), (70000, 256, 256, 4267),
                                      (5000, 128, 200, 64), (33, 8, 4, 5), (1, 16, 8, 3), (37889, 256, 256, 1000)])
def test_fused_scorer_vs_oracle(P, C, Hd, Nn, fused_bwd, monkeypatch):
    monkeypatch.setattr(Fn, "FUSED_SCORE_BWD", fused_bwd)
    g = torch.Generator().manual_seed(P)
    lp = mg.LinkPredictor("mlp", C, Hd, 1, 2, 0.0).to(DEV)
    h = (torch.randn(Nn, C, generator=g) * 0.5).to(DEV).requires_grad_(True)
    src = torch.randint(0, Nn, (P,), generator=g).to(DEV)
    dst = torch.randint(0, Nn, (P,), generator=g).to(DEV)
    out = lp.forward_pairs(h, h, src, dst)
    G = torch.randn(P, Hd, generator=g).to(DEV)
    (out * G).sum().backward()
    hd = h.detach().cpu().double().requires_grad_(True)
    W = [l.weight.detach().cpu().double().requires_grad_(True) for l in lp.lins]
    b = [l.bias.detach().cpu().double().requires_grad_(True) for l in lp.lins]
    ref = O.link_predictor(hd[src.cpu()], hd[dst.cpu()], W, b)
    assert rel_err(_np(out), ref.detach().numpy()) < TOL
    # relu'(x) is discontinuous at 0: among P*Hd pre-activations a handful land within fp32 round-off of 0, where the
    # active set is implementation-defined in fp32 (the reference included).  The gradient oracle therefore uses the
    # kernel's active set for those elements -- after checking that it differs from the exact one only at |pre| < 1e-6.
    pre = (hd[src.cpu()] * hd[dst.cpu()]) @ W[0].t() + b[0]
    mask = torch.from_numpy(_np(out) > 0.5)
    flips = mask != (pre.detach() > 0)
    assert int(flips.sum()) <= 64 and (not flips.any() or float(pre.detach()[flips].abs().max()) < 1e-6)

    class _Relu(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.clamp_min(0)

        @staticmethod
        def backward(ctx, g_):
            return g_ * mask

    ref2 = torch.sigmoid(_Relu.apply(pre))
    (ref2 * G.cpu().double()).sum().backward()
    assert rel_err(_np(h.grad), hd.grad.numpy()) < TOL
    assert rel_err(_np(lp.lins[0].weight.grad), W[0].grad.numpy()) < TOL
    assert rel_err(_np(lp.lins[0].bias.grad), b[0].grad.numpy()) < TOL


@pytest.mark.parametrize("P,C,Hd,Nn", [(1000, 256, 256, 300), (4097, 64, 128, 50), (70001, 256, 256, 4267), (33, 8, 4, 5)])
def test_fused_scorer_nll_matches_separate_ops(P, C, Hd, Nn):
    """nll read-out folded into the scorer backward (d scores generated in the dZ producers) == scorer + nll_loss ops."""
    g = torch.Generator().manual_seed(P + 1)
    lp = mg.LinkPredictor("mlp", C, Hd, 1, 2, 0.0).to(DEV)
    h = (torch.randn(Nn, C, generator=g) * 0.5).to(DEV)
    src = torch.randint(0, Nn, (P,), generator=g).to(DEV)
    dst = torch.randint(0, Nn, (P,), generator=g).to(DEV)
    tgt = torch.randint(0, min(Hd, 2), (P,), generator=g).to(DEV)
    h1 = h.clone().requires_grad_(True)
    loss1 = 3.0 * Fn.nll_loss(lp.forward_pairs(h1, h1, src, dst), tgt)
    loss1.backward()
    ref = [h1.grad.clone(), lp.lins[0].weight.grad.clone(), lp.lins[0].bias.grad.clone()]
    lp.zero_grad(set_to_none=True)
    h2 = h.clone().requires_grad_(True)
    loss2 = 3.0 * lp.nll_loss_pairs(h2, h2, src, dst, tgt)
    loss2.backward()
    assert abs(float(loss1) - float(loss2)) <= 1e-6 * max(1.0, abs(float(loss1)))
    got = [h2.grad, lp.lins[0].weight.grad, lp.lins[0].bias.grad]
    for a, b in zip(got, ref):
        assert rel_err(_np(a), _np(b)) < 1e-5


@pytest.mark.parametrize("P,C,Hd,Nn,labels", [(5000, 256, 256, 300, "random"), (4097, 64, 128, 50, "random"),
                                              (70001, 256, 256, 4267, "runs"), (1000, 128, 32, 77, "bad"),
                                              (257, 256, 256, 9, "one")])
def test_sparse_nll_backward_matches_dense_and_oracle(P, C, Hd, Nn, labels, monkeypatch):
    """The contraction-free backward (d scores is one-hot per row) == the tensor-core backward == the fp64 oracle, for labels
    spread over all Hd classes, in runs, partly out of range, and all equal."""
    g = torch.Generator().manual_seed(P + 7)
    lp = mg.LinkPredictor("mlp", C, Hd, 1, 2, 0.0).to(DEV)
    h = (torch.randn(Nn, C, generator=g) * 0.5)
    src = torch.randint(0, Nn, (P,), generator=g)
    dst = torch.randint(0, Nn, (P,), generator=g)
    if labels == "random":
        tgt = torch.randint(0, Hd, (P,), generator=g)
    elif labels == "runs":
        tgt = torch.cat([torch.ones(P // 2, dtype=torch.int64), torch.zeros(P - P // 2, dtype=torch.int64)])
    elif labels == "one":
        tgt = torch.full((P,), Hd - 1, dtype=torch.int64)
    else:
        tgt = torch.randint(0, Hd, (P,), generator=g)
        tgt[::7] = Hd + 3                                  # out of range: contributes nothing to any gradient
        tgt[3::11] = -1
    grads = {}
    for mode in ("sparse", "sparse_by_src", "dense"):
        monkeypatch.setattr(Fn, "SPARSE_NLL_BWD", mode != "dense")
        monkeypatch.setattr(Fn, "NLL_ORDER_BY_SRC", mode == "sparse_by_src")
        lp.zero_grad(set_to_none=True)
        hd = h.clone().to(DEV).requires_grad_(True)
        loss = 2.0 * lp.nll_loss_pairs(hd, hd, src.to(DEV), dst.to(DEV), tgt.to(DEV))
        loss.backward()
        grads[mode] = [_np(hd.grad), _np(lp.lins[0].weight.grad), _np(lp.lins[0].bias.grad)]
    for a, b, c in zip(grads["sparse"], grads["dense"], grads["sparse_by_src"]):
        assert rel_err(a, b) < 2e-5 and rel_err(c, b) < 2e-5
    # oracle (fp64): autograd through the restated scorer + read-out; bad labels masked out of the mean's numerator
    ho = torch.tensor(h.numpy(), dtype=torch.float64, requires_grad=True)
    W = [torch.tensor(_np(l.weight), dtype=torch.float64, requires_grad=True) for l in lp.lins]
    b = [torch.tensor(_np(l.bias), dtype=torch.float64, requires_grad=True) for l in lp.lins]
    out = O.link_predictor(ho[src], ho[dst], W, b)
    ok = (tgt >= 0) & (tgt < Hd)
    picked = out[torch.arange(P)[ok], tgt[ok]]
    (2.0 * -(picked.sum() / P)).backward()
    assert rel_err(grads["sparse"][0], ho.grad.numpy()) < TOL
    assert rel_err(grads["sparse"][1], W[0].grad.numpy()) < TOL
    assert rel_err(grads["sparse"][2], b[0].grad.numpy()) < TOL


def test_sparse_nll_backward_identity_order():
    """order == NULL walks the pairs as they come (labels already in runs)."""
    from msha_gnn_b200.ops import call, ptr, _stream
    g = torch.Generator().manual_seed(3)
    P, C, Hd, Nn = 3000, 256, 64, 40
    h = (torch.randn(Nn, C, generator=g) * 0.5).to(DEV)
    W0 = (torch.randn(Hd, C, generator=g) * 0.1).to(DEV)
    out = torch.rand(P, Hd, generator=g).to(DEV)                     # any saved activations in (0, 1)
    src = torch.randint(0, Nn, (P,), generator=g).to(DEV)
    dst = torch.randint(0, Nn, (P,), generator=g).to(DEV)
    tgt = torch.sort(torch.randint(0, Hd, (P,), generator=g)).values.to(DEV)
    gl = torch.ones(1, device=DEV)
    res = []
    for order in (None, Fn.nll_label_order(tgt, Hd)):
        dhi, dhj = torch.zeros_like(h), torch.zeros_like(h)
        dW, db = torch.empty_like(W0), torch.empty(Hd, device=DEV)
        call("msha_score_mlp_nll_bwd_sparse", ptr(order, torch.int32), ptr(tgt, torch.int64), ptr(gl), ptr(out), Hd, ptr(h),
             ptr(h), ptr(src, torch.int64), ptr(dst, torch.int64), P, C, ptr(W0), Hd, 3, 0.2, ptr(dhi), ptr(dhj), ptr(dW),
             ptr(db), _stream())
        res.append([_np(dhi), _np(dhj), _np(dW), _np(db)])
    assert np.array_equal(_np(Fn.nll_label_order(tgt, Hd)), np.arange(P))      # stable sort of sorted labels
    by_src = _np(Fn.nll_label_order(tgt, Hd, src, Nn)).astype(np.int64)          # (label, src, original position) order
    want_order = np.lexsort((np.arange(P), _np(src), _np(tgt)))
    assert np.array_equal(by_src, want_order)
    for a, b in zip(*res):
        assert rel_err(a, b) < 1e-5
    # closed form of db0: sum of g_p per label
    y = _np(out)[np.arange(P), _np(tgt)]
    gp = -(1.0 / P) * np.where(y > 0.5, y * (1 - y), 0.0)
    want = np.zeros(Hd)
    np.add.at(want, _np(tgt), gp)
    assert rel_err(res[0][3], want) < 1e-5
