"""TEST INFRASTRUCTURE ONLY -- generate golden vectors by running the reference classes.

Run in the build container (where ``/root/reference`` is mounted):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference has no tests or golden files of its own, so every golden
vector is the output of the *unmodified* reference classes (imported / AST-extracted by
``oracle/ref_import.py``) on small seeded inputs, with dropout inactive (``dropout=0.0`` in
training mode so BatchNorm uses batch statistics, or ``.eval()``).  Gradients are those of the
scalar ``(output * G).sum()`` for a saved random ``G``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def _state(mod, prefix="p."):
    return {prefix + k: _np(v) for k, v in mod.state_dict().items()}


def _rand_adj(N, M, density, gen, iso_rows=(), counts=False):
    adj = (torch.rand(N, M, generator=gen) < density).float()
    if counts:
        adj = adj * torch.randint(1, 5, (N, M), generator=gen).float()
    for i in range(N):          # every non-isolated row gets >= 1 neighbour
        if adj[i].sum() == 0:
            adj[i, int(torch.randint(0, M, (1,), generator=gen))] = 1.0
    for i in iso_rows:
        adj[i] = 0.0
    return adj


def _grads(out, G, tensors):
    loss = (out * G).sum()
    gs = torch.autograd.grad(loss, tensors, allow_unused=True)
    return [torch.zeros_like(t) if g is None else g for g, t in zip(gs, tensors)]


def case_gal(save):
    GATm = ref_import.import_module("GAT")
    gen = torch.Generator().manual_seed(101)
    N, M, Fin = 37, 8, 12
    torch.manual_seed(1)
    layer = GATm.GraphAttentionLayer(Fin, M, dropout=0.0)
    x = torch.rand(N, Fin, generator=gen, requires_grad=True)
    adj = _rand_adj(N, M, 0.3, gen, iso_rows=(5, 20), counts=True)
    out = layer(x, adj)
    G = torch.randn(N, M, generator=gen)
    gx, gW, ga = _grads(out, G, [x, layer.W, layer.a])
    save("gal", dict(x=_np(x), adj=_np(adj), G=_np(G), out=_np(out), gx=_np(gx), gW=_np(gW),
                     ga=_np(ga), **_state(layer)))


def case_gat(save):
    GATm = ref_import.import_module("GAT")
    gen = torch.Generator().manual_seed(102)
    N, M, H = 33, 8, 2
    gdp = {str(i): float(v) for i, v in enumerate(torch.rand(N, generator=gen))}
    torch.manual_seed(2)
    model = GATm.GAT(n_features=M, n_classes=M, n_heads=H, dropout=0.0, gdp=gdp, N=N)
    adj = _rand_adj(N, M, 0.35, gen, iso_rows=(7,))
    model.train()
    out = model(adj)
    G = torch.randn(N, M, generator=gen)
    params = list(model.parameters())
    grads = _grads(out, G, params)
    d = dict(adj=_np(adj), G=_np(G), out=_np(out), gdp=np.array(list(gdp.values()), dtype=np.float64),
             **_state(model))
    for (name, _), g in zip(model.named_parameters(), grads):
        d["g." + name] = _np(g)
    save("gat", d)


def case_gat_h8(save):
    """BASELINE.json configs[1]: GAT(n_features=32, n_classes=32, n_heads=8) -- the head count and widths train.py:199 /
    LLP.py:290-295 use (the `gat` case above has 2 heads of width 8)."""
    GATm = ref_import.import_module("GAT")
    gen = torch.Generator().manual_seed(112)
    N, M, H = 67, 32, 8
    gdp = {str(i): float(v) for i, v in enumerate(torch.rand(N, generator=gen))}
    torch.manual_seed(12)
    model = GATm.GAT(n_features=M, n_classes=M, n_heads=H, dropout=0.0, gdp=gdp, N=N)
    adj = _rand_adj(N, M, 0.08, gen, iso_rows=(5, 40))
    model.train()
    out = model(adj)
    G = torch.randn(N, M, generator=gen)
    grads = _grads(out, G, list(model.parameters()))
    d = dict(adj=_np(adj), G=_np(G), out=_np(out), gdp=np.array(list(gdp.values()), dtype=np.float64), **_state(model))
    for (name, _), g in zip(model.named_parameters(), grads):
        d["g." + name] = _np(g)
    save("gat_h8", d)


def _msha_inputs(gen, N, M, Fin, iso_rows=(3,)):
    S = torch.rand(N, Fin, generator=gen)
    R = torch.rand(M, Fin, generator=gen)
    adj = _rand_adj(N, M, 0.3, gen, iso_rows=iso_rows, counts=True)
    city = torch.randint(0, 6, (N,), generator=gen)
    prov = city // 2
    city_adj = (city[:, None] == city[None, :]).float()
    prov_adj = (prov[:, None] == prov[None, :]).float()
    return S, R, adj, city, prov, city_adj, prov_adj


def _perturb_bn(layer, gen):
    # non-trivial affine / running stats so that eval-mode parity is meaningful
    for bn in (layer.bn1, layer.bn2):
        bn.weight.data = 0.5 + torch.rand(bn.weight.shape, generator=gen)
        bn.bias.data = torch.rand(bn.bias.shape, generator=gen) - 0.5
        bn.running_mean.data = torch.rand(bn.running_mean.shape, generator=gen) - 0.5
        bn.running_var.data = 0.5 + torch.rand(bn.running_var.shape, generator=gen)


def case_ours_layers(save):
    Ab = ref_import.import_module("Ablation")
    for variant, cls in ((1, Ab.OursLayer), (2, Ab.OursLayer2), (3, Ab.OursLayer3)):
        for mode in ("train", "eval"):
            gen = torch.Generator().manual_seed(200 + variant)
            N, M, Fin, d = 41, 8, 16, 8
            S, R, adj, city, prov, city_adj, prov_adj = _msha_inputs(gen, N, M, Fin)
            src = torch.tensor([0, 4, 9, 4, 17, 40, 3])      # duplicate 4; isolated row 3
            torch.manual_seed(10 + variant)
            layer = cls(Fin, d, dropout=0.0)
            _perturb_bn(layer, gen)
            layer.train(mode == "train")
            state0 = _state(layer)
            S.requires_grad_(True)
            R.requires_grad_(True)
            out = layer(S, R, adj, city_adj, prov_adj, src)
            G = torch.randn(N, M, generator=gen)
            names = ["W1", "W2", "a", "a3", "a4", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"]
            tens = [dict(layer.named_parameters())[n] for n in names]
            grads = _grads(out, G, [S, R] + tens)
            dd = dict(S=_np(S), R=_np(R), adj=_np(adj), city=_np(city), prov=_np(prov), src=_np(src),
                      G=_np(G), out=_np(out), gS=_np(grads[0]), gR=_np(grads[1]), **state0)
            for n, g in zip(names, grads[2:]):
                dd["g." + n] = _np(g)
            for k, v in layer.state_dict().items():           # running stats after the call
                if "running" in k or "num_batches" in k:
                    dd["after." + k] = _np(v)
            save(f"ourslayer{variant}_{mode}", dd)


def case_ours_record(save):
    cls = ref_import.ours_classes()
    gen = torch.Generator().manual_seed(301)
    N, M, Fin, d = 29, 6, 10, 4
    S, R, adj, city, prov, city_adj, prov_adj = _msha_inputs(gen, N, M, Fin, iso_rows=())
    src = torch.tensor([1, 5, 8, 13, 21])
    torch.manual_seed(31)
    layer = cls["OursLayer"](Fin, d, dropout=0.0)
    layer.eval()
    C3 = torch.zeros(N, N)
    C4 = torch.zeros(N, N)
    with torch.no_grad():
        out = layer(S, R, adj, city_adj, prov_adj, src, True, None, C3, C4)
    save("ours_record", dict(S=_np(S), R=_np(R), adj=_np(adj), city=_np(city), prov=_np(prov),
                             src=_np(src), out=_np(out), coeff12=_np(cls["train_stub"].Coeff12new),
                             coeff3=_np(C3), coeff4=_np(C4), **_state(layer)))


def case_msha_models(save):
    Ab = ref_import.import_module("Ablation")
    Ou = ref_import.ours_classes()
    for name, cls in (("ablation1", Ab.ablation1), ("ablation2", Ab.ablation2),
                      ("ablation3", Ab.ablation3), ("ours", Ou["Ours"])):
        gen = torch.Generator().manual_seed(400 + len(name))
        N, M, Fin, d, H = 35, 6, 12, 8, 2
        _, _, adj, city, prov, city_adj, prov_adj = _msha_inputs(gen, N, M, Fin, iso_rows=(2,))
        gdp = {str(i): float(v) for i, v in enumerate(torch.rand(N, generator=gen))}
        src = torch.tensor([0, 6, 6, 11, 30, 2])
        rec = torch.randint(0, M, (src.numel(),), generator=gen)
        torch.manual_seed(41)
        model = cls(in_features=Fin, out_features=d, n_classes=M, n_heads=H, dropout=0.0, gdp=gdp,
                    Scount=N, Rcount=M)
        model.train()
        state0 = _state(model)
        out = model(adj, city_adj, prov_adj, src)
        loss = torch.nn.functional.nll_loss(out[src], rec)          # train.py:229
        params = list(model.named_parameters())
        grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
        dd = dict(adj=_np(adj), city=_np(city), prov=_np(prov), src=_np(src), rec=_np(rec),
                  out=_np(out), loss=_np(loss), **state0)
        for (n, p), g in zip(params, grads):
            dd["g." + n] = _np(torch.zeros_like(p) if g is None else g)
        save(name, dd)


def case_hgane(save):
    Hg = ref_import.import_module("HGANE")
    gen = torch.Generator().manual_seed(501)
    Ns, M, Fin, d = 30, 7, 10, 6
    gdp = {str(i): float(v) for i, v in enumerate(torch.rand(Ns, generator=gen))}
    torch.manual_seed(51)
    layer = Hg.GraphAttentionLayer(Fin, d, Ns, M, gdp, dropout=0.0)
    adj_inter = _rand_adj(Ns, M, 0.4, gen)
    grp = torch.randint(0, 4, (Ns,), generator=gen)
    adj_intra = (grp[:, None] == grp[None, :]).float()
    src = torch.tensor([2, 3, 5, 7, 11, 13, 17, 19, 23, 29])
    layer.train()
    state0 = _state(layer)
    out = layer(adj_inter, adj_intra, src)
    G = torch.randn(out.shape, generator=gen)
    params = list(layer.named_parameters())
    grads = _grads(out, G, [p for _, p in params])
    dd = dict(adj_inter=_np(adj_inter), adj_intra=_np(adj_intra), src=_np(src), G=_np(G),
              out=_np(out), **state0)
    for (n, _), g in zip(params, grads):
        dd["g." + n] = _np(g)
    save("hgane", dd)


def case_linkpred(save):
    L = ref_import.llp_classes()
    for tag, predictor, C, Hd, nl in (("mlp2", "mlp", 16, 24, 2), ("mlp3", "mlp", 16, 24, 3),
                                      ("inner", "inner", 16, 24, 2)):
        gen = torch.Generator().manual_seed(600 + nl + len(tag))
        torch.manual_seed(61)
        lp = L["LinkPredictor"](predictor, C, Hd, 1, nl, 0.0)
        P, Nn = 50, 19
        h = torch.randn(Nn, C, generator=gen, requires_grad=True)
        si = torch.randint(0, Nn, (P,), generator=gen)
        di = torch.randint(0, Nn, (P,), generator=gen)
        out = lp(h[si], h[di])
        G = torch.randn(out.shape, generator=gen)
        params = list(lp.named_parameters())
        grads = _grads(out, G, [h] + [p for _, p in params])
        dd = dict(h=_np(h), src=_np(si), dst=_np(di), G=_np(G), out=_np(out), gh=_np(grads[0]),
                  **_state(lp))
        for (n, _), g in zip(params, grads[1:]):
            dd["g." + n] = _np(g)
        save("linkpred_" + tag, dd)


def case_gcn(save):
    Mo = ref_import.import_module("model")
    gen = torch.Generator().manual_seed(701)
    N, M, Fin, Fo = 31, 9, 10, 5
    torch.manual_seed(71)
    gc = Mo.GraphConvolution(Fin, Fo)
    adj = _rand_adj(N, M, 0.3, gen, counts=True)
    adj[:, 0] += 1.0                                     # no empty column -> finite normalisation
    adj_n = Mo.normalize_adjacency_matrix(adj)
    x = torch.rand(N, Fin, generator=gen, requires_grad=True)
    out = gc(x, adj_n)
    G = torch.randn(out.shape, generator=gen)
    gx, gw, gb = _grads(out, G, [x, gc.weight, gc.bias])
    adj_zero = adj.clone()
    adj_zero[:, 3] = 0.0
    save("gcn", dict(x=_np(x), adj=_np(adj), adj_norm=_np(adj_n), G=_np(G), out=_np(out),
                     gx=_np(gx), gw=_np(gw), gb=_np(gb),
                     adj_zero_norm=_np(Mo.normalize_adjacency_matrix(adj_zero)), adj_zero=_np(adj_zero),
                     **_state(gc)))


def case_generic_gat(save):
    """Generic H-head GAT layer == OursLayer3's attention block with S = R, per head
    (Ablation.py:262-271,274).  The reference output used is alpha (dense) and alpha @ h1."""
    Ab = ref_import.import_module("Ablation")
    gen = torch.Generator().manual_seed(801)
    N, Fin, d, H = 40, 12, 8, 3
    X = torch.rand(N, Fin, generator=gen)
    adj = _rand_adj(N, N, 0.2, gen, iso_rows=(9,))
    alphas, aggs, Ws, As = [], [], [], []
    for h in range(H):
        torch.manual_seed(80 + h)
        layer = Ab.OursLayer3(Fin, d, dropout=0.0)
        # replay Ablation.py:262-270 through the module's own tensors (S = R = X, W1 == W2)
        layer.W2.data.copy_(layer.W1.data)
        captured = {}
        orig_softmax = torch.nn.functional.softmax

        def spy(inp, dim=None, **kw):
            o = orig_softmax(inp, dim=dim, **kw)
            captured["alpha"] = o
            return o
        torch.nn.functional.softmax = spy
        try:
            layer.train()
            layer(X, X, adj, None, None, None)
        finally:
            torch.nn.functional.softmax = orig_softmax
        alpha = captured["alpha"]
        h1 = X @ layer.W1
        alphas.append(_np(alpha))
        aggs.append(_np(alpha @ h1))
        Ws.append(_np(layer.W1))
        As.append(_np(layer.a))
    save("generic_gat", dict(X=_np(X), adj=_np(adj), alpha=np.stack(alphas), agg=np.stack(aggs),
                             W=np.stack(Ws), a=np.stack(As)))


def case_dataset(save):
    """dataset.HigherDataset.intra_adjacent / inter_adjacent (dataset.py:260-296) on a small synthetic year: the class
    body is executed verbatim (AST), an instance is made without running the hard-coded-path __init__, and ``open`` is
    shadowed inside the class' globals so the ``inter<year>.json`` dump of dataset.py:293-294 goes to a null sink."""
    import contextlib
    import csv
    import io
    import json
    from torch.utils.data import Dataset

    @contextlib.contextmanager
    def null_open(*a, **k):
        yield io.StringIO()

    cls = ref_import.extract_classes("dataset.py", ["HigherDataset"], extra_globals=dict(
        Dataset=Dataset, json=json, csv=csv, year="0000", open=null_open))["HigherDataset"]
    rng = np.random.default_rng(901)
    N, M, R = 57, 6, 400
    city = rng.integers(0, 9, N)
    province = city // 3                                   # cities nest in provinces, like the real tables
    source = rng.integers(0, N, R)
    source[source == 11] = 12                              # node 11 has no record (isolated inter row)
    recipient = rng.integers(0, M - 1, R)                  # column M-1 never occurs (empty column)
    ds = object.__new__(cls)
    ds.graph_dict = {str(i): [int(i % 7), int(city[i]), int(province[i])] for i in range(N)}   # values[1], values[2]
    ds.N, ds.M, ds.count = N, M, R
    ds.source, ds.recipient = source.tolist(), recipient.tolist()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                    # torch.tensor(tensor) copy-construct warning, dataset.py:277
        city_adj, prov_adj = ds.intra_adjacent()
        inter = ds.inter_adjacent()
    gdp = rng.random(N)
    save("dataset", dict(source=source, recipient=recipient, city=city, province=province, gdp=gdp,
                         inter=_np(inter), city_adj=_np(city_adj), province_adj=_np(prov_adj),
                         n_recipients=np.int64(M)))


def _named_grads(out_scalar, mods):
    named = [(f"{tag}.{n}", p) for tag, m in mods for n, p in m.named_parameters()]
    gs = torch.autograd.grad(out_scalar, [p for _, p in named], allow_unused=True)
    return {"g." + n: _np(torch.zeros_like(p) if g is None else g) for (n, p), g in zip(named, gs)}


def case_llp(save):
    """One LLP training-step loss (LLP.py:230-237) from the reference's own MLP / LinkPredictor / GAT(input, adj) /
    Teacher_LinkPredictor / KD_cosine, dropout 0; plus a 3-layer MLP on its own and KD_cosine / MSELoss on raw rows."""
    L = ref_import.llp_classes()
    Mo = ref_import.import_module("model")
    gen = torch.Generator().manual_seed(1001)
    N, M, P = 41, 8, 30                                   # hidden_channels == Rcount (LLP.py:291-293)
    torch.manual_seed(91)
    model = L["MLP"](2, M, M, M, 0.0)
    predictor = L["LinkPredictor"]("mlp", M, M, 1, 2, 0.0)
    teacher = L["GAT"](n_features=M, n_classes=M, n_heads=2, dropout=0.0, gdp=None, N=N)
    teacher_pred = L["Teacher_LinkPredictor"]("mlp", M, M, 1, 2, 0.0)
    adj = _rand_adj(N, M, 0.3, gen, iso_rows=(4,), counts=True)
    adj[:, 0] += 1.0                                      # no empty column -> finite normalisation
    adj_n = Mo.normalize_adjacency_matrix(adj)
    features = torch.rand(N, M, generator=gen)
    si = torch.randint(0, N, (P,), generator=gen)
    ri = torch.randint(0, M, (P,), generator=gen)
    mse_loss = torch.nn.MSELoss()
    h = model(features)
    t_h = teacher(features, adj_n)
    output = predictor(h[si], h[ri]).squeeze()
    label_loss = torch.nn.functional.nll_loss(output, ri)
    t_out = teacher_pred(t_h[si], t_h[ri]).squeeze().detach()
    kd_f = L["KD_cosine"](h[si], t_h[si])
    kd_p = mse_loss(output, t_out)
    loss = 10.0 * label_loss + 0.1 * kd_f + 100.0 * kd_p
    mods = (("model", model), ("predictor", predictor))
    dd = dict(adj=_np(adj), adj_norm=_np(adj_n), features=_np(features), src=_np(si), rec=_np(ri), loss=_np(loss),
              label_loss=_np(label_loss), kd_f=_np(kd_f), kd_p=_np(kd_p), h=_np(h), t_h=_np(t_h), output=_np(output),
              t_out=_np(t_out), **_named_grads(loss, mods))
    for tag, m in mods + (("teacher", teacher), ("teacher_pred", teacher_pred)):
        dd.update(_state(m, prefix=f"p.{tag}."))
    save("llp_step", dd)

    torch.manual_seed(92)
    mlp = L["MLP"](3, 12, 20, 7, 0.0)
    x = torch.randn(33, 12, generator=gen, requires_grad=True)
    out = mlp(x)
    G = torch.randn(out.shape, generator=gen)
    params = list(mlp.named_parameters())
    grads = _grads(out, G, [x] + [p for _, p in params])
    dd = dict(x=_np(x), G=_np(G), out=_np(out), gx=_np(grads[0]), **_state(mlp))
    for (n, _), g in zip(params, grads[1:]):
        dd["g." + n] = _np(g)
    save("mlp3", dd)

    # KD_cosine / MSELoss on raw operands: duplicate indices, a zero row on each side, C not a multiple of 4
    for tag, C in (("c12", 12), ("c7", 7)):
        s = torch.randn(19, C, generator=gen)
        t = torch.randn(23, C, generator=gen)
        s[3] = 0.0
        t[5] = 0.0
        s.requires_grad_(True)
        t.requires_grad_(True)
        i_s = torch.randint(0, 19, (40,), generator=gen)
        i_t = torch.randint(0, 23, (40,), generator=gen)
        i_s[0], i_t[0] = 3, 1                             # zero student row
        i_s[1], i_t[1] = 2, 5                             # zero teacher row
        cos = L["KD_cosine"](s[i_s], t[i_t])
        (gs,) = torch.autograd.grad(cos, [s])
        full = 1 - torch.nn.functional.cosine_similarity(s[i_s], t[i_t], dim=-1).mean()   # teacher not detached
        gs_full, gt_full = torch.autograd.grad(full, [s, t])
        a = torch.randn(17, C, generator=gen, requires_grad=True)
        b = torch.randn(17, C, generator=gen, requires_grad=True)
        mse = mse_loss(a, b)
        ga, gb = torch.autograd.grad(mse, [a, b])
        save("kd_losses_" + tag, dict(s=_np(s), t=_np(t), idx_s=_np(i_s), idx_t=_np(i_t), kd=_np(cos), gs=_np(gs),
                                      gs_full=_np(gs_full), gt_full=_np(gt_full), a=_np(a), b=_np(b), mse=_np(mse),
                                      ga=_np(ga), gb=_np(gb)))


def case_baselines(save):
    """model.GCN (model.py:48-64) and SGAE.GraphSAGE (SGAE.py:41-56), dropout 0."""
    Mo = ref_import.import_module("model")
    gen = torch.Generator().manual_seed(1101)
    N, M, nfeat, nhid = 31, 9, 6, 5
    gdp = {str(i): float(v) for i, v in enumerate(torch.rand(N, generator=gen))}
    torch.manual_seed(93)
    gcn = Mo.GCN(nfeat, nhid, M, 0.0, gdp, N)
    adj = _rand_adj(N, M, 0.3, gen, iso_rows=(6,), counts=True)
    adj[:, 0] += 1.0
    adj[6, 0] = 0.0                                       # keep row 6 isolated
    adj_n = Mo.normalize_adjacency_matrix(adj)
    gcn.train()
    out = gcn(adj_n)
    G = torch.randn(out.shape, generator=gen)
    params = list(gcn.named_parameters())
    grads = _grads(out, G, [p for _, p in params])
    dd = dict(adj=_np(adj), adj_norm=_np(adj_n), G=_np(G), out=_np(out), **_state(gcn))
    for (n, _), g in zip(params, grads):
        dd["g." + n] = _np(g)
    save("gcn_model", dd)

    S = ref_import.sgae_classes(N)
    torch.manual_seed(94)
    sage = S["GraphSAGE"](7, M, 5, gdp)
    src = torch.randint(0, N, (20,), generator=gen)
    src[0] = 6                                            # an isolated source: its row of adj is all zero
    sage.train()
    out = sage(src, adj_n)
    G = torch.randn(out.shape, generator=gen)
    params = list(sage.named_parameters())
    grads = _grads(out, G, [p for _, p in params])
    dd = dict(adj=_np(adj), adj_norm=_np(adj_n), src=_np(src), G=_np(G), out=_np(out), gdp=np.array(list(gdp.values())),
              **_state(sage))
    for (n, _), g in zip(params, grads):
        dd["g." + n] = _np(g)
    save("graphsage", dd)


def main():
    os.makedirs(OUT, exist_ok=True)

    def save(name, d):
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"wrote {path}  ({os.path.getsize(path)} B, {len(d)} arrays)")

    torch.set_num_threads(1)
    for fn in (case_gal, case_gat, case_ours_layers, case_ours_record, case_msha_models, case_hgane,
               case_linkpred, case_gcn, case_generic_gat, case_dataset, case_llp, case_baselines, case_gat_h8):
        fn(save)


if __name__ == "__main__":
    main()
