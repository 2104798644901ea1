"""Peer-memory data path of the partitioned model (msha_gnn_b200/peer.py, dist.py) on ONE GPU: the ranks are emulated inside
this process (LocalFabric: a rank = its own CUDA stream + its own buffers; "peer" pointers are ordinary device pointers),
so the flag protocol, the pulls, the sums and the block-wise attention kernels are exactly what real ranks run.
Checked against the whole-graph result on one GPU and against the fp64 oracle (Ablation.py:262-274 arithmetic)."""
import copy
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # emulated ranks spin on one another: no false stream serialisation
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")          # (conftest.py sets these before the CUDA context exists)
os.environ.setdefault("MSHA_PEER_TIMEOUT_S", "20")

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import msha_gnn_b200 as mg
from msha_gnn_b200 import dist as md
from msha_gnn_b200 import dist_p2p as mp2p
from msha_gnn_b200 import functional as Fn
from msha_gnn_b200 import graph as mgraph
from msha_gnn_b200 import peer
from msha_gnn_b200.ops import call, ptr, _stream, LRELU_SLOPE
from conftest import rel_err

DEV = "cuda:0"
I32 = torch.int32


def _power_law_graph(N, seed, avg=8, cap=400):
    rng = np.random.default_rng(seed)
    deg = np.minimum(1 + (rng.pareto(1.1, N) * avg).astype(np.int64), cap)
    rows = np.repeat(np.arange(N), deg)
    cols = rng.integers(0, N, rows.size)
    key = np.unique(np.concatenate([rows * N + cols, np.arange(N) * N + np.arange(N)]))      # + self loops
    return key // N, key % N


def test_flags_pull_and_sum_on_emulated_ranks():
    W = 3
    fab = peer.LocalFabric(W, torch.device(DEV))
    n_max, C = 8, 12
    bufs = [g.alloc((W * n_max, C)) for g in fab.groups]
    for r in range(W):
        bufs[r].local.fill_(-1.0)
        bufs[r].local[r * n_max:(r + 1) * n_max] = float(r + 1)
    torch.cuda.synchronize()
    outs = []
    for r, g in enumerate(fab.groups):                       # every rank's work is only enqueued: nobody blocks the host
        with torch.cuda.stream(fab.streams[r]):
            g.signal(5, 1)
            g.wait(5, 1)
            g.pull_blocks_sm(bufs[r], n_max)
            g.barrier()
            out = torch.empty(n_max, C, device=DEV)
            g.sum_into(out, [bufs[r].addr[q] + r * n_max * C * 4 for q in range(W)], n_max * C)
            outs.append(out)
    torch.cuda.synchronize()
    for r in range(W):
        want = torch.cat([torch.full((n_max, C), float(q + 1)) for q in range(W)])
        assert torch.equal(bufs[r].local.cpu(), want)
        assert torch.equal(outs[r].cpu(), torch.full((n_max, C), float(W * (r + 1))))
        fab.groups[r].check()


@pytest.mark.parametrize("seg_limit", [0, 16])
def test_blockwise_forward_equals_fused_forward(seg_limit, monkeypatch):
    """msha_gat_softmax_stats + msha_gat_fwd_block over column blocks == msha_gat_fwd (alpha and aggregate)."""
    monkeypatch.setattr(mgraph, "SEG_LIMIT", seg_limit)
    N, H, D = 500, 8, 8
    C = H * D
    rows, cols = _power_law_graph(N, 3)
    g = mg.Graph.from_coo(torch.from_numpy(rows).to(DEV), torch.from_numpy(cols).to(DEV), N, N)
    torch.manual_seed(0)
    s_nbr, s_self = torch.randn(N, H, device=DEV), torch.randn(N, H, device=DEV)
    feat = torch.randn(N, C, device=DEV)
    ref_out, ref_alpha = Fn.attention_block(g, s_nbr, s_self, feat, heads=H)
    rp, col = g.attention_csr()
    lse = torch.empty(N, 2 * H, device=DEV)
    hub = g.hub_rows()
    assert hub.n_segs > 0                                      # adaptive limit (64) or 16: hub rows either way
    scr = torch.empty(max(1, 2 * H * hub.n_segs), device=DEV)
    call("msha_gat_softmax_stats", ptr(rp, I32), ptr(col, I32), N, ptr(s_nbr), ptr(s_self), H, LRELU_SLOPE, ptr(lse), hub.ptr,
         ptr(scr), _stream())
    # three column blocks, consumed out of order
    cuts = [0, 130, 377, N]
    key = torch.repeat_interleave(torch.arange(N, device=DEV), (rp[1:] - rp[:-1]).long()) * N + col.long()
    q = torch.arange(N, device=DEV).view(-1, 1) * N + torch.tensor(cuts, device=DEV).view(1, -1)
    blk = torch.searchsorted(key, q.reshape(-1)).view(N, 4).t().contiguous().to(I32)
    alpha = torch.zeros_like(ref_alpha)
    out = torch.full((N, C), 7.0, device=DEV)
    first = True
    for b in (1, 0, 2):
        bh = mgraph.Hub(beg=blk[b], end=blk[b + 1], seg_limit=mgraph.default_seg_limit(g.nnz))
        call("msha_gat_fwd_block", ptr(blk[b], I32), ptr(blk[b + 1], I32), ptr(col, I32), N, ptr(s_nbr), ptr(s_self), ptr(lse),
             ptr(feat), H, D, LRELU_SLOPE, ptr(alpha), ptr(out), 0 if first else 1, 0.0, 0, bh.ptr, _stream())
        first = False
    assert rel_err(alpha.cpu().numpy(), ref_alpha.detach().cpu().numpy()) < 2e-6
    assert rel_err(out.cpu().numpy(), ref_out.detach().cpu().numpy()) < 2e-6


def _run_ranks(fab, fn):
    """One host thread per emulated rank (like one process per GPU): a rank whose host blocks -- a device-to-host read, a
    pageable upload -- behind one of its own flag waits does not keep the others from enqueuing the kernels it waits for."""
    import threading
    errs = [None] * fab.world

    def body(r):
        try:
            with torch.cuda.stream(fab.streams[r]):
                fn(r)
            fab.streams[r].synchronize()
        except BaseException as e:      # noqa: BLE001
            errs[r] = e
    ts = [threading.Thread(target=body, args=(r,)) for r in range(fab.world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    torch.cuda.synchronize()
    for e in errs:
        if e is not None:
            raise e
    for g in fab.groups:
        g.check()


def _emulate(world, N, rows, cols, convs_ref, predictor_ref, x_full, src, dst, labels):
    """Runs one training step of the partitioned model on `world` emulated ranks; -> per-rank state (h, loss, grads)."""
    dev = torch.device(DEV)
    fab = peer.LocalFabric(world, dev)
    P = src.numel()
    state = [None] * world
    import threading
    together = threading.Barrier(world)

    def rank_step(r):
        part = md.Partition(N, world, r)
        keep = (rows >= part.lo) & (rows < part.hi)
        pgraph = md.partition_graph(torch.from_numpy(rows[keep]).to(dev), torch.from_numpy(cols[keep]).to(dev), part)
        st = dict(part=part, g=pgraph, p2p=md.P2P(fab.groups[r], part), convs=copy.deepcopy(convs_ref),
                  pred=copy.deepcopy(predictor_ref), x=x_full[part.lo:part.hi].clone().requires_grad_(True))
        lo, hi = (P * r) // world, (P * (r + 1)) // world
        st["h"] = md.gat_encode_p2p(st["convs"], st["x"], pgraph, part, st["p2p"], score_key="score_h")
        st["loss"] = md.score_pairs(st["pred"], st["h"], src[lo:hi], dst[lo:hi], part, target=labels[lo:hi],
                                    global_pairs=P, p2p=st["p2p"])
        together.wait(timeout=60)       # the ranks' backward passes share autograd's one device thread: enter it together
        st["loss"].backward()
        state[r] = st
    _run_ranks(fab, rank_step)
    return state


@pytest.mark.parametrize("world,pipelined", [(2, False), (3, False), (2, "rows"), (3, "rows"), (2, "halo"), (3, "halo"),
                                             (3, "halo-seq")])
def test_emulated_ranks_p2p_match_full_graph(world, pipelined, monkeypatch):
    """flat: fused exchange kernels; rows: copy-engine pulls of whole blocks under row chunks; halo: only the referenced rows
    (compact column numbering, msha_peer_gather_rows / msha_peer_scatter_add_rows)."""
    monkeypatch.setattr(mp2p, "PIPELINE_MIN_BLOCK_BYTES", 0 if pipelined else 1 << 40)
    monkeypatch.setattr(mp2p, "HALO", str(pipelined).startswith("halo"))
    monkeypatch.setattr(mp2p, "PIPELINE_CHUNKS", 1 if pipelined == "halo-seq" else 4)   # 1: sequential halo exchange
    monkeypatch.setattr(mgraph, "SEG_LIMIT", 32)               # hub rows and hub columns on every rank
    N, Fin, H, d, P = 403, 24, 4, 8, 3000
    rows, cols = _power_law_graph(N, 11 + world)
    rng = np.random.default_rng(5)
    torch.manual_seed(2)
    convs = torch.nn.ModuleList([mg.GATConv(Fin, d, H), mg.GATConv(H * d, d, H)]).to(DEV)
    pred = mg.LinkPredictor('mlp', H * d, 16, 1, 2, 0.0).to(DEV)
    x = torch.rand(N, Fin, device=DEV)
    src = torch.from_numpy(rng.integers(0, N, P)).to(DEV)
    dst = torch.from_numpy(rng.integers(0, N, P)).to(DEV)
    labels = torch.from_numpy(rng.integers(0, 2, P)).to(DEV)
    # ---- reference: whole graph on one GPU, global mean nll
    g_full = mg.Graph.from_coo(torch.from_numpy(rows).to(DEV), torch.from_numpy(cols).to(DEV), N, N)
    convs_f, pred_f = copy.deepcopy(convs), copy.deepcopy(pred)
    xf = x.clone().requires_grad_(True)
    hf = convs_f[1](convs_f[0](xf, g_full), g_full)
    loss_f = pred_f.nll_loss_pairs(hf, hf, src, dst, labels)
    loss_f.backward()
    # ---- emulated ranks
    state = _emulate(world, N, rows, cols, convs, pred, x, src, dst, labels)
    h = torch.cat([st["h"].detach() for st in state])
    assert rel_err(h.cpu().numpy(), hf.detach().cpu().numpy()) < 5e-6
    loss = sum(float(st["loss"]) for st in state)              # the ranks' shares add up to the global mean
    assert abs(loss - float(loss_f)) < 1e-5 * abs(float(loss_f))
    gx = torch.cat([st["x"].grad for st in state])
    assert rel_err(gx.cpu().numpy(), xf.grad.cpu().numpy()) < 1e-4
    for (name, pf) in list(convs_f.named_parameters()) + [("pred." + n, p) for n, p in pred_f.named_parameters()]:
        if pf.grad is None:
            continue
        tot = None
        for st in state:
            mod = st["pred"] if name.startswith("pred.") else st["convs"]
            p = dict(mod.named_parameters())[name[5:] if name.startswith("pred.") else name]
            tot = p.grad.clone() if tot is None else tot + p.grad
        assert rel_err(tot.cpu().numpy(), pf.grad.cpu().numpy()) < 1e-4, name


def test_second_step_reuses_buffers(monkeypatch):
    """Two consecutive steps through the same exchanges (reuse guards, sequence numbers) give the same result twice."""
    monkeypatch.setattr(mp2p, "PIPELINE_MIN_BLOCK_BYTES", 0)
    world, N, Fin, H, d = 2, 200, 16, 4, 8
    rows, cols = _power_law_graph(N, 1)
    torch.manual_seed(0)
    convs = torch.nn.ModuleList([mg.GATConv(Fin, d, H)]).to(DEV)
    x = torch.rand(N, Fin, device=DEV)
    dev = torch.device(DEV)
    fab = peer.LocalFabric(world, dev)
    res = [[None, None] for _ in range(world)]

    def rank_steps(r):
        part = md.Partition(N, world, r)
        keep = (rows >= part.lo) & (rows < part.hi)
        g = md.partition_graph(torch.from_numpy(rows[keep]).to(dev), torch.from_numpy(cols[keep]).to(dev), part)
        p2p = md.P2P(fab.groups[r], part)
        my_convs = copy.deepcopy(convs)
        xl = x[part.lo:part.hi].clone().requires_grad_(True)
        for step in range(2):
            out = md.gat_encode_p2p(my_convs, xl, g, part, p2p)
            out.square().sum().backward()
            res[r][step] = (out.detach().clone(), xl.grad.clone())
            xl.grad = None
    _run_ranks(fab, rank_steps)
    for r in range(world):
        assert torch.equal(res[r][0][0], res[r][1][0])
        assert rel_err(res[r][1][1].cpu().numpy(), res[r][0][1].cpu().numpy()) < 1e-6


@pytest.mark.parametrize("variant,world,halo", [(3, 2, True), (1, 2, True), (1, 3, True), (1, 3, False)])
def test_partitioned_msha_layer_matches_single_gpu(variant, world, halo, monkeypatch):
    """dist_msha.ours_encode on emulated ranks (gather of h1, reduce-scatter of alpha.T @ h2, BN statistics and intra-scale
    group tables all-reduced over peer memory) == layers.msha_heads_forward on one GPU: pair scores elu(u_i . v_j), and the
    gradients of the features and of every layer parameter.  (Ours.py:54-109 / Ablation.py:260-277.)"""
    from msha_gnn_b200 import dist_msha as dm
    from msha_gnn_b200.layers import msha_heads_forward
    monkeypatch.setattr(mp2p, "HALO", halo)              # halo: only the referenced recipients cross ranks (H = 2: 64-bit rows)
    dev = torch.device(DEV)
    rng = np.random.default_rng(variant * 10 + world)
    N, M, F, d, H, B, P = 301, 53, 16, 8, 2, 40, 500
    adj = (rng.random((N, M)) < 0.12)
    adj[np.arange(N), rng.integers(0, M, N)] = True                  # every source has a recipient
    rows, cols = np.nonzero(adj)
    city = rng.integers(0, 7, N)
    prov = city % 3
    torch.manual_seed(4)
    cls = mg.OursLayer if variant == 1 else mg.OursLayer3
    layers = torch.nn.ModuleList([cls(F, d, dropout=0.0) for _ in range(H)]).to(dev)
    S = torch.rand(N, F, device=dev)
    R = torch.rand(M, F, device=dev)
    src_b = torch.from_numpy(rng.integers(0, N, B)).to(dev)
    pi = torch.from_numpy(rng.integers(0, N, P)).to(dev)
    pj = torch.from_numpy(rng.integers(0, M, P)).to(dev)
    Gw = torch.from_numpy(rng.standard_normal((P, H)).astype(np.float32)).to(dev)
    city_d, prov_d = torch.from_numpy(city).to(dev), torch.from_numpy(prov).to(dev)
    # ---- single GPU
    g_full = mg.Graph.from_coo(torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev), N, M)
    ref_layers = copy.deepcopy(layers)
    Sf, Rf = S.clone().requires_grad_(True), R.clone().requires_grad_(True)
    out = msha_heads_forward(list(ref_layers), Sf, Rf, g_full, city_d if variant == 1 else None,
                             prov_d if variant == 1 else None, src_b if variant == 1 else None, True)      # (N, H*M)
    ref_scores = torch.stack([out[pi, h * M + pj] for h in range(H)], dim=1)
    (ref_scores * Gw).sum().backward()
    # ---- emulated ranks
    fab = peer.LocalFabric(world, dev)
    n3 = torch.bincount(city_d, minlength=7).float()
    n4 = torch.bincount(prov_d, minlength=3).float()
    res = [None] * world

    def rank_step(r):
        ps, pr = md.Partition(N, world, r), md.Partition(M, world, r)
        keep = (rows >= ps.lo) & (rows < ps.hi)
        pgraph = md.partition_graph(torch.from_numpy(rows[keep]).to(dev), torch.from_numpy(cols[keep]).to(dev), ps, col_part=pr)
        comm = dm.PeerComm({id(ps): mp2p.P2P(fab.groups[r], ps), id(pr): mp2p.P2P(fab.groups[r], pr)}, halo=halo)
        my = copy.deepcopy(layers)
        Sl = S[ps.lo:ps.hi].clone().requires_grad_(True)
        Rl = R[pr.lo:pr.hi].clone().requires_grad_(True)
        mine = (src_b >= ps.lo) & (src_b < ps.hi)
        u, v = dm.ours_encode(list(my), Sl, Rl, pgraph, ps, pr, comm, source_index=(src_b[mine] - ps.lo),
                              city_ids=city_d[ps.lo:ps.hi], province_ids=prov_d[ps.lo:ps.hi], group_sizes=(n3, n4))
        pm = (pi >= ps.lo) & (pi < ps.hi)
        sc = dm.score_uv_pairs(u, v, pi[pm] - ps.lo, pj[pm], pr, comm, heads=H)
        (sc * Gw[pm]).sum().backward()
        res[r] = dict(sc=sc.detach(), pm=pm, dS=Sl.grad, dR=Rl.grad, layers=my)
    _run_ranks(fab, rank_step)
    got = torch.empty_like(ref_scores)
    for r in range(world):
        got[res[r]["pm"]] = res[r]["sc"]
    assert rel_err(got.cpu().numpy(), ref_scores.detach().cpu().numpy()) < 2e-5
    assert rel_err(torch.cat([x["dS"] for x in res]).cpu().numpy(), Sf.grad.cpu().numpy()) < 1e-4
    assert rel_err(torch.cat([x["dR"] for x in res]).cpu().numpy(), Rf.grad.cpu().numpy()) < 1e-4
    for name, pf in ref_layers.named_parameters():
        if pf.grad is None:
            continue
        tot = sum(dict(x["layers"].named_parameters())[name].grad for x in res)
        assert rel_err(tot.cpu().numpy(), pf.grad.cpu().numpy()) < 2e-4, name
    for name, bf in ref_layers.named_buffers():                       # BN running statistics: global, identical on every rank
        if "num_batches" in name:
            continue
        for x in res:
            assert rel_err(dict(x["layers"].named_buffers())[name].cpu().numpy(), bf.cpu().numpy()) < 1e-5, name
