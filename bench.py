#!/usr/bin/env python
"""Benchmark of the MSHA-GNN hot path (GAT message passing fwd+bwd + link scoring) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ddi|rmat|flow2015] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Definitions (identical for the GPU and the CPU arm):
  * step      = one training step of the link-prediction model on the whole graph: L GAT layers forward,
                scoring of P = E positive + E uniform negative pairs with the LinkPredictor MLP, loss,
                backward through everything, Adam update (train.py:221-232 / LLP.py:221-246 shape).
  * value     = layer-edges per second = E * L / step time  (unit "edges/s"); `pairs_per_s` = P / step time.
  * e2e       = the same through the public module API with the step's positive pair batch copied from pinned
                host memory every step and the loss read back to the host.
  * roofline  = dominant kernel: algorithmic bytes per launch (SURVEY.md section 8d model) / CUDA-event time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    # the peer-memory exchange runs kernels that wait on other kernels: nothing may be loaded lazily behind a spinning
    # one, and the side streams get hardware queues of their own (both are read when the CUDA context is created)
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import torch

WORKLOADS = {
    # BASELINE.json configs[2]: OGBL-DDI-shaped graph, 2-layer 8-head GAT (C = 256) + LinkPredictor(256,256)
    "ddi": dict(n_nodes=4267, n_edges=1_334_889, feat=256, hidden=256, heads=8, layers=2, pred_hidden=256, pos_pairs=None,
                graph="erdos-renyi(seed=1), no isolated rows"),
    # BASELINE.json configs[3]: power-law 2M nodes / 100M edges, 3 layers
    "rmat": dict(n_nodes=2_000_000, n_edges=100_000_000, feat=256, hidden=256, heads=8, layers=3, pred_hidden=256,
                 pos_pairs=2_000_000, graph="R-MAT(.57,.19,.19, seed=4) + self loops, ids permuted"),
    "rmat-s": dict(n_nodes=250_000, n_edges=12_500_000, feat=256, hidden=256, heads=8, layers=3, pred_hidden=256,
                   pos_pairs=500_000, graph="R-MAT(.57,.19,.19, seed=4) + self loops, ids permuted"),
    "small": dict(n_nodes=1000, n_edges=50_000, feat=64, hidden=64, heads=4, layers=2, pred_hidden=64, pos_pairs=None,
                  graph="erdos-renyi(seed=1), no isolated rows"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# synthetic graphs (host side, numpy; deterministic)
# ------------------------------------------------------------------------------------------------
def er_graph(n, e, seed, n_cols=None):
    """Directed Erdos-Renyi edge set with exactly e distinct (i, j) pairs, i < n, j < n_cols, no isolated row."""
    m = n if n_cols is None else n_cols
    rng = np.random.default_rng(seed)
    keys = np.unique(rng.integers(0, n * m, int(e * 1.02), dtype=np.int64))
    while keys.size < e:
        keys = np.unique(np.concatenate([keys, rng.integers(0, n * m, e - keys.size + 1024, dtype=np.int64)]))
    keys = rng.permutation(keys)[:e]
    rows, cols = keys // m, keys % m
    missing = np.setdiff1d(np.arange(n), rows)
    if missing.size:                                   # give isolated rows one neighbour (replaces a surplus edge)
        rows = np.concatenate([rows[: e - missing.size], missing])
        cols = np.concatenate([cols[: e - missing.size], rng.integers(0, m, missing.size)])
    return rows.astype(np.int64), cols.astype(np.int64)


def rmat_graph(n, e, seed, a=0.57, b=0.19, c=0.19):
    rng = np.random.default_rng(seed)
    scale = int(np.ceil(np.log2(n)))
    rows = np.zeros(e, dtype=np.int64)
    cols = np.zeros(e, dtype=np.int64)
    for lvl in range(scale):
        r = rng.random(e, dtype=np.float32)
        right = (r >= a) & (r < a + b) | (r >= a + b + c)
        down = r >= a + b
        rows |= down.astype(np.int64) << lvl
        cols |= right.astype(np.int64) << lvl
    perm = rng.permutation(1 << scale)
    rows, cols = perm[rows] % n, perm[cols] % n
    loops = np.arange(n, dtype=np.int64)               # self loops: no isolated rows (standard GAT practice)
    return np.concatenate([rows, loops]), np.concatenate([cols, loops])


def make_graph_host(wl, seed_shift=0):
    if wl["graph"].startswith("R-MAT"):
        return rmat_graph(wl["n_nodes"], wl["n_edges"], 4 + seed_shift)
    return er_graph(wl["n_nodes"], wl["n_edges"], 1 + seed_shift)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons DURING the timed region through NVML (in-process thread;
    an `nvidia-smi -lms` child stalls the driver for tens of ms per query on these hosts).  Every NVML query still
    delays kernel-completion delivery a little: at a 20 ms period the host-synchronised e2e loop of the graph-replayed
    `yearly` step measured 3.3-6.5 ms instead of 2.4 ms, hence the 50 ms default."""

    def __init__(self, index=0, period_s=float(os.environ.get("MSHA_BENCH_NVML_PERIOD", "0.05"))):
        self.index, self.period, self.rows, self._stop, self.t, self.err = index, period_s, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:          # noqa: BLE001
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:       # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, pw, rs))
            except Exception as e:      # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(self.period)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self._stop = True
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = [n for n, bit in names.items() if any(r[2] & bit for r in self.rows)]
        sm = [r[0] for r in self.rows]
        pw = [r[1] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm),
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes per C-ABI call (SURVEY.md section 8d gather model; fp32, int32 indices)
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(wl, E, P, N=None):
    """Per C-ABI call family: SURVEY.md section 8d per-unit bytes x the units one step hands it.  GAT families are per LAYER
    (one launch per layer on a single GPU; several block launches per layer on the pipelined multi-GPU path)."""
    N = wl["n_nodes"] if N is None else N
    C, H, Hd = wl["hidden"], wl["heads"], wl["pred_hidden"]
    fwd = E * (4 + 8 * H + 4 * C) + N * 4 * C
    return {
        "msha_gat_fwd": fwd,
        "msha_gat_fwd_block": fwd,
        "msha_gat_bwd_rows": E * (4 + 8 * H + 4 * C + 4 * H) + N * 12 * C,
        "msha_spmm_csc": E * (8 + 8 * H + 4 * C) + N * 4 * C,
        # scoring MLP, per pair (8d): fwd  8 (idx) + 8C (two row gathers) + 4Hd (scores out);
        # bwd  4Hd (d scores) + 4Hd (scores) + 8C (re-gather) + 16C (two read-modify-write row scatters)
        "msha_score_mlp_fwd": P * (8 + 8 * C + 4 * Hd),
        "msha_score_mlp_bwd": P * (8 + 8 * Hd + 8 * C + 16 * C),
        # nll read-out folded in: d scores is generated in the kernel, not read
        "msha_score_mlp_nll_bwd": P * (8 + 4 * Hd + 8 * C + 16 * C),
        # contraction-free backward: order + label + indices + one score sector, two row gathers and two RMW row scatters
        "msha_score_mlp_nll_bwd_sparse": P * (4 + 8 + 16 + 32 + 8 * C + 16 * C),
        "msha_pair_gather_mul": P * (16 + 8 * C + 4 * C),
        "msha_pair_scatter_mul_add": P * (16 + 4 * C + 8 * C + 16 * C),
    }


GAT_FAMILIES = ("msha_gat_fwd", "msha_gat_fwd_block", "msha_gat_softmax_stats", "msha_gat_bwd_rows", "msha_spmm_csc")


class KernelTimer:
    """Brackets every C-ABI call with CUDA events on the launching (current) stream."""

    def __init__(self, ops):
        self.ops, self.records, self.orig = ops, [], ops.call

    def __enter__(self):
        def timed(fname, *args):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.orig(fname, *args)
            b.record()
            if fname.startswith("msha_gemm"):
                shape = tuple(int(x) for x in args[3:6])          # M, N, K
            else:
                shape = tuple(int(x) for x in args if isinstance(x, int) and not isinstance(x, bool) and 0 < x < (1 << 40))[-6:]
            self.records.append((fname, a, b, shape))
        self.ops.call = timed
        for m in self._mods():
            if getattr(m, "call", None) is self.orig:
                m.call = timed
        return self

    def _mods(self):
        import msha_gnn_b200.functional as f, msha_gnn_b200.graph as g, msha_gnn_b200.intra as i
        import msha_gnn_b200.dist as d, msha_gnn_b200.peer as p, msha_gnn_b200.dist_p2p as dp, msha_gnn_b200.dist_msha as dm
        return [f, g, i, d, p, dp, dm]

    def __exit__(self, *exc):
        self.ops.call = self.orig
        for m in self._mods():
            m.call = self.orig
        torch.cuda.synchronize()

    def summary(self):
        agg = {}
        for fname, a, b, _ in self.records:
            t = a.elapsed_time(b)
            s = agg.setdefault(fname, [0, 0.0])
            s[0] += 1
            s[1] += t
        return agg


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def rmat_graph_device(n, e, seed, dev, a=0.57, b=0.19, c=0.19):
    """R-MAT edge list generated on the GPU (torch's Philox generator: the same seed gives the same list on every rank),
    ids randomly permuted, one self loop per node (no isolated rows).  -> (rows, cols) int64 device tensors."""
    g = torch.Generator(device=dev).manual_seed(seed)
    scale = int(np.ceil(np.log2(n)))
    rows = torch.zeros(e, dtype=torch.int64, device=dev)
    cols = torch.zeros(e, dtype=torch.int64, device=dev)
    for lvl in range(scale):
        r = torch.rand(e, generator=g, device=dev)
        rows |= (r >= a + b).long() << lvl
        cols |= (((r >= a) & (r < a + b)) | (r >= a + b + c)).long() << lvl
        del r
    perm = torch.randperm(1 << scale, generator=g, device=dev)
    rows, cols = perm[rows] % n, perm[cols] % n
    loops = torch.arange(n, dtype=torch.int64, device=dev)
    return torch.cat([rows, loops]), torch.cat([cols, loops])


def measure_tf32_peak(dev):
    """Dense TF32 tensor-pipe peak of THIS GPU, measured like MEASURED_PEAKS.json measures bf16 (library GEMM, 8192^3,
    best of 5): the denominator of the tensor-pipe fractions (kind::tf32 is what the 3xTF32 kernels issue)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / best / 1e9
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def parity_block(mg, mdist, model, x, graph, part, world, dev, n_rows=256, n_pairs=4096, seed=0):
    """Parity at benchmark scale ("check the GPU result on sampled rows"): every GAT layer's output
    on `n_rows` sampled local rows, and the scores of `n_pairs` sampled pairs, against the fp64 oracle evaluated on
    exactly those rows / pairs from the GPU's own layer inputs (oracle.gat_layer_rows / link_predictor).  rel. err =
    max|gpu - oracle| / max|oracle|; the north-star tolerance is 1e-4."""
    from oracle import msha_oracle as O               # checker only
    import torch.distributed as dist
    rng = np.random.default_rng(seed + part.rank)
    res = {"rows_sampled": 0, "pairs_sampled": 0, "layers": [], "tolerance": 1e-4}
    with torch.no_grad():
        n_loc = graph.n_rows
        S = np.sort(rng.choice(n_loc, size=min(n_rows, n_loc), replace=False))
        S_d = torch.from_numpy(S).to(dev)
        rp = graph.rowptr.long()
        beg, end = rp[S_d], rp[S_d + 1]
        cnt = (end - beg)
        idx = torch.repeat_interleave(beg - torch.cumsum(cnt, 0) + cnt, cnt) + torch.arange(int(cnt.sum()), device=dev)
        nbr = graph.col.long()[idx]                                     # padded column ids of the sampled rows' edges
        self_ids = S_d + part.rank * part.n_max
        nodes, inv = torch.unique(torch.cat([nbr, self_ids]), return_inverse=True)
        nbr_ptr = np.concatenate([[0], np.cumsum(cnt.cpu().numpy())])
        nbr_idx, self_idx = inv[:nbr.numel()].cpu().numpy(), inv[nbr.numel():].cpu().numpy()
        h = x.detach()
        worst = 0.0
        for conv in model.convs:
            h_g = mdist.all_gather_rows(h, part) if world > 1 else h     # layer input of every node (padded ids)
            out = mdist.gat_encode([conv], h, graph, part, training=False) if world > 1 else conv(h, graph)
            ref = O.gat_layer_rows(h_g[nodes].double().cpu(), conv.W.detach().double().cpu(), conv.a_nbr.detach().double().cpu(),
                                   conv.a_self.detach().double().cpu(), self_idx, nbr_ptr, nbr_idx, conv.heads)
            got = out[S_d].double().cpu()
            err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
            res["layers"].append(err)
            worst = max(worst, err)
            h = out
        h_g = mdist.all_gather_rows(h, part) if world > 1 else h
        n_all = h_g.shape[0]
        ps = torch.from_numpy(rng.integers(0, n_loc, n_pairs)).to(dev) + part.rank * part.n_max
        pd = part.to_padded(torch.from_numpy(rng.integers(0, part.n_nodes, n_pairs)).to(dev))
        sc = model.predictor.forward_pairs(h_g, h_g, ps, pd)
        lins = model.predictor.lins
        ref = O.link_predictor(h_g[ps].double().cpu(), h_g[pd].double().cpu(), [l.weight.detach().double().cpu() for l in lins],
                               [l.bias.detach().double().cpu() for l in lins])
        err_s = float((sc.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        res.update(rows_sampled=int(S.size), pairs_sampled=int(n_pairs), scores=err_s)
        worst = max(worst, err_s)
    t = torch.tensor([worst], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["worst_over_ranks"] = float(t.item())
    res["ok"] = bool(res["worst_over_ranks"] <= res["tolerance"])
    return res


_FABRIC = [None]


def run_job(args, wl_name, steps, warmup, full=True):
    """One workload on the ranks of this launch.  -> the JSON-line dict (rank 0) or None (other ranks)."""
    import torch.distributed as dist
    import msha_gnn_b200 as mg
    from msha_gnn_b200 import ops
    from msha_gnn_b200 import dist as mdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    wl = WORKLOADS[wl_name]
    hbm_peak, bf16_peak, peak_src = load_peaks()
    mg.functional.SPARSE_NLL_BWD = bool(args.contraction_free)

    # ---- data.  "weak" (ER workloads, N > 1): the graph grows with N -- N*n nodes, N*E edges, same degree distribution --
    # and rank r generates the edges of its own rows.  "strong" (R-MAT): one fixed graph, generated on every GPU from the
    # same seed and partitioned by destination-node range with edge-balanced cut points.
    strong = wl["graph"].startswith("R-MAT")
    rows_h = cols_h = None
    if strong:
        n_glob = wl["n_nodes"]
        rows_d, cols_d = rmat_graph_device(n_glob, wl["n_edges"], 4, dev)
        if world > 1:
            rowptr_h = torch.zeros(n_glob + 1, dtype=torch.int64)
            rowptr_h[1:] = torch.cumsum(torch.bincount(rows_d, minlength=n_glob), 0).cpu()
            part = mdist.Partition.edge_balanced(rowptr_h, world, rank)
            keep = (rows_d >= part.lo) & (rows_d < part.hi)
            rows_d, cols_d = rows_d[keep], cols_d[keep]
            del keep
        else:
            part = mdist.Partition(n_glob, 1, 0)
        n_loc = part.n_local
    else:
        n_loc, n_glob = wl["n_nodes"], wl["n_nodes"] * world
        part = mdist.Partition(n_glob, world, rank)
        if world == 1:
            rows_h, cols_h = make_graph_host(wl)
        else:
            rows_h, cols_h = er_graph(n_loc, wl["n_edges"], 1 + rank, n_cols=n_glob)
            rows_h = rows_h + part.lo
        rows_d, cols_d = torch.from_numpy(rows_h).to(dev), torch.from_numpy(cols_h).to(dev)
    L = wl["layers"]
    n_raw = int(rows_d.numel())
    n_pos = n_raw if wl["pos_pairs"] is None else min(n_raw, wl["pos_pairs"] // (world if strong else 1))
    pos_dev = torch.stack([rows_d[:n_pos], cols_d[:n_pos]]).contiguous()         # (2, n_pos) int64 positives (global ids)
    pos_host = torch.empty(pos_dev.shape, dtype=torch.int64, pin_memory=True)
    pos_host.copy_(pos_dev)
    if strong and full and world == 1 and not args.no_cpu_baseline:              # 1 % node-induced subsample for the CPU leg
        keep = (rows_d % 100 == 0) & (cols_d % 100 == 0)
        rows_h, cols_h = (rows_d[keep] // 100).cpu().numpy(), (cols_d[keep] // 100).cpu().numpy()
    if world == 1:
        graph = mg.Graph.from_coo(rows_d, cols_d, n_loc, n_loc)
    else:
        graph = mdist.partition_graph(rows_d, cols_d, part)
    del rows_d, cols_d
    E = graph.nnz                                      # distinct directed edges owned by this rank
    graph.attention_csc()
    use_p2p = world > 1 and args.comm == "p2p"
    p2p = None
    comm_note = None
    if use_p2p:
        from msha_gnn_b200 import peer
        if _FABRIC[0] is None:
            try:
                _FABRIC[0] = peer.SymmFabric()
            except Exception as e:      # noqa: BLE001  (no peer-mappable memory on this box: say so and use the collectives)
                _FABRIC[0] = repr(e)[:300]
        if isinstance(_FABRIC[0], str):
            use_p2p, comm_note = False, "peer-memory exchange unavailable (" + _FABRIC[0] + "): NCCL collectives used instead"
        else:
            p2p = mdist.P2P(_FABRIC[0].group, part)
    torch.manual_seed(42)                                                        # identical parameters on every rank
    model = mg.GATLinkModel(wl["feat"], wl["hidden"], wl["heads"], L, wl["pred_hidden"], dropout=0.0).to(dev)
    torch.manual_seed(100 + rank)
    x = torch.nn.Parameter(torch.rand(n_loc, wl["feat"], device=dev) * 0.1)      # learnable node features (GAT.py:42)
    params = list(model.parameters())
    # single GPU, small step: the whole step (sampler, forward, loss, backward, Adam) is replayed as one CUDA graph
    use_graph = world == 1 and not strong and not args.no_cuda_graph and not args.unfused_loss
    opt = torch.optim.Adam(params + [x], lr=1e-3, weight_decay=5e-4, fused=True, capturable=use_graph)   # train.py:207
    P = 2 * n_pos
    labels = torch.cat([torch.ones(n_pos, dtype=torch.int64, device=dev), torch.zeros(n_pos, dtype=torch.int64, device=dev)])
    lib = mg._lib.lib()
    e_tot = torch.tensor([E, P], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e_tot)
    E_global, P_global = (int(v) for v in e_tot.tolist())

    def step(it, pos):
        nsrc, ndst = mg.functional.negative_sample(1000 + it * world + rank, n_pos, n_loc, n_glob, dev)
        src = torch.cat([pos[0], nsrc + part.lo])
        dst = torch.cat([pos[1], ndst])
        opt.zero_grad(set_to_none=True)
        if world == 1:
            if args.unfused_loss:
                loss = mg.functional.nll_loss(model(x, graph, src, dst), labels)     # separate read-out op (LLP.py:235)
            else:
                loss = model.loss(x, graph, src, dst, labels)                        # same read-out, fused into the scorer
        else:
            if use_p2p:
                h = mdist.gat_encode_p2p(model.convs, x, graph, part, p2p, score_key="score_h")
            else:
                h = mdist.gat_encode(model.convs, x, graph, part)
            loss = mdist.score_pairs(model.predictor, h, src, dst, part, target=labels, global_pairs=P_global, p2p=p2p)
        loss.backward()
        if world > 1:
            mdist.allreduce_gradients(params, world=world)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph_info = {"cuda_graph": False}
    run_step = step                                      # (it, pos) -> loss

    def capture():
        """-> (callable like `step`, library launches per step) replaying the step as one CUDA graph; None when capture fails."""
        try:
            l0_ = lib.msha_launch_count()
            cap = mg.CapturedStep(lambda pos: step(4242, pos), [pos_dev], warmup=warmup)
            per = (lib.msha_launch_count() - l0_) // (warmup + 1)
            return (lambda it, pos: cap(pos)), int(per)
        except Exception as e:      # noqa: BLE001
            torch.cuda.synchronize()
            graph_info["cuda_graph_error"] = repr(e)[:300]
            return None

    def timed(n, first_it, feed=None):
        """n steps between two CUDA events on the launching stream, barrier + synchronize on both sides -> ms / step."""
        barrier()
        l0 = lib.msha_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for it in range(n):
            last = feed(it) if feed is not None else run_step(first_it + it, pos_dev)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n, lib.msha_launch_count() - l0, last

    launches_per_step = None
    if use_graph:
        got = capture()
        if got is not None:
            run_step, launches_per_step = got
            graph_info = {"cuda_graph": True, "launches_in_graph": launches_per_step}
    if not graph_info["cuda_graph"]:
        for it in range(warmup):
            step(it, pos_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing
    ms_dev, launches, _ = timed(steps, warmup)
    # ---- end-to-end timing: pinned host batch -> device every step, loss back to the host.  The pinned batch of step i+1 is
    # copied on a second stream while step i computes (two device buffers, an event per buffer): every step's host->device
    # copy and its loss read-back stay inside the timed region.
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(pos_dev), torch.empty_like(pos_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    total = [2]

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])              # the step that last read this buffer has finished
            bufs[b].copy_(pos_host, non_blocking=True)
            ready[b].record(copy_stream)

    def host_fed(seed0):
        def run(i):
            b = i & 1
            torch.cuda.current_stream().wait_event(ready[b])
            if i + 1 < total[0]:
                prefetch(i + 1)
            loss_ = run_step(seed0 + i, bufs[b])
            consumed[b].record()
            return float(loss_.item())
        return run

    for b in range(2):
        consumed[b].record()
    prefetch(0)
    run = host_fed(20_000)
    for it in range(2):                # untimed: the first host-fed steps allocate the per-step staging tensors
        run(it)
    barrier()
    total[0] = steps
    prefetch(0)
    ms_e2e, _, loss_host = timed(steps, 0, feed=host_fed(warmup + steps))
    clocks = sampler.stop() if rank == 0 else None
    # ---- the contraction-free scorer backward (exploits the one-hot d scores of the nll read-out) as a variant
    variant = None
    if full and not args.contraction_free and not args.unfused_loss:
        mg.functional.SPARSE_NLL_BWD = True
        keep_run = run_step
        got = capture() if graph_info["cuda_graph"] else None
        if got is not None:
            run_step = got[0]
        else:
            run_step = step
            for it in range(3):
                step(30_000 + it, pos_dev)
        ms_var, _, _ = timed(steps, 30_003)
        mg.functional.SPARSE_NLL_BWD = False
        run_step = keep_run
        variant = ms_var
    t = torch.tensor([ms_dev, ms_e2e, variant if variant is not None else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_var = t.tolist()

    # ---- per-kernel profile pass (outside the timed region; every rank runs it: the steps contain exchanges)
    roofline, kernels, gat_roofline, tf32_peak = None, [], None, None
    with KernelTimer(ops) as kt:
        step(10_000, pos_dev)
        step(10_001, pos_dev)
    if p2p is not None:
        p2p.pg.check()
        if p2p.trace is not None:                      # MSHA_P2P_TRACE=1: phase timeline of one step, per rank, to stderr
            p2p.dump_trace()
            step(10_002, pos_dev)
            tl = p2p.dump_trace()
            for rk in range(world):
                if rk == rank:
                    print(f"[p2p trace rank {rank}] " + " | ".join(f"{lab} {ms}" for lab, ms in tl), file=sys.stderr, flush=True)
                dist.barrier()
    if rank == 0:
        tf32_peak = measure_tf32_peak(dev)
        agg = kt.summary()
        ab = algorithmic_bytes(wl, E, P, n_loc)
        tot = sum(v[1] for v in agg.values())
        for fname, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            per = ms / cnt
            row = {"call": fname, "launches_per_step": cnt // 2, "avg_ms": round(per, 4), "share": round(ms / tot, 4)}
            if fname in ab:
                units = L if fname in GAT_FAMILIES else 1          # GAT families: bytes per layer, time per layer
                per_unit_ms = ms / 2 / units
                row["algorithmic_GB"] = round(ab[fname] / 1e9, 4)
                row["per"] = "layer" if units > 1 else "launch"
                row["ms_per_unit"] = round(per_unit_ms, 4)
                row["GBps"] = round(ab[fname] / 1e6 / per_unit_ms, 1)
                row["frac_hbm"] = round(ab[fname] / 1e6 / per_unit_ms / hbm_peak, 4)
            kernels.append(row)
        # dense-contraction flops per call (fp32-equivalent 2*M*N*K; the 3xTF32 split issues 3x that on the tensor pipe)
        gemm_flops = sum(2.0 * sh[0] * sh[1] * sh[2] for f, a, b, sh in kt.records if f.startswith("msha_gemm") and len(sh) >= 3) / 2
        C_, Hd_ = wl["hidden"], wl["pred_hidden"]
        tensor_flops = {"msha_gemm_tf32x3": gemm_flops, "msha_score_mlp_fwd": 2.0 * P * C_ * Hd_,
                        "msha_score_mlp_bwd": 4.0 * P * C_ * Hd_, "msha_score_mlp_nll_bwd": 4.0 * P * C_ * Hd_}
        for k in kernels:
            if k["call"] in tensor_flops:
                tf = tensor_flops[k["call"]] / (k["avg_ms"] * k["launches_per_step"] * 1e-3) / 1e12
                k["algorithmic_TFLOPs"] = round(tf, 1)
                k["tensor_frac_algorithmic"] = round(tf / tf32_peak, 4)
                k["tensor_frac_issued_3xtf32"] = round(3 * tf / tf32_peak, 4)
        dom = kernels[0] if kernels else None
        traffic = None                     # DRAM bytes per launch of the dominant call, from the committed ncu capture
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f).get(wl_name, {})
            if dom and world == 1 and tr.get("pairs") == P and dom["call"] in tr:
                traffic = tr[dom["call"]]["bytes"]
        except (OSError, ValueError):
            pass
        if dom and "GBps" in dom:
            roofline = {"bound": "hbm", "kernel": dom["call"], "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": dom["frac_hbm"], "traffic": traffic, "peak_source": peak_src, "share_of_step": dom["share"],
                        "algorithmic_bytes": "SURVEY.md section 8d per-unit bytes x units per launch (DESIGN.md section 4)"}
            if "algorithmic_TFLOPs" in dom:
                roofline.update(tensor_TFLOPs_algorithmic=dom["algorithmic_TFLOPs"], tensor_peak_tf32_measured=round(tf32_peak, 1),
                                tensor_frac_algorithmic=dom["tensor_frac_algorithmic"],
                                tensor_frac_issued_3xtf32=dom["tensor_frac_issued_3xtf32"],
                                note="SURVEY 8d classes the scorer as HBM-bound (35 flop/B): frac = algorithmic bytes / time / measured "
                                     "copy peak.  The tensor-pipe view of the same launch: 2*P*C*Hd (fwd) or 4*P*C*Hd (bwd) flops counted "
                                     "once against the TF32 peak measured in this run (torch.matmul, allow_tf32); fp32-grade results "
                                     "(rel. err <= 1e-4) need the 3xTF32 split, which issues three times that")
        # the graph-attention kernels of one layer (fwd + both backward passes) against the HBM roofline
        gat = [k for k in kernels if k["call"] in GAT_FAMILIES]
        if gat:
            gb = (ab["msha_gat_fwd"] + ab["msha_gat_bwd_rows"] + ab["msha_spmm_csc"]) / 1e9
            ms = sum(k["avg_ms"] * k["launches_per_step"] for k in gat) / L
            feat_mb = part.n_padded * wl["hidden"] * 4 / 1e6
            gat_roofline = {"bound": "hbm" if feat_mb > 126 else "l2", "kernels": [k["call"] for k in gat],
                            "algorithmic_GB_per_layer": round(gb, 3), "ms_per_layer": round(ms, 4),
                            "achieved": round(gb / ms * 1e3, 1), "peak": hbm_peak, "unit": "GB/s",
                            "frac": round(gb / ms * 1e3 / hbm_peak, 4),
                            "note": (f"feature table {feat_mb:.0f} MB " + ("exceeds the 126 MB L2: an HBM figure" if feat_mb > 126 else
                                     "is L2-resident: the gathers are served by L2, this is NOT an HBM fraction (it can exceed 1)"))}
    parity = parity_block(mg, mdist, model, x, graph, part, world, dev)
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None
    total_edges = E_global * L
    bwd_kind = ("separate nll_loss op + tensor-core backward (dense d scores written and re-read)" if args.unfused_loss else
                "contraction-free: d scores of the nll read-out is one-hot per pair, dZ = g_p * W0[label], dW0 a per-label sum" if args.contraction_free else
                "general tensor-core backward (dZ = G @ W0, dW0 = G^T @ Z on tcgen05), nll gradient generated in-kernel")
    out = {
        "metric": "gat_fwd_bwd_layer_edges_per_sec", "value": total_edges / (ms_dev / 1e3), "unit": "edges/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl_name, wl, n_glob, E_global, P_global),
        "per_gpu": ("whole graph on one GPU" if world == 1 else
                    f"{'strong' if strong else 'weak'} scaling: graph of {n_glob} nodes / {E_global} edges partitioned by destination-node "
                    f"range; per layer the gather of [Wh | s_nbr] (fwd) and the reduce-scatter of their gradients (bwd) "
                    + ("over NVLink peer memory: own flag kernels + copy-engine / SM pulls, pipelined under the attention kernels for "
                       "large blocks (msha_gnn_b200/peer.py, dist.py)" if use_p2p else "as NCCL all_gather / reduce_scatter")
                    + "; pairs data-parallel; parameter gradients all-reduced (NCCL)"),
        "scorer_backward": bwd_kind,
        **({"comm_note": comm_note} if comm_note else {}),
        "pairs_per_sec": P_global / (ms_dev / 1e3),
        "e2e": {"value": total_edges / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(pos_host.numel() * 8), "d2h_bytes_per_step": 4,
                "input_pipeline": "pinned batch of step i+1 copied on a second stream while step i computes; loss.item() every step"},
        "gpu_launches": int(launches_per_step * steps if graph_info["cuda_graph"] else launches),
        **graph_info,
        "clocks": clocks,
        "roofline": roofline,
        "gat_layer_roofline": gat_roofline,
        "kernels": kernels,
        "parity": parity,
        "loss": loss_host,
    }
    if variant is not None:
        out["contraction_free_scorer_backward"] = {
            "ms_per_step": ms_var, "value": total_edges / (ms_var / 1e3), "unit": "edges/s",
            "note": "same step with msha_score_mlp_nll_bwd_sparse (MSHA_NLL_BWD=sparse, the library default): exploits the one-hot "
                    "d scores of F.nll_loss; not the general backward, hence reported beside the headline"}
    if full and not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args, wl, rows_h, cols_h, subsampled=strong)
    return out


def workload_config(name, wl, n_glob, E_global, P_global):
    """config of the JSON line -- identical for the GPU and the reference arm (same workload, same wording)."""
    return {"workload": f"{name}: {n_glob} nodes, {E_global} directed edges ({wl['graph']}), F={wl['feat']}, "
                        f"{wl['layers']}-layer {wl['heads']}-head GAT C={wl['hidden']} + LinkPredictor(mlp,{wl['hidden']},"
                        f"{wl['pred_hidden']}) over P={P_global} pairs (positives + as many Philox negatives), nll loss, Adam",
            "dropout": 0.0,
            "l2_policy": "per-step working set (>= 4*P*pred_hidden bytes of scores) exceeds the 126 MB L2; no explicit flush"}


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run_job(args, args.workload, args.steps, args.warmup, full=True)
    # BASELINE.json configs[3] next to the headline: the 100 M-edge R-MAT graph on the same N GPUs (north-star: >= 6x from
    # 1 to 8 GPUs, strong scaling).  Fewer steps: a step is 50-300 ms.
    if args.workload == "ddi" and not args.no_strong_scaling:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            ss = run_job(args, "rmat", max(3, min(args.steps, 5)), 3, full=False)
            if rank == 0:
                keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "per_gpu",
                        "pairs_per_sec", "e2e", "gpu_launches", "gat_layer_roofline", "kernels", "parity", "loss")
                out["strong_scaling"] = {k: ss[k] for k in keep if k in ss}
        except Exception as e:      # noqa: BLE001
            if rank == 0:
                out["strong_scaling"] = {"error": repr(e)[:400]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: multi-scale hierarchical attention (Ours) fwd+bwd on a 10 M-node / 500 M-edge graph plus a 50 M-pair
# link-scoring sweep at 8 B200.  Weak definition: every GPU holds 1.25 M sources, 1.25 M recipients and 62.5 M flow edges
# (8 GPUs = the named shape); the batch (65 536 sources) and the sweep (50 M pairs) are global and split over the ranks.
# ------------------------------------------------------------------------------------------------
MSHA_WL = dict(n_per_gpu=1_250_000, e_per_gpu=62_500_000, feat=128, d=64, heads=2, batch=65_536, sweep=50_000_000,
               cities=3000, provinces=30)


def run_msha(args):
    import torch.distributed as dist
    import msha_gnn_b200 as mg
    from msha_gnn_b200 import ops, dist as mdist, dist_msha as dm, dist_p2p as mp2p
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = dict(MSHA_WL)
    if args.scale != 1.0:
        wl["n_per_gpu"] = int(wl["n_per_gpu"] * args.scale)
        wl["e_per_gpu"] = int(wl["e_per_gpu"] * args.scale)
        wl["sweep"] = int(wl["sweep"] * args.scale)
    n_loc, F, d, H = wl["n_per_gpu"], wl["feat"], wl["d"], wl["heads"]
    C = H * d
    n_glob = n_loc * world
    ps, pr = mdist.Partition(n_glob, world, rank), mdist.Partition(n_glob, world, rank)
    g = torch.Generator(device=dev).manual_seed(50 + rank)
    e_loc = wl["e_per_gpu"]
    rows = torch.randint(0, n_loc, (e_loc,), generator=g, device=dev) + ps.lo
    u = torch.rand(e_loc, generator=g, device=dev, dtype=torch.float64)
    cols = ((u * u * u * n_glob).long().clamp_(max=n_glob - 1) * 2654435761) % n_glob      # skewed in-degree, ids scattered
    del u
    loops = torch.arange(ps.lo, ps.hi, device=dev)                                           # every source has a recipient
    rows, cols = torch.cat([rows, loops]), torch.cat([cols, (loops * 40503) % n_glob])
    pgraph = mdist.partition_graph(rows, cols, ps, col_part=pr)
    del rows, cols
    pgraph.attention_csc()
    E = pgraph.nnz
    city = torch.randint(0, wl["cities"], (n_loc,), generator=g, device=dev)
    prov = city % wl["provinces"]
    n3 = torch.bincount(city, minlength=wl["cities"]).float()
    n4 = torch.bincount(prov, minlength=wl["provinces"]).float()
    if world > 1:
        dist.all_reduce(n3)
        dist.all_reduce(n4)
    use_p2p = world > 1 and args.comm == "p2p"
    if use_p2p:
        from msha_gnn_b200 import peer
        if _FABRIC[0] is None:
            _FABRIC[0] = peer.SymmFabric()
        comm = dm.PeerComm({id(ps): mp2p.P2P(_FABRIC[0].group, ps), id(pr): mp2p.P2P(_FABRIC[0].group, pr)})
    else:
        comm = dm.TorchComm()
    torch.manual_seed(42)
    layers = torch.nn.ModuleList([mg.OursLayer(F, d, dropout=0.0) for _ in range(H)]).to(dev)
    torch.manual_seed(100 + rank)
    S = torch.nn.Parameter(torch.rand(n_loc, F, device=dev))
    R = torch.nn.Parameter(torch.rand(n_loc, F, device=dev))
    params = list(layers.parameters())
    opt = torch.optim.Adam(params + [S, R], lr=1e-3, weight_decay=5e-4, fused=True)
    B_loc = wl["batch"] // world
    P_sweep = wl["sweep"] // world
    n_batches = args.warmup + 2 * args.steps + 8
    batch_host = torch.randint(0, n_loc, (n_batches, 2, B_loc), generator=torch.Generator().manual_seed(rank)).pin_memory()
    batch_host[:, 1] = torch.randint(0, n_glob, (n_batches, B_loc), generator=torch.Generator().manual_seed(1000 + rank))
    batch_dev = batch_host.to(dev)
    labels = torch.cat([torch.ones(B_loc, device=dev), torch.zeros(B_loc, device=dev)])
    lib = mg._lib.lib()
    e_tot = torch.tensor([E], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e_tot)
    E_global = int(e_tot.item())
    keep = {}

    def step(it, b):
        src_b, rec_b = b[0], b[1]                                              # batch sources (local rows), their recipients
        nsrc, ndst = mg.functional.negative_sample(7000 + it * world + rank, B_loc, n_loc, n_glob, dev)
        opt.zero_grad(set_to_none=True)
        u_, v_ = dm.ours_encode(list(layers), S, R, pgraph, ps, pr, comm, source_index=src_b, city_ids=city, province_ids=prov,
                                group_sizes=(n3, n4))
        sc = dm.score_uv_pairs(u_, v_, torch.cat([src_b, nsrc]), torch.cat([rec_b, ndst]), pr, comm, heads=H)
        loss = mg.functional.mse_loss(sc.mean(dim=1), labels) * (1.0 / world)   # global mean over the ranks' equal shares
        loss.backward()
        if world > 1:
            mdist.allreduce_gradients(params, world=world)
        opt.step()
        keep["u"], keep["v"] = u_.detach(), v_.detach()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        l0 = lib.msha_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(n):
            last = fn(i)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n, lib.msha_launch_count() - l0, last

    for it in range(args.warmup):
        step(it, batch_dev[it])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches, _ = timed(lambda i: step(args.warmup + i, batch_dev[args.warmup + i]), args.steps)
    base = args.warmup + args.steps

    def host_fed(i):
        b = batch_host[base + i].to(dev, non_blocking=True)
        return float(step(base + i, b).item())
    host_fed(args.steps)
    ms_e2e, _, loss_host = timed(host_fed, args.steps)
    # ---- the link-scoring sweep: forward scores of P_sweep pairs per rank against the trained (u, v)
    gs = torch.Generator(device=dev).manual_seed(900 + rank)
    sw_src = torch.randint(0, n_loc, (P_sweep,), generator=gs, device=dev)
    sw_dst = torch.randint(0, n_glob, (P_sweep,), generator=gs, device=dev)

    def sweep(i):
        with torch.no_grad():
            return dm.score_uv_pairs(keep["u"], keep["v"], sw_src, sw_dst, pr, comm, heads=H)
    sweep(0)
    ms_sweep, _, sc_sw = timed(sweep, max(2, min(args.steps, 5)))
    clocks = sampler.stop() if rank == 0 else None
    # parity of the sweep's scores on a sample: elu(u_i . v_j) recomputed in fp64 from the same (u, v)
    with torch.no_grad():
        idx = torch.randint(0, P_sweep, (4096,), device=dev)
        v_g = comm.gather_rows(keep["v"], pr, "msha.vg")
        ui = keep["u"][sw_src[idx]].double().view(-1, H, d)
        vj = v_g[pr.to_padded(sw_dst[idx])].double().view(-1, H, d)
        ref = torch.nn.functional.elu((ui * vj).sum(-1))
        perr = float((sc_sw[idx].double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    with KernelTimer(ops) as kt:
        step(10_000, batch_dev[0])
        step(10_001, batch_dev[1])
    t = torch.tensor([ms_dev, ms_e2e, ms_sweep, perr], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_sweep, perr = t.tolist()
    if rank == 0:
        hbm_peak, _, peak_src = load_peaks()
        agg = kt.summary()
        tot = sum(v[1] for v in agg.values())
        # GAT layer fwd+bwd bytes per edge + the MSHA extras (one transposed pass fwd, two bwd): 8 + 4H + 4C each
        alg = E * ((16 + 24 * H + 12 * C) + 3 * (8 + 4 * H + 4 * C)) + 2 * n_loc * (12 * F + 20 * C)
        kernels = [{"call": f, "launches_per_step": c // 2, "avg_ms": round(ms / c, 4), "share": round(ms / tot, 4)}
                   for f, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])][:14]
        out = {"metric": "gat_fwd_bwd_layer_edges_per_sec", "value": E_global / (ms_dev / 1e3), "unit": "edges/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": f"ours: {n_glob} sources x {n_glob} recipients, {E_global} flow edges (skewed in-degree), "
                                      f"MSHA layer (OursLayer x {H} heads, in={F}, out={d}, p=0) fwd+bwd with the intra scales over "
                                      f"{wl['cities']} cities / {wl['provinces']} provinces, batch {B_loc * world}, pair read-out "
                                      f"elu(u_i.v_j) + MSE, Adam; then a {P_sweep * world}-pair scoring sweep (BASELINE.json configs[4]"
                                      + ("" if world == 8 and args.scale == 1.0 else f"; the named shape is 8 GPUs at scale 1, this is {world} x scale {args.scale}") + ")",
                          "dropout": 0.0, "l2_policy": "tables of 0.6 GB and more per rank: far beyond the 126 MB L2"},
               "per_gpu": ("one GPU's share" if world == 1 else "sources and recipients split by contiguous range; gather of h1 / s_nbr / v, "
                           "reduce-scatter of alpha.T@h2, all-reduce of BN statistics and intra-scale group tables over "
                           + ("NVLink peer memory (dist_p2p / dist_msha)" if use_p2p else "NCCL")),
               "pairs_per_sec": P_sweep * world / (ms_sweep / 1e3), "sweep_ms": ms_sweep,
               "e2e": {"value": E_global / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": int(2 * B_loc * 8), "d2h_bytes_per_step": 4},
               "gpu_launches": int(launches), "clocks": clocks,
               "roofline": {"bound": "hbm", "kernel": "whole MSHA step (attention fwd/bwd kernels dominate)", "achieved": alg / (ms_dev / 1e3) / 1e9,
                            "peak": hbm_peak, "unit": "GB/s", "frac": alg / (ms_dev / 1e3) / 1e9 / hbm_peak, "traffic": None,
                            "note": "SURVEY 8d algorithmic bytes of one rank's layer (GAT fwd+bwd + three transposed passes + node terms) / step time; " + peak_src},
               "kernels": kernels, "parity": {"sweep_scores_vs_fp64": perr, "pairs_sampled": 4096, "tolerance": 1e-4, "ok": perr <= 1e-4,
                                              "note": "layer parity: tests/test_gpu_p2p.py::test_partitioned_msha_layer_matches_single_gpu"},
               "loss": loss_host}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# "flow" workloads: the reference's own training configuration (BASELINE.json configs[0]/[1]) on a graph with the
# statistics of the shipped 2015 data: N = 39 179 sources, M = 32 recipients, 233 887 flow
# records, 1-30 distinct recipients per source, 291 cities / 25 provinces.  Step = train.py:221-232.
# ------------------------------------------------------------------------------------------------
def flow_graph(seed=2015, N=39179, M=32, n_records=233887):
    """2015-shaped flow records (SURVEY.md section 8d cfg 1/2): every source has k distinct recipients, k ~ geometric with
    mean 2.33 capped at 30 (nnz ~ 2.33 N: 91 283 in the real 2015 files), recipients drawn without replacement in
    proportion to a skewed in-degree (Gumbel top-k), the remaining records repeat existing (source, recipient) pairs."""
    rng = np.random.default_rng(seed)
    k = np.minimum(rng.geometric(1 / 2.33, N), min(30, M))
    pop = rng.pareto(1.0, M) + 0.05
    pop /= pop.sum()                                                          # skewed recipient in-degree
    order = np.argsort(-(np.log(pop)[None, :] + rng.gumbel(size=(N, M))), axis=1)
    src = np.repeat(np.arange(N), k)
    dst = order[np.arange(M)[None, :] < k[:, None]]                           # row-major: matches the repeat above
    extra = rng.integers(0, src.size, max(n_records - src.size, 0))           # repeated records (multiplicities)
    src = np.concatenate([src, src[extra]])
    dst = np.concatenate([dst, dst[extra]])
    city = rng.integers(0, 291, N)
    prov = city % 25
    return src.astype(np.int64), dst.astype(np.int64), city.astype(np.int64), prov.astype(np.int64)


def capture_step(mg, lib, step, example, use_graph, warmup=2):
    """Record `step` (forward, loss, backward, Adam) into one CUDA graph (msha_gnn_b200.graphs.CapturedStep).
    -> (callable, {"cuda_graph": bool, ...}, library kernels per step).  A failed capture is reported, not hidden."""
    if not use_graph:
        return step, {"cuda_graph": False}, 0
    try:
        l0 = lib.msha_launch_count()
        captured = mg.CapturedStep(step, [example], warmup=warmup)
        per_step = (lib.msha_launch_count() - l0) // (warmup + 1)   # eager warm-ups + the recorded step
        return captured, {"cuda_graph": True, "launches_in_graph": int(per_step)}, per_step
    except Exception as e:      # noqa: BLE001
        torch.cuda.synchronize()
        return step, {"cuda_graph": False, "cuda_graph_error": repr(e)[:300]}, 0


def run_flow(args):
    import msha_gnn_b200 as mg
    from msha_gnn_b200 import ops
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    N, M, B = 39179, 32, 64
    src, dst, city, prov = flow_graph()
    graph = mg.Graph.from_coo(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), N, M)
    graph.attention_csc()
    E = graph.nnz
    full = args.workload == "flow-ours"
    gdp = {str(i): 0.05 for i in range(N)}
    torch.manual_seed(42)
    cls = mg.Ours if full else mg.ablation3
    model = cls(in_features=128, out_features=64, n_classes=M, n_heads=2, dropout=0.5, gdp=gdp, Scount=N, Rcount=M).to(dev)
    use_graph = not args.no_cuda_graph
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4, fused=True, capturable=use_graph)   # train.py:207
    city_d, prov_d = torch.from_numpy(city).to(dev), torch.from_numpy(prov).to(dev)
    rng = np.random.default_rng(0)
    rec_idx = rng.integers(0, src.size, (args.warmup + 2 * args.steps + 4, B))
    batches_host = torch.from_numpy(np.stack([src[rec_idx], dst[rec_idx]], axis=1)).pin_memory()   # (steps, 2, B)
    batches_dev = batches_host.to(dev)
    model.train()
    lib = mg._lib.lib()

    def step(b):
        s_i, r_i = b[0], b[1]
        opt.zero_grad(set_to_none=True)
        out = model(graph, city_d, prov_d, s_i) if full else model(graph, None, None, s_i)
        loss = torch.nn.functional.nll_loss(out[s_i], r_i)                     # train.py:229
        loss.backward()
        opt.step()
        return loss

    if not use_graph:
        for i in range(args.warmup):
            step(batches_dev[i])
    torch.cuda.synchronize()
    run, graph_info, per_step_launches = capture_step(mg, lib, step, batches_dev[0], use_graph, warmup=args.warmup)
    if use_graph and not graph_info["cuda_graph"]:
        for i in range(args.warmup):
            step(batches_dev[i])
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = lib.msha_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = run(batches_dev[args.warmup + i])
    e1.record()
    torch.cuda.synchronize()
    launches = per_step_launches * args.steps if graph_info["cuda_graph"] else lib.msha_launch_count() - l0
    ms_dev = e0.elapsed_time(e1) / args.steps
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    trace = [0.0, 0.0] if os.environ.get("MSHA_BENCH_TRACE") else None
    for i in range(args.steps):
        hb = batches_host[args.warmup + args.steps + i]
        t0 = time.perf_counter()
        loss_t = run(hb if graph_info["cuda_graph"] else hb.to(dev, non_blocking=True))
        t1 = time.perf_counter()
        loss_host = float(loss_t.item())
        if trace is not None:
            trace[0] += t1 - t0
            trace[1] += time.perf_counter() - t1
    e3.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    if trace is not None:
        print(f"e2e host trace: enqueue {trace[0] / args.steps * 1e3:.3f} ms, wait for loss {trace[1] / args.steps * 1e3:.3f} ms per step",
              file=sys.stderr)
    with KernelTimer(ops) as kt:
        step(batches_dev[0])
        step(batches_dev[1])
    agg = kt.summary()
    tot = sum(v[1] for v in agg.values())
    kernels = [{"call": f, "launches_per_step": c // 2, "avg_ms": round(ms / c, 4), "share": round(ms / tot, 4)}
               for f, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])][:10]
    layers = 3                                                                    # 2 heads + out_att applications
    # SURVEY 8d algorithmic bytes of one step (fp32, int32 indices): the inter-scale attention of H = 2 heads fwd+bwd with its
    # three transposed passes, the node terms of both sides, BatchNorm (7 passes over (N + M, C)), the dense elu(u @ v.T) score
    # matrix and log-softmax (8 passes over (N, H*M) / (N, M)), the degenerate out_att layer (a-1); the intra scales of the
    # full model add the batch rows' group lists
    Hh, dd, Fin = 2, 64, 128
    Cc = Hh * dd
    alg_bytes = (E * ((16 + 24 * Hh + 12 * Cc) + 3 * (8 + 4 * Hh + 4 * Cc)) + (N + M) * (12 * Fin + 20 * Cc) + 7 * (N + M) * Cc * 4
                 + 8 * N * Hh * M * 4 + N * (12 * Hh * M + 8 * M) + 8 * E)
    if full:
        alg_bytes += 2 * B * (N // 291 + N // 25) * 4 * dd * Hh
    hbm_peak, _, peak_src = load_peaks()
    out = {"metric": "gat_fwd_bwd_layer_edges_per_sec", "value": E * layers / (ms_dev / 1e3), "unit": "edges/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.workload}: 2015-shaped flow graph N={N}, M={M}, nnz={E} ({src.size} records), "
                                  f"{'Ours' if full else 'ablation3'}(in=128,out=64,H=2,p=0.5), batch {B}, nll_loss, Adam "
                                  "(train.py:206-232)",
                      "dropout": 0.5,
                      "l2_policy": "graph and parameters (~50 MB) are L2-resident by nature of the workload; launch/latency bound"},
           "e2e": {"value": E * layers / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": 2 * B * 8, "d2h_bytes_per_step": 4},
           "gpu_launches": int(launches), "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": "whole step (~60 kernels on a graph that fits the L2; launch / latency bound)",
                        "achieved": alg_bytes / (ms_dev / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (ms_dev / 1e3) / 1e9 / hbm_peak, "traffic": None,
                        "note": "algorithmic bytes of the step (SURVEY 8d a-3 / a-1 models) / device step time; " + peak_src},
           "kernels": kernels, "loss": loss_host, **graph_info}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = flow_cpu_baseline(src, dst, N, M, B, reps=2)
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[1]: GAT.py 2-layer 8-head GAT + LinkPredictor MLP scorer on the 2015-2018 yearly graphs
# (SURVEY.md section 8d cfg 2: real node-table sizes, synthetic flows shaped like 2015)
# ------------------------------------------------------------------------------------------------
YEARLY = {"2015": (39179, 233887), "2016": (48032, 286700), "2017": (49700, 296700), "2018": (50015, 298600)}


def yearly_cpu_baseline(src, dst, N, M, H, B, reps=2):
    """GAT.forward (GAT.py:53-58) + LinkPredictor (LLP.py:104-115) + nll, fwd + bwd, oracle port in fp32 on the 2015 graph."""
    from oracle import msha_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    rowptr, col, _ = O.csr_from_coo(src, dst, N, M)
    g = torch.Generator().manual_seed(0)
    dt = torch.float32

    def P(*shape):
        return (torch.rand(*shape, generator=g, dtype=dt) - 0.5).requires_grad_(True)
    heads = [(P(M, M), P(2 * M, 1)) for _ in range(H)]
    out_p = (P(M * H, M), P(2 * M, 1))
    feats = torch.rand(N, M, generator=g, dtype=dt)
    lw, lb = [P(M, M), P(1, M)], [P(M), P(1)]
    orig_t = O._t
    O._t = lambda a, d_=dt: orig_t(a, d_)
    try:
        ts = []
        for _ in range(reps + 1):
            s_i = torch.randint(0, N, (B,), generator=g)
            r_i = torch.randint(0, M, (B,), generator=g)
            t0 = time.perf_counter()
            h = O.gat_model(feats, heads, out_p, rowptr, col)
            out = O.link_predictor(h[s_i], h[r_i], lw, lb)
            loss = -(out[torch.arange(B), r_i]).mean()
            torch.autograd.grad(loss, [w for w, _ in heads] + [out_p[0], lw[0], lb[0]])
            ts.append(time.perf_counter() - t0)
    finally:
        O._t = orig_t
    t = min(ts[1:])
    return {"value": col.size * 2 / t, "unit": "edges/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"the 2015 graph only (1 of the 4 yearly graphs), full step fwd + nll + bwd without Adam, best of {reps} "
                      "after warm-up; oracle port (torch CPU fp32)", "ms_per_step_est": t * 1e3}


def run_yearly(args):
    import msha_gnn_b200 as mg
    from msha_gnn_b200 import ops
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    M, H, B = 32, 8, 4096                                                      # LLP.py:32 batch size
    years = []
    for y, (N, n_rec) in YEARLY.items():
        src, dst, _, _ = flow_graph(seed=int(y), N=N, M=M, n_records=n_rec)
        graph = mg.Graph.from_coo(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), N, M)
        graph.attention_csr()
        years.append(dict(year=y, N=N, src=src, dst=dst, graph=graph))
    torch.manual_seed(42)
    for yr in years:
        gdp = {str(i): 0.05 for i in range(yr["N"])}
        yr["model"] = mg.GAT(n_features=M, n_classes=M, n_heads=H, dropout=0.5, gdp=gdp, N=yr["N"]).to(dev)   # train.py:199
    predictor = mg.LinkPredictor('mlp', M, M, 1, 2, 0.5).to(dev)                # LLP.py:292
    params = [p for yr in years for p in yr["model"].parameters()] + list(predictor.parameters())
    use_graph = not args.no_cuda_graph
    opt = torch.optim.Adam(params, lr=5e-3, fused=True, capturable=use_graph)   # LLP.py:15,297
    rng = np.random.default_rng(0)
    n_batches = args.warmup + 2 * args.steps + 4
    per_year = []
    for yr in years:
        idx = rng.integers(0, yr["src"].size, (n_batches, B))
        per_year.append(np.stack([yr["src"][idx], yr["dst"][idx]], axis=1))      # (batches, 2, B)
        yr["model"].train()
    batches_host = torch.from_numpy(np.stack(per_year, axis=1)).pin_memory()     # (batches, years, 2, B)
    batches_dev = batches_host.to(dev)
    predictor.train()
    lib = mg._lib.lib()
    E_tot = sum(yr["graph"].nnz for yr in years)
    N_tot = sum(yr["N"] for yr in years)

    # the four yearly graphs are independent until the shared scorer's gradient: each year's chain of small kernels (~30
    # launches of 5-35 us, none of which fills the GPU) runs on its own stream, forked from and joined to the step's stream
    # -- inside the captured graph these become four parallel branches
    year_streams = [torch.cuda.Stream() for _ in years] if not args.no_year_streams else None

    def step(batch):
        opt.zero_grad(set_to_none=True)
        cur = torch.cuda.current_stream()
        losses = []
        for k, yr in enumerate(years):
            s_i, r_i = batch[k, 0], batch[k, 1]
            if year_streams is not None:
                year_streams[k].wait_stream(cur)
            with torch.cuda.stream(year_streams[k] if year_streams is not None else cur):
                h = yr["model"](yr["graph"])                                    # (N, M) log-probs, GAT.py:53-58
                losses.append(predictor.nll_loss_pairs(h, h, s_i, r_i, r_i))    # LLP.py:233-235
        if year_streams is not None:
            for st in year_streams:
                cur.wait_stream(st)
        total = losses[0]
        for l_ in losses[1:]:
            total = total + l_
        total.backward()
        opt.step()
        return total

    if not use_graph:
        for i in range(args.warmup):
            step(batches_dev[i])
    torch.cuda.synchronize()
    run, graph_info, per_step_launches = capture_step(mg, lib, step, batches_dev[0], use_graph, warmup=args.warmup)
    if use_graph and not graph_info["cuda_graph"]:
        for i in range(args.warmup):
            step(batches_dev[i])
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = lib.msha_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run(batches_dev[args.warmup + i])
    e1.record()
    torch.cuda.synchronize()
    launches = per_step_launches * args.steps if graph_info["cuda_graph"] else lib.msha_launch_count() - l0
    ms_dev = e0.elapsed_time(e1) / args.steps
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    trace = [0.0, 0.0] if os.environ.get("MSHA_BENCH_TRACE") else None
    for i in range(args.steps):
        hb = batches_host[args.warmup + args.steps + i]
        t0 = time.perf_counter()
        loss_t = run(hb if graph_info["cuda_graph"] else hb.to(dev, non_blocking=True))
        t1 = time.perf_counter()
        loss_host = float(loss_t.item())
        if trace is not None:
            trace[0] += t1 - t0
            trace[1] += time.perf_counter() - t1
    e3.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    if trace is not None:
        print(f"e2e host trace: enqueue {trace[0] / args.steps * 1e3:.3f} ms, wait for loss {trace[1] / args.steps * 1e3:.3f} ms per step",
              file=sys.stderr)
    with KernelTimer(ops) as kt:
        step(batches_dev[0])
        step(batches_dev[1])
    agg = kt.summary()
    tot = sum(v[1] for v in agg.values())
    kernels = [{"call": f, "launches_per_step": c // 2, "avg_ms": round(ms / c, 4), "share": round(ms / tot, 4)}
               for f, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])][:10]
    # a-1 is node-dominated: per head-layer fwd+bwd N*(12F + 8M) + 8E bytes; H heads of F = M plus out_att (F = H*M)
    alg_bytes = sum(yr["N"] * (12 * M + 8 * M) * H + yr["N"] * (12 * H * M + 8 * M) + 8 * yr["graph"].nnz * (H + 1) for yr in years)
    hbm_peak, _, peak_src = load_peaks()
    out = {"metric": "gat_fwd_bwd_layer_edges_per_sec", "value": E_tot * 2 / (ms_dev / 1e3), "unit": "edges/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"yearly: 4 yearly graphs N={[yr['N'] for yr in years]}, M={M}, nnz={[yr['graph'].nnz for yr in years]} "
                                  f"(2015-shaped synthetic flows), GAT(n_features=32,n_classes=32,n_heads=8,p=0.5) per year + shared "
                                  f"LinkPredictor(mlp,32,32,1,2,0.5), batch {B} pairs per year, nll_loss, Adam (BASELINE.json configs[1])",
                      "dropout": 0.5, "year_streams": year_streams is not None,
                      "l2_policy": "per-step working set (4 x N x 256 x 4 B x several tensors, ~0.5 GB) exceeds the 126 MB L2; no explicit flush"},
           "nodes_per_sec": N_tot / (ms_dev / 1e3), "pairs_per_sec": 4 * B / (ms_dev / 1e3),
           "e2e": {"value": E_tot * 2 / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": 4 * 2 * B * 8, "d2h_bytes_per_step": 4},
           "gpu_launches": int(launches), "clocks": clocks,
           "roofline": {"bound": "hbm", "kernel": "whole step (node-dominated a-1 path; launch-bound)", "achieved": alg_bytes / (ms_dev / 1e3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (ms_dev / 1e3) / 1e9 / hbm_peak, "traffic": None,
                        "note": "algorithmic bytes of the step (SURVEY 8d a-1 model) / device step time; peak " + peak_src},
           "kernels": kernels, "loss": loss_host, **graph_info}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = yearly_cpu_baseline(years[0]["src"], years[0]["dst"], years[0]["N"], M, H, B)
    print(json.dumps(out))


def flow_cpu_baseline(src, dst, N, M, B, reps=2):
    """The reference's ablation3 training step (Ablation.py:279-301 via the oracle port, fp32, all host threads)."""
    from oracle import msha_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    rowptr, col, _ = O.csr_from_coo(src, dst, N, M)
    g = torch.Generator().manual_seed(0)
    dt = torch.float32
    def P(*shape):
        return (torch.rand(*shape, generator=g, dtype=dt) - 0.5).requires_grad_(True)
    heads = []
    for _ in range(2):
        heads.append({"W1": P(128, 64), "W2": P(128, 64), "a": P(128, 1), "bn1.weight": torch.ones(64, requires_grad=True),
                      "bn1.bias": torch.zeros(64, requires_grad=True), "bn2.weight": torch.ones(64, requires_grad=True),
                      "bn2.bias": torch.zeros(64, requires_grad=True), "bn1.running_mean": torch.zeros(64),
                      "bn1.running_var": torch.ones(64), "bn2.running_mean": torch.zeros(64), "bn2.running_var": torch.ones(64)})
    S, R, Wo, ao = P(N, 128), P(M, 128), P(2 * M, M), P(2 * M, 1)
    orig_t = O._t
    O._t = lambda a, d_=dt: orig_t(a, d_)
    try:
        ts = []
        for i in range(reps + 1):
            s_i = torch.randint(0, N, (B,), generator=g)
            r_i = torch.randint(0, M, (B,), generator=g)
            t0 = time.perf_counter()
            out = O.msha_model(S, R, heads, (Wo, ao), rowptr, col, training=True, variant=3)
            loss = O.nll_readout(out, s_i, r_i)
            leaves = [S, R, Wo] + [h[k] for h in heads for k in ("W1", "W2", "a")]
            torch.autograd.grad(loss, leaves)
            ts.append(time.perf_counter() - t0)
    finally:
        O._t = orig_t
    t = min(ts[1:])
    return {"value": col.size * 3 / t, "unit": "edges/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"full ablation3 step (fwd + nll + bwd, no Adam) on the same graph, best of {reps} after warm-up; oracle "
                      "port (sparse restatement, torch CPU fp32) -- the dense reference needs ~0.85 s/step (BASELINE.md)",
            "ms_per_step_est": t * 1e3}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (torch CPU, fp32, all host threads) on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
class CpuStep:
    """The oracle port (torch CPU fp32 sparse restatement, all host threads) of the same training step: L GAT layers,
    LinkPredictor scoring of a pair sample, nll read-out, backward, Adam (train.py:221-232 / LLP.py:221-246 shape)."""

    def __init__(self, wl, rows, cols, pair_sample, dtype=torch.float32):
        from oracle import msha_oracle as O
        self.O = O
        torch.set_num_threads(os.cpu_count() or 1)
        N, Fin, C, H, L, Hd = wl["n_nodes"], wl["feat"], wl["hidden"], wl["heads"], wl["layers"], wl["pred_hidden"]
        self.H, self.L, self.dtype = H, L, dtype
        self.rowptr, self.col, _ = O.csr_from_coo(rows, cols, N, N)
        self.E = self.col.size
        g = torch.Generator().manual_seed(0)
        self.x = torch.rand(N, Fin, generator=g, dtype=dtype, requires_grad=True)
        self.Ws = [torch.randn(Fin if l == 0 else C, C, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True) for l in range(L)]
        self.an = [torch.randn(H, C // H, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True) for _ in range(L)]
        self.as_ = [torch.randn(H, C // H, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True) for _ in range(L)]
        self.W0 = torch.randn(Hd, C, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True)
        self.b0 = torch.zeros(Hd, dtype=dtype, requires_grad=True)
        self.W1 = torch.zeros(1, Hd, dtype=dtype)
        self.Ps = min(pair_sample, 2 * self.E)
        self.src = torch.randint(0, N, (self.Ps,), generator=g)
        self.dst = torch.randint(0, N, (self.Ps,), generator=g)
        self.lab = torch.randint(0, 2, (self.Ps,), generator=g)
        self.leaves = [self.x] + self.Ws + self.an + self.as_ + [self.W0, self.b0]
        self.opt = torch.optim.Adam(self.leaves, lr=1e-3, weight_decay=5e-4)          # train.py:207

    def step(self):
        """-> (seconds in the GAT layers fwd+bwd, seconds in scoring + loss fwd+bwd, seconds in Adam)"""
        O, dtype = self.O, self.dtype
        orig_t = O._t
        O._t = lambda a, dt=dtype: orig_t(a, dt)
        try:
            t0 = time.perf_counter()
            h = self.x
            for l in range(self.L):
                h = O.gat_layer(h, self.Ws[l], self.an[l], self.as_[l], self.rowptr, self.col, self.H)
            t1 = time.perf_counter()
            out = O.link_predictor(h[self.src], h[self.dst], [self.W0, self.W1], [self.b0, torch.zeros(1, dtype=dtype)])
            loss = torch.nn.functional.nll_loss(out, self.lab)
            t2 = time.perf_counter()
            gs = torch.autograd.grad(loss, [h, self.W0, self.b0])
            t3 = time.perf_counter()
            gl = torch.autograd.grad(h, [self.x] + self.Ws + self.an + self.as_, gs[0])
            t4 = time.perf_counter()
            for p_, g_ in zip(self.leaves, list(gl) + [gs[1], gs[2]]):
                p_.grad = g_
            self.opt.step()
            t5 = time.perf_counter()
        finally:
            O._t = orig_t
        return (t1 - t0) + (t4 - t3), (t2 - t1) + (t3 - t2), t5 - t4


def cpu_steps(wl, rows, cols, steps, warmup, subsampled, pair_sample=131072):
    """`warmup` untimed + `steps` timed CPU steps on a bounded sample of the workload -> cpu_baseline dict (value = the
    whole-workload rate the sample extrapolates to, from the mean step)."""
    if subsampled:
        sub_n = wl["n_nodes"] // 100 + 1
        wl = dict(wl, n_nodes=sub_n)
        sample_note = "1% node-induced subsample of the graph (every 100th node id); "
    else:
        sample_note = "full graph for the GAT layers; "
    m = (rows < 256) & (cols < 256)        # first-call initialisation of the torch CPU thread pools, off the clock
    CpuStep(dict(wl, n_nodes=256), rows[m], cols[m], pair_sample=1024).step()
    job = CpuStep(wl, rows, cols, pair_sample)
    for _ in range(warmup):
        job.step()
    ts = [job.step() for _ in range(steps)]
    t_gat, t_score, t_adam = (float(np.mean([t[k] for t in ts])) for k in range(3))
    E, Ps = job.E, job.Ps
    P = 2 * E
    t_step = t_gat + t_score * (P / Ps) + t_adam
    return {"value": E * wl["layers"] / t_step, "unit": "edges/s", "cores": os.cpu_count(), "kind": "port",
            "sample": sample_note + f"scoring timed on {Ps} of {P} pairs and scaled; oracle port (sparse restatement, torch CPU "
                      f"fp32, {os.cpu_count()} threads); mean of {steps} steps after {warmup} warm-up: gat {t_gat:.2f}s + score "
                      f"{t_score:.2f}s (sample) + Adam {t_adam:.3f}s",
            "ms_per_step_est": t_step * 1e3}


def cpu_baseline(args, wl, rows, cols, subsampled=False):
    return cpu_steps(wl, rows, cols, steps=2, warmup=1, subsampled=subsampled)


def run_reference(args):
    """CPU arm on the GPU arm's config: the same workload at the same N (weak scaling: the N-times larger graph is N
    copies of one rank's share, so one share is timed and the rate is the share's rate -- the CPU does not scale with N)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    strong = wl["graph"].startswith("R-MAT")
    if strong:
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        r, c = rmat_graph_device(wl["n_nodes"], wl["n_edges"], 4, dev)
        E_glob = int(torch.unique(r * wl["n_nodes"] + c).numel())
        keep = (r % 100 == 0) & (c % 100 == 0)
        rows, cols = (r[keep] // 100).cpu().numpy(), (c[keep] // 100).cpu().numpy()
        n_glob, n_pos = wl["n_nodes"], min(int(r.numel()), wl["pos_pairs"] // world) * world
        del r, c, keep
    else:
        rows, cols = make_graph_host(wl)
        E1 = int(np.unique(rows.astype(np.int64) * wl["n_nodes"] + cols.astype(np.int64)).size)   # Graph.from_coo coalesces
        E_glob, n_glob, n_pos = E1 * world, wl["n_nodes"] * world, int(rows.size) * world
    warm = max(0, min(args.warmup, 3))
    cb = cpu_steps(wl, rows, cols, steps=max(1, args.steps), warmup=warm, subsampled=strong,
                   pair_sample=(2 * rows.size if args.full_pairs else 131072))
    if world > 1 and not strong:
        cb["sample"] += f"; weak scaling at N={world}: one rank's share of the {world}x graph timed (the host does not scale with N)"
    out = {"impl": "reference", "metric": "gat_fwd_bwd_layer_edges_per_sec", "value": cb["value"], "unit": "edges/s",
           "n_gpus": world, "steps": max(1, args.steps), "warmup": warm, "ms_per_step": cb["ms_per_step_est"] * (1 if strong else world),
           "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": workload_config(args.workload, wl, n_glob, E_glob, 2 * n_pos),
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="ddi", choices=list(WORKLOADS) + ["flow", "flow-ours", "yearly", "ours"])
    ap.add_argument("--scale", type=float, default=1.0, help="--workload ours: shrink the per-GPU share (nodes, edges, sweep)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="ddi (1 GPU) / flow / flow-ours / yearly: launch every kernel from the host instead of replaying one "
                         "captured CUDA graph")
    ap.add_argument("--contraction-free", action="store_true",
                    help="headline step with the contraction-free scorer backward (exploits the one-hot d scores of the nll "
                         "read-out); default: the general tensor-core backward, the contraction-free step reported beside it")
    ap.add_argument("--dense-nll-bwd", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: exchange over NVLink peer memory with this library's kernels (default) or NCCL collectives")
    ap.add_argument("--no-year-streams", action="store_true", help="yearly: run the four yearly graphs one after the other")
    ap.add_argument("--no-strong-scaling", action="store_true",
                    help="skip the R-MAT (BASELINE.json configs[3]) block the default workload appends to its line")
    ap.add_argument("--full-pairs", action="store_true", help="--impl reference: score every pair instead of a 131072-pair sample")
    ap.add_argument("--unfused-loss", action="store_true",
                    help="separate nll_loss op: the dense d scores tensor is written and re-read (default: fused into the scorer backward)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.workload == "ours":
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the dense reference (Ours.py) cannot be instantiated at 10 M x 10 M; "
                                  "its CPU arm is timed on --workload flow-ours (BASELINE.json configs[0])"}))
        else:
            run_msha(args)
    elif args.workload == "yearly":
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                src, dst, _, _ = flow_graph(seed=2015)
                cb = yearly_cpu_baseline(src, dst, 39179, 32, 8, 4096, reps=max(1, min(args.steps, 3)))
                print(json.dumps({"impl": "reference", "metric": "gat_fwd_bwd_layer_edges_per_sec", "value": cb["value"],
                                  "unit": "edges/s", "n_gpus": 1, "steps": args.steps, "warmup": 1, "higher_is_better": True,
                                  "ms_per_step": cb["ms_per_step_est"], "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                                  "data": "synthetic", "config": {"workload": args.workload}, "cpu_baseline": cb,
                                  "e2e": {"value": cb["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        else:
            run_yearly(args)
    elif args.workload.startswith("flow"):
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                src, dst, _, _ = flow_graph()
                cb = flow_cpu_baseline(src, dst, 39179, 32, 64, reps=max(1, min(args.steps, 3)))
                print(json.dumps({"impl": "reference", "metric": "gat_fwd_bwd_layer_edges_per_sec", "value": cb["value"],
                                  "unit": "edges/s", "n_gpus": 1, "steps": args.steps, "warmup": 1, "higher_is_better": True,
                                  "ms_per_step": cb["ms_per_step_est"], "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                                  "data": "synthetic", "config": {"workload": args.workload}, "cpu_baseline": cb,
                                  "e2e": {"value": cb["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        else:
            run_flow(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
