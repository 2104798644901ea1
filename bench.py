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
def algorithmic_bytes(wl, E, P):
    N, C, H, F, Hd = wl["n_nodes"], wl["hidden"], wl["heads"], wl["feat"], wl["pred_hidden"]
    return {
        "msha_gat_fwd": E * (4 + 8 * H + 4 * C) + N * 4 * C,
        "msha_gat_bwd_rows": E * (4 + 8 * H + 4 * C + 4 * H) + N * 12 * C,
        "msha_spmm_csc": E * (8 + 8 * H + 4 * C) + N * 4 * C,
        # contraction-free scorer backward under the nll read-out: order + label + indices + one score sector, two row
        # gathers and two read-modify-write row scatters per pair (SURVEY 8d scoring model minus the dense dOut / out / G)
        "msha_score_mlp_nll_bwd_sparse": P * (4 + 8 + 16 + 32 + 8 * C + 16 * C),
        "msha_pair_gather_mul": P * (16 + 8 * C + 4 * C),
        "msha_pair_scatter_mul_add": P * (16 + 4 * C + 8 * C + 16 * C),
    }


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
        return [f, g, i]

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
def run_ours(args):
    import torch.distributed as dist
    import msha_gnn_b200 as mg
    from msha_gnn_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    hbm_peak, _, peak_src = load_peaks()
    if args.dense_nll_bwd:
        mg.functional.SPARSE_NLL_BWD = False

    # ---- data.  N = 1: the workload graph.  N > 1 (weak scaling): the graph grows with N -- N*n nodes, N*E edges,
    # same degree distribution -- and is partitioned by destination-node range; rank r generates the edges of its
    # own rows (columns anywhere), so the per-GPU work is the N = 1 work plus the per-layer all-gather /
    # reduce-scatter of the column-side tensors (msha_gnn_b200/dist.py).
    from msha_gnn_b200 import dist as mdist
    strong = wl["graph"].startswith("R-MAT")          # fixed graph partitioned over the ranks (BASELINE.json configs[3])
    if strong:
        n_glob = wl["n_nodes"]
        rows, cols = make_graph_host(wl)               # identical on every rank (seeded)
        if world > 1:
            rowptr_h = np.zeros(n_glob + 1, dtype=np.int64)
            np.cumsum(np.bincount(rows, minlength=n_glob), out=rowptr_h[1:])
            part = mdist.Partition.edge_balanced(torch.from_numpy(rowptr_h), world, rank)
            keep = (rows >= part.lo) & (rows < part.hi)
            rows, cols = rows[keep], cols[keep]
        else:
            part = mdist.Partition(n_glob, 1, 0)
        n_loc = part.n_local
    else:
        n_loc, n_glob = wl["n_nodes"], wl["n_nodes"] * world
        part = mdist.Partition(n_glob, world, rank)
        if world == 1:
            rows, cols = make_graph_host(wl)
        else:
            rows, cols = er_graph(n_loc, wl["n_edges"], 1 + rank, n_cols=n_glob)
            rows = rows + part.lo
    L = wl["layers"]
    n_pos = rows.size if wl["pos_pairs"] is None else min(rows.size, wl["pos_pairs"] // (world if strong else 1))
    pos_host = torch.from_numpy(np.stack([rows[:n_pos], cols[:n_pos]])).pin_memory()   # (2, n_pos) int64 positives (global ids)
    rows_d, cols_d = torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev)
    if world == 1:
        graph = mg.Graph.from_coo(rows_d, cols_d, n_loc, n_loc)
    else:
        graph = mdist.partition_graph(rows_d, cols_d, part)
    del rows_d, cols_d
    E = graph.nnz                                      # distinct directed edges owned by this rank
    graph.attention_csc()
    torch.manual_seed(42)                                                        # identical parameters on every rank
    model = mg.GATLinkModel(wl["feat"], wl["hidden"], wl["heads"], L, wl["pred_hidden"], dropout=0.0).to(dev)
    torch.manual_seed(100 + rank)
    x = torch.nn.Parameter(torch.rand(n_loc, wl["feat"], device=dev) * 0.1)      # learnable node features (GAT.py:42)
    params = list(model.parameters())
    opt = torch.optim.Adam(params + [x], lr=1e-3, weight_decay=5e-4, fused=True)   # train.py:207
    P = 2 * n_pos
    labels = torch.cat([torch.ones(n_pos, dtype=torch.int64, device=dev), torch.zeros(n_pos, dtype=torch.int64, device=dev)])
    pos_dev = pos_host.to(dev)
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
        if args.unfused_loss:
            if world == 1:
                out = model(x, graph, src, dst)                                  # (P, pred_hidden) sigmoid scores
            else:
                h = mdist.gat_encode(model.convs, x, graph, part)
                out = mdist.score_pairs(model.predictor, h, src, dst, part)
            loss = mg.functional.nll_loss(out, labels)                           # LLP.py:235 read-out shape
        elif world == 1:
            loss = model.loss(x, graph, src, dst, labels)                        # same read-out, d scores generated in-kernel
        else:
            h = mdist.gat_encode(model.convs, x, graph, part)
            loss = mdist.score_pairs(model.predictor, h, src, dst, part, target=labels)
        loss.backward()
        if world > 1:
            mdist.allreduce_gradients(params, world=world)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for it in range(args.warmup):
        step(it, pos_dev)
    # ---- device-resident timing
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.msha_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    trace = [] if os.environ.get("MSHA_BENCH_TRACE") else None
    for it in range(args.steps):
        t0 = time.perf_counter()
        loss = step(args.warmup + it, pos_dev)
        if trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            import gc
            trace.append((time.perf_counter() - t0, e, torch.cuda.memory_reserved() >> 20,
                          sum(g["collections"] for g in gc.get_stats())))
    ev1.record()
    barrier()
    launches = lib.msha_launch_count() - l0
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    if trace is not None and rank == 0:
        prev, gpu = ev0, []
        for rec in trace:
            gpu.append(round(prev.elapsed_time(rec[1]), 2))
            prev = rec[1]
        print("dev loop trace: host enqueue ms/step", [round(r[0] * 1e3, 2) for r in trace], "gpu ms/step", gpu,
              "reserved MiB", [r[2] for r in trace], "gc collections so far", [r[3] for r in trace], file=sys.stderr)
    # ---- end-to-end timing: pinned host batch -> device every step, loss back to the host
    # The pinned batch of step i+1 is copied on a second stream while step i computes (two device buffers, an event per
    # buffer) -- every step's host->device copy and its loss read-back stay inside the timed region.
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(pos_dev), torch.empty_like(pos_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])              # the step that last read this buffer has finished
            bufs[b].copy_(pos_host, non_blocking=True)
            ready[b].record(copy_stream)

    def host_fed_step(i, seed_it):
        b = i & 1
        torch.cuda.current_stream().wait_event(ready[b])
        if i + 1 < host_fed_total[0]:
            prefetch(i + 1)
        loss_ = step(seed_it, bufs[b])
        consumed[b].record()
        return float(loss_.item())

    host_fed_total = [2]
    for b in range(2):
        consumed[b].record()
    prefetch(0)
    for it in range(2):                # untimed: the first host-fed steps allocate the per-step staging tensors
        host_fed_step(it, 20_000 + it)
    barrier()
    host_fed_total[0] = args.steps
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    prefetch(0)
    for it in range(args.steps):
        loss_host = host_fed_step(it, args.warmup + args.steps + it)
    ev3.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = ev2.elapsed_time(ev3) / args.steps
    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    # ---- per-kernel profile pass (outside the timed region)
    roofline, kernels, gat_roofline = None, [], None
    if rank != 0:                      # the profile steps contain collectives: every rank must run them
        step(10_000, pos_dev)
        step(10_001, pos_dev)
    if rank == 0:
        with KernelTimer(ops) as kt:
            step(10_000, pos_dev)
            step(10_001, pos_dev)
        agg = kt.summary()
        ab = algorithmic_bytes(wl, E, P)
        tot = sum(v[1] for v in agg.values())
        for fname, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            per = ms / cnt
            row = {"call": fname, "launches_per_step": cnt // 2, "avg_ms": round(per, 4), "share": round(ms / tot, 4)}
            if fname in ab:
                row["algorithmic_GB"] = round(ab[fname] / 1e9, 4)
                row["GBps"] = round(ab[fname] / 1e6 / per, 1)
                row["frac_hbm"] = round(ab[fname] / 1e6 / per / hbm_peak, 4)
            kernels.append(row)
        # dense-contraction flops per call (fp32-equivalent 2*M*N*K; the 3xTF32 split issues 3x that on the tensor pipe)
        gemm_flops = sum(2.0 * sh[0] * sh[1] * sh[2] for f, a, b, sh in kt.records if f.startswith("msha_gemm") and len(sh) >= 3) / 2
        C_, Hd_ = wl["hidden"], wl["pred_hidden"]
        tensor_flops = {"msha_gemm_tf32x3": gemm_flops, "msha_score_mlp_fwd": 2.0 * P * C_ * Hd_,
                        "msha_score_mlp_bwd": 4.0 * P * C_ * Hd_, "msha_score_mlp_nll_bwd": 4.0 * P * C_ * Hd_}
        _, bf16_peak, _ = load_peaks()
        tf32_peak = bf16_peak / 2.0              # kind::tf32 runs at half the bf16 rate; bf16 peak = measured cuBLAS number
        for k in kernels:
            if k["call"] in tensor_flops:
                tf = tensor_flops[k["call"]] / (k["avg_ms"] * k["launches_per_step"] * 1e-3) / 1e12
                k["algorithmic_TFLOPs"] = round(tf, 1)
                k["issued_tf32_TFLOPs"] = round(3 * tf, 1)
                k["frac_tf32_peak"] = round(3 * tf / tf32_peak, 4)
        dom = kernels[0] if kernels else None
        traffic = None                     # DRAM bytes per launch of the dominant call, from the committed ncu capture
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f).get(args.workload, {})
            if dom and world == 1 and tr.get("pairs") == P and dom["call"] in tr:
                traffic = tr[dom["call"]]["bytes"]
        except (OSError, ValueError):
            pass
        if dom and dom["call"] in tensor_flops:
            roofline = {"bound": "tensor", "kernel": dom["call"], "achieved": dom["issued_tf32_TFLOPs"], "peak": tf32_peak,
                        "unit": "TFLOP/s", "frac": dom["frac_tf32_peak"], "traffic": traffic,
                        "algorithmic_fp32_equiv_TFLOPs": dom["algorithmic_TFLOPs"],
                        "frac_algorithmic": round(dom["algorithmic_TFLOPs"] / tf32_peak, 4),
                        "note": "achieved = tf32 flops issued per launch (3 per fp32-accurate product: 3xTF32 split) / CUDA-event "
                                "time; peak = measured bf16 cuBLAS peak / 2 (kind::tf32 rate); frac_algorithmic counts the "
                                "SURVEY 8d flops (2*P*C*Hd) once -- fp32-grade results (rel. err <= 1e-4) need the 3-way split, so its "
                                "ceiling is 1/3; " + peak_src,
                        "share_of_step": dom["share"]}
        elif dom and "GBps" in dom:
            roofline = {"bound": "hbm", "kernel": dom["call"], "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": dom["frac_hbm"], "traffic": traffic, "peak_source": peak_src,
                        "share_of_step": dom["share"]}
        # the graph-attention kernels of one layer (fwd + both backward passes) against the HBM roofline
        gat = [k for k in kernels if k["call"] in ("msha_gat_fwd", "msha_gat_bwd_rows", "msha_spmm_csc")]
        if gat:
            gb = sum(k["algorithmic_GB"] for k in gat)
            ms = sum(k["avg_ms"] for k in gat)
            gat_roofline = {"bound": "hbm", "kernels": [k["call"] for k in gat], "algorithmic_GB_per_layer": round(gb, 3),
                            "ms_per_layer": round(ms, 4), "achieved": round(gb / ms * 1e3, 1), "peak": hbm_peak,
                            "unit": "GB/s", "frac": round(gb / ms * 1e3 / hbm_peak, 4),
                            "note": "feature matrix (N*C*4 B) is L2-resident on this workload: fraction can exceed 1"}
        else:
            gat_roofline = None
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_edges = E_global * L
    out = {
        "metric": "gat_fwd_bwd_layer_edges_per_sec", "value": total_edges / (ms_dev / 1e3), "unit": "edges/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n_glob} nodes, {E_global} directed edges ({wl['graph']}), F={wl['feat']}, "
                               f"{L}-layer {wl['heads']}-head GAT C={wl['hidden']} + LinkPredictor(mlp,{wl['hidden']},"
                               f"{wl['pred_hidden']}) over P={P_global} pairs (positives + as many Philox negatives), nll loss, Adam",
                   "per_gpu": ("whole graph on one GPU" if world == 1 else
                               f"{'strong' if strong else 'weak'} scaling: graph of {n_glob} nodes / {E_global} edges partitioned by destination-node range, "
                               "per layer one NCCL all-gather (fwd) + reduce-scatter (bwd) of [Wh|s_nbr]; pairs data-parallel; "
                               "parameter gradients all-reduced"),
                   "l2_policy": "per-step working set (>= 4*P*pred_hidden bytes of scores) exceeds the 126 MB L2; no explicit flush",
                   "scorer_backward": ("separate nll_loss op + tensor-core backward" if args.unfused_loss else
                                       "tensor-core backward with the nll gradient generated in-kernel" if not mg.functional.SPARSE_NLL_BWD else
                                       "contraction-free: d scores of the nll read-out is one-hot per pair, so dZ = g_p * W0[label] and "
                                       "dW0 is a per-label sum (same gradients as the dense GEMMs, --dense-nll-bwd runs those)")},
        "pairs_per_sec": P_global / (ms_dev / 1e3),
        "e2e": {"value": total_edges / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(pos_host.numel() * 8), "d2h_bytes_per_step": 4,
                "input_pipeline": "pinned batch of step i+1 copied on a second stream while step i computes; loss.item() every step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "gat_layer_roofline": gat_roofline,
        "kernels": kernels,
        "loss": loss_host,
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(args, wl, rows, cols)

    print(json.dumps(out))
    if world > 1:
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
    out = {"metric": "gat_fwd_bwd_layer_edges_per_sec", "value": E * layers / (ms_dev / 1e3), "unit": "edges/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.workload}: 2015-shaped flow graph N={N}, M={M}, nnz={E} ({src.size} records), "
                                  f"{'Ours' if full else 'ablation3'}(in=128,out=64,H=2,p=0.5), batch {B}, nll_loss, Adam "
                                  "(train.py:206-232)",
                      "l2_policy": "graph and parameters (~50 MB) are L2-resident by nature of the workload; launch/latency bound"},
           "e2e": {"value": E * layers / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": 2 * B * 8, "d2h_bytes_per_step": 4},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": None, "kernels": kernels, "loss": loss_host, **graph_info}
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

    def step(batch):
        opt.zero_grad(set_to_none=True)
        total = None
        for k, yr in enumerate(years):
            s_i, r_i = batch[k, 0], batch[k, 1]
            h = yr["model"](yr["graph"])                                        # (N, M) log-probs, GAT.py:53-58
            loss = predictor.nll_loss_pairs(h, h, s_i, r_i, r_i)                # LLP.py:233-235
            total = loss if total is None else total + loss
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
def cpu_step_time(wl, rows, cols, pair_sample, reps=1, dtype=torch.float32):
    from oracle import msha_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    N, Fin, C, H, L, Hd = wl["n_nodes"], wl["feat"], wl["hidden"], wl["heads"], wl["layers"], wl["pred_hidden"]
    rowptr, col, _ = O.csr_from_coo(rows, cols, N, N)
    E = col.size
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, Fin, generator=g, dtype=dtype, requires_grad=True)
    Ws = [torch.randn(Fin if l == 0 else C, C, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True) for l in range(L)]
    an = [torch.randn(H, C // H, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True) for _ in range(L)]
    as_ = [torch.randn(H, C // H, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True) for _ in range(L)]
    W0 = torch.randn(Hd, C, generator=g, dtype=dtype).mul_(0.1).requires_grad_(True)
    b0 = torch.zeros(Hd, dtype=dtype, requires_grad=True)
    W1 = torch.zeros(1, Hd, dtype=dtype)
    Ps = min(pair_sample, 2 * E)
    src = torch.randint(0, N, (Ps,), generator=g)
    dst = torch.randint(0, N, (Ps,), generator=g)
    lab = torch.randint(0, 2, (Ps,), generator=g)
    orig_t = O._t
    O._t = lambda a, dt=dtype: orig_t(a, dt)
    try:
        t_gat = t_score = 0.0
        for _ in range(reps):
            t0 = time.perf_counter()
            h = x
            for l in range(L):
                h = O.gat_layer(h, Ws[l], an[l], as_[l], rowptr, col, H)
            t1 = time.perf_counter()
            out = O.link_predictor(h[src], h[dst], [W0, W1], [b0, torch.zeros(1, dtype=dtype)])
            loss = torch.nn.functional.nll_loss(out, lab)
            t2 = time.perf_counter()
            gs = torch.autograd.grad(loss, [h, W0, b0])
            t3 = time.perf_counter()
            torch.autograd.grad(h, [x] + Ws + an + as_, gs[0])
            t4 = time.perf_counter()
            t_gat += (t1 - t0) + (t4 - t3)
            t_score += (t2 - t1) + (t3 - t2)
    finally:
        O._t = orig_t
    return t_gat / reps, t_score / reps, E, Ps


def cpu_baseline(args, wl, rows, cols):
    if wl["n_edges"] > 5_000_000:      # 1 % node-induced subsample for the big graphs
        keep = (rows % 100 == 0) & (cols % 100 == 0)
        sub_n = wl["n_nodes"] // 100 + 1
        wl = dict(wl, n_nodes=sub_n)
        rows, cols = rows[keep] // 100, cols[keep] // 100
        sample_note = "1% node-induced subsample of the graph; "
    else:
        sample_note = "full graph for the GAT layers; "
    # untimed warm-up on a sliver of the graph (first-call initialisation of the torch CPU thread pools)
    m = (rows < 256) & (cols < 256)
    cpu_step_time(dict(wl, n_nodes=256), rows[m], cols[m], pair_sample=1024)
    t_gat, t_score, E, Ps = cpu_step_time(wl, rows, cols, pair_sample=131072)
    P = 2 * E
    t_step = t_gat + t_score * (P / Ps)
    return {"value": E * wl["layers"] / t_step, "unit": "edges/s", "cores": os.cpu_count(), "kind": "port",
            "sample": sample_note + f"scoring timed on {Ps} of {P} pairs and scaled; oracle port (sparse restatement, "
                      f"torch CPU fp32); gat {t_gat:.2f}s + score {t_score:.2f}s measured",
            "ms_per_step_est": t_step * 1e3}


def reference_workload_name(name, wl, rows, cols):
    """The GPU arm's config.workload string (run_ours) rebuilt from the host-side graph: same workload, same wording."""
    n = wl["n_nodes"]
    if rows.size <= 20_000_000:
        E = int(np.unique(rows.astype(np.int64) * n + cols.astype(np.int64)).size)      # Graph.from_coo coalesces duplicates
    else:
        E = int(rows.size)                                                             # not coalesced here (a 100 M-key sort)
    n_pos = wl["pos_pairs"] or E
    return (f"{name}: {n} nodes, {E} directed edges ({wl['graph']}), F={wl['feat']}, {wl['layers']}-layer {wl['heads']}-head "
            f"GAT C={wl['hidden']} + LinkPredictor(mlp,{wl['hidden']},{wl['pred_hidden']}) over P={2 * n_pos} pairs "
            "(positives + as many Philox negatives), nll loss, Adam")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    rows, cols = make_graph_host(wl)
    times = []
    n = max(1, min(args.steps, 3))
    for i in range(min(args.warmup, 1) + n):
        cb = cpu_baseline(args, wl, rows, cols)
        if i >= min(args.warmup, 1):
            times.append(cb)
    best = max(times, key=lambda c: c["value"])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": "gat_fwd_bwd_layer_edges_per_sec", "value": best["value"], "unit": "edges/s",
           "n_gpus": world, "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": best["ms_per_step_est"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": reference_workload_name(args.workload, wl, rows, cols),
                      "per_gpu": "CPU arm: oracle port on the host cores (rank 0 only), bounded sample -- see cpu_baseline.sample"},
           "cpu_baseline": best,
           "e2e": {"value": best["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="ddi", choices=list(WORKLOADS) + ["flow", "flow-ours", "yearly"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="flow / flow-ours / yearly: launch every kernel from the host instead of replaying one captured CUDA graph")
    ap.add_argument("--dense-nll-bwd", action="store_true",
                    help="scorer backward on the tensor cores even under the nll read-out (default: the contraction-free "
                         "kernel that exploits the one-hot d scores)")
    ap.add_argument("--unfused-loss", action="store_true",
                    help="separate nll_loss op: the dense d scores tensor is written and re-read (default: fused into the scorer backward)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.workload == "yearly":
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
