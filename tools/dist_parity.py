"""torchrun --nproc-per-node N tools/dist_parity.py : partitioned GAT encode + pair scoring over NCCL (with and without
the early reduce-scatter of d Wh) against the whole-graph result on one GPU -- outputs, input and parameter gradients.

Two constructions.  "positive": non-negative features / weights and a positive scorer bias keep every relu / LeakyReLU
pre-activation away from 0, so the comparison measures arithmetic only (tolerance 1e-4; the attention vectors' gradients
are excluded: with all logits on one LeakyReLU branch softmax shift-invariance makes them exactly 0 / pure cancellation).
"signed": ordinary random parameters; a handful of the 77 M pre-activations lie within fp32 round-off of 0 and fall on
either side in the two summation orders, which moves gradients by ~1e-3 .. 1e-2 of their norm -- more flips the more ranks
the sums are split over; every exchange mode (NCCL or peer memory) shows the SAME figure, the reference has the same
sensitivity -- tolerance 2e-2, all parameters included."""
import os, sys
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")       # peer path: kernels wait on kernels (msha_gnn_b200/peer.py)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msha_gnn_b200 as mg
from msha_gnn_b200 import dist as md
from msha_gnn_b200 import peer, dist_p2p as mp2p

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(7)
N, F, H, d, P = 20011, 64, 8, 32, 300_000
deg = np.minimum(1 + (rng.pareto(1.2, N) * 8).astype(np.int64), 3000)          # power-law rows: hub segments on every rank
rows = np.repeat(np.arange(N), deg); cols = rng.integers(0, N, rows.size)
key = np.unique(rows * N + cols); rows, cols = key // N, key % N
src, dst = rng.integers(0, N, P), rng.integers(0, N, P)
src_d, dst_d = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
G = torch.from_numpy(rng.standard_normal((P, 256)).astype(np.float32)).to(dev)
g_full = mg.Graph.from_coo(torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev), N, N)
part = md.Partition(N, world, rank)
keep = (rows >= part.lo) & (rows < part.hi)
pg = md.partition_graph(torch.from_numpy(rows[keep]).to(dev), torch.from_numpy(cols[keep]).to(dev), part)
lo, hi = (P * rank) // world, (P * (rank + 1)) // world
fabric = peer.SymmFabric() if world > 1 else None
labels = torch.from_numpy(rng.integers(0, 2, P)).to(dev)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


ok = True
for construction, tol in (("positive", 1e-4), ("signed", 2e-2)):
    torch.manual_seed(3)
    model = mg.GATLinkModel(F, H * d, H, 2, 256, dropout=0.0).to(dev)
    if construction == "positive":
        with torch.no_grad():
            for p_ in model.convs.parameters():
                p_.abs_()
            model.predictor.lins[0].weight.abs_().mul_(1e-4)
            model.predictor.lins[0].bias.fill_(0.5)
    x_full = torch.from_numpy(rng.random((N, F)).astype(np.float32) * 0.3).to(dev)
    xf = x_full.clone().requires_grad_(True)
    out_ref = model(xf, g_full, src_d, dst_d)
    (out_ref * G).sum().backward()
    gref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}   # the last Linear is never applied
    gx_ref = xf.grad.clone()
    model.zero_grad(set_to_none=True)
    modes = [("nccl+early-rs", None), ("nccl", None)]
    if fabric is not None:
        modes += [("p2p-flat", (1 << 40, 0, 4)), ("p2p-rows", (0, 0, 4)), ("p2p-halo", (0, 1, 4)), ("p2p-halo-seq", (0, 1, 1))]
    for overlap, thresholds in modes:
        xl = x_full[part.lo:part.hi].clone().requires_grad_(True)
        if thresholds is None:
            h = md.gat_encode(model.convs, xl, pg, part, overlap=(overlap == "nccl+early-rs"))
            out = md.score_pairs(model.predictor, h, src_d[lo:hi], dst_d[lo:hi], part)
        else:
            mp2p.PIPELINE_MIN_BLOCK_BYTES, mp2p.HALO = thresholds[0], bool(thresholds[1])
            p2p = md.P2P(fabric.group, part, chunks=thresholds[2])
            h = md.gat_encode_p2p(model.convs, xl, pg, part, p2p, score_key="score_h")
            out = md.score_pairs(model.predictor, h, src_d[lo:hi], dst_d[lo:hi], part, p2p=p2p)
        (out * G[lo:hi]).sum().backward()
        if thresholds is not None:
            p2p.pg.check()
        md.allreduce_gradients([p for p in model.parameters() if p.grad is not None], world=world)
        errs = {"out": rel(out.detach(), out_ref.detach()[lo:hi]), "dx": rel(xl.grad, gx_ref[part.lo:part.hi])}
        for n, p in model.named_parameters():
            if n in gref:
                errs["d" + n] = rel(p.grad, gref[n])
        checked = {k: v for k, v in errs.items() if not (construction == "positive" and (".a_nbr" in k or ".a_self" in k))}
        worst = torch.tensor([max(checked.values())], device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{construction:8s} mode={overlap} world={world} hub_rows={pg.hub_rows().n_hub} worst checked rel err "
                  f"{float(worst):.3e} (tol {tol:g})", {k: f"{v:.1e}" for k, v in errs.items()}, flush=True)
        ok = ok and float(worst) < tol
        model.zero_grad(set_to_none=True)
    # mean nll read-out: the ranks' shares add up to the single-GPU loss, gradients equal the single-GPU ones
    xf = x_full.clone().requires_grad_(True)
    hf = model.encode(xf, g_full)
    loss_ref = model.predictor.nll_loss_pairs(hf, hf, src_d, dst_d, labels)
    loss_ref.backward()
    gW_ref, gx_ref2 = model.convs[0].W.grad.clone(), xf.grad.clone()
    model.zero_grad(set_to_none=True)
    xl = x_full[part.lo:part.hi].clone().requires_grad_(True)
    h = md.gat_encode(model.convs, xl, pg, part)
    loss = md.score_pairs(model.predictor, h, src_d[lo:hi], dst_d[lo:hi], part, target=labels[lo:hi], global_pairs=P)
    loss.backward()
    md.allreduce_gradients(list(model.parameters()), world=world)
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    e = torch.tensor([abs(float(tot) - float(loss_ref)) / abs(float(loss_ref)), rel(model.convs[0].W.grad, gW_ref),
                      rel(xl.grad, gx_ref2[part.lo:part.hi])], device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{construction:8s} mean-nll over ranks: loss rel err {float(e[0]):.2e}, dW0 {float(e[1]):.2e}, dx {float(e[2]):.2e}", flush=True)
    ok = ok and float(e.max()) < tol
    model.zero_grad(set_to_none=True)
dist.destroy_process_group()
if rank == 0:
    print("PARITY OK" if ok else "PARITY FAILED", flush=True)
sys.exit(0 if ok else 1)
