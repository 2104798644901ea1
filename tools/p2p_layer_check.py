"""One GAT layer on emulated ranks: pipelined peer path vs flat peer path vs whole graph -- out, d x, d W, d a (diagnostic)."""
import os, sys, copy, threading
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import msha_gnn_b200 as mg
from msha_gnn_b200 import dist as md, peer, dist_p2p as mp2p

dev = torch.device("cuda:0")
W_ = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(7)
N, F, H, d = 20011, 64, 8, 32
deg = np.minimum(1 + (rng.pareto(1.2, N) * 8).astype(np.int64), 3000)
rows = np.repeat(np.arange(N), deg); cols = rng.integers(0, N, rows.size)
key = np.unique(rows * N + cols); rows, cols = key // N, key % N
torch.manual_seed(3)
L_ = int(sys.argv[2]) if len(sys.argv) > 2 else 1
conv = torch.nn.ModuleList([mg.GATConv(F if l == 0 else H * d, d, H) for l in range(L_)]).to(dev)
with torch.no_grad():
    for p_ in conv.parameters():
        p_.abs_()
x = torch.from_numpy(rng.random((N, F)).astype(np.float32) * 0.3).to(dev)
G = torch.from_numpy(rng.standard_normal((N, H * d)).astype(np.float32)).to(dev)
g_full = mg.Graph.from_coo(torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev), N, N)
xf = x.clone().requires_grad_(True)
cf = copy.deepcopy(conv)
ref = xf
for c_ in cf:
    ref = c_(ref, g_full)
(ref * G).sum().backward()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for name, thr in (("flat", (1 << 40, 1 << 40)), ("pipelined-rows", (0, 0)), ("pipelined-halo", (0, 1))):
    mp2p.PIPELINE_MIN_BLOCK_BYTES = thr[0]
    mp2p.HALO = bool(thr[1])
    fab = peer.LocalFabric(W_, dev)
    res = [None] * W_

    def body(r):
        with torch.cuda.stream(fab.streams[r]):
            part = md.Partition(N, W_, r)
            keep = (rows >= part.lo) & (rows < part.hi)
            pg = md.partition_graph(torch.from_numpy(rows[keep]).to(dev), torch.from_numpy(cols[keep]).to(dev), part)
            p2p = md.P2P(fab.groups[r], part)
            c = copy.deepcopy(conv)
            xl = x[part.lo:part.hi].clone().requires_grad_(True)
            out = md.gat_encode_p2p(c, xl, pg, part, p2p)
            (out * G[part.lo:part.hi]).sum().backward()
            res[r] = (out.detach(), xl.grad, c)
        fab.streams[r].synchronize()
    ts = [threading.Thread(target=body, args=(r,)) for r in range(W_)]
    [t.start() for t in ts]; [t.join() for t in ts]
    torch.cuda.synchronize()
    [g.check() for g in fab.groups]
    out = torch.cat([r[0] for r in res]); gx = torch.cat([r[1] for r in res])
    msg = [name, "out", f"{rel(out, ref.detach()):.2e}", "dx", f"{rel(gx, xf.grad):.2e}"]
    for l in range(L_):
        gW = sum(r[2][l].W.grad for r in res); ga = sum(r[2][l].a_nbr.grad for r in res); gs = sum(r[2][l].a_self.grad for r in res)
        msg += [f"L{l}: dW {rel(gW, cf[l].W.grad):.2e} da_nbr {rel(ga, cf[l].a_nbr.grad):.2e} da_self {rel(gs, cf[l].a_self.grad):.2e}"]
    print(*msg, flush=True)
