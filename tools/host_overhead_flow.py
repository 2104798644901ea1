"""Host-side cost of one flow (ablation3 / Ours) training step: enqueue time vs total, plus a cProfile of the host side."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench, msha_gnn_b200 as mg
from msha_gnn_b200 import ops
full = len(sys.argv) > 1 and sys.argv[1] == 'ours'
dev = torch.device('cuda:0'); N, M, B = 39179, 32, 64
src, dst, city, prov = bench.flow_graph()
graph = mg.Graph.from_coo(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), N, M); graph.attention_csc()
gdp = {str(i): 0.05 for i in range(N)}
torch.manual_seed(42)
cls = mg.Ours if full else mg.ablation3
model = cls(in_features=128, out_features=64, n_classes=M, n_heads=2, dropout=0.5, gdp=gdp, Scount=N, Rcount=M).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4, fused=True)
city_d, prov_d = torch.from_numpy(city).to(dev), torch.from_numpy(prov).to(dev)
rng = np.random.default_rng(0)
rec_idx = rng.integers(0, src.size, (64, B))
batches = torch.from_numpy(np.stack([src[rec_idx], dst[rec_idx]], axis=1)).to(dev)
model.train()
def step(b):
    s_i, r_i = b[0], b[1]
    opt.zero_grad(set_to_none=True)
    out = model(graph, city_d, prov_d, s_i) if full else model(graph, None, None, s_i)
    loss = torch.nn.functional.nll_loss(out[s_i], r_i)
    loss.backward(); opt.step()
    return loss
for i in range(5): step(batches[i])
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for i in range(n): step(batches[5 + i])
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/n:.3f} ms/step (host), total {1e3*(t2-t0)/n:.3f} ms/step")
with bench.KernelTimer(ops) as kt:
    step(batches[30]); step(batches[31])
agg = kt.summary()
print("device time in C-ABI calls per step: %.3f ms over %d calls" % (sum(v[1] for v in agg.values()) / 2, sum(v[0] for v in agg.values()) // 2))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(10): step(batches[40 + i])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
