"""Host-side cost of one ddi training step: time to ENQUEUE steps without synchronising vs device time."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench, msha_gnn_b200 as mg
wl = bench.WORKLOADS['ddi']; dev = torch.device('cuda:0')
rows, cols = bench.make_graph_host(wl); E = rows.size
pos = torch.from_numpy(np.stack([rows, cols])).to(dev)
graph = mg.Graph.from_coo(pos[0], pos[1], wl['n_nodes'], wl['n_nodes']); graph.attention_csc()
torch.manual_seed(42)
model = mg.GATLinkModel(wl['feat'], wl['hidden'], wl['heads'], wl['layers'], wl['pred_hidden']).to(dev)
x = torch.nn.Parameter(torch.rand(wl['n_nodes'], wl['feat'], device=dev))
opt = torch.optim.Adam(list(model.parameters()) + [x], lr=1e-3, weight_decay=5e-4, fused=True)
labels = torch.cat([torch.ones(E, dtype=torch.int64, device=dev), torch.zeros(E, dtype=torch.int64, device=dev)])
def step(it):
    ns, nd = mg.functional.negative_sample(1000 + it, E, wl['n_nodes'], wl['n_nodes'], dev)
    src = torch.cat([pos[0], ns]); dst = torch.cat([pos[1], nd])
    opt.zero_grad(set_to_none=True)
    out = model(x, graph, src, dst)
    loss = mg.functional.nll_loss(out, labels)
    loss.backward(); opt.step()
    return loss
for i in range(5): step(i)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for i in range(N): step(10 + i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/N:.3f} ms/step (host), total {1e3*(t2-t0)/N:.3f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(10): step(100 + i)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
