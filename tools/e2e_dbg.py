import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench, msha_gnn_b200 as mg
wl = bench.WORKLOADS['ddi']; dev = torch.device('cuda:0')
rows, cols = bench.make_graph_host(wl); E = rows.size
pos_host = torch.from_numpy(np.stack([rows, cols])).pin_memory()
pos = pos_host.to(dev)
graph = mg.Graph.from_coo(pos[0], pos[1], wl['n_nodes'], wl['n_nodes']); graph.attention_csc()
torch.manual_seed(42)
model = mg.GATLinkModel(wl['feat'], wl['hidden'], wl['heads'], wl['layers'], wl['pred_hidden']).to(dev)
x = torch.nn.Parameter(torch.rand(wl['n_nodes'], wl['feat'], device=dev) * 0.1)
opt = torch.optim.Adam(list(model.parameters()) + [x], lr=1e-3, weight_decay=5e-4, fused=True)
labels = torch.cat([torch.ones(E, dtype=torch.int64, device=dev), torch.zeros(E, dtype=torch.int64, device=dev)])
def step(it, pos):
    ns, nd = mg.functional.negative_sample(1000 + it, E, wl['n_nodes'], wl['n_nodes'], dev)
    src = torch.cat([pos[0], ns]); dst = torch.cat([pos[1], nd])
    opt.zero_grad(set_to_none=True)
    out = model(x, graph, src, dst)
    loss = mg.functional.nll_loss(out, labels)
    loss.backward(); opt.step()
    return loss
for i in range(3): step(i, pos)
torch.cuda.synchronize()
for mode in ("async", "sync_item", "h2d+item"):
    ts = []
    for i in range(8):
        t0 = time.perf_counter()
        p = pos_host.to(dev, non_blocking=True) if mode == "h2d+item" else pos
        l = step(10 + i, p)
        t1 = time.perf_counter()
        if mode != "async": l.item()
        t2 = time.perf_counter()
        ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t0) * 1e3, 2)))
    torch.cuda.synchronize()
    print(mode, "cpu-issue ms / total ms per step:", ts)
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.max_memory_allocated() / 1e9)
