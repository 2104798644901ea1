import sys, os, torch, json
sys.path.insert(0, '.')
import bench, numpy as np
import msha_gnn_b200 as mg
from torch.profiler import profile, ProfilerActivity
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
for i in range(3): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3): step(10 + i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
