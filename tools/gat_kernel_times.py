import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench, msha_gnn_b200 as mg
from msha_gnn_b200 import ops
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'ddi']; dev = torch.device('cuda:0')
rows, cols = bench.make_graph_host(wl)
g = mg.Graph.from_coo(torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev), wl['n_nodes'], wl['n_nodes'])
g.attention_csc(); g.hub_rows(); g.hub_cols()
torch.manual_seed(0)
conv = mg.GATConv(256, 32, 8).to(dev)
x = torch.rand(wl['n_nodes'], 256, device=dev, requires_grad=True)
for it in range(3):
    conv(x, g).sum().backward()
torch.cuda.synchronize()
with bench.KernelTimer(ops) as kt:
    for it in range(5):
        conv(x, g).sum().backward()
for f, (c, ms) in sorted(kt.summary().items(), key=lambda kv: -kv[1][1])[:4]:
    print(f"{f}: {ms/c:.4f} ms")
