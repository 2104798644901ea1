import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench, msha_gnn_b200 as mg
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'rmat-s']; dev = torch.device('cuda:0')
rows, cols = bench.make_graph_host(wl)
g = mg.Graph.from_coo(torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev), wl['n_nodes'], wl['n_nodes'])
g.attention_csc(); hr = g.hub_rows(); hc = g.hub_cols()
deg = g.degrees.float()
print("nnz", g.nnz, "max deg", int(deg.max()), "median", float(deg.median()), "hub rows", hr.n_hub, "segs", hr.n_segs, "hub cols", hc.n_hub, hc.n_segs)
torch.manual_seed(0)
conv = mg.GATConv(256, 32, 8).to(dev)
x = torch.rand(wl['n_nodes'], 256, device=dev, requires_grad=True)
for it in range(3):
    out = conv(x, g)
    out.sum().backward()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = conv(x, g); out.sum().backward(); e.record(); torch.cuda.synchronize()
print("layer fwd+bwd ms", a.elapsed_time(e))
