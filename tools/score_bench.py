import sys, torch
sys.path.insert(0, '.')
import msha_gnn_b200 as mg
from msha_gnn_b200 import ops, functional as Fn
DEV='cuda:0'
P, C, Hd, Nn = 2669778, 256, 256, 4267
g0 = torch.Generator().manual_seed(1)
h = (torch.randn(Nn, C, generator=g0) * 0.5).to(DEV).requires_grad_(True)
src = torch.randint(0, Nn, (P,), generator=g0).to(DEV); dst = torch.randint(0, Nn, (P,), generator=g0).to(DEV)
lp = mg.LinkPredictor("mlp", C, Hd, 1, 2, 0.0).to(DEV)
dout = torch.randn(P, Hd, device=DEV)
def timeit(fn, name, reps=5):
    fn(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(e)/reps:.3f} ms")
out = lp.forward_pairs(h, h, src, dst)
timeit(lambda: lp.forward_pairs(h, h, src, dst), "fused fwd")
def bwd():
    h.grad = None
    torch.autograd.grad(out, [h, lp.lins[0].weight, lp.lins[0].bias], dout, retain_graph=True)
timeit(bwd, "fused bwd (dz+scatter kernel + dw kernel + memsets)")
