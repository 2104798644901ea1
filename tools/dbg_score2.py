import sys, torch, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import msha_gnn_b200 as mg
from oracle import msha_oracle as O
DEV='cuda:0'
P, C, Hd, Nn = 70000, 256, 256, 4267
g = torch.Generator().manual_seed(P)
lp = mg.LinkPredictor("mlp", C, Hd, 1, 2, 0.0).to(DEV)
h = (torch.randn(Nn, C, generator=g) * 0.5).to(DEV).requires_grad_(True)
src = torch.randint(0, Nn, (P,), generator=g).to(DEV)
dst = torch.randint(0, Nn, (P,), generator=g).to(DEV)
G = torch.randn(P, Hd, generator=g).to(DEV)
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
res = {}
for fused in (True, False):
    lp.fused = fused
    h.grad = None; lp.zero_grad()
    out = lp.forward_pairs(h, h, src, dst)
    (out * G).sum().backward()
    res[fused] = (out.detach().clone(), h.grad.clone(), lp.lins[0].weight.grad.clone(), lp.lins[0].bias.grad.clone())
# torch fp64 reference on GPU
hd = h.detach().double().requires_grad_(True)
W = lp.lins[0].weight.detach().double().requires_grad_(True); b = lp.lins[0].bias.detach().double().requires_grad_(True)
ref = torch.sigmoid(torch.relu((hd[src] * hd[dst]) @ W.t() + b))
(ref * G.double()).sum().backward()
for fused in (True, False):
    o, gh, gw, gb = res[fused]
    print("fused" if fused else "unfused", "out", rel(o, ref.detach()), "dh", rel(gh, hd.grad), "dW", rel(gw, W.grad), "db", rel(gb, b.grad))
d = (res[True][1].double() - hd.grad).abs()
print("worst rows", torch.topk(d.max(1).values, 5))
pre = ((hd[src] * hd[dst]) @ W.t() + b).detach()
o = res[True][0]
mis = ((pre > 0) != (o > 0.5))
print("misclassified relu:", int(mis.sum()), "of", mis.numel(), " small positive pre among them max:", float(pre[mis].abs().max()) if mis.any() else 0)
