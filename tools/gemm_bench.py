import sys, torch
sys.path.insert(0, '.')
import msha_gnn_b200 as mg
from msha_gnn_b200 import ops
DEV = 'cuda:0'
P, C, Hd = 2669778, 256, 256
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
Z = torch.randn(P, C, device=DEV)
W = torch.randn(Hd, C, device=DEV) * 0.05
b = torch.zeros(Hd, device=DEV)
G = torch.randn(P, Hd, device=DEV)
out = torch.empty(P, Hd, device=DEV)
def t(fn, name, flops):
    fn(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / reps
    print(f"{name}: {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s (fp32-equivalent)  {3*flops/ms/1e9:.1f} TF tf32 issued")
fl = 2.0 * P * C * Hd
if which in ("all", "fwd"): t(lambda: ops.gemm(Z, W, transB=True, bias=b, act=ops.ACT_SIGMOID_RELU, out=out), "fwd  NT act", fl)
if which in ("all", "fwdnoact"): t(lambda: ops.gemm(Z, W, transB=True, out=out), "fwd  NT noact", fl)
if which in ("all", "dz"): t(lambda: ops.gemm(G, W, out=out), "dZ   NN", fl)
if which in ("all", "dw"): t(lambda: ops.gemm(G, Z, transA=True), "dW   TN splitK", fl)
if which in ("all", "small"):
    X = torch.randn(4267, 256, device=DEV); Wg = torch.randn(256, 256, device=DEV)
    t(lambda: ops.gemm(X, Wg), "Wh 4267x256x256", 2.0 * 4267 * 256 * 256)
