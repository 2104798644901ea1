import sys, torch, numpy as np
sys.path.insert(0, '.')
import msha_gnn_b200 as mg
from msha_gnn_b200 import ops
ops.TC_MIN_WORK = 0
DEV = 'cuda:0'
def run(M, N, K, tA, tB):
    g = torch.Generator().manual_seed(1)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(DEV)
    B = torch.randn((N, K) if tB else (K, N), generator=g).to(DEV)
    try:
        out = ops.gemm(A, B, transA=tA, transB=tB)
        torch.cuda.synchronize()
    except Exception as e:
        print(M, N, K, tA, tB, 'EXC', e); return
    a = A.cpu().double().t() if tA else A.cpu().double()
    b = B.cpu().double().t() if tB else B.cpu().double()
    ref = a @ b
    o = out.cpu().double()
    err = (o - ref).abs()
    print(f"M{M} N{N} K{K} tA={tA} tB={tB}: max rel err {float(err.max()/ref.abs().max()):.3e}  out absmax {float(o.abs().max()):.3f} nnz frac {float((o!=0).float().mean()):.3f}")
    # block structure: 32-row x 32-col blocks
    mb, nb = (M + 31) // 32, (N + 31) // 32
    bad = [(i, j) for i in range(mb) for j in range(nb) if float(err[i*32:(i+1)*32, j*32:(j+1)*32].max()) > 1e-4 * float(ref.abs().max())]
    print("   bad 32x32 blocks:", len(bad), "of", mb * nb, bad[:12])
for (tA, tB) in [(False, True), (False, False), (True, True), (True, False)]:
    for (M, N, K) in [(128, 256, 32), (128, 64, 32), (128, 64, 8), (128, 128, 64), (256, 256, 256)]:
        run(M, N, K, tA, tB)
