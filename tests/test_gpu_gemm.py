"""GPU parity of the two GEMM kernels (SIMT fp32 and tcgen05 3xTF32) against fp64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import msha_gnn_b200 as mg
from msha_gnn_b200 import ops
from conftest import rel_err

DEV = "cuda:0"
# fp32 FMA accumulates ~sqrt(K)*2^-24; the 3xTF32 split adds ~2^-21 per product.  North-star tolerance is 1e-4.
GEMM_TOL = 2e-5

SHAPES = [
    # M, N, K
    (128, 256, 32), (128, 64, 256), (4267, 256, 256), (300, 200, 100), (1000, 96, 260), (129, 257, 36), (5000, 32, 64),
    (64, 128, 4096), (70000, 256, 256), (40000, 64, 128),      # several tiles per persistent CTA
]


def _ref(A, B, tA, tB, bias=None):
    a = A.double().t() if tA else A.double()
    b = B.double().t() if tB else B.double()
    r = a @ b
    return r + bias.double() if bias is not None else r


@pytest.mark.parametrize("backend", ["simt", "tcgen05"])
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_layouts(backend, tA, tB, M, N, K, monkeypatch):
    monkeypatch.setattr(ops, "GEMM_BACKEND", backend)
    monkeypatch.setattr(ops, "TC_MIN_WORK", 0)
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).to(DEV)
    B = torch.randn((N, K) if tB else (K, N), generator=g).to(DEV)
    if backend == "tcgen05" and (A.shape[1] % 4 or B.shape[1] % 4):
        pytest.skip("TMA needs 16-byte row strides; ops.gemm routes such shapes to the SIMT kernel")
    out = ops.gemm(A, B, transA=tA, transB=tB)
    ref = _ref(A.cpu(), B.cpu(), tA, tB)
    err = rel_err(out.cpu().numpy(), ref.numpy())
    assert err < GEMM_TOL, (err, backend, tA, tB, M, N, K)


@pytest.mark.parametrize("backend", ["simt", "tcgen05"])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_RELU, ops.ACT_SIGMOID_RELU, ops.ACT_ELU])
def test_gemm_bias_act(backend, act, monkeypatch):
    monkeypatch.setattr(ops, "GEMM_BACKEND", backend)
    monkeypatch.setattr(ops, "TC_MIN_WORK", 0)
    g = torch.Generator().manual_seed(act)
    A = torch.randn(777, 256, generator=g).to(DEV)
    W = (torch.randn(192, 256, generator=g) * 0.1).to(DEV)
    b = torch.randn(192, generator=g).to(DEV)
    out = ops.gemm(A, W, transB=True, bias=b, act=act)
    ref = _ref(A.cpu(), W.cpu(), False, True, b.cpu())
    if act == ops.ACT_RELU:
        ref = ref.clamp_min(0)
    elif act == ops.ACT_SIGMOID_RELU:
        ref = torch.sigmoid(ref.clamp_min(0))
    elif act == ops.ACT_ELU:
        ref = torch.nn.functional.elu(ref)
    assert rel_err(out.cpu().numpy(), ref.numpy()) < GEMM_TOL


def test_gemm_split_k_weight_gradient(monkeypatch):
    """dW = G^T Z with K = P large: both operands MN-major, split-K with atomic accumulation."""
    monkeypatch.setattr(ops, "GEMM_BACKEND", "tcgen05")
    g = torch.Generator().manual_seed(5)
    P, Hd, C = 50_000, 256, 256
    G = torch.randn(P, Hd, generator=g).to(DEV)
    Z = torch.randn(P, C, generator=g).to(DEV)
    out = ops.gemm(G, Z, transA=True)
    ref = G.cpu().double().t() @ Z.cpu().double()
    assert rel_err(out.cpu().numpy(), ref.numpy()) < GEMM_TOL


def test_gemm_strided_views(monkeypatch):
    """Per-head column slices (lda != K) as used by the u @ v.T score (Ours.py:108)."""
    for backend in ("simt", "tcgen05"):
        monkeypatch.setattr(ops, "GEMM_BACKEND", backend)
        monkeypatch.setattr(ops, "TC_MIN_WORK", 0)
        g = torch.Generator().manual_seed(9)
        U = torch.randn(1000, 128, generator=g).to(DEV)
        V = torch.randn(32, 128, generator=g).to(DEV)
        out = torch.empty(1000, 64, device=DEV)
        for h in range(2):
            ops.gemm(U[:, h * 64:(h + 1) * 64], V[:, h * 64:(h + 1) * 64], transB=True, act=ops.ACT_ELU,
                     out=out[:, h * 32:(h + 1) * 32])
            ref = torch.nn.functional.elu(U.cpu().double()[:, h * 64:(h + 1) * 64] @ V.cpu().double()[:, h * 64:(h + 1) * 64].t())
            assert rel_err(out[:, h * 32:(h + 1) * 32].cpu().numpy(), ref.numpy()) < GEMM_TOL
