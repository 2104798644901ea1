"""Thin tensor-level wrappers over the C-ABI: validate tensors, pass raw device pointers, sizes and the
current CUDA stream.  PyTorch is used only for device memory (caching allocator) and streams."""
from __future__ import annotations

import itertools
import os

import torch

from . import _lib

ACT_NONE, ACT_ELU, ACT_RELU, ACT_SIGMOID_RELU, ACT_LRELU, ACT_SIGMOID = 0, 1, 2, 3, 4, 5
LRELU_SLOPE = 0.2   # Ours.py:33 / GAT.py:27

GEMM_BACKEND = os.environ.get("MSHA_GEMM", "tcgen05")   # "tcgen05" (3xTF32 tensor cores) | "simt" (fp32 FMA validation path)
TC_MIN_WORK = 1 << 18                                    # below this many MACs the SIMT kernel's latency wins

launch_count = 0    # number of C-ABI compute calls issued (bench.py reports kernel launches from it)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (raw handle: this sits on every C-ABI call)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t, dtype=torch.float32, name="tensor"):
    """Device pointer of a contiguous CUDA tensor of the expected dtype (None -> NULL)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: msha_b200 ops are CUDA-only (got a {t.device} tensor); there is no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: tensor must be contiguous")
    p = t.data_ptr()
    align = 16 if dtype == torch.float32 else t.element_size()     # float rows are read with 128-bit loads
    if t.numel() and p % align:
        raise RuntimeError(f"{name}: device pointer must be {align}-byte aligned")
    return p


_fn_cache = {}


def call(fname, *args):
    global launch_count
    launch_count += 1
    fn = _fn_cache.get(fname)
    if fn is None:
        fn = _fn_cache[fname] = getattr(_lib.lib(), fname)
    rc = fn(*args)
    if rc:
        _lib.check(rc, fname)


def workspace(nbytes: int, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


_seed_counter = itertools.count(1)


def next_seed() -> int:
    """Per-call Philox key: (torch initial seed, call counter) -- no device synchronisation."""
    base = torch.initial_seed() & 0xFFFFFFFF
    return ((base << 32) | (next(_seed_counter) & 0xFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------ dense
def gemm(A, B, transA=False, transB=False, bias=None, act=ACT_NONE, out=None, beta=0.0, slope=LRELU_SLOPE):
    """act(op(A) @ op(B) + bias).  A, B may be row-strided 2-D views (last dim contiguous)."""
    assert A.dim() == 2 and B.dim() == 2
    M, K = (A.shape[1], A.shape[0]) if transA else (A.shape[0], A.shape[1])
    K2, N = (B.shape[1], B.shape[0]) if transB else (B.shape[0], B.shape[1])
    if K != K2:
        raise RuntimeError(f"gemm: inner dimensions differ ({K} vs {K2})")
    for t, n in ((A, "A"), (B, "B")):
        if not t.is_cuda or t.dtype != torch.float32 or (t.numel() and t.stride(1) != 1):
            raise RuntimeError(f"gemm: {n} must be a CUDA float32 matrix with contiguous rows")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    lda = A.stride(0) if A.shape[0] > 1 else A.shape[1]
    ldb = B.stride(0) if B.shape[0] > 1 else B.shape[1]
    ldc = out.stride(0) if M > 1 else N
    lib = _lib.lib()
    use_tc = (GEMM_BACKEND == "tcgen05" and beta == 0.0 and M * N * K >= TC_MIN_WORK
              and lib.msha_gemm_tf32x3_supported(A.data_ptr(), B.data_ptr(), M, N, K, lda, ldb) == 0)
    if use_tc:
        tiles = -(-M // 128) * -(-N // (256 if N > 128 else (128 if N > 64 else 64)))
        splits = 1
        if bias is None and act == ACT_NONE and K >= 4096 and tiles < 74:
            splits = max(1, min(148 // tiles, K // 1024))
        if splits > 1:
            out.zero_()
        call("msha_gemm_tf32x3", A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldc, int(transA),
             int(transB), ptr(bias, name="bias"), int(act), float(slope), splits, _stream())
    else:
        call("msha_gemm_f32", A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, lda, ldb, ldc, int(transA), int(transB),
             ptr(bias, name="bias"), float(beta), int(act), float(slope), _stream())
    return out


def act_fwd(x, act, slope=LRELU_SLOPE):
    y = torch.empty_like(x)
    call("msha_act_fwd", ptr(x), ptr(y), x.numel(), act, slope, _stream())
    return y


def act_bwd(dy, y, act, slope=LRELU_SLOPE):
    dx = torch.empty_like(y)
    call("msha_act_bwd", ptr(dy), ptr(y), ptr(dx), y.numel(), act, slope, _stream())
    return dx


def dropout_apply(x, p, seed, stream_id=4):
    y = torch.empty_like(x)
    call("msha_dropout_apply", ptr(x), ptr(y), x.numel(), float(p), seed, stream_id, _stream())
    return y


# ------------------------------------------------------------------------------------------ dropout epoch
def dropout_epoch_set(epoch: int):
    """Set the device-side dropout epoch (stream-ordered; 0 = the plain seed stream).  See ``msha_dropout_epoch_set``."""
    call("msha_dropout_epoch_set", int(epoch) & 0xFFFFFFFFFFFFFFFF, _stream())


def dropout_epoch_advance(by: int = 1):
    """Bump the dropout epoch on the current stream -- capturable: first node of a captured training step."""
    call("msha_dropout_epoch_advance", int(by) & 0xFFFFFFFFFFFFFFFF, _stream())
