"""CUDA-graph capture of a whole training step.

The reference's own training configuration (train.py:206-232: ``ablation3`` / ``Ours`` on the 2015 flow graph, batch 64)
is launch-bound on a B200: ~40 C-ABI calls and as many small torch kernels for ~1 ms of device work.  ``CapturedStep``
records one step -- forward, loss, backward, optimiser update -- into a CUDA graph and replays it with a single launch.

Two things make a step capturable here:
  * every C-ABI call is stream-ordered on torch's current stream and allocates nothing, so the library calls are
    recorded like any other kernel launch (graph structures -- CSR/CSC, hub segments -- are built and cached by the
    warm-up steps that run before the capture, because their construction reads sizes back to the host);
  * dropout seeds are by-value kernel arguments and would be frozen into the graph; the first captured node bumps the
    library's device-side *dropout epoch* (``msha_dropout_epoch_advance``), which every Philox draw folds into its key,
    so each replay draws fresh keep-masks while the forward and backward of one replay still agree.

The optimiser must be capture-safe (``torch.optim.Adam(..., capturable=True)``).
"""
from __future__ import annotations

import torch

from . import ops


class CapturedStep:
    """``step = CapturedStep(fn, example_inputs)``; ``out = step(*inputs)`` copies ``inputs`` into the static input
    buffers, replays the graph and returns the static output tensor(s) of ``fn`` (valid until the next call).

    ``fn(*inputs)`` must be a complete step on CUDA tensors (``zero_grad(set_to_none=True)`` ... ``optimizer.step()``) and
    must not synchronise with the host.  ``warmup`` eager calls run first on a side stream."""

    def __init__(self, fn, example_inputs, warmup: int = 3, advance_dropout: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("msha_b200 is CUDA-only: CapturedStep needs a CUDA device")
        self.fn = fn
        self.static_inputs = [x.clone() for x in example_inputs]
        # warm-up and capture share one side stream: autograd pins each parameter's gradient accumulation to the stream
        # it first ran on, and a capture on a different stream would have to cross-synchronise with it
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(max(warmup, 1)):
                if advance_dropout:
                    ops.dropout_epoch_advance(1)
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            if advance_dropout:
                ops.dropout_epoch_advance(1)
            self.static_output = fn(*self.static_inputs)

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_inputs):
            raise TypeError(f"CapturedStep: expected {len(self.static_inputs)} inputs, got {len(inputs)}")
        for s, x in zip(self.static_inputs, inputs):
            if s.shape != x.shape or s.dtype != x.dtype:
                raise ValueError(f"CapturedStep: input {tuple(x.shape)} {x.dtype} does not match the captured "
                                 f"{tuple(s.shape)} {s.dtype} (shapes are frozen into the graph)")
            s.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_output
