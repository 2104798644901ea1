"""CUDA-graph capture of a training step (msha_gnn_b200.graphs.CapturedStep) and the device-side dropout epoch that
keeps replayed dropout masks fresh.  The Philox stream is builder-defined (parity unpinned by the reference);
its oracle is oracle.dropout_keep_mask."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import msha_gnn_b200 as mg
from msha_gnn_b200 import ops
from msha_gnn_b200.ops import call, _stream
from conftest import load_golden, params_of, rel_err
from oracle import msha_oracle as O

DEV = "cuda:0"


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(autouse=True)
def _reset_epoch():
    ops.dropout_epoch_set(0)
    yield
    ops.dropout_epoch_set(0)
    torch.cuda.synchronize()


def test_dropout_epoch_shifts_the_philox_key():
    n, p, seed = 10007, 0.3, 77
    keep = torch.empty(n, dtype=torch.uint8, device=DEV)
    for epoch in (0, 1, 5, (1 << 40) + 3):
        ops.dropout_epoch_set(epoch)
        call("msha_dropout_mask", seed, 2, n, p, keep.data_ptr(), _stream())
        assert np.array_equal(_np(keep).astype(bool), O.dropout_keep_mask(seed, n, p, stream=2, epoch=epoch)), epoch
    ops.dropout_epoch_set(3)
    ops.dropout_epoch_advance(2)                                    # 3 + 2
    x = torch.rand(n, device=DEV)
    y = ops.dropout_apply(x, 0.5, seed)                             # feature dropout draws from stream 4
    want = _np(x) * O.dropout_keep_mask(seed, n, 0.5, stream=4, epoch=5) * np.float32(2.0)
    assert np.array_equal(_np(y), want.astype(np.float32))
    # the other translation units see the same epoch: attention dropout inside the fused GAT kernel
    rng = np.random.default_rng(0)
    adj = (rng.random((60, 60)) < 0.2).astype(np.float32)
    np.fill_diagonal(adj, 1.0)
    graph = mg.Graph.from_dense(torch.tensor(adj, device=DEV))
    torch.manual_seed(0)
    conv = mg.GATConv(8, 4, heads=2, dropout=0.5).to(DEV).train()
    xin = torch.rand(60, 8, device=DEV)
    outs = []
    for epoch in (0, 1, 0):
        ops.dropout_epoch_set(epoch)
        ops._seed_counter = __import__("itertools").count(1)        # same by-value seed for the three calls
        outs.append(_np(conv(xin, graph)))
    assert np.array_equal(outs[0], outs[2]) and not np.array_equal(outs[0], outs[1])


def test_captured_dropout_draws_fresh_masks_per_replay():
    n, seed = 4096, 123
    x = torch.rand(n, device=DEV)
    step = mg.CapturedStep(lambda t: ops.dropout_apply(t, 0.5, seed), [x], warmup=1)   # epoch after construction: 1
    y1 = _np(step(x)).copy()                                        # replay 1 runs at epoch 2
    y2 = _np(step(x)).copy()                                        # replay 2 at epoch 3
    assert not np.array_equal(y1, y2)
    for y, epoch in ((y1, 2), (y2, 3)):
        want = _np(x) * O.dropout_keep_mask(seed, n, 0.5, stream=4, epoch=epoch) * np.float32(2.0)
        assert np.array_equal(y, want.astype(np.float32)), epoch
    x2 = torch.rand(n, device=DEV)
    y3 = _np(step(x2))                                              # new input through the static buffer, epoch 4
    want = _np(x2) * O.dropout_keep_mask(seed, n, 0.5, stream=4, epoch=4) * np.float32(2.0)
    assert np.array_equal(y3, want.astype(np.float32))
    with pytest.raises(ValueError):
        step(torch.rand(n + 1, device=DEV))


@pytest.mark.parametrize("name,cls", [("ablation3", "ablation3"), ("ablation2", "ablation2"), ("ours", "Ours")])
def test_captured_training_step_matches_eager(name, cls):
    """train.py:217-232 as one graph launch: same losses and parameters as the eager loop (dropout 0)."""
    g = load_golden(name)
    p = params_of(g)
    N, M = g["adj"].shape
    Fin, d = p["Sfeatures"].shape[1], p["attention_0.W1"].shape[1]
    gdp = {str(i): 0.0 for i in range(N)}
    graph = mg.Graph.from_dense(torch.tensor(g["adj"], device=DEV))
    city = torch.tensor(g["city"], device=DEV)
    prov = torch.tensor(g["prov"], device=DEV)
    rng = np.random.default_rng(1)
    batches = [torch.tensor(np.stack([rng.integers(0, N, 16), rng.integers(0, M, 16)]), device=DEV) for _ in range(4)]

    def make():
        model = getattr(mg, cls)(Fin, d, M, 2, 0.0, gdp, N, M)
        model.load_state_dict({k: torch.tensor(v) for k, v in p.items()})
        model = model.to(DEV).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4, capturable=True)     # train.py:207

        def fn(batch):
            opt.zero_grad(set_to_none=True)
            out = model(graph, city, prov, batch[0])
            loss = torch.nn.functional.nll_loss(out[batch[0]], batch[1])                             # train.py:229
            loss.backward()
            opt.step()
            return loss
        return model, fn

    eager_model, eager_fn = make()
    warmup = 2
    for _ in range(warmup):
        eager_fn(batches[0])
    eager_losses = [float(eager_fn(b)) for b in batches[1:]]
    cap_model, cap_fn = make()
    step = mg.CapturedStep(cap_fn, [batches[0]], warmup=warmup)
    cap_losses = [float(step(b)) for b in batches[1:]]
    np.testing.assert_allclose(cap_losses, eager_losses, rtol=2e-5)
    for (n1, a), (_, b) in zip(eager_model.named_parameters(), cap_model.named_parameters()):
        assert rel_err(_np(b), _np(a)) < 1e-4, n1
    for k in ("attention_0.bn1.running_mean", "attention_0.bn2.running_var", "attention_0.bn1.num_batches_tracked"):
        np.testing.assert_allclose(_np(cap_model.state_dict()[k]), _np(eager_model.state_dict()[k]), rtol=1e-5)
