"""Partitioned (multi-GPU) path on ONE GPU: the ranks are emulated sequentially in one process -- each rank's local
CSR (padded column indexing) runs through the same kernels with a hand-assembled "gathered" buffer -- and must
reproduce the whole-graph result.  (Kernels that wait on one another are never launched; B200_PROFILING.md.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import msha_gnn_b200 as mg
from msha_gnn_b200 import functional as Fn
from msha_gnn_b200.dist import Partition, partition_graph, gat_encode, all_gather_rows
from msha_gnn_b200.ops import ACT_ELU
from conftest import rel_err

DEV = "cuda:0"


@pytest.mark.parametrize("overlap", [False, True])
def test_world1_partition_equals_plain_conv(overlap):
    """overlap=True runs the split backward (d Wh first, handed to the GradSink, then row pass and d s_nbr sums)."""
    rng = np.random.default_rng(0)
    N, Fin, H, d = 200, 32, 4, 16
    adj = (rng.random((N, N)) < 0.1).astype(np.float32)
    adj[np.arange(N), np.arange(N)] = 1
    r, c = np.nonzero(adj)
    torch.manual_seed(1)
    convs = torch.nn.ModuleList([mg.GATConv(Fin, d, H), mg.GATConv(H * d, d, H)]).to(DEV)
    x = torch.rand(N, Fin, device=DEV, requires_grad=True)
    g_full = mg.Graph.from_dense(torch.from_numpy(adj).to(DEV))
    ref = convs[1](convs[0](x, g_full), g_full)
    part = Partition(N, 1, 0)
    pg = partition_graph(torch.from_numpy(r).to(DEV), torch.from_numpy(c).to(DEV), part)
    x2 = x.detach().clone().requires_grad_(True)
    out = gat_encode(convs, x2, pg, part, overlap=overlap)
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    ref.square().sum().backward()
    gref = x.grad.clone()
    out.square().sum().backward()
    assert rel_err(x2.grad.cpu().numpy(), gref.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_ranks_match_full_graph(world):
    rng = np.random.default_rng(world)
    N, Fin, H, d = 301, 24, 2, 16
    C = H * d
    adj = (rng.random((N, N)) < 0.06).astype(np.float32)
    adj[np.arange(N), np.arange(N)] = 1
    torch.manual_seed(2)
    conv = mg.GATConv(Fin, d, H).to(DEV)
    x = torch.rand(N, Fin, device=DEV)
    G = torch.randn(N, C, device=DEV)
    # ---- reference: whole graph on one GPU
    xf = x.clone().requires_grad_(True)
    ref = conv(xf, mg.Graph.from_dense(torch.from_numpy(adj).to(DEV)))
    (ref * G).sum().backward()
    gx_ref, gW_ref = xf.grad.clone(), conv.W.grad.clone()
    conv.W.grad = None
    conv.a_nbr.grad = None
    conv.a_self.grad = None
    # ---- emulation: per-rank local tensors, gathered buffer assembled by hand (what all_gather_into_tensor returns)
    parts = [Partition(N, world, r) for r in range(world)]
    xs = [x[p.lo:p.hi].clone().requires_grad_(True) for p in parts]
    Whs, sn, ss = [], [], []
    for p, xl in zip(parts, xs):
        Wh = Fn.linear(xl, conv.W)
        a, b = Fn.node_scores(Wh, conv.a_nbr, conv.a_self, H, d)
        Whs.append(Wh); sn.append(a); ss.append(b)
    n_max = parts[0].n_max

    def pad(t):
        return torch.cat([t, t.new_zeros(n_max - t.shape[0], t.shape[1])]) if t.shape[0] < n_max else t
    Wh_g = torch.cat([pad(t) for t in Whs])              # autograd cat == all-gather fwd / reduce-scatter bwd
    sn_g = torch.cat([pad(t) for t in sn])
    outs = []
    for p, s_self in zip(parts, ss):
        r, c = np.nonzero(adj[p.lo:p.hi])
        pg = partition_graph(torch.from_numpy(r + p.lo).to(DEV), torch.from_numpy(c).to(DEV), p)
        assert pg.n_rows == p.n_local and pg.n_cols == p.n_padded
        o, _ = Fn.attention_block(pg, sn_g, s_self, Wh_g, heads=H, act=ACT_ELU)
        outs.append(o)
    out = torch.cat(outs)
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    (out * G).sum().backward()
    gx = torch.cat([t.grad for t in xs])
    assert rel_err(gx.cpu().numpy(), gx_ref.cpu().numpy()) < 1e-5
    assert rel_err(conv.W.grad.cpu().numpy(), gW_ref.cpu().numpy()) < 1e-5
