"""Host-side logic of the multi-GPU path under gloo (world_size 2, CPU): partition bounds, padded id remap,
all-gather / reduce-scatter autograd and gradient all-reduce; and the partition algebra of a GAT layer checked with
the oracle's formulas (no kernels involved)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msha_gnn_b200.dist import Partition, all_gather_rows, allreduce_gradients
from oracle import msha_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_bounds_and_padding():
    p = Partition(10, 3, 1)
    assert p.bounds == [0, 3, 6, 10] and p.n_max == 4 and (p.lo, p.hi, p.n_local) == (3, 6, 3)
    ids = torch.arange(10)
    assert p.owner_of(ids).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 2]
    assert p.to_padded(ids).tolist() == [0, 1, 2, 4, 5, 6, 8, 9, 10, 11]
    rowptr = torch.tensor([0, 10, 10, 11, 12, 20, 40])           # skewed degrees
    q = Partition.edge_balanced(rowptr, 2, 0)
    assert q.bounds[0] == 0 and q.bounds[-1] == 6 and 0 < q.bounds[1] < 6
    e0 = int(rowptr[q.bounds[1]])
    assert abs(e0 - 20) <= 10                                    # close to half of the edges


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        N, Fin, H, d = 11, 5, 2, 3
        C = H * d
        part = Partition(N, world, rank)
        rng = np.random.default_rng(0)
        adj = (rng.random((N, N)) < 0.4).astype(np.float32)
        adj[np.arange(N), np.arange(N)] = 1
        X = torch.tensor(rng.random((N, Fin)), dtype=torch.float64)
        W = torch.tensor(rng.standard_normal((Fin, C)), dtype=torch.float64, requires_grad=True)
        a_n = torch.tensor(rng.standard_normal((H, d)), dtype=torch.float64, requires_grad=True)
        a_s = torch.tensor(rng.standard_normal((H, d)), dtype=torch.float64, requires_grad=True)
        G = torch.tensor(rng.standard_normal((N, C)), dtype=torch.float64)
        # ---- single-process oracle on the whole graph
        rowptr, col, _ = O.csr_from_dense(adj)
        Wf, anf, asf = (t.detach().clone().requires_grad_(True) for t in (W, a_n, a_s))
        full = O.gat_layer(X, Wf, anf, asf, rowptr, col, H, apply_elu=True)
        (full * G).sum().backward()
        # ---- partitioned: local rows, gathered columns (padded indexing)
        x_loc = X[part.lo:part.hi]
        Wh = (x_loc @ W)
        s_n = (Wh.view(-1, H, d) * a_n).sum(-1)
        s_s = (Wh.view(-1, H, d) * a_s).sum(-1)
        gathered = all_gather_rows(torch.cat([Wh, s_n], dim=1), part)
        assert gathered.shape == (part.n_padded, C + H)
        Wh_g, s_n_g = gathered[:, :C].reshape(-1, H, d), gathered[:, C:]
        r_loc, c_glob = np.nonzero(adj[part.lo:part.hi] > 0)
        r = torch.from_numpy(r_loc)
        c = part.to_padded(torch.from_numpy(c_glob))
        e = torch.nn.functional.leaky_relu(s_n_g[c] + s_s[r], 0.2)
        alpha = O.segment_softmax(e, r, part.n_local)
        out = torch.zeros(part.n_local, H, d, dtype=torch.float64).index_add(0, r, alpha[:, :, None] * Wh_g[c])
        out = torch.nn.functional.elu(out.reshape(part.n_local, C))
        assert torch.allclose(out, full[part.lo:part.hi].detach(), atol=1e-12)
        (out * G[part.lo:part.hi]).sum().backward()
        allreduce_gradients([W, a_n, a_s])
        for got, ref in ((W, Wf), (a_n, anf), (a_s, asf)):
            assert torch.allclose(got.grad, ref.grad, atol=1e-10), (got.grad - ref.grad).abs().max()
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_partitioned_gat_layer_equals_full_graph_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
