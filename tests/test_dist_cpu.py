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
    # block stride of the gathered layout: a multiple of 4 rows (128-bit peer pulls / sums)
    assert Partition(10, 2, 0).n_max == 8 and Partition(10, 2, 0).n_padded == 16 and Partition(10, 2, 1).sizes == [5, 5]
    assert Partition(10, 2, 1).to_padded(torch.arange(10)).tolist() == [0, 1, 2, 3, 4, 8, 9, 10, 11, 12]
    # a hub row heavier than total / world must not produce empty ranks
    hub = torch.tensor([0, 1000, 1001, 1002, 1003, 1004])
    for w in (2, 3, 5):
        b = Partition.edge_balanced(hub, w, 0).bounds
        assert all(b[i + 1] > b[i] for i in range(w)), b
    try:
        Partition.edge_balanced(hub, 6, 0)
        raise AssertionError("6 ranks over 5 rows must be rejected")
    except ValueError:
        pass


class _StubPredictor:
    """nll_loss_pairs with plain torch ops (the scaling logic under test is host-side)."""

    def __init__(self, W):
        self.W = W

    def nll_loss_pairs(self, hi, hj, src, dst, target):
        out = torch.sigmoid(torch.relu((hi[src] * hj[dst]) @ self.W.t()))
        return torch.nn.functional.nll_loss(out, target)


def _loss_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from msha_gnn_b200.dist import score_pairs
        rng = np.random.default_rng(3)
        N, C, Hd = 13, 6, 4
        h = torch.tensor(rng.standard_normal((N, C)), dtype=torch.float64)
        W0 = torch.tensor(rng.standard_normal((Hd, C)), dtype=torch.float64)
        P = 37                                                   # uneven shards: 12 and 25 pairs
        src, dst = torch.from_numpy(rng.integers(0, N, P)), torch.from_numpy(rng.integers(0, N, P))
        lab = torch.from_numpy(rng.integers(0, Hd, P))
        cut = [0, 12, P]
        # single-process reference: mean over all pairs
        hf, Wf = h.clone().requires_grad_(True), W0.clone().requires_grad_(True)
        ref = _StubPredictor(Wf).nll_loss_pairs(hf, hf, src, dst, lab)
        ref.backward()
        part = Partition(N, world, rank)
        hl = h[part.lo:part.hi].clone().requires_grad_(True)
        Wl = W0.clone().requires_grad_(True)
        sl = slice(cut[rank], cut[rank + 1])
        for gp in (None, P):                                     # pair count all-reduced inside / given by the caller
            hl.grad = Wl.grad = None
            loss = score_pairs(_StubPredictor(Wl), hl, src[sl], dst[sl], part, target=lab[sl], global_pairs=gp)
            loss.backward()
            allreduce_gradients([Wl])
            tot = loss.detach().clone()
            dist.all_reduce(tot)
            assert torch.allclose(tot, ref.detach(), atol=1e-12), (tot, ref)          # shares add up to the global mean
            assert torch.allclose(Wl.grad, Wf.grad, atol=1e-12)
            assert torch.allclose(hl.grad, hf.grad[part.lo:part.hi], atol=1e-12)
        # a rank without a gradient for one parameter still takes part in the flat all-reduce
        a, b = torch.ones(3, requires_grad=True), torch.ones(2, requires_grad=True)
        (a.sum() * (rank + 1)).backward() if rank == 0 else (a.sum() + b.sum()).backward()
        allreduce_gradients([a, b])
        assert a.grad.tolist() == [2.0] * 3 and b.grad.tolist() == [1.0] * 2
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_mean_nll_over_ranks_equals_single_process_gloo():
    """the per-rank mean losses used to add up to ~world x the single-GPU loss."""
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_loss_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


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


def _msha_worker(rank, world, port, ret):
    """Partition algebra of the MSHA layer (OursLayer3, Ablation.py:260-277) through dist_msha's torch transports: gather of
    h1, reduce-scatter of alpha.T @ h2 partials, BatchNorm statistics all-reduced -- against the oracle on the whole graph."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from msha_gnn_b200 import dist_msha as dm
        F = torch.nn.functional
        rng = np.random.default_rng(11)
        N, M, Fin, d = 13, 7, 5, 4
        adj = (rng.random((N, M)) < 0.4).astype(np.float32)
        adj[np.arange(N), rng.integers(0, M, N)] = 1
        S = torch.tensor(rng.random((N, Fin)))
        R = torch.tensor(rng.random((M, Fin)))
        p = {"W1": torch.tensor(rng.standard_normal((Fin, d))), "W2": torch.tensor(rng.standard_normal((Fin, d))),
             "a": torch.tensor(rng.standard_normal((2 * d, 1)))}
        for bn in ("bn1", "bn2"):
            p[bn + ".weight"], p[bn + ".bias"] = torch.tensor(rng.random(d) + 0.5), torch.tensor(rng.standard_normal(d))
            p[bn + ".running_mean"], p[bn + ".running_var"] = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
        G = torch.tensor(rng.standard_normal((N, M)))
        rowptr, col, _ = O.csr_from_dense(adj)
        pf = {k: (v.clone().requires_grad_(True) if k in ("W1", "W2", "a") else v.clone()) for k, v in p.items()}
        full = O.ours_layer3(S, R, pf, rowptr, col, training=True)
        (full * G).sum().backward()
        # ---- partitioned
        ps, pr = Partition(N, world, rank), Partition(M, world, rank)
        comm = dm.TorchComm()
        W1, W2, a = (p[k].clone().requires_grad_(True) for k in ("W1", "W2", "a"))
        h1 = R[pr.lo:pr.hi] @ W1
        h2 = S[ps.lo:ps.hi] @ W2
        h1_g = comm.gather_rows(h1, pr, "h1")
        s_nbr_g = h1_g @ a[:d, 0]
        s_self = h2 @ a[d:, 0]
        r_loc, c_glob = np.nonzero(adj[ps.lo:ps.hi] > 0)
        r, c = torch.from_numpy(r_loc), pr.to_padded(torch.from_numpy(c_glob))
        alpha = O.segment_softmax(F.leaky_relu(s_nbr_g[c] + s_self[r], 0.2), r, ps.n_local)
        u_in = torch.zeros(ps.n_local, d, dtype=torch.float64).index_add(0, r, alpha[:, None] * h1_g[c])
        v_part = torch.zeros(pr.n_padded, d, dtype=torch.float64).index_add(0, c, alpha[:, None] * h2[r])
        v_in = comm.reduce_scatter_rows(v_part, pr, "v")
        assert v_in.shape == (pr.n_local, d)

        def bn(x, w, b, n_total):                              # batch statistics over the GLOBAL node axis
            sums = dm.all_reduce_sum(torch.cat([x.sum(0), (x * x).sum(0)]), comm)
            mean = sums[:d] / n_total
            var = sums[d:] / n_total - mean * mean
            return F.leaky_relu((x - mean) / torch.sqrt(var + 1e-5) * w + b, 0.2)
        v = bn(v_in, p["bn1.weight"], p["bn1.bias"], M)
        u = bn(u_in, p["bn2.weight"], p["bn2.bias"], N)
        v_g = comm.gather_rows(v, pr, "vg")
        cols_p = pr.to_padded(torch.arange(M))
        out = F.elu(u @ v_g[cols_p].t())
        assert torch.allclose(out, full[ps.lo:ps.hi].detach(), atol=1e-10)
        (out * G[ps.lo:ps.hi]).sum().backward()
        allreduce_gradients([W1, W2, a])
        for got, name in ((W1, "W1"), (W2, "W2"), (a, "a")):
            assert torch.allclose(got.grad, pf[name].grad, atol=1e-9), (name, (got.grad - pf[name].grad).abs().max())
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_partitioned_msha_layer_algebra_gloo():
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_msha_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_halo_numbering_host_logic():
    """Host-side index arithmetic of the halo exchange (dist_p2p.halo_need / halo_offsets / halo_compact_ids) on simulated
    ranks, CPU tensors: the compact numbering reproduces the gathered one (forward gather), and the owners' "give" lists --
    read out of the peers' "need" lists exactly as HaloPlan reads them out of peer memory -- make the scatter-add the adjoint
    of the gather (the gradient's way back equals a reduce-scatter of the dense gathered gradient)."""
    from msha_gnn_b200.dist_p2p import halo_need, halo_offsets, halo_compact_ids
    rng = np.random.default_rng(5)
    N, W, C = 103, 4, 3
    parts = [Partition(N, W, r) for r in range(W)]
    n_max = parts[0].n_max
    feat = torch.tensor(rng.standard_normal((N, C)))
    gathered = torch.zeros(parts[0].n_padded, C, dtype=torch.float64)                 # the whole-block gathered layout
    gathered[parts[0].to_padded(torch.arange(N))] = feat
    cols = [torch.from_numpy(rng.choice(N, size=rng.integers(5, 60), replace=True)) for _ in range(W)]   # each rank's edge columns
    plans = []
    for r in range(W):
        colp = parts[r].to_padded(cols[r])
        uniq, inv, owner, remote, cnt, need = halo_need(colp, W, r, n_max)
        need_ptr = [0]
        for q in range(W):
            need_ptr.append(need_ptr[-1] + (cnt[q] if q != r else 0))
        plans.append(dict(colp=colp, uniq=uniq, inv=inv, owner=owner, remote=remote, cnt=cnt, need=need, need_ptr=need_ptr))
    all_cnt = [p["cnt"] for p in plans]
    n_compact = max(halo_offsets(all_cnt[q], q, n_max)[1] for q in range(W))
    grads, comps = [], []
    for r, p in enumerate(plans):
        hoff, _ = halo_offsets(all_cnt[r], r, n_max)
        comp = halo_compact_ids(p["uniq"], p["owner"], p["remote"], r, n_max, hoff, p["need_ptr"])[p["inv"]]
        comps.append(comp)
        assert int(comp.max()) < n_compact and all(h % 4 == 0 for h in hoff)
        # forward: own block + the listed rows of every peer's own block, packed at hoff[q] (msha_peer_gather_rows)
        buf = torch.full((n_compact, C), float("nan"), dtype=torch.float64)
        buf[: parts[r].n_local] = feat[parts[r].lo:parts[r].hi]
        for q in range(W):
            if q != r:
                rows = p["need"][p["need_ptr"][q]:p["need_ptr"][q + 1]].long()
                assert torch.all(rows[1:] > rows[:-1])                                   # ascending, distinct
                buf[hoff[q]:hoff[q] + rows.numel()] = feat[parts[q].lo + rows]
        assert torch.equal(buf[comp], gathered[p["colp"]])                               # same rows as the whole-block gather
        g = torch.zeros(n_compact, C, dtype=torch.float64)                               # a gradient in the compact numbering
        g.index_add_(0, comp, torch.tensor(rng.standard_normal((comp.numel(), C))))
        grads.append(g)
    # backward: owner r adds, for every peer q, q's halo segment for r onto the rows q listed (msha_peer_scatter_add_rows)
    for r in range(W):
        own = grads[r][: parts[r].n_local].clone()
        for q in range(W):
            if q == r:
                continue
            n_q = all_cnt[q][r]
            start = sum(all_cnt[q][p_] for p_ in range(r) if p_ != q)                    # HaloPlan's read of peer q's list buffer
            give = plans[q]["need"][start:start + n_q].long()
            off = halo_offsets(all_cnt[q], q, n_max)[0][r]
            own.index_add_(0, give, grads[q][off:off + n_q])
        dense = torch.zeros(parts[0].n_padded, C, dtype=torch.float64)                    # reduce-scatter of the dense gathered gradients
        for q in range(W):
            dq = torch.zeros_like(dense)
            # scatter q's compact gradient back to the gathered numbering through its (compact id -> padded id) map
            pad_of = torch.zeros(n_compact, dtype=torch.int64)
            pad_of[comps[q]] = plans[q]["colp"]
            used = torch.unique(comps[q])
            dq.index_add_(0, pad_of[used], grads[q][used])
            dense += dq
        want = dense[r * n_max: r * n_max + parts[r].n_local]
        assert torch.allclose(own, want, atol=1e-12)


def test_backward_passes_of_the_peer_path_never_synchronise_the_host():
    """Source check.  A backward node that blocks its host thread (device-to-host read, pageable upload, lazily built CSC /
    hub table) does so behind kernels that wait for a peer's flag; ranks emulated in one process share autograd's device
    thread, so the peer's signalling node is then never enqueued (profiles/r02_gpu_tests_last_run.txt).  Every
    ``backward`` of dist_p2p / dist_msha therefore works on tables that exist already: HaloPlan / gat_encode_p2p build them."""
    import ast
    import inspect
    from msha_gnn_b200 import dist_p2p, dist_msha
    blocking = {"item", "tolist", "cpu", "synchronize", "nonzero", "unique", "numpy"}
    seen = 0
    for mod in (dist_p2p, dist_msha):
        tree = ast.parse(inspect.getsource(mod))
        for cls in [n for n in ast.walk(tree) if isinstance(n, ast.ClassDef)]:
            for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "backward"]:
                seen += 1
                for node in ast.walk(fn):
                    if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute):
                        assert node.func.attr not in blocking, (mod.__name__, cls.name, node.func.attr, node.lineno)
                        assert not (node.func.attr in ("tensor", "as_tensor") and getattr(node.func.value, "id", "") == "torch"), \
                            (mod.__name__, cls.name, "host list -> device upload", node.lineno)
    assert seen >= 6
    # ... and the halo plan builds the compact graph's derived tables itself
    src = inspect.getsource(dist_p2p.HaloPlan.__init__)
    for built in ("attention_csc()", "hub_rows()", "hub_cols()"):
        assert "self.graph." + built in src
