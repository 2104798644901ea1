"""Probe of the peer-memory data path on real ranks (run under torchrun, >= 2 GPUs):
symmetric allocation, flag round trip, copy-engine and SM pull bandwidth (alone and under a memory-bound kernel),
NCCL all-gather / reduce-scatter of the same size for comparison.  Prints one JSON line per measurement (rank 0)."""
import json
import os
import sys
import time

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def ev_time(fn, reps=5, warm=2, stream=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from msha_gnn_b200 import peer
    out = {}

    def say(k, v):
        out[k] = v
        if rank == 0:
            print(json.dumps({k: v}), flush=True)

    t0 = time.time()
    try:
        fab = peer.SymmFabric()
    except Exception as e:          # noqa: BLE001
        say("symm_fabric_error", repr(e)[:500])
        dist.destroy_process_group()
        return
    pg = fab.group
    say("symm_setup_s", round(time.time() - t0, 2))

    # ---- flags: barrier latency, ping-pong
    for _ in range(3):
        pg.barrier()
    torch.cuda.synchronize()
    ms = ev_time(lambda: pg.barrier(), reps=50)
    say("barrier_us", round(ms * 1e3, 2))

    # ---- data: block of n_max x 256 fp32
    n_max, C = (int(os.environ.get("PROBE_ROWS", "250000")), 256)
    buf = pg.alloc((world * n_max, C))
    buf.local.normal_()
    mine = slice(rank * n_max, (rank + 1) * n_max)
    buf.local[mine] = float(rank + 1)
    torch.cuda.synchronize()
    dist.barrier()
    blk_bytes = n_max * C * 4
    q = (rank + 1) % world
    qs = slice(q * n_max, (q + 1) * n_max)

    ms = ev_time(lambda: pg.pull_block(buf, q, qs))
    say("ce_pull_1peer_GBps", round(blk_bytes / ms / 1e6, 1))
    ok = bool((buf.local[qs] == float(q + 1)).all().item())
    say("ce_pull_correct", ok)

    s2 = torch.cuda.Stream()
    half = n_max // 2

    def two_streams():
        cur = torch.cuda.current_stream()
        s2.wait_stream(cur)
        pg.pull_block(buf, q, slice(q * n_max, q * n_max + half))
        with torch.cuda.stream(s2):
            pg.pull_block(buf, q, slice(q * n_max + half, (q + 1) * n_max))
        cur.wait_stream(s2)
    ms = ev_time(two_streams)
    say("ce_pull_1peer_2streams_GBps", round(blk_bytes / ms / 1e6, 1))

    if world > 2:
        streams = [torch.cuda.Stream() for _ in range(world - 1)]

        def all_peers(nstreams):
            cur = torch.cuda.current_stream()
            for s in range(1, world):
                st = streams[(s - 1) % nstreams]
                st.wait_stream(cur)
                p = (rank + s) % world
                with torch.cuda.stream(st):
                    pg.pull_block(buf, p, slice(p * n_max, (p + 1) * n_max))
            for st in streams[:nstreams]:
                cur.wait_stream(st)
        for ns in (1, 2, 4, world - 1):
            ms = ev_time(lambda: all_peers(ns))
            say(f"ce_allgather_{ns}streams_GBps_in", round(blk_bytes * (world - 1) / ms / 1e6, 1))

    for ctas in (16, 32, 64, 148, 296, 592):
        ms = ev_time(lambda: pg.pull_blocks_sm(buf, n_max, max_ctas=ctas))
        say(f"sm_pull_all_{ctas}ctas_GBps_in", round(blk_bytes * (world - 1) / ms / 1e6, 1))
    buf.local[qs] = 0
    pg.pull_blocks_sm(buf, n_max)
    torch.cuda.synchronize()
    say("sm_pull_correct", bool((buf.local[qs] == float(q + 1)).all().item()))

    # ---- small blocks (latency regime): 4267 rows
    small = pg.alloc((world * 4267, C))
    ms = ev_time(lambda: pg.pull_blocks_sm(small, 4267), reps=20)
    say("sm_pull_small_us", round(ms * 1e3, 1))
    ms = ev_time(lambda: [pg.pull_block(small, (rank + s) % world, slice(((rank + s) % world) * 4267, ((rank + s) % world + 1) * 4267))
                          for s in range(1, world)], reps=20)
    say("ce_pull_small_us", round(ms * 1e3, 1))

    # ---- overlap: a memory-bound kernel on the main stream, the pull on a side stream
    big_a = torch.empty(1 << 28, device=dev)         # 1 GiB
    big_b = torch.empty(1 << 28, device=dev)
    ms_k = ev_time(lambda: big_b.copy_(big_a))
    say("hbm_copy_alone_ms", round(ms_k, 3))
    side = torch.cuda.Stream()

    def overlapped(kind):
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for s in range(1, world):
                p = (rank + s) % world
                if kind == "ce":
                    pg.pull_block(buf, p, slice(p * n_max, (p + 1) * n_max))
            if kind.startswith("sm"):
                pg.pull_blocks_sm(buf, n_max, max_ctas=int(kind[2:]))
        big_b.copy_(big_a)
        big_b.copy_(big_a)
        cur.wait_stream(side)
    for kind in ("ce", "sm32", "sm148"):
        ms = ev_time(lambda: overlapped(kind))
        say(f"overlap_{kind}_total_ms (2 copies alone = {2 * ms_k:.3f})", round(ms, 3))

    # ---- sum kernel over peer memory (small-message reduce-scatter) and over local staging
    addrs = [buf.addr[p] + rank * blk_bytes for p in range(world)]
    res = torch.empty(n_max, C, device=dev)
    ms = ev_time(lambda: pg.sum_into(res, addrs, n_max * C))
    say("sum_over_peers_GBps_in", round(blk_bytes * (world - 1) / ms / 1e6, 1))
    addrs_l = [buf.addr[rank] + p * blk_bytes for p in range(world)]
    ms = ev_time(lambda: pg.sum_into(res, addrs_l, n_max * C))
    say("sum_local_GBps_read", round(blk_bytes * world / ms / 1e6, 1))

    # ---- NCCL for comparison
    x = torch.empty(n_max, C, device=dev)
    g = torch.empty(world * n_max, C, device=dev)
    ms = ev_time(lambda: dist.all_gather_into_tensor(g, x))
    say("nccl_all_gather_GBps_in", round(blk_bytes * (world - 1) / ms / 1e6, 1))
    ms = ev_time(lambda: dist.reduce_scatter_tensor(x, g))
    say("nccl_reduce_scatter_GBps_in", round(blk_bytes * (world - 1) / ms / 1e6, 1))
    xs = torch.empty(4267, C, device=dev)
    gs = torch.empty(world * 4267, C, device=dev)
    ms = ev_time(lambda: dist.all_gather_into_tensor(gs, xs), reps=20)
    say("nccl_all_gather_small_us", round(ms * 1e3, 1))
    ms = ev_time(lambda: dist.reduce_scatter_tensor(xs, gs), reps=20)
    say("nccl_reduce_scatter_small_us", round(ms * 1e3, 1))
    one = torch.zeros(1, device=dev)
    ms = ev_time(lambda: dist.all_reduce(one), reps=20)
    say("nccl_all_reduce_tiny_us", round(ms * 1e3, 1))
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
