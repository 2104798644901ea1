"""Does a flag-wait kernel spinning on a side stream delay kernels of the main stream?  (2 ranks, torchrun.)
Rank 0 parks a wait on a side stream for a flag that rank 1 sets ~40 ms later, and meanwhile times small kernels on its main
stream -- with torch pool streams of both priorities and with the copy-engine pull in between."""
import json, os, sys, time
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from msha_gnn_b200 import peer
fab = peer.SymmFabric(); pg = fab.group
x = torch.zeros(1 << 20, device=dev)
seq = 0
for label, side in (("pool stream prio 0", torch.cuda.Stream()), ("pool stream prio -1", torch.cuda.Stream(priority=-1))):
    for trial in range(2):
        seq += 1
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            with torch.cuda.stream(side):
                pg.wait(7, seq, 1 << 1)                       # spins until rank 1 signals
                ev_side = torch.cuda.Event(enable_timing=True); ev_side.record(side)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                pg.signal(9, seq)                             # small kernels on the main stream
                x.add_(1.0)
            e1.record()
            e1.synchronize()
            ms_main = e0.elapsed_time(e1)
            torch.cuda.synchronize()
            ms_side = e0.elapsed_time(ev_side)
            print(json.dumps({"case": label, "trial": trial, "main_40_small_kernels_ms": round(ms_main, 3),
                              "side_wait_released_after_ms": round(ms_side, 2)}), flush=True)
        else:
            torch.cuda._sleep(int(40e-3 * 1.9e9))             # ~40 ms
            pg.signal(7, seq, 1 << 0)
            torch.cuda.synchronize()
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
