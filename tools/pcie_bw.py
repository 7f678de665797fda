"""Concurrent pinned-memory transfer ceiling of the box: every rank copies an extraction call's worth of results device to
host (8.5 MB) and of frames host to device (2.9 MB) in a loop, both directions at once, and reports GB/s per rank and in
aggregate (torchrun; world 1 = the single-GPU ceiling).  Explains the end-to-end arm's scaling."""
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D2H, H2D, REPS = 8_515_328, 2_887_680, 400
h_out = [torch.empty(D2H, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_in = [torch.empty(H2D, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_out = torch.empty(D2H, dtype=torch.uint8, device="cuda")
d_in = torch.empty(H2D, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def loop(reps):
    for i in range(reps):
        with torch.cuda.stream(s1):
            h_out[i % 4].copy_(d_out, non_blocking=True)
        with torch.cuda.stream(s2):
            d_in.copy_(h_in[i % 4], non_blocking=True)
    torch.cuda.synchronize()


loop(20)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
loop(REPS)
dt = time.perf_counter() - t0
gbs = REPS * (D2H + H2D) / dt / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
lo = t.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: {gbs:.1f} GB/s on rank 0 (d2h + h2d), slowest rank {lo.item():.1f}, aggregate {t.item():.1f} GB/s; "
          f"an 8-frame call's transfers ({(D2H + H2D) / 1e6:.1f} MB) take {1e3 * (D2H + H2D) / (lo.item() * 1e9):.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
