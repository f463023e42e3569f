#!/usr/bin/env python3
"""Host->device copy bandwidth per GPU with all ranks copying at once (torchrun): is the platform the limit of the
N-GPU end-to-end number?  usage: torchrun --nproc-per-node N tools/h2d_probe.py"""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("gloo")
src = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); src.fill_(1)
dst = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
for _ in range(50): dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = 50 * 64 / 1024 / dt
if world > 1:
    lst = [None] * world
    dist.all_gather_object(lst, gbs)
else:
    lst = [gbs]
if rank == 0:
    print(f"H2D from pinned memory, 64 MiB x 50, {world} rank(s) at once: per-GPU GiB/s = {[round(x, 1) for x in lst]}, total {sum(lst):.1f}")
