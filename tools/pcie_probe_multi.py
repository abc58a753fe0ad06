"""Aggregate host <-> device copy rate with N GPUs copying at the same time (one process per GPU under torchrun, pinned host
buffers of the e2e arm's size: 55 MB up, 21 MB down per step), barrier-bracketed: what the box's host side can feed, measured
without any of this repository's code in the timed region.  usage: torchrun --nproc-per-node N tools/pcie_probe_multi.py"""
import os, time
import torch
import torch.distributed as dist
rank = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dev = torch.device("cuda", rank)
up, down = 55 * 2**20 // 8, 21 * 2**20 // 8
h = torch.empty(up, dtype=torch.float64).pin_memory(); d = torch.empty(up, dtype=torch.float64, device=dev)
h2 = torch.empty(down, dtype=torch.float64).pin_memory(); d2 = torch.ones(down, dtype=torch.float64, device=dev)
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
for mode in ("h2d", "d2h", "both"):
    for rep in range(2):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        for _ in range(40):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        dt = (time.perf_counter() - t0) / 40
    gb = ((55 if mode != "d2h" else 0) + (21 if mode != "h2d" else 0)) / 1024
    if rank == 0:
        print(f"{world} GPUs, {mode}: {gb / dt:.1f} GB/s per GPU, {world * gb / dt:.1f} GB/s aggregate ({dt * 1e3:.2f} ms per step)", flush=True)
if world > 1: dist.destroy_process_group()
