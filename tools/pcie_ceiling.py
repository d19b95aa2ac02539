#!/usr/bin/env python
"""Raw host <-> device copy ceiling of the box for the bench's e2e leg: every rank copies the bytes of one C2 message
(419 MB in, 419 MB out, pinned host memory) H2D and D2H at the same time on two streams, no kernel in between.
Run under torchrun with N = 1, 2, 4, 8 ranks; rank 0 prints one JSON line: ms per message, aggregate GB/s each way and the
Msamples/s the e2e leg could reach if rendering were free.  (why e2e does not scale past ~66 GB/s.)"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nb = 100 * (1 << 20) * 4
h_in = torch.empty(nb, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nb, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nb, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nb, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def once():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


for _ in range(2):
    once()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    once()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ms = 1e3 * (time.perf_counter() - t0) / reps
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"ranks": world, "bytes_each_way_per_rank": nb, "ms_per_message": ms, "aggregate_gbs_each_way": world * nb / ms / 1e6,
                      "e2e_ceiling_msamples_s": world * 100 * (1 << 20) / ms / 1e3}), flush=True)
if world > 1:
    dist.destroy_process_group()
