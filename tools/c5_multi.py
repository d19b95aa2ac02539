#!/usr/bin/env python
"""BASELINE.json config 5 as stated: ONE cf32 capture of 2^33 samples (8 GSamples, 64 GiB), FFT N = 65536, hop N
(width 131 072), frame-range sharded with a window-length halo across the GPUs of one box, histograms and min / max
merged with one NCCL all-gather.  STRONG scaling: the capture is fixed, each rank renders 1/world of the frames.

  python tools/c5_multi.py                                            # 1 GPU (the whole capture on one B200)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_multi.py

Checks: colour-histogram total over all ranks == 2^33 pixels (64-bit counters); the first and last frame of every shard
equal the float64 oracle's render of the same samples (a shard edge is where a halo or a position error would show);
merged dBfs_min / dBfs_max printed so runs at different world sizes can be compared.  One JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spectroplot-js_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import spectro_b200
from spectro_b200 import windows, sharding
from oracle import oracle as O
from helpers import injective_cmap, cmap_index_image, bin_to_row

FMT, N, SW, SEED = "CF32", 65536, 8, 0x5EC70005
S = int(os.environ.get("C5_SAMPLES", str(1 << 33)))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    out_fd = os.dup(1)
    os.dup2(2, 1)                                            # library chatter (NCCL banner) off stdout
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = spectro_b200.Engine(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    W = S // N
    sh = sharding.plan_shards(S, N, W, world)[rank]
    width, nbytes = sh["width"], sh["sample_count"] * SW
    cm = injective_cmap(256)
    cm[0] = [0, 0, 0]; cm[-1] = [255, 255, 255]
    w = windows.hannWindow(N)
    ww, wt = np.array(w["window"], np.float64), float(w["weight"])
    d_in = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    eng.synth_fill(d_in.data_ptr(), FMT, sh["sample_first"], sh["sample_count"], S, SEED)
    d_img = torch.empty(4 * width * N, dtype=torch.uint8, device=dev)
    d_g = torch.empty(3 * width, dtype=torch.uint8, device=dev)
    d_stats, d_hist, d_mm, d_gath = sharding.stats_buffers(torch, 1000 + len(cm), world, dev)
    shard = sharding.shard_fields(sh, S, SW, W) if world > 1 else None

    def step():
        rq, keep = eng.make_request(d_in.data_ptr(), FMT, N, width, ww, 1.0 / wt, 6, 30, cm, byte_length=nbytes, shard=shard)
        rp = eng.render_enqueue(rq, d_img.data_ptr(), (d_g.data_ptr(), d_g.data_ptr() + width, d_g.data_ptr() + 2 * width),
                                d_hist.data_ptr(), d_hist.data_ptr() + 8000, d_mm.data_ptr())
        if world > 1:
            sharding.gather_stats(dist, d_stats, d_gath)     # the one exchange step: ~10 KB per rank over NVLink
        return rp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(2):
        rp = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    barrier()
    e0.record(stream)
    for _ in range(steps):
        rp = step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    eng.render_finish(rp)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        hist, mn, mx = sharding.fold_gathered(torch, d_gath, 1000 + len(cm))
    else:
        hist, mn, mx = d_hist, rp.dBfs_min, rp.dBfs_max
    ms = float(t.item())
    c_total = int(hist[1000:].sum().item())
    # first and last frame of this shard against the oracle (global positions, global stride)
    stride = (S - N) / (W - 1)
    img = d_img.view(N, width, 4)
    rows = bin_to_row(N)
    worst = 0
    for xl in (0, width - 1):
        xg = sh["frame_first"] + xl
        pos = int(0.5 + stride * xg)
        fb = O.synth(FMT, pos, N, S, SEED).tobytes()
        o = O.render(fb + fb, FMT, N, 2, ww, 1.0 / wt, 6, 30, cm, taps=True)
        col = img[:, xl, :].cpu().numpy()
        gi = cmap_index_image(col[None], cm)[0][rows]
        d_ = gi.astype(int) - o.gray[0].astype(int)
        assert np.abs(d_).max() <= 1, (rank, xg)
        worst = max(worst, int((d_ != 0).sum()))
    tw = torch.tensor([worst], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    if rank == 0:
        alg = S * SW + 4.0 * W * N
        line = dict(case="C5", world=world, scaling="strong", samples=S, n=N, width=W, frames_per_rank=width,
                    shard_bytes=nbytes, halo_samples=sh["sample_count"] - width * N, ms_per_render=ms, msamples_s=S / ms / 1e3,
                    alg_gbs_total=alg / ms / 1e6, c_hist_total=c_total, pixels=W * N, hist_ok=c_total == W * N,
                    worst_pixels_off_in_an_edge_frame=int(tw.item()), dBfs_min=mn, dBfs_max=mx, launches_per_rank=rp.kernel_launches)
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
