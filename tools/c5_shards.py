#!/usr/bin/env python
"""BASELINE.json config 5 through the C ABI alone (no torch, no launcher): ONE cf32 capture of C5_SAMPLES samples
(default 2^33), FFT N = 65536, hop N, frame-range sharded with a window-length halo across the GPUs of ONE multi-device
engine (sp_create with ndev = C5_GPUS), every shard generated on its own device, rendered there and merged by the
engine's grouped NCCL all-reduce (sp_render_shards).  STRONG scaling: the capture is fixed.

  C5_GPUS=2 python tools/c5_shards.py        (under `gpurun --gpus 2`)

One JSON line: ms per render (wall clock around sp_render_shards, all devices synchronised; best of 3), Msamples/s,
histogram total == 2^33 pixels on EVERY device, merged dBfs range."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spectroplot-js_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import spectro_b200
from spectro_b200 import windows, sharding, _lib
from helpers import injective_cmap

FMT, N, SW, SEED = "CF32", int(os.environ.get("C5_N", "65536")), 8, 0x5EC70005
S = int(os.environ.get("C5_SAMPLES", str(1 << 33)))
G = int(os.environ.get("C5_GPUS", "2"))


def main():
    eng = spectro_b200.Engine(list(range(G))) if G > 1 else None
    assert eng is not None, "C5_GPUS >= 2 (one GPU: tools/fullsize.py)"
    W = S // N
    plan = sharding.plan_shards(S, N, W, G)
    cm = injective_cmap(256)
    cm[0] = [0, 0, 0]; cm[-1] = [255, 255, 255]
    w = windows.hannWindow(N)
    ww, wt = np.array(w["window"], np.float64), float(w["weight"])
    rqs, rps, bufs, keep = [], [], [], []
    p = lambda v: C.c_void_p(int(v))
    for g, sh in enumerate(plan):
        eng.select_device(g)
        nb = sh["sample_count"] * SW
        d_in = eng.alloc(nb + 256)
        eng.synth_fill(d_in, FMT, sh["sample_first"], sh["sample_count"], S, SEED)
        d_img, d_g = eng.alloc(4 * sh["width"] * N), eng.alloc(3 * sh["width"])
        d_hist, d_mm = eng.alloc(8 * (1000 + len(cm))), eng.alloc(16)
        rq, k = eng.make_request(d_in, FMT, N, sh["width"], ww, 1 / wt, 6, 30, cm, byte_length=nb, shard=sharding.shard_fields(sh, S, SW, W))
        keep.append(k)
        rqs.append(rq)
        rps.append(_lib.Reply(p(d_img), p(d_g), p(d_g + sh["width"]), p(d_g + 2 * sh["width"]), p(d_hist), p(d_hist + 8000), 0.0, 0.0, 0.0, 0, p(d_mm)))
        bufs.append((d_in, d_img, d_g, d_hist, d_mm))
    times = []
    for i in range(4):
        t0 = time.perf_counter()
        out = eng.render_shards(rqs, rps)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * min(times[1:])
    totals, launches = [], 0
    for g in range(G):
        eng.select_device(g)
        hist = np.empty(1000 + len(cm), np.uint64)
        eng.d2h(hist, bufs[g][3])
        totals.append(int(hist[1000:].sum()))
        launches += out[g].kernel_launches
    line = dict(case="C5 strong scaling through sp_render_shards (C ABI, NCCL merge inside the engine, no torch)", gpus=G, samples=S, n=N, width=W,
                halo_samples=N, ms_per_render=ms, device_ms_max=max(float(o.device_ms) for o in out), msamples_s=S / ms / 1e3,
                c_hist_total_per_device=totals, hist_ok=all(t == W * N for t in totals), dBfs_min=out[0].dBfs_min, dBfs_max=out[0].dBfs_max,
                stats_equal_on_all_devices=all(o.dBfs_min == out[0].dBfs_min and o.dBfs_max == out[0].dBfs_max for o in out),
                kernel_launches=launches)
    print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
