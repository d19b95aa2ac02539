#!/usr/bin/env python
"""BASELINE.json config 1 exactly as stated: cu8, 10 s at 1 MS/s (10 M complex samples), FFT N = 1024, Hann, zoom x1, Cube1 -
the reference's own CPU-runnable case.  Two widths: hop N (9 765 frames over the first 9 999 360 samples) and the reference's default
canvas (3 000 px - 200 px of axes = 2 800 frames over all 10 M samples, fractional stride, data skipped).  For each: the C float64
port of the worker on every host core (the stand-in for the Node.js worker pool), the engine host -> host (pinned) and device
resident, and the parity of the two pictures.  usage (under gpurun): python tools/c1_exact.py > gpurun_out/c1_exact.jsonl"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spectroplot-js_b200")):
    sys.path.insert(0, p)
import torch
import spectro_b200
from spectro_b200 import windows, cmaps
from oracle import oracle as O

n, S_all = 1024, 10_000_000
raw = O.synth("CU8", 0, S_all, S_all, 0x5EC70001)
w = windows.hannWindow(n)
ww, wt = np.array(w["window"], np.float64), float(w["weight"])
cm = [list(c) for c in cmaps.cmaps["cube1_cmap"]]
cm[0] = [0, 0, 0]; cm[-1] = [255, 255, 255]
cmb = cmaps.cmap_bytes(cm)
eng = spectro_b200.Engine(0)
cores = os.cpu_count() or 1
for name, S, width in (("hop N", 9765 * n, 9765), ("canvas 2800", S_all, 2800)):
    buf = raw[:2 * S]
    pin = spectro_b200.PinnedBuffer(2 * S); pin.array[:] = buf
    img = spectro_b200.PinnedBuffer(4 * width * n)
    ts = []
    for i in range(6):
        t0 = time.perf_counter(); out = eng.render(pin.array, "CU8", n, width, ww, 1 / wt, 6, 30, cmb, out_image=img.array); ts.append(time.perf_counter() - t0)
    e2e_ms = 1e3 * min(ts[1:])
    t0 = time.perf_counter(); ora = O.render(buf.tobytes(), "CU8", n, width, ww, 1 / wt, 6, 30, cmb, workers=cores); cpu_ms = 1e3 * (time.perf_counter() - t0)
    one = O.render(buf.tobytes(), "CU8", n, width, ww, 1 / wt, 6, 30, cmb)            # the unsliced message, for parity
    bad = int((out["image"] != one.image).any(axis=2).sum())
    print(json.dumps(dict(case="C1 " + name, samples=S, width=width, hop=(S - n) / (width - 1), engine_host_to_host_ms=e2e_ms,
                          engine_device_ms=out["device_ms"], engine_msamples_s_e2e=S / e2e_ms / 1e3, cpu_port_ms=cpu_ms, cpu_cores=cores,
                          cpu_msamples_s=S / cpu_ms / 1e3, speedup_e2e=cpu_ms / e2e_ms, pixels=width * n, pixels_differing=bad,
                          dBfs_max_gpu=out["dBfs_max"], dBfs_max_cpu=one.dBfs_max, plan=eng.kernel_plan("CU8", n))), flush=True)
    pin.free(); img.free()
eng.close()
