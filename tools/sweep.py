#!/usr/bin/env python
"""Device-resident throughput sweep over the BASELINE.json configs other than the headline one
(C1 cu8/N=1024, C3 cf32/N=32768 zoom x1..x8, C4 packed formats x N=128..65536, C5 cf32/N=65536).
One JSON line per case: Msamples/s, algorithmic GB/s (sampleWidth + 4*n/H bytes per sample)
and the fraction of the measured HBM peak.  usage (under gpurun): python tools/sweep.py > gpurun_out/sweep.jsonl"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spectroplot-js_b200")):
    sys.path.insert(0, p)
import torch
import spectro_b200
from spectro_b200 import windows, cmaps, _lib

def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0

def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    eng = spectro_b200.Engine(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    steps = 5
    cases = []
    cases.append(("C1", "CU8", 1024, 1, "hann", "cube1", 1 << 26))
    cases.append(("C2", "CS16", 4096, 1, "blackmanHarris", "viridis", 100 << 20))
    cases.append(("C2-hann", "CS16", 4096, 1, "hann", "viridis", 100 << 20))
    for z in (2, 4, 8):
        cases.append((f"C2-z{z}", "CS16", 4096, z, "hann", "viridis", 1 << 26))
    for z in (1, 2, 4, 8):
        cases.append((f"C3-z{z}", "CF32", 32768, z, "hann", "inferno", 1 << 27))
    for fmt in ("CS4", "CU12"):
        for n in (128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
            cases.append(("C4", fmt, n, 1, "blackmanHarris", "viridis", 1 << 26))
    for fmt in ("CU4", "CS12", "CS8", "CU16", "CF32"):
        cases.append(("C4", fmt, 1024, 1, "hann", "viridis", 1 << 26))
    cases.append(("C5", "CF32", 65536, 1, "hann", "viridis", 1 << 28))
    pk = peak()
    only = sys.argv[1].split(",") if len(sys.argv) > 1 else None      # e.g. "C5,C3-z1", or ad-hoc cases "X:CS16:65536:1:28" (format, n, zoom, log2 samples)
    for spec in (only or []):
        if spec.startswith("X:"):
            _, f_, n_, z_, l_ = spec.split(":")
            cases.append((spec, f_, int(n_), int(z_), "hann", "viridis", 1 << int(l_)))
    if len(sys.argv) > 2:
        steps = int(sys.argv[2])
    for tag, fmt, n, z, win, cmname, S in cases:
        if only and tag not in only:
            continue
        sw = _lib.load().sp_sample_width(_lib.format_id(fmt))
        width = z * S // n
        if z > 1:
            width = width // 8 * 8
        w = getattr(windows, win + "Window")(n)
        cm = [list(c) for c in cmaps.cmaps[cmname + "_cmap"]]
        cm[0] = [0, 0, 0]; cm[-1] = [255, 255, 255]
        cmb = cmaps.cmap_bytes(cm)
        nbytes = S * sw
        d_in = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        eng.synth_fill(d_in.data_ptr(), fmt, 0, S, S, 0x5EC70010)
        d_img = torch.empty(4 * width * n, dtype=torch.uint8, device=dev)
        d_g = torch.empty(3 * width, dtype=torch.uint8, device=dev)
        d_hist = torch.zeros(1000 + len(cmb), dtype=torch.int64, device=dev)
        d_mm = torch.zeros(2, dtype=torch.float64, device=dev)
        ww = np.array(w["window"], np.float64)
        def step():
            rq, keep = eng.make_request(d_in.data_ptr(), fmt, n, width, ww, 1.0 / float(w["weight"]), 6, 30, cmb, byte_length=nbytes)
            return eng.render_enqueue(rq, d_img.data_ptr(), (d_g.data_ptr(), d_g.data_ptr() + width, d_g.data_ptr() + 2 * width),
                                      d_hist.data_ptr(), d_hist.data_ptr() + 8000, d_mm.data_ptr())
        for _ in range(3):
            rp = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            rp = step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        eng.render_finish(rp)
        total = int(d_hist[1000:].sum().item())
        hop = (S - n) / (width - 1)
        alg = S * sw + 4.0 * width * n
        line = dict(case=tag, fmt=fmt, n=n, zoom=z, window=win, samples=S, width=width, hop=hop, ms_per_render=ms,
                    msamples_s=S / ms / 1e3, alg_gb=alg / 1e9, alg_gbs=alg / ms / 1e6, frac_of_measured_hbm=alg / ms / 1e6 / pk,
                    launches=rp.kernel_launches, hist_ok=(total == width * n), plan=eng.kernel_plan(fmt, n))
        print(json.dumps(line), flush=True)
        del d_in, d_img, d_g
        torch.cuda.empty_cache()
    eng.close()

if __name__ == "__main__":
    main()
