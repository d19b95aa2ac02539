"""One small N=4096 render through the C ABI (for compute-sanitizer runs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spectroplot-js_b200")]
import spectro_b200
from oracle import oracle as O
n, width = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = n * width
buf = O.synth("CS16", 0, S, S, 1).tobytes()
w, wt = O.window("hann", n)
cm = np.stack([np.arange(256)] * 3, 1).astype(np.uint8)
eng = spectro_b200.Engine(0)
r = eng.render(buf, "CS16", n, width, w, 1 / wt, 6, 30, cm)
print("ok", int(r["c_hist"].sum()), r["dBfs_min"], r["dBfs_max"], eng.kernel_plan("CS16", n))
