"""Extract the colormap DATA tables from the reference checkout into a compact binary file.

Run once in the build container (reads /root/reference, which does not exist on the GPU box):
    python tools/gen_cmap_data.py
Writes spectroplot-js_b200/spectro_b200/cmap_tables.npz (uint8 [len,3] per table).
Only the numeric tables are taken (data, not code):
  cube1      — Matteo Niccoli, mycarta.wordpress.com (credit required; lib/cube1cmap.js:7-14)
  viridis, plasma, inferno, magma, hot, afmhot, gist_heat — matplotlib, CC0 (lib/matplotlibcmaps.js:6-15)
  parabola   — 64 entries (lib/parabolacmap.js:2)
The computed maps (sox, naive, grayscale, roentgen, phosphor) are re-derived in code instead.
"""
import re, sys, os
import numpy as np

REF = "/root/reference/lib"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "spectroplot-js_b200", "spectro_b200", "cmap_tables.npz")

def tables(path):
    src = open(path).read()
    for m in re.finditer(r"export const (\w+)_cmap = \[(.*?)\n\]", src, re.S):
        rows = re.findall(r"\[\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\]", m.group(2))
        yield m.group(1), np.array(rows, dtype=np.uint8)

out = {}
for f in ("cube1cmap.js", "matplotlibcmaps.js", "parabolacmap.js"):
    for name, arr in tables(os.path.join(REF, f)):
        out[name] = arr
        print(name, arr.shape, arr[0], arr[-1])
np.savez_compressed(OUT, **out)
print("wrote", os.path.abspath(OUT))
