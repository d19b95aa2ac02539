"""Small renders through every kernel family of the C ABI, for compute-sanitizer (memcheck / racecheck / synccheck).
usage (under gpurun): compute-sanitizer --tool memcheck python tools/san_multi.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spectroplot-js_b200")]
import spectro_b200
from oracle import oracle as O
cm = np.stack([np.arange(256)] * 3, 1).astype(np.uint8)
eng = spectro_b200.Engine(0)
#        fmt    n     width waterfall channelMode   kernel family
cases = [("CS16", 4096, 32, False, False),      # render_r64_kernel, full tiles
         ("CS16", 4096, 21, False, False),      # r64 partial tile + generic remainder, unaligned rows
         ("CU8", 1024, 64, False, False),       # render_rc_kernel
         ("CF32", 256, 72, False, False),       # rc, C = 4
         ("CS4", 128, 40, False, False),        # render_kernel (generic, two passes)
         ("CU12", 512, 24, False, True),        # split-real
         ("CS8", 2048, 13, True, False),        # waterfall layout
         ("CF32", 8192, 16, False, False),      # four-step: prepass_kernel + r64 SUB
         ("CS16", 16384, 9, False, True),       # four-step + spectrum epilogue
         ("CS16", 4096, 40, True, False),       # waterfall rows from the r64 store warps (full + partial tile)
         ("CU8", 1024, 72, True, False),        # waterfall rows from the rc store warps
         ("CF32", 256, 300, True, False),       # rc C = 4, waterfall
         # round 2
         ("CS16", 4096, 64, False, False),      # r64 with tensor-TMA row stores (width % 8 == 0: RGBA tiles, UTMASTG)
         ("CS16", 128, 296, False, False),      # render_w_kernel 16 x 8, 4 frames per warp, span staging (overlapping hops), partial tile
         ("CU8", 512, 104, False, False),       # render_w_kernel 32 x 16 (twiddles from shared memory)
         ("CF32", 1024, 56, False, False),      # render_w_kernel 32 x 32
         ("CU12", 64, 520, False, False),       # render_w_kernel 8 x 8
         ("CF32", 8192, 40, False, False),      # four-step, L2-ring form (render_big_kernel; SP_FOURSTEP=ring below)
         ("CS16", 65536, 24, False, False),     # ring form, R = 16
         # round 2, second half
         ("CU8", 1024, 72, True, False),        # render_w_kernel 32 x 32, waterfall rows from its store warps (case 10 ran on rc in round 1)
         ("CS16", 128, 296, True, True),        # render_w_kernel 16 x 8, split-real + waterfall, partial tile
         ("CF32", 512, 136, False, True),       # render_w_kernel 32 x 16, split-real
         ("CS8", 2048, 40, True, False)]        # render_rc_kernel N = 2048 with the interleaved exchange stores, waterfall
os.environ.setdefault("SP_FOURSTEP", "ring")   # cases 7, 8 predate the ring kernel and ask for it too now; the HBM form keeps its round-1 record
only = [int(a) for a in sys.argv[1:]]
for i, (fmt, n, width, wf, chm) in enumerate(cases):
    if only and i not in only:
        continue
    S = n * (width // 2 + 2) + 5
    buf = O.synth(fmt, 0, S, S, 7 + i).tobytes()
    w, wt = O.window("hann", n)
    r = eng.render(buf, fmt, n, width, w, 1 / wt, 6, 30, cm, channel_mode=chm, waterfall=wf)
    assert int(r["c_hist"].sum()) == n * width
    print("ok", i, fmt, n, width, wf, chm, r["kernel_launches"], eng.kernel_plan(fmt, n)[:60], flush=True)
eng.close()
