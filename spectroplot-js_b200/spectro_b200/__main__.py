"""python -m spectro_b200 CAPTURE [options] — render a capture file to PNG tiles on the GPU.

The headless equivalent of dropping a file on the reference's page: the file name supplies format, centre frequency and
sample rate (lib/parseFreqRate.js), the options are the reference's (`fftN`, `windowF`, `cmap`, `gain`, `range`, `zoom`,
`channelMode`, `turnFlip`), the picture width is `--width` frames (the page's canvas width).  Needs a B200: there is no
CPU path."""
from __future__ import annotations

import argparse
import json
import os
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m spectro_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("capture", help="raw I/Q file (.cu8 .cs16 .cf32 ...) or PCM .wav; e.g. g001_433.92M_250k.cu8")
    ap.add_argument("--out", default=None, help="output prefix (default: capture name without extension)")
    ap.add_argument("--fftN", default="512")
    ap.add_argument("--windowF", default="blackmanHarris")
    ap.add_argument("--cmap", default="cube1")
    ap.add_argument("--gain", default="6")
    ap.add_argument("--range", default="30")
    ap.add_argument("--zoom", default="1")
    ap.add_argument("--channelMode", default="I/Q")
    ap.add_argument("--turnFlip", default="spectrogram", help="'flip' renders the waterfall layout")
    ap.add_argument("--width", type=int, default=3000, help="frames (pixel columns) at zoom 1")
    ap.add_argument("--devices", default="0", help="comma-separated GPU indices; several GPUs shard the capture by frame range")
    ap.add_argument("--tile", type=int, default=4096, help="PNG tile width in frames")
    a = ap.parse_args(argv)

    from . import Spectroplot, egress
    from ._lib import Engine
    from .worker import GpuWorker
    devs = [int(d) for d in a.devices.split(",")]

    class Worker(GpuWorker):                       # one worker over all listed GPUs (multi-device engine) or one GPU
        def __init__(self):
            self.engine = Engine(devs if len(devs) > 1 else devs[0])
            self.onmessage = None
    extra = 40 + 60 + 100                          # freqWidth + dbfsWidth + histWidth of the page layout (lib/spectroplot.js:1103-1104)
    sp = Spectroplot({"fftN": 512, "clientWidth": a.width + extra, "innerHeight": a.width + extra, "workerOrUrl": Worker})
    sp.setOptions({"fftN": a.fftN, "windowF": a.windowF, "cmap": a.cmap, "gain": a.gain, "range": a.range, "zoom": a.zoom,
                   "channelMode": a.channelMode, "turnFlip": a.turnFlip})
    res = sp.setData(a.capture)
    prefix = a.out or os.path.splitext(a.capture)[0]
    img = res["image"]
    tiles = egress.write_tiles(os.path.dirname(prefix) or ".", img, a.tile, os.path.basename(prefix)) if img.shape[1] > a.tile else \
        [egress.write_png(prefix + ".png", img)]
    meta = {"capture": a.capture, "format": sp.sampleView.format, "center_freq": sp.center_freq, "sample_rate": sp.sample_rate,
            "fftN": sp.fftN, "width": res["width"], "height": res["height"], "dBfs_min": res["dBfs_min"], "dBfs_max": res["dBfs_max"],
            "tiles": tiles, "cB_hist": [int(v) for v in res["cB_hist"]], "c_hist": [int(v) for v in res["c_hist"]]}
    with open(prefix + ".json", "w") as f:
        json.dump(meta, f)
    sp.destroy()
    print(json.dumps({k: meta[k] for k in ("format", "center_freq", "sample_rate", "fftN", "width", "height", "dBfs_min", "dBfs_max", "tiles")}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
