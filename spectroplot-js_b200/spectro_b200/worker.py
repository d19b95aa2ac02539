"""GpuWorker — the render worker behind the reference's message protocol.

Stands where lib/worker.js stands: any object with `postMessage(msg, transfer)` and a settable
`onmessage` can be handed to `startWorkers` (reference lib/spectroplot.js:100-130).  Request and
reply carry exactly the fields of lib/spectroplot.js:1213-1226 and lib/worker.js:140-149; one
reply per request, in order; messages without `.buffer` are ignored (lib/worker.js:159).
The work itself is one call through the C ABI (sp_render) on one GPU.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .cmaps import cmap_bytes
from .samples import SampleView


def renderFft(engine: _lib.Engine, ctx: dict) -> dict:
    """Replaces renderFft(ctx) of reference lib/worker.js:23-156; returns the reply's `data`."""
    view = SampleView(ctx["format"])                       # alias table + unknown -> CU8
    buf = ctx["buffer"]
    if len(buf) % view.elementSize:
        raise ValueError("RangeError: byte length of typed array should be a multiple of %d" % view.elementSize)
    cmap = cmap_bytes(ctx["cmap"])
    n = int(ctx["n"])
    width = int(ctx["width"])
    out = engine.render(buf, view.format_id(), n, width, ctx["windowc"], ctx["block_norm"], ctx["gain"],
                        ctx["range"], cmap, ctx.get("channelMode", False), ctx.get("waterfall", False),
                        shard=ctx.get("_shard"))
    return {
        "cB_hist": out["cB_hist"],
        "c_hist": out["c_hist"],
        "dBfs_min": out["dBfs_min"],
        "dBfs_max": out["dBfs_max"],
        "offset": ctx.get("offset"),
        "gauge_mins": out["gauge_mins"],
        "gauge_maxs": out["gauge_maxs"],
        "gauge_amps": out["gauge_amps"],
        "imageData": {"data": out["image"].reshape(-1)},
        "_device_ms": out["device_ms"],
    }


class GpuWorker:
    """Worker-shaped front door: `w = GpuWorker(); w.onmessage = cb; w.postMessage(msg)`."""

    _next_device = 0

    def __init__(self, device: int | None = None):
        if device is None:
            device = 0
        self.engine = _lib.Engine(device)
        self.onmessage = None

    def postMessage(self, msg, transfer=None):
        if not msg or msg.get("buffer") is None:          # lib/worker.js:159
            return
        data = renderFft(self.engine, msg)
        if self.onmessage:
            self.onmessage({"data": data})

    def terminate(self):
        self.engine.close()
