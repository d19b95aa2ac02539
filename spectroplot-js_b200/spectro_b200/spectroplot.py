"""Headless Spectroplot — the reference's public surface for the render path.

`Spectroplot(options)`, `setOption`, `setOptions`, `setData`, `zoomIn/Out/Fit` keep their
names, defaults, `constrain` parsers and single-flight behaviour (reference
lib/spectroplot.js:212-306, 461-527, 1096-1285).  Everything that needs a DOM (canvases, axes,
events, themes) is dropped: results are exposed as buffers (`image`, `cB_hist`, `c_hist`,
gauges, `dBfs_min/max`).  `processData` builds the worker messages exactly like
lib/spectroplot.js:1206-1228 (disjoint halo-free slices, one per worker) and merges the replies
like :1229-1268.
"""
from __future__ import annotations

import numpy as np

from . import windows as _windows
from .cmaps import cmaps
from .parse_freq_rate import parseFormat, parseFreqRate
from .samples import SampleView
from .utils import js_parse_int, lookup
from .worker import GpuWorker

cB_hist_size = 1000


class Spectroplot:
    def __init__(self, options=None):
        W = _windows.windows
        self.constrain = {                                                    # lib/spectroplot.js:238-251
            "fftN": lambda v: js_parse_int(v, 512),
            "height": lambda v: js_parse_int(v, 0),
            "windowF": lambda v: lookup(W, v) or W["blackmanHarrisWindow"],
            "zoom": lambda v: js_parse_int(v, 1),
            "gain": lambda v: js_parse_int(v, 0),
            "range": lambda v: js_parse_int(v, 30),
            "cmap": lambda v: lookup(cmaps, v) or cmaps["cube1_cmap"],
            "ampHeight": lambda v: js_parse_int(v, 0),
            "minmaxHeight": lambda v: js_parse_int(v, 0),
            "histWidth": lambda v: js_parse_int(v, 0),
            "channelMode": lambda v: (not v.lower().startswith("i")) if isinstance(v, str) else v,
            "turnFlip": lambda v: (not v.lower().startswith("s")) if isinstance(v, str) else v,
        }
        defaults = dict(fftN=512, width=3000, height=512, zoom=1, windowF=W["blackmanHarrisWindow"], gain=6,
                        range=30, cmap=cmaps["cube1_cmap"], ampHeight=0, minmaxHeight=20, channelMode=False,
                        turnFlip=False, dbfsWidth=60, dbfsHeight=0, freqWidth=40, timeHeight=20, rampHeight=0,
                        rampTop=10, rampWidth=15, histWidth=100, histLeft=55,
                        # headless stand-ins for parent.clientWidth / window.innerHeight
                        clientWidth=3200, innerHeight=3200, workerCount=1, workerOrUrl=None, devices=None)
        options = {**defaults, **(options or {})}
        self.buffer = None
        self.fileinfo = None
        self.fftN = options["fftN"]
        self.width = options["width"]
        self.height = options["height"]
        self.zoom = options["zoom"]
        self.windowF = lookup(W, options["windowF"])
        self.gain = options["gain"]
        self.range = options["range"]
        self.cmap = lookup(cmaps, options["cmap"])
        self.ampHeight = options["ampHeight"]
        self.minmaxHeight = options["minmaxHeight"]
        self.channelMode = options["channelMode"]
        self.turnFlip = options["turnFlip"]
        self.histWidth = options["histWidth"]
        self.opts = options
        self.inProcess = False
        self.sampleView = None
        self.result = None
        self._workers = []
        self._start_workers(options["workerOrUrl"], options["workerCount"], options["devices"])
        if options.get("filedata"):
            self.setData(options["filedata"])

    # lib/spectroplot.js:100-130 — the worker pool; a constructor may be supplied
    def _start_workers(self, workerOrUrl, count, devices):
        ctor = workerOrUrl or GpuWorker
        for i in range(count):
            w = ctor(devices[i % len(devices)]) if devices else ctor()
            self._workers.append(w)

    def destroy(self):
        for w in self._workers:
            if hasattr(w, "terminate"):
                w.terminate()
        self._workers = []

    def setOption(self, opt, value):                                          # :461-464
        setattr(self, opt, self.constrain[opt](value))
        return self.processData()

    def setOptions(self, opts):                                               # :471-476
        for opt in opts:
            setattr(self, opt, self.constrain[opt](opts[opt]))
        return self.processData()

    def setData(self, filedata):                                              # :483-511
        if isinstance(filedata, str):                                         # :486-487 loads a URL; headless: a local file
            import os
            if not os.path.exists(filedata):
                raise NotImplementedError("URL loading needs XHR (out of scope); pass a local path or {fileBuffer, name, size, type}")
            from .ingest import load_capture
            filedata = load_capture(filedata)
        self.fileinfo = filedata
        self.buffer = filedata["fileBuffer"]
        self.sampleFormat = parseFormat(filedata.get("name", ""))
        nameInfo = parseFreqRate(filedata.get("name", ""))
        self.center_freq = nameInfo["freq"]
        self.sample_rate = nameInfo["rate"]
        self.sampleView = SampleView(self.sampleFormat, None, self.sample_rate, self.center_freq)
        self.sampleView.loadBuffer(self.buffer)
        self.sample_rate = self.sampleView.sampleRate
        return self.processData()

    def zoomOut(self):                                                        # :513-517
        if self.zoom <= 1:
            return None
        self.zoom -= 0.5
        return self.processData()

    def zoomFit(self):
        if self.zoom == 1:
            return None
        self.zoom = 1
        return self.processData()

    def zoomIn(self):                                                         # :523-527
        if self.zoom >= 8:
            return None
        self.zoom += 0.5
        return self.processData()

    def processData(self):                                                    # :1096-1285
        if self.buffer is None or len(self.buffer) == 0:
            return None
        if not self.sampleView or self.sampleView.buffer is None or len(self.sampleView.buffer) == 0:
            return None
        if self.inProcess:
            return self.inProcess                                             # single flight (:1099)
        self.inProcess = True
        try:
            waterfall = self.turnFlip
            extraWidth = self.opts["freqWidth"] + self.opts["dbfsWidth"] + self.histWidth
            self.width = int((self.opts["innerHeight"] if waterfall else self.opts["clientWidth"]) * self.zoom - extraWidth)
            sv = self.sampleView
            n = self.fftN
            ww = self.windowF(n)
            windowc, weight = ww["window"], ww["weight"]
            block_norm = 1.0 / weight                                         # :1116
            self.dBfs_min = 0.0
            self.dBfs_max = -200.0
            cmap = self.cmap
            cmap[0] = [0, 0, 0]                                               # :1129-1130 (mutates the table)
            cmap[len(cmap) - 1] = [255, 255, 255]
            cB_hist = np.zeros(cB_hist_size, np.uint64)
            c_hist = np.zeros(len(cmap), np.uint64)
            width = self.width
            height = n
            count = len(self._workers)
            startSample = 0
            endSample = int(len(self.buffer) / sv.sampleWidth)                # :1207
            sliceWidth = int(width / count)                                   # :1208
            image = np.zeros((width, height, 4) if waterfall else (height, width, 4), np.uint8)
            gmin = np.zeros(width, np.uint8); gmax = np.zeros(width, np.uint8); gamp = np.zeros(width, np.uint8)
            for i, worker in enumerate(self._workers):
                bufferSlice = sv.slice(i, count, startSample, endSample)      # :1211
                fftCtx = dict(block_norm=block_norm, gain=self.gain, range=self.range, cmap=cmap, n=n,
                              windowc=windowc, width=sliceWidth, offset=i * sliceWidth, buffer=bufferSlice,
                              format=sv.format, channelMode=self.channelMode, waterfall=waterfall)
                replies = []
                worker.onmessage = replies.append
                worker.postMessage(fftCtx, [bufferSlice])
                d = replies[0]["data"]
                if d["dBfs_min"] < self.dBfs_min: self.dBfs_min = d["dBfs_min"]   # :1230
                if d["dBfs_max"] > self.dBfs_max: self.dBfs_max = d["dBfs_max"]   # :1231
                cB_hist += d["cB_hist"]
                c_hist += d["c_hist"]
                off = d["offset"]
                tile = np.asarray(d["imageData"]["data"])
                if waterfall:                                                 # putImageData(.., 0, width - sliceWidth - offset)
                    row0 = width - sliceWidth - off
                    image[row0:row0 + sliceWidth] = tile.reshape(sliceWidth, height, 4)
                else:                                                         # putImageData(.., offset, 0)
                    image[:, off:off + sliceWidth] = tile.reshape(height, sliceWidth, 4)
                gmin[off:off + sliceWidth] = d["gauge_mins"]
                gmax[off:off + sliceWidth] = d["gauge_maxs"]
                gamp[off:off + sliceWidth] = d["gauge_amps"]
            self.result = dict(image=image, cB_hist=cB_hist, c_hist=c_hist, gauge_mins=gmin, gauge_maxs=gmax,
                               gauge_amps=gamp, dBfs_min=self.dBfs_min, dBfs_max=self.dBfs_max, width=width,
                               height=height)
            return self.result
        finally:
            self.inProcess = False
