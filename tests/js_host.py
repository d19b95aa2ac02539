"""Runs the repository's JavaScript host layer (spectroplot-js_b200/js/*.js) under oracle/jsmini.py (no Node.js in the
image).  The N-API addon those files `require` is replaced by an object with the same three functions
(create / render / destroy, see js/spectro_napi.c) that calls a Python render function: the C-ABI engine on a GPU
box, the float64 oracle in the CPU suite."""
import os

import numpy as np

from oracle.jsmini import Interp, JSObject, JSArray, JSTypedArray, JSArrayBuffer, NativeFunction, UNDEF, JSThrow

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
JS = os.path.join(ROOT, "spectroplot-js_b200", "js")
REFERENCE = "/root/reference/lib"
FORMATS = ["CU4", "CS4", "CU8", "CS8", "CU12", "CS12", "CU16", "CS16", "CU32", "CS32", "CU64", "CS64", "CF32", "CF64"]


def typed(I, kind, arr):
    buf = JSArrayBuffer(I, bytearray(np.ascontiguousarray(arr).tobytes()))
    return I.construct(I.globals.vars[kind], [buf])


def make_addon(I, render_fn, log):
    """render_fn(buf, fmt, n, width, windowc, block_norm, gain, range, cmap[len,3], channel_mode, waterfall) -> dict."""
    addon = JSObject(I.object_proto)
    engines = {}

    def create(this, a):
        h = JSObject(I.object_proto)
        engines[id(h)] = int(a[0]) if a and a[0] is not UNDEF else 0
        log.append(("create", engines[id(h)]))
        return h

    def destroy(this, a):
        log.append(("destroy", engines.pop(id(a[0]), None)))
        return UNDEF

    def render(this, a):
        ctx = a[1]
        g = lambda k: I.get_prop(ctx, k)
        assert id(a[0]) in engines, "render on a destroyed engine"
        buf = g("buffer")
        assert isinstance(buf, JSArrayBuffer), "ctx.buffer must be an ArrayBuffer (napi_get_arraybuffer_info)"
        wc, cm = g("windowc"), g("cmap")
        assert isinstance(wc, JSTypedArray) and wc.kind == "Float64Array", "ctx.windowc must be a Float64Array"
        assert isinstance(cm, JSTypedArray) and cm.kind in ("Uint8Array", "Uint8ClampedArray"), "ctx.cmap must be a Uint8Array or Uint8ClampedArray(len*3)"
        n, width = int(g("n")), int(g("width"))
        log.append(("render", FORMATS[int(g("format"))], n, width))
        try:
            r = render_fn(bytes(buf.data), FORMATS[int(g("format"))], n, width, wc.arr.copy(), float(g("block_norm")), float(g("gain")),
                          float(g("range")), cm.arr.reshape(-1, 3).copy(), bool(g("channelMode")), bool(g("waterfall")))
        except Exception as ex:                      # the addon throws a JS Error with sp_last_error()
            err = JSObject(I.object_proto)
            err.props["message"] = str(ex)
            raise JSThrow(err)
        out = JSObject(I.object_proto)
        out.props["image"] = typed(I, "Uint8ClampedArray", np.asarray(r["image"], np.uint8).reshape(-1))
        for k in ("gauge_mins", "gauge_maxs", "gauge_amps"):
            out.props[k] = typed(I, "Uint8ClampedArray", np.asarray(r[k], np.uint8))
        out.props["cB_hist"] = typed(I, "BigUint64Array", np.asarray(r["cB_hist"], np.uint64))
        out.props["c_hist"] = typed(I, "BigUint64Array", np.asarray(r["c_hist"], np.uint64))
        out.props["dBfs_min"], out.props["dBfs_max"] = float(r["dBfs_min"]), float(r["dBfs_max"])
        out.props["device_ms"] = float(r.get("device_ms", 0.0))
        return out
    for name, fn in (("create", create), ("destroy", destroy), ("render", render)):
        addon.props[name] = NativeFunction(I, name, fn)
    return addon


class JsHost:
    def __init__(self, render_fn):
        self.I = I = Interp(JS)
        self.log = []
        addon = make_addon(I, render_fn, self.log)
        self.worker_mod = I.require_file(os.path.join(JS, "gpu_worker.js"), lambda spec: addon if spec.endswith("spectro_napi.node") else None)
        self.GpuWorker = I.get_prop(self.worker_mod, "GpuWorker")
        self.formatId = I.get_prop(self.worker_mod, "formatId")

    def reference_deps(self):
        """The reference's pure modules, loaded from its own files (only where /root/reference exists)."""
        I = self.I
        deps = JSObject(I.object_proto)
        win = JSObject(I.object_proto)
        win.props.update(I.load_module("./windows", REFERENCE))
        cm = JSObject(I.object_proto)
        for m in ("cube1cmap", "matplotlibcmaps", "parabolacmap", "soxcmap", "naivecmap"):   # import order of lib/spectroplot.js
            cm.props.update(I.load_module("./" + m, REFERENCE))
        pfr = I.load_module("./parseFreqRate", REFERENCE)
        deps.props.update(windows=win, cmaps=cm, lookup=I.load_module("./utils", REFERENCE)["lookup"],
                          parseFreqRate=pfr["parseFreqRate"], parseFormat=pfr["parseFormat"],
                          SampleView=I.load_module("./samples", REFERENCE)["default"], Worker=self.GpuWorker)
        return deps

    def spectroplot_class(self, deps):
        create = self.I.require_file(os.path.join(JS, "spectroplot_headless.js"))
        return self.I.call(create, UNDEF, [deps])

    def await_(self, promise):
        self.I.drain()
        state, value = self.I.promise_state(promise)
        assert state != "pending", "promise still pending after the job queue drained"
        if state == "rejected":
            raise JSThrow(value)
        return value
