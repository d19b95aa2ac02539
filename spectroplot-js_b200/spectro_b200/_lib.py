"""ctypes binding of libspectro_b200.so (the C ABI in include/spectro_b200.h).

This is the Python stand-in for the N-API addon of INTEGRATION.md (no Node.js toolchain in
this image): it unwraps buffers to pointers, calls the C entry points and wraps the reply.
There is NO CPU fallback: a missing library or a missing sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SP_LIB: an experiment build of the same C ABI (csrc/Makefile XFLAGS); the product default is the in-tree library
LIB_PATH = os.environ.get("SP_LIB") or os.path.join(_HERE, "..", "lib", "libspectro_b200.so")

CB_HIST_SIZE = 1000
MAX_CMAP = 4096
F_BUFFER_ON_DEVICE = 1
F_REPLY_ON_DEVICE = 2
F_NO_IMAGE = 4

FORMATS = ["CU4", "CS4", "CU8", "CS8", "CU12", "CS12", "CU16", "CS16",
           "CU32", "CS32", "CU64", "CS64", "CF32", "CF64"]

ERRORS = {0: "SP_OK", -1: "SP_E_INVAL", -2: "SP_E_BAD_N", -3: "SP_E_BAD_FORMAT", -4: "SP_E_TOO_SHORT",
          -5: "SP_E_BAD_WIDTH", -6: "SP_E_RAGGED", -7: "SP_E_BAD_CMAP", -8: "SP_E_CUDA", -9: "SP_E_NO_DEVICE",
          -10: "SP_E_RANGE", -11: "SP_E_ALIGN", -12: "SP_E_NCCL"}

# every symbol include/spectro_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ["sp_abi_version", "sp_format_from_name", "sp_format_name", "sp_sample_width", "sp_element_size",
           "sp_create", "sp_destroy", "sp_last_error", "sp_set_stream", "sp_render", "sp_render_enqueue",
           "sp_render_finish", "sp_render_zooms", "sp_decode", "sp_render_db", "sp_device_alloc", "sp_device_free", "sp_memcpy_h2d",
           "sp_memcpy_d2h", "sp_host_alloc_pinned", "sp_host_free_pinned", "sp_device_sync", "sp_synth_fill",
           "sp_synth_lut", "sp_device_count", "sp_sm_count", "sp_kernel_plan", "sp_profile_enable", "sp_profile_read",
           "sp_render_shards", "sp_select_device", "sp_build_id", "sp_profile_sample", "sp_render_async", "sp_render_wait"]


class SpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code
        self.name = ERRORS.get(code, str(code))


class Request(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("byte_length", C.c_uint64), ("format", C.c_int32), ("n", C.c_int32),
                ("width", C.c_int64), ("block_norm", C.c_double), ("gain", C.c_double), ("range", C.c_double),
                ("windowc", C.c_void_p), ("cmap_rgb", C.c_void_p), ("cmap_len", C.c_int32),
                ("channel_mode", C.c_int32), ("waterfall", C.c_int32), ("flags", C.c_uint32),
                ("total_byte_length", C.c_uint64), ("total_width", C.c_int64), ("frame_first", C.c_int64),
                ("buffer_first_sample", C.c_uint64)]


class Reply(C.Structure):
    _fields_ = [("image", C.c_void_p), ("gauge_mins", C.c_void_p), ("gauge_maxs", C.c_void_p),
                ("gauge_amps", C.c_void_p), ("cB_hist", C.c_void_p), ("c_hist", C.c_void_p),
                ("dBfs_min", C.c_double), ("dBfs_max", C.c_double), ("device_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("minmax_dev", C.c_void_p)]


_lib = None


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python __graft_entry__.py` "
                          "(or make -C spectroplot-js_b200/csrc); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.sp_format_name.restype = C.c_char_p
    lib.sp_build_id.restype = C.c_char_p
    lib.sp_last_error.restype = C.c_char_p
    lib.sp_last_error.argtypes = [C.c_void_p]
    lib.sp_kernel_plan.restype = C.c_char_p
    lib.sp_kernel_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.sp_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
    lib.sp_destroy.argtypes = [C.c_void_p]
    lib.sp_destroy.restype = None
    lib.sp_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    for fn in ("sp_render", "sp_render_enqueue"):
        getattr(lib, fn).argtypes = [C.c_void_p, C.POINTER(Request), C.POINTER(Reply)]
    lib.sp_render_finish.argtypes = [C.c_void_p, C.POINTER(Reply)]
    lib.sp_render_async.argtypes = [C.c_void_p, C.POINTER(Request), C.POINTER(Reply), C.POINTER(C.c_int)]
    lib.sp_render_wait.argtypes = [C.c_void_p, C.c_int]
    lib.sp_render_shards.argtypes = [C.c_void_p, C.POINTER(Request), C.POINTER(Reply)]
    lib.sp_select_device.argtypes = [C.c_void_p, C.c_int]
    lib.sp_render_db.argtypes = [C.c_void_p, C.POINTER(Request), C.c_void_p]
    lib.sp_render_zooms.argtypes = [C.c_void_p, C.POINTER(Request), C.c_int, C.POINTER(C.c_int64), C.POINTER(Reply)]
    lib.sp_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.sp_device_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.sp_device_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.sp_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.sp_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.sp_host_alloc_pinned.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
    lib.sp_host_free_pinned.argtypes = [C.c_void_p]
    lib.sp_device_sync.argtypes = [C.c_void_p]
    lib.sp_synth_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.sp_synth_lut.argtypes = [C.c_void_p]
    lib.sp_synth_lut.restype = None
    lib.sp_format_from_name.argtypes = [C.c_char_p]
    lib.sp_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.sp_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.sp_profile_sample.argtypes = [C.c_void_p, C.c_int]
    lib.sp_device_count.argtypes = [C.c_void_p]
    lib.sp_sm_count.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def build_id() -> str:
    return load().sp_build_id().decode()


def format_id(fmt) -> int:
    if isinstance(fmt, str):
        return load().sp_format_from_name(fmt.encode())
    return int(fmt)


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        rc = load().sp_host_alloc_pinned(int(nbytes), C.byref(self.ptr))
        if rc:
            raise SpError(rc, "cudaMallocHost failed")
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr.value))

    def free(self):
        if self.ptr:
            load().sp_host_free_pinned(self.ptr)
            self.ptr = C.c_void_p()
            self.array = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One sp_engine == one reference worker (sequential requests).  `device` is one GPU index, or a list of indices for a
    multi-device engine: `render` of a host-buffer message is then sharded by frame range across those GPUs inside the
    C ABI (sp_create with ndev > 1) and returns the single-device result."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        devs = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
        ids = (C.c_int * len(devs))(*devs)
        rc = self.lib.sp_create(C.byref(self.h), ids, len(devs))
        if rc:
            raise SpError(rc, (self.lib.sp_last_error(None) or b"").decode())
        self.device = devs[0]
        self.devices = devs

    def close(self):
        if self.h:
            self.lib.sp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SpError(rc, (self.lib.sp_last_error(self.h) or b"").decode())

    # ---- device memory
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.sp_device_alloc(self.h, int(nbytes), C.byref(p)))
        return p.value

    def free(self, dptr: int):
        self._check(self.lib.sp_device_free(self.h, C.c_void_p(dptr)))

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._check(self.lib.sp_memcpy_h2d(self.h, C.c_void_p(dptr), _vp(arr), arr.nbytes))

    def d2h(self, arr: np.ndarray, dptr: int):
        self._check(self.lib.sp_memcpy_d2h(self.h, _vp(arr), C.c_void_p(dptr), arr.nbytes))

    def sync(self):
        self._check(self.lib.sp_device_sync(self.h))

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.sp_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def sm_count(self) -> int:
        return self.lib.sp_sm_count(self.h)

    def kernel_plan(self, fmt, n, channel_mode=False) -> str:
        return self.lib.sp_kernel_plan(self.h, format_id(fmt), int(n), int(bool(channel_mode))).decode()

    def synth_fill(self, dptr: int, fmt, first: int, count: int, total: int, seed: int):
        self._check(self.lib.sp_synth_fill(self.h, C.c_void_p(dptr), format_id(fmt), first, count, total, seed))

    def profile_enable(self, slots: int):
        self._check(self.lib.sp_profile_enable(self.h, int(slots)))

    def profile_sample(self, every: int):
        """Bracket only every `every`-th render-kernel launch with events."""
        self._check(self.lib.sp_profile_sample(self.h, int(every)))

    def profile_read(self, max_n: int = 4096) -> np.ndarray:
        out = np.zeros(max_n, np.float32)
        n = self.lib.sp_profile_read(self.h, _vp(out), max_n)
        if n < 0:
            self._check(n)
        return out[:n]

    # ---- device-resident path (bench / multi-GPU host layer)
    def render_enqueue(self, rq: "Request", image_dev=0, gauges_dev=(0, 0, 0), cB_dev=0, c_dev=0, minmax_dev=0) -> "Reply":
        """Enqueue one render whose request buffer and reply buffers all live in HBM; no host sync."""
        rq.flags |= F_BUFFER_ON_DEVICE | F_REPLY_ON_DEVICE
        p = lambda v: C.c_void_p(int(v)) if v else None
        rp = Reply(p(image_dev), p(gauges_dev[0]), p(gauges_dev[1]), p(gauges_dev[2]), p(cB_dev), p(c_dev),
                   0.0, 0.0, 0.0, 0, p(minmax_dev))
        self._check(self.lib.sp_render_enqueue(self.h, C.byref(rq), C.byref(rp)))
        return rp

    def render_finish(self, rp: "Reply") -> "Reply":
        self._check(self.lib.sp_render_finish(self.h, C.byref(rp)))
        return rp

    # ---- multi-device engine: device-resident shards, NCCL merge inside the C ABI
    def select_device(self, index: int):
        """Which device of a multi-device engine alloc / free / h2d / d2h / synth_fill address."""
        self._check(self.lib.sp_select_device(self.h, int(index)))

    def render_shards(self, requests, replies):
        """sp_render_shards: requests[g] / replies[g] live on device g; every reply comes back with the merged histograms
        (device) and dBfs_min / dBfs_max (host fields) of the whole message."""
        n = len(self.devices)
        assert len(requests) == n and len(replies) == n
        rq, rp = (Request * n)(*requests), (Reply * n)(*replies)
        for r in rq:
            r.flags |= F_BUFFER_ON_DEVICE | F_REPLY_ON_DEVICE
        self._check(self.lib.sp_render_shards(self.h, rq, rp))
        return list(rp)

    # ---- taps
    def decode(self, fmt, buf, first: int = 0, count: int | None = None) -> np.ndarray:
        f = format_id(fmt)
        b = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf).view(np.uint8).ravel()
        if count is None:
            count = len(b) // self.lib.sp_sample_width(f) - first
        out = np.empty((count, 2), np.float32)
        self._check(self.lib.sp_decode(self.h, f, _vp(b), len(b), first, count, _vp(out)))
        return out

    # ---- the path
    def make_request(self, buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode=False,
                     waterfall=False, flags=0, byte_length=None, shard=None):
        """Returns (Request, keepalive list).  `buf` is bytes / numpy (host) or an int device pointer."""
        keep = []
        if isinstance(buf, (int, np.integer)):
            bptr = C.c_void_p(int(buf))
            assert byte_length is not None
            flags |= F_BUFFER_ON_DEVICE
        else:
            b = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf).view(np.uint8).ravel()
            keep.append(b)
            bptr = _vp(b)
            byte_length = len(b) if byte_length is None else byte_length
        w = np.ascontiguousarray(windowc, dtype=np.float64)
        cm = np.ascontiguousarray(cmap, dtype=np.uint8).reshape(-1, 3)
        keep += [w, cm]
        rq = Request(bptr, int(byte_length), format_id(fmt), int(n), int(width), float(block_norm), float(gain),
                     float(range_), _vp(w), _vp(cm), len(cm), int(bool(channel_mode)), int(bool(waterfall)),
                     int(flags), 0, 0, 0, 0)
        if shard is not None:
            rq.total_byte_length = int(shard["total_byte_length"])
            rq.total_width = int(shard["total_width"])
            rq.frame_first = int(shard["frame_first"])
            rq.buffer_first_sample = int(shard["buffer_first_sample"])
        return rq, keep

    def render(self, buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode=False,
               waterfall=False, image=True, shard=None, byte_length=None, out_image: np.ndarray | None = None):
        """Host-buffer render: one worker message in, one reply dict out."""
        rq, keep = self.make_request(buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode,
                                     waterfall, 0 if image else F_NO_IMAGE, byte_length, shard)
        width = int(width)
        n = int(n)
        img = None
        if image:
            img = out_image if out_image is not None else np.empty(4 * width * n, np.uint8)
        gmin = np.empty(width, np.uint8); gmax = np.empty(width, np.uint8); gamp = np.empty(width, np.uint8)
        cb = np.zeros(CB_HIST_SIZE, np.uint64); ch = np.zeros(rq.cmap_len, np.uint64)
        rp = Reply(_vp(img), _vp(gmin), _vp(gmax), _vp(gamp), _vp(cb), _vp(ch), 0.0, 0.0, 0.0, 0, None)
        self._check(self.lib.sp_render(self.h, C.byref(rq), C.byref(rp)))
        if img is not None:
            img = img[: 4 * width * n].reshape((width, n, 4) if waterfall else (n, width, 4))
        return dict(image=img, gauge_mins=gmin, gauge_maxs=gmax, gauge_amps=gamp, cB_hist=cb, c_hist=ch,
                    dBfs_min=rp.dBfs_min, dBfs_max=rp.dBfs_max, device_ms=rp.device_ms,
                    kernel_launches=rp.kernel_launches)

    def render_async(self, buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode=False,
                     waterfall=False, shard=None, byte_length=None, out_image: np.ndarray | None = None):
        """sp_render_async: enqueue a host-buffer message and return a handle for `wait`; up to two may be in flight."""
        rq, keep = self.make_request(buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode,
                                     waterfall, 0, byte_length, shard)
        width, n = int(width), int(n)
        img = out_image if out_image is not None else np.empty(4 * width * n, np.uint8)
        gmin = np.empty(width, np.uint8); gmax = np.empty(width, np.uint8); gamp = np.empty(width, np.uint8)
        cb = np.zeros(CB_HIST_SIZE, np.uint64); ch = np.zeros(rq.cmap_len, np.uint64)
        rp = Reply(_vp(img), _vp(gmin), _vp(gmax), _vp(gamp), _vp(cb), _vp(ch), 0.0, 0.0, 0.0, 0, None)
        ticket = C.c_int(-1)
        self._check(self.lib.sp_render_async(self.h, C.byref(rq), C.byref(rp), C.byref(ticket)))
        return dict(ticket=ticket.value, rq=rq, rp=rp, keep=keep, img=img, gmin=gmin, gmax=gmax, gamp=gamp, cb=cb, ch=ch,
                    shape=(width, n, 4) if waterfall else (n, width, 4))

    def wait(self, handle) -> dict:
        """sp_render_wait: block until the message of `handle` is complete; returns the reply dict of `render`."""
        self._check(self.lib.sp_render_wait(self.h, int(handle["ticket"])))
        rp, shape = handle["rp"], handle["shape"]
        img = handle["img"][: shape[0] * shape[1] * 4].reshape(shape)
        return dict(image=img, gauge_mins=handle["gmin"], gauge_maxs=handle["gmax"], gauge_amps=handle["gamp"],
                    cB_hist=handle["cb"], c_hist=handle["ch"], dBfs_min=rp.dBfs_min, dBfs_max=rp.dBfs_max,
                    device_ms=rp.device_ms, kernel_launches=rp.kernel_launches)

    def render_zooms(self, buf, fmt, n, widths, windowc, block_norm, gain, range_, cmap, channel_mode=False,
                     waterfall=False):
        """Several zoom levels of one capture in one pass over its bytes (sp_render_zooms): the capture is
        uploaded once; level i is the message with width = widths[i] and its own stride.  -> list of reply dicts."""
        rq, keep = self.make_request(buf, fmt, n, widths[0], windowc, block_norm, gain, range_, cmap, channel_mode, waterfall)
        n = int(n)
        nl = len(widths)
        wl = (C.c_int64 * nl)(*[int(w) for w in widths])
        rps = (Reply * nl)()
        outs = []
        for i, width in enumerate(int(w) for w in widths):
            img = np.empty(4 * width * n, np.uint8)
            gmin = np.empty(width, np.uint8); gmax = np.empty(width, np.uint8); gamp = np.empty(width, np.uint8)
            cb = np.zeros(CB_HIST_SIZE, np.uint64); ch = np.zeros(rq.cmap_len, np.uint64)
            rps[i] = Reply(_vp(img), _vp(gmin), _vp(gmax), _vp(gamp), _vp(cb), _vp(ch), 0.0, 0.0, 0.0, 0, None)
            outs.append(dict(image=img, gauge_mins=gmin, gauge_maxs=gmax, gauge_amps=gamp, cB_hist=cb, c_hist=ch))
        self._check(self.lib.sp_render_zooms(self.h, C.byref(rq), nl, wl, rps))
        for i, width in enumerate(int(w) for w in widths):
            o = outs[i]
            o["image"] = o["image"].reshape((width, n, 4) if waterfall else (n, width, 4))
            o.update(dBfs_min=rps[i].dBfs_min, dBfs_max=rps[i].dBfs_max, device_ms=rps[i].device_ms,
                     kernel_launches=rps[i].kernel_launches)
        return outs

    def render_db(self, buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode=False) -> np.ndarray:
        rq, keep = self.make_request(buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode)
        out = np.empty((int(width), int(n)), np.float32)
        self._check(self.lib.sp_render_db(self.h, C.byref(rq), _vp(out)))
        return out
