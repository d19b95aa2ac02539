"""ctypes front-end of the CPU oracle (oracle/spectro_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of spectro_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product path under spectroplot-js_b200/.

PARITY PIN: the reference has no tests / golden vectors and the image has no JavaScript engine; the pin
is the reference's own source executed by oracle/jsmini.py (tools/make_ref_golden.py ->
tests/golden/ref_js/), which this oracle reproduces exactly (tests/test_reference_js.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libspectro_oracle.so")

FORMATS = ["CU4", "CS4", "CU8", "CS8", "CU12", "CS12", "CU16", "CS16",
           "CU32", "CS32", "CU64", "CS64", "CF32", "CF64"]
SAMPLE_WIDTH = [1, 1, 2, 2, 3, 3, 4, 4, 8, 8, 16, 16, 8, 16]
WINDOWS = ["rectangular", "bartlett", "hamming", "hann", "blackman", "blackmanHarris"]
CB_HIST = 1000


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (Makefile in this directory)."""
    src = os.path.join(_HERE, "spectro_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libspectro_oracle.so"])
    return _SO


class _Req(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("byte_length", C.c_uint64),
                ("format", C.c_int32), ("n", C.c_int32), ("width", C.c_int64),
                ("block_norm", C.c_double), ("gain", C.c_double), ("range", C.c_double),
                ("windowc", C.c_void_p), ("cmap_rgb", C.c_void_p),
                ("cmap_len", C.c_int32), ("channel_mode", C.c_int32),
                ("waterfall", C.c_int32), ("pad_", C.c_int32)]


class _Rep(C.Structure):
    _fields_ = [("image", C.c_void_p), ("gauge_mins", C.c_void_p), ("gauge_maxs", C.c_void_p),
                ("gauge_amps", C.c_void_p), ("cB_hist", C.c_void_p), ("c_hist", C.c_void_p),
                ("dBfs_min", C.c_double), ("dBfs_max", C.c_double),
                ("db", C.c_void_p), ("gray", C.c_void_p), ("cbk", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.spo_window.restype = C.c_double
        _lib.spo_format_name.restype = C.c_char_p
    return _lib


def fmt_id(fmt) -> int:
    if isinstance(fmt, str):
        return lib().spo_format_from_name(fmt.encode())
    return int(fmt)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@dataclass
class Result:
    image: np.ndarray | None
    gauge_mins: np.ndarray
    gauge_maxs: np.ndarray
    gauge_amps: np.ndarray
    cB_hist: np.ndarray
    c_hist: np.ndarray
    dBfs_min: float
    dBfs_max: float
    db: np.ndarray | None = None
    gray: np.ndarray | None = None
    cbk: np.ndarray | None = None


def window(kind, n: int):
    """-> (window[n] float64, weight).  reference lib/windows.js:14-88"""
    k = WINDOWS.index(kind) if isinstance(kind, str) else int(kind)
    w = np.empty(n, dtype=np.float64)
    weight = lib().spo_window(k, n, _ptr(w))
    return w, float(weight)


def decode(fmt, buf: bytes | np.ndarray, first: int = 0, count: int | None = None) -> np.ndarray:
    """-> float64 [count, 2] (I, Q).  reference lib/samples.js:313-400"""
    f = fmt_id(fmt)
    b = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).ravel()
    if count is None:
        count = len(b) // SAMPLE_WIDTH[f] - first
    out = np.empty((count, 2), dtype=np.float64)
    rc = lib().spo_decode(f, _ptr(b), C.c_uint64(len(b)), C.c_int64(first), C.c_int64(count), _ptr(out))
    if rc:
        raise ValueError(f"spo_decode rc={rc}")
    return out


def fft(re: np.ndarray, im: np.ndarray):
    re = np.array(re, dtype=np.float64); im = np.array(im, dtype=np.float64)
    rc = lib().spo_fft(len(re), _ptr(re), _ptr(im))
    if rc:
        raise ValueError("Length is not a power of 2")
    return re, im


def splitreal(re, im):
    re = np.array(re, dtype=np.float64); im = np.array(im, dtype=np.float64)
    lib().spo_splitreal(len(re), _ptr(re), _ptr(im))
    return re, im


def render(buf, fmt, n, width, windowc, block_norm, gain, range_, cmap, channel_mode=False,
           waterfall=False, taps=False, image=True, workers: int = 0) -> Result:
    """One worker message (reference lib/worker.js:23-156).  workers > 0 runs the
    caller-side fan-out of lib/spectroplot.js:1206-1238 on that many threads instead."""
    f = fmt_id(fmt)
    b = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).ravel()
    windowc = np.ascontiguousarray(windowc, dtype=np.float64)
    cmap = np.ascontiguousarray(cmap, dtype=np.uint8).reshape(-1, 3)
    width = int(width)
    rq = _Req(_ptr(b), len(b), f, int(n), width, float(block_norm), float(gain), float(range_),
              _ptr(windowc), _ptr(cmap), len(cmap), int(bool(channel_mode)), int(bool(waterfall)), 0)
    w = max(width, 0)
    img = np.zeros(4 * w * n, dtype=np.uint8) if image else None
    gmin = np.zeros(w, np.uint8); gmax = np.zeros(w, np.uint8); gamp = np.zeros(w, np.uint8)
    cb = np.zeros(CB_HIST, np.uint64); ch = np.zeros(len(cmap), np.uint64)
    db = np.empty((w, n), np.float64) if taps else None
    gray = np.empty((w, n), np.uint16) if taps else None
    cbk = np.empty((w, n), np.int32) if taps else None
    rp = _Rep(_ptr(img), _ptr(gmin), _ptr(gmax), _ptr(gamp), _ptr(cb), _ptr(ch), 0.0, 0.0,
              _ptr(db), _ptr(gray), _ptr(cbk))
    if workers > 0:
        rc = lib().spo_render_fanout(C.byref(rq), C.byref(rp), int(workers))
    else:
        rc = lib().spo_render(C.byref(rq), C.byref(rp))
    if rc:
        raise ValueError(f"oracle render failed rc={rc}")
    if img is not None:
        img = img.reshape((w, n, 4) if waterfall else (n, w, 4))
    return Result(img, gmin, gmax, gamp, cb, ch, rp.dBfs_min, rp.dBfs_max, db, gray, cbk)


def cmap_sox(stops=256):
    out = np.empty((stops, 3), np.uint8); lib().spo_cmap_sox(stops, _ptr(out)); return out


def cmap_naive(kind, stops=256):
    k = ["naive", "grayscale", "roentgen", "phosphor"].index(kind) if isinstance(kind, str) else kind
    out = np.empty((stops, 3), np.uint8); lib().spo_cmap_naive(k, stops, _ptr(out)); return out


def synth(fmt, first: int, count: int, total: int, seed: int) -> np.ndarray:
    """Deterministic synthetic capture bytes (uint8 array)."""
    f = fmt_id(fmt)
    out = np.empty(count * SAMPLE_WIDTH[f], np.uint8)
    rc = lib().spo_synth_fill(_ptr(out), f, C.c_uint64(first), C.c_uint64(count), C.c_uint64(total), C.c_uint64(seed))
    if rc:
        raise ValueError(f"spo_synth_fill rc={rc}")
    return out


def synth_lut() -> np.ndarray:
    out = np.empty(4096, np.int16); lib().spo_synth_lut(_ptr(out)); return out
