"""Second, independent float64 restatement of the render worker in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/spectro_oracle.c).  The parity pin is described there
(reference source run by oracle/jsmini.py -> tests/golden/ref_js/).  This one is vectorised and uses numpy's pocketfft instead of the reference's
radix-2 transform, so an error in either restatement of lib/fft_nayuki.js / lib/worker.js
shows up as a disagreement between the two.  It also generates tests/golden/*.npz
(tools/make_golden.py).

Citations are reference file:line.
"""
from __future__ import annotations

import numpy as np

FORMATS = ["CU4", "CS4", "CU8", "CS8", "CU12", "CS12", "CU16", "CS16",
           "CU32", "CS32", "CU64", "CS64", "CF32", "CF64"]
ALIASES = {"DATA": "CU8", "COMPLEX16U": "CU8", "COMPLEX16S": "CS8", "CFILE": "CF32", "COMPLEX": "CF32"}
SAMPLE_WIDTH = dict(zip(FORMATS, [1, 1, 2, 2, 3, 3, 4, 4, 8, 8, 16, 16, 8, 16]))


def canon(fmt: str) -> str:
    f = fmt.upper()                                   # lib/samples.js:22
    f = ALIASES.get(f, f)
    return f if f in FORMATS else "CU8"               # lib/samples.js:149-155


def toint32(v):
    """JS `~~v` on an array of doubles."""
    v = np.asarray(v, dtype=np.float64)
    t = np.where(np.isfinite(v), np.trunc(v), 0.0)
    m = np.mod(t, 4294967296.0)
    m = np.where(m >= 2147483648.0, m - 4294967296.0, m)
    return m.astype(np.int64)


def u8clamped(v):
    """Store into a Uint8ClampedArray: clamp, round half to even, NaN -> 0."""
    v = np.asarray(v, dtype=np.float64)
    v = np.where(np.isnan(v), 0.0, v)
    return np.rint(np.clip(v, 0.0, 255.0)).astype(np.uint8)


def decode_all(fmt: str, buf: bytes) -> np.ndarray:
    """All whole samples of `buf` as float64 [count, 2].  lib/samples.js:30-139,313-400"""
    f = canon(fmt)
    raw = np.frombuffer(buf, dtype=np.uint8)
    sw = SAMPLE_WIDTH[f]
    cnt = len(raw) // sw
    raw = raw[: cnt * sw]
    if f in ("CU4", "CS4"):
        hi = (raw >> 4).astype(np.int64); lo = (raw & 15).astype(np.int64)
        if f == "CU4":
            return np.stack([(hi - 7.5) * (1.0 / 7.5), (lo - 7.5) * (1.0 / 7.5)], 1)
        hi = np.where(hi >= 8, hi - 16, hi); lo = np.where(lo >= 8, lo - 16, lo)
        return np.stack([hi * (1.0 / 8.0), lo * (1.0 / 8.0)], 1)
    if f in ("CU12", "CS12"):
        g = raw.reshape(cnt, 3).astype(np.int64)
        i = ((g[:, 1] & 15) << 8) | g[:, 0]
        q = (g[:, 2] << 4) | (g[:, 1] >> 4)
        if f == "CU12":
            return np.stack([(i - 2047.5) * (1.0 / 2047.5), (q - 2047.5) * (1.0 / 2047.5)], 1)
        i = np.where(i >= 2048, i - 4096, i); q = np.where(q >= 2048, q - 4096, q)
        return np.stack([i * (1.0 / 2048.0), q * (1.0 / 2048.0)], 1)
    if f in ("CU64", "CS64"):
        w = raw.view("<u4").reshape(cnt, 4).astype(np.float64)
        hi_i, hi_q = w[:, 1], w[:, 3]
        if f == "CS64":
            hi_i = np.where(hi_i >= 2 ** 31, hi_i - 2 ** 32, hi_i)
            hi_q = np.where(hi_q >= 2 ** 31, hi_q - 2 ** 32, hi_q)
        i = hi_i / 2 ** 31 + w[:, 0] / 2 ** 64
        q = hi_q / 2 ** 31 + w[:, 2] / 2 ** 64
        if f == "CU64":
            i = i - 1.0; q = q - 1.0
        return np.stack([i, q], 1)
    dt, bias, scale = {
        "CU8": ("u1", 127.5, 1.0 / 127.5), "CS8": ("i1", 0, 1.0 / 128.0),
        "CU16": ("<u2", 32767.5, 1.0 / 32768.0), "CS16": ("<i2", 0, 1.0 / 32768.0),
        "CU32": ("<u4", 2147483647.5, 1.0 / 2147483648.0), "CS32": ("<i4", 0, 1.0 / 2147483648.0),
        "CF32": ("<f4", 0, 1.0), "CF64": ("<f8", 0, 1.0)}[f]
    v = raw.view(dt).astype(np.float64).reshape(cnt, 2)
    return (v - bias) * scale


def window(kind: str, n: int):
    """lib/windows.js:14-88 -> (window, weight accumulated in index order)"""
    i = np.arange(n, dtype=np.float64)
    pi = np.pi
    if kind == "rectangular":
        w = np.ones(n)
    elif kind == "bartlett":
        w = 1.0 - np.abs((i - 0.5 * (n - 1)) / (0.5 * (n - 1)))
    elif kind == "hamming":
        w = 0.54 - 0.46 * np.cos(2.0 * pi * i / (n - 1))
    elif kind == "hann":
        w = 0.5 * (1.0 - np.cos(2.0 * pi * i / (n - 1)))
    elif kind == "blackman":
        w = 0.42 - (0.5 * np.cos((2.0 * pi * i) / (n - 1))) + (0.08 * np.cos((4.0 * pi * i) / (n - 1)))
    elif kind == "blackmanHarris":
        w = (0.35875 - (0.48829 * np.cos((2.0 * pi * i) / (n - 1)))
             + (0.14128 * np.cos((4.0 * pi * i) / (n - 1))) - (0.01168 * np.cos((6.0 * pi * i) / (n - 1))))
    else:
        raise KeyError(kind)
    weight = 0.0
    for x in w:                    # index-order accumulation like the reference
        weight += float(x)
    return w, weight


def splitreal(X: np.ndarray) -> np.ndarray:
    """lib/fft_nayuki.js:103-119 on rows of complex X[..., n]."""
    n = X.shape[-1]
    re = X.real.copy(); im = X.imag.copy()
    im[..., 0] = 0
    re[..., n // 2] = 0            # real[n/2] = imag[0] (already zeroed)
    im[..., n // 2] = 0
    i = np.arange(1, n // 2)
    a_re, a_im = X.real[..., i], X.imag[..., i]
    b_re, b_im = X.real[..., n - i], X.imag[..., n - i]
    re[..., i] = 0.5 * (a_re + b_re)
    im[..., i] = 0.5 * (a_im - b_im)
    re[..., n - i] = 0.5 * (a_im + b_im)
    im[..., n - i] = 0.5 * (-a_re + b_re)
    return re + 1j * im


def render(buf: bytes, fmt: str, n: int, width: int, windowc, block_norm, gain, range_, cmap,
           channel_mode=False, waterfall=False):
    """lib/worker.js:23-156 for well-formed messages (no out-of-range reads)."""
    f = canon(fmt)
    iq = decode_all(f, buf)
    z = iq[:, 0] + 1j * iq[:, 1]
    sample_count = len(buf) / SAMPLE_WIDTH[f]                      # lib/samples.js:167
    stride = (sample_count - n) / (width - 1)                      # lib/worker.js:50
    x = np.arange(width, dtype=np.float64)
    p0 = toint32(0.5 + stride * x)                                 # :72
    idx = p0[:, None] + np.arange(n)[None, :]
    if idx.max() >= len(z) or idx.min() < 0:
        raise ValueError("np_restatement only handles in-range messages")
    frames = z[idx] * np.asarray(windowc, dtype=np.float64)[None, :]
    X = np.fft.fft(frames, axis=1)                                 # forward, unscaled (:54-86)
    if channel_mode:
        X = splitreal(X)
    abs2 = X.real * X.real + X.imag * X.imag                       # :92
    with np.errstate(divide="ignore"):
        dbfs = 5 * np.log10(abs2) + 10 * np.log10(block_norm) + gain   # :93
    d0 = dbfs - gain
    cmap = np.asarray(cmap, dtype=np.uint8).reshape(-1, 3)
    clen = len(cmap); cmax = clen - 1
    color_norm = clen / -range_                                    # :39
    min_i = np.minimum(0.0, np.nanmin(d0, axis=1))                 # :82,102
    max_i = np.maximum(-200.0, np.nanmax(d0, axis=1))              # :83,103
    cb = toint32(0.5 + d0 * -10)                                   # :105
    cb = np.where(cb >= 1000, 999, cb)                             # :106
    cB_hist = np.bincount(cb[cb >= 0].ravel(), minlength=1000).astype(np.uint64)
    grayu = cmax - dbfs * color_norm                               # :111
    gray = toint32(0.5 + np.clip(np.where(np.isnan(grayu), 0, grayu), 0, cmax))   # :112
    c_hist = np.bincount(gray.ravel(), minlength=clen).astype(np.uint64)
    i = np.arange(n)
    y = np.where(i <= n // 2, n // 2 - i, n // 2 + n - i)          # :90
    rgba = np.concatenate([cmap[gray], np.full(gray.shape + (1,), 255, np.uint8)], axis=2)  # [width][n][4]
    if waterfall:                                                  # :116
        img = np.zeros((width, n, 4), np.uint8)
        img[(width - 1 - np.arange(width))[:, None], (n - 1 - y)[None, :]] = rgba
    else:                                                          # :117
        img = np.zeros((n, width, 4), np.uint8)
        img[y[None, :], np.arange(width)[:, None]] = rgba
    gmin = u8clamped(0.5 + (range_ + min_i) * 256 / range_)        # :128
    gmax = u8clamped(0.5 + (range_ + max_i) * 256 / range_)        # :129
    mid = iq[p0 + n // 2]                                          # :131-133
    with np.errstate(divide="ignore"):
        amp = 5 * np.log10(mid[:, 0] ** 2 + mid[:, 1] ** 2) + gain  # :135
    gamp = u8clamped(0.5 + (range_ + amp) * 256 / range_)          # :136
    return dict(image=img, gauge_mins=gmin, gauge_maxs=gmax, gauge_amps=gamp, cB_hist=cB_hist,
                c_hist=c_hist, dBfs_min=float(min(0.0, min_i.min())), dBfs_max=float(max(-200.0, max_i.max())),
                db=d0, gray=gray.astype(np.uint16), p0=p0)
