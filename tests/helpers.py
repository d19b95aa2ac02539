"""Shared helpers of the parity tests: the tolerances of BASELINE.json's north_star."""
import numpy as np

DB_TOL = 0.01            # dB, wherever the bin is above the floor
DB_FLOOR_REF = -60.0     # "-120 dBFS" conventional 20*log10 == -60 on the reference's 10*log10|X| scale
PIXEL_FRAC = 1e-3        # <= 0.1 % of pixels may differ, by one colour step (quantisation ties)


def injective_cmap(n):
    i = np.arange(n)
    return np.stack([i & 255, (i * 7 + 3) & 255, ((i >> 8) * 16 + (i * 37 & 15)) & 255], 1).astype(np.uint8)


def cmap_index_image(image, cmap):
    """RGBA image -> colour index image for an injective cmap (-1 where the pixel is not in the cmap)."""
    key = lambda a: (a[..., 0].astype(np.int64) << 16) | (a[..., 1].astype(np.int64) << 8) | a[..., 2].astype(np.int64)
    table = {int(k): i for i, k in enumerate(key(cmap))}
    assert len(table) == len(cmap), "cmap not injective"
    k = key(image)
    out = np.full(k.shape, -1, np.int64)
    for kk, i in table.items():
        out[k == kk] = i
    return out


def bin_to_row(n):
    i = np.arange(n)
    return np.where(i <= n // 2, n // 2 - i, n // 2 + n - i)      # lib/worker.js:90


def gray_from_image(image, cmap, n, width, waterfall=False):
    """-> [width][n] colour indices in FFT bin order (the oracle's `gray` tap layout)."""
    idx = cmap_index_image(image, cmap)
    y = bin_to_row(n)
    if waterfall:
        return idx[(width - 1 - np.arange(width))[:, None], (n - 1 - y)[None, :]]
    return idx[y[None, :], np.arange(width)[:, None]]


def check_parity(gpu, ora, cmap, n, width, waterfall=False, gpu_db=None, label=""):
    """gpu: dict from Engine.render; ora: oracle Result with taps.  Asserts the north_star bars."""
    img = gpu["image"]
    assert img.shape == ora.image.shape, label
    assert (img[..., 3] == 255).all(), label + ": alpha"
    g = gray_from_image(img, cmap, n, width, waterfall)
    assert (g >= 0).all(), label + ": pixel colour not in cmap"
    diff = g - ora.gray.astype(np.int64)
    nbad = int((diff != 0).sum())
    assert np.abs(diff).max() <= 1, f"{label}: colour index off by more than one step (max {np.abs(diff).max()})"
    assert nbad <= max(1, int(PIXEL_FRAC * diff.size)), f"{label}: {nbad}/{diff.size} pixels differ"
    # histograms: identical up to those ties
    assert int(gpu["c_hist"].sum()) == int(ora.c_hist.sum()) == width * n, label
    assert int(np.abs(gpu["c_hist"].astype(np.int64) - ora.c_hist.astype(np.int64)).sum()) <= 2 * nbad, label + ": c_hist"
    cb_d = np.abs(gpu["cB_hist"].astype(np.int64) - ora.cB_hist.astype(np.int64)).sum()
    assert cb_d <= 2 * max(2, int(PIXEL_FRAC * diff.size)), f"{label}: cB_hist differs by {cb_d}"
    assert abs(int(gpu["cB_hist"].sum()) - int(ora.cB_hist.sum())) <= max(2, int(PIXEL_FRAC * diff.size)), label
    # gauges: +-1 count (fp32 min/max feeding a rounding)
    for k in ("gauge_mins", "gauge_maxs", "gauge_amps"):
        d = np.abs(gpu[k].astype(int) - getattr(ora, k).astype(int))
        assert d.max() <= 1, f"{label}: {k} max diff {d.max()}"
    for k in ("dBfs_min", "dBfs_max"):
        a, b = gpu[k], getattr(ora, k)
        if np.isfinite(b) and b > DB_FLOOR_REF:
            assert abs(a - b) <= DB_TOL, f"{label}: {k} {a} vs {b}"
        elif np.isfinite(b):
            assert abs(a - b) <= 1.0, f"{label}: {k} {a} vs {b}"
        else:
            assert a == b, f"{label}: {k} {a} vs {b}"
    if gpu_db is not None:
        m = np.isfinite(ora.db) & (ora.db > DB_FLOOR_REF)
        err = np.abs(gpu_db.astype(np.float64) - ora.db)[m]
        assert err.size == 0 or err.max() <= DB_TOL, f"{label}: dB error {err.max()} above the floor"
    return nbad
