"""Image egress — the step after the render path: the worker's RGBA `imageData` to files.

The reference paints the reply into a canvas (lib/spectroplot.js:1240-1268); a headless host writes it out instead.
`write_png` is a dependency-free PNG encoder (8-bit RGBA, zlib), `write_tiles` cuts a long spectrogram into fixed-width
PNG tiles (a 25 600-frame C2 image is 25 600 px wide), `write_pyramid` renders and writes the zoom levels of one capture
(Engine.render_zooms: one upload, one render per level with its own stride), `write_reply` stores the
histograms, gauges and dB range next to the picture.
"""
from __future__ import annotations

import json
import os
import struct
import zlib

import numpy as np


def _chunk(tag: bytes, body: bytes) -> bytes:
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)


def png_bytes(image: np.ndarray, level: int = 3) -> bytes:
    """image: uint8 [height][width][4] RGBA (the reply's imageData reshaped) -> PNG file bytes."""
    img = np.ascontiguousarray(image, np.uint8)
    if img.ndim != 3 or img.shape[2] != 4:
        raise ValueError("expected an RGBA image [height][width][4]")
    h, w = img.shape[:2]
    raw = np.empty((h, 1 + 4 * w), np.uint8)
    raw[:, 0] = 0                                                   # filter type 0 (None) on every scanline
    raw[:, 1:] = img.reshape(h, 4 * w)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
            + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b""))


def write_png(path: str, image: np.ndarray, level: int = 3) -> str:
    with open(path, "wb") as f:
        f.write(png_bytes(image, level))
    return path


def read_png_rgba(data: bytes) -> np.ndarray:
    """Decoder for the files written above (8-bit RGBA, filter 0) — used by the tests and by tools."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack_from(">I", data, pos)[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack_from(">I", data, pos + 8 + n)[0] == zlib.crc32(tag + body) & 0xFFFFFFFF, "PNG chunk CRC"
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack_from(">IIBB", body, 0)
            assert (depth, ctype) == (8, 6)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 4).copy()


def write_tiles(directory: str, image: np.ndarray, tile_width: int = 4096, prefix: str = "tile") -> list:
    """Spectrogram [n][width][4] -> PNG tiles of `tile_width` columns, named by their first frame."""
    os.makedirs(directory, exist_ok=True)
    out = []
    for x0 in range(0, image.shape[1], tile_width):
        out.append(write_png(os.path.join(directory, "%s_%08d.png" % (prefix, x0)), image[:, x0:x0 + tile_width]))
    return out


def write_reply(prefix: str, reply: dict) -> dict:
    """Everything a reply (lib/worker.js:140-149) carries besides the picture, as JSON next to `<prefix>.png`."""
    image = np.asarray(reply["image"])
    write_png(prefix + ".png", image)
    meta = {"width": int(image.shape[1]), "height": int(image.shape[0]), "dBfs_min": float(reply["dBfs_min"]),
            "dBfs_max": float(reply["dBfs_max"]), "cB_hist": [int(v) for v in reply["cB_hist"]],
            "c_hist": [int(v) for v in reply["c_hist"]]}
    for k in ("gauge_mins", "gauge_maxs", "gauge_amps"):
        meta[k] = [int(v) for v in reply[k]]
    with open(prefix + ".json", "w") as f:
        json.dump(meta, f)
    return meta


def write_pyramid(directory: str, engine, buf, fmt, n, base_width, zooms, windowc, block_norm, gain, range_, cmap,
                  tile_width: int = 4096) -> list:
    """Zoom pyramid of one capture on disk: level z has width z * base_width and its own stride (one upload, C3 shape)."""
    widths = [int(z * base_width) for z in zooms]
    replies = engine.render_zooms(buf, fmt, n, widths, windowc, block_norm, gain, range_, cmap)
    out = []
    for z, r in zip(zooms, replies):
        d = os.path.join(directory, "zoom_x%g" % z)
        out.append({"zoom": z, "tiles": write_tiles(d, np.asarray(r["image"]), tile_width), "meta": write_reply(os.path.join(d, "reply"), r)})
    return out
