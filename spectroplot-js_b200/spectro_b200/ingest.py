"""Capture ingestion — the step before the render path.

* `load_capture(path)`: file -> the `filedata` object `Spectroplot.setData` takes (reference lib/spectroplot.js:480-511:
  `{fileBuffer, name, size, type}`), read straight into page-locked host memory when a GPU engine is available so that the
  pipelined host path (H2D / render / D2H overlap, csrc/sp_engine.cu::render_pipelined) runs at PCIe speed.
* `decode_wav(data)`: the audio front end.  The reference hands WAV / FLAC / MP3 ... to the browser's
  `decodeAudioData` and then interleaves the first two channels into CF32 (lib/samples.js:141-148, 260-302).  There is no
  browser codec here; PCM WAV / BWF (8 / 16 / 24 / 32-bit integer, 32 / 64-bit float, WAVE_FORMAT_EXTENSIBLE) is decoded
  directly with WebAudio's scaling, mono is duplicated and extra channels are dropped exactly like `interleaved()`.
  Compressed formats raise NotImplementedError (decode them to WAV / CF32 first).
"""
from __future__ import annotations

import os
import struct

import numpy as np

AUDIO_PCM = {"WAV", "BWF"}


def decode_wav(data: bytes):
    """-> (interleaved float32 array [2 * frames], sampleRate, channels in the file)."""
    if len(data) < 12 or data[:4] not in (b"RIFF", b"RF64") or data[8:12] != b"WAVE":
        raise ValueError("decodeAudioData error: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, rate, _bps, _align, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == 0xFFFE and len(body) >= 26:                  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the tag
                tag = struct.unpack_from("<H", body, 24)[0]
            fmt = (tag, ch, rate, bits)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError("decodeAudioData error: missing fmt or data chunk")
    tag, ch, rate, bits = fmt
    if ch < 1:
        raise ValueError("AudioBuffer wrong numberOfChannels (%d)" % ch)
    if tag == 1:                                                    # integer PCM
        if bits == 8:
            x = (np.frombuffer(pcm, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(pcm[:len(pcm) // 2 * 2], "<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(pcm[:len(pcm) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = (v / 8388608.0).astype(np.float32)
        elif bits == 32:
            x = (np.frombuffer(pcm[:len(pcm) // 4 * 4], "<i4") / 2147483648.0).astype(np.float32)
        else:
            raise NotImplementedError("PCM WAV with %d bits per sample" % bits)
    elif tag == 3:                                                  # IEEE float
        x = np.frombuffer(pcm[:len(pcm) // (bits // 8) * (bits // 8)], "<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise NotImplementedError("compressed WAV (format tag %d) needs a codec; decode to PCM first" % tag)
    frames = x.size // ch
    x = x[:frames * ch].reshape(frames, ch)
    out = np.empty((frames, 2), np.float32)                          # lib/samples.js:276-302 interleaved()
    out[:, 0] = x[:, 0]
    out[:, 1] = x[:, 1] if ch > 1 else x[:, 0]                       # mono: channel duplicated; > 2 channels: first two
    return out.reshape(-1), int(rate), int(ch)


def load_capture(path: str, pinned: bool = True) -> dict:
    """File -> filedata for Spectroplot.setData.  With `pinned` (and the CUDA library present) the bytes are read into
    page-locked memory; the returned dict keeps the owner alive under 'pinned'."""
    size = os.path.getsize(path)
    name = os.path.basename(path)
    owner = None
    if pinned and size:
        try:
            from . import _lib
            owner = _lib.PinnedBuffer(size)
            buf = owner.array
        except Exception:                                           # no GPU library on this host: plain memory
            owner, buf = None, np.empty(size, np.uint8)
    else:
        buf = np.empty(size, np.uint8)
    with open(path, "rb") as f:
        got = f.readinto(memoryview(buf)) if size else 0
    assert got == size
    return {"fileBuffer": buf, "name": name, "size": size, "type": "", "pinned": owner}
