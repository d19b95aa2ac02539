#!/usr/bin/env python
"""Is BASELINE's 0.01 dB bar reachable in fp32 in the -120 .. -100 dBFS band?  (VERDICT r1, weak #1.)

A numpy fp32 model of the N = 4096 transform exactly as render_r64_kernel factors it (64 x 64; each 64-point transform
8 x 8 with fp32 constants; one multiplication per element by W_4096^{t k0} between the passes) on the C2 signal (two tones
at -6 / -20 dBFS + noise at -50 dBFS, Blackman-Harris), compared bin by bin with the float64 transform, for two ways of
getting the 63 inter-pass twiddles of a thread:
  table    every W^{t k0} rounded ONCE from double (what a 32 KB shared-memory table would hold);
  product  14 table entries per thread (W^{t j}, j = 1..7, and W^{8 t i}, i = 1..7), the other 49 formed as fp32 products
           (what the kernel does: the table does not fit next to the exchange buffers).
Prints max / rms dB error above -100 dBFS and in the -120 .. -100 dBFS band for both.  CPU only; ~1 min."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

N = 4096
f32, c64 = np.float32, np.complex64


def w_exact(num, den):
    a = 2.0 * np.pi * (np.asarray(num) % den) / den
    return (np.cos(a) - 1j * np.sin(a))


def dft8(v):                       # v: [..., 8] complex64, radix-2 DIT in fp32 with exact-rounded W8 constants
    w8 = w_exact(np.arange(8), 8).astype(c64)
    e = v[..., 0::2]; o = v[..., 1::2]
    def dft4(u):
        a, b, c, d = u[..., 0], u[..., 1], u[..., 2], u[..., 3]
        t0, t1, t2, t3 = a + c, a - c, b + d, (b - d) * c64(-1j)
        return np.stack([t0 + t2, t1 + t3, t0 - t2, t1 - t3], -1)
    E, Od = dft4(e), dft4(o)
    Od = Od * w8[:4]
    return np.concatenate([E + Od, E - Od], -1).astype(c64)


def dft64(v):                      # [..., 64]: n = 8 n1 + n0, k = k0 + 8 k1
    v = v.reshape(v.shape[:-1] + (8, 8))                       # [n1][n0]
    a = dft8(np.swapaxes(v, -1, -2))                           # over n1 -> [n0][k0]
    tw = w_exact(np.outer(np.arange(8), np.arange(8)), 64).astype(c64)   # W64^{n0 k0}
    a = (a * tw).astype(c64)
    b = dft8(np.swapaxes(a, -1, -2))                           # over n0 -> [k0][k1]
    return np.swapaxes(b, -1, -2).reshape(v.shape[:-2] + (64,))   # index k0 + 8 k1


def fft4096_fp32(x, mode):
    # x: [frames, 4096] complex64 (already windowed); thread t holds x[t + 64 a]
    X = x.reshape(-1, 64, 64)                                  # [a][t]
    A = dft64(np.swapaxes(X, -1, -2))                          # per t over a -> [t][k0]
    t = np.arange(64)[:, None]; k0 = np.arange(64)[None, :]
    if mode == "table":
        tw = w_exact(t * k0, N).astype(c64)
    else:
        j = k0 % 8; i = k0 // 8
        wj = w_exact(t * j, N).astype(c64); hi = w_exact(t * 8 * i, N).astype(c64)
        tw = np.where(i == 0, wj, np.where(j == 0, hi, (hi * wj).astype(c64))).astype(c64)
    A = (A * tw).astype(c64)
    B = dft64(np.swapaxes(A, -1, -2))                          # per k0 over t -> [k0][k1]: bin k0 + 64 k1
    return np.swapaxes(B, -1, -2).reshape(-1, N)


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    S = frames * N
    raw = O.synth("CS16", 0, S, 100 << 20, 0x5EC70002).view("<i2").reshape(-1, 2).astype(np.float64) / 32768.0
    w, wt = O.window("blackmanHarris", N)
    x = (raw[:, 0] + 1j * raw[:, 1]).reshape(frames, N) * w
    ref = np.fft.fft(x, axis=1)
    d_ref = 10 * np.log10(np.abs(ref) / wt)                    # the reference's scale: 5 log10|X|^2 + 10 log10(1/weight)
    print(f"{frames} frames of the C2 capture; bins above -100 dBFS: {(d_ref > -50).mean():.3f}, in -120..-100: {((d_ref <= -50) & (d_ref > -60)).mean():.4f}")
    for mode in ("table", "product"):
        got = fft4096_fp32(x.astype(c64), mode)
        d = 10 * np.log10(np.abs(got.astype(np.complex128)) / wt)
        e = np.abs(d - d_ref)
        hi, lo = d_ref > -50, (d_ref <= -50) & (d_ref > -60)
        print(f"{mode:8s} above -100 dBFS: max {e[hi].max():.5f} dB rms {np.sqrt((e[hi] ** 2).mean()):.6f} | -120..-100 dBFS: max {e[lo].max():.5f} dB "
              f"rms {np.sqrt((e[lo] ** 2).mean()):.6f}  (bins over 0.01 dB: {(e[lo] > 0.01).sum()} of {lo.sum()})")


if __name__ == "__main__":
    main()
