"""Window generators — host side of the path (reference lib/windows.js:14-88).

The reference computes the window on the main thread in float64 and ships it in the worker
message as `windowc`; `block_norm = 1 / weight` (lib/spectroplot.js:1114-1116).  Same symmetric
(n-1) forms, same accumulation order of `weight`.  Each returns {"window": list, "weight": float}.
"""
import math


def _gen(n, f):
    window = [0.0] * n
    weight = 0.0
    for i in range(n):
        window[i] = f(i)
        weight += window[i]
    return {"window": window, "weight": weight}


def rectangularWindow(n):
    return _gen(n, lambda i: 1.0)


def bartlettWindow(n):
    return _gen(n, lambda i: 1.0 - abs((i - 0.5 * (n - 1)) / (0.5 * (n - 1))))


def hammingWindow(n):
    return _gen(n, lambda i: 0.54 - 0.46 * math.cos(2.0 * math.pi * i / (n - 1)))


def hannWindow(n):
    return _gen(n, lambda i: 0.5 * (1.0 - math.cos(2.0 * math.pi * i / (n - 1))))


def blackmanWindow(n):
    return _gen(n, lambda i: 0.42 - (0.5 * math.cos((2.0 * math.pi * i) / (n - 1)))
                + (0.08 * math.cos((4.0 * math.pi * i) / (n - 1))))


def blackmanHarrisWindow(n):
    return _gen(n, lambda i: 0.35875 - (0.48829 * math.cos((2.0 * math.pi * i) / (n - 1)))
                + (0.14128 * math.cos((4.0 * math.pi * i) / (n - 1)))
                - (0.01168 * math.cos((6.0 * math.pi * i) / (n - 1))))


# key order matters for lookup()'s prefix match (lib/utils.js:36-38): module export order
windows = {
    "rectangularWindow": rectangularWindow,
    "bartlettWindow": bartlettWindow,
    "hammingWindow": hammingWindow,
    "hannWindow": hannWindow,
    "blackmanWindow": blackmanWindow,
    "blackmanHarrisWindow": blackmanHarrisWindow,
}
