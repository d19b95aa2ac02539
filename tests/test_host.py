"""Host-side mirror of the reference interface (no GPU): name lookup, constrain parsers,
file-name metadata, slicing, windows and colormaps vs the oracle's restatements."""
import numpy as np
import pytest

from oracle import oracle as O
from spectro_b200 import windows as W, cmaps as CM
from spectro_b200.parse_freq_rate import parseFormat, parseFreqRate
from spectro_b200.samples import SampleView
from spectro_b200.utils import js_parse_int, lookup
from spectro_b200 import sharding


def test_lookup_semantics():                                     # lib/utils.js:25-40
    t = W.windows
    assert lookup(t, "hannWindow") is W.hannWindow
    assert lookup(t, "hann") is W.hannWindow
    assert lookup(t, "HAMMING") is W.hammingWindow
    assert lookup(t, "blackman") is W.blackmanWindow               # first prefix match in key order
    assert lookup(t, "blackmanHarris") is W.blackmanHarrisWindow
    assert lookup(CM.cmaps, "hot") is CM.cmaps["hot_cmap"]
    assert lookup(CM.cmaps, "parula") is None                      # demo's 'parula' falls back to cube1
    assert lookup(t, None) is None and lookup(t, W.hannWindow) is W.hannWindow


def test_js_parse_int():                                         # lib/spectroplot.js:239-250
    assert js_parse_int("1024", 512) == 1024 and js_parse_int("12px", 0) == 12
    assert js_parse_int("abc", 512) == 512 and js_parse_int("0", 30) == 30 and js_parse_int(6.7, 0) == 6
    assert js_parse_int(None, 7) == 7 and js_parse_int("-3", 1) == -3


@pytest.mark.parametrize("name", ["rectangular", "bartlett", "hamming", "hann", "blackman", "blackmanHarris"])
@pytest.mark.parametrize("n", [8, 128, 4096])
def test_windows_match_oracle(name, n):                          # lib/windows.js:14-88
    r = lookup(W.windows, name + "Window")(n)
    w, wt = O.window(name, n)
    assert r["weight"] == wt and np.array_equal(np.array(r["window"]), w)


def test_computed_cmaps_match_oracle():                          # lib/soxcmap.js, lib/naivecmap.js
    assert np.array_equal(CM.cmap_bytes(CM.cmaps["sox_cmap"]), O.cmap_sox())
    for k in ("naive", "grayscale", "roentgen", "phosphor"):
        assert np.array_equal(CM.cmap_bytes(CM.cmaps[k + "_cmap"]), O.cmap_naive(k)), k
    assert len(CM.cmaps["parabola_cmap"]) == 64 and all(len(CM.cmaps[k]) == 256 for k in CM.cmaps if k != "parabola_cmap")
    assert CM.cmaps["cube1_cmap"][0] == [116, 0, 129] and CM.cmaps["viridis_cmap"][0] == [68, 1, 84]
    assert list(CM.cmaps)[:2] == ["cube1_cmap", "sox_cmap"]


def test_parse_freq_rate():                                      # lib/parseFreqRate.js:16-70
    assert parseFreqRate("g001_433.92M_250k.cu8") == {"freq": 433920000.0, "rate": 250000.0}
    assert parseFreqRate("a/b/c_868M_1000k.cs16")["freq"] == 868000000.0
    assert parseFreqRate("plain.cu8") == {"freq": 0, "rate": 1}
    assert parseFreqRate("") == {"freq": 0, "rate": 0}
    assert parseFormat("x_433M_250k.cu8") == "CU8" and parseFormat("noext") == "?" and parseFormat("") == "?"


def test_sample_view_table_and_slice():                          # lib/samples.js:30-169,253-258
    v = SampleView("cs16", bytes(4 * 1001))
    assert v.sampleWidth == 4 and v.sampleCount == 1001
    assert SampleView("whatever").canonical == "CU8" and SampleView("cfile").canonical == "CF32"
    s = v.slice(1, 4, 0, 1001)
    assert len(s) == 4 * 250                                     # sliceLength = 4 * ~~(1001/4)
    with pytest.raises(ValueError):
        SampleView("cs16", bytes(7))
    assert SampleView("cu12", bytes(10)).sampleCount == 10 / 3


def test_shard_plan_covers_every_frame_once():
    total_samples, n, width = 1 << 20, 4096, 1000
    stride = (total_samples - n) / (width - 1)
    for world in (1, 2, 3, 8):
        shards = sharding.plan_shards(total_samples, n, width, world)
        assert [s["frame_first"] for s in shards][0] == 0
        assert sum(s["width"] for s in shards) == width
        for a, b in zip(shards, shards[1:]):
            assert a["frame_first"] + a["width"] == b["frame_first"]
        for s in shards:
            p_first = int(0.5 + stride * s["frame_first"])
            p_last = int(0.5 + stride * (s["frame_first"] + s["width"] - 1)) + n
            assert s["sample_first"] <= p_first and s["sample_first"] + s["sample_count"] >= p_last
            assert s["sample_first"] % 4 == 0                     # 16-byte alignment for <= 4-byte... samples
            assert s["sample_first"] + s["sample_count"] <= total_samples


def test_cu8_two_instruction_division_is_correctly_rounded():
    """csrc/sp_device.cuh decodes cu8 as x / 255 = fma(x, r_hi, x * r_lo) with 1/255 split into two fp32 constants (one packed
    instruction less than the Markstein form).  The claim that this is the correctly rounded quotient - i.e. fround((c - 127.5) /
    127.5) of lib/samples.js:313-330 - for every code is checked here exactly, with the constants read from the source."""
    import os
    import re
    from fractions import Fraction
    src = open(os.path.join(os.path.dirname(__file__), "..", "spectroplot-js_b200", "csrc", "sp_device.cuh")).read()
    m = re.search(r"r_hi = ([0-9.eE+-]+)f, r_lo = ([0-9.eE+-]+)f", src)
    assert m, "constants not found"
    r_hi, r_lo = np.float32(m.group(1)), np.float32(m.group(2))
    assert r_hi == np.float32(1.0 / 255.0) and r_lo == np.float32(1.0 / 255.0 - float(r_hi))

    def rn32(fr):                                   # correctly rounded fp32 of an exact rational (no double rounding)
        lo, hi = np.float32(float(fr)), None
        cands = {lo, np.nextafter(lo, np.float32(np.inf)), np.nextafter(lo, np.float32(-np.inf))}
        best = min(cands, key=lambda v: (abs(Fraction(float(v)) - fr), int(np.float32(v).view(np.uint32)) & 1))
        return np.float32(best)

    for c in range(256):
        x = np.float32(2 * c - 255)
        lo = np.float32(x * r_lo)                                                  # FMUL2
        q = rn32(Fraction(float(x)) * Fraction(float(r_hi)) + Fraction(float(lo)))  # FFMA2: one rounding of the exact sum
        ref = np.float32((c - 127.5) / 127.5)                                      # Math.fround of the reference's double value
        assert q == ref, (c, float(q), float(ref))
