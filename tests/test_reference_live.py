"""Randomised differential tests of the Python host mirror against the reference's own helper code, executed live by
oracle/jsmini.py.  Needs /root/reference (present in the build container, absent on the GPU box): skipped elsewhere; the
committed fixtures of tests/test_reference_js.py cover the same helpers without it."""
import os
import random

import numpy as np
import pytest

REFERENCE = "/root/reference/lib"
pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs the reference sources")


@pytest.fixture(scope="module")
def js():
    from oracle.jsmini import Interp
    return Interp(REFERENCE)


def test_parse_freq_rate_random_names(js):
    from oracle.jsmini import UNDEF
    from spectro_b200 import parse_freq_rate as P
    m = js.load_module("./parseFreqRate", REFERENCE)
    rnd = random.Random(20261017)
    parts = ["g", "001", "433.92M", "868M", "250k", "1024k", "2.4K", "10.7m", "x", "1e3k", "-", "_", ".", " ", "/", "cu8", "cs16", "CF32", "wav", "3.5", "0", "M", "k", "7.", ".5M"]
    for _ in range(400):
        name = "".join(rnd.choice(parts) + rnd.choice(["_", "-", ".", "", " ", "/"]) for _ in range(rnd.randint(0, 7)))
        assert P.parseFreqRate(name) == js.to_py(js.call(m["parseFreqRate"], UNDEF, [name])), name
        assert P.parseFormat(name) == js.to_py(js.call(m["parseFormat"], UNDEF, [name])), name


def test_lookup_random_keys(js):
    from oracle.jsmini import UNDEF, JSObject
    from spectro_b200 import utils
    lookup = js.load_module("./utils", REFERENCE)["lookup"]
    names = ["rectangularWindow", "bartlettWindow", "hammingWindow", "hannWindow", "blackmanWindow", "blackmanHarrisWindow",
             "cube1_cmap", "viridis_cmap", "hot_cmap", "afmhot_cmap", "gist_heat_cmap", "parabola_cmap", "sox_cmap", "naive_cmap"]
    tab = JSObject(js.object_proto)
    for k in names:
        tab.props[k] = k
    py_tab = {k: k for k in names}
    rnd = random.Random(7)
    for _ in range(500):
        base = rnd.choice(names)
        key = base[:rnd.randint(0, len(base))]
        key = "".join(c.upper() if rnd.random() < 0.3 else c for c in key) + rnd.choice(["", "", "x", "_"])
        got = js.to_py(js.call(lookup, UNDEF, [tab, key]))
        assert utils.lookup(py_tab, key) == got, key


def test_windows_random_sizes(js):
    from oracle.jsmini import UNDEF
    from spectro_b200 import windows
    m = js.load_module("./windows", REFERENCE)
    rnd = random.Random(3)
    for kind in ("rectangular", "bartlett", "hamming", "hann", "blackman", "blackmanHarris"):
        for n in (2, 3, 5, 16, 100, rnd.randint(6, 600), 1 << rnd.randint(3, 11)):
            r = js.call(m[kind + "Window"], UNDEF, [n])
            ref_w, ref_weight = js.to_py(js.get_prop(r, "window")), js.get_prop(r, "weight")
            mine = getattr(windows, kind + "Window")(n)
            assert mine["weight"] == ref_weight and list(mine["window"]) == ref_w, (kind, n)


def test_sampleview_decode_and_slice_random(js):
    from oracle.jsmini import UNDEF
    from oracle import oracle as O
    from spectro_b200.samples import SampleView
    SV = js.load_module("./samples", REFERENCE)["default"]
    rnd = random.Random(11)
    for fmt in ("CU4", "CS4", "CU8", "CS8", "CU12", "CS12", "CU16", "CS16", "CU32", "CS32", "CF32", "CU64", "CS64", "CF64"):
        sw = SampleView(fmt).sampleWidth
        count = rnd.randint(5, 40)
        raw = bytes(rnd.randrange(256) for _ in range(sw * count))
        if fmt in ("CF32", "CF64"):                      # keep the floats finite
            raw = np.random.default_rng(5).standard_normal(2 * count).astype("<f4" if fmt == "CF32" else "<f8").tobytes()
        sv = js.construct(SV, [fmt, js.from_py(raw)])
        ref = [[js.call(js.get_prop(sv, "sampleI"), sv, [i]), js.call(js.get_prop(sv, "sampleQ"), sv, [i])] for i in range(count)]
        assert np.array_equal(O.decode(fmt, raw), np.array(ref, np.float64)), fmt
        py = SampleView(fmt, raw)
        for _ in range(5):
            k = rnd.randint(1, 6)
            end = int(len(raw) / sw)
            for i in range(k):
                b = js.call(js.get_prop(sv, "slice"), sv, [i, k, 0, end])
                assert bytes(py.slice(i, k, 0, end)) == bytes(b.data), (fmt, i, k)
