"""The C-ABI library loads and exports every symbol include/spectro_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "spectro_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from spectro_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in spectro_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.sp_abi_version() == 1


def test_format_table_matches_reference():
    from spectro_b200 import _lib
    lib = _lib.load()
    f = lambda s: lib.sp_format_from_name(s.encode())
    assert [f(n) for n in _lib.FORMATS] == list(range(14))
    assert f("cu8") == f("DATA") == f("Complex16U") == f("???") == 2       # lib/samples.js:48,149-155
    assert f("COMPLEX16S") == 3 and f("cfile") == f("complex") == 12       # lib/samples.js:55,126
    assert [lib.sp_sample_width(i) for i in range(14)] == [1, 1, 2, 2, 3, 3, 4, 4, 8, 8, 16, 16, 8, 16]
    assert [lib.sp_element_size(i) for i in range(14)] == [1, 1, 1, 1, 1, 1, 2, 2, 4, 4, 4, 4, 4, 8]
    assert lib.sp_sample_width(99) < 0
    assert lib.sp_format_name(7) == b"CS16"


def test_struct_layout_matches_header():
    from spectro_b200 import _lib
    # sizes implied by the header on LP64: request 14 x 8 = 120 bytes (with the 4 packed int32 pairs), reply 72
    assert C.sizeof(_lib.Request) == 120
    assert C.sizeof(_lib.Reply) == 80
    assert _lib.Request.windowc.offset == 56 and _lib.Request.total_byte_length.offset == 88
    assert _lib.Reply.dBfs_min.offset == 48 and _lib.Reply.device_ms.offset == 64


def test_synth_lut_identical_to_oracle():
    from spectro_b200 import _lib
    from oracle import oracle as O
    lut = np.empty(4096, np.int16)
    _lib.load().sp_synth_lut(lut.ctypes.data_as(C.c_void_p))
    assert np.array_equal(lut, O.synth_lut())


def test_no_cpu_fallback():
    """Without a usable sm_100 device the engine refuses to exist."""
    import spectro_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(spectro_b200.SpError) as ei:
        spectro_b200.Engine(0)
    assert ei.value.name == "SP_E_NO_DEVICE"


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, link or call it."""
    pat = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|libspectro_oracle|\bspo_[a-z_]+\s*\(|np_restatement")
    pkg = os.path.join(ROOT, "spectroplot-js_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".js", ".cc", ".h", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not pat.search(src), f"{f} reaches into the oracle"
