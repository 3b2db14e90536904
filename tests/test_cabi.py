"""The C-ABI library loads without a GPU and exports every symbol include/s2d_b200.h declares;
the ctypes mirror of s2d_video_desc has the C layout. No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "s2d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from s2d_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/s2d_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in s2d_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_descriptor_layout_and_version():
    from s2d_b200 import _lib
    lib = _lib.load()
    assert lib.s2d_desc_size() == C.sizeof(_lib.VideoDesc) == 128
    assert lib.s2d_version() >= 100
    assert lib.s2d_last_error() is not None


def test_argument_errors_are_reported_without_a_gpu():
    from s2d_b200 import _lib
    _lib.load()
    with pytest.raises(_lib.S2DError, match="null pointer"):
        _lib.call("s2d_vis_reduce", None, 1, 10, None, None, None)
    with pytest.raises(_lib.S2DError, match="1024 frames"):
        _lib.call("s2d_windows", 1, 1, 10, 33, 1, 1, 1, 1, 1, 0.3, 1, 1, 1, 1, 1, 1, 1, 1, 1, None)


def test_missing_library_fails_loudly(monkeypatch):
    from s2d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libs2d_b200.so")
    with pytest.raises(_lib.S2DError, match="no CPU fallback"):
        _lib.load()


def test_gram_tiling_query_is_host_only():
    """which Gram tiling a shape gets (2: 256x256 two-m-tile kernel, 1: 128x256, 0: 128x128) - no GPU needed."""
    from s2d_b200 import _lib
    _lib.load()
    got = {}
    for F, L in [(36, 21), (64, 31), (30, 18), (24, 14), (24, 11), (48, 6), (100, 3), (3, 4)]:
        t = C.c_int(-1)
        _lib.call("s2d_overlap_gram_tiling", F, L, C.byref(t))
        got[(F, L)] = t.value
    # (24, 14), (24, 11), (48, 6): the 256 x 256 kernel with two operand stages instead of three, so that the label ring
    # of shapes with few labels per frame fits (round 1 fell back to 128-row tiles there: the C1 shape ran at 18 %)
    assert got == {(36, 21): 2, (64, 31): 2, (30, 18): 2, (24, 14): 2, (24, 11): 2, (48, 6): 2, (100, 3): 0, (3, 4): 0}
    n = C.c_int64()
    for F, L in got:                       # the scratch size covers whichever tiling is chosen
        _lib.call("s2d_overlap_gram_work_ints", F, L, 64 * 48, C.byref(n))
        assert n.value >= (F * L) ** 2
