"""The C-ABI library loads without a GPU and exports exactly what include/rqk.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from generative_ranking_recommender_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def header_names():
    hdr = open(os.path.join(ROOT, "include", "rqk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(rqk_[a-z0-9_]+)\s*\(", hdr))


def test_library_is_built_and_loads_without_gpu():
    assert os.path.exists(_lib.LIB_PATH), "run python -m generative_ranking_recommender_b200.build"
    L = _lib.lib()
    assert L.rqk_version() >= 100
    assert L.rqk_last_error() is not None


def test_every_declared_symbol_is_exported(header_names):
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_names:
        assert hasattr(L, name), f"{name} declared in include/rqk.h but not exported"


def test_bindings_and_header_agree(header_names):
    assert header_names == set(_lib.SIGNATURES), (header_names ^ set(_lib.SIGNATURES))


def test_exported_symbols_are_c_linkage():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in _lib.SIGNATURES:
        assert name in exported


def test_sizes_only_entry_points_work_on_cpu():
    L = _lib.lib()
    assert L.rqk_score_workspace_bytes(1000, 128, 512) > 0
    assert L.rqk_auction_workspace_bytes(100000, 128) > 100000 * 4
    assert L.rqk_centroid_workspace_bytes(100000, 128, 512) > 0
    assert L.rqk_encode_workspace_bytes(100000, 512, 256) > 0
    ks = (ctypes.c_int32 * 3)(128, 128, 256)
    assert L.rqk_encode_fused_supported(512, 3, ks, ks, 0) == 1 and L.rqk_encode_fused_supported(512, 3, ks, ks, 1) == 1
    assert 0 < L.rqk_encode_fused_workspace_bytes(1000000, 512, 3, ks) < 16 << 20      # no N x D scratch
    odd = (ctypes.c_int32 * 3)(8, 8, 16)
    assert L.rqk_encode_fused_supported(64, 3, odd, odd, 0) == 0 and L.rqk_encode_fused_workspace_bytes(10, 64, 3, odd) == 0
    lay = _lib.AuctionLayout()
    assert L.rqk_auction_layout_query(100000, 128, ctypes.byref(lay)) == 0
    assert lay.reduce_count == 128 * 256 + 2 * 128 + 2
    assert lay.total_bytes == L.rqk_auction_workspace_bytes(100000, 128)
    # argument errors come back as codes + message, never as exceptions or crashes
    assert L.rqk_auction_layout_query(0, 128, ctypes.byref(lay)) < 0
    assert b"rqk_auction_layout_query" in L.rqk_last_error()


def test_sass_is_blackwell_native():
    """tcgen05 / TMA / TMEM mnemonics must be present in the shipped cubin (B200_PROFILING.md)."""
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True, stderr=subprocess.DEVNULL)
    except (FileNotFoundError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "UTMALDG" in sass, "no TMA tensor load in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"
