"""CPU: the C-ABI library loads without a GPU and exports every symbol include/re2e_b200.h declares."""
import ctypes
import os
import subprocess
import sys

import pytest

from robust_e2e_gan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_declares_what_the_binding_binds():
    hdr = set(_lib.header_symbols())
    assert hdr == set(_lib.SIGNATURES), hdr ^ set(_lib.SIGNATURES)
    assert len(hdr) >= 20


def test_library_exports_every_header_symbol(built):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.header_symbols():
        assert hasattr(raw, name), "missing export %s" % name


def test_version_and_error_strings(built):
    assert built.re2e_abi_version() >= 1
    assert b"sm_100a" in built.re2e_build_info()
    assert built.re2e_error_string(0) == b"ok"
    assert b"argument" in built.re2e_error_string(-1)
    assert b"support" in built.re2e_error_string(-2)
    assert built.re2e_launch_count() == 0       # nothing has been launched in a CPU-only process


def test_no_silent_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from robust_e2e_gan_b200 import fbank
    with pytest.raises(RuntimeError, match="no CUDA device|not found"):
        fbank(torch.rand(1, 4, 257), torch.rand(257, 40))


def test_sass_is_sm100a_only_and_uses_tma(built):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if l.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass            # cp.async.bulk (TMA 1-D) in the AttLoc kernels
    assert "UBLKRED" in sass           # cp.reduce.async.bulk: d_pre accumulated by the TMA unit
    assert "UCGABAR" in sass           # cluster barrier
