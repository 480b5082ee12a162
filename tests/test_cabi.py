"""The C-ABI library loads and exports every symbol include/ai_b200.h declares.  CPU only:
no compute is launched; argument validation that precedes any CUDA call is exercised."""
import ctypes
import os
import re

import numpy as np
import pytest

from adaptive_interpolation_b200 import _cabi, build
from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.lib()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ai_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ai_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_cabi.EXPORTS)


def test_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_version_string(lib):
    assert b"sm_100a" in lib.ai_version()


def test_only_sm100a_code_in_library():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _cabi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_needs_no_gpu(lib):
    h = ctypes.c_void_p()
    assert lib.ai_table_create(7, 3, 0, 1, None, None, 0, ctypes.byref(h)) == 1      # AI_ERR_INVALID: basis
    assert b"basis" in lib.ai_last_error()
    assert lib.ai_table_create(0, 99, 0, 1, None, None, 0, ctypes.byref(h)) == 1     # order
    assert lib.ai_table_create(0, 3, 5, 1, None, None, 0, ctypes.byref(h)) == 1      # dtype
    assert lib.ai_table_create(0, 3, 0, 1, None, None, 0, None) == 1                 # out == NULL
    assert lib.ai_eval_f32(None, None, None, 0, None) == 1                           # NULL table
    assert lib.ai_eval_index_f64(None, None, None, 0, None) == 1
    assert lib.ai_table_get_info(None, None) == 1
    assert lib.ai_table_destroy(None) == 0


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libai_b200.so")
    with pytest.raises(ImportError):
        _cabi.lib()


def test_host_copy_pool_copies_every_byte(lib):
    """ai_host_copy: the threaded, non-temporal-store copy the staged host pipeline runs on pageable caller
    buffers (page-aligned shares, 16-byte-aligned streaming stores with memcpy head / tail).  Any size,
    any alignment of source and destination, any thread count; nothing outside [dst, dst + bytes) written."""
    L = _cabi.lib()
    rng = np.random.default_rng(0)
    src_buf = rng.integers(0, 256, (1 << 22) + 256, dtype=np.uint8)
    for nbytes in (0, 1, 15, 16, 17, 63, 64, 65, 4095, 4096, 4097, 65536 + 3, (1 << 20) + 11, (1 << 22) - 7):
        for so, do, threads in ((0, 0, 1), (1, 0, 3), (0, 5, 4), (7, 13, 16), (64, 32, 0)):
            dst_buf = np.full(nbytes + 128, 0xA5, dtype=np.uint8)
            src = src_buf[so:so + nbytes]
            _cabi.check(L.ai_host_copy(dst_buf[do:].ctypes.data, src.ctypes.data, nbytes, threads))
            assert np.array_equal(dst_buf[do:do + nbytes], src), (nbytes, so, do, threads)
            assert np.all(dst_buf[:do] == 0xA5) and np.all(dst_buf[do + nbytes:] == 0xA5), (nbytes, so, do, threads)
    assert L.ai_host_copy(None, src_buf.ctypes.data, 8, 1) == 1                  # AI_ERR_INVALID
    assert L.ai_host_copy(src_buf.ctypes.data, src_buf.ctypes.data, -1, 1) == 1
