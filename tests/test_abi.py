"""CPU: the C-ABI library loads and exports every symbol include/rmx.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from radio_mapper_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, "include", "rmx.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"RMX_API\s+[\w\s\*]+?\b(rmx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = _header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_native.SIGNATURES) == names        # the ctypes table covers the header, no more, no less
    assert lib.rmx_version() == 201


def test_header_cites_reference_lines():
    with open(os.path.join(ROOT, "include", "rmx.h")) as f:
        text = f.read()
    for cite in ("buoy_node.py:392-398", "tdoa_processor.py:51", "signal_analyzer.py:92-99", "buoy_node.py:411-415"):
        assert cite in text


def test_bad_arguments_return_negative_status_with_message():
    lib = _native.load()
    h = ctypes.c_void_p()
    assert lib.rmx_plan_create(ctypes.byref(h), 0, 16, 16, 0) == -1
    assert b"n_signals" in lib.rmx_last_error()
    assert lib.rmx_plan_create(ctypes.byref(h), 1, 16, 24, 0) == -2          # not a power of two
    assert b"power of two" in lib.rmx_last_error()
    assert lib.rmx_plan_create(ctypes.byref(h), 1, 32, 16, 0) == -1          # n_samples > fft_len
    assert lib.rmx_plan_create(ctypes.byref(h), 1, 4, 8, 0) == -2            # below the minimum length
    assert lib.rmx_unpack_cu8(None, None, 5, None) == -1
    assert lib.rmx_unpack_cu8(None, None, 0, None) == 0                      # empty input is a no-op
    with pytest.raises(_native.RmxError):
        _native.check(lib.rmx_threshold_peaks(None, 10, 0.0, None, None, 4, None), "rmx_threshold_peaks")


def test_single_stage_plan_needs_no_device():
    """Plans whose transform fits one radix stage own no device tables: creating one, querying its
    layout and workspace size are pure host calls."""
    lib = _native.load()
    h = ctypes.c_void_p()
    assert lib.rmx_plan_create(ctypes.byref(h), 3, 8, 16, 0) == 0
    buf = (ctypes.c_int32 * 8)()
    assert lib.rmx_plan_layout(h, buf, 8) == 1 and buf[0] == 16
    assert lib.rmx_plan_workspace_bytes(h, 3) >= 3 * 16 * 8
    assert lib.rmx_plan_set_max_lag(h, 3) == 0
    # tuning knobs and developer flags are plan state (nothing on the launch path reads the environment)
    assert lib.rmx_plan_set_option(h, b"pair_run", 16) == 0 and lib.rmx_plan_set_option(h, b"pair_prefetch", 0) == 0
    assert lib.rmx_plan_set_option(h, b"pair_run", 5) == -1 and lib.rmx_plan_set_option(h, b"no_such_knob", 1) == -1
    assert b"no_such_knob" in lib.rmx_last_error()
    for knob, good, bad in ((b"pair_groups", (0, 2, 3), (1, 4)), (b"pair_ctas", (4, 5, 6), (3, 7)), (b"pair_store", (0, 1, 2), (3,))):
        assert all(lib.rmx_plan_set_option(h, knob, v) == 0 for v in good), knob
        assert all(lib.rmx_plan_set_option(h, knob, v) == -1 for v in bad), knob
    assert lib.rmx_welch_path(h, None) == 0                                  # 16-point plan: no cluster kernel
    assert lib.rmx_plan_destroy(h) == 0
    flags = _native.PLAN_NO_TMA | _native.PLAN_NO_PAIR_RUN | _native.PLAN_ROW_E8 | _native.plan_row_logn(12)
    assert lib.rmx_plan_create(ctypes.byref(h), 3, 8, 16, flags) == 0 and lib.rmx_plan_destroy(h) == 0


def test_library_never_reads_the_environment():
    """Developer switches travel as plan flags: no source of librmx calls getenv (the statically linked CUDA
    runtime imports it for its own CUDA_* variables, so the check is on the sources, not the symbol table)."""
    csrc = os.path.join(ROOT, "radio_mapper_b200", "csrc")
    for name in os.listdir(csrc):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(csrc, name)) as f:
                assert "getenv" not in f.read(), name
