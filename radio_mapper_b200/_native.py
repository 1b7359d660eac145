"""ctypes binding of librmx.so (include/rmx.h).  There is NO fallback: if the CUDA library
is missing the import raises, and every product path that needs it fails loudly."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RMX_LIB_PATH: developer override used by tools/variant_build.py to A/B kernel build variants
LIB_PATH = os.environ.get("RMX_LIB_PATH") or os.path.join(_HERE, "librmx.so")

c_void_p, c_int, c_size_t, c_uint, c_float, c_double, c_longlong = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint, ctypes.c_float, ctypes.c_double,
    ctypes.c_longlong)

# name -> (restype, argtypes); must list every symbol include/rmx.h declares
SIGNATURES = {
    "rmx_last_error": (ctypes.c_char_p, []),
    "rmx_version": (c_int, []),
    "rmx_unpack_cu8": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmx_plan_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_size_t, c_size_t, c_uint]),
    "rmx_plan_destroy": (c_int, [c_void_p]),
    "rmx_plan_set_option": (c_int, [c_void_p, ctypes.c_char_p, c_longlong]),
    "rmx_welch_path": (c_int, [c_void_p, c_void_p]),
    "rmx_plan_layout": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_int32), c_int]),
    "rmx_plan_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "rmx_plan_set_max_lag": (c_int, [c_void_p, c_longlong]),
    "rmx_plan_set_search_mode": (c_int, [c_void_p, c_int]),
    "rmx_fft_forward_cu8": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "rmx_fft_forward_c64": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "rmx_signal_stats_c64": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "rmx_xcorr_full": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "rmx_bluestein_prepare": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "rmx_bluestein_finish": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "rmx_abs_db": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "rmx_profile_enable": (c_int, [c_void_p, c_int]),
    "rmx_profile_collect": (c_int, [c_void_p, c_void_p, c_int]),
    "rmx_spectrum_natural": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "rmx_xcorr_pairs_peak": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmx_spectrum_db": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rmx_welch_psd": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_size_t, c_void_p]),
    "rmx_welch_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "rmx_power_db": (c_int, [c_void_p, c_void_p, c_size_t, c_float, c_void_p]),
    "rmx_threshold_peaks": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "rmx_select_by_distance_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rmx_mean_median": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmx_signal_stats": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "rmx_peak_bandwidth_batch": (c_int, [c_void_p, c_int, c_int, c_size_t, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "rmx_find_peaks_batch": (c_int, [c_void_p, c_int, c_int, c_size_t, c_float, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                     c_void_p, c_int, c_void_p, c_void_p]),
    "rmx_signal_energy": (c_int, [c_void_p, c_size_t, c_int, c_size_t, c_void_p, c_void_p]),
}

# plan flags (include/rmx.h)
PLAN_TWIDDLE_IN_COL, PLAN_NO_TMA, PLAN_NO_PAIR_RUN, PLAN_NO_WELCH_CLUSTER, PLAN_ROW_E8 = 0x01, 0x02, 0x04, 0x08, 0x10


def plan_row_logn(n: int) -> int:
    return (int(n) & 0x1F) << 8


def flags_from_env() -> int:
    """Developer switches: the RMX_* environment variables are read HERE, once per plan creation on the Python
    side, and handed to rmx_plan_create as flags; the library itself never reads the environment."""
    f = 0
    if os.environ.get("RMX_TWIDDLE_IN_COL"):
        f |= PLAN_TWIDDLE_IN_COL
    if os.environ.get("RMX_NO_TMA"):
        f |= PLAN_NO_TMA
    if os.environ.get("RMX_NO_PAIR_RUN"):
        f |= PLAN_NO_PAIR_RUN
    if os.environ.get("RMX_NO_WELCH_CLUSTER"):
        f |= PLAN_NO_WELCH_CLUSTER
    if os.environ.get("RMX_CONTIG_LOGN"):
        f |= plan_row_logn(int(os.environ["RMX_CONTIG_LOGN"]))
    return f


_lib = None


class RmxError(RuntimeError):
    """Raised when a librmx call returns a negative status."""


def load():
    """Load librmx.so once.  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "radio_mapper_b200: %s is missing — build it with `python radio_mapper_b200/csrc/build.py` "
            "(there is no CPU fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what=""):
    if rc < 0:
        msg = load().rmx_last_error()
        raise RmxError("%s failed (%d): %s" % (what or "librmx call", rc, msg.decode() if msg else ""))
    return rc
