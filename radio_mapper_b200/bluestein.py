"""Arbitrary-length DFT on the GPU (Bluestein chirp-z over the power-of-two tile FFT).

The reference transforms whole captures of any length in one call (signal_analyzer.py:62-63,
e.g. 10 240 000 points); pocketfft handles that with mixed radices.  Here
    X[k] = w[k] * sum_n (x[n] w[n]) * conj(w[k-n]),   w[n] = exp(-i*pi*n^2/N)
and the convolution is one `rmx_xcorr_full` call on two padded signals.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _native, engine

_lib = _native.load()


def dft(x: torch.Tensor, plans: Optional[dict] = None) -> torch.Tensor:
    """complex64[N] (device, any N >= 1) -> complex64[N] unnormalised forward DFT, natural order."""
    engine._require_cuda(x, torch.complex64, "x")
    n = x.numel()
    lp = max(16, engine.next_pow2(2 * n - 1))
    key = ("bluestein", lp)
    plan = plans.get(key) if plans is not None else None
    if plan is None:
        plan = engine.Plan(2, lp, lp, device=x.device)
        if plans is not None:
            plans[key] = plan
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    sig = torch.empty((2, lp), dtype=torch.complex64, device=x.device)
    with torch.cuda.device(x.device):
        _native.check(_lib.rmx_bluestein_prepare(ctypes.c_void_p(x.data_ptr()), n, lp, ctypes.c_void_p(sig[0].data_ptr()),
                                                 ctypes.c_void_p(sig[1].data_ptr()), stream), "rmx_bluestein_prepare")
        spectra = plan.forward_c64(sig)
        pairs = torch.tensor([[1, 0]], dtype=torch.int32, device=x.device)     # i = chirp, j = a
        conv = plan.xcorr_full(spectra, pairs)
        out = torch.empty(n, dtype=torch.complex64, device=x.device)
        _native.check(_lib.rmx_bluestein_finish(ctypes.c_void_p(conv.data_ptr()), n, ctypes.c_void_p(out.data_ptr()), stream),
                      "rmx_bluestein_finish")
    return out


def spectrum_db(x: torch.Tensor, shift: bool = False, plans: Optional[dict] = None) -> torch.Tensor:
    """20*log10(|DFT(x)| + 1e-12), float32[N]."""
    X = dft(x, plans)
    out = torch.empty(X.numel(), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _native.check(_lib.rmx_abs_db(ctypes.c_void_p(X.data_ptr()), X.numel(), ctypes.c_void_p(out.data_ptr()),
                                      int(bool(shift)), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "rmx_abs_db")
    return out
