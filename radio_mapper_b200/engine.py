"""Device-side engine: torch tensors in, librmx (hand-written sm_100a kernels) underneath.

torch is used for device memory, streams and (elsewhere) torch.distributed — plumbing only.
Every numeric stage runs in librmx.so through the C ABI of include/rmx.h.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native

_lib = _native.load()


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (the hot path has no CPU fallback)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must have dtype %s (got %s)" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def next_pow2(n: int) -> int:
    return 1 << max(0, int(n - 1).bit_length())


def correlation_fft_len(n_samples: int) -> int:
    """Smallest power of two >= 2N-1 (and >= 16): linear correlation without aliasing."""
    return max(16, next_pow2(2 * int(n_samples) - 1))


def pair_table(n: int) -> np.ndarray:
    """All i<j in the enumeration order of tdoa_processor.py:156-157, int32[P, 2]."""
    i, j = np.triu_indices(int(n), k=1)
    return np.stack([i, j], axis=1).astype(np.int32)


def unpack_cu8(iq_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8[..., 2N] interleaved I,Q -> complex64[..., N], bit-exact with the reference unpack."""
    _require_cuda(iq_u8, torch.uint8, "iq_u8")
    if iq_u8.shape[-1] % 2:
        raise ValueError("cu8 buffers hold an even number of bytes (I,Q pairs)")
    n = iq_u8.numel() // 2
    if out is None:
        out = torch.empty(iq_u8.shape[:-1] + (iq_u8.shape[-1] // 2,), dtype=torch.complex64, device=iq_u8.device)
    _require_cuda(out, torch.complex64, "out")
    with torch.cuda.device(iq_u8.device):
        _native.check(_lib.rmx_unpack_cu8(_ptr(iq_u8), _ptr(out), n, _stream_ptr()), "rmx_unpack_cu8")
    return out


class Plan:
    """Batched FFT / correlation plan for `n_signals` signals of `n_samples` samples each,
    zero-padded to `fft_len` (power of two)."""

    def __init__(self, n_signals: int, n_samples: int, fft_len: Optional[int] = None, device=None,
                 flags: Optional[int] = None, options: Optional[dict] = None):
        """flags: RMX_PLAN_* developer switches (default: taken once from the RMX_* environment variables, see
        _native.flags_from_env); options: {name: value} for rmx_plan_set_option."""
        if not torch.cuda.is_available():
            raise RuntimeError("radio_mapper_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_signals = int(n_signals)
        self.n_samples = int(n_samples)
        self.fft_len = int(fft_len) if fft_len is not None else correlation_fft_len(n_samples)
        self._h = ctypes.c_void_p()
        self.flags = _native.flags_from_env() if flags is None else int(flags)
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_plan_create(ctypes.byref(self._h), self.n_signals, self.n_samples, self.fft_len, self.flags),
                          "rmx_plan_create")
        for name, value in (options or {}).items():
            self.set_option(name, value)
        buf = (ctypes.c_int32 * 8)()
        n = _native.check(_lib.rmx_plan_layout(self._h, buf, 8), "rmx_plan_layout")
        self.pass_lengths = [int(buf[i]) for i in range(n)]
        self._workspace: Optional[torch.Tensor] = None
        self.max_lag: Optional[int] = None

    def __del__(self):
        # rmx_plan_destroy synchronises the device before freeing the plan's tables (launches that read them
        # may still be queued)
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                with torch.cuda.device(self.device):
                    _lib.rmx_plan_destroy(h)
            except Exception:
                pass

    def set_option(self, name: str, value: int):
        """Tuning knob of the plan (include/rmx.h: rmx_plan_set_option)."""
        _native.check(_lib.rmx_plan_set_option(self._h, name.encode(), int(value)), "rmx_plan_set_option(%s)" % name)

    # ---- layout ------------------------------------------------------------------------
    def layout_freq_index(self) -> np.ndarray:
        """freq bin held at each position of the plan's spectrum layout (host, for tests/tools)."""
        pos = np.arange(self.fft_len, dtype=np.int64)
        freq = np.zeros(self.fft_len, dtype=np.int64)
        weight, m = 1, self.fft_len
        for n in self.pass_lengths:
            s = m // n
            freq += ((pos // s) % n) * weight
            weight *= n
            m = s
        return freq

    # ---- workspace -----------------------------------------------------------------------
    def _get_workspace(self, nbytes: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._workspace

    def workspace_bytes(self, n_pairs: int) -> int:
        return int(_lib.rmx_plan_workspace_bytes(self._h, int(n_pairs)))

    def set_max_lag(self, max_lag: Optional[int]):
        self.max_lag = None if max_lag is None else int(max_lag)
        _native.check(_lib.rmx_plan_set_max_lag(self._h, -1 if max_lag is None else int(max_lag)), "rmx_plan_set_max_lag")

    def set_search_mode(self, force_full: bool):
        """force_full=True disables the one-pass windowed search (full inverse + masked arg-max)."""
        _native.check(_lib.rmx_plan_set_search_mode(self._h, int(bool(force_full))), "rmx_plan_set_search_mode")

    # ---- stages --------------------------------------------------------------------------
    def forward(self, iq_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """cu8[n_signals, 2N] -> spectra complex64[n_signals, L] in the plan layout.

        `iq_u8` may be a row-strided view (stride(0) >= 2N bytes, stride(1) == 1), e.g. one
        window cut out of a [buoy, stream] buffer."""
        if not isinstance(iq_u8, torch.Tensor) or not iq_u8.is_cuda or iq_u8.dtype != torch.uint8:
            raise TypeError("iq_u8 must be a CUDA uint8 tensor (the hot path has no CPU fallback)")
        if iq_u8.ndim != 2 or iq_u8.shape[0] != self.n_signals or iq_u8.shape[1] != 2 * self.n_samples:
            raise ValueError("iq_u8 must be uint8[%d, %d] (got %s)" % (self.n_signals, 2 * self.n_samples, tuple(iq_u8.shape)))
        if iq_u8.stride(1) != 1 or (self.n_signals > 1 and (iq_u8.stride(0) < 2 * self.n_samples or iq_u8.stride(0) % 2)):
            raise ValueError("iq_u8 rows must be contiguous with an even row stride")
        stride = iq_u8.stride(0) if self.n_signals > 1 else 0
        if out is None:
            out = torch.empty((self.n_signals, self.fft_len), dtype=torch.complex64, device=self.device)
        _require_cuda(out, torch.complex64, "out")
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_fft_forward_cu8(self._h, _ptr(iq_u8), stride, _ptr(out), _stream_ptr()),
                          "rmx_fft_forward_cu8")
        return out

    def forward_c64(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """complex64[n_signals, N] (device) -> spectra complex64[n_signals, L] in the plan layout."""
        _require_cuda(x, torch.complex64, "x")
        if x.numel() != self.n_signals * self.n_samples:
            raise ValueError("x has %d samples, plan expects %d x %d" % (x.numel(), self.n_signals, self.n_samples))
        if out is None:
            out = torch.empty((self.n_signals, self.fft_len), dtype=torch.complex64, device=self.device)
        _require_cuda(out, torch.complex64, "out")
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_fft_forward_c64(self._h, _ptr(x), 0, _ptr(out), _stream_ptr()), "rmx_fft_forward_c64")
        return out

    # ---- profiling -----------------------------------------------------------------------
    def profile(self, enable: bool = True):
        _native.check(_lib.rmx_profile_enable(self._h, int(bool(enable))), "rmx_profile_enable")

    def profile_collect(self):
        """{kernel name: (launches, total_ms)} since profile(True); synchronises the recorded events."""
        class _Entry(ctypes.Structure):
            _fields_ = [("name", ctypes.c_char * 32), ("launches", ctypes.c_int32), ("total_ms", ctypes.c_float)]
        buf = (_Entry * 32)()
        n = _native.check(_lib.rmx_profile_collect(self._h, ctypes.cast(buf, ctypes.c_void_p), 32), "rmx_profile_collect")
        return {buf[i].name.decode(): (int(buf[i].launches), float(buf[i].total_ms)) for i in range(n)}

    def spectrum_natural(self, spectra: torch.Tensor) -> torch.Tensor:
        _require_cuda(spectra, torch.complex64, "spectra")
        out = torch.empty_like(spectra)
        n = spectra.numel() // self.fft_len
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_spectrum_natural(self._h, _ptr(spectra), _ptr(out), n, _stream_ptr()), "rmx_spectrum_natural")
        return out

    def spectrum_db(self, spectra: torch.Tensor, shift: bool = False) -> torch.Tensor:
        _require_cuda(spectra, torch.complex64, "spectra")
        n = spectra.numel() // self.fft_len
        out = torch.empty((n, self.fft_len), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_spectrum_db(self._h, _ptr(spectra), _ptr(out), n, int(bool(shift)), _stream_ptr()),
                          "rmx_spectrum_db")
        return out

    def xcorr_pairs_peak(self, spectra: torch.Tensor, pairs: torch.Tensor, out: Optional[torch.Tensor] = None,
                         max_pairs_in_flight: Optional[int] = None) -> torch.Tensor:
        """For each (i, j) row of `pairs` (int32[P, 2], device): peak of ifft(X_j conj X_i).

        Returns int32[P, 4] records (lag, peak, frac, search value): use `peaks_to_numpy`."""
        _require_cuda(spectra, torch.complex64, "spectra")
        _require_cuda(pairs, torch.int32, "pairs")
        if pairs.ndim != 2 or pairs.shape[1] != 2:
            raise ValueError("pairs must be int32[P, 2]")
        n_pairs = pairs.shape[0]
        if out is None:
            out = torch.empty((n_pairs, 4), dtype=torch.int32, device=self.device)
        _require_cuda(out, torch.int32, "out")
        if n_pairs == 0:
            return out
        chunk = n_pairs if max_pairs_in_flight is None else max(1, min(n_pairs, int(max_pairs_in_flight)))
        ws = self._get_workspace(self.workspace_bytes(chunk))
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_xcorr_pairs_peak(self._h, _ptr(spectra), _ptr(pairs), n_pairs, _ptr(out), _ptr(ws),
                                                    ws.numel(), _stream_ptr()), "rmx_xcorr_pairs_peak")
        return out

    def xcorr_full(self, spectra: torch.Tensor, pairs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Full correlation ifft(X_j conj X_i) per pair, natural order: complex64[P, L]."""
        _require_cuda(spectra, torch.complex64, "spectra")
        _require_cuda(pairs, torch.int32, "pairs")
        n_pairs = pairs.shape[0]
        if out is None:
            out = torch.empty((n_pairs, self.fft_len), dtype=torch.complex64, device=self.device)
        _require_cuda(out, torch.complex64, "out")
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_xcorr_full(self._h, _ptr(spectra), _ptr(pairs), n_pairs, _ptr(out), _stream_ptr()),
                          "rmx_xcorr_full")
        return out

    def welch_psd(self, iq_u8: torch.Tensor, sample_rate: float, segments_in_flight: Optional[int] = None) -> torch.Tensor:
        """Welch PSD of n_signals segments of nperseg = fft_len samples each (float32[L], natural order).
        segments_in_flight bounds the spectra workspace (8*fft_len bytes per segment); the default takes as many
        segments per launch as fit 1 GiB -- fewer, larger launches measured faster on B200 (1000 x 64k bins:
        0.36 ms with all segments in flight, 0.45 ms with 256, 0.69 ms with 64)."""
        _require_cuda(iq_u8, torch.uint8, "iq_u8")
        if segments_in_flight is None:
            # nperseg = 2/4/8 * 8192 runs as one thread-block-cluster kernel that keeps each segment in
            # distributed shared memory and needs no spectra workspace at all; the library says which path this
            # plan and input pointer take (an unaligned view falls back to two passes and then needs the workspace)
            single_kernel = _lib.rmx_welch_path(self._h, _ptr(iq_u8)) == 1
            segments_in_flight = 1 if single_kernel else max(1, (1 << 30) // (8 * self.fft_len))
        if self.n_samples != self.fft_len:
            raise ValueError("Welch plans need n_samples == fft_len")
        if iq_u8.numel() != self.n_signals * 2 * self.fft_len:
            raise ValueError("iq_u8 has %d bytes, expected %d" % (iq_u8.numel(), self.n_signals * 2 * self.fft_len))
        psd = torch.empty(self.fft_len, dtype=torch.float32, device=self.device)
        nb = int(_lib.rmx_welch_workspace_bytes(self._h, max(1, min(self.n_signals, int(segments_in_flight)))))
        ws = self._get_workspace(nb)
        with torch.cuda.device(self.device):
            _native.check(_lib.rmx_welch_psd(self._h, _ptr(iq_u8), _ptr(psd), float(sample_rate), _ptr(ws), nb, _stream_ptr()),
                          "rmx_welch_psd")
        return psd


def peaks_to_numpy(records: torch.Tensor) -> np.ndarray:
    """int32[P, 4] device records -> host structured array (lag, peak, frac, search)."""
    host = records.detach().cpu().numpy()
    return host.view(np.dtype([("lag", "<i4"), ("peak", "<f4"), ("frac", "<f4"), ("search", "<f4")])).reshape(-1)


def power_db(p: torch.Tensor, eps: float = 1e-24) -> torch.Tensor:
    _require_cuda(p, torch.float32, "p")
    out = torch.empty_like(p)
    with torch.cuda.device(p.device):
        _native.check(_lib.rmx_power_db(_ptr(p), _ptr(out), p.numel(), float(eps), _stream_ptr()), "rmx_power_db")
    return out


def threshold_peaks(db: torch.Tensor, height: float, cap: Optional[int] = None) -> np.ndarray:
    """Sorted bin indices of local maxima (scipy `_local_maxima_1d` semantics) with db >= height."""
    _require_cuda(db, torch.float32, "db")
    n = db.numel()
    cap = n // 2 + 1 if cap is None else int(cap)
    idx = torch.empty(max(cap, 1), dtype=torch.int32, device=db.device)
    count = torch.zeros(1, dtype=torch.int32, device=db.device)
    with torch.cuda.device(db.device):
        _native.check(_lib.rmx_threshold_peaks(_ptr(db), n, float(height), _ptr(idx), _ptr(count), cap, _stream_ptr()),
                      "rmx_threshold_peaks")
    c = int(count.item())
    if c > cap:
        raise _native.RmxError("threshold_peaks: %d candidates exceed cap %d" % (c, cap))
    return np.sort(idx[:c].cpu().numpy())


_pinned_peaks: dict = {}


def find_peaks_batch(db: torch.Tensor, height: float, distance: int = 0, height_above_mean: bool = False, cap: int = 1024,
                     flat: bool = False, gate_dc_bins: int = 0, gate_conf_min: float = 0.0, bandwidth_drop_db: Optional[float] = None):
    """scipy.signal.find_peaks(row, height=, distance=) for every row of db[n_rows, n] in one launch, plus each
    row's mean and median.  height_above_mean=True uses mean(row) + height (signal_analyzer.py:75).
    Returns (peaks: list of int32 arrays (ascending bins), heights: list of float32 arrays, mean[n_rows],
    median[n_rows]) on the host; only the peak lists cross PCIe, not the spectra.  flat=True returns
    (bins, heights, offsets[n_rows + 1], mean, median) with row r in [offsets[r], offsets[r+1]).
    gate_dc_bins / gate_conf_min apply the buoy detector's gates on the device (see include/rmx.h).
    bandwidth_drop_db (flat=True only) appends the per-peak width in bins of the -drop_db walk of
    iq_stream_client.py:254-278 as a sixth return value."""
    _require_cuda(db, torch.float32, "db")
    if db.ndim != 2:
        raise ValueError("db must be [n_rows, n]")
    n_rows, n = db.shape
    if db.stride(1) != 1:
        db = db.contiguous()
    cap = max(1, int(cap))
    dev = db.device
    idx = torch.empty((n_rows, cap), dtype=torch.int32, device=dev)
    hts = torch.empty((n_rows, cap), dtype=torch.float32, device=dev)
    count = torch.empty(n_rows, dtype=torch.int32, device=dev)
    stats = torch.empty((n_rows, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _native.check(_lib.rmx_find_peaks_batch(_ptr(db), n_rows, n, db.stride(0), float(height), int(bool(height_above_mean)),
                                                int(np.ceil(distance)) if distance else 0, int(gate_dc_bins), float(gate_conf_min),
                                                _ptr(idx), _ptr(hts), _ptr(count), cap, _ptr(stats), _stream_ptr()),
                      "rmx_find_peaks_batch")
    width = None
    if bandwidth_drop_db is not None:
        if not flat:
            raise ValueError("bandwidth_drop_db needs flat=True")
        width = torch.empty((n_rows, cap), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_lib.rmx_peak_bandwidth_batch(_ptr(db), n_rows, n, db.stride(0), _ptr(idx), _ptr(count), cap,
                                                        float(bandwidth_drop_db), _ptr(width), _stream_ptr()),
                          "rmx_peak_bandwidth_batch")
    c = count.cpu().numpy()
    st = stats.cpu().numpy()
    most = int(min(cap, max(0, c.max(initial=0))))
    if most:
        # page-locked staging (cached per shape): the peak lists are tens of MB for a large batch
        key = (n_rows, cap)
        stage = _pinned_peaks.get(key)
        if stage is None:
            if len(_pinned_peaks) > 4:
                _pinned_peaks.clear()
            stage = _pinned_peaks[key] = (torch.empty((n_rows, cap), dtype=torch.int32).pin_memory(),
                                          torch.empty((n_rows, cap), dtype=torch.float32).pin_memory())
        with torch.cuda.device(dev):
            stage[0][:, :most].copy_(idx[:, :most], non_blocking=True)
            stage[1][:, :most].copy_(hts[:, :most], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        idx_h, hts_h = stage[0][:, :most].numpy(), stage[1][:, :most].numpy()
    else:
        idx_h, hts_h = np.empty((n_rows, 0), np.int32), np.empty((n_rows, 0), np.float32)
    ok = (c >= 0) & (c <= cap)
    if flat:
        # one boolean mask over the [n_rows, most] block instead of a Python loop over rows
        cc = np.where(ok, c, 0)
        mask = np.arange(most)[None, :] < cc[:, None]
        flat_bins, flat_h = idx_h[mask], hts_h[mask]
        offsets = np.concatenate([[0], np.cumsum(cc)])
        if not ok.all():
            raise _native.RmxError("find_peaks_batch(flat=True): %d rows exceed the on-chip candidate limit or cap"
                                   % int((~ok).sum()))
        if width is not None:
            wid = width[:, :most].cpu().numpy()[mask] if most else np.empty(0, np.int32)
            return flat_bins, flat_h, offsets, st[:, 0].copy(), st[:, 1].copy(), wid
        return flat_bins, flat_h, offsets, st[:, 0].copy(), st[:, 1].copy()
    peaks, heights = [], []
    for r in range(n_rows):
        if not ok[r]:
            # more candidates than the kernel keeps on chip (or more peaks than `cap`): this row alone goes
            # through the single-row entry points
            thr = float(st[r, 0]) + float(height) if height_above_mean else float(height)
            cand = threshold_peaks(db[r], thr)
            row = db[r].cpu().numpy()
            kept = select_by_distance(cand, row[cand], distance) if distance else cand
            if gate_dc_bins > 0 or gate_conf_min > 0:
                conf = np.clip((row[kept] - np.float32(st[r, 1])).astype(np.float32) / np.float32(20.0), 0.0, 1.0)
                kept = kept[(np.minimum(kept, n - kept) >= gate_dc_bins) & (conf >= np.float32(gate_conf_min))]
            peaks.append(kept.astype(np.int32)); heights.append(row[kept])
        else:
            peaks.append(idx_h[r, :c[r]].copy()); heights.append(hts_h[r, :c[r]].copy())
    return peaks, heights, st[:, 0].copy(), st[:, 1].copy()


def select_by_distance(positions: np.ndarray, heights: np.ndarray, distance: int) -> np.ndarray:
    """find_peaks(distance=) greedy rule on host arrays; returns the kept positions."""
    positions = np.ascontiguousarray(positions, dtype=np.int32)
    heights = np.ascontiguousarray(heights, dtype=np.float32)
    keep = np.ones(len(positions), dtype=np.uint8)
    if len(positions):
        _native.check(_lib.rmx_select_by_distance_host(positions.ctypes.data_as(ctypes.c_void_p),
                                                       heights.ctypes.data_as(ctypes.c_void_p), len(positions),
                                                       int(np.ceil(distance)), keep.ctypes.data_as(ctypes.c_void_p)),
                      "rmx_select_by_distance_host")
    return positions[keep.astype(bool)]


def mean_median(db: torch.Tensor):
    """(mean, median) of a float32 device vector as Python floats."""
    _require_cuda(db, torch.float32, "db")
    out = torch.empty(2, dtype=torch.float32, device=db.device)
    with torch.cuda.device(db.device):
        _native.check(_lib.rmx_mean_median(_ptr(db), db.numel(), _ptr(out), None, 0, _stream_ptr()), "rmx_mean_median")
    m = out.cpu().numpy()
    return np.float32(m[0]), np.float32(m[1])


def signal_stats_c64(x: torch.Tensor):
    """(mean |x|^2 as float64, max |x| as float32) of complex64 device samples."""
    _require_cuda(x, torch.complex64, "x")
    out = torch.empty(32, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _native.check(_lib.rmx_signal_stats_c64(_ptr(x), x.numel(), _ptr(out), ctypes.c_void_p(out.data_ptr() + 16),
                                                _stream_ptr()), "rmx_signal_stats_c64")
    raw = out.cpu().numpy()
    return float(raw[:8].view(np.float64)[0]), np.float32(raw[8:12].view(np.float32)[0])


def signal_stats(iq_u8: torch.Tensor):
    """(mean |x|^2 as float64, max |x| as float32) straight from the cu8 bytes."""
    _require_cuda(iq_u8, torch.uint8, "iq_u8")
    out = torch.empty(16, dtype=torch.uint8, device=iq_u8.device)
    with torch.cuda.device(iq_u8.device):
        _native.check(_lib.rmx_signal_stats(_ptr(iq_u8), iq_u8.numel() // 2, _ptr(out), _stream_ptr()), "rmx_signal_stats")
    raw = out.cpu().numpy()
    return float(raw[:8].view(np.float64)[0]), np.float32(raw[8:12].view(np.float32)[0])


def is_pow2(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


def spectrum_db_c64(x: torch.Tensor, shift: bool = False, plans: Optional[dict] = None) -> torch.Tensor:
    """20*log10(|fft(x)| + 1e-12) of one complex64 device vector, natural (or fftshifted) order.
    Power-of-two lengths use the tile FFT directly; other lengths go through `bluestein`."""
    _require_cuda(x, torch.complex64, "x")
    n = x.numel()
    if n < 16 or not is_pow2(n):
        from . import bluestein
        return bluestein.spectrum_db(x.reshape(-1), shift=shift, plans=plans)
    key = (1, n, n)
    plan = plans.get(key) if plans is not None else None
    if plan is None:
        plan = Plan(1, n, n, device=x.device)
        if plans is not None:
            plans[key] = plan
    return plan.spectrum_db(plan.forward_c64(x.reshape(1, n)), shift=shift)[0]
