"""GPU-backed equivalents of the reference's two block detectors.

  * `BuoySignalDetector.detect_block`   <- buoy_node.py:391-455  (SignalDetector._detect_real_signals
                                          after the rtl_sdr capture)
  * `StreamSignalDetector.detect_signals` <- iq_stream_client.py:181-252 (SignalDetector.detect_signals)
  * `unpack_iq_samples`                 <- iq_stream_client.py:148-159 (RealTimeSDRCapture.read_iq_samples
                                          after the pipe read)

The FFT, dB spectrum, local-maximum / threshold scan and the median noise floor run on the
GPU (librmx); the greedy `distance=10` rule and the per-peak bookkeeping (a few hundred
peaks) are host code, as SURVEY §7 prescribes.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from datetime import datetime, timezone
from typing import List, Optional

import numpy as np

from .tdoa_processor import SignalDetection


def _engine():
    from . import engine
    return engine


def unpack_iq_samples(raw_bytes) -> np.ndarray:
    """bytes / uint8 array of interleaved I,Q -> complex64 numpy array (GPU unpack, bit-exact)."""
    import torch
    eng = _engine()
    raw = np.frombuffer(raw_bytes, dtype=np.uint8) if not isinstance(raw_bytes, np.ndarray) else raw_bytes
    n = raw.size // 2
    dev = eng.unpack_cu8(torch.from_numpy(np.ascontiguousarray(raw[: 2 * n])).cuda())
    return dev.cpu().numpy()


def classify_buoy(frequency: float) -> str:
    """Band table of buoy_node.py:342-355 (MHz)."""
    if frequency in (121.5, 243.0):
        return "emergency"
    for lo, hi, name in ((118.0, 136.0, "aviation"), (144.0, 148.0, "amateur"), (156.0, 162.0, "marine"),
                         (406.0, 406.1, "emergency_beacon")):
        if lo <= frequency <= hi:
            return name
    return "unknown"


def classify_stream(frequency_hz: float) -> str:
    """Band table of iq_stream_client.py:280-304 (Hz)."""
    if abs(frequency_hz - 121500000) < 1000 or abs(frequency_hz - 243000000) < 1000:
        return "emergency"
    table = ((155000000, 156000000, "public_safety"), (406000000, 406100000, "emergency"),
             (88000000, 108000000, "fm_radio"), (118000000, 136000000, "aviation"),
             (144000000, 148000000, "amateur"), (420000000, 450000000, "amateur"))
    for lo, hi, name in table:
        if lo <= frequency_hz <= hi:
            return name
    return "unknown"


class _BlockSpectrum:
    """dB spectrum of one block on the device + the scalars/arrays the detectors need."""

    def __init__(self, db_dev, height, distance):
        eng = _engine()
        self.db_dev = db_dev
        _, self.median = eng.mean_median(db_dev)
        cand = eng.threshold_peaks(db_dev, float(height))
        self.db = db_dev.cpu().numpy()
        self.peaks = eng.select_by_distance(cand, self.db[cand], distance) if distance else cand


def _spectrum_from_cu8(iq_u8, plans):
    import torch
    eng = _engine()
    t = torch.as_tensor(iq_u8).reshape(1, -1).cuda()
    n = t.shape[1] // 2
    if eng.is_pow2(n) and n >= 16:
        key = (1, n, n)
        plan = plans.get(key)
        if plan is None:
            plan = plans[key] = eng.Plan(1, n, n)
        return plan.spectrum_db(plan.forward(t))[0], n
    return eng.spectrum_db_c64(eng.unpack_cu8(t).reshape(-1), plans=plans), n


class BuoySignalDetector:
    """buoy_node.SignalDetector's real-signal path on raw cu8 blocks."""

    def __init__(self, buoy_id: str, lat: float = 0.0, lng: float = 0.0, sample_rate: int = 2048000):
        self.buoy_id = buoy_id
        self.lat, self.lng = lat, lng
        self.sample_rate = sample_rate
        self.detection_threshold_dbm = -70           # buoy_node.py:141
        self._plans = {}

    def detect_block(self, iq_u8, center_freq_mhz: float, iso_timestamp: Optional[str] = None,
                     gps_ns: Optional[int] = None, timestamps=None) -> List[SignalDetection]:
        """Detections of one raw cu8 block, in increasing bin order (like the reference loop)."""
        return self.detect_block_indexed(iq_u8, center_freq_mhz, iso_timestamp, gps_ns, timestamps)[1]

    def detect_block_indexed(self, iq_u8, center_freq_mhz: float, iso_timestamp: Optional[str] = None,
                             gps_ns: Optional[int] = None, timestamps=None):
        """-> (FFT bin of each detection, detections).  `timestamps`: optional callable returning
        (iso_timestamp, gps_ns), called once per detection like the reference's
        gps_source.get_precise_timestamp() (buoy_node.py:436)."""
        center_freq_hz = int(center_freq_mhz * 1e6)                       # :365
        db_dev, n = _spectrum_from_cu8(iq_u8, self._plans)
        spec = _BlockSpectrum(db_dev, self.detection_threshold_dbm, 10)   # :411-415
        abs_freqs = np.fft.fftfreq(n, 1.0 / self.sample_rate) + center_freq_hz   # :402,408
        if iso_timestamp is None:
            iso_timestamp = datetime.now(timezone.utc).isoformat()
        if gps_ns is None:
            gps_ns = time.time_ns()
        out: List[SignalDetection] = []
        bins: List[int] = []
        for k in spec.peaks:
            f_hz = abs_freqs[k]
            if abs(f_hz - center_freq_hz) < 10000:                        # :423
                continue
            power = spec.db[k]
            confidence = min(max((power - spec.median) / 20.0, 0.0), 1.0)  # :427-429
            if confidence < 0.3:                                          # :432
                continue
            f_mhz = f_hz / 1e6
            bins.append(int(k))
            if timestamps is not None:
                iso_timestamp, gps_ns = timestamps()
            out.append(SignalDetection(buoy_id=self.buoy_id, frequency_mhz=round(f_mhz, 3),
                                       signal_strength_dbm=round(power, 1), timestamp_utc=iso_timestamp,
                                       gps_timestamp_ns=gps_ns, lat=self.lat, lng=self.lng,
                                       confidence=round(confidence, 2), signal_type=classify_buoy(f_mhz)))
        return bins, out

    def detect_blocks_arrays(self, iq_u8, center_freq_mhz: float):
        """Numeric half of `detect_blocks`: for uint8[n_blocks, 2N] (N a power of two >= 16; host or CUDA)
        returns, per block, (bins int32[], frequency_hz float64[], power_db float32[], confidence float32[]) of
        the detections that pass the reference's gates (|f - fc| >= 10 kHz, confidence >= 0.3;
        buoy_node.py:423-433), in increasing bin order.  One batched forward FFT, one dB pass and one
        find_peaks/median launch cover all blocks; only the peak lists come back to the host."""
        import torch
        eng = _engine()
        t = torch.as_tensor(iq_u8)
        if t.ndim != 2 or t.dtype != torch.uint8 or t.shape[1] % 2:
            raise ValueError("iq_u8 must be uint8[n_blocks, 2N]")
        nb, n = t.shape[0], t.shape[1] // 2
        if not (eng.is_pow2(n) and n >= 16):
            raise ValueError("detect_blocks_arrays needs a power-of-two block length >= 16 (use detect_block)")
        key = (nb, n, n)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = eng.Plan(nb, n, n)
        db = plan.spectrum_db(plan.forward(t.cuda(non_blocking=True)))         # [nb, n] dB, natural order
        center_freq_hz = int(center_freq_mhz * 1e6)                            # :365
        abs_freqs = np.fft.fftfreq(n, 1.0 / self.sample_rate) + center_freq_hz   # :402,408
        # The |f - fc| < 10 kHz gate (:423) is a symmetric band of bins around DC; when the float64 mask really has
        # that shape both gates run on the device and only the detections come back.
        dc_mask = np.abs(abs_freqs - center_freq_hz) < 10000
        ar = np.arange(n)
        dc_bins = int(dc_mask[:n // 2].sum())
        on_device = bool(np.array_equal(dc_mask, np.minimum(ar, n - ar) < dc_bins))
        kw = dict(gate_dc_bins=dc_bins, gate_conf_min=0.3) if on_device else {}
        try:
            bins, power, off, _, medians = eng.find_peaks_batch(db, self.detection_threshold_dbm, 10,   # :411-415, :427
                                                                cap=n // 10 + 2, flat=True, **kw)
        except eng._native.RmxError:
            peaks, heights, _, medians = eng.find_peaks_batch(db, self.detection_threshold_dbm, 10, cap=n // 10 + 2, **kw)
            bins = np.concatenate(peaks) if peaks else np.empty(0, np.int32)
            power = np.concatenate(heights) if heights else np.empty(0, np.float32)
            off = np.concatenate([[0], np.cumsum([len(p) for p in peaks])])
        # same float32 arithmetic as the per-block loop (:427-429)
        f_hz = abs_freqs[bins]
        med = np.repeat(medians, np.diff(off))
        conf = np.minimum(np.maximum((power - med) / np.float32(20.0), np.float32(0.0)), np.float32(1.0))
        if on_device:
            cuts = off[1:-1]
            return list(zip(np.split(bins, cuts), np.split(f_hz, cuts), np.split(power, cuts), np.split(conf, cuts)))
        keep = (~dc_mask[bins]) & (conf >= np.float32(0.3))
        block_of = np.repeat(np.arange(nb), np.diff(off))
        cuts = np.cumsum(np.bincount(block_of[keep], minlength=nb))[:-1]
        return list(zip(np.split(bins[keep], cuts), np.split(f_hz[keep], cuts), np.split(power[keep], cuts),
                        np.split(conf[keep], cuts)))

    def detect_blocks(self, iq_u8, center_freq_mhz: float, iso_timestamps: Optional[List[str]] = None,
                      gps_ns: Optional[List[int]] = None) -> List[List[SignalDetection]]:
        """`detect_block` for many raw cu8 blocks of one length at once (see `detect_blocks_arrays`)."""
        import torch
        eng = _engine()
        t = torch.as_tensor(iq_u8)
        nb = t.shape[0]
        n = t.shape[1] // 2 if t.ndim == 2 else 0
        if t.ndim != 2 or not (eng.is_pow2(n) and n >= 16):
            return [self.detect_block(t[b], center_freq_mhz, iso_timestamps[b] if iso_timestamps else None,
                                      gps_ns[b] if gps_ns else None) for b in range(nb)]
        results: List[List[SignalDetection]] = []
        for b, (bins, f_hz, power, conf) in enumerate(self.detect_blocks_arrays(t, center_freq_mhz)):
            stamp = iso_timestamps[b] if iso_timestamps else datetime.now(timezone.utc).isoformat()
            ns = gps_ns[b] if gps_ns else time.time_ns()
            results.append([SignalDetection(buoy_id=self.buoy_id, frequency_mhz=round(f / 1e6, 3),
                                            signal_strength_dbm=round(pw, 1), timestamp_utc=stamp,
                                            gps_timestamp_ns=ns, lat=self.lat, lng=self.lng,
                                            confidence=round(c, 2), signal_type=classify_buoy(f / 1e6))
                            for f, pw, c in zip(f_hz, power, conf)])
        return results


@dataclass
class StreamDetection:
    """Fields of iq_stream_client.SignalDetection (:46-60)."""
    node_id: str
    frequency_mhz: float
    signal_strength_dbm: float
    bandwidth_hz: float
    timestamp_utc: str
    gps_timestamp_ns: int
    lat: float
    lng: float
    confidence: float
    signal_type: str
    iq_samples: Optional[list] = None
    detection_method: str = "power_threshold"


class StreamSignalDetector:
    """iq_stream_client.SignalDetector on complex64 blocks."""

    def __init__(self, node_id: str, sample_rate: int = 2048000):
        self.node_id = node_id
        self.sample_rate = sample_rate
        self.detection_threshold = -70
        self.lat, self.lng = 35.4676, -97.5164          # :173-174
        self._plans = {}

    def _estimate_bandwidth(self, power_spectrum_db: np.ndarray, peak_idx: int) -> float:
        """-3 dB walk left and right of the peak (:254-278)."""
        p_db = power_spectrum_db
        thr = p_db[peak_idx] - 3.0
        left = right = int(peak_idx)
        last = len(p_db) - 1
        while left > 0 and p_db[left] > thr:
            left -= 1
        while right < last and p_db[right] > thr:
            right += 1
        return (right - left) * (self.sample_rate / len(p_db))

    def detect_signals(self, iq_samples: np.ndarray, center_freq_hz: float) -> List[StreamDetection]:
        return self.detect_signals_indexed(iq_samples, center_freq_hz)[1]

    def detect_blocks_arrays(self, iq_u8, center_freq_hz: float):
        """`detect_signals` for many raw cu8 blocks at once: iq_u8 is uint8[n_blocks, 2N] as read from the
        rtl_sdr pipe (`read_iq_samples` unpacks exactly these bytes, :148-159), N a power of two >= 16.  Returns per
        block (bins, frequency_hz, power_db, bandwidth_hz, confidence) arrays: FFT, dB, find_peaks(height=-70,
        distance=10), median and the -3 dB bandwidth walk (:254-278) all run on the device."""
        import torch
        eng = _engine()
        t = torch.as_tensor(iq_u8)
        if t.ndim != 2 or t.dtype != torch.uint8 or t.shape[1] % 2:
            raise ValueError("iq_u8 must be uint8[n_blocks, 2N]")
        nb, n = t.shape[0], t.shape[1] // 2
        if not (eng.is_pow2(n) and n >= 16):
            raise ValueError("detect_blocks_arrays needs a power-of-two block length >= 16")
        key = (nb, n, n)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = eng.Plan(nb, n, n)
        db = plan.spectrum_db(plan.forward(t.cuda(non_blocking=True)))
        bins, power, off, _, medians, width = eng.find_peaks_batch(db, self.detection_threshold, 10, cap=n // 10 + 2,
                                                                   flat=True, bandwidth_drop_db=3.0)
        abs_freqs = np.fft.fftfreq(n, 1.0 / self.sample_rate) + center_freq_hz
        f_hz = abs_freqs[bins]
        med = np.repeat(medians, np.diff(off))
        conf = np.minimum((power - med) / np.float32(20.0), np.float32(1.0))          # :215-217, no lower clamp
        bw = width * (self.sample_rate / n)
        cuts = off[1:-1]
        return list(zip(np.split(bins, cuts), np.split(f_hz, cuts), np.split(power, cuts), np.split(bw, cuts),
                        np.split(conf, cuts)))

    def detect_signals_indexed(self, iq_samples: np.ndarray, center_freq_hz: float):
        """-> (FFT bin of each detection, detections)."""
        import torch
        eng = _engine()
        x = np.ascontiguousarray(iq_samples, dtype=np.complex64)
        n = x.size
        db_dev = eng.spectrum_db_c64(torch.from_numpy(x).cuda(), plans=self._plans)
        spec = _BlockSpectrum(db_dev, self.detection_threshold, 10)       # :197-201
        abs_freqs = np.fft.fftfreq(n, 1.0 / self.sample_rate) + center_freq_hz
        out: List[StreamDetection] = []
        for k in spec.peaks:
            f_hz = abs_freqs[k]
            power = spec.db[k]
            start = max(0, int(k) - 128)
            snippet = x[start:min(n, start + 256)]                         # :306-316
            out.append(StreamDetection(
                node_id=self.node_id, frequency_mhz=f_hz / 1e6, signal_strength_dbm=power,
                bandwidth_hz=self._estimate_bandwidth(spec.db, k), timestamp_utc=datetime.now(timezone.utc).isoformat(),
                gps_timestamp_ns=time.time_ns(), lat=self.lat, lng=self.lng,
                confidence=min((power - spec.median) / 20.0, 1.0),           # :215-217
                signal_type=classify_stream(f_hz), iq_samples=snippet.tolist(), detection_method="fft_peak"))
        return [int(k) for k in spec.peaks], out
