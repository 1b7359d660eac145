"""Drop-in for the hot path of the reference's `buoy_node` module.

Only the block-detection path is mirrored: `SignalDetector._detect_real_signals(center_freq_mhz)`
(reference buoy_node.py:357-468) with its capture step, its error behaviour (fallback detection,
:461-468) and the types it returns.  The WebSocket transport, the scan scheduler and the CLI of
the reference module are control plane and out of scope.

The numeric part — cu8 unpack (:392-398), FFT (:401), dB (:405), find_peaks(height=-70,
distance=10) (:411-415), median / confidence / gates (:423-433) — runs on the GPU through
`detectors.BuoySignalDetector`; there is no CPU path for it.
"""
from __future__ import annotations

import logging
import random
import subprocess
import time
from dataclasses import dataclass
from datetime import datetime, timezone
from typing import Callable, List, Optional, Tuple

from .detectors import BuoySignalDetector, classify_buoy

logger = logging.getLogger(__name__)


@dataclass
class SignalDetection:
    """Fields of buoy_node.SignalDetection (:34-47)."""
    buoy_id: str
    frequency_mhz: float
    signal_strength_dbm: float
    timestamp_utc: str
    gps_timestamp_ns: int
    lat: float
    lng: float
    confidence: float
    signal_type: str = "unknown"
    iq_sample_file: Optional[str] = None
    correlation_id: Optional[str] = None


class GPSTimeSource:
    """The two calls the detector makes on the reference's GPSTimeSource (:100-122): a timestamp pair and the
    buoy position.  `initialize_gps` is device bring-up and is not mirrored; set lat/lng directly."""

    def __init__(self, development_mode: bool = False):
        self.gps_locked = False
        self.timing_accuracy_ns = 1000000
        self.lat = 0.0
        self.lng = 0.0
        self.last_gps_update = None
        self.development_mode = development_mode

    def get_precise_timestamp(self) -> Tuple[str, int]:
        return datetime.now(timezone.utc).isoformat(), int(time.time_ns())

    def get_position(self) -> Tuple[float, float]:
        return self.lat, self.lng


class CaptureError(RuntimeError):
    """rtl_sdr ran but exited non-zero (reference :383-386)."""


def rtl_sdr_capture(center_freq_hz: int, sample_rate: int, num_samples: int) -> bytes:
    """One blocking capture of `num_samples` IQ samples: `rtl_sdr -f F -s S -n 2N -` exactly as the reference
    builds it (:368-381).  Returns the raw interleaved cu8 bytes."""
    cmd = ["rtl_sdr", "-f", str(center_freq_hz), "-s", str(sample_rate), "-n", str(num_samples * 2), "-"]
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, stdin=subprocess.DEVNULL)
    raw, err = proc.communicate(timeout=5)
    if proc.returncode != 0:
        raise CaptureError("rtl_sdr exit code %s: %s" % (proc.returncode, err.decode() if err else "Unknown error"))
    return raw


class SignalDetector:
    """buoy_node.SignalDetector (:134): same constructor, attributes and `_detect_real_signals` signature.

    `capture` (new, optional) replaces the rtl_sdr subprocess: a callable
    (center_freq_hz, sample_rate, num_samples) -> bytes, e.g. a file reader or a test vector."""

    def __init__(self, buoy_id: str, gps_source: GPSTimeSource, development_mode: bool = False,
                 capture: Optional[Callable[[int, int, int], bytes]] = None):
        self.buoy_id = buoy_id
        self.gps_source = gps_source
        self.monitoring = False
        self.detection_threshold_dbm = -70
        self.latest_signal_timestamp: Optional[str] = None
        self.development_mode = development_mode
        self._capture = capture or rtl_sdr_capture
        self._gpu: Optional[BuoySignalDetector] = None

    def _classify_signal(self, frequency: float) -> str:
        return classify_buoy(frequency)

    def _detect_real_signals(self, center_freq_mhz: float) -> List[SignalDetection]:
        """Capture 16384 samples at 2.048 Msps and detect peaks (reference :357-468).  Capture failures take the
        reference's fallback path; a short read returns [] (:388-390)."""
        sample_rate, num_samples = 2048000, 16384                     # :362-364
        center_freq_hz = int(center_freq_mhz * 1e6)
        try:
            raw = self._capture(center_freq_hz, sample_rate, num_samples)
        except subprocess.TimeoutExpired:
            logger.error("SDR capture timeout - using fallback detection")
            return self._fallback_signal_detection(center_freq_mhz)
        except FileNotFoundError:
            logger.error("rtl_sdr command not found - using fallback detection")
            return self._fallback_signal_detection(center_freq_mhz)
        except Exception as exc:                                       # :466-468 (CaptureError lands here too)
            logger.error("SDR hardware inaccessible (%s) - using fallback detection", exc)
            return self._fallback_signal_detection(center_freq_mhz)
        if len(raw) < num_samples * 2:
            logger.warning("Incomplete SDR data: got %d bytes, expected %d", len(raw), num_samples * 2)
            return []
        import numpy as np
        import torch
        from . import engine  # noqa: F401  (ImportError if librmx.so is missing: the numeric path has no CPU fallback,
        if not torch.cuda.is_available():  # and a missing GPU must not be mistaken for missing SDR hardware)
            raise RuntimeError("radio_mapper_b200.buoy_node needs a CUDA device (no CPU fallback)")
        try:
            if self._gpu is None:
                self._gpu = BuoySignalDetector(self.buoy_id, sample_rate=sample_rate)
            self._gpu.detection_threshold_dbm = self.detection_threshold_dbm
            lat, lng = self.gps_source.get_position()
            self._gpu.lat, self._gpu.lng = lat, lng
            found = self._gpu.detect_block(np.frombuffer(raw, dtype=np.uint8), center_freq_mhz,
                                           timestamps=self.gps_source.get_precise_timestamp)
        except Exception as exc:
            logger.error("SDR hardware inaccessible (%s) - using fallback detection", exc)
            return self._fallback_signal_detection(center_freq_mhz)
        out = [SignalDetection(buoy_id=d.buoy_id, frequency_mhz=d.frequency_mhz, signal_strength_dbm=d.signal_strength_dbm,
                               timestamp_utc=d.timestamp_utc, gps_timestamp_ns=d.gps_timestamp_ns, lat=d.lat, lng=d.lng,
                               confidence=d.confidence, signal_type=d.signal_type) for d in found]
        logger.debug("Found %d signals at %s MHz", len(out), center_freq_mhz)
        return out

    def _fallback_signal_detection(self, center_freq_mhz: float) -> List[SignalDetection]:
        """The reference's behaviour without SDR hardware (:470-522): with probability 0.25 one synthetic
        detection at the centre frequency with a band-dependent strength."""
        if random.random() >= 0.25:
            return []
        bands = {121.5: (-85, -65, "emergency"), 243.0: (-90, -70, "emergency"),
                 105.7: (-50, -35, "commercial"), 101.9: (-55, -40, "commercial")}
        if center_freq_mhz in bands:
            lo, hi, signal_type = bands[center_freq_mhz]
        else:
            lo, hi, signal_type = -80, -50, self._classify_signal(center_freq_mhz)
        strength = random.uniform(lo, hi)
        stamp, gps_ns = self.gps_source.get_precise_timestamp()
        lat, lng = self.gps_source.get_position()
        confidence = min(max((strength + 90) / 40.0, 0.3), 0.95)
        return [SignalDetection(buoy_id=self.buoy_id, frequency_mhz=center_freq_mhz, signal_strength_dbm=round(strength, 1),
                                timestamp_utc=stamp, gps_timestamp_ns=gps_ns, lat=lat, lng=lng,
                                confidence=round(confidence, 2), signal_type=signal_type)]
