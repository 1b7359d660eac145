"""Drop-in for the hot path of the reference's `iq_stream_client` module.

Mirrored: `RealTimeSDRCapture.read_iq_samples(num_samples=8192)` (reference iq_stream_client.py:134-163) and
`SignalDetector.detect_signals(iq_samples, center_freq_hz)` (:181-252) with `_estimate_bandwidth` (:254-278),
`_classify_signal` (:280-304), `_extract_signal_samples` (:306-316) and the `SignalDetection` record (:46-60).
The WebSocket client, the request/history bookkeeping and the CLI are control plane and out of scope.

Unpack, FFT, dB, peak search and the median run on the GPU (`detectors`); there is no CPU path for them.
"""
from __future__ import annotations

import logging
import subprocess
from typing import List, Optional

import numpy as np

from .detectors import StreamDetection as SignalDetection
from .detectors import StreamSignalDetector, classify_stream, unpack_iq_samples

logger = logging.getLogger(__name__)


class RealTimeSDRCapture:
    """iq_stream_client.RealTimeSDRCapture (:72): the rtl_sdr pipe and the cu8 -> complex64 unpack."""

    def __init__(self, device_index: int = 0, sample_rate: int = 2048000):
        self.device_index = device_index
        self.sample_rate = sample_rate
        self.center_freq_hz = 100000000
        self.running = False
        self.capture_process = None
        self.detection_threshold_db = -70
        self.fft_size = 1024
        self.overlap = 0.5
        self.emergency_frequencies = [121500000, 243000000, 155160000, 406000000]

    def start_capture(self, center_freq_mhz: float = 100.0) -> bool:
        """Continuous `rtl_sdr -f F -s S -` into a pipe (:95-123)."""
        self.center_freq_hz = int(center_freq_mhz * 1e6)
        try:
            self.capture_process = subprocess.Popen(["rtl_sdr", "-f", str(self.center_freq_hz), "-s", str(self.sample_rate), "-"],
                                                    stdout=subprocess.PIPE, stderr=subprocess.PIPE, bufsize=0)
            self.running = True
            return True
        except Exception as exc:
            logger.error("Failed to start SDR capture: %s", exc)
            return False

    def stop_capture(self):
        self.running = False
        if self.capture_process:
            self.capture_process.terminate()
            self.capture_process.wait()
            self.capture_process = None

    def read_raw(self, num_samples: int = 8192) -> Optional[bytes]:
        """The pipe read of `read_iq_samples` alone: 2*num_samples raw cu8 bytes, or None (not running / short
        read).  Batched consumers (`SignalDetector.detect_blocks_arrays`, `ingest.Cu8PipeSource`) take these
        bytes to the GPU without the complex64 round trip through the host."""
        if not self.running or not self.capture_process:
            return None
        num_bytes = num_samples * 2
        raw = self.capture_process.stdout.read(num_bytes)
        if len(raw) != num_bytes:
            logger.warning("Incomplete read: got %d bytes, expected %d", len(raw), num_bytes)
            return None
        return raw

    def read_iq_samples(self, num_samples: int = 8192) -> Optional[np.ndarray]:
        """complex64[num_samples] from the rtl_sdr stream, or None (:134-163); the unpack runs on the GPU and is
        bit-exact with the reference's numpy expression (:149-157)."""
        raw = self.read_raw(num_samples)
        if raw is None:
            return None
        return unpack_iq_samples(raw)


class SignalDetector(StreamSignalDetector):
    """iq_stream_client.SignalDetector (:165): same constructor, attributes and method signatures."""

    def __init__(self, node_id: str, sample_rate: int = 2048000):
        super().__init__(node_id, sample_rate)
        self.signal_history: List[SignalDetection] = []
        self.max_history_size = 1000

    def detect_signals(self, iq_samples: np.ndarray, center_freq_hz: float) -> List[SignalDetection]:
        """Errors in the numeric path are logged and yield [] like the reference (:249-252) — except a missing
        CUDA library / device, which must fail loudly."""
        import torch
        from . import engine  # noqa: F401
        if not torch.cuda.is_available():
            raise RuntimeError("radio_mapper_b200.iq_stream_client needs a CUDA device (no CPU fallback)")
        try:
            return super().detect_signals(iq_samples, center_freq_hz)
        except Exception as exc:
            logger.error("Error in signal detection: %s", exc)
            return []

    def _classify_signal(self, frequency_hz: float) -> str:
        return classify_stream(frequency_hz)

    def _extract_signal_samples(self, iq_samples: np.ndarray, peak_idx: int, num_samples: int = 256) -> Optional[np.ndarray]:
        start = max(0, peak_idx - num_samples // 2)
        return iq_samples[start:min(len(iq_samples), start + num_samples)]
