"""Binary `cu8_window` message: raw IQ windows from the buoys to the central processor.

The reference's buoys send only JSON `signal_detection` text frames (reference central_processor.py:305-335,
buoy_node.py CentralCommunicator), so the central `TDoAProcessor` never sees samples and can only subtract
timestamps (tdoa_processor.py:166).  This module defines the frame that carries the samples themselves next to
those text frames — WebSocket frames are typed, so a BINARY frame is a cu8 window and a TEXT frame stays the
reference's JSON — and the assembler that turns the frames of B buoys into the uint8[B, 1, 2N] block
`TDoAProcessor.correlate_iq` consumes.  Host-only code: no arithmetic on the samples happens here.

Frame layout (little endian, 56-byte fixed header, then the buoy id, then the payload):

    off  size  field
      0     4  magic  b"RMXW"
      4     2  version (1)
      6     2  header_len (56 + len(buoy_id)): payload offset
      8     8  window_index            u64   windows of one capture are numbered from 0
     16     8  gps_timestamp_ns        i64   GPS time of the first sample (buoy_node.py:108-118 timestamp pair)
     24     8  sample_rate_hz          f64
     32     8  center_freq_hz          f64
     40     4  n_samples               u32   complex samples; payload is 2*n_samples bytes of rtl_sdr cu8 (I,Q,I,Q...)
     44     4  payload_crc32           u32   zlib.crc32 of the payload
     48     2  buoy_id_len             u16
     50     6  reserved (0)
     56     -  buoy_id (utf-8), payload
"""
from __future__ import annotations

import json
import struct
import zlib
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

MAGIC = b"RMXW"
VERSION = 1
_HEADER = struct.Struct("<4sHHQqddIIH6x")
assert _HEADER.size == 56


class WireError(ValueError):
    """Malformed or corrupted cu8_window frame."""


@dataclass
class Cu8Window:
    buoy_id: str
    window_index: int
    gps_timestamp_ns: int
    sample_rate_hz: float
    center_freq_hz: float
    iq_u8: np.ndarray            # uint8[2*n_samples], a zero-copy view of the frame

    @property
    def n_samples(self) -> int:
        return self.iq_u8.size // 2


def pack_cu8_window(buoy_id: str, window_index: int, gps_timestamp_ns: int, sample_rate_hz: float,
                    center_freq_hz: float, iq_u8) -> bytes:
    """One binary frame for `iq_u8` (uint8, interleaved I,Q: exactly what `rtl_sdr ... -` writes)."""
    payload = np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1)
    if payload.size % 2:
        raise WireError("cu8 payload must hold an even number of bytes (I,Q pairs)")
    name = buoy_id.encode("utf-8")
    if len(name) > 0xFFFF - _HEADER.size:
        raise WireError("buoy_id too long")
    raw = payload.tobytes()
    head = _HEADER.pack(MAGIC, VERSION, _HEADER.size + len(name), int(window_index), int(gps_timestamp_ns),
                        float(sample_rate_hz), float(center_freq_hz), payload.size // 2, zlib.crc32(raw) & 0xFFFFFFFF,
                        len(name))
    return head + name + raw


def unpack_cu8_window(frame: Union[bytes, bytearray, memoryview], verify: bool = True) -> Cu8Window:
    buf = memoryview(frame)
    if len(buf) < _HEADER.size:
        raise WireError("frame shorter than the header (%d bytes)" % len(buf))
    magic, version, header_len, widx, gps_ns, fs, fc, n, crc, name_len = _HEADER.unpack_from(buf, 0)
    if magic != MAGIC:
        raise WireError("not a cu8_window frame (magic %r)" % bytes(magic))
    if version != VERSION:
        raise WireError("unsupported cu8_window version %d" % version)
    if header_len != _HEADER.size + name_len or len(buf) != header_len + 2 * n:
        raise WireError("frame length %d does not match its header (header %d + payload %d)" % (len(buf), header_len, 2 * n))
    payload = np.frombuffer(buf, dtype=np.uint8, count=2 * n, offset=header_len)
    if verify and (zlib.crc32(payload) & 0xFFFFFFFF) != crc:
        raise WireError("payload checksum mismatch")
    return Cu8Window(bytes(buf[_HEADER.size:header_len]).decode("utf-8"), widx, gps_ns, fs, fc, payload)


def dispatch(frame) -> Tuple[str, object]:
    """Classify one WebSocket frame of the buoy link: TEXT frames are the reference's JSON messages
    (`signal_detection`, `heartbeat`, ... central_processor.py:270-345) and are returned as (type, dict); BINARY
    frames are cu8 windows and are returned as ("cu8_window", Cu8Window)."""
    if isinstance(frame, str):
        data = json.loads(frame)
        return str(data.get("type", "")), data
    return "cu8_window", unpack_cu8_window(frame)


class WindowAssembler:
    """Collects cu8_window frames of a fixed set of buoys and releases a window once every buoy has delivered it.

    Windows land in a page-locked ring uint8[depth, B, 2N] (pinned when torch + CUDA are available, so the block
    goes to the GPU by plain async DMA); a released window is valid until `depth - 1` later windows were released.
    Windows older than the ring (a buoy that fell `depth` windows behind) are dropped and counted in `dropped`."""

    def __init__(self, buoy_ids: Sequence[str], samples_per_window: int, depth: int = 4, pinned: bool = True):
        self.buoy_ids = list(buoy_ids)
        self._index = {b: k for k, b in enumerate(self.buoy_ids)}
        if len(self._index) != len(self.buoy_ids):
            raise ValueError("duplicate buoy ids")
        self.samples_per_window = int(samples_per_window)
        self.depth = int(depth)
        shape = (self.depth, len(self.buoy_ids), 2 * self.samples_per_window)
        self._torch_ring = None
        if pinned:
            try:
                import torch
                self._torch_ring = torch.empty(shape, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
                self._ring = self._torch_ring.numpy()
            except Exception:
                self._torch_ring = None
        if self._torch_ring is None:
            self._ring = np.empty(shape, dtype=np.uint8)
        self._have: Dict[int, set] = {}
        self._stamps: Dict[int, List[int]] = {}
        self._meta: Dict[int, Tuple[float, float]] = {}
        self.released = -1           # highest window index handed out
        self._done: set = set()      # recently released windows (late duplicates of these are dropped)
        self._newest = -1            # highest window index seen so far
        self.dropped = 0

    def add(self, msg: Cu8Window) -> Optional[int]:
        """Store one frame; returns the window index if this frame completed a window, else None."""
        if msg.buoy_id not in self._index:
            raise WireError("frame from unknown buoy %r" % msg.buoy_id)
        if msg.n_samples != self.samples_per_window:
            raise WireError("frame holds %d samples, expected %d" % (msg.n_samples, self.samples_per_window))
        w = msg.window_index
        self._newest = max(self._newest, w)
        # windows complete in any order (frames of different windows interleave freely); a frame is dropped only
        # if its window was already handed out or has fallen out of the ring
        if w in self._done or w <= self._newest - self.depth:
            self.dropped += 1
            return None
        # a window that pushes the ring forward evicts the incomplete windows it overwrites
        for old in [k for k in self._have if k <= w - self.depth]:
            del self._have[old], self._stamps[old], self._meta[old]
            self.dropped += 1
        got = self._have.setdefault(w, set())
        meta = self._meta.setdefault(w, (msg.sample_rate_hz, msg.center_freq_hz))
        if meta != (msg.sample_rate_hz, msg.center_freq_hz):
            raise WireError("buoys disagree on sample rate / centre frequency for window %d" % w)
        b = self._index[msg.buoy_id]
        self._ring[w % self.depth, b] = msg.iq_u8
        self._stamps.setdefault(w, [0] * len(self.buoy_ids))[b] = int(msg.gps_timestamp_ns)
        got.add(b)
        return w if len(got) == len(self.buoy_ids) else None

    def take(self, w: int):
        """-> (block uint8[B, 1, 2N] (torch tensor when available, else numpy), gps_timestamp_ns per buoy,
        sample_rate_hz, center_freq_hz) of a completed window."""
        if len(self._have.get(w, ())) != len(self.buoy_ids):
            raise KeyError("window %d is not complete" % w)
        stamps = self._stamps.pop(w)
        fs, fc = self._meta.pop(w)
        del self._have[w]
        self.released = max(self.released, w)
        self._done.add(w)
        self._done = {k for k in self._done if k > self._newest - 4 * self.depth}
        slot = w % self.depth
        block = self._torch_ring[slot] if self._torch_ring is not None else self._ring[slot]
        return block[:, None, :], stamps, fs, fc

    def feed(self, frames) -> Iterator[Tuple[int, object, List[int], float, float]]:
        """Convenience: iterate over binary frames, yielding (window_index, block, stamps, fs, fc) per completed window."""
        for frame in frames:
            w = self.add(frame if isinstance(frame, Cu8Window) else unpack_cu8_window(frame))
            if w is not None:
                yield (w,) + self.take(w)
