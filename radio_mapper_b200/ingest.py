"""Streaming ingest of raw cu8 IQ for the correlator.

The reference's on-disk / on-wire format is headerless interleaved unsigned 8-bit I,Q
(`rtl_sdr` output: `Code/src/rtl_sdr.c:95`; capture files `iq_capture_<f>MHz_<ts>.bin` of exactly
2 bytes per sample, `sdr_capture.py:26,58`; the live pipe `rtl_sdr ... -` read in
`iq_stream_client.py:101-116,142`).  This module feeds that format to the GPU without staging the
whole capture:

    source  ->  pinned host ring (K windows)  --H2D on a copy stream-->  device ring  ->  Correlator

The copy of window w+1 overlaps the FFT / correlate kernels of window w; only the 16-byte peak
records come back.  Sources:

    Cu8FileSource   one headerless .bin capture per buoy (memory-mapped, no read-ahead copies)
    Cu8PipeSource   one readable binary stream per buoy (e.g. the stdout of `rtl_sdr ... -`)
    ArraySource     an in-memory uint8[B, n_bytes] array (tests, synthetic data)

`StreamingCorrelator.run(source)` yields one host record array [P] (RECORD_DTYPE) per window, in
order.  There is no CPU fallback: the arithmetic is `Correlator.run_device`.
"""
from __future__ import annotations

import os
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch

BYTES_PER_SAMPLE = 2        # cu8: one unsigned byte I, one unsigned byte Q (sdr_capture.py:58)

_pool = None


def _copy_rows(pairs):
    """dst[:] = src for every (dst, src) numpy pair, spread over a small thread pool: numpy releases the GIL
    in the copy loop, and one core moves ~10 GB/s where the DMA engine takes 50 GB/s."""
    global _pool
    pairs = list(pairs)
    if len(pairs) <= 1 or sum(d.nbytes for d, _ in pairs) < (8 << 20):
        for d, s_ in pairs:
            d[...] = s_
        return
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4), thread_name_prefix="rmx-ingest")

    def one(ds):
        ds[0][...] = ds[1]
    list(_pool.map(one, pairs))


class ArraySource:
    """uint8[B, n_bytes] held in memory; windows are consecutive, non-overlapping slices."""

    def __init__(self, iq_u8, samples_per_window: int):
        # a pinned torch tensor is used in place: its windows are DMA'd straight from the caller's memory
        self._pinned = iq_u8 if isinstance(iq_u8, torch.Tensor) and iq_u8.is_pinned() else None
        iq_u8 = iq_u8.numpy() if isinstance(iq_u8, torch.Tensor) else np.asarray(iq_u8)
        if iq_u8.dtype != np.uint8 or iq_u8.ndim != 2:
            raise TypeError("expected uint8[B, n_bytes]")
        self._data = iq_u8
        self.n_buoys = iq_u8.shape[0]
        self.samples_per_window = int(samples_per_window)
        self.n_windows = iq_u8.shape[1] // (BYTES_PER_SAMPLE * self.samples_per_window)

    def read_window(self, w: int, out: np.ndarray) -> bool:
        """Fill out[B, 2N] with window w; False when the stream is exhausted."""
        if w >= self.n_windows:
            return False
        nb = BYTES_PER_SAMPLE * self.samples_per_window
        _copy_rows((out[b], self._data[b, w * nb:(w + 1) * nb]) for b in range(self.n_buoys))
        return True

    def pinned_window(self, w: int):
        """uint8[B, 2N] view of window w in page-locked memory, or None (then read_window fills a ring slot)."""
        if self._pinned is None or w >= self.n_windows:
            return None
        nb = BYTES_PER_SAMPLE * self.samples_per_window
        return self._pinned[:, w * nb:(w + 1) * nb]

    def close(self):
        pass


class Cu8FileSource:
    """One raw capture file per buoy (the format `sdr_capture.capture_iq_data` writes).  Files are
    memory-mapped; the stream ends at the shortest file's last complete window.  Odd trailing bytes
    (half a sample) are ignored, like `load_iq_data` does (`signal_analyzer.py:28-36`)."""

    def __init__(self, paths: Sequence[str], samples_per_window: int, offset_samples: int = 0):
        if len(paths) < 1:
            raise ValueError("need at least one capture file")
        self.paths = list(paths)
        self.samples_per_window = int(samples_per_window)
        if self.samples_per_window <= 0:
            raise ValueError("samples_per_window must be positive")
        self._maps = []
        n_samples = None
        for p in self.paths:
            size = os.path.getsize(p)
            if size < BYTES_PER_SAMPLE:
                raise ValueError("%s holds no complete sample" % p)
            m = np.memmap(p, dtype=np.uint8, mode="r")
            self._maps.append(m)
            s = size // BYTES_PER_SAMPLE - int(offset_samples)
            n_samples = s if n_samples is None else min(n_samples, s)
        self._offset = BYTES_PER_SAMPLE * int(offset_samples)
        self.n_buoys = len(self.paths)
        self.n_windows = max(0, n_samples) // self.samples_per_window

    def read_window(self, w: int, out: np.ndarray) -> bool:
        if w >= self.n_windows:
            return False
        nb = BYTES_PER_SAMPLE * self.samples_per_window
        a = self._offset + w * nb
        _copy_rows((out[b], m[a:a + nb]) for b, m in enumerate(self._maps))
        return True

    def close(self):
        self._maps = []


class Cu8PipeSource:
    """One readable binary stream per buoy (file objects opened 'rb', sockets wrapped with makefile,
    `subprocess.Popen(['rtl_sdr', ..., '-'], stdout=PIPE).stdout`).  A window is complete when every
    stream has delivered 2N bytes; the stream ends at the first short read."""

    def __init__(self, streams: Sequence, samples_per_window: int):
        self.streams = list(streams)
        self.n_buoys = len(self.streams)
        self.samples_per_window = int(samples_per_window)
        self.n_windows = None            # unknown

    def read_window(self, w: int, out: np.ndarray) -> bool:
        nb = BYTES_PER_SAMPLE * self.samples_per_window
        for b, s in enumerate(self.streams):
            view = memoryview(out[b]).cast("B")
            got = 0
            while got < nb:
                k = s.readinto(view[got:nb])
                if not k:
                    return False
                got += k
        return True

    def close(self):
        pass


class StreamingCorrelator:
    """Ring-buffered front end of `Correlator` for sources that deliver one window at a time."""

    def __init__(self, correlator, depth: int = 3):
        if depth < 2:
            raise ValueError("depth must be at least 2 (one window in flight, one being filled)")
        self.cor = correlator
        self.depth = int(depth)
        B, N = correlator.n_buoys, correlator.n_samples
        self.device = correlator.device
        self._host = torch.empty((self.depth, B, BYTES_PER_SAMPLE * N), dtype=torch.uint8).pin_memory()
        self._dev = torch.empty((B, self.depth, BYTES_PER_SAMPLE * N), dtype=torch.uint8, device=self.device)
        self._copy_stream = torch.cuda.Stream(device=self.device)
        # at most depth-1 windows are pending, so depth+1 result slots are never overwritten before they are read
        self._rec_host = torch.empty((self.depth + 1, correlator.n_pairs, 4), dtype=torch.int32).pin_memory()
        self._en_host = torch.empty((self.depth + 1, B), dtype=torch.int64).pin_memory()
        self.windows_done = 0
        self.h2d_bytes = 0
        self._dirty = False             # True while a run() is in progress or was abandoned with windows in flight

    def run(self, source, max_lag: Optional[int] = None, max_windows: Optional[int] = None) -> Iterator[np.ndarray]:
        cor = self.cor
        if source.n_buoys != cor.n_buoys or source.samples_per_window != cor.n_samples:
            raise ValueError("source delivers %d buoys x %d samples, the correlator expects %d x %d"
                             % (source.n_buoys, source.samples_per_window, cor.n_buoys, cor.n_samples))
        if max_lag != cor.plan.max_lag:
            cor.plan.set_max_lag(max_lag)
        D = self.depth
        host_np = self._host.numpy()
        slot_free = [None] * D          # event: the kernels that read device slot s have finished
        copied = [None] * D             # event: the H2D copy into device slot s has finished
        host_free = [None] * D          # event: the H2D copy out of host slot s has finished
        pending: List = []              # (result slot, done_event) in window order
        with torch.cuda.device(self.device):
            compute = torch.cuda.current_stream()
            # A consumer that stopped iterating an earlier run() early left windows in flight: their kernels and
            # result copies may still be using the device / pinned slots this run is about to overwrite.
            self._copy_stream.wait_stream(compute)
            if self._dirty:
                compute.synchronize()
                self._copy_stream.synchronize()
            self._dirty = True
            w = 0
            while max_windows is None or w < max_windows:
                s = w % D
                view = source.pinned_window(w) if hasattr(source, "pinned_window") else None
                if view is None:
                    if host_free[s] is not None:
                        host_free[s].synchronize()             # pinned slot may still be feeding the DMA engine
                    if not source.read_window(w, host_np[s]):
                        break
                    view = self._host[s]
                with torch.cuda.stream(self._copy_stream):
                    if slot_free[s] is not None:
                        self._copy_stream.wait_event(slot_free[s])
                    for b in range(cor.n_buoys):               # contiguous rows: plain async memcpys
                        self._dev[b, s].copy_(view[b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                    copied[s] = ev
                    host_free[s] = ev
                self.h2d_bytes += view.numel()
                compute.wait_event(copied[s])
                rec, en = cor.run_device(self._dev, [s])
                # results go to page-locked slots right behind this window's kernels, so handing a window back
                # never waits for the kernels of the windows queued after it
                r = w % (D + 1)
                self._rec_host[r].copy_(rec[0], non_blocking=True)
                self._en_host[r].copy_(en[0], non_blocking=True)
                done = torch.cuda.Event()
                done.record(compute)
                slot_free[s] = done
                pending.append((r, done))
                w += 1
                # hand back every window whose kernels have already finished, keeping at most depth-1 in flight
                while pending and (len(pending) >= D - 1 or pending[0][1].query()):
                    yield self._finish(pending.pop(0))
            while pending:
                yield self._finish(pending.pop(0))
            self._dirty = False                                # ran to completion: nothing left in flight

    def _finish(self, item) -> np.ndarray:
        r, done = item
        done.synchronize()
        out = self.cor._finish(self._rec_host[r].numpy()[None].copy(), self._en_host[r].numpy()[None].copy())[0]
        self.windows_done += 1
        return out
