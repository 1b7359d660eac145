"""Streaming ingest: raw cu8 sources on the CPU, and on the GPU the streamed
records against the batched call on the same bytes."""
import io
import os

import numpy as np
import pytest

from radio_mapper_b200 import synth


def _ingest():
    from radio_mapper_b200 import ingest          # imports torch only
    return ingest


def test_file_source_windows_and_ragged_tail(tmp_path):
    ingest = _ingest()
    rng = np.random.default_rng(5)
    n = 1000
    data = [rng.integers(0, 256, size=2 * (3 * n + extra) + odd, dtype=np.uint8)
            for extra, odd in ((0, 0), (17, 1), (999, 0))]
    paths = []
    for b, d in enumerate(data):
        p = tmp_path / ("iq_capture_100.0MHz_%d.bin" % b)     # the name sdr_capture.py:26 gives its files
        d.tofile(p)
        paths.append(str(p))
    src = ingest.Cu8FileSource(paths, n)
    assert src.n_buoys == 3 and src.n_windows == 3            # the shortest file has exactly 3 windows
    out = np.empty((3, 2 * n), np.uint8)
    for w in range(3):
        assert src.read_window(w, out)
        for b in range(3):
            assert np.array_equal(out[b], data[b][2 * n * w:2 * n * (w + 1)])
    assert not src.read_window(3, out)
    off = ingest.Cu8FileSource(paths, n, offset_samples=500)
    assert off.n_windows == 2
    assert off.read_window(0, out) and np.array_equal(out[1], data[1][1000:1000 + 2 * n])
    with pytest.raises(ValueError):
        ingest.Cu8FileSource([], n)
    empty = tmp_path / "empty.bin"
    empty.write_bytes(b"\x01")
    with pytest.raises(ValueError):
        ingest.Cu8FileSource([str(empty)], n)


def test_pipe_source_short_reads_and_end_of_stream():
    ingest = _ingest()

    class Dribble(io.RawIOBase):
        """delivers at most 333 bytes per readinto, like a pipe would"""
        def __init__(self, payload):
            self.buf, self.pos = payload, 0

        def readable(self):
            return True

        def readinto(self, b):
            k = min(len(b), 333, len(self.buf) - self.pos)
            b[:k] = self.buf[self.pos:self.pos + k]
            self.pos += k
            return k

    rng = np.random.default_rng(6)
    n = 512
    payload = [rng.integers(0, 256, size=2 * n * 2 + 100, dtype=np.uint8).tobytes() for _ in range(2)]
    src = ingest.Cu8PipeSource([Dribble(p) for p in payload], n)
    out = np.empty((2, 2 * n), np.uint8)
    for w in range(2):
        assert src.read_window(w, out)
        for b in range(2):
            assert out[b].tobytes() == payload[b][2 * n * w:2 * n * (w + 1)]
    assert not src.read_window(2, out)                        # 100 trailing bytes: not a full window


def test_array_source_shapes():
    ingest = _ingest()
    a = np.arange(2 * 3 * 2 * 8, dtype=np.uint8).reshape(2, -1)
    src = ingest.ArraySource(a, 8)
    assert src.n_windows == 3
    out = np.empty((2, 16), np.uint8)
    assert src.read_window(2, out) and np.array_equal(out, a[:, 32:48])
    assert not src.read_window(3, out)
    with pytest.raises(TypeError):
        ingest.ArraySource(a.astype(np.int16), 8)


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [2, 3, 5])
def test_streamed_records_equal_batched(tmp_path, depth):
    import torch
    from radio_mapper_b200.tdoa_processor import TDoAProcessor
    from radio_mapper_b200 import ingest
    B, W, n = 4, 7, 1 << 15
    iq, delays = synth.delayed_buoys_torch(31, B, W, n, torch.device("cpu"))
    iq = iq.numpy()                                            # [B, W, 2n]
    paths = []
    for b in range(B):
        p = tmp_path / ("iq_capture_433.9MHz_%d.bin" % b)
        iq[b].reshape(-1).tofile(p)
        paths.append(str(p))
    ids = ["BUOY_%d" % b for b in range(B)]
    proc = TDoAProcessor()
    batched = proc.correlate_iq(torch.from_numpy(iq), ids, 2_048_000, 433.9)
    streamed = []
    for meas in proc.correlate_stream(ingest.Cu8FileSource(paths, n), ids, 2_048_000, 433.9, depth=depth):
        streamed.extend(meas)
    assert len(streamed) == len(batched) == W * 6
    assert streamed == batched                                 # same kernels on the same bytes: identical dataclasses
    pairs = [(i, j) for i in range(B) for j in range(i + 1, B)]
    for w in range(W):
        got = [round(m.time_difference_ns * 2_048_000 / 1e9) for m in streamed[w * 6:(w + 1) * 6]]
        assert got == [int(delays[w, j] - delays[w, i]) for i, j in pairs]
    # a second pass over an in-memory source reuses the ring
    again = [m for meas in proc.correlate_stream(ingest.ArraySource(iq.reshape(B, -1), n), ids, 2_048_000, 433.9,
                                                 depth=depth, max_windows=3) for m in meas]
    assert again == batched[:18]
