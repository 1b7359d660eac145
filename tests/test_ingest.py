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


@pytest.mark.gpu
def test_cu8_window_frames_reach_the_correlator():
    """binary cu8_window frames from the buoys (any interleaving) -> WindowAssembler -> GPU correlate.
    Measurements equal correlate_iq on the same bytes; capture-start offsets enter time_difference_ns exactly."""
    import random
    from radio_mapper_b200 import wire
    from radio_mapper_b200.tdoa_processor import TDOAProcessor
    n, fs, B, W = 1 << 15, 2048000, 4, 3
    ids = ["B%d" % b for b in range(B)]
    blocks = [synth.delayed_buoys(40 + w, B, n)[0] for w in range(W)]          # uint8[B, 2N] per window
    base = 1_735_689_600_000_000_000
    start = {(b, w): base + w * int(n / fs * 1e9) + (137 * b if w == 1 else 0) for b in range(B) for w in range(W)}
    frames = [wire.pack_cu8_window(ids[b], w, start[(b, w)], fs, 121.5e6, blocks[w][b]) for w in range(W) for b in range(B)]
    random.Random(3).shuffle(frames)
    proc = TDOAProcessor()
    got = dict(proc.correlate_window_frames(frames, ids, n, depth=4))
    assert sorted(got) == [0, 1, 2]
    for w in range(W):
        direct = proc.correlate_iq(blocks[w][:, None, :], ids, fs, 121.5)
        assert len(got[w]) == len(direct) == 6
        for g, d in zip(got[w], direct):
            i, j = ids.index(g.buoy1_id), ids.index(g.buoy2_id)
            assert (g.buoy1_id, g.buoy2_id, g.frequency_mhz) == (d.buoy1_id, d.buoy2_id, 121.5)
            assert g.time_difference_ns == d.time_difference_ns + (start[(j, w)] - start[(i, w)])
            assert g.confidence == d.confidence
