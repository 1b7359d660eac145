"""CPU: the binary cu8_window frame and the window assembler, plus the start-time seam."""
import json

import numpy as np
import pytest

from radio_mapper_b200 import wire
from radio_mapper_b200 import tdoa_processor as T


def _frame(buoy, w, n=64, seed=0, gps=1_735_689_600_000_000_000):
    rng = np.random.default_rng(seed)
    iq = rng.integers(0, 256, size=2 * n, dtype=np.uint8)
    return wire.pack_cu8_window(buoy, w, gps, 2.048e6, 121.5e6, iq), iq


def test_roundtrip_and_layout():
    frame, iq = _frame("BUOY_Ä", 7, n=100, seed=3)
    assert frame[:4] == b"RMXW" and len(frame) == 56 + len("BUOY_Ä".encode()) + 200
    msg = wire.unpack_cu8_window(frame)
    assert (msg.buoy_id, msg.window_index, msg.gps_timestamp_ns, msg.sample_rate_hz, msg.center_freq_hz, msg.n_samples) == \
           ("BUOY_Ä", 7, 1_735_689_600_000_000_000, 2.048e6, 121.5e6, 100)
    assert np.array_equal(msg.iq_u8, iq) and msg.iq_u8.dtype == np.uint8
    empty = wire.unpack_cu8_window(wire.pack_cu8_window("B", 0, 0, 1.0, 0.0, np.empty(0, np.uint8)))
    assert empty.n_samples == 0                                   # an empty window is a valid frame
    kind, obj = wire.dispatch(frame)
    assert kind == "cu8_window" and obj.window_index == 7
    kind, obj = wire.dispatch(json.dumps({"type": "signal_detection", "data": {"buoy_id": "B"}}))
    assert kind == "signal_detection" and obj["data"]["buoy_id"] == "B"   # text frames stay the reference's JSON


def test_malformed_frames_are_rejected():
    frame, _ = _frame("B0", 1)
    with pytest.raises(wire.WireError):
        wire.unpack_cu8_window(frame[:20])
    with pytest.raises(wire.WireError):
        wire.unpack_cu8_window(b"XXXX" + frame[4:])
    with pytest.raises(wire.WireError):
        wire.unpack_cu8_window(frame[:-1])                        # truncated payload
    bad = bytearray(frame)
    bad[-1] ^= 0xFF
    with pytest.raises(wire.WireError):
        wire.unpack_cu8_window(bytes(bad))                        # checksum
    assert wire.unpack_cu8_window(bytes(bad), verify=False).n_samples == 64
    with pytest.raises(wire.WireError):
        wire.pack_cu8_window("B0", 0, 0, 1.0, 1.0, np.zeros(3, np.uint8))   # odd byte count


def test_assembler_releases_complete_windows_in_any_interleaving():
    ids = ["A", "B", "C"]
    asm = wire.WindowAssembler(ids, 64, depth=3, pinned=False)
    frames = {(b, w): _frame(b, w, seed=10 * w + k, gps=1000 * w + k) for k, b in enumerate(ids) for w in range(4)}
    order = [("A", 0), ("B", 0), ("A", 1), ("C", 0), ("C", 1), ("B", 1), ("C", 2), ("B", 2), ("A", 2)]
    done = []
    for key in order:
        w = asm.add(wire.unpack_cu8_window(frames[key][0]))
        if w is not None:
            block, stamps, fs, fc = asm.take(w)
            assert block.shape == (3, 1, 128) and (fs, fc) == (2.048e6, 121.5e6)
            for k, b in enumerate(ids):
                assert np.array_equal(np.asarray(block[k, 0]), frames[(b, w)][1])
            assert stamps == [1000 * w + k for k in range(3)]
            done.append(w)
    assert done == [0, 1, 2] and asm.dropped == 0
    # out-of-order completion: window 3 completes before window... (a fresh assembler, windows 1 then 0)
    asm2 = wire.WindowAssembler(ids, 64, depth=3, pinned=False)
    seq = [("A", 1), ("B", 1), ("A", 0), ("C", 1), ("B", 0), ("C", 0)]
    done2 = []
    for key in seq:
        w = asm2.add(wire.unpack_cu8_window(frames[key][0]))
        if w is not None:
            asm2.take(w)
            done2.append(w)
    assert done2 == [1, 0] and asm2.dropped == 0
    assert asm.add(wire.unpack_cu8_window(frames[("A", 1)][0])) is None and asm.dropped == 1   # late duplicate
    with pytest.raises(KeyError):
        asm.take(3)
    with pytest.raises(wire.WireError):
        asm.add(wire.unpack_cu8_window(_frame("Z", 3)[0]))         # unknown buoy
    with pytest.raises(wire.WireError):
        asm.add(wire.unpack_cu8_window(_frame("A", 3, n=32)[0]))   # wrong window length


def test_assembler_drops_windows_a_slow_buoy_never_completes():
    asm = wire.WindowAssembler(["A", "B"], 16, depth=2, pinned=False)
    for w in range(5):                                            # B never sends: A's old windows are evicted
        assert asm.add(wire.unpack_cu8_window(_frame("A", w, n=16)[0])) is None
    assert asm.dropped == 3 and sorted(asm._have) == [3, 4]
    assert asm.add(wire.unpack_cu8_window(_frame("B", 1, n=16)[0])) is None and asm.dropped == 4   # too old
    assert asm.add(wire.unpack_cu8_window(_frame("B", 4, n=16)[0])) == 4
    out = list(asm.feed([]))
    assert out == []


def test_start_time_offsets_enter_the_time_difference():
    calc = T.TDoACalculator()
    pairs = np.array([[0, 1], [0, 2], [1, 2]], dtype=np.int32)
    lag = np.array([100, -50, 7], dtype=np.int32)
    frac = np.array([0.25, -0.5, 0.0], dtype=np.float32)
    conf = np.ones(3, dtype=np.float32)
    base = calc.measurements_from_lags(["A", "B", "C"], pairs, lag, frac, conf, 2.048e6, 121.5)
    t0 = [1_000_000_000, 1_000_000_300, 999_999_000]
    moved = calc.measurements_from_lags(["A", "B", "C"], pairs, lag, frac, conf, 2.048e6, 121.5, start_ns=t0)
    assert [m.time_difference_ns - b.time_difference_ns for m, b in zip(moved, base)] == [300, -1000, -1300]
    assert all(isinstance(m.time_difference_ns, int) for m in moved)
    assert base[0].time_difference_ns == int(round(100.25 / 2.048e6 * 1e9))
