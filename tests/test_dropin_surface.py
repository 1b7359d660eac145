"""CPU: the drop-in modules keep the reference's signatures (SURVEY §8b "preserve verbatim").

tests/golden/signatures.json was recorded by tests/golden/make_golden.py from the imported reference modules
(inspect.signature / dataclasses.fields); here the same introspection runs on the root-level shims."""
import dataclasses
import importlib
import inspect
import json
import os

import pytest


@pytest.fixture(scope="module")
def surface(golden_dir):
    with open(os.path.join(golden_dir, "signatures.json")) as f:
        return json.load(f)


def _params(obj):
    return [[q.name, q.kind.name, None if q.default is inspect.Parameter.empty else repr(q.default)]
            for q in inspect.signature(obj).parameters.values()]


def _resolve(module, dotted):
    obj = module
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


@pytest.mark.parametrize("modname", ["buoy_node", "iq_stream_client", "tdoa_processor", "signal_analyzer"])
def test_signatures_match_reference(surface, modname):
    mod = importlib.import_module(modname)              # the root-level shim, as a user of the reference imports it
    assert mod.__name__ == modname
    for name, want in surface[modname].items():
        obj = _resolve(mod, name)
        if isinstance(want, list):                       # dataclass: field names, order and defaults
            got = [[f.name, None if f.default is dataclasses.MISSING else repr(f.default)] for f in dataclasses.fields(obj)]
            assert got == want, "%s.%s fields" % (modname, name)
            continue
        got = _params(obj)
        ref = want["params"]
        # the reference's parameters come first, same names / kinds / defaults; the drop-in may only ADD optional ones
        assert got[:len(ref)] == ref, "%s.%s: %s vs reference %s" % (modname, name, got, want["text"])
        for extra in got[len(ref):]:
            assert extra[2] is not None or extra[1].startswith("VAR_"), "%s.%s adds a required parameter %s" % (modname, name, extra)


def test_buoy_detector_error_paths_follow_reference(monkeypatch):
    """rtl_sdr missing / failing / timing out -> the reference's fallback generator (buoy_node.py:461-468);
    a short read -> [] (:388-390).  None of these touch the GPU."""
    import subprocess
    import buoy_node as bn
    gps = bn.GPSTimeSource()
    gps.lat, gps.lng = 35.4676, -97.5164

    def raising(exc):
        def capture(fc, fs, n):
            raise exc
        return capture

    monkeypatch.setattr(bn.SignalDetector, "_fallback_signal_detection", lambda self, f: ["fallback", f])
    for exc in (FileNotFoundError("rtl_sdr"), subprocess.TimeoutExpired("rtl_sdr", 5), bn.CaptureError("exit 1"), OSError("usb")):
        det = bn.SignalDetector("BUOY_X", gps, capture=raising(exc))
        assert det._detect_real_signals(121.5) == ["fallback", 121.5]
    det = bn.SignalDetector("BUOY_X", gps, capture=lambda fc, fs, n: b"\x80" * 100)
    assert det._detect_real_signals(121.5) == []
    # default capture = the rtl_sdr subprocess, built like the reference (:368-380); the binary is absent here
    det = bn.SignalDetector("BUOY_X", gps)
    assert det._capture is bn.rtl_sdr_capture
    assert det._detect_real_signals(243.0) == ["fallback", 243.0]


def test_fallback_generator_matches_reference_distribution():
    import random
    import buoy_node as bn
    gps = bn.GPSTimeSource()
    det = bn.SignalDetector("BUOY_X", gps)
    random.seed(5)
    out = [d for _ in range(400) for d in det._fallback_signal_detection(121.5)]
    assert 60 <= len(out) <= 140                                   # p = 0.25
    assert all(d.signal_type == "emergency" and -85 <= d.signal_strength_dbm <= -65 and 0.3 <= d.confidence <= 0.95
               and d.frequency_mhz == 121.5 for d in out)
    out = [d for _ in range(200) for d in det._fallback_signal_detection(150.0)]
    assert all(d.signal_type == "unknown" and -80 <= d.signal_strength_dbm <= -50 for d in out)


def test_stream_capture_not_running_returns_none():
    import iq_stream_client as sc
    cap = sc.RealTimeSDRCapture()
    assert cap.read_iq_samples() is None and cap.read_iq_samples(16) is None
    assert cap.sample_rate == 2048000 and cap.center_freq_hz == 100000000 and cap.fft_size == 1024

    class _Proc:
        def __init__(self, data):
            import io
            self.stdout = io.BytesIO(data)
    cap.running, cap.capture_process = True, _Proc(b"\x00" * 10)
    assert cap.read_iq_samples(8) is None                          # short read (:143-145)
    det = sc.SignalDetector("NODE")
    assert (det.node_id, det.sample_rate, det.detection_threshold, det.lat, det.lng) == ("NODE", 2048000, -70, 35.4676, -97.5164)
    assert det.signal_history == [] and det.max_history_size == 1000
    assert det._classify_signal(121.5e6) == "emergency" and det._classify_signal(100e6) == "fm_radio"
