#!/usr/bin/env python3
"""Generate golden vectors by RUNNING THE REFERENCE's own functions.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It imports the unmodified reference modules (stubbing only the absent third-party
imports `matplotlib`, `websocket`, `requests` and the `rtl_sdr` subprocess / pipe, which
are I/O and not arithmetic), feeds them seeded synthetic cu8 bytes from
`radio_mapper_b200.synth`, and stores inputs + outputs under tests/golden/*.npz / *.json.
`tests/test_oracle_golden.py` pins the oracle to these files; the GPU parity tests then
compare the CUDA path with the oracle AND with these files.
"""
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RADIO_MAPPER_REF", "/root/reference")
sys.path.insert(0, ROOT)

from radio_mapper_b200 import synth  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _stub("websocket", WebSocketApp=object)
    if "requests" not in sys.modules:
        try:
            import requests  # noqa: F401
        except Exception:
            _stub("requests")
    sys.path.insert(0, REF)
    import signal_analyzer as ref_sa
    import tdoa_processor as ref_tdoa
    import buoy_node as ref_buoy
    import iq_stream_client as ref_stream
    return ref_sa, ref_tdoa, ref_buoy, ref_stream


class _FakeProc:
    """Stands in for the rtl_sdr subprocess (buoy_node.py:379-381, iq_stream_client.py:110)."""

    def __init__(self, data):
        self._data = data
        self.stdout = io.BytesIO(data)
        self.returncode = 0

    def communicate(self, timeout=None):
        return self._data, b""


def main():
    ref_sa, ref_tdoa, ref_buoy, ref_stream = import_reference()
    import contextlib
    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---- a1: unpack through load_iq_data (file) and read_iq_samples (pipe) -------------
    rng = np.random.default_rng(101)
    raw = rng.integers(0, 256, size=2 * 4096, dtype=np.uint8)
    raw[:8] = [0, 255, 127, 128, 1, 254, 255, 0]
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(raw.tobytes())
        path = f.name
    with quiet:
        x_file, fs = ref_sa.load_iq_data(path)
    os.unlink(path)
    cap = ref_stream.RealTimeSDRCapture()
    cap.running = True
    cap.capture_process = _FakeProc(raw.tobytes())
    x_pipe = cap.read_iq_samples(4096)
    assert x_file.dtype == np.complex64 and x_pipe.dtype == np.complex64
    np.savez_compressed(os.path.join(HERE, "unpack.npz"), raw=raw, x_file=x_file, x_pipe=x_pipe,
                        sample_rate=fs)

    # ---- a2,a3,a4(mean+10),a6,a7: signal_analyzer on an 8192-sample tone block ----------
    iq, bins = synth.tones_block(202, 8192)
    x = np.asarray(ref_stream_unpack(ref_stream, iq))
    with quiet:
        freqs, p_db, peak_freqs = ref_sa.analyze_spectrum(x, 2048000, 121.5)
        stats = ref_sa.calculate_signal_stats(x)
    np.savez_compressed(os.path.join(HERE, "analyze_spectrum.npz"), iq=iq, tone_bins=bins,
                        freqs=freqs, p_db=p_db, peak_freqs=peak_freqs,
                        power_db=stats["power_db"], peak_amplitude=stats["peak_amplitude"],
                        rms_amplitude=stats["rms_amplitude"], num_samples=stats["num_samples"],
                        sample_rate=2048000, center_mhz=121.5)

    # ---- a2..a5 buoy_node._detect_real_signals on 32768 samples (rtl_sdr mocked) -------
    iq_b, bins_b = synth.tones_block(303, 32768)
    gps = ref_buoy.GPSTimeSource.__new__(ref_buoy.GPSTimeSource)
    gps.gps_locked = True
    gps.lat, gps.lng = 35.4676, -97.5164
    gps.get_precise_timestamp = lambda: ("2025-01-01T00:00:00+00:00", 1735689600000000000)
    det = ref_buoy.SignalDetector("BUOY_T", gps)
    orig_popen = ref_buoy.subprocess.Popen
    ref_buoy.subprocess.Popen = lambda *a, **k: _FakeProc(iq_b.tobytes())
    try:
        dets = det._detect_real_signals(121.5)
    finally:
        ref_buoy.subprocess.Popen = orig_popen
    buoy_json = [dict(frequency_mhz=d.frequency_mhz, signal_strength_dbm=float(d.signal_strength_dbm),
                      confidence=float(d.confidence), signal_type=d.signal_type) for d in dets]
    np.savez_compressed(os.path.join(HERE, "buoy_detect.npz"), iq=iq_b, tone_bins=bins_b,
                        center_mhz=121.5, sample_rate=2048000)
    with open(os.path.join(HERE, "buoy_detect.json"), "w") as f:
        json.dump(buoy_json, f, indent=1)

    # ---- a2..a5,a8 iq_stream_client.SignalDetector.detect_signals on 8192 samples ------
    iq_s, bins_s = synth.tones_block(404, 8192)
    xs = ref_stream_unpack(ref_stream, iq_s)
    sdet = ref_stream.SignalDetector("NODE_T")
    sdets = sdet.detect_signals(xs, 100e6)
    stream_json = [dict(frequency_mhz=float(d.frequency_mhz), signal_strength_dbm=float(d.signal_strength_dbm),
                        bandwidth_hz=float(d.bandwidth_hz), confidence=float(d.confidence),
                        signal_type=d.signal_type) for d in sdets]
    np.savez_compressed(os.path.join(HERE, "stream_detect.npz"), iq=iq_s, tone_bins=bins_s,
                        center_hz=100e6, sample_rate=2048000)
    with open(os.path.join(HERE, "stream_detect.json"), "w") as f:
        json.dump(stream_json, f, indent=1)

    # ---- a9 + geodesy + grouping: tdoa_processor -------------------------------------
    T = ref_tdoa
    buoys = [T.BuoyPosition("BUOY_A", 35.4676, -97.5164, 0.0, 50000),
             T.BuoyPosition("BUOY_B", 35.5276, -97.5164, 10.0, 75000),
             T.BuoyPosition("BUOY_C", 35.4676, -97.4464, 0.0, 60000),
             T.BuoyPosition("BUOY_D", 35.4076, -97.5864, 5.0, 100000)]
    base = 1735689600000000000
    dets = [T.SignalDetection("BUOY_A", 121.5, -55.0, "2025-01-01T00:00:00Z", base, 35.4676, -97.5164, 0.9, "emergency"),
            T.SignalDetection("BUOY_B", 121.5, -60.0, "2025-01-01T00:00:00Z", base + 15000, 35.5276, -97.5164, 0.85, "emergency"),
            T.SignalDetection("BUOY_C", 121.505, -58.0, "2025-01-01T00:00:00Z", base - 7300, 35.4676, -97.4464, 0.88, "aviation"),
            T.SignalDetection("BUOY_D", 121.52, -61.0, "2025-01-01T00:00:00Z", base + 22100, 35.4076, -97.5864, 0.5, "emergency"),
            T.SignalDetection("BUOY_B", 243.0, -50.0, "2025-01-01T00:00:00Z", base + 40, 35.5276, -97.5164, 0.7, "emergency"),
            T.SignalDetection("BUOY_A", 121.5, -70.0, "2025-01-01T00:00:00Z", base - 11_000_000_000, 35.4676, -97.5164, 0.4, "emergency")]
    proc = T.TDoAProcessor()
    for b in buoys:
        proc.register_buoy(b)
    meas = proc.tdoa_calculator.calculate_tdoa_measurements(dets[:4], proc.buoy_positions)
    groups = proc._group_by_frequency(dets)
    filt = proc._filter_by_time_window([dets[0], dets[1], dets[5]])
    G = T.GeodeticCalculator
    geo = dict(
        xyz=[list(G.lat_lng_to_xyz(b.lat, b.lng, b.altitude)) for b in buoys],
        back=[list(G.xyz_to_lat_lng(*G.lat_lng_to_xyz(b.lat, b.lng, b.altitude))) for b in buoys],
        d3=[G.distance_3d(buoys[0].lat, buoys[0].lng, buoys[0].altitude, b.lat, b.lng, b.altitude) for b in buoys],
        bearing=[list(G.bearing_distance(buoys[0].lat, buoys[0].lng, b.lat, b.lng)) for b in buoys[1:]],
    )
    # a solvable multilateration: exact dt from a transmitter inside the network
    tx = (35.47, -97.50, 0.0)
    sdets = []
    for b in buoys:
        dist = G.distance_3d(tx[0], tx[1], tx[2], b.lat, b.lng, b.altitude)
        sdets.append(T.SignalDetection(b.buoy_id, 121.5, -55.0, "2025-01-01T00:00:00Z",
                                       base + int(dist / T.TDoACalculator.SPEED_OF_LIGHT * 1e9),
                                       b.lat, b.lng, 0.9, "emergency"))
    smeas = proc.tdoa_calculator.calculate_tdoa_measurements(sdets, proc.buoy_positions)
    res = proc.hyperbolic_positioner.triangulate_position(smeas, proc.buoy_positions)
    tdoa_json = dict(
        buoys=[[b.buoy_id, b.lat, b.lng, b.altitude, b.timing_accuracy_ns] for b in buoys],
        detections=[[d.buoy_id, d.frequency_mhz, d.signal_strength_dbm, d.timestamp_utc, d.gps_timestamp_ns,
                     d.lat, d.lng, d.confidence, d.signal_type] for d in dets],
        measurements=[[m.buoy1_id, m.buoy2_id, m.time_difference_ns, m.distance_difference_m, m.confidence,
                       m.frequency_mhz] for m in meas],
        groups={str(k): [dets.index(d) for d in v] for k, v in groups.items()},
        time_filtered=[[dets[0], dets[1], dets[5]].index(d) for d in filt],
        geodesy=geo,
        status=proc.get_buoy_network_status(),
        solve=dict(tx=list(tx),
                   detections=[[d.buoy_id, d.gps_timestamp_ns] for d in sdets],
                   measurements=[[m.buoy1_id, m.buoy2_id, m.time_difference_ns, m.distance_difference_m,
                                  m.confidence] for m in smeas],
                   result=None if res is None else dict(lat=res.estimated_lat, lng=res.estimated_lng,
                                                        alt=res.estimated_altitude, accuracy=res.accuracy_meters,
                                                        confidence=res.confidence, method=res.method,
                                                        contributing=sorted(res.contributing_buoys))),
    )
    with open(os.path.join(HERE, "tdoa.json"), "w") as f:
        json.dump(tdoa_json, f, indent=1)

    # ---- the reference's own worked example (tdoa_processor.py:472-490): three buoys, dt = 150 000 / 300 000 ns ----
    ex_buoys = [T.BuoyPosition("BUOY_ALPHA", 51.505, -0.09, 0.0, 50000),
                T.BuoyPosition("BUOY_BETA", 51.51, -0.1, 0.0, 75000),
                T.BuoyPosition("BUOY_GAMMA", 51.5, -0.12, 0.0, 60000)]
    ex_proc = T.TDoAProcessor()
    for b in ex_buoys:
        ex_proc.register_buoy(b)
    t0 = 1737217800000000000
    ex_dets = [T.SignalDetection("BUOY_ALPHA", 121.5, -55, "2025-01-18T16:30:00Z", t0, 51.505, -0.09, 0.9, "emergency"),
               T.SignalDetection("BUOY_BETA", 121.5, -60, "2025-01-18T16:30:00Z", t0 + 150000, 51.51, -0.1, 0.85, "emergency"),
               T.SignalDetection("BUOY_GAMMA", 121.5, -58, "2025-01-18T16:30:00Z", t0 + 300000, 51.5, -0.12, 0.88, "emergency")]
    ex_meas = ex_proc.tdoa_calculator.calculate_tdoa_measurements(ex_dets, ex_proc.buoy_positions)
    with quiet:
        ex_res = ex_proc.process_signal_detections(ex_dets)
    example = dict(
        buoys=[[b.buoy_id, b.lat, b.lng, b.altitude, b.timing_accuracy_ns] for b in ex_buoys],
        base_time_ns=t0, sample_rate=2048000, frequency_mhz=121.5,
        measurements=[[m.buoy1_id, m.buoy2_id, m.time_difference_ns, m.distance_difference_m, m.confidence, m.frequency_mhz]
                      for m in ex_meas],
        results=[dict(lat=r.estimated_lat, lng=r.estimated_lng, accuracy=r.accuracy_meters, confidence=r.confidence,
                      signal_type=r.signal_type, contributing=sorted(r.contributing_buoys)) for r in ex_res])
    with open(os.path.join(HERE, "example_main.json"), "w") as f:
        json.dump(example, f, indent=1)

    # ---- (b) the drop-in surface: signatures and dataclass fields of the reference, as strings ----
    import dataclasses
    import inspect

    def sig(obj):
        sg = inspect.signature(obj)
        return {"text": str(sg),
                "params": [[q.name, q.kind.name, None if q.default is inspect.Parameter.empty else repr(q.default)]
                           for q in sg.parameters.values()]}

    def fields(cls):
        return [[f.name, None if f.default is dataclasses.MISSING else repr(f.default)] for f in dataclasses.fields(cls)]

    surface = {
        "buoy_node": {
            "SignalDetector.__init__": sig(ref_buoy.SignalDetector.__init__),
            "SignalDetector._detect_real_signals": sig(ref_buoy.SignalDetector._detect_real_signals),
            "SignalDetector._classify_signal": sig(ref_buoy.SignalDetector._classify_signal),
            "SignalDetector._fallback_signal_detection": sig(ref_buoy.SignalDetector._fallback_signal_detection),
            "GPSTimeSource.__init__": sig(ref_buoy.GPSTimeSource.__init__),
            "GPSTimeSource.get_precise_timestamp": sig(ref_buoy.GPSTimeSource.get_precise_timestamp),
            "GPSTimeSource.get_position": sig(ref_buoy.GPSTimeSource.get_position),
            "SignalDetection": fields(ref_buoy.SignalDetection),
        },
        "iq_stream_client": {
            "RealTimeSDRCapture.__init__": sig(ref_stream.RealTimeSDRCapture.__init__),
            "RealTimeSDRCapture.read_iq_samples": sig(ref_stream.RealTimeSDRCapture.read_iq_samples),
            "RealTimeSDRCapture.start_capture": sig(ref_stream.RealTimeSDRCapture.start_capture),
            "RealTimeSDRCapture.stop_capture": sig(ref_stream.RealTimeSDRCapture.stop_capture),
            "SignalDetector.__init__": sig(ref_stream.SignalDetector.__init__),
            "SignalDetector.detect_signals": sig(ref_stream.SignalDetector.detect_signals),
            "SignalDetector._estimate_bandwidth": sig(ref_stream.SignalDetector._estimate_bandwidth),
            "SignalDetector._classify_signal": sig(ref_stream.SignalDetector._classify_signal),
            "SignalDetector._extract_signal_samples": sig(ref_stream.SignalDetector._extract_signal_samples),
            "SignalDetection": fields(ref_stream.SignalDetection),
        },
        "tdoa_processor": {
            "TDoAProcessor.__init__": sig(T.TDoAProcessor.__init__),
            "TDoAProcessor.register_buoy": sig(T.TDoAProcessor.register_buoy),
            "TDoAProcessor.process_signal_detections": sig(T.TDoAProcessor.process_signal_detections),
            "TDoAProcessor._group_by_frequency": sig(T.TDoAProcessor._group_by_frequency),
            "TDoAProcessor._filter_by_time_window": sig(T.TDoAProcessor._filter_by_time_window),
            "TDoAProcessor.get_buoy_network_status": sig(T.TDoAProcessor.get_buoy_network_status),
            "TDoACalculator.calculate_tdoa_measurements": sig(T.TDoACalculator.calculate_tdoa_measurements),
            "HyperbolicPositioning.triangulate_position": sig(T.HyperbolicPositioning.triangulate_position),
            "GeodeticCalculator.lat_lng_to_xyz": sig(T.GeodeticCalculator.lat_lng_to_xyz),
            "GeodeticCalculator.xyz_to_lat_lng": sig(T.GeodeticCalculator.xyz_to_lat_lng),
            "GeodeticCalculator.distance_3d": sig(T.GeodeticCalculator.distance_3d),
            "GeodeticCalculator.bearing_distance": sig(T.GeodeticCalculator.bearing_distance),
            "BuoyPosition": fields(T.BuoyPosition), "SignalDetection": fields(T.SignalDetection),
            "TDoAMeasurement": fields(T.TDoAMeasurement), "TriangulationResult": fields(T.TriangulationResult),
        },
        "signal_analyzer": {name: sig(getattr(ref_sa, name)) for name in
                            ("load_iq_data", "analyze_spectrum", "calculate_signal_stats", "plot_spectrum", "analyze_iq_file")},
    }
    with open(os.path.join(HERE, "signatures.json"), "w") as f:
        json.dump(surface, f, indent=1)
    print("golden vectors written to", HERE)
    print("  buoy detections:", len(buoy_json), " stream detections:", len(stream_json),
          " tdoa measurements:", len(meas), " solve:", tdoa_json["solve"]["result"])


def ref_stream_unpack(ref_stream, iq_u8):
    cap = ref_stream.RealTimeSDRCapture()
    cap.running = True
    cap.capture_process = _FakeProc(np.asarray(iq_u8).tobytes())
    return cap.read_iq_samples(len(iq_u8) // 2)


if __name__ == "__main__":
    main()
