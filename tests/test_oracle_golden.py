"""Pin the CPU oracle to golden vectors produced by the reference's own functions
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np

import oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_unpack_bit_exact(golden_dir):
    g = _load(golden_dir, "unpack.npz")
    x = oracle.unpack_cu8(g["raw"])
    assert x.dtype == np.complex64
    assert np.array_equal(x.view(np.uint32), g["x_file"].view(np.uint32))    # load_iq_data
    assert np.array_equal(x.view(np.uint32), g["x_pipe"].view(np.uint32))    # read_iq_samples
    # known answers: 0 -> -127.5, 255 -> +127.5, 127 -> -0.5, 128 -> +0.5
    assert x[0] == np.complex64(-127.5 + 127.5j) and x[1] == np.complex64(-0.5 + 0.5j)


def test_analyze_spectrum_matches_reference(golden_dir):
    g = _load(golden_dir, "analyze_spectrum.npz")
    x = oracle.unpack_cu8(g["iq"])
    freqs, p_db, peak_freqs = oracle.analyze_spectrum(x, int(g["sample_rate"]), float(g["center_mhz"]))
    assert p_db.dtype == g["p_db"].dtype == np.float32
    assert np.array_equal(freqs, g["freqs"])
    assert np.array_equal(p_db, g["p_db"])
    assert np.array_equal(peak_freqs, g["peak_freqs"])
    # the injected tones are among the detected peaks
    shifted_bins = (g["tone_bins"] + len(x) // 2) % len(x)
    assert set(freqs[shifted_bins]).issubset(set(peak_freqs))
    st = oracle.signal_stats(x)
    assert st["power_db"] == g["power_db"] and st["peak_amplitude"] == g["peak_amplitude"]
    assert st["rms_amplitude"] == g["rms_amplitude"] and st["num_samples"] == int(g["num_samples"])


def test_buoy_detection_matches_reference(golden_dir):
    g = _load(golden_dir, "buoy_detect.npz")
    with open(os.path.join(golden_dir, "buoy_detect.json")) as f:
        want = json.load(f)
    fs, fc_mhz = int(g["sample_rate"]), float(g["center_mhz"])
    fc_hz = int(fc_mhz * 1e6)                                   # buoy_node.py:365
    x = oracle.unpack_cu8(g["iq"])
    p = oracle.spectrum_db(oracle.forward_fft(x))
    peaks = oracle.detect_peaks_fixed(p, height=-70, distance=10)
    got = oracle.score_peaks_buoy(p, peaks, oracle.freq_axis_hz(len(x), fs, fc_hz), fc_hz)
    assert len(got) == len(want) > 100
    for a, b in zip(got, want):
        assert a["frequency_mhz"] == b["frequency_mhz"]
        assert a["signal_strength_dbm"] == b["signal_strength_dbm"]
        assert a["confidence"] == b["confidence"]
        assert a["signal_type"] == b["signal_type"]


def test_stream_detection_matches_reference(golden_dir):
    g = _load(golden_dir, "stream_detect.npz")
    with open(os.path.join(golden_dir, "stream_detect.json")) as f:
        want = json.load(f)
    fs, fc = int(g["sample_rate"]), float(g["center_hz"])
    x = oracle.unpack_cu8(g["iq"])
    p = oracle.spectrum_db(oracle.forward_fft(x))
    peaks = oracle.detect_peaks_fixed(p)
    got = oracle.score_peaks_stream(p, peaks, oracle.freq_axis_hz(len(x), fs, fc), fs)
    assert len(got) == len(want) > 100
    for a, b in zip(got, want):
        for key in ("frequency_mhz", "signal_strength_dbm", "bandwidth_hz", "confidence", "signal_type"):
            assert a[key] == b[key], key


def test_tdoa_seam_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "tdoa.json")) as f:
        g = json.load(f)
    acc = {b[0]: b[4] for b in g["buoys"]}
    dets = [(d[0], d[1], d[4], d[7]) for d in g["detections"][:4]]
    got = oracle.tdoa_measurements(dets, acc)
    assert len(got) == len(g["measurements"])
    for a, b in zip(got, g["measurements"]):
        assert list(a) == b
    # the example in tdoa_processor.py:476-490: 150 000 ns -> 44 968.87 m
    assert abs(150000 / 1e9 * oracle.SPEED_OF_LIGHT - 44968.8687) < 1e-3


def test_lag_sign_convention():
    """If buoy j hears the waveform d samples later the peak lag is +d (tdoa_processor.py:51)."""
    from radio_mapper_b200 import synth
    iq, d, _ = synth.delayed_buoys(7, 3, 4096, delays=[0, 25, -40], snr_db=20)
    r = oracle.xcorr_pairs_peak(iq)
    assert list(r["lag"]) == [25, -40, -65]
    r2 = oracle.xcorr_pairs_peak(iq, max_lag=100)
    assert list(r2["lag"]) == [25, -40, -65]
    assert np.allclose(r2["frac"], r["frac"]) and np.array_equal(r2["peak"], r["peak"])
    assert oracle.lag_to_tdoa_ns(25, 0.0, 2048000) == 12207


def test_welch_definition():
    """welch_psd equals the textbook average of Hann-windowed periodograms."""
    from radio_mapper_b200 import synth
    u, bins = synth.welch_stream(11, 4, 1024, 2_400_000, n_tones=3)
    x = oracle.unpack_cu8(u)
    f, pxx = oracle.welch_psd(x, 2_400_000, nperseg=1024)
    w = np.hanning(1025)[:1024].astype(np.float64)              # periodic hann == scipy 'hann' sym=False
    seg = x.reshape(4, 1024).astype(np.complex128) * w
    ref = (np.abs(np.fft.fft(seg, axis=1)) ** 2).mean(axis=0) / (2_400_000 * (w ** 2).sum())
    assert np.allclose(pxx, ref, rtol=2e-5)
    assert set(bins).issubset(set(np.argsort(pxx)[-3 * 3:]))
