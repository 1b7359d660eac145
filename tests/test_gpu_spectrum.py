"""GPU parity of the PSD / detection stages and the drop-in host modules."""
import json
import os

import numpy as np
import pytest
import scipy.signal

import oracle
from radio_mapper_b200 import synth

pytestmark = pytest.mark.gpu


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _db_close(got, ref, atol=1e-3, floor_margin=30.0):
    """dB spectra agree to `atol` for every bin within `floor_margin` dB of the median level.  Both
    sides are float32 FFTs whose absolute error is ~2e-7 of the spectrum's rms, so weaker bins are
    compared in linear amplitude instead (error <= 2e-6 of the rms amplitude)."""
    ok = ref > np.median(ref) - floor_margin
    assert ok.mean() > 0.99
    assert np.max(np.abs(got[ok] - ref[ok])) <= atol, np.max(np.abs(got[ok] - ref[ok]))
    if (~ok).any():
        a, b = 10.0 ** (got[~ok].astype(np.float64) / 20), 10.0 ** (ref[~ok].astype(np.float64) / 20)
        rms = 10.0 ** (np.median(ref) / 20)
        assert np.max(np.abs(a - b)) <= 2e-6 * rms, np.max(np.abs(a - b)) / rms


def _same_peaks(got, want, db, height, tol=1e-3):
    """Identical detected-bin sets, allowing bins whose height is within `tol` dB of the threshold."""
    diff = set(map(int, got)) ^ set(map(int, want))
    for k in diff:
        near_threshold = abs(db[k] - height) <= tol
        near_tie = (0 < k < len(db) - 1) and min(abs(db[k] - db[k - 1]), abs(db[k] - db[k + 1])) <= tol
        assert near_threshold or near_tie, (k, db[k], height)


@pytest.mark.parametrize("n", [8192, 32768, 1 << 16])
def test_spectrum_db_parity(rmx, n):
    iq, _ = synth.tones_block(9 + n, n)
    plan = rmx.Plan(1, n, n)
    S = plan.forward(_cuda(iq[None]))
    ref = oracle.spectrum_db(oracle.forward_fft(oracle.unpack_cu8(iq)))
    _db_close(plan.spectrum_db(S)[0].cpu().numpy(), ref)
    _db_close(plan.spectrum_db(S, shift=True)[0].cpu().numpy(), np.fft.fftshift(ref))


@pytest.mark.parametrize("n", [8192, 32768])
def test_peak_candidates_and_distance_rule(rmx, n):
    iq, _ = synth.tones_block(3 + n, n)
    ref = oracle.spectrum_db(oracle.forward_fft(oracle.unpack_cu8(iq)))
    dev = _cuda(ref)
    for height in (-70.0, float(np.mean(ref) + 10)):
        cand = rmx.threshold_peaks(dev, height)
        want, _ = scipy.signal.find_peaks(ref, height=height)
        assert np.array_equal(cand, want)
    cand = rmx.threshold_peaks(dev, -70.0)
    assert np.array_equal(rmx.select_by_distance(cand, ref[cand], 10), oracle.detect_peaks_fixed(ref))
    mean, med = rmx.mean_median(dev)
    assert abs(float(mean) - float(np.mean(ref))) < 1e-4 and med == np.median(ref)


def test_peak_candidates_plateaus_and_edges(rmx):
    x = np.array([0, 1, 1, 1, 0, 2, 2, 3, 3, 1, 5, 5, 0, 4, 4], dtype=np.float32)
    want, _ = scipy.signal.find_peaks(x, height=0.5)
    assert np.array_equal(rmx.threshold_peaks(_cuda(x), 0.5), want)
    for tiny in (np.zeros(1, np.float32), np.array([1, 2], np.float32), np.array([1, 3, 2], np.float32)):
        want, _ = scipy.signal.find_peaks(tiny, height=0.0)
        assert np.array_equal(rmx.threshold_peaks(_cuda(tiny), 0.0), want)
    odd = np.random.default_rng(0).standard_normal(1001).astype(np.float32)
    assert rmx.mean_median(_cuda(odd))[1] == np.median(odd)


def test_signal_stats_parity(rmx):
    iq, _ = synth.tones_block(21, 1 << 16)
    x = oracle.unpack_cu8(iq)
    ref = oracle.signal_stats(x)
    mp, pk = rmx.signal_stats(_cuda(iq))
    exact = float(np.sum((2 * iq.astype(np.int64) - 255) ** 2)) / 4.0 / (iq.size // 2)
    assert mp == exact                                                  # exact integer arithmetic
    ulp = float(np.spacing(np.float32(ref["peak_amplitude"])))
    assert abs(np.float32(mp) / ref["rms_amplitude"] ** 2 - 1) < 1e-5 and abs(float(pk) - float(ref["peak_amplitude"])) <= ulp
    mp2, pk2 = rmx.signal_stats_c64(_cuda(x))
    assert abs(mp2 / mp - 1) < 1e-6 and pk2 == pk


@pytest.mark.parametrize("nperseg,n_seg", [(4096, 8), (8192, 5), (16384, 9), (32768, 7), (65536, 12), (65536, 70)])
def test_welch_parity(rmx, nperseg, n_seg):
    iq, bins = synth.welch_stream(3, n_seg, nperseg)
    plan = rmx.Plan(n_seg, nperseg, nperseg)
    psd = plan.welch_psd(_cuda(iq), 2.4e6, segments_in_flight=32).cpu().numpy()
    f, ref = oracle.welch_psd(oracle.unpack_cu8(iq), 2.4e6, nperseg)
    assert np.max(np.abs(psd / ref - 1)) < 2e-5
    db = rmx.power_db(_cuda(psd)).cpu().numpy()
    assert np.max(np.abs(db - oracle.welch_db(ref))) < 1e-3


@pytest.mark.parametrize("nperseg", [16384, 32768, 65536])
def test_welch_cluster_kernel_matches_two_pass(rmx, nperseg):
    """nperseg = C*8192 runs as ONE kernel on a cluster of C CTAs (segment held in distributed shared
    memory); a plan created with RMX_PLAN_NO_WELCH_CLUSTER takes the two-pass path through HBM, and so does an
    input pointer that is not 8-byte aligned (the engine then sizes the workspace for many segments in
    flight).  Same transform, different factorisation: the PSDs agree to float32 rounding."""
    from radio_mapper_b200 import _native as nat
    n_seg = 45
    iq, _ = synth.welch_stream(11, n_seg, nperseg)
    a = rmx.Plan(n_seg, nperseg, nperseg, flags=0).welch_psd(_cuda(iq), 2.4e6).cpu().numpy()
    b = rmx.Plan(n_seg, nperseg, nperseg, flags=nat.PLAN_NO_WELCH_CLUSTER).welch_psd(_cuda(iq), 2.4e6).cpu().numpy()
    assert np.max(np.abs(a / b - 1)) < 2e-5
    # unaligned view (2 bytes into a larger buffer): two-pass fallback, same numbers as the forced two-pass plan
    import torch
    big = torch.empty(iq.size + 2, dtype=torch.uint8, device="cuda")
    big[2:].copy_(torch.from_numpy(iq.reshape(-1)))
    plan = rmx.Plan(n_seg, nperseg, nperseg, flags=0)
    c = plan.welch_psd(big[2:], 2.4e6).cpu().numpy()
    assert plan._workspace.numel() > 8 * nperseg * min(n_seg, 8)        # sized for many segments, not one
    assert np.max(np.abs(c / b - 1)) < 2e-6                            # same kernels; only the atomicAdd order differs


def test_cfg2_size_welch_against_oracle_and_parseval(rmx):
    """BASELINE config 2 at full size: 1000 segments x 65536 bins (131 MB of cu8).  Against scipy.signal.welch
    on the same bytes, and Parseval: the PSD integrates to the mean windowed power (size-independent)."""
    nperseg, n_seg, fs = 65536, 1000, 2.4e6
    iq, bins = synth.welch_stream(21, n_seg, nperseg, int(fs))
    plan = rmx.Plan(n_seg, nperseg, nperseg)
    psd = plan.welch_psd(_cuda(iq), fs).cpu().numpy()
    x = oracle.unpack_cu8(iq)
    _, ref = oracle.welch_psd(x, fs, nperseg)
    assert np.max(np.abs(psd / ref - 1)) < 2e-5
    w = scipy.signal.get_window("hann", nperseg).astype(np.float32)
    seg = x.reshape(n_seg, nperseg)
    mean_windowed_power = float(np.mean(np.sum((np.abs(seg) ** 2).astype(np.float64) * (w.astype(np.float64) ** 2), axis=1)))
    integral = float(psd.astype(np.float64).sum()) * fs / nperseg * float(np.sum(w.astype(np.float64) ** 2))
    assert abs(integral / mean_windowed_power - 1) < 1e-5
    db = rmx.power_db(_cuda(psd)).cpu().numpy()
    got = set(rmx.threshold_peaks(_cuda(db), float(np.mean(db) + 10)).tolist())
    assert set(map(int, bins)).issubset(got)


def test_welch_detect_finds_the_tones(rmx):
    from radio_mapper_b200.signal_analyzer import SignalAnalyzer
    nperseg, n_seg = 65536, 40
    iq, bins = synth.welch_stream(5, n_seg, nperseg, 2_400_000)
    res = SignalAnalyzer(verbose=False).welch_detect(iq, 2_400_000, 100.0, nperseg=nperseg, threshold_db=10.0)
    f, ref = oracle.welch_psd(oracle.unpack_cu8(iq), 2_400_000, nperseg)
    ref_db = oracle.welch_db(ref)
    want, _ = scipy.signal.find_peaks(ref_db, height=np.mean(ref_db) + 10)
    _same_peaks(res["peak_bins"], want, ref_db, np.mean(ref_db) + 10)
    assert set(map(int, bins)).issubset(set(map(int, res["peak_bins"])))
    assert res["n_segments"] == n_seg and len(res["peak_bins"]) == len(bins)    # averaging removes the noise peaks


@pytest.mark.parametrize("n", [17, 1000, 10000, 2_048_000 // 8])
def test_bluestein_arbitrary_length(rmx, n):
    from radio_mapper_b200 import bluestein
    rng = np.random.default_rng(n)
    x = oracle.unpack_cu8(rng.integers(0, 256, size=2 * n, dtype=np.uint8))
    got = bluestein.dft(_cuda(x)).cpu().numpy()
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-6


# ---- drop-in host modules against the reference's golden outputs -----------------------------
def test_signal_analyzer_matches_reference_golden(golden_dir):
    from radio_mapper_b200 import signal_analyzer as sa
    g = np.load(os.path.join(golden_dir, "analyze_spectrum.npz"))
    an = sa.SignalAnalyzer(verbose=False)
    x = oracle.unpack_cu8(g["iq"])
    freqs, p_db, peak_freqs = an.analyze_spectrum(x, int(g["sample_rate"]), float(g["center_mhz"]))
    assert np.array_equal(freqs, g["freqs"]) and p_db.dtype == np.float32
    _db_close(p_db, g["p_db"])
    height = float(np.mean(g["p_db"]) + 10)
    got_bins = np.searchsorted(freqs, peak_freqs)
    want_bins = np.searchsorted(freqs, g["peak_freqs"])
    _same_peaks(got_bins, want_bins, g["p_db"], height)
    st = an.calculate_signal_stats(x)
    assert abs(st["power_db"] - float(g["power_db"])) < 1e-4
    # numpy's hypotf-based np.abs is not correctly rounded (here it is 1 ulp high); the GPU returns
    # the correctly rounded sqrt(I^2+Q^2), so the peak amplitude is compared to 1 ulp
    assert abs(float(st["peak_amplitude"]) - float(g["peak_amplitude"])) <= float(np.spacing(np.float32(g["peak_amplitude"])))
    assert st["num_samples"] == int(g["num_samples"])
    assert abs(st["rms_amplitude"] / float(g["rms_amplitude"]) - 1) < 1e-6
    # class alias + module-level functions exist with the reference names
    for name in ("load_iq_data", "analyze_spectrum", "calculate_signal_stats", "plot_spectrum", "analyze_iq_file"):
        assert callable(getattr(sa, name)) and callable(getattr(an, name))


def test_load_iq_data_roundtrip(tmp_path, golden_dir):
    from radio_mapper_b200 import signal_analyzer as sa
    g = np.load(os.path.join(golden_dir, "unpack.npz"))
    path = tmp_path / "iq_capture_100.0MHz_test.bin"
    g["raw"].tofile(path)
    x, fs = sa.SignalAnalyzer(verbose=False).load_iq_data(str(path))
    assert fs == 2048000 and np.array_equal(x.view(np.uint32), g["x_file"].view(np.uint32))
    bad, none = sa.SignalAnalyzer(verbose=False).load_iq_data(str(tmp_path / "missing.bin"))
    assert bad is None and none is None                       # same error contract as the reference
    odd = tmp_path / "odd.bin"
    g["raw"][:-1].tofile(odd)                                 # odd byte count: the reference's I + 1j*Q raises -> (None, None)
    bad, none = sa.SignalAnalyzer(verbose=False).load_iq_data(str(odd))
    assert bad is None and none is None


def _explained_peak_flips(diff, p, height, distance, tol=2e-3):
    """Which bins of `diff` (symmetric difference of two find_peaks(height=, distance=) results computed from dB
    spectra that agree to tol/2) can legitimately differ?  A bin may flip when
      (a) its height is within tol of the height threshold,
      (b) it ties an immediate neighbour within tol (strict-local-maximum test flips),
      (c) a candidate within `distance` bins has a height within tol of it (the greedy rule's priority flips), or
      (d) a bin within `distance` of it is itself an explained flip (its removal / survival cascades).
    Returns the set of explained bins; the caller asserts it equals `diff`."""
    n = len(p)
    diff = set(map(int, diff))
    explained = set()
    for k in diff:
        near_height = abs(p[k] - height) <= tol
        near_tie = 0 < k < n - 1 and min(abs(p[k] - p[k - 1]), abs(p[k] - p[k + 1])) <= tol
        lo, hi = max(0, k - distance + 1), min(n, k + distance)
        near_priority = any(j != k and abs(p[j] - p[k]) <= tol for j in range(lo, hi))
        if near_height or near_tie or near_priority:
            explained.add(k)
    changed = True
    while changed:
        changed = False
        for k in diff - explained:
            if any(abs(j - k) < distance for j in explained):
                explained.add(k)
                changed = True
    return explained


def _rounds_differently(value, decimals, err):
    """True when an error of `err` on `value` can change round(value, decimals)."""
    scaled = value * 10.0 ** decimals
    return abs(scaled - np.floor(scaled) - 0.5) <= err * 10.0 ** decimals + 1e-9


def _check_buoy_detections(got_by_bin, want_by_bin, p):
    """Every difference from the reference's golden detections must be PROVEN threshold-adjacent: no percentage
    allowances.  p = oracle dB spectrum of the block (the GPU's agrees to 1e-3 dB, test_spectrum_db_parity)."""
    med = float(np.median(p))
    diff = set(want_by_bin) ^ set(got_by_bin)
    peak_flips = _explained_peak_flips(diff, p, -70.0, 10)
    for k in diff - peak_flips:
        # not a find_peaks flip: then the confidence must sit on the 0.3 gate (snr within 2e-3 dB of 6 dB)
        assert abs((p[k] - med) / 20.0 - 0.3) <= 1e-4, (k, p[k], med)
    for k in set(want_by_bin) & set(got_by_bin):
        d, w = got_by_bin[k], want_by_bin[k]
        assert d.frequency_mhz == w["frequency_mhz"] and d.signal_type == w["signal_type"]
        if float(d.signal_strength_dbm) != w["signal_strength_dbm"]:
            assert abs(float(d.signal_strength_dbm) - w["signal_strength_dbm"]) <= 0.1001
            assert _rounds_differently(float(p[k]), 1, 1e-3), (k, p[k])                 # 1-decimal rounding boundary
        if float(d.confidence) != w["confidence"]:
            assert abs(float(d.confidence) - w["confidence"]) <= 0.0101
            assert _rounds_differently(min(max((float(p[k]) - med) / 20.0, 0.0), 1.0), 2, 2e-3 / 20.0), (k, p[k])
    return len(diff)


def _buoy_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "buoy_detect.npz"))
    with open(os.path.join(golden_dir, "buoy_detect.json")) as f:
        want = json.load(f)
    fs, fc_hz = int(g["sample_rate"]), int(float(g["center_mhz"]) * 1e6)
    x = oracle.unpack_cu8(g["iq"])
    p = oracle.spectrum_db(oracle.forward_fft(x))
    ora = oracle.score_peaks_buoy(p, oracle.detect_peaks_fixed(p), oracle.freq_axis_hz(len(x), fs, fc_hz), fc_hz)
    assert len(ora) == len(want)
    return g, p, {o["index"]: w for o, w in zip(ora, want)}


def test_buoy_detector_matches_reference_golden(golden_dir):
    """BuoySignalDetector vs the detections the reference's _detect_real_signals produced on the
    same bytes (golden), aligned by FFT bin through the oracle (itself pinned to that golden)."""
    from radio_mapper_b200.detectors import BuoySignalDetector
    g, p, want_by_bin = _buoy_golden(golden_dir)
    bins, got = BuoySignalDetector("BUOY_T", 35.4676, -97.5164).detect_block_indexed(g["iq"], float(g["center_mhz"]))
    _check_buoy_detections(dict(zip(bins, got)), want_by_bin, p)


def test_buoy_node_module_matches_reference_golden(golden_dir):
    """The drop-in `buoy_node.SignalDetector._detect_real_signals(center_freq_mhz)` with the capture step injected
    (the golden was recorded from the reference with rtl_sdr mocked by the same bytes)."""
    import buoy_node as bn
    g, p, want_by_bin = _buoy_golden(golden_dir)
    gps = bn.GPSTimeSource()
    gps.gps_locked, gps.lat, gps.lng = True, 35.4676, -97.5164
    gps.get_precise_timestamp = lambda: ("2025-01-01T00:00:00+00:00", 1735689600000000000)
    seen = []

    def capture(fc_hz, fs, n):
        seen.append((fc_hz, fs, n))
        return g["iq"].tobytes()

    det = bn.SignalDetector("BUOY_T", gps, capture=capture)
    got = det._detect_real_signals(float(g["center_mhz"]))
    assert seen == [(121500000, 2048000, 16384)]                       # the reference's capture parameters (:362-365)
    assert all(isinstance(d, bn.SignalDetection) and d.buoy_id == "BUOY_T" and d.lat == 35.4676 and d.lng == -97.5164
               and d.timestamp_utc == "2025-01-01T00:00:00+00:00" and d.gps_timestamp_ns == 1735689600000000000
               and d.iq_sample_file is None for d in got)
    # align by frequency: frequency_mhz is round(f, 3) of the bin frequency, unique per bin at 62.5 Hz spacing? no --
    # several bins share a rounded MHz value, so align by order instead: both lists are in increasing bin order
    fs, fc_hz = int(g["sample_rate"]), int(float(g["center_mhz"]) * 1e6)
    from radio_mapper_b200.detectors import BuoySignalDetector
    bins, ref_objs = BuoySignalDetector("BUOY_T", 35.4676, -97.5164).detect_block_indexed(g["iq"], float(g["center_mhz"]))
    assert len(bins) == len(got)
    for d, r in zip(got, ref_objs):
        assert (d.frequency_mhz, d.signal_strength_dbm, d.confidence, d.signal_type) == \
               (r.frequency_mhz, r.signal_strength_dbm, r.confidence, r.signal_type)
    _check_buoy_detections(dict(zip(bins, got)), want_by_bin, p)


def _bandwidth_walk_is_marginal(p, k, tol=2e-3):
    """The -3 dB walk of iq_stream_client.py:254-278 from peak k compares p[bin] > p[k] - 3 bin by bin; its result
    can differ between two spectra that agree to tol/2 only if some comparison on the walk is within tol."""
    thr = p[k] - 3.0
    last = len(p) - 1
    left = right = int(k)
    marginal = False
    while left > 0 and p[left] > thr:
        marginal |= abs(p[left] - thr) <= tol
        left -= 1
    marginal |= abs(p[left] - thr) <= tol
    while right < last and p[right] > thr:
        marginal |= abs(p[right] - thr) <= tol
        right += 1
    marginal |= abs(p[right] - thr) <= tol
    return marginal


def _check_stream_detections(got_by_bin, want_by_bin, p):
    diff = set(want_by_bin) ^ set(got_by_bin)
    assert _explained_peak_flips(diff, p, -70.0, 10) == diff, sorted(diff)
    for k in set(want_by_bin) & set(got_by_bin):
        d, w = got_by_bin[k], want_by_bin[k]
        assert d.frequency_mhz == w["frequency_mhz"] and d.signal_type == w["signal_type"]
        assert abs(float(d.signal_strength_dbm) - w["signal_strength_dbm"]) <= 1e-3
        assert abs(float(d.confidence) - w["confidence"]) <= 1e-4
        if d.bandwidth_hz != w["bandwidth_hz"]:
            assert _bandwidth_walk_is_marginal(p, k), (k, d.bandwidth_hz, w["bandwidth_hz"])


def _stream_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "stream_detect.npz"))
    with open(os.path.join(golden_dir, "stream_detect.json")) as f:
        want = json.load(f)
    x = oracle.unpack_cu8(g["iq"])
    p = oracle.spectrum_db(oracle.forward_fft(x))
    peaks = oracle.detect_peaks_fixed(p)
    assert len(peaks) == len(want)
    return g, x, p, dict(zip(map(int, peaks), want))


def test_stream_detector_matches_reference_golden(golden_dir):
    from radio_mapper_b200.detectors import StreamSignalDetector
    g, x, p, want_by_bin = _stream_golden(golden_dir)
    bins, got = StreamSignalDetector("NODE_T").detect_signals_indexed(x, float(g["center_hz"]))
    _check_stream_detections(dict(zip(bins, got)), want_by_bin, p)


def test_iq_stream_client_module_matches_reference_golden(golden_dir):
    """The drop-in `iq_stream_client`: read_iq_samples on a mocked rtl_sdr pipe is bit-exact with the reference's
    (golden unpack.npz x_pipe), and SignalDetector.detect_signals reproduces the reference's detections."""
    import io
    import iq_stream_client as sc

    class _Proc:
        def __init__(self, data):
            self.stdout = io.BytesIO(data)

    u = np.load(os.path.join(golden_dir, "unpack.npz"))
    cap = sc.RealTimeSDRCapture()
    cap.running, cap.capture_process = True, _Proc(u["raw"].tobytes())
    x = cap.read_iq_samples(4096)
    assert x.dtype == np.complex64 and np.array_equal(x.view(np.uint32), u["x_pipe"].view(np.uint32))
    assert cap.read_iq_samples(4096) is None                           # pipe exhausted: short read -> None

    g, xs, p, want_by_bin = _stream_golden(golden_dir)
    cap.capture_process = _Proc(g["iq"].tobytes())
    block = cap.read_iq_samples(len(g["iq"]) // 2)
    assert np.array_equal(block.view(np.uint32), xs.view(np.uint32))
    det = sc.SignalDetector("NODE_T")
    got = det.detect_signals(block, float(g["center_hz"]))
    bins, _ = det.detect_signals_indexed(block, float(g["center_hz"]))
    assert len(bins) == len(got)
    n = len(block)
    for k, d in zip(bins, got):
        assert isinstance(d, sc.SignalDetection) and d.node_id == "NODE_T" and d.detection_method == "fft_peak"
        start = max(0, k - 128)                                     # _extract_signal_samples (:306-316)
        assert len(d.iq_samples) == min(n, start + 256) - start
        assert d.iq_samples[0] == complex(block[start])
    _check_stream_detections(dict(zip(bins, got)), want_by_bin, p)
    assert det.detect_signals(np.zeros(0, np.complex64), 100e6) == []  # errors are logged and yield [] (:249-252)


def test_non_power_of_two_blocks_and_files(tmp_path):
    """Arbitrary block lengths (the reference FFTs whatever it is given) go through Bluestein."""
    from radio_mapper_b200 import signal_analyzer as sa
    from radio_mapper_b200.detectors import BuoySignalDetector, StreamSignalDetector
    n = 10000
    iq, bins = synth.tones_block(77, n)
    x = oracle.unpack_cu8(iq)
    p = oracle.spectrum_db(oracle.forward_fft(x))
    # stream detector
    peaks = oracle.detect_peaks_fixed(p)
    want = oracle.score_peaks_stream(p, peaks, oracle.freq_axis_hz(n, 2048000, 100e6), 2048000)
    kbins, got = StreamSignalDetector("N").detect_signals_indexed(x, 100e6)
    diff = set(kbins) ^ set(map(int, peaks))
    assert _explained_peak_flips(diff, p, -70.0, 10, tol=4e-3) == diff, sorted(diff)
    wb = {w["index"]: w for w in want}
    for k, d in zip(kbins, got):
        if k in wb:
            assert abs(float(d.signal_strength_dbm) - wb[k]["signal_strength_dbm"]) <= 2e-3
            assert d.signal_type == wb[k]["signal_type"]
    # buoy detector straight from bytes
    fc_hz = int(121.5 * 1e6)
    ora = oracle.score_peaks_buoy(p, peaks, oracle.freq_axis_hz(n, 2048000, fc_hz), fc_hz)
    bbins, bgot = BuoySignalDetector("B").detect_block_indexed(iq, 121.5)
    diff = set(bbins) ^ {o["index"] for o in ora}
    flips = _explained_peak_flips(diff, p, -70.0, 10, tol=4e-3)
    for k in diff - flips:                                     # otherwise the confidence sits on the 0.3 gate
        assert abs((p[k] - np.median(p)) / 20.0 - 0.3) <= 2e-4, k
    # whole-file analysis of a capture whose length is not a power of two
    path = tmp_path / "iq_capture_100.0MHz_20250101.bin"
    big, _ = synth.tones_block(78, 250_000)
    big.tofile(path)
    res = sa.SignalAnalyzer(verbose=False).analyze_iq_file(str(path), plot=False)
    xb = oracle.unpack_cu8(big)
    freqs, pdb, peak_freqs = oracle.analyze_spectrum(xb, 2048000, 100.0)
    assert res["center_freq_mhz"] == 100.0 and res["stats"]["num_samples"] == 250_000
    got_bins = np.searchsorted(freqs, res["peak_frequencies"])
    want_bins = np.searchsorted(freqs, peak_freqs)
    _same_peaks(got_bins, want_bins, pdb, float(np.mean(pdb) + 10), tol=2e-3)
    assert abs(res["stats"]["power_db"] - oracle.signal_stats(xb)["power_db"]) < 1e-4


def test_correlate_iq_end_to_end():
    import torch
    from radio_mapper_b200.tdoa_processor import TDOAProcessor, BuoyPosition
    n, fs = 1 << 15, 2048000
    iq0, d0, _ = synth.delayed_buoys(1, 4, n)
    iq1, d1, _ = synth.delayed_buoys(2, 4, n)
    block = np.stack([iq0, iq1], axis=1)                       # [B, W, 2N]
    proc = TDOAProcessor()
    ids = ["B0", "B1", "B2", "B3"]
    for k, b in enumerate(ids):
        proc.register_buoy(BuoyPosition(b, 35.4 + 0.05 * k, -97.5 + 0.03 * (k % 2), 0.0, 1000))
    meas = proc.correlate_iq(torch.from_numpy(block).pin_memory(), ids, fs, 121.5)
    assert len(meas) == 2 * 6
    ref = [oracle.xcorr_pairs_peak(iq0), oracle.xcorr_pairs_peak(iq1)]
    k = 0
    for w in range(2):
        for p, (i, j) in enumerate(oracle.pair_list(4)):
            m = meas[k]; k += 1
            assert (m.buoy1_id, m.buoy2_id) == (ids[i], ids[j]) and m.frequency_mhz == 121.5
            want_ns = oracle.lag_to_tdoa_ns(ref[w]["lag"][p], ref[w]["frac"][p], fs)
            assert abs(m.time_difference_ns - want_ns) <= 1          # frac tolerance 1e-3 samples = 0.5 ns
            assert m.distance_difference_m == m.time_difference_ns / 1e9 * 299792458.0
            assert 0.0 < m.confidence <= 1.0
    # numpy input and a device tensor give the same records
    a = proc.correlate_iq_records(block)
    b = proc.correlate_iq_records(torch.from_numpy(block).cuda())
    assert np.array_equal(a["lag"], b["lag"]) and a.shape == (2, 6)
    with pytest.raises(ValueError):
        proc.correlate_iq(block, ids[:3], fs, 121.5)


# ---- batched block detection ----------------------------------------------
@pytest.mark.parametrize("distance", [0, 1, 3, 10, 57])
def test_find_peaks_batch_matches_scipy(rmx, distance):
    """Every row of a batch against scipy.signal.find_peaks itself: noise rows, rows quantised to few
    levels (plateaus and exact ties in height), a constant row, a monotone row, and both height modes."""
    rng = np.random.default_rng(40 + distance)
    n = 4096
    rows = [rng.standard_normal(n).astype(np.float32) for _ in range(5)]
    rows += [np.round(rng.standard_normal(n) * 2).astype(np.float32) for _ in range(4)]      # plateaus, ties
    rows += [np.zeros(n, np.float32), np.arange(n, dtype=np.float32), -np.arange(n, dtype=np.float32)]
    rows.append(np.repeat(rng.standard_normal(n // 8), 8).astype(np.float32))                # wide plateaus
    db = np.stack(rows)
    for above_mean, height in ((False, 0.25), (True, 0.5)):
        peaks, heights, mean, median = rmx.find_peaks_batch(_cuda(db), height, distance, height_above_mean=above_mean,
                                                            cap=n // 2 + 1)
        for r, row in enumerate(db):
            h = np.float32(np.float32(row.mean(dtype=np.float64)) + height) if above_mean else height
            kw = {"distance": distance} if distance >= 1 else {}
            want, _ = scipy.signal.find_peaks(row, height=h, **kw)
            ties = len(np.unique(row[want])) != len(want) if distance > 1 else False
            if not ties:
                assert np.array_equal(peaks[r], want), (r, distance, above_mean)
                assert np.array_equal(heights[r], row[want])
            else:
                # equal heights: scipy's argsort order is unspecified; both answers obey the distance rule
                assert len(peaks[r]) == 0 or np.all(np.diff(peaks[r]) >= distance)
            assert abs(mean[r] - row.mean(dtype=np.float64)) <= 1e-6 * max(1.0, abs(row.mean()))
            assert median[r] == np.median(row)


def test_find_peaks_batch_overflow_row_falls_back(rmx):
    """A row with more candidates than fit in shared memory (> 16384: only rows longer than 32768 bins can)
    is reported by the kernel and finished by the single-row path."""
    rng = np.random.default_rng(3)
    n = 65536
    saw = np.tile(np.array([0.0, 1.0], np.float32), n // 2) + rng.random(n).astype(np.float32) * 0.01   # ~16k maxima
    quiet = rng.standard_normal(n).astype(np.float32) - 10.0
    db = np.stack([saw, quiet])
    peaks, heights, _, _ = rmx.find_peaks_batch(_cuda(db), 0.5, 10, cap=8192)
    assert len(peaks[0]) > 3000 and len(peaks[1]) == 0
    for r in range(2):
        want, _ = scipy.signal.find_peaks(db[r], height=0.5, distance=10)
        assert np.array_equal(peaks[r], want)
        assert np.array_equal(heights[r], db[r][want])


def test_buoy_detect_blocks_equals_per_block(golden_dir):
    """detect_blocks (one batched launch chain) == detect_block on each block, including the reference's
    golden block; mixed content so blocks differ in peak count."""
    from radio_mapper_b200.detectors import BuoySignalDetector
    g = np.load(os.path.join(golden_dir, "buoy_detect.npz"))
    n2 = g["iq"].size
    rng = np.random.default_rng(8)
    blocks = [g["iq"]]
    for k in range(6):
        u, _ = synth.welch_stream(100 + k, 1, n2 // 2)
        blocks.append(u[:n2])
    blocks.append(rng.integers(0, 256, n2, dtype=np.uint8))
    iq = np.stack(blocks)
    det = BuoySignalDetector("BUOY_T", 35.4676, -97.5164)
    stamps = ["2026-01-01T00:00:%02dZ" % b for b in range(len(blocks))]
    ns = list(range(1000, 1000 + len(blocks)))
    batched = det.detect_blocks(iq, float(g["center_mhz"]), stamps, ns)
    for b in range(len(blocks)):
        single = det.detect_block(iq[b], float(g["center_mhz"]), stamps[b], ns[b])
        assert batched[b] == single, b
    assert sum(len(x) for x in batched) > 0


@pytest.mark.parametrize("n_buoys", [8, 9, 12])
def test_correlate_iq_split_first_window_matches_device_path(n_buoys):
    """From pinned host memory the first window is processed in two halves while its second half is still
    being copied; the records must be bit-identical to the all-at-once device path, for every window."""
    import torch
    from radio_mapper_b200.tdoa_processor import TDoAProcessor
    n, W = 1 << 14, 3
    iq, delays = synth.delayed_buoys_torch(50 + n_buoys, n_buoys, W, n, torch.device("cpu"))
    proc = TDoAProcessor()
    pinned = proc.correlate_iq_records(iq.pin_memory())
    device = proc.correlate_iq_records(iq.cuda())
    pageable = proc.correlate_iq_records(iq.numpy())
    assert pinned.tobytes() == device.tobytes() == pageable.tobytes()
    # the lag-window (one-pass) search goes through the same grouping; its row chunks depend on the number of
    # pairs per launch, so sums are re-associated: same lags, peaks / offsets within the north_star tolerances
    win_p = proc.correlate_iq_records(iq.pin_memory(), max_lag=700)
    win_d = proc.correlate_iq_records(iq.cuda(), max_lag=700)
    assert np.array_equal(win_p["lag"], win_d["lag"]) and np.array_equal(win_p["lag"], pinned["lag"])
    assert np.max(np.abs(win_p["peak"] / win_d["peak"] - 1)) <= 1e-4
    assert np.max(np.abs(win_p["frac"] - win_d["frac"])) <= 1e-3
    pairs = [(i, j) for i in range(n_buoys) for j in range(i + 1, n_buoys)]
    for w in range(W):
        assert list(pinned["lag"][w]) == [int(delays[w, j] - delays[w, i]) for i, j in pairs]


def test_stream_detect_blocks_arrays_equal_per_block():
    """Batched stream detector (device bandwidth walk included) against detect_signals on each block."""
    from radio_mapper_b200.detectors import StreamSignalDetector
    nb, n = 6, 8192                                              # iq_stream_client.py:459 block size
    u, _ = synth.welch_stream(31, nb, n)
    blocks = u.reshape(nb, 2 * n)
    det = StreamSignalDetector("NODE_T")
    fc = 100.0e6
    batched = det.detect_blocks_arrays(blocks, fc)
    for b in range(nb):
        bins, dets = det.detect_signals_indexed(oracle.unpack_cu8(blocks[b]), fc)
        kb, f_hz, power, bw, conf = batched[b]
        assert list(kb) == bins
        assert np.array_equal(power, np.array([d.signal_strength_dbm for d in dets], np.float32))
        assert np.array_equal(f_hz / 1e6, np.array([d.frequency_mhz for d in dets]))
        assert np.array_equal(bw, np.array([d.bandwidth_hz for d in dets]))
        assert np.allclose(conf, np.array([d.confidence for d in dets], np.float32), rtol=0, atol=1e-7)
