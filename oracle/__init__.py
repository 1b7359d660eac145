"""CPU oracle for the radio-mapper hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, with numpy/scipy, the arithmetic of the reference
(physiii/radio-mapper) for the path named in BASELINE.json: cu8 unpack ->
FFT -> dB spectrum -> peak detection, plus the pairwise cross-correlation /
lag search / Welch stages the reference only *imports* primitives for.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  The product (`radio_mapper_b200`)
never does: its hot path is CUDA-only and fails loudly without the
extension.

Parity status (see DESIGN.md "Oracle"):
  * a1..a9 (unpack, FFT, dB, find_peaks, scoring, stats, axes, bandwidth,
    pair enumeration / dt->dd) are PINNED: `tests/golden/*.npz` were produced
    by running the reference's own functions (`tests/golden/make_golden.py`)
    and `tests/test_oracle_golden.py` checks this oracle against them.
  * a10..a12 (cross-correlation, argmax + parabolic, Welch) have NO
    implementation in the reference ("parity unpinned" by the reference);
    they are defined here by the scipy primitives the reference imports
    (`scipy.signal.correlate`, `scipy.signal.welch`, `find_peaks`).
"""
from .spectrum import (  # noqa: F401
    unpack_cu8, forward_fft, forward_fft_np, spectrum_db, freq_axis_hz, freq_axis_mhz_shifted,
    detect_peaks_fixed, detect_peaks_mean, score_peaks_buoy, score_peaks_stream,
    estimate_bandwidth, signal_stats, analyze_spectrum, classify_buoy, classify_stream,
    welch_psd, welch_db, strict_local_maxima,
)
from .xcorr import (  # noqa: F401
    pair_list, xcorr_full, peak_lag, xcorr_pairs_peak, lag_to_tdoa_ns, tdoa_measurements,
    timing_confidence, SPEED_OF_LIGHT,
)
