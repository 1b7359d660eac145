"""Oracle: cu8 unpack, FFT, dB spectrum, peak detection, scoring, stats, Welch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it restates; paths are relative to the reference checkout.
Pinned to this image's numpy 2.3 / scipy 1.18 (the reference pins only lower
bounds, requirements.txt:2-3).
"""
import numpy as np
import scipy.fft
import scipy.signal


def unpack_cu8(raw):
    """uint8[2N] interleaved I,Q -> complex64[N].

    buoy_node.py:392-398, iq_stream_client.py:149-157, signal_analyzer.py:28-36:
    `a = raw.astype(float32) - 127.5; a[0::2] + 1j*a[1::2]`.
    (`float32 + 1j*float32` is complex64 under numpy>=2 promotion rules.)
    """
    a = np.asarray(raw, dtype=np.uint8).astype(np.float32) - np.float32(127.5)
    return (a[0::2] + 1j * a[1::2]).astype(np.complex64)


def forward_fft(x):
    """Unwindowed, unpadded forward DFT, complex64 (buoy_node.py:401,
    iq_stream_client.py:187 use scipy.fft.fft; signal_analyzer.py:63 uses np.fft.fft,
    which is also complex64 for complex64 input on numpy>=2)."""
    return scipy.fft.fft(np.asarray(x, dtype=np.complex64))


def forward_fft_np(x):
    """signal_analyzer.py:63 calls np.fft.fft (numpy's own pocketfft build): complex64 out on
    numpy>=2 (complex128 on numpy 1.x — this oracle pins the image's numpy 2.3).  It differs
    from scipy.fft.fft in the last ulp, so the analyzer path keeps its own entry point."""
    return np.fft.fft(np.asarray(x, dtype=np.complex64))


def spectrum_db(X):
    """20*log10(|X| + 1e-12) (buoy_node.py:405, iq_stream_client.py:191,
    signal_analyzer.py:67).  float32 for complex64 input."""
    return 20 * np.log10(np.abs(X) + 1e-12)


def freq_axis_hz(n, sample_rate, center_hz):
    """fftfreq(n, 1/fs) + fc (buoy_node.py:402,408; iq_stream_client.py:188,194)."""
    return scipy.fft.fftfreq(n, 1.0 / sample_rate) + center_hz


def freq_axis_mhz_shifted(n, sample_rate, center_mhz):
    """fftshift(fftfreq)/1e6 + fc_mhz (signal_analyzer.py:70-72)."""
    return np.fft.fftshift(np.fft.fftfreq(n, 1 / sample_rate)) / 1e6 + center_mhz


def detect_peaks_fixed(p_db, height=-70, distance=10):
    """find_peaks(P, height=-70, distance=10) (buoy_node.py:411-415,
    iq_stream_client.py:197-201)."""
    peaks, _ = scipy.signal.find_peaks(p_db, height=height, distance=distance)
    return peaks


def detect_peaks_mean(p_db):
    """find_peaks(P, height=mean(P)+10) (signal_analyzer.py:75)."""
    peaks, _ = scipy.signal.find_peaks(p_db, height=np.mean(p_db) + 10)
    return peaks


def strict_local_maxima(p):
    """Index set scipy's `_local_maxima_1d` returns (plateau midpoints, endpoints
    excluded) — the first stage of find_peaks; used to check the GPU candidate kernel."""
    peaks, _ = scipy.signal.find_peaks(p)
    return peaks


def classify_buoy(freq_mhz):
    """buoy_node.py:342-355 band table (argument in MHz)."""
    if freq_mhz in (121.5, 243.0):
        return "emergency"
    if 118.0 <= freq_mhz <= 136.0:
        return "aviation"
    if 144.0 <= freq_mhz <= 148.0:
        return "amateur"
    if 156.0 <= freq_mhz <= 162.0:
        return "marine"
    if 406.0 <= freq_mhz <= 406.1:
        return "emergency_beacon"
    return "unknown"


def classify_stream(freq_hz):
    """iq_stream_client.py:280-304 band table (argument in Hz)."""
    if abs(freq_hz - 121500000) < 1000 or abs(freq_hz - 243000000) < 1000:
        return "emergency"
    if 155000000 <= freq_hz <= 156000000:
        return "public_safety"
    if 406000000 <= freq_hz <= 406100000:
        return "emergency"
    if 88000000 <= freq_hz <= 108000000:
        return "fm_radio"
    if 118000000 <= freq_hz <= 136000000:
        return "aviation"
    if 144000000 <= freq_hz <= 148000000:
        return "amateur"
    if 420000000 <= freq_hz <= 450000000:
        return "amateur"
    return "unknown"


def score_peaks_buoy(p_db, peaks, abs_freqs_hz, center_hz):
    """Per-peak scoring loop of buoy_node.py:418-455.

    Returns a list of dicts (frequency_mhz rounded to 3, strength rounded to 1,
    confidence rounded to 2, signal_type) for the peaks that survive the +-10 kHz DC skip
    (:423) and the confidence>=0.3 gate (:432).
    """
    out = []
    noise_floor = np.median(p_db)                       # :427 (loop-invariant)
    for k in peaks:
        f_hz = abs_freqs_hz[k]
        if abs(f_hz - center_hz) < 10000:               # :423
            continue
        snr = p_db[k] - noise_floor                     # :428
        conf = min(max(snr / 20.0, 0.0), 1.0)           # :429
        if conf < 0.3:                                  # :432
            continue
        f_mhz = f_hz / 1e6
        # round() is applied to the numpy float32 scalars, as the reference does (:447-452)
        out.append(dict(index=int(k), frequency_mhz=round(f_mhz, 3),
                        signal_strength_dbm=float(round(p_db[k], 1)),
                        confidence=float(round(conf, 2)),
                        signal_type=classify_buoy(f_mhz)))
    return out


def estimate_bandwidth(p_db, k, sample_rate):
    """-3 dB walk of iq_stream_client.py:254-278."""
    thr = p_db[k] - 3.0
    lo = hi = int(k)
    n = len(p_db)
    while lo > 0 and p_db[lo] > thr:
        lo -= 1
    while hi < n - 1 and p_db[hi] > thr:
        hi += 1
    return (hi - lo) * (sample_rate / n)


def score_peaks_stream(p_db, peaks, abs_freqs_hz, sample_rate):
    """Per-peak loop of iq_stream_client.py:204-217 (no DC skip, no lower clamp)."""
    out = []
    noise_floor = np.median(p_db)
    for k in peaks:
        snr = p_db[k] - noise_floor
        out.append(dict(index=int(k), frequency_mhz=float(abs_freqs_hz[k] / 1e6),
                        signal_strength_dbm=float(p_db[k]),
                        bandwidth_hz=float(estimate_bandwidth(p_db, k, sample_rate)),
                        confidence=float(min(snr / 20.0, 1.0)),
                        signal_type=classify_stream(abs_freqs_hz[k])))
    return out


def signal_stats(x):
    """signal_analyzer.py:92-99: mean |x|^2 -> dB, max |x|, rms."""
    p = np.mean(np.abs(x) ** 2)
    return dict(power_db=10 * np.log10(p + 1e-12),
                peak_amplitude=np.max(np.abs(x)),
                rms_amplitude=np.sqrt(np.mean(np.abs(x) ** 2)),
                num_samples=len(x))


def analyze_spectrum(x, sample_rate, center_mhz):
    """signal_analyzer.py:61-76 without the prints: (frequencies_mhz, P_shifted_db, peak_freqs)."""
    X = np.fft.fftshift(forward_fft_np(x))
    p = spectrum_db(X)
    freqs = freq_axis_mhz_shifted(len(x), sample_rate, center_mhz)
    peaks = detect_peaks_mean(p)
    return freqs, p, freqs[peaks]


def welch_psd(x, sample_rate, nperseg=65536):
    """Stage-5 definition (absent in the reference): two-sided Welch PSD,
    Hann window, no overlap, no detrend, density scaling, via scipy.signal.welch."""
    f, pxx = scipy.signal.welch(np.asarray(x, dtype=np.complex64), fs=sample_rate, window="hann",
                                nperseg=nperseg, noverlap=0, detrend=False,
                                return_onesided=False, scaling="density")
    return f, pxx


def welch_db(pxx):
    """10*log10 of *power* (+1e-12 guard scaled like the reference's amplitude guard)."""
    return 10 * np.log10(pxx + 1e-24)
