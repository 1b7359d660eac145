"""Oracle: pairwise cross-correlation, lag search, and the dt->dd seam.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference has no cross-correlation (tdoa_processor.py:20 imports
scipy.signal.correlate and never calls it; Documents/ROADMAP.md:33), so stages
(3)/(4) are defined here from that imported primitive, following the conventions
the reference does fix:
  * pair order i<j in list order           tdoa_processor.py:156-157
  * dt = t_buoy2 - t_buoy1, positive when buoy2 hears the signal later    :51, :166
  * dd = dt/1e9 * 299792458.0              tdoa_processor.py:141, :169-170
"""
import math
import numpy as np
import scipy.signal

from .spectrum import unpack_cu8

SPEED_OF_LIGHT = 299792458.0            # tdoa_processor.py:141


def pair_list(n):
    """All i<j in enumeration order of tdoa_processor.py:156-157."""
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def xcorr_full(x_i, x_j, dtype=np.complex64):
    """c[k] = sum_n x_j[n+k] * conj(x_i[n]), k = -(N-1)..N-1.

    `scipy.signal.correlate(x_j, x_i, 'full', 'fft')` — if buoy j hears the waveform d
    samples later than buoy i the peak is at lag +d (matches tdoa_processor.py:51).
    dtype=complex128 gives the float64 'truth' used to arbitrate near-ties.
    """
    a = np.asarray(x_j, dtype=dtype)
    b = np.asarray(x_i, dtype=dtype)
    c = scipy.signal.correlate(a, b, mode="full", method="fft")
    lags = scipy.signal.correlation_lags(len(a), len(b), mode="full")
    return c, lags


def peak_lag(c, lags, max_lag=None):
    """argmax |c| (first maximum, like np.argmax) + 3-point parabolic vertex on |c|.

    Returns (lag:int, peak:float, frac:float).  frac = 0 when the maximum sits on the
    edge of the searched range.  `max_lag` restricts the search to |lag| <= max_lag.
    """
    mag = np.abs(c)
    if max_lag is not None:
        keep = np.abs(lags) <= max_lag
        idx = np.flatnonzero(keep)
    else:
        idx = np.arange(len(c))
    k = idx[int(np.argmax(mag[idx]))]
    frac = 0.0
    if idx[0] < k < idx[-1]:
        ym, y0, yp = (np.float64(mag[k - 1]), np.float64(mag[k]), np.float64(mag[k + 1]))
        den = ym - 2.0 * y0 + yp
        if den != 0.0:
            frac = float(0.5 * (ym - yp) / den)
    return int(lags[k]), float(mag[k]), frac


def xcorr_pairs_peak(iq_u8, pairs=None, max_lag=None, dtype=np.complex64):
    """iq_u8: uint8[B, 2N].  Returns structured array (lag, peak, frac) per pair."""
    iq_u8 = np.asarray(iq_u8, dtype=np.uint8)
    x = [unpack_cu8(row) for row in iq_u8]
    if pairs is None:
        pairs = pair_list(len(x))
    out = np.zeros(len(pairs), dtype=[("lag", "i4"), ("peak", "f4"), ("frac", "f4")])
    for p, (i, j) in enumerate(pairs):
        c, lags = xcorr_full(x[i], x[j], dtype=dtype)
        out[p] = peak_lag(c, lags, max_lag=max_lag)
    return out


def lag_to_tdoa_ns(lag, frac, sample_rate):
    """Seam into TDoAMeasurement.time_difference_ns (int, tdoa_processor.py:50):
    round((lag+frac)/fs * 1e9)."""
    return int(round((float(lag) + float(frac)) / float(sample_rate) * 1e9))


def timing_confidence(acc1_ns, acc2_ns):
    """tdoa_processor.py:200-210."""
    return min(math.exp(-math.sqrt(acc1_ns ** 2 + acc2_ns ** 2) / 100000), 1.0)


def tdoa_measurements(detections, timing_accuracy_ns):
    """Restates tdoa_processor.py:156-193 on plain tuples.

    detections: list of (buoy_id, frequency_mhz, gps_timestamp_ns, confidence);
    timing_accuracy_ns: {buoy_id: ns}.  Returns list of
    (buoy1, buoy2, dt_ns, dd_m, confidence, frequency_mhz).
    """
    out = []
    for i in range(len(detections)):
        for j in range(i + 1, len(detections)):
            b1, f1, t1, c1 = detections[i]
            b2, f2, t2, c2 = detections[j]
            if abs(f1 - f2) > 0.01:
                continue
            if b1 not in timing_accuracy_ns or b2 not in timing_accuracy_ns:
                continue
            dt = t2 - t1
            dd = dt / 1e9 * SPEED_OF_LIGHT
            conf = min(c1, c2) * timing_confidence(timing_accuracy_ns[b1], timing_accuracy_ns[b2])
            out.append((b1, b2, dt, dd, conf, f1))
    return out
