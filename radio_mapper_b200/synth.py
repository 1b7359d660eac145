"""Seeded synthetic cu8 IQ generators (host/numpy) shared by tests, smoke and bench.

The wire format is rtl_sdr's raw output: unsigned 8-bit interleaved I,Q,I,Q...
(reference: Code/src/rtl_sdr.c:95, sdr_capture.py:58), centred on 127.5
(buoy_node.py:393).  Generation is plain numpy so the identical bytes can be fed
to the CPU oracle and to the CUDA path.
"""
from __future__ import annotations

import numpy as np

RMS_LSB = 30.0


def _lowpass_noise(rng, n, sample_rate, bandwidth_hz):
    """Complex Gaussian noise band-limited to +-bandwidth_hz/2 (brick-wall in frequency)."""
    spec = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    f = np.fft.fftfreq(n, 1.0 / sample_rate)
    spec[np.abs(f) > bandwidth_hz / 2] = 0
    s = np.fft.ifft(spec)
    return s / np.sqrt(np.mean(np.abs(s) ** 2))


def quantize_cu8(x, rms_lsb=RMS_LSB):
    """complex array (unit RMS) -> uint8[2N] interleaved, clip(rint(x*rms + 127.5), 0, 255)."""
    x = np.asarray(x)
    out = np.empty(x.shape[:-1] + (2 * x.shape[-1],), dtype=np.uint8)
    scale = rms_lsb / np.sqrt(2.0)
    out[..., 0::2] = np.clip(np.rint(x.real * scale + 127.5), 0, 255).astype(np.uint8)
    out[..., 1::2] = np.clip(np.rint(x.imag * scale + 127.5), 0, 255).astype(np.uint8)
    return out


def delayed_buoys(seed, n_buoys, n_samples, sample_rate=2_048_000, bandwidth_hz=200_000.0,
                  snr_db=10.0, max_delay=342, delays=None, frac_delays=None):
    """One window of B buoys hearing the same band-limited source with known delays.

    Buoy b receives  g_b * s[n - d_b] + w_b[n]  (SURVEY §8d).  Returns
    (iq_u8[B, 2N], delays[B] int, frac[B] float): the expected peak lag of pair (i, j)
    is (d_j + f_j) - (d_i + f_i)  (sign convention of tdoa_processor.py:51).
    """
    rng = np.random.default_rng(seed)
    pad = max_delay + 2
    total = n_samples + 2 * pad
    s = _lowpass_noise(rng, total, sample_rate, bandwidth_hz)
    if delays is None:
        delays = rng.integers(-max_delay, max_delay + 1, size=n_buoys)
    delays = np.asarray(delays, dtype=np.int64)
    if frac_delays is None:
        frac_delays = np.zeros(n_buoys)
    frac_delays = np.asarray(frac_delays, dtype=np.float64)
    gains = rng.uniform(0.7, 1.0, size=n_buoys)
    noise_amp = 10.0 ** (-snr_db / 20.0)
    out = np.empty((n_buoys, n_samples), dtype=np.complex128)
    if np.any(frac_delays != 0):
        S = np.fft.fft(s)
        f = np.fft.fftfreq(total)
    for b in range(n_buoys):
        if frac_delays[b] != 0:
            sb = np.fft.ifft(S * np.exp(-2j * np.pi * f * frac_delays[b]))
        else:
            sb = s
        start = pad - int(delays[b])
        w = (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples)) / np.sqrt(2)
        x = gains[b] * sb[start:start + n_samples] + noise_amp * w
        out[b] = x / np.sqrt(np.mean(np.abs(x) ** 2))
    return quantize_cu8(out), delays, frac_delays


def tones_block(seed, n_samples, sample_rate=2_048_000, tone_bins=None, tone_snr_db=None,
                rms_lsb=20.0):
    """Noise floor + CW tones at exact FFT bins (for the PSD / detection stages).

    Returns (iq_u8[2N], tone_bins).  Tones are placed away from DC so that the
    reference's +-10 kHz skip (buoy_node.py:423) does not hide them.
    """
    rng = np.random.default_rng(seed)
    if tone_bins is None:
        lo = int(0.05 * n_samples)
        tone_bins = np.sort(rng.choice(np.arange(lo, n_samples - lo), size=5, replace=False))
    tone_bins = np.asarray(tone_bins)
    if tone_snr_db is None:
        tone_snr_db = rng.uniform(15, 40, size=len(tone_bins))
    n = np.arange(n_samples)
    x = (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples)) / np.sqrt(2)
    for b, snr in zip(tone_bins, tone_snr_db):
        # per-bin SNR: tone power N*A^2 against noise power per bin ~ 1
        amp = 10.0 ** (snr / 20.0) / np.sqrt(n_samples)
        x = x + amp * np.exp(2j * np.pi * (b * n % n_samples) / n_samples + 1j * rng.uniform(0, 2 * np.pi))
    x = x / np.sqrt(np.mean(np.abs(x) ** 2))
    return quantize_cu8(x, rms_lsb), tone_bins


def welch_stream(seed, n_segments, nperseg, sample_rate=2_400_000, n_tones=5):
    """W*L samples: noise + CW tones at known bins of the nperseg-point PSD (config 2)."""
    rng = np.random.default_rng(seed)
    total = n_segments * nperseg
    lo = int(0.05 * nperseg)
    bins = np.sort(rng.choice(np.arange(lo, nperseg - lo), size=n_tones, replace=False))
    snrs = rng.uniform(15, 40, size=n_tones)
    out = np.empty(2 * total, dtype=np.uint8)
    chunk = max(1, (1 << 22) // nperseg) * nperseg
    for start in range(0, total, chunk):
        m = min(chunk, total - start)
        n = np.arange(start, start + m)
        x = (rng.standard_normal(m) + 1j * rng.standard_normal(m)) / np.sqrt(2)
        for b, snr in zip(bins, snrs):
            amp = 10.0 ** (snr / 20.0) / np.sqrt(nperseg)
            x = x + amp * np.exp(2j * np.pi * ((b * n) % nperseg) / nperseg)
        x = x / np.sqrt(1.0 + np.sum((10.0 ** (snrs / 20.0)) ** 2) / nperseg)
        out[2 * start:2 * (start + m)] = quantize_cu8(x, 20.0)
    return out, bins


def delayed_buoys_torch(seed, n_buoys, n_windows, n_samples, device, sample_rate=2_048_000,
                        bandwidth_hz=200_000.0, snr_db=10.0, max_delay=342, frac_delays=None):
    """GPU-side generator for full-size bench inputs (not bit-identical to `delayed_buoys`).

    Returns (iq_u8 CUDA uint8[B, W, 2N], delays int64[W, B]).  torch.fft is used here only to
    shape the synthetic source; it is data generation, not the product path.
    frac_delays: optional per-buoy sub-sample delays in (-0.5, 0.5), applied to the band-limited source as a
    phase ramp (buoy b hears the source delays[w, b] + frac_delays[b] samples late).
    """
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    pad = max_delay + 2
    total = n_samples + 2 * pad
    out = torch.empty((n_buoys, n_windows, 2 * n_samples), dtype=torch.uint8, device=device)
    delays = torch.randint(-max_delay, max_delay + 1, (n_windows, n_buoys), generator=g, device=device)
    delays_h = delays.cpu()
    freqs = torch.fft.fftfreq(total, 1.0 / sample_rate, device=device)
    keep = (freqs.abs() <= bandwidth_hz / 2)
    noise_amp = 10.0 ** (-snr_db / 20.0)
    scale = RMS_LSB / (2.0 ** 0.5)
    for w in range(n_windows):
        spec = torch.randn(total, 2, generator=g, device=device)
        spec = torch.view_as_complex(spec) * keep
        s0 = torch.fft.ifft(spec)
        norm = s0.abs().pow(2).mean().sqrt()
        s = s0 / norm
        for b in range(n_buoys):
            if frac_delays is not None and frac_delays[b] != 0.0:
                ramp = torch.exp(-2j * torch.pi * (freqs / sample_rate) * float(frac_delays[b]))
                s = torch.fft.ifft(spec * ramp) / norm
            elif frac_delays is not None:
                s = s0 / norm
            start = pad - int(delays_h[w, b])
            gain = 0.7 + 0.3 * float(torch.rand(1, generator=g, device=device))
            nz = torch.view_as_complex(torch.randn(n_samples, 2, generator=g, device=device)) * (noise_amp / 2.0 ** 0.5)
            x = gain * s[start:start + n_samples] + nz
            x = x / x.abs().pow(2).mean().sqrt()
            q = torch.view_as_real(x).mul(scale).add(127.5).round().clamp_(0, 255).to(torch.uint8)
            out[b, w] = q.reshape(-1)
        del spec, s
    return out, delays_h.numpy()
