"""B200-native drop-in for the reference's `signal_analyzer` module.

Same function names, arguments and return values as /root/reference/signal_analyzer.py
(load_iq_data :14, analyze_spectrum :47, calculate_signal_stats :88, plot_spectrum :114,
analyze_iq_file :136), exposed both at module level and as methods of `SignalAnalyzer` (the
class name BASELINE.json's north_star uses).  The numeric stages run on the GPU through
librmx (cu8 unpack, FFT, dB spectrum, mean / threshold / local-maximum detection, signal
statistics); frequency axes are closed-form host arithmetic.  New batched entry points:
`SignalAnalyzer.analyze_cu8` (fused from raw bytes) and `SignalAnalyzer.welch_detect`.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np


def _engine():
    from . import engine            # imports torch + librmx; raises without the CUDA library
    return engine


def _torch():
    import torch
    return torch


def plt_backend_is_set() -> bool:
    """True when the caller (or the environment) has already chosen a matplotlib backend."""
    import os
    import sys
    return "matplotlib.pyplot" in sys.modules or bool(os.environ.get("MPLBACKEND"))


class SignalAnalyzer:
    """Spectrum analysis of RTL-SDR cu8 captures on the GPU."""

    def __init__(self, sample_rate: int = 2048000, device=None, verbose: bool = True):
        self.sample_rate = sample_rate
        self.device = device
        self.verbose = verbose
        self._plans = {}

    def _say(self, *a):
        if self.verbose:
            print(*a)

    def _plan(self, n_signals: int, n_samples: int, fft_len: int):
        key = (n_signals, n_samples, fft_len)
        if key not in self._plans:
            self._plans[key] = _engine().Plan(n_signals, n_samples, fft_len, device=self.device)
        return self._plans[key]

    # ---- reference API ---------------------------------------------------------------------
    def load_iq_data(self, filename, sample_rate=2048000):
        """Raw cu8 file -> (complex64 samples, sample_rate); (None, None) on error like the
        reference (:43-45).  The unpack runs on the GPU and is bit-exact with :28-36."""
        try:
            raw = np.fromfile(filename, dtype=np.uint8)
            torch = _torch()
            if raw.size % 2:
                # the reference's i_samples + 1j*q_samples raises on the I/Q length mismatch of an odd byte
                # count and lands in its except branch (:43-45): same contract here
                raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,)"
                                 % ((raw.size + 1) // 2, raw.size // 2))
            n = raw.size // 2
            dev = _engine().unpack_cu8(torch.from_numpy(raw[: 2 * n]).to(self._device()))
            samples = dev.cpu().numpy()
            self._say(f"Loaded {len(samples)} complex samples from {filename}")
            self._say(f"Duration: {len(samples) / sample_rate:.2f} seconds")
            return samples, sample_rate
        except Exception as exc:
            print(f"Error loading IQ data: {exc}")
            return None, None

    def analyze_spectrum(self, complex_samples, sample_rate, center_freq_mhz):
        """(frequencies MHz, power spectrum dB (fftshifted), peak frequencies) as :47-86:
        P = 20*log10(|fftshift(fft(x))| + 1e-12), peaks = find_peaks(P, height=mean(P)+10)."""
        eng, torch = _engine(), _torch()
        x = np.ascontiguousarray(complex_samples, dtype=np.complex64)
        n = x.size
        db_dev = eng.spectrum_db_c64(torch.from_numpy(x).to(self._device()), shift=True, plans=self._plans)
        mean, _ = eng.mean_median(db_dev)
        peaks = eng.threshold_peaks(db_dev, float(mean) + 10.0)
        power_spectrum = db_dev.cpu().numpy()
        freq_shifted = np.fft.fftshift(np.fft.fftfreq(n, 1 / sample_rate))
        frequencies = (freq_shifted / 1e6) + center_freq_mhz
        peak_freqs = frequencies[peaks]
        peak_powers = power_spectrum[peaks]
        self._say("\nSpectrum Analysis Results:")
        self._say(f"  Frequency range: {frequencies[0]:.2f} to {frequencies[-1]:.2f} MHz")
        self._say(f"  Peak frequencies found: {len(peak_freqs)}")
        for i, (f, p) in enumerate(zip(peak_freqs, peak_powers)):
            self._say(f"    Peak {i + 1}: {f:.3f} MHz ({p:.1f} dB)")
        return frequencies, power_spectrum, peak_freqs

    def calculate_signal_stats(self, complex_samples):
        """power_db = 10*log10(mean|x|^2 + 1e-12), peak |x|, rms (:88-112)."""
        eng, torch = _engine(), _torch()
        x = np.ascontiguousarray(complex_samples, dtype=np.complex64)
        mean_power, peak = eng.signal_stats_c64(torch.from_numpy(x).to(self._device()))
        mean_power = np.float32(mean_power)
        stats = {
            "power_db": 10 * np.log10(mean_power + 1e-12),
            "peak_amplitude": peak,
            "rms_amplitude": np.sqrt(mean_power),
            "num_samples": len(x),
        }
        self._say("\nSignal Statistics:")
        self._say(f"  Signal Power: {stats['power_db']:.2f} dB")
        self._say(f"  Peak Amplitude: {stats['peak_amplitude']:.2f}")
        self._say(f"  RMS Amplitude: {stats['rms_amplitude']:.2f}")
        self._say(f"  Total Samples: {stats['num_samples']}")
        return stats

    def plot_spectrum(self, frequencies, power_spectrum, center_freq_mhz, output_file=None):
        """Plot helper (:114-134).  matplotlib is optional; without it this raises ImportError."""
        import matplotlib
        if output_file and not plt_backend_is_set():
            matplotlib.use("Agg")                              # headless file output only; never override a caller's backend
        import matplotlib.pyplot as plt
        plt.figure(figsize=(12, 6))
        plt.plot(frequencies, power_spectrum)
        plt.xlabel("Frequency (MHz)")
        plt.ylabel("Power (dB)")
        plt.title(f"Power Spectrum - Center Frequency: {center_freq_mhz} MHz")
        plt.grid(True, alpha=0.3)
        plt.axvline(x=center_freq_mhz, color="red", linestyle="--", alpha=0.7, label="Center Freq")
        plt.legend()
        if output_file:
            plt.savefig(output_file, dpi=150, bbox_inches="tight")
            self._say(f"Spectrum plot saved to {output_file}")
        else:
            plt.show()
        plt.close()

    def analyze_iq_file(self, filename, plot: bool = True):
        """load -> stats -> spectrum (-> plot) for one capture (:136-176)."""
        # Same parse as the reference (:140-146): the 2nd '_'-separated token, falling back to 100.0.
        # (For the stock 'iq_capture_<f>MHz_<ts>.bin' names that token is 'capture', so the
        # reference — and therefore this drop-in — reports 100.0 MHz.)
        center_freq_mhz = None
        if "MHz" in filename:
            try:
                center_freq_mhz = float(filename.split("_")[1].replace("MHz", ""))
            except Exception:
                center_freq_mhz = 100.0
        self._say(f"\n{'=' * 60}\nAnalyzing: {filename}\nCenter Frequency: {center_freq_mhz} MHz\n{'=' * 60}")
        samples, sample_rate = self.load_iq_data(filename)
        if samples is None:
            return None
        stats = self.calculate_signal_stats(samples)
        frequencies, power_spectrum, peak_freqs = self.analyze_spectrum(samples, sample_rate, center_freq_mhz)
        plot_filename = filename.replace(".bin", "_spectrum.png")
        if plot:
            try:
                self.plot_spectrum(frequencies, power_spectrum, center_freq_mhz, plot_filename)
            except ImportError:
                plot_filename = None
        return {"filename": filename, "center_freq_mhz": center_freq_mhz, "stats": stats,
                "peak_frequencies": peak_freqs, "spectrum_plot": plot_filename}

    # ---- new batched entry points ------------------------------------------------------------
    def _device(self):
        torch = _torch()
        return torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)

    def analyze_cu8(self, iq_u8, sample_rate, center_freq_mhz):
        """Fused path from raw bytes: iq_u8 uint8[B, 2N] (N a power of two) -> dict with the
        shifted dB spectra (device tensor [B, N]), the per-block mean and the peak bins."""
        eng, torch = _engine(), _torch()
        t = torch.as_tensor(iq_u8)
        if t.ndim == 1:
            t = t[None]
        t = t.to(self._device())
        n = t.shape[1] // 2
        plan = self._plan(t.shape[0], n, n)
        db = plan.spectrum_db(plan.forward(t), shift=True)
        out = []
        for b in range(t.shape[0]):
            mean, median = eng.mean_median(db[b])
            out.append(dict(mean_db=float(mean), median_db=float(median),
                            peak_bins=eng.threshold_peaks(db[b], float(mean) + 10.0)))
        freqs = np.fft.fftshift(np.fft.fftfreq(n, 1 / sample_rate)) / 1e6 + center_freq_mhz
        return dict(power_db=db, frequencies_mhz=freqs, blocks=out)

    def welch_detect(self, iq_u8, sample_rate, center_freq_mhz, nperseg: int = 65536, threshold_db: float = 10.0,
                     segments_in_flight: Optional[int] = None):
        """Welch PSD (Hann, no overlap, density scaling == scipy.signal.welch) of a cu8 stream
        and threshold detection over the frequency bins: bins that are local maxima of the dB
        spectrum and exceed mean + threshold_db.  iq_u8: uint8[2*W*nperseg] (host or device).

        Returns dict(psd (device float32[nperseg], natural order), psd_db, frequencies_hz,
        peak_bins, peak_freqs_hz, mean_db)."""
        eng, torch = _engine(), _torch()
        t = torch.as_tensor(iq_u8).reshape(-1)
        n_seg = t.numel() // (2 * nperseg)
        if n_seg < 1:
            raise ValueError("need at least one full segment of %d samples" % nperseg)
        t = t[: n_seg * 2 * nperseg].to(self._device(), non_blocking=True)
        plan = self._plan(n_seg, nperseg, nperseg)
        psd = plan.welch_psd(t, float(sample_rate), segments_in_flight=segments_in_flight)
        psd_db = eng.power_db(psd, 1e-24)
        mean, _ = eng.mean_median(psd_db)
        bins = eng.threshold_peaks(psd_db, float(mean) + float(threshold_db))
        freqs = np.fft.fftfreq(nperseg, 1.0 / sample_rate) + center_freq_mhz * 1e6
        return dict(psd=psd, psd_db=psd_db, frequencies_hz=freqs, peak_bins=bins, peak_freqs_hz=freqs[bins],
                    mean_db=float(mean), n_segments=n_seg)


# ---- module-level functions with the reference's names -------------------------------------
_default = SignalAnalyzer()


def load_iq_data(filename, sample_rate=2048000):
    return _default.load_iq_data(filename, sample_rate)


def analyze_spectrum(complex_samples, sample_rate, center_freq_mhz):
    return _default.analyze_spectrum(complex_samples, sample_rate, center_freq_mhz)


def calculate_signal_stats(complex_samples):
    return _default.calculate_signal_stats(complex_samples)


def plot_spectrum(frequencies, power_spectrum, center_freq_mhz, output_file=None):
    return _default.plot_spectrum(frequencies, power_spectrum, center_freq_mhz, output_file)


def analyze_iq_file(filename):
    return _default.analyze_iq_file(filename)


if __name__ == "__main__":
    import sys
    print("=== RTL-SDR IQ Data Analysis Tool (B200) ===")
    iq_files = [f for f in os.listdir(".") if f.startswith("iq_capture_") and f.endswith(".bin")]
    if not iq_files:
        print("No IQ capture files found. Run sdr_capture.py first.")
        sys.exit(1)
    for name in iq_files:
        try:
            res = analyze_iq_file(name)
            if res:
                print(f"File: {res['filename']}  peaks: {len(res['peak_frequencies'])}  power: {res['stats']['power_db']:.2f} dB")
        except Exception as exc:
            print(f"Error analyzing {name}: {exc}")
