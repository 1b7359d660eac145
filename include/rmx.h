/* rmx.h — C ABI of the B200-native radio-mapper hot path (librmx.so).
 *
 * The reference (physiii/radio-mapper) has NO native/FFI boundary for this path: its
 * arithmetic is in-process numpy/scipy calls inside Python modules.  Each entry point below
 * therefore cites the reference *Python* lines whose arithmetic it replaces; the host-side
 * drop-in (radio_mapper_b200/tdoa_processor.py, signal_analyzer.py, detectors.py) keeps the
 * reference's class / function signatures and calls these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the name says host;
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it;
 *   - return 0 on success, negative on error; `rmx_last_error()` returns the thread-local text;
 *   - the library allocates device memory only inside a plan (twiddle tables, window);
 *   - no exceptions cross the boundary; no torch types appear in signatures.
 *
 * Spectrum layout: `rmx_fft_forward_cu8` writes spectra in the plan's "digit-transposed" order
 * (see rmx_plan_layout); `rmx_xcorr_pairs_peak` consumes that order; `rmx_spectrum_db` and
 * `rmx_spectrum_natural` convert to natural bin order for anything that needs it.
 */
#ifndef RMX_H_
#define RMX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMX_VERSION 201

#if defined(__GNUC__)
#define RMX_API __attribute__((visibility("default")))
#else
#define RMX_API
#endif

typedef struct rmx_plan rmx_plan;

/* per-pair result of the lag search (16 bytes) */
typedef struct rmx_peak {
    int32_t lag;   /* integer lag of max |c|, samples; >0: signal j arrived later (tdoa_processor.py:51) */
    float peak;    /* |c[lag]|, unnormalised like scipy.signal.correlate */
    float frac;    /* parabolic vertex offset in (-0.5, 0.5), 0 at the range edge */
    float pad;     /* |c|^2 found by the tile search (diagnostic) */
} rmx_peak;

typedef struct rmx_complex64 { float re, im; } rmx_complex64;
typedef struct rmx_pair { int32_t i, j; } rmx_pair;      /* signal indices, i != j */

/* signal statistics (signal_analyzer.py:92-99) */
typedef struct rmx_stats {
    double mean_power;     /* mean |x|^2 (exact integer arithmetic on the cu8 samples) */
    float peak_amplitude;  /* max |x| */
    float pad;
} rmx_stats;

enum {
    RMX_OK = 0,
    RMX_ERR_ARG = -1,        /* bad argument */
    RMX_ERR_UNSUPPORTED = -2,/* size not supported by the kernels */
    RMX_ERR_CUDA = -3,       /* CUDA runtime error (text in rmx_last_error) */
    RMX_ERR_WORKSPACE = -4   /* workspace too small */
};

RMX_API const char* rmx_last_error(void);
RMX_API int rmx_version(void);

/* Stage 1 — cu8 unpack:  out[n] = (in[2n]-127.5) + i*(in[2n+1]-127.5)   (bit-exact)
 * replaces buoy_node.py:392-398, iq_stream_client.py:149-157, signal_analyzer.py:28-36 */
RMX_API int rmx_unpack_cu8(const uint8_t* in, rmx_complex64* out, size_t n_samples, void* stream);

/* Plan: n_signals signals of n_samples complex samples each, zero-padded to fft_len (a power of
 * two >= 16, >= n_samples).  flags: 0 for the defaults, or an OR of the developer switches below — fixed for
 * the life of the plan, so a process A/Bs kernel variants by creating two plans (nothing on the launch path
 * reads the environment). */
#define RMX_PLAN_TWIDDLE_IN_COL   0x01u  /* inter-pass twiddles on the input of the column pass, not the row pass output */
#define RMX_PLAN_NO_TMA           0x02u  /* arg-max pass through per-thread strided loads instead of the TMA-fed kernel */
#define RMX_PLAN_NO_PAIR_RUN      0x04u  /* one pair per CTA in the 4096-point row pass (no X_i-stationary walk) */
#define RMX_PLAN_NO_WELCH_CLUSTER 0x08u  /* Welch PSD through the two-pass path even where the cluster kernel applies */
#define RMX_PLAN_ROW_E8           0x10u  /* two-pass plans on 2048-point rows held 8 values per thread (more resident CTAs) */
#define RMX_PLAN_ROW_LOGN(n)      (((unsigned)(n) & 0x1fu) << 8)  /* force log2 of the row length of multi-pass plans */
RMX_API int rmx_plan_create(rmx_plan** plan, int n_signals, size_t n_samples, size_t fft_len, unsigned flags);
/* Tuning knobs (defaults are the measured best on B200): "pair_run" = 8 | 16 pairs walked by one CTA of the
 * X_i-stationary row pass; "pair_prefetch" = 0 | 1 next X_j row by bulk copy into shared memory;
 * "fwd_group_bytes" = forward passes run over groups of signals whose spectra fit this many bytes, so a pass
 * reads the previous one's output from L2 (0 = all signals per launch); "welch_clusters" = resident clusters of
 * the Welch kernel (0 = occupancy query); "fuse_outer" = 0 | 1 three-pass plans (fft_len > 2^23): middle and outer inverse pass + arg-max as separate
 * launches through the workspace (default) or in one persistent kernel through an L2-resident scratch ring (measured
 * slightly slower on B200: see DESIGN.md);
 * "fwd_tma" = 1 | 0 forward pass 0 through the persistent kernel that stages
 * the raw cu8 tiles in shared memory by 3-D TMA box loads (default; taken for whole-row windows and 16-byte aligned
 * input, otherwise -- and with 0 -- the per-thread 128-bit staging kernel runs).
 * Measured-slower alternatives of the 4096-point row pass, kept selectable for A/B runs:
 * "pair_store" = 0 | 1 | 2 finished rows leave by per-thread stores (default), through a dedicated staging buffer + one
 * bulk copy (1: takes the prefetch buffer's place), or staged in the exchange buffer + one bulk copy (2: keeps the prefetch);
 * "pair_xi_early" = 1 | 0 a new X_i row is loaded one pair ahead of its first use (default; neutral: -3 % on the row pass
 * of 3 buoys, +-1 % at 8 and 64); "pair_xi_smem" = 0 | 1 the stationary X_i row in shared memory instead of registers (takes the prefetch buffer's place);
 * "pair_groups" = 0 | 2 | 3 one CTA of 2 or 3 warp groups that hand the FP32 pipe round on a ring of named barriers
 * instead of independent CTAs; "pair_ctas" = 4 | 5 | 6 resident CTAs per SM of the RMX_PLAN_ROW_E8 row kernel. */
RMX_API int rmx_plan_set_option(rmx_plan* plan, const char* name, long long value);
RMX_API int rmx_plan_destroy(rmx_plan* plan);
/* number of passes and their lengths n_t (outermost first); returns n_passes */
RMX_API int rmx_plan_layout(const rmx_plan* plan, int32_t* pass_lengths, int cap);
/* bytes of workspace needed to correlate `n_pairs` pairs in one chunk */
RMX_API size_t rmx_plan_workspace_bytes(const rmx_plan* plan, int n_pairs);
/* restrict the lag search to [-max_lag, +max_lag]; max_lag < 0 restores the full range
 * -(n_samples-1) .. n_samples-1 of scipy.signal.correlate(mode='full') */
RMX_API int rmx_plan_set_max_lag(rmx_plan* plan, long long max_lag);

/* With a lag window much narrower than a row of the innermost pass (max_lag < 2048) the search
 * runs as ONE pass over the spectra (row iFFT + twiddled accumulation of the kept lags, no
 * correlation workspace).  force_full != 0 disables that path (the full inverse transform with a
 * masked arg-max is used instead); results are identical up to float rounding. */
RMX_API int rmx_plan_set_search_mode(rmx_plan* plan, int force_full);

/* Stages 1+2 — fused unpack + zero-pad + batched forward FFT of all signals.
 * iq: signal s starts at iq + s*signal_stride_bytes (0 = densely packed, 2*n_samples) and holds
 * 2*n_samples bytes; spectra: complex64[n_signals][fft_len] (plan layout).
 * replaces fft(iq_samples) of buoy_node.py:401 / iq_stream_client.py:187 / signal_analyzer.py:63 */
RMX_API int rmx_fft_forward_cu8(const rmx_plan* plan, const uint8_t* iq, size_t signal_stride_bytes,
                                rmx_complex64* spectra, void* stream);

/* Same transform for signals that are already complex64 (signal s at x + s*signal_stride_elems,
 * 0 = densely packed); used where the reference hands complex samples around
 * (signal_analyzer.py:47 analyze_spectrum, iq_stream_client.py:181 detect_signals). */
RMX_API int rmx_fft_forward_c64(const rmx_plan* plan, const rmx_complex64* x, size_t signal_stride_elems,
                                rmx_complex64* spectra, void* stream);

/* plan layout -> natural bin order (out may not alias in) */
RMX_API int rmx_spectrum_natural(const rmx_plan* plan, const rmx_complex64* spectra, rmx_complex64* out,
                         int n_signals, void* stream);

/* Stages 3+4 — for every pair (i, j): c = ifft(X_j * conj(X_i)) == scipy.signal.correlate(x_j, x_i,
 * 'full', 'fft'); arg-max of |c| over the lag range, 3-point parabolic vertex.  One launch per
 * pass covers all pairs of a chunk.  pairs: device rmx_pair[n_pairs]; out: device rmx_peak[n_pairs].
 * (absent in the reference — tdoa_processor.py:20 imports correlate but only subtracts timestamps,
 * :166; this produces the time difference that line computes) */
RMX_API int rmx_xcorr_pairs_peak(const rmx_plan* plan, const rmx_complex64* spectra, const rmx_pair* pairs,
                         int n_pairs, rmx_peak* out, void* workspace, size_t workspace_bytes, void* stream);

/* Full correlation instead of its peak: out[p][m] = ifft(X_j conj X_i)[m], natural order
 * (lag = m for m < L/2, m - L otherwise).  out: complex64[n_pairs][fft_len]. */
RMX_API int rmx_xcorr_full(const rmx_plan* plan, const rmx_complex64* spectra, const rmx_pair* pairs, int n_pairs,
                           rmx_complex64* out, void* stream);

/* Arbitrary-length DFT (the reference FFTs whole captures of any length, signal_analyzer.py:62-63)
 * by Bluestein's chirp-z on top of the power-of-two engine:
 *   prepare: a[i] = x[i]*w[i] (i < n, else 0), chirp_circ[m] = w[min(m, padded_len-m)] for |m| < n,
 *            w[i] = exp(-i*pi*i^2/n);  then conv = ifft(fft(a) * conj(fft(chirp_circ)))  (rmx_xcorr_full)
 *   finish : X[k] = w[k]*conv[k], k < n. */
RMX_API int rmx_bluestein_prepare(const rmx_complex64* x, size_t n, size_t padded_len, rmx_complex64* a,
                                  rmx_complex64* chirp_circ, void* stream);
RMX_API int rmx_bluestein_finish(const rmx_complex64* conv, size_t n, rmx_complex64* out, void* stream);
/* out[k'] = 20*log10(|x[k]| + 1e-12) on a natural-order vector; shift != 0: k' = (k + n/2) mod n */
RMX_API int rmx_abs_db(const rmx_complex64* x, size_t n, float* out_db, int shift, void* stream);

/* Stage 5a — dB spectrum in natural order: out[k] = 20*log10(|X[k]| + 1e-12); shift != 0 applies
 * fftshift.  replaces buoy_node.py:405, iq_stream_client.py:191, signal_analyzer.py:64-67 */
RMX_API int rmx_spectrum_db(const rmx_plan* plan, const rmx_complex64* spectra, float* out_db, int n_signals,
                    int shift, void* stream);

/* Stage 5b — Welch PSD (Hann, no overlap, no detrend, two-sided, density scaling): the plan's
 * n_signals are the segments, n_samples == fft_len == nperseg.  psd: float[fft_len], natural order.
 * == scipy.signal.welch(x, fs, 'hann', nperseg, 0, detrend=False, return_onesided=False)
 * nperseg = 16384 / 32768 / 65536 (and iq 8-byte aligned) runs as ONE kernel on thread-block clusters of
 * 2 / 4 / 8 CTAs that hold each segment in distributed shared memory; it touches only the first
 * 4*fft_len bytes of the workspace.  Other sizes take two passes through a spectra workspace of
 * 8*fft_len bytes per segment in flight. */
RMX_API int rmx_welch_psd(rmx_plan* plan, const uint8_t* iq, float* psd, double sample_rate,
                  void* workspace, size_t workspace_bytes, void* stream);
RMX_API size_t rmx_welch_workspace_bytes(const rmx_plan* plan, int segments_in_flight);
/* which path rmx_welch_psd takes for this plan and input pointer: 1 = one cluster kernel (no spectra workspace),
 * 0 = two passes (size the workspace for as many segments in flight as memory allows) */
RMX_API int rmx_welch_path(const rmx_plan* plan, const uint8_t* iq);

/* out[k] = 10*log10(in[k] + eps) */
RMX_API int rmx_power_db(const float* in, float* out, size_t n, float eps, void* stream);

/* Stage 5c — peak candidates: scipy's _local_maxima_1d (plateau midpoints, end points excluded)
 * with db[k] >= height, compacted in unspecified order into idx[0..min(count,cap)).
 * replaces the first two stages of find_peaks (buoy_node.py:411-415, signal_analyzer.py:75) */
RMX_API int rmx_threshold_peaks(const float* db, int n, float height, int32_t* idx, int32_t* count, int cap,
                        void* stream);

/* HOST helper: find_peaks' greedy minimum-distance rule on sorted candidate positions.
 * keep[i] = 1 if candidate i survives.  All pointers are host pointers. */
RMX_API int rmx_select_by_distance_host(const int32_t* positions, const float* heights, int n, int distance,
                                uint8_t* keep);

/* mean(db) (double accumulation) and median(db) (np.median: mean of the two middle order
 * statistics for even n).  out: device float[2] = {mean, median}.  workspace: >= 4 KiB.
 * replaces np.mean(...) signal_analyzer.py:75 and np.median(...) buoy_node.py:427 */
RMX_API int rmx_mean_median(const float* db, int n, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Stage 5c/5d batched — block detection for many dB spectra at once (one CTA per row):
 * scipy.signal.find_peaks(db_row, height=H, distance=D) with H = height (height_mode 0; buoy_node.py:411-415,
 * iq_stream_client.py:197-201) or H = mean(db_row) + height (height_mode 1; signal_analyzer.py:75), plus the row's
 * mean and median (buoy_node.py:427).  idx/heights: [n_rows][cap] kept peaks in ascending bin order and their
 * dB values; count[row] = number of kept peaks (only the first cap are stored), or -(candidates) when a row has
 * more than 16384 candidates before the distance rule (rows longer than 32768 bins can) (take that row through rmx_threshold_peaks instead);
 * stats: [n_rows][2] = mean, median.  distance <= 1 disables the distance rule.
 * Optional gates applied after the distance rule (buoy_node.py:423-433): gate_dc_bins > 0 drops peaks whose bin is
 * closer than that to DC (bins k and n-k; the reference's |f - fc| < 10 kHz), gate_conf_min > 0 drops peaks with
 * clip((p - median)/20, 0, 1) < gate_conf_min (float32 arithmetic). */
RMX_API int rmx_find_peaks_batch(const float* db, int n_rows, int n, size_t row_stride, float height, int height_mode,
                         int distance, int gate_dc_bins, float gate_conf_min, int32_t* idx, float* heights,
                         int32_t* count, int cap, float* stats, void* stream);

/* Stage 5e batched — the -3 dB bandwidth walk of iq_stream_client.py:254-278 for the peaks rmx_find_peaks_batch
 * returned: width[row][k] = right - left with left/right walked away from peak k while db > db[peak] - drop_db
 * (stopping at bins 0 and n-1); bandwidth_hz = width * fs / n. */
RMX_API int rmx_peak_bandwidth_batch(const float* db, int n_rows, int n, size_t row_stride, const int32_t* idx,
                             const int32_t* count, int cap, float drop_db, int32_t* width, void* stream);

/* Per-launch timing with CUDA events on the launching stream (off by default).  enable != 0 clears
 * earlier records and starts recording; collect synchronises the recorded events and returns the
 * number of distinct kernel names written to out[0..cap). */
typedef struct rmx_prof_entry {
    char name[32];
    int32_t launches;
    float total_ms;
} rmx_prof_entry;
RMX_API int rmx_profile_enable(rmx_plan* plan, int enable);
RMX_API int rmx_profile_collect(rmx_plan* plan, rmx_prof_entry* out, int cap);

/* signal statistics straight from cu8 (signal_analyzer.py:92-99); out: device rmx_stats */
RMX_API int rmx_signal_stats(const uint8_t* iq, size_t n_samples, rmx_stats* out, void* stream);

/* same statistics for complex64 samples (calculate_signal_stats, signal_analyzer.py:88);
 * workspace: 16 bytes of device scratch */
RMX_API int rmx_signal_stats_c64(const rmx_complex64* x, size_t n_samples, rmx_stats* out, void* workspace, void* stream);

/* Exact per-signal energy for normalising correlation peaks: out[s] = sum_n (2I-255)^2 + (2Q-255)^2
 * = 4 * sum |x_s[n]|^2 (unpack of buoy_node.py:392-398 in integers).  out: device uint64[n_signals]. */
RMX_API int rmx_signal_energy(const uint8_t* iq, size_t signal_stride_bytes, int n_signals, size_t n_samples,
                              unsigned long long* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RMX_H_ */
