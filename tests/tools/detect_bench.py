#!/usr/bin/env python3
"""Block detection throughput: the reference's per-block path
(buoy_node.py:391-433: unpack, FFT, dB, find_peaks(height=-70, distance=10), median, scoring) on batches of
raw cu8 blocks.

    python tools/detect_bench.py [N_BLOCKS] [LOG2_SAMPLES] [ITERS]

GPU arm: BuoySignalDetector.detect_blocks on a pinned host uint8[n_blocks, 2N] (H2D, batched FFT + dB +
find_peaks/median kernels, peak lists D2H, host scoring) and, for comparison, detect_block in a loop.
CPU arm: the oracle (numpy/scipy, one thread) on a bounded sample of the same blocks.  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import oracle
from radio_mapper_b200 import synth
from radio_mapper_b200.detectors import BuoySignalDetector

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 15)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
fs, fc_mhz = 2_048_000, 100.0
u, _ = synth.welch_stream(9, nb, n, fs)                   # noise + CW tones, [nb*2n]
iq = torch.from_numpy(np.ascontiguousarray(u.reshape(nb, 2 * n))).pin_memory()
det = BuoySignalDetector("BUOY_T", 35.4676, -97.5164, fs)
stamps = ["2026-01-01T00:00:00Z"] * nb
ns = [0] * nb
res = det.detect_blocks_arrays(iq, fc_mhz)                # warm-up (plan, kernels)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    res = det.detect_blocks_arrays(iq, fc_mhz)            # host cu8 -> per-block detection arrays
torch.cuda.synchronize()
batched_s = (time.perf_counter() - t0) / iters
t0 = time.perf_counter()
objs = det.detect_blocks(iq, fc_mhz, stamps, ns)          # + one SignalDetection dataclass per detection
objects_s = time.perf_counter() - t0
k = min(nb, 32)
det.detect_block(iq[0], fc_mhz, stamps[0], 0)
t0 = time.perf_counter()
for b in range(k):
    single = det.detect_block(iq[b], fc_mhz, stamps[b], 0)
torch.cuda.synchronize()
loop_s = (time.perf_counter() - t0) / k
fc_hz = int(fc_mhz * 1e6)
t0 = time.perf_counter()
for b in range(k):
    x = oracle.unpack_cu8(iq[b].numpy())
    p = oracle.spectrum_db(oracle.forward_fft(x))
    ora = oracle.score_peaks_buoy(p, oracle.detect_peaks_fixed(p), oracle.freq_axis_hz(n, fs, fc_hz), fc_hz)
cpu_s = (time.perf_counter() - t0) / k
print(json.dumps({"workload": "%d blocks x %d samples (buoy_node block detection)" % (nb, n),
                  "batched_blocks_per_s": nb / batched_s, "batched_samples_per_s": nb * n / batched_s,
                  "batched_ms_per_batch": 1e3 * batched_s, "with_dataclasses_blocks_per_s": nb / objects_s,
                  "per_block_call_blocks_per_s": 1.0 / loop_s,
                  "cpu_oracle_blocks_per_s_1_thread": 1.0 / cpu_s, "cpu_sample_blocks": k,
                  "detections_per_block_mean": float(np.mean([len(r[0]) for r in res])),
                  
                  "h2d_bytes_per_batch": int(iq.numel())}))
