#!/usr/bin/env python3
"""BASELINE config 2: Welch PSD + threshold detection, 2.4 Msps, 64k-bin FFT, 1000 segments."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from radio_mapper_b200 import engine, synth
from radio_mapper_b200.signal_analyzer import SignalAnalyzer

nperseg, n_seg, fs = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 2_400_000
inflight = int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = torch.Generator(device="cuda"); g.manual_seed(1)
iq = torch.randint(96, 160, (2 * nperseg * n_seg,), dtype=torch.uint8, device="cuda", generator=g)
plan = engine.Plan(n_seg, nperseg, nperseg)
for _ in range(3):
    psd = plan.welch_psd(iq, fs, segments_in_flight=inflight)
torch.cuda.synchronize()
plan.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    psd = plan.welch_psd(iq, fs, segments_in_flight=inflight)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
prof = plan.profile_collect()
an = SignalAnalyzer(verbose=False)
t0 = time.perf_counter(); res = an.welch_detect(iq, fs, 100.0, nperseg=nperseg); torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({"metric": "welch_samples_per_sec", "value": n_seg * nperseg / (ms * 1e-3), "ms": ms, "passes": plan.pass_lengths,
                  "segments_in_flight": inflight, "hbm_frac_of_6453": (2.0 * n_seg * nperseg + 4 * nperseg) / (ms * 1e-3) / 1e9 / 6453.1,
                  "kernels_ms": {k: v[1] / K for k, v in prof.items()}, "welch_detect_s": t1 - t0}))
