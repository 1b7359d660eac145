#!/usr/bin/env python3
"""BASELINE config 5: sub-sample delay accuracy against SNR (8 buoys, long windows).

    python tests/tools/snr_sweep.py [LOG2_SAMPLES=22] [BUOYS=8]

For SNR in {+20, +10, 0, -10, -20} dB per buoy: synthetic cu8 with known integer + fractional delays is generated
on the device, correlated by librmx, and the measured delay (lag + parabolic offset) is compared with the truth.
Prints one JSON line per SNR: integer lags exact?, RMS / max error of the sub-sample delay in samples, ms."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from radio_mapper_b200 import engine, synth

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = 1 << logn
dev = torch.device("cuda")
rng = np.random.default_rng(5)
frac = rng.uniform(-0.45, 0.45, size=B)
frac[0] = 0.0
plan = engine.Plan(B, N)
pairs_h = engine.pair_table(B)
pairs = torch.from_numpy(pairs_h).to(dev)
for snr in (20, 10, 0, -10, -20):
    iq, d = synth.delayed_buoys_torch(100 + snr, B, 1, N, dev, snr_db=float(snr), frac_delays=frac)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec = engine.peaks_to_numpy(plan.xcorr_pairs_peak(plan.forward(iq[:, 0, :]), pairs))
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    true = np.array([(d[0, j] + frac[j]) - (d[0, i] + frac[i]) for i, j in pairs_h])
    got = rec["lag"] + rec["frac"]
    err = got - true
    print(json.dumps({"snr_db": snr, "samples": N, "buoys": B, "pairs": len(pairs_h), "passes": plan.pass_lengths,
                      "integer_lags_exact": bool(np.array_equal(rec["lag"], np.rint(true).astype(np.int64))),
                      "delay_err_rms_samples": float(np.sqrt(np.mean(err ** 2))), "delay_err_max_samples": float(np.max(np.abs(err))),
                      "ms": round(ms, 2)}), flush=True)
    del iq
