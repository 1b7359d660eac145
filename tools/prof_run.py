#!/usr/bin/env python3
"""Small driver for ncu: `python tools/prof_run.py BUOYS LOG2_SAMPLES [ITERS]` runs forward + correlate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from radio_mapper_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
max_lag = int(sys.argv[4]) if len(sys.argv) > 4 else -1
iq, delays = synth.delayed_buoys_torch(7, B, 1, N, torch.device("cuda"))
plan = engine.Plan(B, N)
if max_lag >= 0:
    plan.set_max_lag(max_lag)
pairs_h = engine.pair_table(B)
pairs = torch.from_numpy(pairs_h).cuda()
for _ in range(iters):
    S = plan.forward(iq[:, 0, :])
    rec = plan.xcorr_pairs_peak(S, pairs)
torch.cuda.synchronize()
got = engine.peaks_to_numpy(rec)["lag"]
want = np.array([delays[0, j] - delays[0, i] for i, j in pairs_h])
print("passes", plan.pass_lengths, "lags ok:", np.array_equal(got, want))
