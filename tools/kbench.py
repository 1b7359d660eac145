#!/usr/bin/env python3
"""Per-kernel timing of forward + correlate on one window (developer aid for A/B runs).

    [RMX_LIB_PATH=...] python tools/kbench.py [BUOYS] [LOG2_SAMPLES] [ITERS] [MAX_LAG]

Prints one line: total pair-stage ms per window and the per-kernel ms (CUDA events from librmx)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from radio_mapper_b200 import engine, synth, _native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
max_lag = int(sys.argv[4]) if len(sys.argv) > 4 else -1
iq, delays = synth.delayed_buoys_torch(7, B, 1, N, torch.device("cuda"))
plan = engine.Plan(B, N)
if max_lag >= 0:
    plan.set_max_lag(max_lag)
pairs_h = engine.pair_table(B)
pairs = torch.from_numpy(pairs_h).cuda()
for _ in range(2):
    S = plan.forward(iq[:, 0, :])
    rec = plan.xcorr_pairs_peak(S, pairs)
torch.cuda.synchronize()
got = engine.peaks_to_numpy(rec)["lag"]
want = np.array([delays[0, j] - delays[0, i] for i, j in pairs_h])
plan.profile(True)
for _ in range(iters):
    S = plan.forward(iq[:, 0, :])
    rec = plan.xcorr_pairs_peak(S, pairs)
torch.cuda.synchronize()
prof = plan.profile_collect()
per = {k: round(v[1] / iters, 4) for k, v in prof.items()} if isinstance(prof, dict) else prof
print(json.dumps({"lib": os.path.basename(_native.LIB_PATH), "B": B, "N": N, "passes": plan.pass_lengths,
                  "lags_ok": bool(np.array_equal(got, want)), "ms_per_window": per}))
