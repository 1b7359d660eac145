#!/usr/bin/env python3
"""Soak check (developer aid): the same window correlated ITERS times must give bit-identical peak records every
time -- the kernels are deterministic (no atomics on the data path of the correlate stage), so any difference is a
race (mbarrier phases, bulk-copy ordering, exchange-buffer reuse).

    python tools/soak.py BUOYS LOG2_SAMPLES ITERS ['name=value,...' plan options]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from radio_mapper_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
opts = {k.strip(): int(v, 0) for k, v in (kv.split("=") for kv in filter(None, (sys.argv[4] if len(sys.argv) > 4 else "").split(",")))}
iq, delays = synth.delayed_buoys_torch(11, B, 1, N, torch.device("cuda"))
pairs_h = engine.pair_table(B)
pairs = torch.from_numpy(pairs_h).cuda()
want = np.array([delays[0, j] - delays[0, i] for i, j in pairs_h])
plan = engine.Plan(B, N, options=opts)
S = plan.forward(iq[:, 0, :])
first_S = S.clone()
first = plan.xcorr_pairs_peak(S, pairs).clone()
bad_records = bad_spectra = 0
for it in range(iters):
    S = plan.forward(iq[:, 0, :])
    rec = plan.xcorr_pairs_peak(S, pairs)
    bad_spectra += int(not torch.equal(S.view(torch.int32) if S.dtype != torch.complex64 else torch.view_as_real(S).view(torch.int32),
                                       torch.view_as_real(first_S).view(torch.int32)))
    bad_records += int(not torch.equal(rec, first))
torch.cuda.synchronize()
got = engine.peaks_to_numpy(first)
print(json.dumps({"B": B, "N": N, "passes": plan.pass_lengths, "options": opts, "iterations": iters, "lags_ok": bool(np.array_equal(got["lag"], want)),
                  "iterations_with_different_records": bad_records, "iterations_with_different_spectra": bad_spectra}), flush=True)
sys.exit(1 if (bad_records or bad_spectra) else 0)
