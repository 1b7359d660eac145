#!/usr/bin/env python3
"""Small, fast coverage of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from radio_mapper_b200 import engine, synth, bluestein
import oracle

def run(n, B, max_lag=None, force_full=False):
    iq, d, _ = synth.delayed_buoys(n, B, n, max_delay=min(300, n // 4))
    plan = engine.Plan(B, n)
    if max_lag is not None:
        plan.set_max_lag(max_lag)
    plan.set_search_mode(force_full)
    S = plan.forward(torch.from_numpy(iq).cuda())
    got = engine.peaks_to_numpy(plan.xcorr_pairs_peak(S, torch.from_numpy(engine.pair_table(B)).cuda()))
    ref = oracle.xcorr_pairs_peak(iq, max_lag=max_lag)
    print(n, B, max_lag, plan.pass_lengths, np.array_equal(got["lag"], ref["lag"]), flush=True)

for n in (100, 2048, 4096, 5000, 1 << 14, 1 << 16, (1 << 16) - 3, 1 << 18):
    run(n, 3)
run(1 << 16, 3, 342); run(1 << 16, 3, 342, True); run(1 << 16, 3, 900); run(1 << 18, 3, 1500)
run(1 << 22, 2)                      # [512, 8192]
run(1 << 23, 2)                      # three passes
u, bins = synth.welch_stream(3, 6, 8192)
print("welch", float(engine.Plan(6, 8192, 8192).welch_psd(torch.from_numpy(u).cuda(), 2.4e6).sum()))
u, bins = synth.welch_stream(3, 5, 65536)
print("welch64k", float(engine.Plan(5, 65536, 65536).welch_psd(torch.from_numpy(u).cuda(), 2.4e6, segments_in_flight=2).sum()))
x = oracle.unpack_cu8(np.random.default_rng(0).integers(0, 256, 2 * 1000, dtype=np.uint8))
print("bluestein", float(bluestein.spectrum_db(torch.from_numpy(x).cuda()).sum()))
db = torch.from_numpy(np.random.default_rng(1).standard_normal(5000).astype(np.float32)).cuda()
print("peaks", len(engine.threshold_peaks(db, 0.0)), engine.mean_median(db), engine.signal_stats(torch.from_numpy(u[:20000]).cuda()))
print("unpack", engine.unpack_cu8(torch.from_numpy(u[:4098]).cuda()).abs().sum().item())
torch.cuda.synchronize(); print("sanitize_run done")
