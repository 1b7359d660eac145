#!/usr/bin/env python3
"""A/B of plan options / flags in ONE process (developer aid): per-kernel ms per window for each setting.

    python tools/optbench.py BUOYS LOG2_SAMPLES ITERS 'name=value,...;name=value,...;...' [IDLE_MS]

IDLE_MS > 0 synchronises and sleeps that long between iterations (per-kernel times then show what the kernels do
on a GPU that is not at its power limit; ms_per_window_total includes the sleeps and is meaningless).

Each ';'-separated setting is a comma list of plan options (rmx_plan_set_option) and/or `flags=<int>`; an empty
setting is the default plan.  Prints one JSON line per setting."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from radio_mapper_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
settings = [x.strip() for x in (sys.argv[4] if len(sys.argv) > 4 else "").split(";")]
idle_ms = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
iq, delays = synth.delayed_buoys_torch(7, B, 1, N, torch.device("cuda"))
pairs_h = engine.pair_table(B)
pairs = torch.from_numpy(pairs_h).cuda()
want = np.array([delays[0, j] - delays[0, i] for i, j in pairs_h])
ref = None
for setting in settings:
    opts, flags = {}, 0
    for kv in filter(None, setting.split(",")):
        k, v = (t.strip() for t in kv.split("="))
        if k == "flags":
            flags = int(v, 0)
        else:
            opts[k] = int(v, 0)
    plan = engine.Plan(B, N, flags=flags, options=opts)
    for _ in range(2):
        S = plan.forward(iq[:, 0, :])
        rec = plan.xcorr_pairs_peak(S, pairs)
    torch.cuda.synchronize()
    got = engine.peaks_to_numpy(rec)
    if ref is None:
        ref = got.copy()
    plan.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        S = plan.forward(iq[:, 0, :])
        rec = plan.xcorr_pairs_peak(S, pairs)
        if idle_ms > 0:
            torch.cuda.synchronize()
            import time
            time.sleep(idle_ms * 1e-3)
    e1.record()
    torch.cuda.synchronize()
    prof = plan.profile_collect()
    plan.profile(False)
    per = {k: round(v[1] / iters, 4) for k, v in prof.items()}
    print(json.dumps({"setting": setting or "default", "idle_ms_between_iterations": idle_ms, "B": B, "N": N, "passes": plan.pass_lengths,
                      "lags_ok": bool(np.array_equal(got["lag"], want)),
                      "peak_maxrel_vs_first": float(np.max(np.abs(got["peak"] / ref["peak"] - 1))),
                      "ms_per_window_total": round(e0.elapsed_time(e1) / iters, 4), "ms_per_kernel": per}), flush=True)
    del plan, S, rec
