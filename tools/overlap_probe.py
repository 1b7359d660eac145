#!/usr/bin/env python3
"""Probe: does running the row pass of one chunk of pairs concurrently with the column/arg-max pass of another
(two streams, two workspaces) beat the serial chain?  Developer aid; prints one JSON line per chunk count.

    python tools/overlap_probe.py BUOYS LOG2_SAMPLES ITERS 'K1,K2,...'"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from radio_mapper_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ks = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "1,2,4,8").split(",")]
iq, delays = synth.delayed_buoys_torch(7, B, 1, N, torch.device("cuda"))
pairs_h = engine.pair_table(B)
pairs = torch.from_numpy(pairs_h).cuda()
P = len(pairs_h)
want = np.array([delays[0, j] - delays[0, i] for i, j in pairs_h])
plans = [engine.Plan(B, N), engine.Plan(B, N)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
S = plans[0].forward(iq[:, 0, :])
out = torch.empty((P, 4), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()


def run(k):
    cuts = [round(c * P / k) for c in range(k + 1)]
    if k == 1:
        plans[0].xcorr_pairs_peak(S, pairs, out=out)
        return
    main = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(main)
    for c in range(k):
        with torch.cuda.stream(streams[c % 2]):
            plans[c % 2].xcorr_pairs_peak(S, pairs[cuts[c]:cuts[c + 1]], out=out[cuts[c]:cuts[c + 1]])
    for s in streams:
        main.wait_stream(s)


for k in ks:
    for _ in range(2):
        run(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run(k)
    e1.record()
    torch.cuda.synchronize()
    got = engine.peaks_to_numpy(out)
    print(json.dumps({"chunks": k, "B": B, "N": N, "lags_ok": bool(np.array_equal(got["lag"], want)),
                      "ms_xcorr": round(e0.elapsed_time(e1) / iters, 4)}), flush=True)
