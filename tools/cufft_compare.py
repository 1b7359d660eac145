#!/usr/bin/env python3
"""Library baseline for the same window: torch (cuFFT + ATen element-wise kernels) against librmx.

    python tools/cufft_compare.py [BUOYS] [LOG2_SAMPLES] [ITERS]

The library pipeline is what a straightforward GPU port of the oracle would be: unpack + zero-pad,
torch.fft.fft, X_j * conj(X_i) for all pairs, torch.fft.ifft, abs, argmax.  Both arms start from the same cu8
tensor resident in HBM and are timed with CUDA events after a warm-up.  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from radio_mapper_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda")
iq, delays = synth.delayed_buoys_torch(7, B, 1, N, dev)
iq = iq[:, 0, :].contiguous()
pairs_h = engine.pair_table(B)
P = len(pairs_h)
want = np.array([delays[0, j] - delays[0, i] for i, j in pairs_h])
L = 2 * N
ii = torch.from_numpy(pairs_h[:, 0].astype(np.int64)).to(dev)
jj = torch.from_numpy(pairs_h[:, 1].astype(np.int64)).to(dev)


def library_window(chunk):
    x = torch.view_as_complex((iq.view(B, N, 2).to(torch.float32) - 127.5).contiguous())
    S = torch.fft.fft(x, n=L, dim=1)
    lags = []
    for p0 in range(0, P, chunk):
        prod = S[jj[p0:p0 + chunk]] * S[ii[p0:p0 + chunk]].conj()
        c = torch.fft.ifft(prod, dim=1)
        k = torch.argmax(c.abs(), dim=1)
        lags.append(torch.where(k >= N, k - L, k))
    return torch.cat(lags)


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


chunk = max(1, min(P, int(6e9 // (L * 8 * 3))))        # keep the three [chunk, L] temporaries under ~6 GB each
lib_ms, lib_lags = timed(lambda: library_window(chunk))
plan = engine.Plan(B, N)
pairs = torch.from_numpy(pairs_h).to(dev)


def ours():
    return plan.xcorr_pairs_peak(plan.forward(iq), pairs)


our_ms, rec = timed(ours)
got = engine.peaks_to_numpy(rec)["lag"]
print(json.dumps({"buoys": B, "pairs": P, "samples": N, "fft_len": L, "passes": plan.pass_lengths,
                  "library_ms_per_window": round(lib_ms, 3), "librmx_ms_per_window": round(our_ms, 3),
                  "speedup": round(lib_ms / our_ms, 2), "library_pairs_per_chunk": chunk,
                  "library_lags_ok": bool(np.array_equal(lib_lags.cpu().numpy(), want)),
                  "librmx_lags_ok": bool(np.array_equal(got, want)),
                  "library": "torch %s (cuFFT c2c + ATen kernels)" % torch.__version__}))
