#!/usr/bin/env python3
"""Developer aid: what does this B200 sustain for pure writes, pure reads and copies?  (the
row pass of the correlate stage writes 8 bytes per element and reads almost nothing from HBM.)
Prints one JSON line per pattern; 4 GiB buffers (far above the 126 MB L2), best of 10 after warm-up, CUDA events."""
import json
import torch

n = 1 << 30                       # float32 elements = 4 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
a.fill_(1.0); b.fill_(2.0)
torch.cuda.synchronize()


def best(fn, reps=10):
    t = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


for name, fn, nbytes in [("fill (pure write, elementwise kernel)", lambda: a.fill_(3.0), 4 * n),
                         ("memset (pure write, cudaMemsetAsync)", lambda: a.zero_(), 4 * n),
                         ("sum (pure read)", lambda: torch.sum(a), 4 * n),
                         ("copy (read + write)", lambda: b.copy_(a), 8 * n),
                         ("add into third of the two (2 reads + 1 write)", lambda: torch.add(a, b, out=b), 12 * n)]:
    ms = best(fn)
    print(json.dumps({"pattern": name, "bytes": nbytes, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
