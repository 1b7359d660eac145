import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radio_mapper_b200 import engine, _native
nperseg, n_seg, fs = 65536, 1000, 2_400_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
iq = torch.randint(96, 160, (2 * nperseg * n_seg,), dtype=torch.uint8, device="cuda", generator=g)
for flags, name in ((0, "cluster"), (_native.PLAN_NO_WELCH_CLUSTER, "two_pass")):
    for inflight in (None,) if flags == 0 else (1000, 500, 250):
        plan = engine.Plan(n_seg, nperseg, nperseg, flags=flags)
        for _ in range(3):
            plan.welch_psd(iq, fs, segments_in_flight=inflight)
        torch.cuda.synchronize()
        plan.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            plan.welch_psd(iq, fs, segments_in_flight=inflight)
        e1.record(); torch.cuda.synchronize()
        prof = plan.profile_collect()
        print(json.dumps({"path": name, "inflight": inflight, "ms": e0.elapsed_time(e1) / 10, "passes": plan.pass_lengths,
                          "kernels_ms": {k: round(v[1] / 10, 4) for k, v in prof.items()}}), flush=True)
        del plan
