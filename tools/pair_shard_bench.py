#!/usr/bin/env python3
"""BASELINE config 4 as worded: 64 buoys (2016 pairs), 2^20-sample windows, ONE window at a time with the PAIRS
sharded over the ranks and the per-pair peak records all-gathered over NCCL (strong scaling).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        tools/pair_shard_bench.py [STEPS] [WARMUP]          (or plain `python tools/pair_shard_bench.py` for 1 GPU)

Every rank holds the window's cu8 (64 x 2 MB), recomputes the 64 forward FFTs (cheaper than receiving 8L-byte
spectra over NVLink) and correlates its contiguous slice of the i<j pair list.  Timed like bench.py:
barrier + synchronize on both sides, CUDA events, max over ranks.  Rank 0 prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from radio_mapper_b200 import sharding, synth
from radio_mapper_b200.correlator import Correlator

args = [a for a in sys.argv[1:] if not a.startswith("--")]
tiled = "--tiled" in sys.argv          # blocks of the pair matrix per rank instead of contiguous slices of the pair list
steps = int(args[0]) if len(args) > 0 else 10
warmup = int(args[1]) if len(args) > 1 else 3
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, N = 64, 1 << 20
iq, delays = synth.delayed_buoys_torch(4000, B, 1, N, dev)         # the same window on every rank (same seed)
cor = Correlator(B, N, device=dev)
P = cor.n_pairs
windows, pair_slice = sharding.shard_units(1, P, world, rank)
tiles = sharding.tile_pairs(B, world) if tiled else None


def step():
    if tiled:
        rec, en = cor.run_device_tile(iq, windows, tiles[rank])
        return sharding.gather_tiled_records(rec, tiles, P, world) if world > 1 else rec
    rec, en = cor.run_device(iq, windows, pair_slice)
    if world > 1:
        rec, en = sharding.gather_records(rec, en, 1, P, world, rank)
    return rec


for _ in range(max(3, warmup)):
    rec = step()
torch.cuda.synchronize()
got = rec[0].cpu().numpy()[:, 0]
want = np.array([delays[0, j] - delays[0, i] for i, j in cor.pairs_host])
assert np.array_equal(got, want), "gathered lags do not match the synthetic delays"
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    a, b = (0, P) if pair_slice is None else (pair_slice.start, pair_slice.stop)
    if tiled:
        a, b = 0, len(tiles[0]["global_index"])
    print(json.dumps({"metric": "correlated_pair_samples_per_sec", "value": P * N * steps / (ms * 1e-3), "unit": "pair-samples/s",
                      "n_gpus": world, "steps": steps, "ms_per_step": ms / steps, "scaling": "strong",
                      "config": {"workload": "cfg4 pair-sharded", "buoys": B, "pairs": P, "pairs_on_rank0": b - a,
                                 "samples_per_window": N, "passes": cor.plan.pass_lengths,
                                 "sharding": ("blocks of the upper-triangular pair matrix (sharding.tile_pairs): rank 0 transforms %d of the "
                                              "%d buoys; NCCL all_gather of the 16-byte peak records" % (len(tiles[0]["buoys"]), B)) if tiled else
                                             "contiguous slices of the i<j pair list; every rank recomputes the 64 forward FFTs; "
                                             "NCCL all_gather of the 16-byte peak records"},
                      "lags_match_known_delays": True}))
if world > 1:
    dist.destroy_process_group()
