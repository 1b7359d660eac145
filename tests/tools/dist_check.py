#!/usr/bin/env python3
"""torchrun --nproc-per-node 2 tools/dist_check.py : the distributed=True API on real GPUs (NCCL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from radio_mapper_b200 import synth
from radio_mapper_b200.tdoa_processor import TDOAProcessor
import oracle

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n_windows in (5, 1):
    blocks = [synth.delayed_buoys(40 + w, 5, 1 << 15)[0] for w in range(n_windows)]
    iq = np.stack(blocks, axis=1)
    proc = TDOAProcessor()
    rec = proc.correlate_iq_records(torch.from_numpy(iq).pin_memory(), distributed=True)
    ok = all(np.array_equal(rec["lag"][w], oracle.xcorr_pairs_peak(blocks[w])["lag"]) for w in range(n_windows))
    # the distributed result must be BYTE-identical to the single-GPU one (windows dealt to ranks when there are
    # at least as many windows as ranks, blocks of the pair matrix otherwise): same kernels on the same spectra
    single = TDOAProcessor().correlate_iq_records(torch.from_numpy(iq).pin_memory(), distributed=False)
    same = all(np.array_equal(rec[f], single[f]) for f in ("lag", "peak", "frac", "coherence"))
    dev = TDOAProcessor().correlate_iq_records(torch.from_numpy(iq).cuda(), distributed=True)     # device-resident input
    same_dev = all(np.array_equal(dev[f], single[f]) for f in ("lag", "peak", "frac", "coherence"))
    print(f"rank {rank} windows={n_windows} shape={rec.shape} lags_ok={ok} identical_to_single_gpu={same} device_input_identical={same_dev} "
          f"coherence_min={rec['coherence'].min():.3f}", flush=True)
    assert ok and same and same_dev
dist.destroy_process_group()
