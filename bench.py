#!/usr/bin/env python3
"""Benchmark of the radio-mapper hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic cu8 IQ: for each of the
workload's windows, fused unpack + forward FFT of all buoys, then conj-multiply + inverse FFT
+ arg-max/parabolic lag for all buoy pairs.  Metric (BASELINE.json): correlated
pair-samples/s = pairs * samples_per_window * windows / time, whole job over all ranks.

One JSON line is printed by rank 0 (see the contract in the task statement).  For N > 1 it is
launched by torchrun, one rank per GPU; each rank processes its own windows (weak scaling)
and the 16-byte peak records are all-gathered over NCCL inside the timed step.

The same line carries `pair_sharded`: BASELINE config 4 as worded (64 buoys / 2016 pairs, ONE 2^20-sample
window whose pairs are sharded over the N ranks, strong scaling) with its own 1-GPU reference time.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "correlated_pair_samples_per_sec"
UNIT = "pair-samples/s"

WORKLOADS = {
    # BASELINE.json configs[2] — the default single-GPU correlate workload
    "cfg3": dict(desc="16 buoys (120 pairs), 2^22-sample windows, 5 windows (10.24 s @ 2.048 Msps) of synthetic cu8 IQ",
                 buoys=16, samples=1 << 22, windows=5, fs=2_048_000),
    # BASELINE.json configs[3]
    "cfg4": dict(desc="64 buoys (2016 pairs), 2^20-sample windows, 4 windows per rank", buoys=64, samples=1 << 20,
                 windows=4, fs=2_048_000),
    # BASELINE.json configs[0]
    "cfg1": dict(desc="3 buoys (3 pairs), 2.048 Msps, 1 s window", buoys=3, samples=2_048_000, windows=1, fs=2_048_000),
    # BASELINE.json configs[4]
    "cfg5": dict(desc="8 buoys (28 pairs), 2^26-sample windows", buoys=8, samples=1 << 26, windows=1, fs=2_048_000),
    # small smoke-sized case
    "tiny": dict(desc="4 buoys, 2^14-sample windows, 2 windows", buoys=4, samples=1 << 14, windows=2, fs=2_048_000),
}


def algorithmic_bytes(B, P, N, L):
    """per window  B*(2N + 8L) + P*(16L + 16)."""
    return B * (2 * N + 8 * L) + P * (16 * L + 16), B * (2 * N + 8 * L), P * (16 * L + 16)


def traffic_table():
    """Per-unit DRAM traffic of the pair-stage kernels from the newest committed ncu --set full capture."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                return json.load(f)
    raise FileNotFoundError("no profiles/r0*_traffic.json")


def measured_traffic(workload, B, P, L):
    """DRAM bytes per launch of the correlate+peak stage, scaled from the ncu --set full capture recorded in
    profiles/r01_traffic.json (per-unit ratios measured on the same kernels at the same FFT length)."""
    try:
        t = traffic_table()[workload]
        unit = 8 * L
        return int(unit * (B * t["pair_pass_read_per_buoy_unit8L"] + P * (t["pair_pass_write_per_pair_unit8L"] +
                                                                          t["argmax_pass_read_per_pair_unit8L"]))), t["source"]
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled every 50 ms from before the warm-up until after the timed region; the
    summary uses the samples whose timestamps fall inside the timed region (falling back to every
    sample taken under load if the region was shorter than the sampling period)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.idx = set(int(i) for i in gpu_indices)
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        # a timed region of a few ms (cfg1) can end before nvidia-smi has printed its first sample
        time.sleep(0.12 if (self.t1 or 0) - (self.t0 or 0) > 0.2 else 0.6)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not f[1].isdigit() or int(f[1]) not in self.idx:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), float(f[4]), [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= self.t1 + 0.05]
        scope = "timed region"
        if len(inside) < 2:
            loaded = [r for r in rows if r[3] > 250.0]                  # under load (warm-up runs the same kernels)
            inside, scope = (loaded or rows), "warm-up + timed region (timed region shorter than the sampling period)"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        reasons = sorted({n for r in inside for n in r[4]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)),
                "power_w_max": float(max(r[3] for r in inside)), "samples": len(inside), "scope": scope, "reasons": reasons}


# ----------------------------------------------------------------------------------------
# CPU legs (the only places that may execute oracle/)
# ----------------------------------------------------------------------------------------
def cpu_sample_run(iq_sample, threads):
    """Oracle (reference arithmetic) on uint8[b, 2N]: unpack + correlate + lag search for all
    pairs of the sample.  Returns (seconds, n_pairs)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    b = iq_sample.shape[0]
    pairs = oracle.pair_list(b)
    t0 = time.perf_counter()
    x = [oracle.unpack_cu8(row) for row in iq_sample]

    def one(pr):
        c, lags = oracle.xcorr_full(x[pr[0]], x[pr[1]])
        return oracle.peak_lag(c, lags)

    if threads <= 1:
        res = [one(p) for p in pairs]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            res = list(ex.map(one, pairs))
    return time.perf_counter() - t0, len(pairs), res


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    import platform
    return platform.processor() or platform.machine()


def cpu_baseline_cfg1():
    """the oracle on config 1 EXACTLY (3 buoys, 2.048 Msps cu8, 1 s window = 2 048 000 samples,
    3 pairs, seeded synthetic input), in this process on this box's host cores -- once as the reference would run
    it (scipy default workers=1) and once under scipy.fft.set_workers(os.cpu_count()); per-stage milliseconds,
    CPU model and library versions in the record."""
    import numpy
    import scipy
    import scipy.fft
    import oracle
    from radio_mapper_b200 import synth
    B, N = 3, 2_048_000
    iq, delays, _ = synth.delayed_buoys(1000, B, N, sample_rate=2_048_000)
    pairs = oracle.pair_list(B)

    def once():
        st = {"unpack_ms": 0.0, "correlate_ms": 0.0, "argmax_ms": 0.0}
        t0 = time.perf_counter()
        x = [oracle.unpack_cu8(row) for row in iq]
        t1 = time.perf_counter()
        st["unpack_ms"] = 1e3 * (t1 - t0)
        lags = []
        for i, j in pairs:
            ta = time.perf_counter()
            c, lg = oracle.xcorr_full(x[i], x[j])
            tb = time.perf_counter()
            lags.append(oracle.peak_lag(c, lg)[0])
            tc = time.perf_counter()
            st["correlate_ms"] += 1e3 * (tb - ta)
            st["argmax_ms"] += 1e3 * (tc - tb)
        st["total_ms"] = 1e3 * (time.perf_counter() - t0)
        return st, lags

    def best_of(k):
        runs = [once() for _ in range(k)]
        st, lags = min(runs, key=lambda r: r[0]["total_ms"])
        return {k2: round(v, 2) for k2, v in st.items()}, lags

    cores = os.cpu_count() or 1
    st1, lags1 = best_of(2)
    with scipy.fft.set_workers(cores):
        stn, lagsn = best_of(2)
    want = [int(delays[j] - delays[i]) for i, j in pairs]
    ps = len(pairs) * N
    return {"value": ps / (st1["total_ms"] * 1e-3), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "BASELINE config 1 exactly: 3 buoys x 2 048 000 samples, 3 pairs; numpy unpack + "
                      "scipy.signal.correlate('full','fft') + argmax/parabolic; best of 2 runs, scipy workers=1",
            "stages_ms": st1,
            "all_workers": {"value": ps / (stn["total_ms"] * 1e-3), "workers": cores, "stages_ms": stn,
                            "note": "the same run under scipy.fft.set_workers(os.cpu_count())"},
            "host_cpus": cores, "cpu_model": cpu_model(), "numpy": numpy.__version__, "scipy": scipy.__version__,
            "lags_match_known_delays": lags1 == want and lagsn == want}, iq, lags1


def run_reference(args, wl):
    """`--impl reference`: the reference's CPU arithmetic (oracle port: numpy unpack +
    scipy.signal.correlate + argmax; the reference has no compiled implementation of this
    path) on a bounded sample of the workload, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from radio_mapper_b200 import synth
    N = wl["samples"]
    cores = os.cpu_count() or 1
    # bounded sample: one window, as many buoys as keep a step to a few seconds
    b = min(wl["buoys"], 8 if N <= (1 << 22) else 3)
    iq, _, _ = synth.delayed_buoys(4242, b, N, sample_rate=wl["fs"])
    n_pairs = b * (b - 1) // 2
    threads = max(1, min(cores, n_pairs))
    for _ in range(args.warmup):
        cpu_sample_run(iq, threads)
    t_total = 0.0
    for _ in range(args.steps):
        t, _, _ = cpu_sample_run(iq, threads)
        t_total += t
    ps = n_pairs * N * args.steps / t_total
    sample = "one window, first %d buoys (%d pairs) of the workload per step" % (b, n_pairs)
    line = {
        "impl": "reference", "metric": METRIC, "value": ps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (complex64 numpy/scipy)", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "buoys": wl["buoys"], "samples_per_window": N,
                   "sample": sample},
        "cpu_baseline": {"value": ps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": cores},
        "e2e": {"value": ps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def welch_roofline(n_seg, nperseg, ms, clocks):
    """fp32 roofline of the Welch PSD (compute-bound): flops = W * 5 L log2 L (the radix-2 count the
    survey uses), peak = 148 SMs x 128 lanes x 2 flop x SM clock -- derived, at the clock sampled in this run and
    at the maximum clock; no measured fp32 peak exists in MEASURED_PEAKS.json."""
    import math
    flops = n_seg * 5.0 * nperseg * math.log2(nperseg)
    tf = flops / (ms * 1e-3) / 1e12
    mhz_max = float((clocks or {}).get("sm_max_mhz") or 1965.0)
    mhz_run = float((clocks or {}).get("sm_mhz") or mhz_max)
    peak_max = 148 * 128 * 2 * mhz_max * 1e6 / 1e12
    peak_run = 148 * 128 * 2 * mhz_run * 1e6 / 1e12
    return {"bound": "fp32", "achieved": tf, "unit": "TFLOP/s", "flops": flops, "peak": peak_max, "frac": tf / peak_max,
            "peak_at_sampled_clock": peak_run, "frac_at_sampled_clock": tf / peak_run,
            "peak_source": "derived: 148 SMs x 128 fp32 lanes x 2 x clock (max %.0f MHz, sampled %.0f MHz)" % (mhz_max, mhz_run),
            "note": "5 L log2 L counts a radix-2 FFT; the radix-16/32 kernels execute ~0.53x that many flops, so frac is "
                    "an upper bound of the ALU time actually spent"}


def pair_sharded_cfg4(device, world, rank, steps):
    """BASELINE config 4 as worded: 64 buoys (2016 pairs), 2^20-sample windows, ONE window whose pairs are sharded
    over the ranks as blocks of the pair matrix (sharding.tile_pairs: a rank transforms only its blocks' buoys),
    peak records assembled by one NCCL all-gather (strong scaling).  Timed like the headline: barrier +
    synchronize on both sides, CUDA events, max over ranks.  Rank 0 also times the same window on its GPU alone
    (the other ranks wait at a barrier), so the record carries its own 1-GPU reference."""
    import torch
    import torch.distributed as dist
    from radio_mapper_b200 import sharding, synth
    from radio_mapper_b200.correlator import Correlator
    B, N = 64, 1 << 20
    iq, delays = synth.delayed_buoys_torch(4000, B, 1, N, device)       # the same window on every rank (same seed)
    cor = Correlator(B, N, device=device)
    P = cor.n_pairs
    tiles = sharding.tiles_for(B, world)

    def step():
        rec, _ = cor.run_device_tile(iq, [0], tiles[rank])
        return sharding.gather_tiled_records(rec, tiles, P, world) if world > 1 else rec

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(3):
        rec = step()
    torch.cuda.synchronize()
    want = np.array([delays[0, j] - delays[0, i] for i, j in cor.pairs_host])
    got = np.empty(P, dtype=np.int64)
    if world > 1:
        got[:] = rec[0].cpu().numpy()[:, 0]
    else:
        got[tiles[0]["global_index"]] = rec[0].cpu().numpy()[:, 0]
    ok = bool(np.array_equal(got, want))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = timed(step, steps)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    one_ms = None
    if rank == 0:
        for _ in range(2):
            cor.run_device(iq, [0])
        torch.cuda.synchronize()
        one_ms = timed(lambda: cor.run_device(iq, [0]), steps)
    if world > 1:
        dist.barrier()
    out = None
    if rank == 0:
        out = {"workload": "cfg4 pair-sharded: 64 buoys (2016 pairs), 2^20-sample window, ONE window, pairs sharded over the ranks",
               "value": P * N / (ms * 1e-3), "unit": UNIT, "ms_per_window": ms, "n_gpus": world, "scaling": "strong",
               "steps": steps, "one_gpu_ms_per_window": one_ms, "speedup_vs_one_gpu": one_ms / ms,
               "buoys_transformed_on_rank0": int(len(tiles[0]["buoys"])), "pairs_on_rank0": int(len(tiles[0]["global_index"])),
               "passes": cor.plan.pass_lengths, "lags_match_known_delays": ok,
               "sharding": "blocks of the upper-triangular pair matrix (sharding.tile_pairs); each rank recomputes the "
                           "forward FFTs of its blocks' buoys from cu8; one NCCL all_gather_into_tensor of the 16-byte "
                           "peak records per window"}
    del cor, iq
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from radio_mapper_b200 import sharding, synth
    from radio_mapper_b200.tdoa_processor import TDoAProcessor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B, N, W, fs = wl["buoys"], wl["samples"], wl["windows"], wl["fs"]
    P = B * (B - 1) // 2

    # synthetic input, resident in HBM (device-timed arm) and in pinned host memory (e2e arm)
    iq_dev, delays = synth.delayed_buoys_torch(1000 * 3 + rank, B, W, N, device, sample_rate=fs)
    proc = TDoAProcessor()
    buoy_ids = ["BUOY_%02d" % b for b in range(B)]
    cor = proc._correlator(B, N, device)
    plan = cor.plan
    L = plan.fft_len
    windows = list(range(W))

    def step():
        rec, en = cor.run_device(iq_dev, windows)
        if world > 1:
            rec, en = sharding.gather_records(rec, en, W * world, P, world, rank)
        return rec, en

    sampler = ClockSampler(range(world) if rank == 0 else [])
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        rec, en = step()
    torch.cuda.synchronize()

    # correctness guard on the timed configuration: lags must equal the generator's delays
    got = rec[:W].cpu().numpy()[..., 0] if world == 1 else rec[rank * W:(rank + 1) * W].cpu().numpy()[..., 0]
    want = np.stack([delays[:, j] - delays[:, i] for i, j in cor.pairs_host], axis=1)
    if not np.array_equal(got, want):
        raise SystemExit("bench: lags do not match the synthetic delays (%d mismatches)" % int((got != want).sum()))

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_begin()
    plan.profile(True)
    cor.launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    if world > 1:
        dist.barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = cor.launches
    prof = plan.profile_collect()
    plan.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = P * N * W * world * args.steps / (elapsed_ms * 1e-3)

    # ---- one-pass windowed search (|lag| <= 2*342 samples: every lag two buoys within a 50 km
    #      radius of the source can produce, reference config.yaml:145) — reported beside the
    #      full-range headline, never instead of it ------------------------------------------
    win_lag = 2 * 342
    plan.set_max_lag(win_lag)
    for _ in range(3):
        rec_w, _ = step()
    torch.cuda.synchronize()
    got_w = rec_w[:W].cpu().numpy()[..., 0] if world == 1 else rec_w[rank * W:(rank + 1) * W].cpu().numpy()[..., 0]
    windowed_ok = bool(np.array_equal(got_w, want))
    if world > 1:
        dist.barrier()
    plan.profile(True)
    evw0, evw1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evw0.record()
    for _ in range(args.steps):
        step()
    evw1.record()
    torch.cuda.synchronize()
    win_ms = evw0.elapsed_time(evw1)
    prof_w = plan.profile_collect()
    plan.profile(False)
    plan.set_max_lag(None)
    if world > 1:
        t = torch.tensor([win_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        win_ms = float(t.item())

    # ---- end to end through the public API: pinned host cu8 -> TDoAMeasurement list ----------
    iq_host = torch.empty(iq_dev.shape, dtype=torch.uint8, pin_memory=True)
    iq_host.copy_(iq_dev)
    torch.cuda.synchronize()
    for _ in range(2):
        meas = proc.correlate_iq(iq_host, buoy_ids, fs, 121.5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        meas = proc.correlate_iq(iq_host, buoy_ids, fs, 121.5)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert len(meas) == W * P
    e2e_value = P * N * W * world * args.steps / e2e_s
    # the bus under the e2e number: pinned host -> device bandwidth of this box, and the copy time of one step
    # (only the first window's copy cannot hide behind kernels within a correlate_iq call)
    h2d_gbs = None
    try:
        scratch = torch.empty_like(iq_dev)
        hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scratch.copy_(iq_host, non_blocking=True)
        torch.cuda.synchronize()
        hb0.record()
        scratch.copy_(iq_host, non_blocking=True)
        hb1.record()
        torch.cuda.synchronize()
        h2d_gbs = iq_host.numel() / (hb0.elapsed_time(hb1) * 1e-3) / 1e9
        del scratch
    except Exception:
        h2d_gbs = None

    # ---- the same windows through the streaming API (ingest.ArraySource -> pinned ring -> copy stream ->
    #      correlator): the copy of window w+1 overlaps the kernels of window w across step boundaries too ----
    stream_value = None
    if rank == 0 and world == 1:
        from radio_mapper_b200 import ingest
        src_np = iq_host.view(B, -1)                                    # pinned uint8[B, W*2N]: windows are consecutive
        n_meas = sum(len(m) for m in proc.correlate_stream(ingest.ArraySource(src_np, N), buoy_ids, fs, 121.5, depth=3))
        torch.cuda.synchronize()

        class _Repeat:                                                  # args.steps passes over the same bytes
            def __init__(self, a, n, reps):
                self.inner = ingest.ArraySource(a, n)
                self.n_buoys, self.samples_per_window = self.inner.n_buoys, n
                self.n_windows = self.inner.n_windows * reps

            def read_window(self, w, out):
                return w < self.n_windows and self.inner.read_window(w % self.inner.n_windows, out)

            def pinned_window(self, w):
                return self.inner.pinned_window(w % self.inner.n_windows) if w < self.n_windows else None

            def close(self):
                pass

        t0 = time.perf_counter()
        n_meas = 0
        for m in proc.correlate_stream(_Repeat(src_np, N, args.steps), buoy_ids, fs, 121.5, depth=3):
            n_meas += len(m)
        torch.cuda.synchronize()
        stream_s = time.perf_counter() - t0
        assert n_meas == W * P * args.steps
        stream_value = {"value": P * N * W * args.steps / stream_s, "unit": UNIT, "ms_per_step": 1e3 * stream_s / args.steps,
                        "api": "TDoAProcessor.correlate_stream(ingest source, ...) -> TDoAMeasurements per window",
                        "note": "pinned source: windows are DMA'd in place; file / pipe sources add one host copy into the pinned ring"}

    # ---- BASELINE config 4 as worded: one window, pairs sharded over the ranks (strong scaling) -----------------
    pair_sharded = None
    if args.workload == "cfg3":
        try:
            pair_sharded = pair_sharded_cfg4(device, world, rank, max(10, args.steps))
        except Exception as exc:                    # secondary measurement: never fail the headline line
            pair_sharded = {"error": repr(exc)}
            if world > 1:
                raise

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage (fused correlate + peak) -------------------------------
    peak, peak_src = measured_peak()
    total_b, fwd_b, pair_b = algorithmic_bytes(B, P, N, L)
    pair_names = [k for k in prof if k.startswith(("contig_inv", "col_inv", "finalize"))]
    fwd_names = [k for k in prof if k not in pair_names]
    pair_ms = sum(prof[k][1] for k in pair_names)
    fwd_ms = sum(prof[k][1] for k in fwd_names)
    n_calls = args.steps * W                               # one correlate call per window
    pair_gbs = pair_b * n_calls / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
    whole_gbs = total_b * W * args.steps / (elapsed_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(args.workload, B, P, L)
    per_kernel = None
    try:
        tr = traffic_table()[args.workload]
        unit = 8 * L
        kb = {"contig_inv_pair": unit * (B * tr["pair_pass_read_per_buoy_unit8L"] + P * tr["pair_pass_write_per_pair_unit8L"]),
              "col_inv_argmax": unit * P * tr["argmax_pass_read_per_pair_unit8L"]}
        per_kernel = {}
        for name, nbytes in kb.items():
            if name in prof and prof[name][1] > 0:
                ms = prof[name][1] / n_calls
                per_kernel[name] = {"ms_per_launch": ms, "dram_bytes_per_launch": int(nbytes),
                                    "dram_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_peak": nbytes / (ms * 1e-3) / 1e9 / peak}
    except Exception:
        per_kernel = None
    roofline = {
        "bound": "hbm", "kernel": "correlate+peak stage (" + " + ".join(sorted(pair_names)) + ")",
        "achieved": pair_gbs, "peak": peak, "unit": "GB/s", "frac": pair_gbs / peak, "traffic": traffic,
        "traffic_source": traffic_src, "per_kernel_dram": per_kernel, "peak_source": peak_src, "algorithmic_bytes_per_launch": pair_b,
        "avg_launch_ms": pair_ms / max(1, n_calls), "share_of_step": pair_ms / max(1e-9, pair_ms + fwd_ms),
        "whole_step": {"achieved": whole_gbs, "frac": whole_gbs / peak, "algorithmic_bytes_per_window": total_b},
        "kernels_ms_per_step": {k: v[1] / args.steps for k, v in sorted(prof.items())},
    }

    win_pair_ms = sum(v[1] for k, v in prof_w.items() if k.startswith(("contig_inv", "finalize")))
    windowed = {
        "max_lag": win_lag, "lags_match_full_search": windowed_ok,
        "value": P * N * W * world * args.steps / (win_ms * 1e-3), "unit": UNIT, "ms_per_step": win_ms / args.steps,
        "pair_stage": {"achieved": pair_b * n_calls / (win_pair_ms * 1e-3) / 1e9 if win_pair_ms > 0 else 0.0, "unit": "GB/s",
                       "frac": (pair_b * n_calls / (win_pair_ms * 1e-3) / 1e9 / peak) if win_pair_ms > 0 else 0.0,
                       "note": "algorithmic bytes / time; spectra are re-read from L2, so this can exceed what HBM alone delivers"},
        "kernels_ms_per_step": {k: v[1] / args.steps for k, v in sorted(prof_w.items())},
    }

    # ---- BASELINE config 2 (secondary): Welch PSD, 64k bins, 1000 segments, device-resident cu8 ----
    welch = None
    try:
        from radio_mapper_b200 import engine as _eng
        del iq_dev
        nperseg, n_seg = 65536, 1000
        gen = torch.Generator(device=device)
        gen.manual_seed(2)
        wiq = torch.randint(96, 160, (2 * nperseg * n_seg,), dtype=torch.uint8, device=device, generator=gen)
        wplan = _eng.Plan(n_seg, nperseg, nperseg, device=device)
        for _ in range(3):
            wplan.welch_psd(wiq, 2.4e6)
        torch.cuda.synchronize()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(10):
            wplan.welch_psd(wiq, 2.4e6)
        w1.record()
        torch.cuda.synchronize()
        wms = w0.elapsed_time(w1) / 10
        welch = {"workload": "2.4 Msps, 64k-bin Welch PSD, 1000 segments", "value": n_seg * nperseg / (wms * 1e-3),
                 "unit": "samples/s", "ms": wms, "passes": wplan.pass_lengths,
                 "hbm_frac": (2.0 * n_seg * nperseg + 4 * nperseg) / (wms * 1e-3) / 1e9 / peak,
                 "roofline": welch_roofline(n_seg, nperseg, wms, clocks),
                 "kernel": "one thread-block-cluster kernel (8 CTAs hold a 64k segment in distributed shared memory)",
                 "note": "fp32-ALU / DSMEM bound (SURVEY §8d): 2 B/sample of traffic against ~80 flop/sample"}
        del wiq, wplan
    except Exception as exc:  # secondary measurement: never fail the headline line
        welch = {"error": repr(exc)}

    # ---- secondary: block detection (buoy_node.py:391-433) batched over raw cu8 blocks from pinned host memory ----
    detect = None
    try:
        from radio_mapper_b200 import synth as _synth
        from radio_mapper_b200.detectors import BuoySignalDetector
        nblk, nlen = 256, 32768
        ub, _ = _synth.welch_stream(9, nblk, nlen, fs)
        blocks = torch.from_numpy(np.ascontiguousarray(ub.reshape(nblk, 2 * nlen))).pin_memory()
        det = BuoySignalDetector("BUOY_B", 35.4676, -97.5164, fs)
        det.detect_blocks_arrays(blocks, 100.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            res = det.detect_blocks_arrays(blocks, 100.0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        import oracle as _oracle
        t0 = time.perf_counter()
        for bq in range(4):
            xo = _oracle.unpack_cu8(blocks[bq].numpy())
            po = _oracle.spectrum_db(_oracle.forward_fft(xo))
            oo = _oracle.score_peaks_buoy(po, _oracle.detect_peaks_fixed(po), _oracle.freq_axis_hz(nlen, fs, 100_000_000), 100_000_000)
        dt_cpu = (time.perf_counter() - t0) / 4
        detect = {"workload": "%d raw cu8 blocks x %d samples: unpack, FFT, dB, find_peaks(height=-70, distance=10), median, gates" % (nblk, nlen),
                  "value": nblk / dt, "unit": "blocks/s", "ms_per_batch": 1e3 * dt, "h2d_bytes_per_batch": int(blocks.numel()),
                  "cpu_oracle_blocks_per_s_1_thread": 1.0 / dt_cpu,
                  "block3_detections_match_oracle": sorted(int(k) for k in res[3][0]) == sorted(int(o["index"]) for o in oo)}
    except Exception as exc:
        detect = {"error": repr(exc)}

    # ---- CPU baseline: the oracle on config 1 exactly, this box's host cores; the same bytes go
    #      through the GPU path and the lags must agree ---------------------------------------------------------
    cpu_baseline, cfg1_iq, cfg1_cpu_lags = cpu_baseline_cfg1()
    try:
        from radio_mapper_b200 import engine as _eng1
        p1 = _eng1.Plan(3, 2_048_000, device=device)
        r1 = _eng1.peaks_to_numpy(p1.xcorr_pairs_peak(p1.forward(torch.from_numpy(cfg1_iq).to(device)),
                                                       torch.from_numpy(_eng1.pair_table(3)).to(device)))
        cpu_baseline["lags_match_gpu"] = [int(v) for v in r1["lag"]] == cfg1_cpu_lags
        del p1
    except Exception as exc:
        cpu_baseline["lags_match_gpu"] = repr(exc)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (complex64)", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "buoys": B, "pairs": P, "samples_per_window": N,
                   "fft_len": L, "windows_per_step_per_gpu": W, "passes": plan.pass_lengths,
                   "l2": "inputs larger than L2 (%.1f GB of spectra+workspace touched per window)" % (total_b / 1e9),
                   "sharding": "windows across ranks; NCCL all_gather of peak records" if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(iq_host.numel()),
                "d2h_bytes_per_step": int(W * P * 16 + W * B * 8), "ms_per_step": 1e3 * e2e_s / args.steps,
                "h2d_gbs_measured": h2d_gbs,
                "h2d_ms_per_step": (iq_host.numel() / (h2d_gbs * 1e9) * 1e3) if h2d_gbs else None,
                "h2d_ms_first_window": (iq_host.numel() / W / (h2d_gbs * 1e9) * 1e3) if h2d_gbs else None,
                "streaming": stream_value,
                "api": "TDoAProcessor.correlate_iq(pinned host uint8[B,W,2N]) -> List[TDoAMeasurement]"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "windowed_search": windowed,
        "welch_psd": welch, "block_detect": detect, "pair_sharded": pair_sharded,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def _keep_stdout_for_the_result():
    """stdout carries ONE JSON line.  Libraries write there too (NCCL prints its version line to fd 1 when the first
    communicator comes up), so everything but the result goes to stderr: fd 1 is parked and points at stderr until
    _emit() prints the line."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _RESULT_FD is not None:
        os.dup2(_RESULT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("RMX_WORKLOAD", "cfg3"), choices=sorted(WORKLOADS))
    args = ap.parse_args()
    _keep_stdout_for_the_result()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
