#!/usr/bin/env python3
"""Non-fail-fast numerical diagnostics of the CUDA path against the oracle (run on the GPU box)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.fft
import torch

import oracle
from radio_mapper_b200 import engine, synth


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def diag_forward():
    rng = np.random.default_rng(1)
    for logL in list(range(4, 19)) + [20, 22]:
        L = 1 << logL
        for N in sorted({L, L // 2, max(1, L // 2 - 3)}):
            B = 3 if logL <= 18 else 2
            u = rng.integers(0, 256, size=(B, 2 * N), dtype=np.uint8)
            try:
                plan = engine.Plan(B, N, L)
                S = plan.forward(torch.from_numpy(u).cuda())
                nat = plan.spectrum_natural(S).cpu().numpy()
                lay = S.cpu().numpy()
                x = np.zeros((B, L), np.complex64)
                for b in range(B):
                    x[b, :N] = oracle.unpack_cu8(u[b])
                ref = scipy.fft.fft(x.astype(np.complex128), axis=1)
                fi = plan.layout_freq_index()
                e1 = rel_l2(nat, ref)
                e2 = rel_l2(lay, ref[:, fi])
                print(f"fwd logL={logL:2d} N={N:8d} passes={plan.pass_lengths} rel_l2 nat={e1:.2e} layout={e2:.2e}"
                      + ("  <<< BAD" if max(e1, e2) > 5e-6 else ""))
            except Exception as ex:
                print(f"fwd logL={logL} N={N} EXC {type(ex).__name__}: {ex}")
                if "CUDA" in str(ex) or "illegal" in str(ex):
                    raise


def diag_xcorr():
    for N, B in [(8, 3), (100, 3), (1000, 4), (2048, 4), (4096, 4), (5000, 3), (1 << 14, 4), (1 << 16, 4), (1 << 18, 3), (1 << 20, 3)]:
        try:
            iq, d, _ = synth.delayed_buoys(100 + N, B, N, max_delay=min(342, max(1, N // 4)))
            t0 = time.time()
            ref = oracle.xcorr_pairs_peak(iq)
            t_or = time.time() - t0
            plan = engine.Plan(B, N)
            S = plan.forward(torch.from_numpy(iq).cuda())
            pairs = torch.from_numpy(engine.pair_table(B)).cuda()
            got = engine.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs))
            lag_ok = np.array_equal(got["lag"], ref["lag"])
            pk = np.max(np.abs(got["peak"] / ref["peak"] - 1))
            fr = np.max(np.abs(got["frac"] - ref["frac"]))
            print(f"xcorr N={N:8d} L={plan.fft_len} passes={plan.pass_lengths} lag_exact={lag_ok} peak_rel={pk:.2e} frac_abs={fr:.2e} oracle_s={t_or:.2f}"
                  + ("  <<< BAD" if (not lag_ok or pk > 1e-4 or fr > 1e-3) else ""))
            if not lag_ok:
                print("   got", got["lag"], "ref", ref["lag"], "true", [d[j] - d[i] for i, j in oracle.pair_list(B)])
        except Exception as ex:
            print(f"xcorr N={N} EXC {type(ex).__name__}: {ex}")
            traceback.print_exc()
            if "CUDA" in str(ex) or "illegal" in str(ex):
                raise


def diag_misc():
    rng = np.random.default_rng(5)
    for n in (1, 7, 8, 4099, 1 << 20):
        u = rng.integers(0, 256, size=2 * n, dtype=np.uint8)
        got = engine.unpack_cu8(torch.from_numpy(u).cuda()).cpu().numpy()
        print(f"unpack n={n} exact={np.array_equal(got.view(np.uint32), oracle.unpack_cu8(u).view(np.uint32))}")
    for n in (8192, 32768):
        u, bins = synth.tones_block(9, n)
        plan = engine.Plan(1, n, n)
        S = plan.forward(torch.from_numpy(u[None]).cuda())
        db = plan.spectrum_db(S)[0]
        ref = oracle.spectrum_db(oracle.forward_fft(oracle.unpack_cu8(u)))
        print(f"db n={n} max_abs={np.max(np.abs(db.cpu().numpy() - ref)):.2e}")
        refdb = torch.from_numpy(ref).cuda()
        cand = engine.threshold_peaks(refdb, -70.0)
        import scipy.signal
        want, _ = scipy.signal.find_peaks(ref, height=-70)
        print(f"   candidates equal={np.array_equal(cand, want)} ({len(cand)})")
        kept = engine.select_by_distance(cand, ref[cand], 10)
        want2, _ = scipy.signal.find_peaks(ref, height=-70, distance=10)
        print(f"   distance equal={np.array_equal(kept, want2)} ({len(kept)})")
        mean, med = engine.mean_median(refdb)
        print(f"   mean {mean} vs {np.mean(ref)}  median {med} vs {np.median(ref)}")
        mp, pk = engine.signal_stats(torch.from_numpy(u).cuda())
        st = oracle.signal_stats(oracle.unpack_cu8(u))
        print(f"   stats mean_power {mp} vs {np.mean(np.abs(oracle.unpack_cu8(u)).astype(np.float64)**2)} peak {pk} vs {st['peak_amplitude']}")
    for nperseg, W in [(4096, 8), (8192, 5), (65536, 12)]:
        u, bins = synth.welch_stream(3, W, nperseg)
        plan = engine.Plan(W, nperseg, nperseg)
        psd = plan.welch_psd(torch.from_numpy(u).cuda(), 2.4e6, segments_in_flight=5).cpu().numpy()
        f, ref = oracle.welch_psd(oracle.unpack_cu8(u), 2.4e6, nperseg)
        print(f"welch nperseg={nperseg} W={W} max_rel={np.max(np.abs(psd / ref - 1)):.2e}")


if __name__ == "__main__":
    torch.cuda.init()
    print(torch.cuda.get_device_name(0))
    for fn in (diag_misc, diag_forward, diag_xcorr):
        try:
            fn()
        except Exception:
            traceback.print_exc()
            break
    torch.cuda.synchronize()
    print("diag done")
