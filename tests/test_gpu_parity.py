"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, the golden
vectors produced by the reference, and size-independent properties at BASELINE sizes.

Tolerances (BASELINE.json north_star): cu8 unpack and integer lags bit-exact; correlation peak
values 1e-4 relative; sub-sample delays 1e-3 samples.  Spectra: rel-L2 <= 1e-5 (measured ~2e-7);
dB spectra 1e-3 dB away from deep nulls; identical detected-bin sets away from the threshold.
"""
import json
import os

import numpy as np
import pytest
import scipy.fft
import scipy.signal

import oracle
from radio_mapper_b200 import synth

pytestmark = pytest.mark.gpu

PEAK_RTOL = 1e-4
FRAC_ATOL = 1e-3


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check_records(got, ref):
    assert np.array_equal(got["lag"], ref["lag"]), (got["lag"], ref["lag"])
    assert np.max(np.abs(got["peak"] / ref["peak"] - 1)) <= PEAK_RTOL
    assert np.max(np.abs(got["frac"] - ref["frac"])) <= FRAC_ATOL


# ---- stage 1 -----------------------------------------------------------------------------
def test_unpack_bit_exact_against_reference_golden(rmx, golden_dir):
    g = np.load(os.path.join(golden_dir, "unpack.npz"))
    got = rmx.unpack_cu8(_cuda(g["raw"])).cpu().numpy()
    assert got.dtype == np.complex64
    assert np.array_equal(got.view(np.uint32), g["x_file"].view(np.uint32))


@pytest.mark.parametrize("n", [1, 7, 8, 9, 4099, 1 << 16, (1 << 20) + 5])
def test_unpack_bit_exact_random(rmx, n):
    rng = np.random.default_rng(n)
    raw = rng.integers(0, 256, size=2 * n, dtype=np.uint8)
    got = rmx.unpack_cu8(_cuda(raw)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), oracle.unpack_cu8(raw).view(np.uint32))
    # unaligned view (scalar path)
    if n > 8:
        import torch
        t = _cuda(np.concatenate([[0, 0], raw]).astype(np.uint8))[2:]
        got2 = rmx.unpack_cu8(t.clone() if False else t.contiguous()).cpu().numpy()
        assert np.array_equal(got2.view(np.uint32), oracle.unpack_cu8(raw).view(np.uint32))


def test_unpack_empty(rmx):
    import torch
    out = rmx.unpack_cu8(torch.empty(0, dtype=torch.uint8, device="cuda"))
    assert out.numel() == 0


# ---- stage 2 -----------------------------------------------------------------------------
@pytest.mark.parametrize("logL", [4, 5, 7, 9, 12, 13, 14, 16, 17, 19, 21, 22, 24])
def test_forward_fft_vs_scipy(rmx, logL):
    L = 1 << logL
    rng = np.random.default_rng(logL)
    B = 3 if logL <= 20 else 2
    for N in sorted({L, L // 2, max(1, L // 2 - 3)}):
        raw = rng.integers(0, 256, size=(B, 2 * N), dtype=np.uint8)
        plan = rmx.Plan(B, N, L)
        S = plan.forward(_cuda(raw))
        nat = plan.spectrum_natural(S).cpu().numpy()
        x = np.zeros((B, L), np.complex64)
        for b in range(B):
            x[b, :N] = oracle.unpack_cu8(raw[b])
        ref = scipy.fft.fft(x.astype(np.complex128), axis=1)
        err = np.linalg.norm(nat - ref) / np.linalg.norm(ref)
        assert err < 1e-6, (logL, N, plan.pass_lengths, err)
        # the documented layout: position p holds bin layout_freq_index()[p]
        lay = S.cpu().numpy()
        assert np.array_equal(lay, nat[:, plan.layout_freq_index()])
        # complex64 entry point gives the same spectrum
        S2 = plan.forward_c64(_cuda(x[:, :N].copy()))
        assert np.linalg.norm(S2.cpu().numpy() - lay) / np.linalg.norm(lay) < 1e-6
        # and matches the reference's own complex64 scipy.fft.fft within fp32 rounding
        ref32 = oracle.forward_fft(x[0])
        assert np.linalg.norm(nat[0] - ref32) / np.linalg.norm(ref32) < 2e-6


def test_strided_input_rows(rmx):
    """One window cut out of a [buoy, stream] buffer (row stride > 2N)."""
    rng = np.random.default_rng(5)
    B, W, N = 3, 4, 4096
    raw = rng.integers(0, 256, size=(B, W, 2 * N), dtype=np.uint8)
    dev = _cuda(raw)
    plan = rmx.Plan(B, N)
    a = plan.forward(dev[:, 2, :]).cpu().numpy()
    b = plan.forward(_cuda(raw[:, 2, :].copy())).cpu().numpy()
    assert np.array_equal(a, b)


# ---- stages 3+4 ---------------------------------------------------------------------------
@pytest.mark.parametrize("n_samples,n_buoys", [(8, 3), (100, 3), (1000, 4), (2048, 4), (4096, 4), (5000, 3),
                                               (1 << 14, 5), (1 << 16, 4), (100000, 3), (1 << 18, 3), (1 << 20, 3)])
def test_xcorr_peak_parity(rmx, n_samples, n_buoys):
    iq, delays, _ = synth.delayed_buoys(100 + n_samples, n_buoys, n_samples, max_delay=min(342, max(1, n_samples // 4)))
    ref = oracle.xcorr_pairs_peak(iq)
    plan = rmx.Plan(n_buoys, n_samples)
    S = plan.forward(_cuda(iq))
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(rmx.pair_table(n_buoys))))
    _check_records(got, ref)
    if n_samples >= 1000:          # and both equal the ground truth the generator used
        assert list(got["lag"]) == [delays[j] - delays[i] for i, j in oracle.pair_list(n_buoys)]


def test_xcorr_chunked_workspace_and_pair_subsets(rmx):
    iq, delays, _ = synth.delayed_buoys(77, 6, 1 << 15)
    ref = oracle.xcorr_pairs_peak(iq)
    plan = rmx.Plan(6, 1 << 15)
    S = plan.forward(_cuda(iq))
    pairs = rmx.pair_table(6)
    a = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(pairs), max_pairs_in_flight=4))
    _check_records(a, ref)
    sel = np.array([14, 3, 3, 0], dtype=np.int64)            # arbitrary order, duplicates allowed
    b = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(pairs[sel])))
    _check_records(b, ref[sel])
    # reversed pair (j, i) negates the lag
    c = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(pairs[:, ::-1].copy())))
    assert np.array_equal(c["lag"], -ref["lag"])
    # empty pair list
    import torch
    assert plan.xcorr_pairs_peak(S, torch.empty((0, 2), dtype=torch.int32, device="cuda")).shape[0] == 0


def test_correlate_chain_is_deterministic(rmx):
    """No atomics on the data path of the correlate stage: the same window must give bit-identical spectra and peak
    records every time; a difference would be a race (mbarrier phases, bulk-copy ordering, exchange-buffer reuse).
    tools/soak.py runs the same check at the BASELINE sizes (profiles/r02_soak_determinism.jsonl)."""
    import torch
    for n, b in ((1 << 17, 12), (1 << 19, 5)):
        iq, _, _ = synth.delayed_buoys(31 + b, b, n, max_delay=200)
        plan = rmx.Plan(b, n)
        dev, pairs = _cuda(iq), _cuda(rmx.pair_table(b))
        first_s = plan.forward(dev).clone()
        first = plan.xcorr_pairs_peak(first_s, pairs).clone()
        for _ in range(150):
            s = plan.forward(dev)
            assert torch.equal(torch.view_as_real(s), torch.view_as_real(first_s))
            assert torch.equal(plan.xcorr_pairs_peak(s, pairs), first)


@pytest.mark.parametrize("options", [{}, {"pair_prefetch": 0}, {"pair_run": 16}, {"pair_groups": 2}, {"pair_store": 2}])
def test_xcorr_arbitrary_pair_lists(rmx, options):
    """The X_i-stationary row pass walks runs of consecutive pairs and re-reads X_i when i changes: any pair list must
    do -- shuffled, reversed (j, i), repeated and auto-correlation (i, i) entries, lengths that are no multiple of the
    run -- with records bit-identical to the same pairs taken from the ordered table (two-pass plan 64 x 4096)."""
    n, b = 1 << 17, 7
    iq, delays, _ = synth.delayed_buoys(4242, b, n, max_delay=300)
    plan = rmx.Plan(b, n, options=options)
    S = plan.forward(_cuda(iq))
    table = rmx.pair_table(b)
    ordered = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(table))).copy()
    assert list(ordered["lag"]) == [delays[j] - delays[i] for i, j in oracle.pair_list(b)]
    rng = np.random.default_rng(5)
    sel = rng.permutation(len(table))[:17]                       # 17 pairs: two full runs of 8 and one of 1
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(table[sel].copy())))
    assert got.tobytes() == ordered[sel].tobytes()
    mixed = np.concatenate([table[sel[:5]], table[sel[:5], ::-1], np.array([[3, 3], [0, 0], [6, 6]], dtype=table.dtype),
                            table[sel[:3]]]).astype(np.int32)
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(mixed)))
    assert got[:5].tobytes() == ordered[sel[:5]].tobytes() and got[13:].tobytes() == ordered[sel[:3]].tobytes()
    assert np.array_equal(got["lag"][5:10], -ordered["lag"][sel[:5]])           # (j, i): mirrored correlation
    assert np.allclose(got["peak"][5:10], ordered["peak"][sel[:5]], rtol=1e-6)
    assert np.array_equal(got["lag"][10:13], [0, 0, 0]) and np.all(np.abs(got["frac"][10:13]) < 1e-3)   # auto-correlation


@pytest.mark.parametrize("max_lag", [0, 5, 342, 5000])
def test_xcorr_max_lag_window(rmx, max_lag):
    iq, delays, _ = synth.delayed_buoys(31, 4, 1 << 14, max_delay=300)
    ref = oracle.xcorr_pairs_peak(iq, max_lag=max_lag)
    plan = rmx.Plan(4, 1 << 14)
    plan.set_max_lag(max_lag)
    S = plan.forward(_cuda(iq))
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(rmx.pair_table(4))))
    assert np.array_equal(got["lag"], ref["lag"])
    assert np.max(np.abs(got["peak"] / ref["peak"] - 1)) <= PEAK_RTOL
    assert np.max(np.abs(got["frac"] - ref["frac"])) <= FRAC_ATOL
    plan.set_max_lag(None)
    full = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(rmx.pair_table(4))))
    assert np.array_equal(full["lag"], oracle.xcorr_pairs_peak(iq)["lag"])


@pytest.mark.parametrize("log_n,max_lag", [(15, 342), (16, 0), (16, 5), (16, 342), (16, 511), (16, 512), (16, 2047),
                                           (16, 2048), (18, 342), (20, 342), (22, 342)])
def test_xcorr_one_pass_windowed_search(rmx, log_n, max_lag):
    """max_lag << row length takes the one-pass windowed kernel; it must agree with the oracle and
    with the full inverse transform + masked arg-max (set_search_mode(force_full=True))."""
    n = 1 << log_n
    iq, delays, _ = synth.delayed_buoys(900 + log_n, 4, n, max_delay=300)
    plan = rmx.Plan(4, n)
    S = plan.forward(_cuda(iq))
    pairs = _cuda(rmx.pair_table(4))
    plan.set_max_lag(max_lag)
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs))
    plan.set_search_mode(True)
    full = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs))
    plan.set_search_mode(False)
    assert np.array_equal(got["lag"], full["lag"])
    assert np.max(np.abs(got["peak"] / full["peak"] - 1)) <= PEAK_RTOL
    assert np.max(np.abs(got["frac"] - full["frac"])) <= FRAC_ATOL
    if log_n <= 20:
        ref = oracle.xcorr_pairs_peak(iq, max_lag=max_lag)
        assert np.array_equal(got["lag"], ref["lag"])
        assert np.max(np.abs(got["peak"] / ref["peak"] - 1)) <= PEAK_RTOL
        assert np.max(np.abs(got["frac"] - ref["frac"])) <= FRAC_ATOL
    if max_lag >= 600:
        assert list(got["lag"]) == [delays[j] - delays[i] for i, j in oracle.pair_list(4)]
    # the windowed path (taken while the window is narrower than half a row of the innermost pass,
    # i.e. max_lag < 1024 for 4096-point rows) needs (much) less workspace than the full one
    if 0 < max_lag < 1024 and log_n >= 16:
        small = plan.workspace_bytes(6)
        plan.set_search_mode(True)
        assert small < plan.workspace_bytes(6)
        plan.set_search_mode(False)


@pytest.mark.parametrize("log_n", [17, 20, 22])
def test_kernel_variants_agree(rmx, log_n):
    """The TMA-fed arg-max pass, the X_i-stationary row pass and the twiddle placement are
    re-arrangements of the same arithmetic: every variant must return the oracle's lags, and peaks /
    sub-sample offsets inside the north_star tolerances of each other (plans 32x4096, 256x4096, 1024x8192)."""
    n = 1 << log_n
    iq, delays, _ = synth.delayed_buoys(1200 + log_n, 5, n, max_delay=300)
    want = np.array([delays[j] - delays[i] for i, j in oracle.pair_list(5)])
    pairs = _cuda(rmx.pair_table(5))
    results = {}
    from radio_mapper_b200 import _native as nat
    for name, flags, options in [("default", 0, {}), ("no_tma", nat.PLAN_NO_TMA, {}), ("no_pair_run", nat.PLAN_NO_PAIR_RUN, {}),
                                 ("twiddle_in_col", nat.PLAN_TWIDDLE_IN_COL, {}),
                                 ("no_prefetch", 0, {"pair_prefetch": 0}), ("run16", 0, {"pair_run": 16}),
                                 ("staged_store", 0, {"pair_store": 1}), ("xstaged_store", 0, {"pair_store": 2}), ("xi_smem", 0, {"pair_xi_smem": 1}), ("xi_late", 0, {"pair_xi_early": 0}),
                                 ("groups2", 0, {"pair_groups": 2}), ("groups3", 0, {"pair_groups": 3}),
                                 ("row_e8", nat.PLAN_ROW_E8, {}), ("row_e8_ctas4", nat.PLAN_ROW_E8, {"pair_ctas": 4, "pair_prefetch": 0}),
                                 ("fwd_groups", 0, {"fwd_group_bytes": 3 * 8 * 2 * n}), ("fwd_ldg", 0, {"fwd_tma": 0}),
                                 ("all_off", nat.PLAN_NO_TMA | nat.PLAN_NO_PAIR_RUN | nat.PLAN_TWIDDLE_IN_COL,
                                  {"pair_prefetch": 0, "fwd_tma": 0})]:
        plan = rmx.Plan(5, n, flags=flags, options=options)
        S = plan.forward(_cuda(iq))
        results[name] = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs)).copy()
        full = plan.xcorr_full(S, pairs[:2].contiguous()).cpu().numpy()
        results[name + "_full"] = full
    ref = results["default"]
    assert np.array_equal(ref["lag"], want)
    # the TMA-fed forward pass (default) and the per-thread 128-bit staging stage the same bytes differently:
    # identical spectra, hence identical records
    assert np.array_equal(results["fwd_ldg"], ref) and np.array_equal(results["fwd_ldg_full"], results["default_full"])
    for name in ("no_tma", "no_pair_run", "twiddle_in_col", "no_prefetch", "run16", "staged_store", "xstaged_store", "xi_smem", "xi_late", "groups2", "groups3", "row_e8", "row_e8_ctas4", "fwd_groups", "fwd_ldg", "all_off"):
        _check_records(results[name], ref)
        a, b = results[name + "_full"], results["default_full"]
        assert np.linalg.norm(a - b) <= 2e-6 * np.linalg.norm(b), name
    # same arithmetic, different data movement: bit-identical records
    assert results["no_tma"].tobytes() == ref.tobytes()
    assert results["no_pair_run"].tobytes() == ref.tobytes()


def test_xcorr_edge_inputs(rmx):
    """Saturated / constant inputs and peaks on the edge of the lag range."""
    n = 4096
    rng = np.random.default_rng(9)
    base = rng.integers(0, 256, size=2 * n, dtype=np.uint8)
    iq = np.stack([base,
                   np.roll(base, 2 * (n - 1)),                  # circular shift: peak near the range edge
                   np.full(2 * n, 255, np.uint8),               # saturated constant
                   np.where(rng.random(2 * n) < 0.5, 0, 255).astype(np.uint8)])   # full-scale square noise
    ref = oracle.xcorr_pairs_peak(iq)
    plan = rmx.Plan(4, n)
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(plan.forward(_cuda(iq)), _cuda(rmx.pair_table(4))))
    truth = oracle.xcorr_pairs_peak(iq, dtype=np.complex128)
    for k in range(len(ref)):
        # where float32 and float64 oracles disagree the arg-max is numerically ambiguous
        if ref["lag"][k] == truth["lag"][k]:
            assert got["lag"][k] == ref["lag"][k], k
            assert abs(got["peak"][k] / ref["peak"][k] - 1) <= PEAK_RTOL
    # identical signals: peak exactly at lag 0 with the signal energy as the peak value
    same = np.stack([base, base])
    plan2 = rmx.Plan(2, n)
    r = rmx.peaks_to_numpy(plan2.xcorr_pairs_peak(plan2.forward(_cuda(same)), _cuda(rmx.pair_table(2))))
    energy = float(np.sum(np.abs(oracle.unpack_cu8(base).astype(np.complex128)) ** 2))
    assert r["lag"][0] == 0 and abs(r["peak"][0] / energy - 1) < 1e-5 and abs(r["frac"][0]) < 1e-3


def test_xcorr_full_output_matches_scipy(rmx):
    iq, _, _ = synth.delayed_buoys(5, 3, 3000)
    plan = rmx.Plan(3, 3000)
    S = plan.forward(_cuda(iq))
    c = plan.xcorr_full(S, _cuda(rmx.pair_table(3))).cpu().numpy()
    L = plan.fft_len
    for p, (i, j) in enumerate(oracle.pair_list(3)):
        ref, lags = oracle.xcorr_full(oracle.unpack_cu8(iq[i]), oracle.unpack_cu8(iq[j]), dtype=np.complex128)
        got = c[p][lags % L]
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-6


def test_xcorr_full_into_unaligned_output(rmx):
    """The 8192-point row pass writes each row with one 16-byte-aligned bulk copy; an output tensor that is
    only 8-byte aligned must take the per-thread store path and give the same bits."""
    import torch
    n = 1 << 22                                       # plan 1024 x 8192
    iq, _, _ = synth.delayed_buoys(77, 2, n)
    plan = rmx.Plan(2, n)
    S = plan.forward(_cuda(iq))
    pairs = _cuda(rmx.pair_table(2))
    aligned = plan.xcorr_full(S, pairs)
    backing = torch.empty(plan.fft_len + 1, dtype=torch.complex64, device="cuda")
    view = backing[1:].view(1, plan.fft_len)
    assert view.data_ptr() % 16 == 8
    out = plan.xcorr_full(S, pairs, out=view)
    assert torch.equal(out, aligned)


def test_fractional_delay_accuracy(rmx):
    """Sub-sample delays: GPU frac == oracle frac within 1e-3, and both track the true delay."""
    fr = [0.0, 0.3, -0.2, 0.12]
    iq, d, f = synth.delayed_buoys(11, 4, 1 << 16, snr_db=20, frac_delays=fr)
    ref = oracle.xcorr_pairs_peak(iq)
    plan = rmx.Plan(4, 1 << 16)
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(plan.forward(_cuda(iq)), _cuda(rmx.pair_table(4))))
    _check_records(got, ref)
    true = np.array([(d[j] + f[j]) - (d[i] + f[i]) for i, j in oracle.pair_list(4)])
    assert np.max(np.abs(got["lag"] + got["frac"] - true)) < 0.1


# ---- BASELINE-sized properties ----------------------------------------------------------------
def _delays_ok(rmx, n_buoys, log_samples, seed, n_windows=1):
    import torch
    iq, delays = synth.delayed_buoys_torch(seed, n_buoys, n_windows, 1 << log_samples, torch.device("cuda"))
    plan = rmx.Plan(n_buoys, 1 << log_samples)
    pairs_h = rmx.pair_table(n_buoys)
    pairs = _cuda(pairs_h)
    for w in range(n_windows):
        S = plan.forward(iq[:, w, :])
        got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs))
        want = delays[w, pairs_h[:, 1]] - delays[w, pairs_h[:, 0]]
        assert np.array_equal(got["lag"], want)
        assert np.all(np.abs(got["frac"]) < 0.2) and np.all(got["peak"] > 0)
    return plan


def test_cfg3_size_known_delays(rmx):
    """16 buoys / 120 pairs / N = 2^22 (L = 2^23): every lag equals the generator's delay."""
    plan = _delays_ok(rmx, 16, 22, 303)
    assert plan.fft_len == 1 << 23


def test_cfg4_size_known_delays(rmx):
    """64 buoys / 2016 pairs / N = 2^20."""
    _delays_ok(rmx, 64, 20, 404)


def test_cfg5_size_known_delays(rmx):
    """N = 2^26 (L = 2^27, three passes): 3 buoys of the 8 to bound memory and time."""
    plan = _delays_ok(rmx, 3, 26, 505)
    assert len(plan.pass_lengths) == 3


def _oracle_subset_parity(rmx, n_buoys, log_samples, seed, subset, expect_passes):
    """All pairs of all `n_buoys` buoys on the GPU at the BASELINE plan size; the pairs among the buoys in `subset`
    are also run through the CPU oracle on the same bytes: lags bit-exact, peak within 1e-4 relative, sub-sample
    offset within 1e-3 samples (north_star tolerances) -- plus every GPU lag against the generator's delay."""
    import torch
    n = 1 << log_samples
    iq, delays = synth.delayed_buoys_torch(seed, n_buoys, 1, n, torch.device("cuda"))
    plan = rmx.Plan(n_buoys, n)
    assert plan.pass_lengths == expect_passes
    pairs_h = rmx.pair_table(n_buoys)
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(plan.forward(iq[:, 0, :]), _cuda(pairs_h)))
    assert np.array_equal(got["lag"], delays[0, pairs_h[:, 1]] - delays[0, pairs_h[:, 0]])
    host = iq[list(subset), 0, :].cpu().numpy()
    ref = oracle.xcorr_pairs_peak(host)                       # pairs of the subset, i<j in subset order
    index = {(int(i), int(j)): k for k, (i, j) in enumerate(pairs_h)}
    sel = [index[(subset[a], subset[b])] for a, b in oracle.pair_list(len(subset))]
    _check_records(got[sel], ref)
    return got, ref


def test_cfg3_plan_matches_oracle(rmx):
    """BASELINE config 3 as benchmarked: 16 buoys / 120 pairs / N = 2^22, plan 1024 x 8192.  Six pairs (four
    buoys) against scipy.signal.correlate on the CPU."""
    _oracle_subset_parity(rmx, 16, 22, 3303, (0, 5, 10, 15), [1024, 8192])


def test_cfg4_plan_matches_oracle(rmx):
    """BASELINE config 4: 64 buoys / 2016 pairs / N = 2^20, plan 512 x 4096 (X_i-stationary row pass with the
    bulk-copy prefetch, TMA arg-max pass).  Six pairs drawn across the 64 buoys against the oracle."""
    _oracle_subset_parity(rmx, 64, 20, 4404, (0, 17, 42, 63), [512, 4096])


def test_cfg5_plan_matches_oracle(rmx):
    """BASELINE config 5: ALL 8 buoys / 28 pairs at N = 2^26 (three-pass plan 128 x 256 x 4096): every lag against
    the generator, and one pair against the oracle (a 2^27-point scipy correlation takes ~1 min of CPU)."""
    got, ref = _oracle_subset_parity(rmx, 8, 26, 5505, (2, 5), [128, 256, 4096])
    assert len(got) == 28 and len(ref) == 1


@pytest.mark.parametrize("n_buoys,world,log_samples", [(16, 3, 20), (64, 8, 16), (5, 8, 17)])
def test_tiled_records_identical_to_untiled(rmx, n_buoys, world, log_samples):
    """sharding.tile_pairs deals blocks of the pair matrix to ranks; a rank transforms only its tile's buoys and
    correlates only its pairs (Correlator.run_device_tile).  Running EVERY tile on one GPU and assembling the
    records by global pair index must reproduce Correlator.run_device byte for byte: same kernels, same
    spectra, only the launch grouping differs."""
    import torch
    from radio_mapper_b200 import sharding
    from radio_mapper_b200.correlator import Correlator
    n = 1 << log_samples
    iq, delays = synth.delayed_buoys_torch(90 + n_buoys, n_buoys, 2, n, torch.device("cuda"))
    cor = Correlator(n_buoys, n)
    full, energy = cor.run_device(iq, [0, 1])
    full = full.clone()
    tiles = sharding.tile_pairs(n_buoys, world)
    assert sorted(np.concatenate([t["global_index"] for t in tiles]).tolist()) == list(range(cor.n_pairs))
    assembled = torch.full_like(full, -1)
    for t in tiles:
        rec, en = cor.run_device_tile(iq, [0, 1], t)
        assert torch.equal(en, energy)
        if len(t["global_index"]):
            assembled[:, torch.from_numpy(t["global_index"]).cuda()] = rec
    assert torch.equal(assembled, full)
    want = np.stack([delays[:, j] - delays[:, i] for i, j in cor.pairs_host], axis=1)
    assert np.array_equal(full.cpu().numpy()[..., 0], want)


def test_cfg1_exact_reference_case(rmx):
    """3 buoys, 2 048 000 samples (not a power of two), 3 pairs — compared with the oracle."""
    iq, delays, _ = synth.delayed_buoys(1000, 3, 2_048_000)
    ref = oracle.xcorr_pairs_peak(iq)
    plan = rmx.Plan(3, 2_048_000)
    assert plan.fft_len == 1 << 22
    got = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(plan.forward(_cuda(iq)), _cuda(rmx.pair_table(3))))
    _check_records(got, ref)
    assert list(got["lag"]) == [delays[j] - delays[i] for i, j in oracle.pair_list(3)]


def test_linearity_and_shift_properties(rmx):
    """Size-independent properties at 2^20: the correlation of a signal with its own circular
    shift peaks at the shift; swapping the pair negates the lag; scaling the input scales |c|."""
    import torch
    n = 1 << 20
    rng = np.random.default_rng(1)
    a = rng.integers(96, 160, size=2 * n, dtype=np.uint8)
    shift = 12345
    b = np.roll(a, 2 * shift)
    iq = np.stack([a, b])
    plan = rmx.Plan(2, n)
    S = plan.forward(_cuda(iq))
    r = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, _cuda(np.array([[0, 1], [1, 0]], dtype=np.int32))))
    assert r["lag"][0] == shift and r["lag"][1] == -shift
    assert abs(r["peak"][0] / r["peak"][1] - 1) < 1e-5
    # Parseval on the spectra: sum|X|^2 == L * sum|x|^2
    x = oracle.unpack_cu8(a).astype(np.complex128)
    e_spec = float((S[0].abs().double() ** 2).sum().item())
    assert abs(e_spec / (plan.fft_len * np.sum(np.abs(x) ** 2)) - 1) < 1e-5


def test_subsample_delay_accuracy_vs_snr(rmx):
    """BASELINE config 5's sweep at a test-sized window: known integer + fractional delays, per-buoy SNR from
    +20 to -10 dB.  The measured delay (lag + parabolic offset) tracks the truth, and the error grows as the
    SNR falls (`tests/tools/snr_sweep.py` runs the same sweep at 2^26 samples)."""
    import torch
    n, B = 1 << 20, 4
    frac = np.array([0.0, 0.31, -0.22, 0.4])
    plan = rmx.Plan(B, n)
    pairs_h = rmx.pair_table(B)
    pairs = _cuda(pairs_h)
    rms = {}
    for snr in (20, 0, -10):
        iq, d = synth.delayed_buoys_torch(300 + snr, B, 1, n, torch.device("cuda"), snr_db=float(snr), frac_delays=frac)
        rec = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(plan.forward(iq[:, 0, :]), pairs))
        true = np.array([(d[0, j] + frac[j]) - (d[0, i] + frac[i]) for i, j in pairs_h])
        err = (rec["lag"] + rec["frac"]) - true
        rms[snr] = float(np.sqrt(np.mean(err ** 2)))
        assert np.max(np.abs(err)) < 0.5                       # always within half a sample of the truth
    assert rms[20] < 5e-3 and rms[0] < 5e-2 and rms[-10] < 0.3
    assert rms[20] < rms[0] < rms[-10]


def test_reference_worked_example_from_iq(golden_dir):
    """The reference's own worked example (tdoa_processor.py:472-490): three buoys whose detections are 150 000 and
    300 000 ns apart.  tests/golden/example_main.json holds what the reference's calculate_tdoa_measurements makes
    of those TIMESTAMPS; here the same arrival-time differences are put into synthetic cu8 IQ (307.2 and 614.4
    samples at 2.048 Msps) and measured by the GPU correlation path: same pairs, order, frequency and timing
    confidence, time / distance differences equal to the reference's within the sub-sample accuracy of the lag
    search -- the seam a10/a11 feed (tdoa_processor.py:166-170)."""
    import torch
    from radio_mapper_b200.tdoa_processor import TDOAProcessor, BuoyPosition
    with open(os.path.join(golden_dir, "example_main.json")) as f:
        ex = json.load(f)
    fs = ex["sample_rate"]
    dt_ns = [0, 150000, 300000]
    true = [t * 1e-9 * fs for t in dt_ns]                                 # 0, 307.2, 614.4 samples
    iq, _, _ = synth.delayed_buoys(77, 3, 1 << 20, sample_rate=fs, snr_db=20.0, max_delay=700,
                                   delays=[int(np.floor(t)) for t in true], frac_delays=[t - np.floor(t) for t in true])
    proc = TDOAProcessor()
    ids = []
    for b in ex["buoys"]:
        proc.register_buoy(BuoyPosition(*b))
        ids.append(b[0])
    meas = proc.correlate_iq(torch.from_numpy(iq[:, None, :]).pin_memory(), ids, fs, ex["frequency_mhz"])
    rec = proc.correlate_iq_records(torch.from_numpy(iq[:, None, :]).pin_memory())
    assert len(meas) == len(ex["measurements"]) == 3
    for m, coh, (b1, b2, dt, dd, conf, fmhz) in zip(meas, rec["coherence"][0], ex["measurements"]):
        assert (m.buoy1_id, m.buoy2_id, m.frequency_mhz) == (b1, b2, fmhz)
        assert isinstance(m.time_difference_ns, int)
        assert abs(m.time_difference_ns - dt) <= 1                       # 1 ns = 0.002 samples (the CPU oracle hits dt exactly)
        assert abs(m.distance_difference_m - dd) <= 0.31                 # 1 ns of light
        assert m.distance_difference_m == (m.time_difference_ns / 1e9) * 299792458.0      # :169-170 verbatim
        # confidence = strength term x the reference's timing term (:200-210); the reference's strength term is
        # min(c_i, c_j) of the detections, ours the measured coherence: the timing terms must be identical
        ci = {"BUOY_ALPHA": 0.9, "BUOY_BETA": 0.85, "BUOY_GAMMA": 0.88}
        assert abs(m.confidence / float(coh) - conf / min(ci[b1], ci[b2])) < 1e-6


def test_correlate_iq_edge_shapes():
    """Host-facing call on degenerate shapes: no windows -> empty record table / no measurements; two buoys -> one
    pair; the per-window generator (run_iter) hands back exactly the rows of run(); an unaligned (sliced) spectra
    buffer takes the load path without the bulk-copy prefetch and returns the same records."""
    import torch
    from radio_mapper_b200 import engine
    from radio_mapper_b200.tdoa_processor import TDOAProcessor
    n = 1 << 17                                                        # plan 32 x 4096: X_i-stationary row pass
    iq, d, _ = synth.delayed_buoys(5, 2, n)
    proc = TDOAProcessor()
    empty = proc.correlate_iq_records(torch.empty((2, 0, 2 * n), dtype=torch.uint8).pin_memory())
    assert empty.shape == (0, 1)
    assert proc.correlate_iq(torch.empty((2, 0, 2 * n), dtype=torch.uint8).pin_memory(), ["A", "B"]) == []
    block = np.stack([iq, iq[::-1]], axis=1)                           # [B=2, W=2, 2N]; window 1 swaps the buoys
    rec = proc.correlate_iq_records(torch.from_numpy(np.ascontiguousarray(block)).pin_memory())
    assert rec.shape == (2, 1) and rec["lag"][0, 0] == d[1] - d[0] and rec["lag"][1, 0] == d[0] - d[1]
    cor = proc._correlator(2, n, None)
    rows = list(cor.run_iter(torch.from_numpy(np.ascontiguousarray(block)).cuda()))
    assert len(rows) == 2 and all(np.array_equal(rows[w][f], rec[w][f]) for w in range(2) for f in ("lag", "peak", "frac"))
    # spectra at an address that is 8- but not 16-byte aligned
    plan = engine.Plan(2, n)
    S = plan.forward(_cuda(iq))
    pairs = _cuda(engine.pair_table(2))
    a = engine.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs)).copy()
    buf = torch.empty(S.numel() + 1, dtype=torch.complex64, device="cuda")
    view = buf[1:].view(2, plan.fft_len)
    view.copy_(S)
    assert view.data_ptr() % 16 == 8
    b = engine.peaks_to_numpy(plan.xcorr_pairs_peak(view, pairs))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("log_n", [23, 24])
def test_fused_outer_passes_match_unfused(rmx, log_n):
    """Three-pass plans (L = 2^24: 64 x 64 x 4096, L = 2^25: 64 x 128 x 4096): the fused middle + outer pass
    (`fuse_outer`, slabs through an L2-resident scratch ring, arg-max on the fly, peak values from the row-pass
    output) against the separate launches through the workspace -- same lags, peaks and offsets inside the
    north_star tolerances of each other and of the generator / oracle; also with a restricted lag range (the
    masked arg-max) and with the pairs walked in chunks (ring and counters reused across launches)."""
    import torch
    n = 1 << log_n
    iq, delays = synth.delayed_buoys_torch(600 + log_n, 3, 1, n, torch.device("cuda"))
    pairs_h = rmx.pair_table(3)
    pairs = _cuda(pairs_h)
    want = delays[0, pairs_h[:, 1]] - delays[0, pairs_h[:, 0]]
    res = {}
    for name, opts in (("fused", {"fuse_outer": 1}), ("unfused", {"fuse_outer": 0})):
        plan = rmx.Plan(3, n, options=opts)
        assert len(plan.pass_lengths) == 3
        S = plan.forward(iq[:, 0, :])
        res[name] = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs)).copy()
        res[name + "_chunked"] = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs, max_pairs_in_flight=1)).copy()
        plan.set_max_lag(5000)
        plan.set_search_mode(True)                                  # full inverse + masked arg-max, not the one-pass search
        res[name + "_lag5000"] = rmx.peaks_to_numpy(plan.xcorr_pairs_peak(S, pairs)).copy()
        del plan
    assert np.array_equal(res["fused"]["lag"], want)
    for key in ("", "_chunked", "_lag5000"):
        _check_records(res["fused" + key], res["unfused" + key])
    assert np.array_equal(res["fused"]["lag"], res["fused_chunked"]["lag"])
    assert np.array_equal(res["fused"]["peak"], res["fused_chunked"]["peak"])
    if log_n == 23:                                                  # one pair against the CPU oracle (2^24-point correlation)
        ref = oracle.xcorr_pairs_peak(iq[:2, 0, :].cpu().numpy())
        _check_records(res["fused"][:1], ref)
