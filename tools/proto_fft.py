"""numpy prototype of the multi-pass FFT scheme used by the CUDA kernels.

Checks the index algebra only (not performance):
  * Stockham autosort stages with mixed radices (natural in -> natural out)
  * forward DIF passes:  x[natural] -> S[k_a][k_b][k_c]  (freq = k_a + n_a*(k_b + n_b*k_c))
  * inverse DIT passes from that layout back to natural-order lags
Run:  python tools/proto_fft.py
"""
import numpy as np


def stockham(x, radices, inverse=False):
    """x: (..., n) natural order. Emulates the per-stage formulas of the kernel."""
    n = x.shape[-1]
    sign = +1.0 if inverse else -1.0
    cur = x.astype(np.complex128).copy()
    p = 1
    for R in radices:
        T = n // R
        nxt = np.empty_like(cur)
        i = np.arange(T)
        k = i & (p - 1)
        jbase = (i - k) * R + k
        u = np.stack([cur[..., i + q * T] for q in range(R)], axis=-1)      # (..., T, R)
        q = np.arange(R)
        tw = np.exp(sign * 2j * np.pi * (k[:, None] * q[None, :]) / (p * R))
        u = u * tw
        # R-point DFT, natural order out
        W = np.exp(sign * 2j * np.pi * np.outer(q, q) / R)
        U = u @ W.T
        for qq in range(R):
            nxt[..., jbase + qq * p] = U[..., qq]
        cur = nxt
        p *= R
    assert p == n
    return cur


def plan_passes(logL, max_contig=12, max_col=9, min_n=4):
    """Return list of (n_t) for passes t=0..; last pass is the contiguous one."""
    if logL <= max_contig:
        return [1 << logL]
    rem = logL
    contig = min(max_contig, logL - min_n)
    rem -= contig
    cols = []
    while rem > 0:
        ncol_passes = -(-rem // max_col)
        a = -(-rem // ncol_passes)
        cols.append(1 << a)
        rem -= a
    return cols + [1 << contig]


def forward_passes(x, ns):
    """x natural (L,), returns S in 'digit-transposed' layout (flat array of L)."""
    L = x.shape[0]
    y = x.astype(np.complex128).copy()
    M = L
    for n in ns:
        s = M // n
        blocks = L // M
        v = y.reshape(blocks, n, s)                      # [beta][r][j]
        v = np.fft.fft(v, axis=1)                         # over r -> k
        if s > 1:
            k = np.arange(n)[:, None]
            j = np.arange(s)[None, :]
            v = v * np.exp(-2j * np.pi * (k * j) / M)
        y = v.reshape(L)
        M = s
    return y


def layout_freq_index(L, ns):
    """freq index held at each flat position of the transposed layout."""
    pos = np.arange(L)
    freq = np.zeros(L, dtype=np.int64)
    weight = 1
    M = L
    for n in ns:
        s = M // n
        digit = (pos // s) % n
        freq += digit * weight
        weight *= n
        M = s
    return freq


def inverse_passes(C, ns):
    """C in transposed layout -> natural-order ifft (unnormalised, times L)."""
    L = C.shape[0]
    y = C.astype(np.complex128).copy()
    # run passes in reverse: last pass (s=1) first
    Ms = []
    M = L
    for n in ns:
        Ms.append(M)
        M //= n
    for n, M in reversed(list(zip(ns, Ms))):
        s = M // n
        blocks = L // M
        v = y.reshape(blocks, n, s)                      # [beta][k][j]
        if s > 1:
            k = np.arange(n)[:, None]
            j = np.arange(s)[None, :]
            v = v * np.exp(+2j * np.pi * (k * j) / M)
        v = np.fft.ifft(v, axis=1) * n
        y = v.reshape(L)
    return y


def main():
    rng = np.random.default_rng(0)
    for n, rad in [(4096, [16, 16, 16]), (2048, [16, 16, 8]), (512, [16, 16, 2]),
                   (64, [16, 4]), (8192, [32, 32, 8]), (1024, [32, 32]), (128, [16, 8])]:
        x = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
        for inv in (False, True):
            got = stockham(x, rad, inv)
            ref = np.fft.ifft(x, axis=-1) * n if inv else np.fft.fft(x, axis=-1)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            assert err < 1e-12, (n, rad, inv, err)
    print("stockham stages OK")

    for logL in (10, 13, 16, 18, 20):
        for kw in ({}, {"max_contig": 6, "max_col": 4}):
            ns = plan_passes(logL, **kw)
            L = 1 << logL
            assert int(np.prod(ns)) == L, ns
            x = rng.standard_normal(L) + 1j * rng.standard_normal(L)
            S = forward_passes(x, ns)
            X = np.fft.fft(x)
            fi = layout_freq_index(L, ns)
            assert np.array_equal(np.sort(fi), np.arange(L))
            err = np.abs(S - X[fi]).max() / np.abs(X).max()
            assert err < 1e-10, (logL, ns, err)
            back = inverse_passes(S, ns) / L
            err = np.abs(back - x).max()
            assert err < 1e-10, (logL, ns, err)
            # correlation through the transposed layout
            N = L // 2
            a = np.zeros(L, complex); b = np.zeros(L, complex)
            a[:N] = rng.standard_normal(N) + 1j * rng.standard_normal(N)
            d = 37
            b[d:N] = a[:N - d]
            Sa, Sb = forward_passes(a, ns), forward_passes(b, ns)
            c = inverse_passes(Sb * np.conj(Sa), ns) / L
            m = int(np.argmax(np.abs(c)))
            lag = m if m < N else m - L
            assert lag == d, (lag, d)
            print(f"logL={logL} passes={ns} OK")


if __name__ == "__main__":
    main()
