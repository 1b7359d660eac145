// rmx_fft_soa.cuh — two transforms per thread in packed fp32x2 registers (sm_100a FADD2 / FMUL2 / FFMA2).
//
// A thread of the scalar tile code (rmx_fft_core.cuh) holds E complex values of ONE transform as float2 (re, im):
// its butterfly adds are packed (one FADD2 per complex add) but every complex multiply — stage twiddles, inter-pass
// twiddles, the constant twiddles inside a radix-16 butterfly — costs four scalar FMUL/FFMA.  ncu shows the row
// passes balanced at ~55 % of issue slots and FMA pipe with 85 % of their instructions floating point, so the way to
// make them faster is fewer instructions per point.
//
// Here a thread holds the same E positions of TWO transforms that share every twiddle (two buoy pairs of one
// spectrum row, or the even / odd half-transforms of one 8192-point row): value u is C2{re = (re_A, re_B),
// im = (im_A, im_B)}.  Complex adds stay two packed instructions for two transforms, and a complex multiply by a
// twiddle becomes FMUL2, FFMA2, FMUL2, FFMA2 for BOTH transforms (operand negation and scalar immediates are free in
// the packed instructions), i.e. half the issue slots per point for all multiplies.  The exchange between radix
// stages goes through two shared-memory planes (one of (re_A, re_B), one of (im_A, im_B)) with the same padded,
// conflict-free 64-bit addressing as the scalar code, so shared-memory traffic per point is unchanged.
#pragma once
#include "rmx_fft_core.cuh"

namespace rmx {

struct C2 {
    float2 re, im;      // .x: transform A, .y: transform B
};

__device__ __forceinline__ C2 add2(C2 a, C2 b) { return C2{__fadd2_rn(a.re, b.re), __fadd2_rn(a.im, b.im)}; }
__device__ __forceinline__ C2 sub2(C2 a, C2 b) {
    const float2 m1 = make_float2(-1.f, -1.f);
    return C2{__ffma2_rn(b.re, m1, a.re), __ffma2_rn(b.im, m1, a.im)};
}
// a * w, lane-wise (w may differ between the lanes)
__device__ __forceinline__ C2 mul2(C2 a, C2 w) {
    C2 r;
    r.re = __ffma2_rn(a.im, make_float2(-w.im.x, -w.im.y), __fmul2_rn(a.re, w.re));
    r.im = __ffma2_rn(a.im, w.re, __fmul2_rn(a.re, w.im));
    return r;
}
// a * conj(w)
__device__ __forceinline__ C2 mul2_conj(C2 a, C2 w) {
    C2 r;
    r.re = __ffma2_rn(a.im, w.im, __fmul2_rn(a.re, w.re));
    r.im = __ffma2_rn(a.re, make_float2(-w.im.x, -w.im.y), __fmul2_rn(a.im, w.re));
    return r;
}
// the same twiddle (c + i s) on both lanes
__device__ __forceinline__ C2 dup2(float2 w) { return C2{make_float2(w.x, w.x), make_float2(w.y, w.y)}; }

// a * exp(-2*pi*i*EXP/32) on both lanes (constants become immediates of the packed instructions)
template <int EXP>
__device__ __forceinline__ C2 mul_w32_fwd2(C2 a) {
    constexpr int e = EXP & 31;
    constexpr float h = 0.70710678118654757f;
    const float2 m1 = make_float2(-1.f, -1.f);
    if constexpr (e == 0) return a;
    else if constexpr (e == 8) return C2{a.im, __fmul2_rn(a.re, m1)};
    else if constexpr (e == 16) return C2{__fmul2_rn(a.re, m1), __fmul2_rn(a.im, m1)};
    else if constexpr (e == 24) return C2{__fmul2_rn(a.im, m1), a.re};
    else if constexpr (e == 4) return C2{__fmul2_rn(__fadd2_rn(a.re, a.im), make_float2(h, h)),
                                         __fmul2_rn(__ffma2_rn(a.re, m1, a.im), make_float2(h, h))};
    else if constexpr (e == 12) return C2{__fmul2_rn(__ffma2_rn(a.re, m1, a.im), make_float2(h, h)),
                                          __fmul2_rn(__fadd2_rn(a.re, a.im), make_float2(-h, -h))};
    else if constexpr (e == 20) return C2{__fmul2_rn(__fadd2_rn(a.re, a.im), make_float2(-h, -h)),
                                          __fmul2_rn(__ffma2_rn(a.im, m1, a.re), make_float2(h, h))};
    else if constexpr (e == 28) return C2{__fmul2_rn(__ffma2_rn(a.im, m1, a.re), make_float2(h, h)),
                                          __fmul2_rn(__fadd2_rn(a.re, a.im), make_float2(h, h))};
    else {
        constexpr float c = kCos32[e], s = kSin32[e];
        return C2{__ffma2_rn(a.im, make_float2(s, s), __fmul2_rn(a.re, make_float2(c, c))),
                  __ffma2_rn(a.re, make_float2(-s, -s), __fmul2_rn(a.im, make_float2(c, c)))};
    }
}
template <int EXP, bool INV>
__device__ __forceinline__ C2 mul_w32_2(C2 a) {
    return mul_w32_fwd2<INV ? (32 - (EXP & 31)) & 31 : (EXP & 31)>(a);
}

template <int R, int LEN, bool INV>
__device__ __forceinline__ void dif_layers2(C2 (&x)[R]) {
    constexpr int half = LEN / 2;
    static_for<0, R / 2>([&](auto I) {
        constexpr int idx = decltype(I)::value;
        constexpr int blk = idx / half, k = idx % half;
        constexpr int a = blk * LEN + k, b = a + half;
        const C2 u = x[a], v = x[b];
        x[a] = add2(u, v);
        x[b] = mul_w32_2<(32 / LEN) * k, INV>(sub2(u, v));
    });
    if constexpr (LEN > 2) dif_layers2<R, LEN / 2, INV>(x);
}

// R-point DFTs of both transforms, natural order in and out
template <int R, bool INV>
__device__ __forceinline__ void dft_regs2(C2 (&x)[R]) {
    if constexpr (R > 1) {
        dif_layers2<R, R, INV>(x);
        C2 t[R];
        static_for<0, R>([&](auto I) { constexpr int q = decltype(I)::value; t[q] = x[bitrev(q, ilog2(R))]; });
        static_for<0, R>([&](auto I) { constexpr int q = decltype(I)::value; x[q] = t[q]; });
    }
}

// shared memory needed by fft_tile2: two planes of NP*G 8-byte units
template <class GEO>
__host__ __device__ constexpr size_t soa_smem_bytes() { return 2 * size_t(GEO::NP) * GEO::G * sizeof(float2); }

// All Stockham stages of BOTH transforms.  Same stage structure, thread map and shared-memory addressing as
// fft_tile; `plane_re` / `plane_im` are the two exchange planes (GEO::NP*GEO::G float2 each).  Stage twiddles are
// identical for the two transforms: the power-of-two table entries are read once (TWTREE as in fft_tile), duplicated
// onto both lanes and combined by packed products.  All 256 threads must call (contains barriers).
template <class GEO, bool INV, class Hook = NoHook>
__device__ __forceinline__ void fft_tile2(C2 (&r)[GEO::E], float2* plane_re, float2* plane_im, int g, int i0,
                                          const StageTables& tabs, Hook after_last_gather = Hook{}) {
    constexpr int E = GEO::E, LOGE = GEO::LOGE, LOGN = GEO::LOGN, NT = GEO::NT;
    constexpr int UNIT = GEO::COLUMN ? GEO::G : 1;
    static_for<0, GEO::NSTAGES>([&](auto S_) {
        constexpr int S = decltype(S_)::value;
        constexpr int LOGP = S * LOGE;
        constexpr int LOGR = cmin(LOGE, LOGN - LOGP);
        constexpr int R = 1 << LOGR;
        constexpr int NB = E / R;
        constexpr int P = 1 << LOGP;
        if constexpr (S > 0) {
            const float2* __restrict__ tw = tabs.tw[S];
            static_for<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                const int k = (i0 + b * NT) & (P - 1);
                if constexpr (LOGR >= 3) {
                    constexpr int LO = 4;                        // w^q = wl[q % LO] * wh[q / LO]
                    C2 pw[LOGR];
                    static_for<0, LOGR>([&](auto Z_) {
                        constexpr int z = decltype(Z_)::value;
                        float2 t = __ldg(tw + ((1 << z) - 1) * P + k);
                        if (INV) t.y = -t.y;
                        pw[z] = dup2(t);
                    });
                    C2 wl[LO];
                    wl[1] = pw[0]; wl[2] = pw[1]; wl[3] = mul2(pw[0], pw[1]);
                    C2 wh[R / LO];
                    static_for<1, R / LO>([&](auto M_) {
                        constexpr int m = decltype(M_)::value;
                        constexpr int top = ilog2(m + 1) - ((1 << (ilog2(m + 1))) > m ? 1 : 0);   // floor(log2 m)
                        if constexpr ((m & (m - 1)) == 0) wh[m] = pw[2 + top];
                        else wh[m] = mul2(wh[m - (1 << top)], pw[2 + top]);
                    });
                    static_for<1, R>([&](auto Q_) {
                        constexpr int q = decltype(Q_)::value;
                        constexpr int lo = q % LO, hi = q / LO;
                        C2 w;
                        if constexpr (hi == 0) w = wl[lo];
                        else if constexpr (lo == 0) w = wh[hi];
                        else w = mul2(wl[lo], wh[hi]);
                        r[b + q * NB] = mul2(r[b + q * NB], w);
                    });
                } else {
                    static_for<1, R>([&](auto Q_) {
                        constexpr int q = decltype(Q_)::value;
                        float2 t = __ldg(tw + (q - 1) * P + k);
                        if (INV) t.y = -t.y;
                        r[b + q * NB] = mul2(r[b + q * NB], dup2(t));
                    });
                }
            });
        }
        static_for<0, NB>([&](auto B_) {
            constexpr int b = decltype(B_)::value;
            C2 x[R];
            static_for<0, R>([&](auto Q_) { constexpr int q = decltype(Q_)::value; x[q] = r[b + q * NB]; });
            dft_regs2<R, INV>(x);
            static_for<0, R>([&](auto Q_) { constexpr int q = decltype(Q_)::value; r[b + q * NB] = x[q]; });
        });
        if constexpr (S + 1 < GEO::NSTAGES) {
            constexpr int LOGP2 = LOGP + LOGE;
            constexpr int LOGR2 = cmin(LOGE, LOGN - LOGP2);
            constexpr int R2 = 1 << LOGR2;
            constexpr int NB2 = E / R2;
            constexpr int T2 = 1 << (LOGN - LOGR2);
            static_assert(P == 1 || P % (1 << GEO::LOGR0) == 0, "stage stride must keep the padding additive");
            static_assert(T2 % (1 << GEO::LOGR0) == 0, "stage stride must keep the padding additive");
            constexpr int PSTEP = (P + (P >> GEO::LOGR0)) * UNIT;
            constexpr int TSTEP = (T2 + (T2 >> GEO::LOGR0)) * UNIT;
            if constexpr (S > 0) __syncthreads();
            static_for<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                const int i = i0 + b * NT;
                const int k = i & (P - 1);
                const int jbase = ((i >> LOGP) << (LOGP + LOGR)) | k;
                const int off = GEO::saddr(g, jbase);
                float2* __restrict__ dre = plane_re + off;
                float2* __restrict__ dim = plane_im + off;
                static_for<0, R>([&](auto Q_) {
                    constexpr int q = decltype(Q_)::value;
                    dre[q * PSTEP] = r[b + q * NB].re;
                    dim[q * PSTEP] = r[b + q * NB].im;
                });
            });
            __syncthreads();
            static_for<0, NB2>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                const int off = GEO::saddr(g, i0 + b * NT);
                const float2* __restrict__ sre = plane_re + off;
                const float2* __restrict__ sim = plane_im + off;
                static_for<0, R2>([&](auto Q_) {
                    constexpr int q = decltype(Q_)::value;
                    r[b + q * NB2].re = sre[q * TSTEP];
                    r[b + q * NB2].im = sim[q * TSTEP];
                });
            });
            if constexpr (S + 2 == GEO::NSTAGES) after_last_gather();
        }
    });
}

}  // namespace rmx
