// rmx_fft_core.cuh — register/shared-memory FFT building blocks for sm_100a.
//
// One CTA (256 threads) transforms one TILE of 256*E complex points held E-per-thread in
// registers.  A tile is G independent FFTs of length n (n*G == TILE).  Each FFT is a
// Stockham autosort pipeline of radix-E register DFTs; between stages the tile is
// exchanged through padded shared memory (conflict-free for 64-bit accesses).  The first
// stage loads straight from global memory into registers and the last stage stores
// straight from registers, so shared memory carries only the inter-stage exchanges.
//
// Thread <-> data map (both first-stage loads and last-stage stores):
//      thread (g, i0) holds rows  i0 + u*(n/E),  u = 0..E-1   of FFT g.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace rmx {

constexpr int kThreads = 256;
constexpr int kLogThreads = 8;
constexpr int kMaxStages = 4;

// ---------------------------------------------------------------------------------------
// small complex helpers
// ---------------------------------------------------------------------------------------
// Blackwell (sm_100) packed fp32x2 arithmetic: one FADD2 / FFMA2 instruction per complex add / sub
// halves the issue slots of the butterfly adds (the kernels are issue-bound, not FMA-pipe-bound).
#ifndef RMX_NO_PACKED_F32X2
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float cnorm2(float2 a) { return a.x * a.x + a.y * a.y; }

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2(int v) {
    int r = 0;
    while ((1 << r) < v) ++r;
    return r;
}

// cos/sin(2*pi*k/32), k = 0..31, correctly rounded to float.
constexpr float kCos32[32] = {
    1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
    0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f,
    0.0f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f,
    -0.70710678118654757f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f,
    -1.0f, -0.98078528040323043f, -0.92387953251128674f, -0.83146961230254524f,
    -0.70710678118654757f, -0.55557023301960218f, -0.38268343236508978f, -0.19509032201612825f,
    0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
    0.70710678118654757f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f};
constexpr float kSin32[32] = {
    0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
    0.70710678118654757f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
    1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
    0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f,
    0.0f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f,
    -0.70710678118654757f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f,
    -1.0f, -0.98078528040323043f, -0.92387953251128674f, -0.83146961230254524f,
    -0.70710678118654757f, -0.55557023301960218f, -0.38268343236508978f, -0.19509032201612825f};

// a * exp(-2*pi*i*EXP/32)   (forward twiddle; the inverse uses EXP -> 32-EXP)
template <int EXP>
__device__ __forceinline__ float2 mul_w32_fwd(float2 a) {
    constexpr int e = EXP & 31;
    constexpr float h = 0.70710678118654757f;
    if constexpr (e == 0) return a;
    else if constexpr (e == 8) return make_float2(a.y, -a.x);
    else if constexpr (e == 16) return make_float2(-a.x, -a.y);
    else if constexpr (e == 24) return make_float2(-a.y, a.x);
    else if constexpr (e == 4) return make_float2(h * (a.x + a.y), h * (a.y - a.x));
    else if constexpr (e == 12) return make_float2(h * (a.y - a.x), -h * (a.x + a.y));
    else if constexpr (e == 20) return make_float2(-h * (a.x + a.y), h * (a.x - a.y));
    else if constexpr (e == 28) return make_float2(h * (a.x - a.y), h * (a.x + a.y));
    else {
        constexpr float c = kCos32[e], s = kSin32[e];
        return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
    }
}
template <int EXP, bool INV>
__device__ __forceinline__ float2 mul_w32(float2 a) {
    return mul_w32_fwd<INV ? (32 - (EXP & 31)) & 31 : (EXP & 31)>(a);
}

// One radix-2 DIF layer over sub-transforms of length LEN inside an R-point register DFT.
template <int R, int LEN, bool INV>
__device__ __forceinline__ void dif_layers(float2 (&x)[R]) {
    constexpr int half = LEN / 2;
    static_for<0, R / 2>([&](auto I) {
        constexpr int idx = decltype(I)::value;
        constexpr int blk = idx / half, k = idx % half;
        constexpr int a = blk * LEN + k, b = a + half;
        const float2 u = x[a], v = x[b];
        x[a] = cadd(u, v);
        x[b] = mul_w32<(32 / LEN) * k, INV>(csub(u, v));
    });
    if constexpr (LEN > 2) dif_layers<R, LEN / 2, INV>(x);
}

// R-point DFT (R = 2..32) of a register array, natural order in and out.
template <int R, bool INV>
__device__ __forceinline__ void dft_regs(float2 (&x)[R]) {
    if constexpr (R > 1) {
        dif_layers<R, R, INV>(x);
        float2 t[R];
        static_for<0, R>([&](auto I) { constexpr int q = decltype(I)::value; t[q] = x[bitrev(q, ilog2(R))]; });
        static_for<0, R>([&](auto I) { constexpr int q = decltype(I)::value; x[q] = t[q]; });
    }
}

// ---------------------------------------------------------------------------------------
// tile geometry
// ---------------------------------------------------------------------------------------
template <int LOGN_, int LOGE_, bool COLUMN_>
struct TileGeom {
    static constexpr int LOGN = LOGN_, LOGE = LOGE_;
    static constexpr bool COLUMN = COLUMN_;
    static constexpr int E = 1 << LOGE;
    static constexpr int N = 1 << LOGN;
    static constexpr int LOGTILE = kLogThreads + LOGE;
    static constexpr int TILE = 1 << LOGTILE;
    static_assert(LOGN <= LOGTILE, "FFT longer than a tile");
    static constexpr int LOGG = LOGTILE - LOGN;          // FFTs per tile
    static constexpr int G = 1 << LOGG;
    static constexpr int LOGNT = cmax(LOGN - LOGE, 0);   // threads per FFT
    static constexpr int NT = 1 << LOGNT;
    static constexpr int ROWSTEP = NT;                   // rows held by a thread: i0 + u*ROWSTEP
    static constexpr int NSTAGES = (LOGN + LOGE - 1) / LOGE;
    static constexpr int LOGR0 = cmin(LOGE, LOGN);
    static constexpr int NP = N + (N >> LOGR0);          // padded FFT length
    static constexpr size_t SMEM_BYTES = NSTAGES > 1 ? size_t(NP) * G * sizeof(float2) : 0;
    static_assert(LOGN >= LOGE, "FFT shorter than a thread's register tile is not supported");

    __device__ static __forceinline__ int stage_logr(int s) { return cmin(LOGE, LOGN - s * LOGE); }
    // shared-memory element index of (fft g, position pos)
    __device__ static __forceinline__ int saddr(int g, int pos) {
        const int pp = pos + (pos >> LOGR0);
        if constexpr (COLUMN) return (pp << LOGG) + g;
        else return g * NP + pp;
    }
    // thread -> (fft g, first row i0)
    __device__ static __forceinline__ void thread_map(int v, int& g, int& i0) {
        if constexpr (COLUMN) { g = v & (G - 1); i0 = v >> LOGG; }
        else { i0 = v & (NT - 1); g = v >> LOGNT; }
    }
};

struct StageTables {
    // tw[s] : table for stage s (s >= 1), entries [(q-1)*p + k] = exp(-2*pi*i*q*k/(p*R)), q=1..R-1, k<p
    const float2* tw[kMaxStages];
};

struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};

// Run all Stockham stages on the register tile.  On entry r[u] = x[i0 + u*NT]; on exit
// r[u] = X[i0 + u*NT] (natural order).  All 256 threads must call (contains barriers).
//
// TWTREE: read only the power-of-two entries of each stage's twiddle table and build the rest.
// `after_last_gather` runs once every thread's last read of the exchange buffer has been ISSUED;
// a caller that wants to refill the buffer puts its own barrier inside the hook.
template <class GEO, bool INV, bool TWTREE = false, class Hook = NoHook>
__device__ __forceinline__ void fft_tile(float2 (&r)[GEO::E], float2* smem, int g, int i0, const StageTables& tabs,
                                         Hook after_last_gather = Hook{}) {
    constexpr int E = GEO::E, LOGE = GEO::LOGE, LOGN = GEO::LOGN, NT = GEO::NT;
    // Shared-memory addressing: element (g, pos) lives at unit*(pos + (pos >> LOGR0)) (+ g terms).
    // Every stride used below (T, P) is a multiple of 2^LOGR0 or the stage-0 row stride, so
    // pad(base + q*stride) == pad(base) + q*pad(stride): one address per butterfly, the rest
    // are compile-time immediates.
    constexpr int UNIT = GEO::COLUMN ? GEO::G : 1;
    static_for<0, GEO::NSTAGES>([&](auto S_) {
        constexpr int S = decltype(S_)::value;
        constexpr int LOGP = S * LOGE;
        constexpr int LOGR = cmin(LOGE, LOGN - LOGP);
        constexpr int R = 1 << LOGR;
        constexpr int NB = E / R;                 // butterflies per thread in this stage
        constexpr int P = 1 << LOGP;
        if constexpr (S > 0) {
            // twiddle by w_{P*R}^{q*k}, k = i mod P (the inputs were gathered at the end of stage S-1)
            const float2* __restrict__ tw = tabs.tw[S];
            static_for<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                const int k = (i0 + b * NT) & (P - 1);
                if constexpr (TWTREE && LOGR >= 3) {
                    // Only the power-of-two exponents are read from the table (each correctly rounded);
                    // the other w^q are products of at most log2(R) of them.  Trades table loads (L1
                    // wavefronts, long-scoreboard latency after the barrier) for FMA work.
                    constexpr int LO = 4;                        // w^q = wl[q % LO] * wh[q / LO]
                    float2 pw[LOGR];
                    static_for<0, LOGR>([&](auto Z_) {
                        constexpr int z = decltype(Z_)::value;
                        pw[z] = __ldg(tw + ((1 << z) - 1) * P + k);
                        if (INV) pw[z].y = -pw[z].y;
                    });
                    float2 wl[LO];
                    wl[1] = pw[0]; wl[2] = pw[1]; wl[3] = cmul(pw[0], pw[1]);
                    float2 wh[R / LO];
                    static_for<1, R / LO>([&](auto M_) {
                        constexpr int m = decltype(M_)::value;
                        constexpr int top = ilog2(m + 1) - ((1 << (ilog2(m + 1))) > m ? 1 : 0);   // floor(log2 m)
                        if constexpr ((m & (m - 1)) == 0) wh[m] = pw[2 + top];
                        else wh[m] = cmul(wh[m - (1 << top)], pw[2 + top]);
                    });
                    static_for<1, R>([&](auto Q_) {
                        constexpr int q = decltype(Q_)::value;
                        constexpr int lo = q % LO, hi = q / LO;
                        float2 w;
                        if constexpr (hi == 0) w = wl[lo];
                        else if constexpr (lo == 0) w = wh[hi];
                        else w = cmul(wl[lo], wh[hi]);
                        r[b + q * NB] = cmul(r[b + q * NB], w);
                    });
                } else {
                    static_for<1, R>([&](auto Q_) {
                        constexpr int q = decltype(Q_)::value;
                        float2 w = __ldg(tw + (q - 1) * P + k);
                        if (INV) w.y = -w.y;
                        r[b + q * NB] = cmul(r[b + q * NB], w);
                    });
                }
            });
        }
        // R-point DFTs
        static_for<0, NB>([&](auto B_) {
            constexpr int b = decltype(B_)::value;
            float2 x[R];
            static_for<0, R>([&](auto Q_) { constexpr int q = decltype(Q_)::value; x[q] = r[b + q * NB]; });
            dft_regs<R, INV>(x);
            static_for<0, R>([&](auto Q_) { constexpr int q = decltype(Q_)::value; r[b + q * NB] = x[q]; });
        });
        if constexpr (S + 1 < GEO::NSTAGES) {
            // exchange: scatter this stage's outputs, gather the next stage's inputs x[i + q*T2]
            constexpr int LOGP2 = LOGP + LOGE;
            constexpr int LOGR2 = cmin(LOGE, LOGN - LOGP2);
            constexpr int R2 = 1 << LOGR2;
            constexpr int NB2 = E / R2;
            constexpr int T2 = 1 << (LOGN - LOGR2);   // butterflies per FFT in the next stage
            static_assert(P == 1 || P % (1 << GEO::LOGR0) == 0, "stage stride must keep the padding additive");
            static_assert(T2 % (1 << GEO::LOGR0) == 0, "stage stride must keep the padding additive");
            // P == 1 (stage 0): jbase = i*R with R == 2^LOGR0, so pad(jbase + q) = pad(jbase) + q
            constexpr int PSTEP = (P + (P >> GEO::LOGR0)) * UNIT;
            constexpr int TSTEP = (T2 + (T2 >> GEO::LOGR0)) * UNIT;
            if constexpr (S > 0) __syncthreads();      // everyone has finished reading the previous exchange
            static_for<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                const int i = i0 + b * NT;
                const int k = i & (P - 1);
                const int jbase = ((i >> LOGP) << (LOGP + LOGR)) | k;
                float2* __restrict__ dst = smem + GEO::saddr(g, jbase);
                static_for<0, R>([&](auto Q_) {
                    constexpr int q = decltype(Q_)::value;
                    dst[q * PSTEP] = r[b + q * NB];
                });
            });
            __syncthreads();
            static_for<0, NB2>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                const float2* __restrict__ src = smem + GEO::saddr(g, i0 + b * NT);
                static_for<0, R2>([&](auto Q_) {
                    constexpr int q = decltype(Q_)::value;
                    r[b + q * NB2] = src[q * TSTEP];
                });
            });
            if constexpr (S + 2 == GEO::NSTAGES) after_last_gather();
        }
    });
}

// ---------------------------------------------------------------------------------------
// exact unit roots
// ---------------------------------------------------------------------------------------
// 2^e as a float, e in [-126, 127]
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }

// exp(sign * 2*pi*i * p / 2^logm), 0 <= p < 2^logm, sign = -1 forward / +1 inverse.
// The angle is formed exactly (integer p, power-of-two scaling) before sincospif.
__device__ __forceinline__ float2 unit_root(uint32_t p, int logm, bool inverse) {
    float c, s;
    if (logm <= 24) {
        // fold to [-m/2, m/2): |p| <= 2^23 is exact in float and 2p/m is an exact scaling
        const int32_t m = 1 << logm;
        int32_t sp = (int32_t)p;
        if (sp >= (m >> 1)) sp -= m;
        sincospif((float)sp * pow2f(1 - logm), &s, &c);
    } else {
        const uint32_t lo = p & 4095u, hi = p >> 12;
        float c1, s1, c2, s2;
        sincospif((float)lo * pow2f(1 - logm), &s1, &c1);
        const int32_t mh = 1 << (logm - 12);
        int32_t sh = (int32_t)hi;
        if (sh >= (mh >> 1)) sh -= mh;
        sincospif((float)sh * pow2f(13 - logm), &s2, &c2);
        c = c1 * c2 - s1 * s2;
        s = c1 * s2 + s1 * c2;
    }
    return make_float2(c, inverse ? s : -s);
}

// tw[u] = root^(j*(i0 + u*step_rows)) for u = 0..E-1 with root = exp(sign*2*pi*i/2^logm):
// the inter-pass twiddles of the rows a thread holds in column j.  step^(2^z) is evaluated
// exactly for even z and by one squaring for odd z; tw[] is built as a product tree of depth
// log2(E), so the worst-case error stays at a few ulp.
template <int E>
__device__ __forceinline__ void row_twiddles(float2 (&tw)[E], uint32_t j, uint32_t i0, uint32_t step_rows,
                                             int logm, bool inverse, float scale) {
    const uint32_t mask = (logm >= 32) ? 0xffffffffu : ((1u << logm) - 1u);
    tw[0] = unit_root((j * i0) & mask, logm, inverse);
    tw[0].x *= scale;                 // a power of two: exact, and it propagates through the product tree
    tw[0].y *= scale;
    float2 pw = make_float2(1.f, 0.f);
    static_for<0, ilog2(E)>([&](auto Z_) {
        constexpr int z = decltype(Z_)::value;
        if constexpr ((z & 1) == 0) pw = unit_root((j * (step_rows << z)) & mask, logm, inverse);
        else pw = make_float2(pw.x * pw.x - pw.y * pw.y, 2.0f * pw.x * pw.y);
        static_for<0, (1 << z)>([&](auto U_) {
            constexpr int u = decltype(U_)::value;
            tw[u + (1 << z)] = cmul(tw[u], pw);
        });
    });
}

// Same result as row_twiddles when the step powers pw[z] = root^(j*step_rows*2^z), z < log2(E), are
// shared by the whole CTA (one row per tile): they are read from shared memory, each an exactly
// reduced root, so a thread evaluates one sincospif instead of 1 + log2(E)/2.
template <int E>
__device__ __forceinline__ float2 row_twiddle_base(uint32_t j, uint32_t i0, int logm, bool inverse, float scale) {
    // tw[0] of row_twiddles_shared: depends on the row and the thread only, so a CTA that walks several pairs of
    // one row evaluates it once
    const uint32_t mask = (logm >= 32) ? 0xffffffffu : ((1u << logm) - 1u);
    float2 t = unit_root((j * i0) & mask, logm, inverse);
    t.x *= scale;
    t.y *= scale;
    return t;
}

template <int E>
__device__ __forceinline__ void row_twiddles_from_base(float2 (&tw)[E], float2 base, const float2* pw_shared) {
    tw[0] = base;
    static_for<0, ilog2(E)>([&](auto Z_) {
        constexpr int z = decltype(Z_)::value;
        const float2 pw = pw_shared[z];
        static_for<0, (1 << z)>([&](auto U_) {
            constexpr int u = decltype(U_)::value;
            tw[u + (1 << z)] = cmul(tw[u], pw);
        });
    });
}

template <int E>
__device__ __forceinline__ void row_twiddles_shared(float2 (&tw)[E], uint32_t j, uint32_t i0, int logm, bool inverse,
                                                    float scale, const float2* pw_shared) {
    row_twiddles_from_base<E>(tw, row_twiddle_base<E>(j, i0, logm, inverse, scale), pw_shared);
}

}  // namespace rmx
