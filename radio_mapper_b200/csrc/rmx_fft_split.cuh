// rmx_fft_split.cuh — fft_tile with SPLIT-PHASE exchange barriers (included by rmx_kernels.cuh after the mbarrier helpers).
//
// fft_tile separates "scatter my outputs" from "gather my next inputs" with __syncthreads(), and the next stage then
// starts by building its twiddles (table loads + products) before it can touch the gathered data: every warp of the
// CTA idles through the barrier and again through the table-load latency.  Here the two block barriers per exchange
// are mbarriers used in split phase (one elected lane per warp arrives, everybody waits later):
//
//     wait(FREE)                  -- all warps have finished READING the previous exchange   (arrived long ago:
//     scatter outputs                the whole radix butterfly lies between that arrive and this wait)
//     arrive(FULL)
//     build the next stage's twiddles   <- independent of the exchange; the registers of the scattered tile are free
//     wait(FULL)
//     gather inputs
//     arrive(FREE)
//     multiply by the prepared twiddles, radix butterfly, ...
//
// so the barrier latency and the twiddle-table latency overlap with the twiddle arithmetic, and the FREE barrier
// costs nothing.  State (two mbarriers in shared memory + this thread's phase parities) persists across the tiles
// a CTA processes, so the loop barrier between tiles disappears as well.
#pragma once

namespace rmx {

struct SplitBarriers {
    unsigned long long* bar;     // [0] FULL, [1] FREE (shared memory, 8-byte aligned)
    uint32_t par_full = 0, par_free = 0;
    bool free_pending = false;   // a FREE arrive of this CTA has not been waited for yet
};

// all threads call; followed by a __syncthreads() of the caller before first use
__device__ __forceinline__ void split_init(SplitBarriers& sb, unsigned long long* bars) {
    sb.bar = bars;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], kThreads / 32);
        mbar_init(&bars[1], kThreads / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
__device__ __forceinline__ void split_arrive(unsigned long long* bar) {
    __syncwarp();                                           // orders the warp's shared-memory accesses before the arrive
    if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class GEO, bool INV>
__device__ __forceinline__ void fft_tile_split(float2 (&r)[GEO::E], float2* smem, int g, int i0, const StageTables& tabs,
                                               SplitBarriers& sb) {
    constexpr int E = GEO::E, LOGE = GEO::LOGE, LOGN = GEO::LOGN, NT = GEO::NT;
    constexpr int UNIT = GEO::COLUMN ? GEO::G : 1;
    float2 wtw[E];                                          // twiddles of the NEXT stage, built while the exchange settles
    static_for<0, GEO::NSTAGES>([&](auto S_) {
        constexpr int S = decltype(S_)::value;
        constexpr int LOGP = S * LOGE;
        constexpr int LOGR = cmin(LOGE, LOGN - LOGP);
        constexpr int R = 1 << LOGR;
        constexpr int NB = E / R;
        constexpr int P = 1 << LOGP;
        if constexpr (S > 0) {
            static_for<0, NB>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                static_for<1, R>([&](auto Q_) {
                    constexpr int q = decltype(Q_)::value;
                    r[b + q * NB] = cmul(r[b + q * NB], wtw[b + q * NB]);
                });
            });
        }
        static_for<0, NB>([&](auto B_) {
            constexpr int b = decltype(B_)::value;
            float2 x[R];
            static_for<0, R>([&](auto Q_) { constexpr int q = decltype(Q_)::value; x[q] = r[b + q * NB]; });
            dft_regs<R, INV>(x);
            static_for<0, R>([&](auto Q_) { constexpr int q = decltype(Q_)::value; r[b + q * NB] = x[q]; });
        });
        if constexpr (S + 1 < GEO::NSTAGES) {
            constexpr int LOGP2 = LOGP + LOGE;
            constexpr int LOGR2 = cmin(LOGE, LOGN - LOGP2);
            constexpr int R2 = 1 << LOGR2;
            constexpr int NB2 = E / R2;
            constexpr int P2 = 1 << LOGP2;
            constexpr int T2 = 1 << (LOGN - LOGR2);
            static_assert(P == 1 || P % (1 << GEO::LOGR0) == 0, "stage stride must keep the padding additive");
            static_assert(T2 % (1 << GEO::LOGR0) == 0, "stage stride must keep the padding additive");
            constexpr int PSTEP = (P + (P >> GEO::LOGR0)) * UNIT;
            constexpr int TSTEP = (T2 + (T2 >> GEO::LOGR0)) * UNIT;
            if (sb.free_pending) {                          // everyone has finished reading the previous exchange
                mbar_wait(&sb.bar[1], sb.par_free);
                sb.par_free ^= 1u;
                sb.free_pending = false;
            }
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
            split_arrive(&sb.bar[0]);
            // ---- twiddles of stage S+1: w_{P2*R2}^{q*k}, k = i mod P2, from the power-of-two table entries ----
            {
                const float2* __restrict__ tw = tabs.tw[S + 1];
                static_for<0, NB2>([&](auto B_) {
                    constexpr int b = decltype(B_)::value;
                    const int k = (i0 + b * NT) & (P2 - 1);
                    if constexpr (LOGR2 >= 3) {
                        constexpr int LO = 4;
                        float2 pw[LOGR2];
                        static_for<0, LOGR2>([&](auto Z_) {
                            constexpr int z = decltype(Z_)::value;
                            pw[z] = __ldg(tw + ((1 << z) - 1) * P2 + k);
                            if (INV) pw[z].y = -pw[z].y;
                        });
                        float2 wl[LO];
                        wl[1] = pw[0]; wl[2] = pw[1]; wl[3] = cmul(pw[0], pw[1]);
                        float2 wh[R2 / LO];
                        static_for<1, R2 / LO>([&](auto M_) {
                            constexpr int m = decltype(M_)::value;
                            constexpr int top = ilog2(m + 1) - ((1 << (ilog2(m + 1))) > m ? 1 : 0);   // floor(log2 m)
                            if constexpr ((m & (m - 1)) == 0) wh[m] = pw[2 + top];
                            else wh[m] = cmul(wh[m - (1 << top)], pw[2 + top]);
                        });
                        static_for<1, R2>([&](auto Q_) {
                            constexpr int q = decltype(Q_)::value;
                            constexpr int lo = q % LO, hi = q / LO;
                            if constexpr (hi == 0) wtw[b + q * NB2] = wl[lo];
                            else if constexpr (lo == 0) wtw[b + q * NB2] = wh[hi];
                            else wtw[b + q * NB2] = cmul(wl[lo], wh[hi]);
                        });
                    } else {
                        static_for<1, R2>([&](auto Q_) {
                            constexpr int q = decltype(Q_)::value;
                            float2 w = __ldg(tw + (q - 1) * P2 + k);
                            if (INV) w.y = -w.y;
                            wtw[b + q * NB2] = w;
                        });
                    }
                });
            }
            mbar_wait(&sb.bar[0], sb.par_full);
            sb.par_full ^= 1u;
            static_for<0, NB2>([&](auto B_) {
                constexpr int b = decltype(B_)::value;
                const float2* __restrict__ src = smem + GEO::saddr(g, i0 + b * NT);
                static_for<0, R2>([&](auto Q_) {
                    constexpr int q = decltype(Q_)::value;
                    r[b + q * NB2] = src[q * TSTEP];
                });
            });
            split_arrive(&sb.bar[1]);
            sb.free_pending = true;
        }
    });
}

}  // namespace rmx
