// rmx_pair_pp.cuh — X_i-stationary 4096-point row pass with the FP32 pipe handed round between warp groups
// (included by rmx_kernels.cuh after rmx_fft_split.cuh).
//
// k_contig_pair_run spends a tile ~2000 cycles on the FP32 pipe and ~1800 cycles on shared-memory / L1 wavefronts
// (DESIGN.md section 3), and with three independent CTAs per SM the two are used almost one after the other: the
// CTAs drift into the same phase, fight for the FP32 pipe together and then queue for the LSU together (ncu: 44 % of
// the issue slots empty while "math pipe throttle" + "not selected" are the top stall reasons).  Here ONE CTA of
// NG * 256 threads holds NG tiles, one per warp group, and a token goes round the groups: a group runs a butterfly
// region (pair product / stage twiddles + radix-16 butterflies / inter-pass twiddles) only while it holds the token
// and does everything else -- landing-buffer reads, exchange scatter / gather, next-stage twiddle tree, row store --
// without it.  The regions of the groups are thereby interleaved by construction: while one group computes, the
// others move data.  The token is a ring of named barriers (bar.sync by the group that enters, bar.arrive by the
// group that leaves); everything inside a group is the same code as k_contig_pair_run<.., PREFETCH> with the block
// barriers replaced by a named barrier of the group and per-group mbarriers.
#pragma once

namespace rmx {

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// exchange between stage S and S+1 of the tile (split-phase, as fft_tile_split): scatter, arrive(FULL), build the
// twiddles of stage S+1 into wtw, wait(FULL), gather, arrive(FREE)
template <class GEO, int S, bool INV>
__device__ __forceinline__ void pp_exchange(float2 (&r)[GEO::E], float2 (&wtw)[GEO::E], float2* smem, int i0,
                                            const StageTables& tabs, SplitBarriers& sb) {
    constexpr int E = GEO::E, LOGE = GEO::LOGE, LOGN = GEO::LOGN, NT = GEO::NT;
    constexpr int LOGP = S * LOGE;
    constexpr int LOGR = cmin(LOGE, LOGN - LOGP);
    constexpr int R = 1 << LOGR;
    constexpr int NB = E / R;
    constexpr int P = 1 << LOGP;
    constexpr int LOGP2 = LOGP + LOGE;
    constexpr int LOGR2 = cmin(LOGE, LOGN - LOGP2);
    constexpr int R2 = 1 << LOGR2;
    constexpr int NB2 = E / R2;
    constexpr int P2 = 1 << LOGP2;
    constexpr int T2 = 1 << (LOGN - LOGR2);
    static_assert(!GEO::COLUMN && LOGR2 >= 3, "row tiles with radix >= 8 stages");
    constexpr int PSTEP = P + (P >> GEO::LOGR0);
    constexpr int TSTEP = T2 + (T2 >> GEO::LOGR0);
    if (sb.free_pending) {
        mbar_wait(&sb.bar[1], sb.par_free);
        sb.par_free ^= 1u;
        sb.free_pending = false;
    }
    static_for<0, NB>([&](auto B_) {
        constexpr int b = decltype(B_)::value;
        const int i = i0 + b * NT;
        const int k = i & (P - 1);
        const int jbase = ((i >> LOGP) << (LOGP + LOGR)) | k;
        float2* __restrict__ dst = smem + GEO::saddr(0, jbase);
        static_for<0, R>([&](auto Q_) {
            constexpr int q = decltype(Q_)::value;
            dst[q * PSTEP] = r[b + q * NB];
        });
    });
    split_arrive(&sb.bar[0]);
    {
        const float2* __restrict__ tw = tabs.tw[S + 1];
        static_for<0, NB2>([&](auto B_) {
            constexpr int b = decltype(B_)::value;
            const int k = (i0 + b * NT) & (P2 - 1);
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
        });
    }
    mbar_wait(&sb.bar[0], sb.par_full);
    sb.par_full ^= 1u;
    static_for<0, NB2>([&](auto B_) {
        constexpr int b = decltype(B_)::value;
        const float2* __restrict__ src = smem + GEO::saddr(0, i0 + b * NT);
        static_for<0, R2>([&](auto Q_) {
            constexpr int q = decltype(Q_)::value;
            r[b + q * NB2] = src[q * TSTEP];
        });
    });
    split_arrive(&sb.bar[1]);
    sb.free_pending = true;
}

// twiddle multiply (S > 0) + the radix butterflies of stage S
template <class GEO, int S, bool INV>
__device__ __forceinline__ void pp_butterflies(float2 (&r)[GEO::E], const float2 (&wtw)[GEO::E]) {
    constexpr int E = GEO::E, LOGE = GEO::LOGE, LOGN = GEO::LOGN;
    constexpr int LOGP = S * LOGE;
    constexpr int LOGR = cmin(LOGE, LOGN - LOGP);
    constexpr int R = 1 << LOGR;
    constexpr int NB = E / R;
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
}

// Grid: ceil(rows * ceil(n_items / RUN) / NG) CTAs of NG * 256 threads; warp group `grp` of CTA b does exactly what CTA
// b * NG + grp of k_contig_pair_run does.  Every group runs RUN slots of NSTAGES token regions each, whether or not
// a slot holds a pair, so the token ring never waits for a group that has run out of work.
template <int LOGN, int LOGE, int RUN, int NG>
__global__ void __launch_bounds__(kThreads * NG, 1) k_contig_pair_run_pp(const PassParams p) {
    using GEO = TileGeom<LOGN, LOGE, false>;
    constexpr int E = GEO::E, NT = GEO::NT, NS = GEO::NSTAGES;
    static_assert(GEO::G == 1 && NS >= 2 && NG >= 2 && NG <= 4, "one row per tile, 2..4 groups");
    extern __shared__ float2 smem_all[];
    constexpr int LAND_OFF = (GEO::NP + 15) & ~15;
    constexpr int GROUP_F2 = LAND_OFF + GEO::N;               // exchange area + landing buffer, 128-byte multiples
    constexpr uint32_t ROW_BYTES = (uint32_t)(GEO::N * sizeof(float2));
    __shared__ float2 s_pw_all[NG][8];
    __shared__ __align__(8) unsigned long long s_bars[NG][4];  // [0] landing buffer full, [1] exchange FULL, [2] exchange FREE
    const int grp = threadIdx.x / kThreads, i0 = threadIdx.x % kThreads;
    float2* smem = smem_all + grp * GROUP_F2;
    float2* land = smem + LAND_OFF;
    float2* s_pw = s_pw_all[grp];
    unsigned long long* mbar = &s_bars[grp][0];
    const int TOKEN = 1 + grp, TOKEN_NEXT = 1 + (grp + 1) % NG, GSYNC = 1 + NG + grp;

    const unsigned n_blocks = ((unsigned)p.n_items + RUN - 1) / RUN;
    const unsigned long long total = (unsigned long long)n_blocks << (p.logL - LOGN);
    const unsigned long long vb = (unsigned long long)blockIdx.x * NG + grp;
    const bool active = vb < total;
    const unsigned blk = active ? (unsigned)(vb % n_blocks) : 0u;
    const long long row = active ? (long long)(vb / n_blocks) : 0;
    const uint32_t rr = (uint32_t)row & ((1u << p.post_logn) - 1u);
    if (p.post_logm > 0 && i0 < LOGE) {
        const uint32_t mask = (p.post_logm >= 32) ? 0xffffffffu : ((1u << p.post_logm) - 1u);
        s_pw[i0] = unit_root((rr * ((uint32_t)NT << i0)) & mask, p.post_logm, true);
    }
    const int first = active ? (int)blk * RUN : 0;
    const int last = active ? min(first + RUN, p.n_items) : 0;
    SplitBarriers sb;
    sb.bar = &s_bars[grp][1];
    if (i0 == 0) {
        mbar_init(mbar, 1);
        mbar_init(&sb.bar[0], kThreads / 32);
        mbar_init(&sb.bar[1], kThreads / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (first < last) {
            const int2 pr0 = __ldg(p.pairs + first);
            mbar_expect_tx(mbar, ROW_BYTES);
            bulk_load_1d(land, p.spectra + ((long long)pr0.y << p.logL) + (row << LOGN), ROW_BYTES, mbar);
        }
    }
    const float2 tw_base = p.post_logm > 0 ? row_twiddle_base<E>(rr, (uint32_t)i0, p.post_logm, true, p.post_scale) : make_float2(1.f, 0.f);
    named_sync(GSYNC, kThreads);                              // barriers initialised, s_pw written
    if (grp == NG - 1) named_arrive(TOKEN_NEXT, 2 * kThreads); // the ring starts at group 0

    float2 a[E];
    int cur_i = -1;
    uint32_t parity = 0;
#pragma unroll 1
    for (int s = 0; s < RUN; ++s) {
        const int pidx = first + s;
        const bool have = pidx < last;                        // group-uniform
        float2 r[E], wtw[E];
        if (have) {
            const int2 pr = __ldg(p.pairs + pidx);
            if (pr.x != cur_i) {
                const float2* __restrict__ xi = p.spectra + ((long long)pr.x << p.logL) + (row << LOGN);
#pragma unroll
                for (int u = 0; u < E; ++u) a[u] = RMX_X_LOAD(xi + i0 + u * NT);
                cur_i = pr.x;
            }
            mbar_wait(mbar, parity);
            parity ^= 1u;
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = land[i0 + u * NT];
            named_sync(GSYNC, kThreads);                      // landing buffer consumed
            if (i0 == 0 && pidx + 1 < last) {
                const int2 prn = __ldg(p.pairs + pidx + 1);
                fence_proxy_async();
                mbar_expect_tx(mbar, ROW_BYTES);
                bulk_load_1d(land, p.spectra + ((long long)prn.y << p.logL) + (row << LOGN), ROW_BYTES, mbar);
            }
        }
        static_for<0, NS>([&](auto S_) {
            constexpr int S = decltype(S_)::value;
            named_sync(TOKEN, 2 * kThreads);                  // ---- token held from here ...
            if (have) {
                if constexpr (S == 0) {
#pragma unroll
                    for (int u = 0; u < E; ++u) r[u] = cmul_conj(r[u], a[u]);             // X_j * conj(X_i)
                }
                pp_butterflies<GEO, S, true>(r, wtw);
                if constexpr (S == NS - 1) {
                    if (p.post_logm > 0) {
                        float2 tw[E];
                        row_twiddles_from_base<E>(tw, tw_base, s_pw);
#pragma unroll
                        for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
                    }
                    if (p.scale != 1.0f) {
#pragma unroll
                        for (int u = 0; u < E; ++u) { r[u].x *= p.scale; r[u].y *= p.scale; }
                    }
                }
            }
            if (!(grp == NG - 1 && s == RUN - 1 && S == NS - 1))
                named_arrive(TOKEN_NEXT, 2 * kThreads);       // ---- ... to here
            if (have) {
                if constexpr (S + 1 < NS) {
                    pp_exchange<GEO, S, true>(r, wtw, smem, i0, p.tabs, sb);
                } else {
                    float2* __restrict__ out = p.dst + (long long)pidx * p.src_item_stride + (row << LOGN);
#pragma unroll
                    for (int u = 0; u < E; ++u) RMX_D_STORE(out + i0 + u * NT, r[u]);
                }
            }
        });
    }
}

}  // namespace rmx
