// rmx_kernels.cuh — the tile kernels of the radio-mapper hot path (sm_100a).
//
// Data layout in HBM
//   cu8 input   : uint8 [signal][2*N]   interleaved I,Q (rtl_sdr raw; reference buoy_node.py:392-398)
//   spectrum    : float2[signal][L]     "digit-transposed": for pass lengths n_0..n_{m-1}
//                 (n_0*...*n_{m-1} = L) position ((k_0*n_1 + k_1)*n_2 + ...) + k_{m-1} holds
//                 frequency bin k_0 + n_0*(k_1 + n_1*(k_2 + ...)).  Both FFT directions run
//                 in place on this layout with no transposes; element-wise products are
//                 layout-agnostic.  Single-pass plans (L <= TILE) are natural order.
//   correlation : float2[pair][L]       workspace between the inverse passes
//
// Forward  (DIF):  pass t = column FFTs of length n_t at stride s_t = n_{t+1}*...*n_{m-1},
//                  then multiply by w_{M_t}^{j*k} (M_t = n_t*s_t, j = column, k = output row);
//                  the last pass (s = 1) is the contiguous kernel.
// Inverse  (DIT):  the same passes in reverse order with conjugate twiddles applied on the
//                  INPUT of each column pass; the last pass to run (pass 0) yields natural-order
//                  lags and is fused with the |c|^2 arg-max.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the driver entry point is resolved at run time)
#include "rmx_fft_core.cuh"

#ifndef RMX_X_LOAD
#define RMX_X_LOAD __ldg      // spectrum loads of the pair passes (experiment: __ldcg / __ldcs / __ldlu)
#endif
#ifndef RMX_D_STORE_MODE
#define RMX_D_STORE_MODE 0    // workspace stores of the pair passes: 0 plain, 1 __stcs (streaming), 2 __stcg
#endif
#if RMX_D_STORE_MODE == 1
#define RMX_D_STORE(ptr, v) __stcs((ptr), (v))
#elif RMX_D_STORE_MODE == 2
#define RMX_D_STORE(ptr, v) __stcg((ptr), (v))
#else
#define RMX_D_STORE(ptr, v) (*(ptr) = (v))
#endif
#ifndef RMX_PAIR_RUN_BULK_STORE
#define RMX_PAIR_RUN_BULK_STORE 0   // same for the X_i-stationary 4096-point pass: measured 2-3 % slower (3 CTAs/SM hide the stores)
#endif
#ifndef RMX_FWD_BULK_STORE
#define RMX_FWD_BULK_STORE 0    // forward row pass (C_FWD): measured 2 % slower with the bulk store
#endif
#ifndef RMX_PAIR_BULK_STORE
#define RMX_PAIR_BULK_STORE 1   // row pass output through shared memory + cp.async.bulk (one row per tile)
#endif
#ifndef RMX_PAIR_TWTREE
#define RMX_PAIR_TWTREE 1   // row passes and forward column passes: build stage twiddles from their power-of-two entries
#endif
#ifndef RMX_DBG_PAIR_NOLOAD
#define RMX_DBG_PAIR_NOLOAD 0    // timing experiments only: the row pass without its spectrum loads ...
#endif
#ifndef RMX_DBG_PAIR_NOSTORE
#define RMX_DBG_PAIR_NOSTORE 0   // ... and / or without its workspace stores (results are wrong by construction)
#endif
#ifndef RMX_PAIR_SPLIT
#define RMX_PAIR_SPLIT 1      // X_i-stationary row pass: split-phase exchange barriers (rmx_fft_split.cuh)
#endif
#ifndef RMX_PAIR_RUN_CTAS
#define RMX_PAIR_RUN_CTAS 3   // resident CTAs per SM for the X_i-stationary pair pass
#endif
#ifndef RMX_PAIR_CTAS
#define RMX_PAIR_CTAS 2     // resident CTAs per SM requested for the contiguous pair pass (32 values/thread)
#endif
#ifndef RMX_ARGMAX_CTAS
#define RMX_ARGMAX_CTAS 2   // resident CTAs per SM requested for the pre-twiddled arg-max column pass (32 values/thread)
#endif
#ifndef RMX_ARGMAX_TMA_CTAS
#define RMX_ARGMAX_TMA_CTAS 3   // same for the TMA-fed kernel: its tile arrives through shared memory, so 80 registers suffice
#endif

namespace rmx {

struct Partial {
    float val;      // |c|^2 of the best lag in this tile (-1 = none)
    uint32_t rank;  // lag + lag_neg_max (position in the lag-ordered 'full' output); lowest wins ties
};

struct PassParams {
    const float2* src;        // float2 input of in-place passes / correlation workspace
    float2* dst;              // float2 output
    const uint8_t* cu8;       // raw IQ input (FWD_CU8 modes)
    const float* window;      // optional per-sample window (Welch), nullptr = none
    const float2* spectra;    // spectra base for the pair product
    const int2* pairs;        // (i, j) signal indices per pair
    Partial* partials;        // arg-max partials [item][tiles_per_item]
    float* accum;             // PSD accumulator (PSD mode)
    StageTables tabs;
    long long n_samples;      // valid samples per signal (N); the rest of L is zero padding
    long long cu8_stride;     // bytes between consecutive signals in cu8
    long long src_item_stride;  // elements between consecutive items of src/dst (normally L)
    int n_items;              // signals or pairs covered by this launch
    int logL;                 // log2 of the full transform length
    int logS;                 // log2 of the column stride (column kernels)
    int lag_pos_max;          // arg-max searches lags in [-lag_neg_max, lag_pos_max]
    int lag_neg_max;
    int items_per_cta;        // PSD mode: signals accumulated by one CTA; windowed mode: rows per CTA
    float scale;              // scale folded into the pair product (1/L)
    // windowed lag search (C_INV_PAIR_WIN*): per-(pair, row chunk) partial sums of the kept lags
    float2* win_partials;     // [pair][chunk][2*WU*NT]
    int win_lag_max;          // keep |lag| <= win_lag_max  (< WU*NT)
    int row_npass;            // number of outer (column) passes whose digits make up a row index
    int row_logn[kMaxStages]; // their lengths (log2), outermost first
    // Inter-pass twiddles of the inverse transform ride on the OUTPUT of the contiguous pass
    // (C_INV_PAIR), which is bound by shared-memory bandwidth and has FMA slots to spare, instead of
    // the input of the column pass that follows, which is FMA-bound:
    int post_logm;            // C_INV_PAIR: log2(M) of the next column pass (0 = do not twiddle)
    int post_logn;            // C_INV_PAIR: log2 of that pass's transform length (row index = row mod n)
    float post_scale;         // C_INV_PAIR: scale folded into those twiddles (a power of two)
    int prefetch;             // C_FWD_PSD: next segment's row by bulk copy into a landing buffer behind the exchange area
    int xi_early;             // k_contig_pair_run: a new X_i row is loaded one pair ahead (right after the pair product that last used the old one)
};

enum ContigMode { C_FWD = 0, C_FWD_CU8 = 1, C_INV_PAIR = 2, C_FWD_PSD = 3, C_INV_PAIR_WIN2 = 4, C_INV_PAIR_WIN4 = 5, C_INV_PAIR_WIN8 = 6 };
__host__ __device__ constexpr int window_wu(int mode) { return mode == C_INV_PAIR_WIN2 ? 2 : mode == C_INV_PAIR_WIN4 ? 4 : mode == C_INV_PAIR_WIN8 ? 8 : 0; }

// frequency offset phi(rho) of row rho of the digit-transposed layout: row rho = (k_0*n_1 + k_1)*...
// holds the bins  phi + (n_0*n_1*...)*k_c  with  phi = k_0 + n_0*(k_1 + n_1*(...)).
__device__ __forceinline__ uint32_t row_frequency(const PassParams& p, uint32_t rho) {
    uint32_t phi = 0;
    int shift = 0;
    for (int t = 0; t < p.row_npass; ++t) shift += p.row_logn[t];
    int weight = 0;
    for (int t = 0; t < p.row_npass; ++t) {
        shift -= p.row_logn[t];
        phi |= ((rho >> shift) & ((1u << p.row_logn[t]) - 1u)) << weight;
        weight += p.row_logn[t];
    }
    return phi;
}
// *_PRE: the input already carries this pass's twiddles and scale (applied by the contiguous pass)
enum ColMode { K_FWD_CU8 = 0, K_FWD = 1, K_INV = 2, K_INV_ARGMAX = 3, K_INV_PRE = 4, K_INV_ARGMAX_PRE = 5 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global bulk copy (TMA engine); the issuing thread must wait for the read of shared memory
// (bulk_store_wait_read) before the buffer is reused or the CTA exits
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// TMA helpers (sm_100a): 2-D tiled bulk-tensor loads completing on an mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#ifndef RMX_MBAR_SLEEP_NS
#define RMX_MBAR_SLEEP_NS 100   // back-off between polls of a barrier that is not ready yet: without it 7.6 % of the row pass's
                                // issued instructions were try_wait + branch (ncu source page); measured -1.5 % at cfg4, neutral elsewhere
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (RMX_MBAR_SLEEP_NS > 0 && !done) __nanosleep(RMX_MBAR_SLEEP_NS);
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy (TMA engine, 1-D): `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load_1d(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace rmx
#include "rmx_fft_split.cuh"
namespace rmx {


__device__ __forceinline__ float2 load_cu8_sample(const uint8_t* base, long long idx) {
    // (float)u8 - 127.5f, I then Q: exactly the reference's unpack (buoy_node.py:392-398)
    const uchar2 b = *reinterpret_cast<const uchar2*>(base + 2 * idx);
    return make_float2((float)b.x - 127.5f, (float)b.y - 127.5f);
}

__device__ __forceinline__ bool better(float v, uint32_t rank, float bv, uint32_t brank) {
    // np.argmax semantics on the lag-ordered 'full' output: first (lowest-lag) maximum wins
    return v > bv || (v == bv && rank < brank);
}

// Block-wide arg-max.  Every thread passes its best value `v` (-1 if none) and a functor that
// returns the lowest rank among the thread's elements equal to a given value.  The common path
// costs one float max-reduction; only threads holding the block maximum evaluate ranks.
// Result valid in thread 0.
template <class RankOf>
__device__ __forceinline__ void block_argmax(float& v, uint32_t& rank, RankOf rank_of) {
    __shared__ float s_v[kThreads / 32];
    __shared__ float s_max;
    __shared__ unsigned s_rank;
    float m = v;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = m;
    if (threadIdx.x == 0) s_rank = 0xffffffffu;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = s_v[0];
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w) t = fmaxf(t, s_v[w]);
        s_max = t;
    }
    __syncthreads();
    const float bm = s_max;
    if (v == bm && bm >= 0.f) atomicMin(&s_rank, rank_of(bm));
    __syncthreads();
    v = bm;
    rank = s_rank;
}

// ---------------------------------------------------------------------------------------
// contiguous pass: FFTs over rows of n consecutive elements
// ---------------------------------------------------------------------------------------
// resident CTAs per SM the register allocator must leave room for: 32 values per thread need
// ~128 registers (2 CTAs), 16 values per thread fit in 80 (3 CTAs)
__host__ __device__ constexpr int min_ctas(int loge) { return loge >= 5 ? 2 : loge == 4 ? 3 : 4; }
__host__ __device__ constexpr int argmax_tma_ctas(int loge) { return loge >= 5 ? RMX_ARGMAX_TMA_CTAS : 4; }

template <int LOGN, int LOGE, int MODE>
__global__ void __launch_bounds__(kThreads, (MODE == 3 /* C_FWD_PSD keeps E accumulators */ || MODE >= 4) ? 2 : (MODE == 2 && LOGE == 5) ? RMX_PAIR_CTAS : min_ctas(LOGE))
k_contig(const PassParams p) {
    using GEO = TileGeom<LOGN, LOGE, false>;
    constexpr int E = GEO::E, NT = GEO::NT, G = GEO::G;
    constexpr bool INV = (MODE == C_INV_PAIR || window_wu(MODE) > 0);
    extern __shared__ float2 smem[];

    int g, i0;
    GEO::thread_map(threadIdx.x, g, i0);
    const int log_rows = p.logL - LOGN;               // rows per item (log2)
    float2 r[E];

    if constexpr (window_wu(MODE) > 0) {
        // Windowed lag search in ONE pass over the spectra.  For |lag| <= win_lag_max only the first
        // and last WU*NT outputs of every row iFFT matter:
        //     c[lag] = sum_rows  w_L^{+phi(row)*lag} * ifft_row(X_j conj X_i)[lag mod n]
        // Each CTA owns a chunk of consecutive rows of one pair, keeps the 2*WU kept outputs per
        // thread in registers across rows, and writes one partial vector per chunk; no correlation
        // workspace is written.  Tiles are pair-fastest so spectrum rows are shared through L2.
        static_assert(GEO::G == 1, "windowed mode needs one row per tile");
        constexpr int WU = window_wu(MODE);
        static_assert(2 * WU <= E, "window wider than the row");
        const unsigned pair_idx = blockIdx.x % (unsigned)p.n_items;
        const unsigned chunk = blockIdx.x / (unsigned)p.n_items;
        const int2 pr = p.pairs[pair_idx];
        const float2* __restrict__ xi = p.spectra + ((long long)pr.x << p.logL);
        const float2* __restrict__ xj = p.spectra + ((long long)pr.y << p.logL);
        float2 acc[2 * WU];
#pragma unroll
        for (int a = 0; a < 2 * WU; ++a) acc[a] = make_float2(0.f, 0.f);
        const uint32_t lmask = (1u << p.logL) - 1u;
        const unsigned row0 = chunk * (unsigned)p.items_per_cta;
        for (int rr = 0; rr < p.items_per_cta; ++rr) {
            const unsigned row = row0 + rr;
            const long long off = ((long long)row << LOGN) + i0;
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const float2 a = __ldg(xi + off + u * NT), b = __ldg(xj + off + u * NT);
                r[u] = cmul_conj(b, a);
            }
            fft_tile<GEO, true, (RMX_PAIR_TWTREE != 0)>(r, smem, g, i0, p.tabs);
            // twiddle of (row, lag): w_L^{+phi*lag} with lag = a*NT + i0  ->  A * B^a,
            // A = w^{phi*i0} (per thread), B = w^{phi*NT} (per row); exact roots, short product chains
            const uint32_t phi = row_frequency(p, row);
            const float2 A = unit_root((phi * (uint32_t)i0) & lmask, p.logL, true);
            const float2 B = unit_root((phi * (uint32_t)NT) & lmask, p.logL, true);
            float2 t = A;
#pragma unroll
            for (int a = 0; a < WU; ++a) {                    // lags  a*NT + i0  >= 0   (outputs u = a)
                if (a * NT <= p.win_lag_max) {                // warp-uniform
                    if (a * NT + i0 <= p.win_lag_max) {
                        const float2 v = cmul(r[a], t);
                        acc[a].x += v.x;
                        acc[a].y += v.y;
                    }
                    t = cmul(t, B);
                }
            }
            t = A;
#pragma unroll
            for (int a = 1; a <= WU; ++a) {                   // lags  i0 - a*NT  < 0   (outputs u = E - a)
                if ((a - 1) * NT < p.win_lag_max) {           // warp-uniform
                    t = cmul_conj(t, B);
                    if (i0 - a * NT >= -p.win_lag_max) {
                        const float2 v = cmul(r[E - a], t);
                        acc[2 * WU - a].x += v.x;
                        acc[2 * WU - a].y += v.y;
                    }
                }
            }
            __syncthreads();                                  // exchange buffer is reused by the next row
        }
        const unsigned n_chunks = gridDim.x / (unsigned)p.n_items;
        float2* __restrict__ out = p.win_partials + ((long long)pair_idx * n_chunks + chunk) * (2 * WU * NT);
#pragma unroll
        for (int a = 0; a < 2 * WU; ++a) out[a * NT + i0] = acc[a];
        return;
    } else if constexpr (MODE == C_FWD_PSD) {
        // Welch: accumulate |X|^2 of the same rows over items_per_cta consecutive signals.
        // grid = (tiles_per_signal, signal chunks)
        float acc[E];
#pragma unroll
        for (int u = 0; u < E; ++u) acc[u] = 0.f;
        const long long row = (long long)blockIdx.x * G + g;          // row within a signal
        const int first = blockIdx.y * p.items_per_cta;
        const int last = min(first + p.items_per_cta, p.n_items);
        if constexpr (G == 1 && GEO::NSTAGES > 1) {
            if (p.prefetch) {
                // the row of the NEXT segment arrives by one bulk copy (TMA engine, mbarrier-completed) while this
                // one is transformed: the pass streams the spectra workspace from HBM, so every iteration would
                // otherwise start with a full-latency load
                __shared__ __align__(8) unsigned long long mbar;
                constexpr uint32_t ROW_BYTES = (uint32_t)(GEO::N * sizeof(float2));
                float2* land = smem + ((GEO::NP + 15) & ~15);
                const float2* __restrict__ base = p.src + (row << LOGN);
                if (threadIdx.x == 0) {
                    mbar_init(&mbar, 1);
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                    if (first < last) {
                        mbar_expect_tx(&mbar, ROW_BYTES);
                        bulk_load_1d(land, base + (long long)first * p.src_item_stride, ROW_BYTES, &mbar);
                    }
                }
                __syncthreads();
                uint32_t parity = 0;
                for (int it = first; it < last; ++it) {
                    mbar_wait(&mbar, parity);
                    parity ^= 1u;
#pragma unroll
                    for (int u = 0; u < E; ++u) r[u] = land[i0 + u * NT];
                    __syncthreads();      // landing buffer consumed; everyone is past the previous exchange reads
                    if (threadIdx.x == 0 && it + 1 < last) {
                        fence_proxy_async();
                        mbar_expect_tx(&mbar, ROW_BYTES);
                        bulk_load_1d(land, base + (long long)(it + 1) * p.src_item_stride, ROW_BYTES, &mbar);
                    }
                    fft_tile<GEO, false>(r, smem, g, i0, p.tabs);
#pragma unroll
                    for (int u = 0; u < E; ++u) acc[u] += cnorm2(r[u]);
                }
                float* __restrict__ outp = p.accum + (row << LOGN);
#pragma unroll
                for (int u = 0; u < E; ++u) atomicAdd(outp + i0 + u * NT, acc[u]);
                return;
            }
        }
        for (int it = first; it < last; ++it) {
            const float2* __restrict__ in = p.src + (long long)it * p.src_item_stride + (row << LOGN);
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = in[i0 + u * NT];
            fft_tile<GEO, false>(r, smem, g, i0, p.tabs);
#pragma unroll
            for (int u = 0; u < E; ++u) acc[u] += cnorm2(r[u]);
            if constexpr (GEO::NSTAGES > 1) __syncthreads();          // exchange buffer is reused
        }
        float* __restrict__ out = p.accum + (row << LOGN);
#pragma unroll
        for (int u = 0; u < E; ++u) atomicAdd(out + i0 + u * NT, acc[u]);
        return;
    } else {
        // decode (item, row)
        long long item, row;
        if (log_rows >= GEO::LOGG) {
            // item-fastest tile order: CTAs that run together touch the same rows of
            // different pairs, so each spectrum row is fetched from HBM once and then hit in L2
            item = blockIdx.x % (unsigned)p.n_items;
            row = (long long)(blockIdx.x / (unsigned)p.n_items) * G + g;
        } else {
            const long long flat = (long long)blockIdx.x * G + g;
            item = flat >> log_rows;
            row = flat & ((1LL << log_rows) - 1);
        }
        const bool active = item < p.n_items;
        __shared__ float2 s_pw[8];
        if constexpr (MODE == C_INV_PAIR && G == 1) {
            if (p.post_logm > 0 && threadIdx.x < LOGE) {
                const uint32_t rr = (uint32_t)row & ((1u << p.post_logn) - 1u);
                const uint32_t mask = (p.post_logm >= 32) ? 0xffffffffu : ((1u << p.post_logm) - 1u);
                s_pw[threadIdx.x] = unit_root((rr * ((uint32_t)NT << threadIdx.x)) & mask, p.post_logm, true);
            }
        }

        if constexpr (MODE == C_FWD) {
            const float2* __restrict__ in = p.src + item * p.src_item_stride + (row << LOGN);
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = active ? in[i0 + u * NT] : make_float2(0.f, 0.f);
        } else if constexpr (MODE == C_FWD_CU8) {
            // single-pass plans only (n == L, one row per signal)
            const uint8_t* __restrict__ in = p.cu8 + item * p.cu8_stride;
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const long long idx = (row << LOGN) + i0 + u * NT;
                r[u] = (active && idx < p.n_samples) ? load_cu8_sample(in, idx) : make_float2(0.f, 0.f);
            }
        } else {  // C_INV_PAIR
            int2 pr = make_int2(0, 0);
            if (active) pr = p.pairs[item];
            const float2* __restrict__ xi = p.spectra + ((long long)pr.x << p.logL) + (row << LOGN);
            const float2* __restrict__ xj = p.spectra + ((long long)pr.y << p.logL) + (row << LOGN);
#pragma unroll
            for (int u = 0; u < E; ++u) {
                if (active) {
                    const float2 a = RMX_X_LOAD(xi + i0 + u * NT), b = RMX_X_LOAD(xj + i0 + u * NT);
                    r[u] = cmul_conj(b, a);                        // X_j * conj(X_i)
                } else {
                    r[u] = make_float2(0.f, 0.f);
                }
            }
        }

        fft_tile<GEO, INV, (RMX_PAIR_TWTREE != 0)>(r, smem, g, i0, p.tabs);

        if constexpr (MODE == C_INV_PAIR) {
            if (p.post_logm > 0) {
                // input twiddles w_M^{+r*j} of the column pass that consumes this row (r = row within
                // its block, j = column = position in the row), times its scale
                const uint32_t rr = (uint32_t)row & ((1u << p.post_logn) - 1u);
                float2 tw[E];
                if constexpr (G == 1) {
                    // the step powers w_M^{r*NT*2^z} are the same for the whole CTA: s_pw was filled
                    // before the first barrier of fft_tile
                    row_twiddles_shared<E>(tw, rr, (uint32_t)i0, p.post_logm, true, p.post_scale, s_pw);
                } else {
                    row_twiddles<E>(tw, rr, (uint32_t)i0, (uint32_t)NT, p.post_logm, true, p.post_scale);
                }
#pragma unroll
                for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
            }
            // single-pass plans apply the 1/L here; multi-pass plans fold it into the twiddles of
            // the outermost column pass (p.scale == 1 for this launch)
            if (p.scale != 1.0f) {
#pragma unroll
                for (int u = 0; u < E; ++u) { r[u].x *= p.scale; r[u].y *= p.scale; }
            }
        }
        constexpr bool BULK = G == 1 && ((MODE == C_INV_PAIR && RMX_PAIR_BULK_STORE && (LOGE == 5 || RMX_PAIR_BULK_STORE > 1)) ||
                                         (MODE == C_FWD && RMX_FWD_BULK_STORE && (LOGE == 5 || RMX_FWD_BULK_STORE > 1)));
        // cp.async.bulk needs a 16-byte aligned destination; rows are multiples of 32 KB apart, so only the
        // base pointer matters (launch-uniform)
        if (BULK && (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0) {
            // the row leaves through the exchange buffer and ONE bulk copy (TMA engine) instead of 32
            // STG per thread: same shared-memory wavefronts as the stores' L1 wavefronts, but the warps
            // do not sit in the LSU queue behind 64 KB of write-through traffic
            __syncthreads();                                   // every thread is past its last exchange read
#pragma unroll
            for (int u = 0; u < E; ++u) smem[i0 + u * NT] = r[u];
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0 && active) {
                bulk_store_1d(p.dst + item * p.src_item_stride + (row << LOGN), smem, (uint32_t)(GEO::N * sizeof(float2)));
                bulk_store_wait_read();
            }
        } else if (active) {
            float2* __restrict__ out = p.dst + item * p.src_item_stride + (row << LOGN);
#pragma unroll
            for (int u = 0; u < E; ++u) {
                if constexpr (MODE == C_INV_PAIR) RMX_D_STORE(out + i0 + u * NT, r[u]);
                else out[i0 + u * NT] = r[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// column pass: FFTs of length n at element stride s = 2^logS, G adjacent columns per tile
// ---------------------------------------------------------------------------------------
template <int LOGN, int LOGE, int MODE>
__global__ void __launch_bounds__(kThreads, (MODE == K_INV_ARGMAX_PRE && LOGE == 5) ? RMX_ARGMAX_CTAS : min_ctas(LOGE)) k_col(const PassParams p) {
    using GEO = TileGeom<LOGN, LOGE, true>;
    constexpr int E = GEO::E, NT = GEO::NT, G = GEO::G, LOGG = GEO::LOGG;
    constexpr bool INV = (MODE == K_INV || MODE == K_INV_ARGMAX || MODE == K_INV_PRE || MODE == K_INV_ARGMAX_PRE);
    constexpr bool PRE = (MODE == K_INV_PRE || MODE == K_INV_ARGMAX_PRE);
    constexpr bool ARGMAX = (MODE == K_INV_ARGMAX || MODE == K_INV_ARGMAX_PRE);
    extern __shared__ float2 smem[];

    int g, i0;
    GEO::thread_map(threadIdx.x, g, i0);
    const int logS = p.logS;
    const int logM = LOGN + logS;
    const int log_tpb = logS - LOGG;                       // tiles per block (log2)
    const int log_bpi = p.logL - logM;                     // blocks per item (log2)
    const unsigned idx = blockIdx.x;
    const unsigned jt = idx & ((1u << log_tpb) - 1u);
    const unsigned rest = idx >> log_tpb;
    const unsigned beta = rest & ((1u << log_bpi) - 1u);
    const long long item = rest >> log_bpi;
    const unsigned j = (jt << LOGG) + g;                   // column within the block
    const long long base = ((long long)beta << logM) + j;  // element offset inside the item

    float2 r[E];
    if constexpr (MODE == K_FWD_CU8) {
        const uint8_t* __restrict__ in = p.cu8 + item * p.cu8_stride;
        const long long tile0 = base - g;                   // sample index of (row 0, column 0) of this tile
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(in) | (uintptr_t)p.cu8_stride) & 15) == 0;
        if (vec_ok) {
            // Stage the tile's raw bytes in shared memory with 128-bit loads: every row of the tile is
            // 2*G contiguous bytes (G >= 8), i.e. whole 16-byte chunks; rows past the valid samples
            // (zero padding) are never touched.
            constexpr int CPR = (2 * G) / 16;               // 16-byte chunks per row
            constexpr int CHUNKS = GEO::N * CPR;
            uint4* sb4 = reinterpret_cast<uint4*>(smem);
#pragma unroll
            for (int c = threadIdx.x; c < CHUNKS; c += kThreads) {
                const int row = c / CPR, part = c % CPR;
                const long long sidx = tile0 + ((long long)row << logS) + part * 8;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (sidx + 8 <= p.n_samples) {
                    v = __ldg(reinterpret_cast<const uint4*>(in + 2 * sidx));
                } else if (sidx < p.n_samples) {              // ragged tail of the last valid row
                    unsigned w[4] = {0u, 0u, 0u, 0u};
                    for (int b = 0; b < 16 && sidx * 2 + b < 2 * p.n_samples; ++b) w[b >> 2] |= (unsigned)in[2 * sidx + b] << (8 * (b & 3));
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                sb4[c] = v;
            }
            __syncthreads();
            const uchar2* sb2 = reinterpret_cast<const uchar2*>(smem);
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const int row = i0 + u * NT;
                const long long sidx = base + ((long long)row << logS);
                const uchar2 b = sb2[row * G + g];
                float2 v = make_float2(0.f, 0.f);
                if (sidx < p.n_samples) v = make_float2((float)b.x - 127.5f, (float)b.y - 127.5f);
                r[u] = v;
            }
            __syncthreads();                                 // the buffer becomes the exchange area
        } else {
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const long long sidx = base + ((long long)(i0 + u * NT) << logS);
                r[u] = sidx < p.n_samples ? load_cu8_sample(in, sidx) : make_float2(0.f, 0.f);
            }
        }
        if (p.window != nullptr) {
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const long long sidx = base + ((long long)(i0 + u * NT) << logS);
                if (sidx < p.n_samples) { const float w = __ldg(p.window + sidx); r[u].x *= w; r[u].y *= w; }
            }
        }
    } else {
        const float2* __restrict__ in = p.src + item * p.src_item_stride + base + ((long long)i0 << logS);
        const long long rstride = (long long)NT << logS;
#pragma unroll
        for (int u = 0; u < E; ++u) { r[u] = *in; in += rstride; }
    }

    if constexpr (INV && !PRE) {
        float2 tw[E];
        row_twiddles<E>(tw, j, (uint32_t)i0, (uint32_t)NT, logM, true, p.scale);
#pragma unroll
        for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
    }

    // (twiddle tree in the forward passes only: the inverse column kernels are register-tight)
    fft_tile<GEO, INV, (!INV && RMX_PAIR_TWTREE != 0)>(r, smem, g, i0, p.tabs);

    if constexpr (!INV) {
        float2 tw[E];
        row_twiddles<E>(tw, j, (uint32_t)i0, (uint32_t)NT, logM, false, 1.0f);
#pragma unroll
        for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
    }

    if constexpr (ARGMAX) {
        // pass 0 of the inverse: row m1, column j  ->  lag index m = m1*s + j (natural order).
        // rank = (m + lag_neg_max) mod L orders the lags like scipy's 'full' output; a lag is
        // searched iff rank <= lag_pos_max + lag_neg_max.
        const uint32_t lmask = (1u << p.logL) - 1u;
        const uint32_t span = (uint32_t)p.lag_pos_max + (uint32_t)p.lag_neg_max;
        const uint32_t rank0 = ((uint32_t)base + ((uint32_t)i0 << logS) + (uint32_t)p.lag_neg_max) & lmask;
        const uint32_t rstep = (uint32_t)NT << logS;
        float v[E];
        float bv = -1.f;
        // does any lag of this thread's column fall outside the searched range?  The column holds
        // m = j + row*s for every row; the excluded lags are the ranks in (span, L).
        const uint32_t col_rank0 = (j + (uint32_t)p.lag_neg_max) & lmask;          // rank of row 0
        const uint32_t excl = lmask - span;                                          // number of excluded ranks
        // first excluded rank congruent to col_rank0 modulo s, if there is one
        const uint32_t smask = (1u << logS) - 1u;
        const uint32_t first_excl = span + 1u + ((col_rank0 - (span + 1u)) & smask);
        const bool all_valid = excl == 0u || first_excl > lmask || first_excl < span + 1u;
        if (all_valid) {
#pragma unroll
            for (int u = 0; u < E; ++u) { v[u] = cnorm2(r[u]); bv = fmaxf(bv, v[u]); }
        } else {
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const uint32_t rank = (rank0 + (uint32_t)u * rstep) & lmask;
                v[u] = rank <= span ? cnorm2(r[u]) : -1.f;
                bv = fmaxf(bv, v[u]);
            }
        }
        if constexpr (GEO::NSTAGES > 1) __syncthreads();
        uint32_t brank;
        block_argmax(bv, brank, [&](float target) {
            uint32_t best = 0xffffffffu;
#pragma unroll
            for (int u = 0; u < E; ++u)
                if (v[u] == target) best = min(best, (rank0 + (uint32_t)u * rstep) & lmask);
            return best;
        });
        if (threadIdx.x == 0) {
            Partial out;
            out.val = bv;
            out.rank = brank;
            p.partials[(item << log_tpb) + jt] = out;
        }
    } else {
        float2* __restrict__ out = p.dst + item * p.src_item_stride + base + ((long long)i0 << logS);
        const long long rstride = (long long)NT << logS;
#pragma unroll
        for (int u = 0; u < E; ++u) { *out = r[u]; out += rstride; }
    }
}

// ---------------------------------------------------------------------------------------
// forward pass 0 from cu8, persistent and TMA-fed (north_star stage 2: "window tiles staged into SMEM by TMA")
// ---------------------------------------------------------------------------------------
// Same arithmetic as k_col<..., K_FWD_CU8>: column FFTs of length n at stride s over the raw cu8 samples of each
// signal, then the inter-pass twiddle.  The n x 2G-byte tile of raw bytes (n rows of the [row][2s bytes] view of
// the window, G adjacent columns) is brought in by 3-D TMA box loads {2G bytes, <=256 rows, 1 signal} completing
// on an mbarrier, double-buffered: the load of tile t+1 is issued before tile t is unpacked, so no thread ever
// waits on a global load of raw samples and no registers stage them.  Rows past the valid samples (zero padding)
// are outside the tensor map and arrive as zero bytes; the unpack still masks them (a zero BYTE is -127.5).
// tmap: UINT16 tensor (one element = one I,Q byte pair) {s, n_valid_rows, n_items}, strides {2s, cu8_stride} bytes,
// box {G, min(n,256), 1}.
template <int LOGN, int LOGE>
__global__ void __launch_bounds__(kThreads, min_ctas(LOGE)) k_col_fwd_cu8_tma(const PassParams p, const __grid_constant__ CUtensorMap tmap,
                                                                             const unsigned n_tiles) {
    using GEO = TileGeom<LOGN, LOGE, true>;
    constexpr int E = GEO::E, NT = GEO::NT, G = GEO::G, LOGG = GEO::LOGG, N = GEO::N;
    constexpr int BOX_ROWS = N < 256 ? N : 256;
    constexpr uint32_t STAGE_BYTES = (uint32_t)N * 2u * G;
    extern __shared__ float2 smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2];
    // [exchange area | stage 0 | stage 1], stages 128-byte aligned
    float2* smem = smem_raw;
    unsigned char* stage0 = reinterpret_cast<unsigned char*>(smem_raw) + ((GEO::SMEM_BYTES + 127) & ~size_t(127));
    stage0 += (128u - (smem_u32(stage0) & 127u)) & 127u;

    int g, i0;
    GEO::thread_map(threadIdx.x, g, i0);
    const int logS = p.logS;
    const int logM = LOGN + logS;
    const int log_tpb = logS - LOGG;                       // tiles per block (log2)
    const int log_bpi = p.logL - logM;                     // blocks per item (log2); 0 for pass 0 of a plan

    auto issue = [&](unsigned idx, int buf) {              // one thread
        const unsigned jt = idx & ((1u << log_tpb) - 1u);
        const unsigned item = idx >> (log_tpb + log_bpi);
        mbar_expect_tx(&mbar[buf], STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < N / BOX_ROWS; ++c)
            tma_load_3d(stage0 + buf * STAGE_BYTES + c * BOX_ROWS * 2 * G, &tmap, (int)(jt << LOGG), c * BOX_ROWS, (int)item, &mbar[buf]);
    };

    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned idx = blockIdx.x;
    if (threadIdx.x == 0 && idx < n_tiles) issue(idx, 0);
    uint32_t parity[2] = {0u, 0u};
    int buf = 0;

    for (; idx < n_tiles; idx += gridDim.x, buf ^= 1) {
        const unsigned next = idx + gridDim.x;
        // the other stage was last read two iterations ago, before that iteration's __syncthreads
        if (threadIdx.x == 0 && next < n_tiles) {
            fence_proxy_async();
            issue(next, buf ^ 1);
        }
        const unsigned jt = idx & ((1u << log_tpb) - 1u);
        const long long item = idx >> (log_tpb + log_bpi);
        const unsigned j = (jt << LOGG) + g;               // column within the block
        const long long base = j;                          // pass 0: one block per item

        mbar_wait(&mbar[buf], parity[buf]);
        parity[buf] ^= 1u;
        const uchar2* sb2 = reinterpret_cast<const uchar2*>(stage0 + buf * STAGE_BYTES);
        float2 r[E];
#pragma unroll
        for (int u = 0; u < E; ++u) {
            const int row = i0 + u * NT;
            const long long sidx = base + ((long long)row << logS);
            const uchar2 b = sb2[row * G + g];
            float2 v = make_float2(0.f, 0.f);
            if (sidx < p.n_samples) {
                v = make_float2((float)b.x - 127.5f, (float)b.y - 127.5f);
            }
            r[u] = v;
        }
        __syncthreads();                                   // stage consumed (it is refilled one iteration later); exchange area free

        fft_tile<GEO, false, (RMX_PAIR_TWTREE != 0)>(r, smem, g, i0, p.tabs);
        {
            float2 tw[E];
            row_twiddles<E>(tw, j, (uint32_t)i0, (uint32_t)NT, logM, false, 1.0f);
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
        }
        float2* __restrict__ out = p.dst + item * p.src_item_stride + base + ((long long)i0 << logS);
        const long long rstride = (long long)NT << logS;
#pragma unroll
        for (int u = 0; u < E; ++u) { *out = r[u]; out += rstride; }
        if constexpr (GEO::NSTAGES > 1) __syncthreads();   // exchange area is reused by the next tile
    }
}

// ---------------------------------------------------------------------------------------
// innermost inverse pass, X_i-stationary: one CTA walks RUN consecutive pairs of one row
// ---------------------------------------------------------------------------------------
// Same arithmetic as k_contig<..., C_INV_PAIR> for one row per tile (n == TILE).  Pair lists are
// ordered by first buoy (i < j, i-major), so consecutive pairs share X_i: the CTA keeps the X_i row
// in registers and re-reads it only when i changes.  That removes ~8*(1 - 1/RUN) of the 24 bytes
// per element this pass moves through L2 and L1 (it is bound by exactly that traffic), and the
// per-row twiddle powers are computed once per CTA instead of once per pair.
// PREFETCH: the X_j row of the NEXT pair of the run is brought into a 32 KB landing buffer by one bulk copy
// (TMA engine, completes on an mbarrier) while the CTA transforms the current pair, so only the first pair of a
// run waits for its spectrum row and no registers are tied up by loads in flight.
// STAGED (instead of PREFETCH; the same 32 KB buffer): the finished row leaves through a dedicated staging buffer and
// ONE bulk copy (TMA engine) that drains while the CTA already transforms the next pair -- no burst of 16 global
// stores per thread through the LSU queue that the next pair's shared-memory exchanges share.
// STAGED == 2 (with PREFETCH): the finished row is staged in the EXCHANGE buffer (free between the last gather of a
// pair and the first scatter of the next) and leaves by one bulk copy; the landing buffer keeps prefetching.
// XI_SMEM (without PREFETCH / STAGED; the same 32 KB buffer): the stationary X_i row lives in shared memory, each thread
// re-reading exactly the 16 values it wrote (no barrier), instead of in 32 registers next to the tile.
template <int LOGN, int LOGE, int RUN, bool PREFETCH, int STAGED = 0, int CTAS = RMX_PAIR_RUN_CTAS, bool XI_SMEM = false>
__global__ void __launch_bounds__(kThreads, CTAS) k_contig_pair_run(const PassParams p) {
    static_assert(!XI_SMEM || (!PREFETCH && STAGED == 0), "one 32 KB buffer: landing zone, store staging or the X_i row");
    static_assert(!(PREFETCH && STAGED == 1), "one 32 KB buffer: landing zone or store staging");
    static_assert(STAGED != 2 || (PREFETCH && RMX_PAIR_SPLIT), "exchange-buffer staging rides on the prefetch path's barriers");
    using GEO = TileGeom<LOGN, LOGE, false>;
    constexpr int E = GEO::E, NT = GEO::NT;
    static_assert(GEO::G == 1 && GEO::NSTAGES >= 2, "one row per tile");
    extern __shared__ float2 smem[];
    __shared__ float2 s_pw[8];
    __shared__ __align__(8) unsigned long long mbar;
    constexpr uint32_t ROW_BYTES = (uint32_t)(GEO::N * sizeof(float2));
    // landing buffer behind the exchange area, 128-byte aligned
    float2* land = smem + ((GEO::NP + 15) & ~15);

    const int i0 = threadIdx.x, g = 0;
    const unsigned n_blocks = ((unsigned)p.n_items + RUN - 1) / RUN;
    const unsigned blk = blockIdx.x % n_blocks;              // pair-block fastest: CTAs that run together share rows
    const long long row = blockIdx.x / n_blocks;
    const uint32_t rr = (uint32_t)row & ((1u << p.post_logn) - 1u);
    if (p.post_logm > 0 && threadIdx.x < LOGE) {
        const uint32_t mask = (p.post_logm >= 32) ? 0xffffffffu : ((1u << p.post_logm) - 1u);
        s_pw[threadIdx.x] = unit_root((rr * ((uint32_t)NT << threadIdx.x)) & mask, p.post_logm, true);
    }
    float2 a[E];
    int cur_i = -1;
    const int first = (int)blk * RUN;
    const int last = min(first + RUN, p.n_items);
    uint32_t parity = 0;
#if RMX_PAIR_SPLIT
    __shared__ __align__(8) unsigned long long s_split[2];
    SplitBarriers sb;
    split_init(sb, s_split);
    if constexpr (!PREFETCH) __syncthreads();                 // (the prefetch path has its own barrier below)
#endif
    // w_M^{row*i0} (times the scale): the same for every pair of the run -- one sincospif per CTA walk, not per pair
    const float2 tw_base = p.post_logm > 0 ? row_twiddle_base<E>(rr, (uint32_t)i0, p.post_logm, true, p.post_scale) : make_float2(1.f, 0.f);
    if constexpr (PREFETCH) {
        if (threadIdx.x == 0) {
            mbar_init(&mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const int2 pr0 = __ldg(p.pairs + first);
            mbar_expect_tx(&mbar, ROW_BYTES);
            bulk_load_1d(land, p.spectra + ((long long)pr0.y << p.logL) + (row << LOGN), ROW_BYTES, &mbar);
        }
        __syncthreads();                                      // the barrier is initialised before anyone waits on it
    }
    for (int pidx = first; pidx < last; ++pidx) {
        const int2 pr = __ldg(p.pairs + pidx);
        if (pr.x != cur_i) {                                  // CTA-uniform
            const float2* __restrict__ xi = p.spectra + ((long long)pr.x << p.logL) + (row << LOGN);
            if constexpr (XI_SMEM) {
#pragma unroll
                for (int u = 0; u < E; ++u) land[i0 + u * NT] = RMX_X_LOAD(xi + i0 + u * NT);
            } else {
#pragma unroll
                for (int u = 0; u < E; ++u) a[u] = RMX_X_LOAD(xi + i0 + u * NT);
            }
            cur_i = pr.x;
        }
        float2 r[E];
        if constexpr (PREFETCH) {
            mbar_wait(&mbar, parity);
            parity ^= 1u;
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = cmul_conj(land[i0 + u * NT], a[u]);            // X_j * conj(X_i)
            if constexpr (STAGED == 2) {
                if (threadIdx.x == 0 && pidx != first) bulk_store_wait_read();      // previous row has left the exchange buffer
            }
            __syncthreads();          // landing buffer consumed; also: everyone is past the previous pair's last exchange read
            if (pidx + 1 < last) {
                const int2 prn = __ldg(p.pairs + pidx + 1);
                if (threadIdx.x == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&mbar, ROW_BYTES);
                    bulk_load_1d(land, p.spectra + ((long long)prn.y << p.logL) + (row << LOGN), ROW_BYTES, &mbar);
                }
                // the next pair starts a new X_i row: its loads go out now -- the registers are free until the next pair
                // product -- and arrive during this pair's transform instead of stalling the next pair
                if (p.xi_early && prn.x != cur_i) {
                    const float2* __restrict__ xi = p.spectra + ((long long)prn.x << p.logL) + (row << LOGN);
#pragma unroll
                    for (int u = 0; u < E; ++u) a[u] = RMX_X_LOAD(xi + i0 + u * NT);
                    cur_i = prn.x;
                }
            }
        } else {
            const float2* __restrict__ xj = p.spectra + ((long long)pr.y << p.logL) + (row << LOGN);
#if RMX_DBG_PAIR_NOLOAD
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = cmul_conj(make_float2((float)(pidx + u), (float)(i0 - u)), a[u]);
            (void)xj;
#else
            if constexpr (XI_SMEM) {
#pragma unroll
                for (int u = 0; u < E; ++u) r[u] = RMX_X_LOAD(xj + i0 + u * NT);
#pragma unroll
                for (int u = 0; u < E; ++u) r[u] = cmul_conj(r[u], land[i0 + u * NT]);
            } else {
#pragma unroll
                for (int u = 0; u < E; ++u) r[u] = cmul_conj(RMX_X_LOAD(xj + i0 + u * NT), a[u]);     // X_j * conj(X_i)
            }
#endif
#if !RMX_PAIR_SPLIT
            if (pidx != first) {
                if (RMX_PAIR_RUN_BULK_STORE && threadIdx.x == 0) bulk_store_wait_read();   // previous row has left the buffer
                __syncthreads();
            }
#endif
        }
#if RMX_PAIR_SPLIT
        fft_tile_split<GEO, true>(r, smem, g, i0, p.tabs, sb);      // the FREE barrier protects the exchange area across pairs
#else
        fft_tile<GEO, true, (RMX_PAIR_TWTREE != 0)>(r, smem, g, i0, p.tabs);
#endif
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
        float2* __restrict__ out = p.dst + (long long)pidx * p.src_item_stride + (row << LOGN);
        if constexpr (STAGED == 2) {
#if RMX_PAIR_SPLIT
            if (sb.free_pending) {                            // every thread is past its last exchange read
                mbar_wait(&sb.bar[1], sb.par_free);
                sb.par_free ^= 1u;
                sb.free_pending = false;
            }
#endif
#pragma unroll
            for (int u = 0; u < E; ++u) smem[i0 + u * NT] = r[u];
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) bulk_store_1d(out, smem, ROW_BYTES);
        } else if constexpr (STAGED == 1) {
            if (pidx != first) {
                if (threadIdx.x == 0) bulk_store_wait_read();        // the previous row has left the staging buffer (long ago)
                __syncthreads();
            }
#pragma unroll
            for (int u = 0; u < E; ++u) land[i0 + u * NT] = r[u];
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) bulk_store_1d(out, land, ROW_BYTES);
        } else if constexpr (RMX_PAIR_RUN_BULK_STORE && !PREFETCH) {
            // row -> exchange buffer -> one bulk copy (TMA engine); it drains while the next pair loads
            __syncthreads();                                 // every thread is past its last exchange read
#pragma unroll
            for (int u = 0; u < E; ++u) smem[i0 + u * NT] = r[u];
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) bulk_store_1d(out, smem, (uint32_t)(GEO::N * sizeof(float2)));
        } else {
#if RMX_DBG_PAIR_NOSTORE
            float chk = 0.f;
#pragma unroll
            for (int u = 0; u < E; ++u) chk += r[u].x * r[u].y;
            if (chk == 1.2345e-30f) out[i0] = r[0];                  // keeps the arithmetic alive, never true in practice
#else
#pragma unroll
            for (int u = 0; u < E; ++u) RMX_D_STORE(out + i0 + u * NT, r[u]);
#endif
        }
    }
    if ((STAGED != 0 || (RMX_PAIR_RUN_BULK_STORE && !PREFETCH)) && threadIdx.x == 0) bulk_store_wait_read();
}

// ---------------------------------------------------------------------------------------
// Welch PSD, one kernel: a thread-block cluster holds a whole segment on chip
// ---------------------------------------------------------------------------------------
// nperseg = C * n (C = 2, 4 or 8 CTAs per cluster; n = 8192 points per CTA).  The two passes of
// the forward transform (column FFTs of length C at stride n, twiddle, row FFTs of length n) run
// back to back inside the cluster: CTA c computes the length-C column transforms of its n/C
// columns straight from cu8 (unpack, Hann window), multiplies by w_L^{j*k} and scatters output
// row k to CTA k through distributed shared memory; after a cluster barrier every CTA owns one
// complete row, transforms it and accumulates |X|^2 in registers across the segments its cluster
// walks.  The 8*nperseg-byte spectrum never goes to HBM (the two-pass path writes and re-reads it).
// Layout of `accum`: position k*n + m holds bin k + C*m (digit-transposed, passes {C, n}).
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_count_x() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrive without release semantics: orders nothing but the barrier itself (no MEMBAR); for "I am done READING" hand-offs
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f2(uint32_t local_addr, unsigned rank, float2 v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(v.x), "f"(v.y) : "memory");
}

struct WelchClusterParams {
    const uint8_t* cu8;       // [n_segments][2*nperseg]
    const float* window;      // [nperseg]
    float* accum;             // [nperseg], zeroed by the caller
    StageTables tabs;         // stage tables of the n-point row transform (32 values per thread)
    int n_segments;
    int logL;                 // log2(nperseg)
};

template <int LOGC>
__global__ void __launch_bounds__(kThreads, 2) k_welch_cluster(const WelchClusterParams p) {
    using GEO = TileGeom<13, 5, false>;
    constexpr int E = GEO::E, NT = GEO::NT, N = GEO::N, C = 1 << LOGC;
    constexpr int CPT = E / C;                               // columns per thread in the column phase
    extern __shared__ float2 smem[];
    // window values this CTA needs (its N/C columns of all C rows), behind the row / exchange buffer: they are the
    // same for every segment, so they are read from global memory once instead of once per segment
    float* s_win = reinterpret_cast<float*>(smem + GEO::NP);
    const unsigned c = cluster_ctarank();
    const int t = threadIdx.x, i0 = threadIdx.x, g = 0;
    const uint32_t lmask = (1u << p.logL) - 1u;
    const uint32_t smem_base = smem_u32(smem);
#pragma unroll
    for (int v = 0; v < CPT; ++v)
#pragma unroll
        for (int r = 0; r < C; ++r)
            s_win[(v * C + r) * kThreads + t] = __ldg(p.window + (uint32_t)r * N + c * (uint32_t)(N / C) + (uint32_t)t + 256u * v);
    // every CTA of the cluster has started before anyone stores into a peer's shared memory
    cluster_arrive_relaxed();
    cluster_wait();

    float acc[E];
#pragma unroll
    for (int u = 0; u < E; ++u) acc[u] = 0.f;

    bool first = true;
    for (int seg = (int)cluster_id_x(); seg < p.n_segments; seg += (int)cluster_count_x()) {
        const uint8_t* __restrict__ in = p.cu8 + ((size_t)seg << (p.logL + 1));
        // ---- column phase: CPT columns x C rows per thread --------------------------------------
        // (columns t + 256*v: 8-byte lanes of a warp are contiguous in the destination row, which is what
        // the distributed-shared-memory stores want; 16-byte stores of adjacent columns measured 1.45x slower)
        float2 y[CPT][C];
#pragma unroll
        for (int v = 0; v < CPT; ++v) {
            const uint32_t j = c * (uint32_t)(N / C) + (uint32_t)t + 256u * v;
#pragma unroll
            for (int r = 0; r < C; ++r) {
                const uint32_t n = (uint32_t)r * N + j;
                const uchar2 b = *reinterpret_cast<const uchar2*>(in + 2 * (size_t)n);
                const float w = s_win[(v * C + r) * kThreads + t];
                y[v][r] = make_float2(((float)b.x - 127.5f) * w, ((float)b.y - 127.5f) * w);
            }
            dft_regs<C, false>(y[v]);
            // twiddle w_L^{-j*k}, k = 1..C-1: exact base root, powers by a product tree
            float2 pw[C];
            pw[1] = unit_root(j & lmask, p.logL, false);
            static_for<2, C>([&](auto K_) {
                constexpr int k = decltype(K_)::value;
                pw[k] = cmul(pw[k / 2], pw[k - k / 2]);
            });
            static_for<1, C>([&](auto K_) {
                constexpr int k = decltype(K_)::value;
                y[v][k] = cmul(y[v][k], pw[k]);
            });
        }
        if (!first) cluster_wait();                          // every CTA of the cluster is done with its row buffer
        first = false;
#pragma unroll
        for (int v = 0; v < CPT; ++v) {
            const uint32_t j = c * (uint32_t)(N / C) + (uint32_t)t + 256u * v;
#pragma unroll
            for (int k = 0; k < C; ++k) st_cluster_f2(smem_base + j * (uint32_t)sizeof(float2), (unsigned)k, y[v][k]);
        }
        cluster_arrive();
        cluster_wait();                                      // all rows are complete
        // ---- row phase: this CTA's row k = c ------------------------------------------------------
        float2 r[E];
#pragma unroll
        for (int u = 0; u < E; ++u) r[u] = smem[i0 + u * NT];
        __syncthreads();                                     // dense row consumed; buffer becomes the exchange area
        fft_tile<GEO, false>(r, smem, g, i0, p.tabs);
#pragma unroll
        for (int u = 0; u < E; ++u) acc[u] += cnorm2(r[u]);
        // my READS of the buffer are done (matched at the top): nothing this CTA wrote has to become visible to
        // the others here, so the arrive carries no release fence (saves a MEMBAR.ALL.GPU + ERRBAR per segment)
        cluster_arrive_relaxed();
    }
    if (!first) cluster_wait();
    float* __restrict__ out = p.accum + (size_t)c * N;
#pragma unroll
    for (int u = 0; u < E; ++u) atomicAdd(out + i0 + u * NT, acc[u]);
}

// ---------------------------------------------------------------------------------------
// outermost inverse pass + arg-max, persistent and TMA-fed
// ---------------------------------------------------------------------------------------
// Same arithmetic as k_col<..., K_INV_ARGMAX[_PRE]>, but each CTA loops over tiles and the n x G
// tile (n rows of G*8 contiguous bytes, row stride s*8 bytes) is brought into shared memory by
// 2-D TMA box loads instead of per-thread strided LDGs.  The load of tile t+1 is issued as soon
// as the last exchange read of tile t is done, so it overlaps the final radix stage, the |c|^2
// arg-max and the start of the next iteration -- and the LSU queue carries no bulk global loads
// that would stall the other resident CTA's shared-memory exchanges.
// tmap: FLOAT32 tensor {2*s, n_items*n}, box {2*G, min(n, 256)}, no swizzle.
template <int LOGN, int LOGE, bool PRE>
__global__ void __launch_bounds__(kThreads, argmax_tma_ctas(LOGE)) k_col_argmax_tma(const PassParams p, const __grid_constant__ CUtensorMap tmap,
                                                               const unsigned n_tiles) {
    using GEO = TileGeom<LOGN, LOGE, true>;
    constexpr int E = GEO::E, NT = GEO::NT, G = GEO::G, LOGG = GEO::LOGG, N = GEO::N;
    static_assert(GEO::NSTAGES >= 2, "needs an exchange buffer");
    constexpr int BOX_ROWS = N < 256 ? N : 256;
    constexpr uint32_t TILE_BYTES = (uint32_t)GEO::TILE * sizeof(float2);
    extern __shared__ float2 smem_raw[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ float s_am_v[2][kThreads / 32];
    __shared__ unsigned s_am_rank[2];
    int am_par = 0;
    bool am_have = false;
    float am_prev_val = -1.f;
    long long am_prev_slot = 0;
    if (threadIdx.x == 0) { s_am_rank[0] = 0xffffffffu; s_am_rank[1] = 0xffffffffu; }
    // 128-byte aligned TMA destination, kept in the shared address space (LDS/STS, not generic LD/ST)
    float2* smem = smem_raw + (((128u - (smem_u32(smem_raw) & 127u)) & 127u) >> 3);

    int g, i0;
    GEO::thread_map(threadIdx.x, g, i0);
    const int logS = p.logS;
    const int logM = LOGN + logS;                          // == logL for the outermost pass
    const int log_tpb = logS - LOGG;                       // tiles per item (log2)
    const uint32_t lmask = (1u << p.logL) - 1u;

    auto issue = [&](unsigned idx) {                       // one thread
        const unsigned jt = idx & ((1u << log_tpb) - 1u);
        const unsigned item = idx >> log_tpb;
        mbar_expect_tx(&mbar, TILE_BYTES);
#pragma unroll
        for (int c = 0; c < N / BOX_ROWS; ++c)
            tma_load_2d(smem + c * BOX_ROWS * G, &tmap, (int)(jt << (LOGG + 1)), (int)(item << LOGN) + c * BOX_ROWS, &mbar);
    };

    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned idx = blockIdx.x;
    if (threadIdx.x == 0 && idx < n_tiles) issue(idx);
    uint32_t parity = 0;

    for (; idx < n_tiles; idx += gridDim.x) {
        const unsigned jt = idx & ((1u << log_tpb) - 1u);
        const long long item = idx >> log_tpb;
        const unsigned j = (jt << LOGG) + g;               // column
        const long long base = j;                          // element offset inside the item (row 0)

        mbar_wait(&mbar, parity);
        parity ^= 1u;
        float2 r[E];
#pragma unroll
        for (int u = 0; u < E; ++u) r[u] = smem[((i0 + u * NT) << LOGG) + g];
        const unsigned next = idx + gridDim.x;
        __syncthreads();                                   // dense tile consumed; buffer becomes the exchange area

        if constexpr (!PRE) {
            float2 tw[E];
            row_twiddles<E>(tw, j, (uint32_t)i0, (uint32_t)NT, logM, true, p.scale);
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
        }

        fft_tile<GEO, true, false>(r, smem, g, i0, p.tabs, [&]() {      // no twiddle tree: 80 registers (measured +22 % with it)
            // every generic-proxy access to the buffer is ordered before the async-proxy refill
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0 && next < n_tiles) {
                fence_proxy_async();
                issue(next);
            }
        });

        // rank = (m + lag_neg_max) mod L orders the lags like scipy's 'full' output (see k_col)
        const uint32_t span = (uint32_t)p.lag_pos_max + (uint32_t)p.lag_neg_max;
        const uint32_t rank0 = ((uint32_t)base + ((uint32_t)i0 << logS) + (uint32_t)p.lag_neg_max) & lmask;
        const uint32_t rstep = (uint32_t)NT << logS;
        float v[E];
        float bv = -1.f;
        const uint32_t col_rank0 = (j + (uint32_t)p.lag_neg_max) & lmask;
        const uint32_t excl = lmask - span;
        const uint32_t smask = (1u << logS) - 1u;
        const uint32_t first_excl = span + 1u + ((col_rank0 - (span + 1u)) & smask);
        const bool all_valid = excl == 0u || first_excl > lmask || first_excl < span + 1u;
        if (all_valid) {
#pragma unroll
            for (int u = 0; u < E; ++u) { v[u] = cnorm2(r[u]); bv = fmaxf(bv, v[u]); }
        } else {
#pragma unroll
            for (int u = 0; u < E; ++u) {
                const uint32_t rank = (rank0 + (uint32_t)u * rstep) & lmask;
                v[u] = rank <= span ? cnorm2(r[u]) : -1.f;
                bv = fmaxf(bv, v[u]);
            }
        }
        // Block arg-max with ONE barrier per tile: the per-warp maxima and the winning rank live in buffers
        // indexed by tile parity, and thread 0 writes the partial of tile t after the barrier of tile t+1 (which
        // every thread passes only after its atomicMin of tile t).
        {
            float m = bv;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
            if ((threadIdx.x & 31) == 0) s_am_v[am_par][threadIdx.x >> 5] = m;
            __syncthreads();
            if (threadIdx.x == 0) {
                if (am_have) {                              // partial of the previous tile
                    Partial out;
                    out.val = am_prev_val;
                    out.rank = s_am_rank[am_par ^ 1];
                    p.partials[am_prev_slot] = out;
                }
                s_am_rank[am_par ^ 1] = 0xffffffffu;        // free for the tile after this one
            }
            float bm = s_am_v[am_par][0];
#pragma unroll
            for (int w = 1; w < kThreads / 32; ++w) bm = fmaxf(bm, s_am_v[am_par][w]);
            if (bv == bm && bm >= 0.f) {
                uint32_t best = 0xffffffffu;
#pragma unroll
                for (int u = 0; u < E; ++u)
                    if (v[u] == bm) best = min(best, (rank0 + (uint32_t)u * rstep) & lmask);
                atomicMin(&s_am_rank[am_par], best);
            }
            am_prev_val = bm;
            am_prev_slot = (item << log_tpb) + jt;
            am_have = true;
            am_par ^= 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && am_have) {
        Partial out;
        out.val = am_prev_val;
        out.rank = s_am_rank[am_par ^ 1];
        p.partials[am_prev_slot] = out;
    }
}

}  // namespace rmx
#include "rmx_pair_pp.cuh"
