// rmx_fused_outer.cuh — three-pass plans (L > 2^23, BASELINE config 5): the two OUTER inverse passes in one kernel.
//
// A three-pass inverse transform runs rows (n2 = 4096) -> middle columns (n1) -> outer columns (n0) + arg-max.
// As separate launches the middle pass reads and re-writes the whole correlation workspace (2 x 8L bytes per pair) at
// the HBM roofline just to be read once more by the outer pass.  Here both passes run inside one persistent kernel on
// SLABS of the workspace that fit L2: a slab is G1 adjacent columns (G1 = 4096/n1) of one pair = n0*n1*G1 points
// (4 MB at cfg5).  Per slab the kernel first transforms the n0 middle tiles (n1 rows x G1 columns each, read from the
// row-pass output) and writes them into a ring of scratch slabs laid out [m1][k0][c] -- small enough to stay in L2 --
// and then the n1*G1/G0 outer tiles, each a CONTIGUOUS 32 KB block of that scratch slab (n0 rows x G0 entries,
// G0/G1 values of m1 x G1 columns), followed by the |c|^2 arg-max.  The middle-pass output therefore never goes to
// HBM: the two passes move 8L bytes per pair instead of 24L.
//
// Scheduling: CTAs pull work items from one global counter.  Item order is  M(s) tiles, then F(s - LAG) tiles, for
// slab s = 0, 1, ... so every dependency points to a LOWER item index: an F tile waits (acquire-spin on a per-slot
// counter) until all M tiles of its slab have published their writes (threadfence + atomic), an M tile waits until
// the F tiles of the slab that used its ring slot NS slabs earlier are done.  The lowest unfinished item never
// waits on anything unfinished and every pulled item belongs to a running CTA, so the scheme cannot deadlock for
// any grid size; no cooperative launch is needed.
#pragma once
#include "rmx_kernels.cuh"

namespace rmx {

constexpr int kFusedBatch = 2;      // tiles per work item

struct FusedOuterParams {
    const float2* src;        // row-pass output (pre-twiddled for the middle pass), [pair][L]
    float2* scratch;          // ring of `ring_slots` slabs, n0*n1*G1 float2 each
    Partial* partials;        // [pair][tiles_per_pair]
    unsigned* counters;       // [0] work queue, [1 .. 1+ring_slots) M-done, [1+ring_slots .. 1+2*ring_slots) F-done
    StageTables tabs1;        // middle pass (n1)
    StageTables tabs0;        // outer pass (n0)
    int n_pairs;
    int logL;
    int logn2;                // log2 of the row length
    int lag_pos_max, lag_neg_max;
    int ring_slots;           // NS
    int f_lag;                // F(s - f_lag) follows M(s) in the queue (1 <= f_lag < ring_slots - 1)
    float scale;              // 1/L, folded into the outer twiddles
};

// L2 cache policies: the row-pass output is read exactly once (evict first), the scratch ring is written and re-read
// within a few hundred microseconds and must not be pushed out by that stream (evict last)
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float2 ld_nc_hint(const float2* p, unsigned long long pol) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ld_cg_hint(const float2* p, unsigned long long pol) {
    float2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st_hint(float2* p, float2 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int LOGN1, int LOGN0>
__global__ void __launch_bounds__(kThreads, 3) k_outer_fused(const FusedOuterParams p) {
    using GEO1 = TileGeom<LOGN1, 4, true>;      // middle tile: n1 rows x G1 columns
    using GEO0 = TileGeom<LOGN0, 4, true>;      // outer tile:  n0 rows x G0 entries
    static_assert(LOGN0 <= LOGN1, "outer transform must not be longer than the middle one");
    constexpr int E = 16, N1 = GEO1::N, N0 = GEO0::N, G1 = GEO1::G, G0 = GEO0::G, LOGG1 = GEO1::LOGG;
    constexpr int NT1 = GEO1::NT, NT0 = GEO0::NT;
    constexpr int M_TILES = N0;                               // middle tiles per slab (one per k0)
    constexpr int F_TILES = (N1 * G1) / G0;                   // outer tiles per slab
    constexpr int M1SUB = G0 / G1;                            // values of m1 covered by one outer tile
    constexpr long long SLAB = (long long)N0 * N1 * G1;       // points per slab
    // A work item is a BATCH of BT consecutive tiles of one phase of one slab: one queue atomic, one dependency
    // check and one publication per batch (each is a global round trip with the whole CTA waiting on thread 0).
    constexpr int BT = kFusedBatch;
    static_assert(M_TILES % BT == 0 && F_TILES % BT == 0, "batch must divide the tile counts");
    constexpr unsigned MB = M_TILES / BT, FB = F_TILES / BT;
    extern __shared__ float2 smem[];
    __shared__ unsigned s_item[2];

    const int logn2 = p.logn2;
    const int logM1 = LOGN1 + logn2;                          // middle block length
    const unsigned spp = 1u << (logn2 - LOGG1);               // slabs per pair
    const unsigned n_slabs = (unsigned)p.n_pairs * spp;
    const unsigned per_round = MB + FB;
    const unsigned long long n_items = (unsigned long long)(n_slabs + (unsigned)p.f_lag) * per_round;
    unsigned* done_m = p.counters + 1;
    unsigned* done_f = p.counters + 1 + p.ring_slots;
    const uint32_t lmask = (1u << p.logL) - 1u;
    const uint32_t span = (uint32_t)p.lag_pos_max + (uint32_t)p.lag_neg_max;
    const unsigned long long pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();

    // the index of the NEXT item is fetched while the current one is processed (double-buffered slot)
    if (threadIdx.x == 0) s_item[0] = atomicAdd(p.counters, 1u);
    int par = 0;
    for (;;) {
        __syncthreads();                                      // s_item[par] is written; the exchange area is free again
        const unsigned q = s_item[par];
        if ((unsigned long long)q >= n_items) break;
        if (threadIdx.x == 0) s_item[par ^ 1] = atomicAdd(p.counters, 1u);      // latency hidden behind this item
        par ^= 1;
        const unsigned round = q / per_round, within = q % per_round;
        if (within < MB) {
            // ---------------- BT middle tiles (k0 = b*BT ...) of slab `round` ----------------
            const unsigned s = round;
            if (s >= n_slabs) continue;
            const unsigned pair = s / spp, cb = s % spp;
            const unsigned slot = s % (unsigned)p.ring_slots;
            int g, i0;
            GEO1::thread_map(threadIdx.x, g, i0);
            // ring slot free?  (the outer tiles of slab s - ring_slots have finished reading it) -- the check is
            // issued before the data loads so that its round trip overlaps them
            if (threadIdx.x == 0) {
                const unsigned need = (unsigned)F_TILES * (s / (unsigned)p.ring_slots);
                while (ld_acquire_u32(done_f + slot) < need) __nanosleep(100);
            }
            for (int t = 0; t < BT; ++t) {
                const unsigned k0 = within * BT + t;
                const float2* __restrict__ in = p.src + ((long long)pair << p.logL) + ((long long)k0 << logM1) +
                                                ((long long)cb << LOGG1) + g + ((long long)i0 << logn2);
                const long long rstride = (long long)NT1 << logn2;
                float2 r[E];
#pragma unroll
                for (int u = 0; u < E; ++u) r[u] = ld_nc_hint(in + u * rstride, pol_stream);
                if (t > 0) __syncthreads();                   // exchange area free (previous tile's last gather is done)
                fft_tile<GEO1, true, false>(r, smem, g, i0, p.tabs1);
                if (t == 0) __syncthreads();                  // thread 0 is past the slot check
                float2* __restrict__ out = p.scratch + (long long)slot * SLAB + (long long)k0 * G1 + g;
#pragma unroll
                for (int u = 0; u < E; ++u) st_hint(out + (long long)(i0 + u * NT1) * (N0 * G1), r[u], pol_keep);      // [m1][k0][c]
            }
            __syncthreads();                                  // every thread has issued its stores
            if (threadIdx.x == 0) {
                __threadfence();                              // cumulative: publishes the whole CTA's stores (ordered by the barrier)
                atomicAdd(done_m + slot, (unsigned)BT);
            }
        } else {
            // ---------------- BT outer tiles of slab `round - f_lag`, arg-max ----------------
            if (round < (unsigned)p.f_lag) continue;
            const unsigned s = round - (unsigned)p.f_lag;
            if (s >= n_slabs) continue;
            const unsigned pair = s / spp, cb = s % spp;
            const unsigned slot = s % (unsigned)p.ring_slots;
            int g, i0;
            GEO0::thread_map(threadIdx.x, g, i0);
            if (threadIdx.x == 0) {
                const unsigned need = (unsigned)M_TILES * (s / (unsigned)p.ring_slots + 1u);
                while (ld_acquire_u32(done_m + slot) < need) __nanosleep(100);
            }
            __syncthreads();
            for (int t = 0; t < BT; ++t) {
                const unsigned f = (within - MB) * BT + t;
                const unsigned m1 = f * M1SUB + ((unsigned)g >> LOGG1);
                const unsigned c = (unsigned)g & (G1 - 1);
                const float2* __restrict__ in = p.scratch + (long long)slot * SLAB + ((long long)m1 * N0 + i0) * G1 + c;
                float2 r[E];
#pragma unroll
                for (int u = 0; u < E; ++u) r[u] = ld_cg_hint(in + (long long)u * NT0 * G1, pol_keep);   // written by other SMs: L2
                // outer-pass column index j' = m1*n2 + column, twiddles w_L^{k0*j'} (times 1/L) on the input
                const uint32_t jp = (m1 << logn2) + (cb << LOGG1) + c;
                {
                    float2 tw[E];
                    row_twiddles<E>(tw, jp, (uint32_t)i0, (uint32_t)NT0, p.logL, true, p.scale);
#pragma unroll
                    for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
                }
                fft_tile<GEO0, true, false>(r, smem, g, i0, p.tabs0);
                // lag index m = m0*M1 + j'; rank orders the lags like scipy's 'full' output (see k_col)
                const uint32_t rank0 = (jp + ((uint32_t)i0 << logM1) + (uint32_t)p.lag_neg_max) & lmask;
                const uint32_t rstep = (uint32_t)NT0 << logM1;
                float v[E];
                float bv = -1.f;
#pragma unroll
                for (int u = 0; u < E; ++u) {
                    const uint32_t rank = (rank0 + (uint32_t)u * rstep) & lmask;
                    v[u] = rank <= span ? cnorm2(r[u]) : -1.f;
                    bv = fmaxf(bv, v[u]);
                }
                uint32_t brank;
                block_argmax(bv, brank, [&](float target) {   // (its barriers also free the exchange area for the next tile)
                    uint32_t best = 0xffffffffu;
#pragma unroll
                    for (int u = 0; u < E; ++u)
                        if (v[u] == target) best = min(best, (rank0 + (uint32_t)u * rstep) & lmask);
                    return best;
                });
                if (threadIdx.x == 0) {
                    Partial o;
                    o.val = bv;
                    o.rank = brank;
                    p.partials[(long long)pair * ((long long)spp * F_TILES) + (long long)cb * F_TILES + f] = o;
                }
            }
            // every scratch load of the batch has been consumed (block_argmax ends with a barrier): the slot may be reused
            if (threadIdx.x == 0) atomicAdd(done_f + slot, (unsigned)BT);
        }
    }
}

}  // namespace rmx
