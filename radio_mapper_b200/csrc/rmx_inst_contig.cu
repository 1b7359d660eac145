// Instantiations of the contiguous-pass kernels.
#include "rmx_dispatch.h"

namespace rmx {

template <int LOGN, int LOGE, int MODE>
static KernelEntry contig_entry() {
    using GEO = TileGeom<LOGN, LOGE, false>;
    return KernelEntry{(PassKernel)k_contig<LOGN, LOGE, MODE>, GEO::SMEM_BYTES, GEO::LOGG};
}

template <int LOGE, int MODE>
static KernelEntry contig_by_logn(int logn) {
    switch (logn - LOGE) {
        case 0: return contig_entry<LOGE + 0, LOGE, MODE>();
        case 1: return contig_entry<LOGE + 1, LOGE, MODE>();
        case 2: return contig_entry<LOGE + 2, LOGE, MODE>();
        case 3: return contig_entry<LOGE + 3, LOGE, MODE>();
        case 4: return contig_entry<LOGE + 4, LOGE, MODE>();
        case 5: return contig_entry<LOGE + 5, LOGE, MODE>();
        case 6: return contig_entry<LOGE + 6, LOGE, MODE>();
        case 7: return contig_entry<LOGE + 7, LOGE, MODE>();
        case 8: return contig_entry<LOGE + 8, LOGE, MODE>();
        default: return KernelEntry{nullptr, 0, 0};
    }
}

KernelEntry get_contig_kernel32(int logn, int mode);   // rmx_inst_contig32.cu

KernelEntry get_contig_kernel8(int logn, int mode);    // rmx_inst_contig8.cu
PairRunEntry get_pair_run_kernel8(int run, int mem, int ctas);

KernelEntry get_contig_kernel(int logn, int loge, int mode) {
    if (loge == 5) return get_contig_kernel32(logn, mode);
    if (loge == 3) return get_contig_kernel8(logn, mode);
    if (loge != 4) return KernelEntry{nullptr, 0, 0};
    switch (mode) {
        case C_FWD: return contig_by_logn<4, C_FWD>(logn);
        case C_FWD_CU8: return contig_by_logn<4, C_FWD_CU8>(logn);
        case C_INV_PAIR: return contig_by_logn<4, C_INV_PAIR>(logn);
        case C_FWD_PSD: return contig_by_logn<4, C_FWD_PSD>(logn);
        case C_INV_PAIR_WIN2: return logn == 12 ? contig_entry<12, 4, C_INV_PAIR_WIN2>() : KernelEntry{nullptr, 0, 0};
        case C_INV_PAIR_WIN4: return logn == 12 ? contig_entry<12, 4, C_INV_PAIR_WIN4>() : KernelEntry{nullptr, 0, 0};
        case C_INV_PAIR_WIN8: return logn == 12 ? contig_entry<12, 4, C_INV_PAIR_WIN8>() : KernelEntry{nullptr, 0, 0};
        default: return KernelEntry{nullptr, 0, 0};
    }
}

template <int RUN, bool PREFETCH, int STAGED>
static PairRunEntry pair_run_entry() {
    using GEO = TileGeom<12, 4, false>;
    const size_t exchange = size_t((GEO::NP + 15) & ~15) * sizeof(float2);
    return PairRunEntry{(PassKernel)k_contig_pair_run<12, 4, RUN, PREFETCH, STAGED>,
                        (PREFETCH || STAGED) ? exchange + size_t(GEO::N) * sizeof(float2) : GEO::SMEM_BYTES, RUN};
}

template <int RUN>
static PairRunEntry pair_run_xi_smem_entry() {
    using GEO = TileGeom<12, 4, false>;
    const size_t exchange = size_t((GEO::NP + 15) & ~15) * sizeof(float2);
    return PairRunEntry{(PassKernel)k_contig_pair_run<12, 4, RUN, false, 0, RMX_PAIR_RUN_CTAS, true>, exchange + size_t(GEO::N) * sizeof(float2), RUN};
}

// run: pairs walked by one CTA (8 or 16); mem: 4 = X_i row in shared memory instead of registers (no prefetch); 0 = per-thread loads and stores, 1 = next X_j row prefetched by a bulk
// copy into shared memory, 2 = finished row staged in shared memory and stored by a bulk copy, 3 = prefetch + the
// finished row staged in the exchange buffer and stored by a bulk copy
PairRunEntry get_pair_run_kernel(int logn, int loge, int run, int mem, int ctas) {
    if (logn == 11 && loge == 3) return get_pair_run_kernel8(run, mem == 1 || mem == 3 ? 1 : 0, ctas);
    if (logn == 12 && loge == 4) {
        if (mem == 4) return run >= 16 ? pair_run_xi_smem_entry<16>() : pair_run_xi_smem_entry<8>();
        if (run >= 16) return mem == 1 ? pair_run_entry<16, true, 0>() : mem == 2 ? pair_run_entry<16, false, 1>() : mem == 3 ? pair_run_entry<16, true, 2>() : pair_run_entry<16, false, 0>();
        return mem == 1 ? pair_run_entry<8, true, 0>() : mem == 2 ? pair_run_entry<8, false, 1>() : mem == 3 ? pair_run_entry<8, true, 2>() : pair_run_entry<8, false, 0>();
    }
    return PairRunEntry{nullptr, 0, 0};
}

template <int RUN, int NG>
static PairRunEntry pair_run_pp_entry() {
    using GEO = TileGeom<12, 4, false>;
    const size_t group = (size_t((GEO::NP + 15) & ~15) + size_t(GEO::N)) * sizeof(float2);
    return PairRunEntry{(PassKernel)k_contig_pair_run_pp<12, 4, RUN, NG>, NG * group, RUN};
}

PairRunEntry get_pair_run_pp_kernel(int logn, int loge, int run, int groups) {
    if (logn == 12 && loge == 4) {
        if (groups == 2) return run >= 16 ? pair_run_pp_entry<16, 2>() : pair_run_pp_entry<8, 2>();
        if (groups == 3) return run >= 16 ? pair_run_pp_entry<16, 3>() : pair_run_pp_entry<8, 3>();
    }
    return PairRunEntry{nullptr, 0, 0};
}

}  // namespace rmx
