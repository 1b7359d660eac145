// Instantiations of the contiguous-pass kernels with 8 elements per thread (n = 2048: RMX_PLAN_ROW_E8).
#include "rmx_dispatch.h"

namespace rmx {

template <int MODE>
static KernelEntry contig8_entry() {
    using GEO = TileGeom<11, 3, false>;
    return KernelEntry{(PassKernel)k_contig<11, 3, MODE>, GEO::SMEM_BYTES, GEO::LOGG};
}

KernelEntry get_contig_kernel8(int logn, int mode) {
    if (logn != 11) return KernelEntry{nullptr, 0, 0};
    switch (mode) {
        case C_FWD: return contig8_entry<C_FWD>();
        case C_INV_PAIR: return contig8_entry<C_INV_PAIR>();
        default: return KernelEntry{nullptr, 0, 0};
    }
}

template <int RUN, bool PREFETCH, int CTAS>
static PairRunEntry pair_run8_entry() {
    using GEO = TileGeom<11, 3, false>;
    const size_t exchange = size_t((GEO::NP + 15) & ~15) * sizeof(float2);
    return PairRunEntry{(PassKernel)k_contig_pair_run<11, 3, RUN, PREFETCH, 0, CTAS>,
                        PREFETCH ? exchange + size_t(GEO::N) * sizeof(float2) : GEO::SMEM_BYTES, RUN};
}

template <int RUN, bool PREFETCH>
static PairRunEntry pair_run8_by_ctas(int ctas) {
    switch (ctas) {
        case 4: return pair_run8_entry<RUN, PREFETCH, 4>();
        case 6: return pair_run8_entry<RUN, PREFETCH, 6>();
        default: return pair_run8_entry<RUN, PREFETCH, 5>();
    }
}

PairRunEntry get_pair_run_kernel8(int run, int mem, int ctas) {
    if (run >= 16) return mem == 1 ? pair_run8_by_ctas<16, true>(ctas) : pair_run8_by_ctas<16, false>(ctas);
    return mem == 1 ? pair_run8_by_ctas<8, true>(ctas) : pair_run8_by_ctas<8, false>(ctas);
}

}  // namespace rmx
