// Instantiations of the contiguous-pass kernels with 32 elements per thread (n = 8192).
#include "rmx_dispatch.h"

namespace rmx {

template <int MODE>
static KernelEntry contig32_entry() {
    using GEO = TileGeom<13, 5, false>;
    return KernelEntry{(PassKernel)k_contig<13, 5, MODE>, GEO::SMEM_BYTES, GEO::LOGG};
}

KernelEntry get_contig_kernel32(int logn, int mode) {
    if (logn != 13) return KernelEntry{nullptr, 0, 0};
    switch (mode) {
        case C_FWD: return contig32_entry<C_FWD>();
        case C_FWD_CU8: return contig32_entry<C_FWD_CU8>();
        case C_INV_PAIR: return contig32_entry<C_INV_PAIR>();
        case C_FWD_PSD: return contig32_entry<C_FWD_PSD>();
        case C_INV_PAIR_WIN2: return contig32_entry<C_INV_PAIR_WIN2>();
        case C_INV_PAIR_WIN4: return contig32_entry<C_INV_PAIR_WIN4>();
        case C_INV_PAIR_WIN8: return contig32_entry<C_INV_PAIR_WIN8>();
        default: return KernelEntry{nullptr, 0, 0};
    }
}

WelchClusterEntry get_welch_cluster_kernel(int logc) {
    using GEO = TileGeom<13, 5, false>;
    switch (logc) {
        // + the CTA's window slice: E values per thread
        case 1: return WelchClusterEntry{(WelchClusterKernel)k_welch_cluster<1>, GEO::SMEM_BYTES + GEO::TILE * sizeof(float), 2};
        case 2: return WelchClusterEntry{(WelchClusterKernel)k_welch_cluster<2>, GEO::SMEM_BYTES + GEO::TILE * sizeof(float), 4};
        case 3: return WelchClusterEntry{(WelchClusterKernel)k_welch_cluster<3>, GEO::SMEM_BYTES + GEO::TILE * sizeof(float), 8};
        default: return WelchClusterEntry{nullptr, 0, 0};
    }
}

}  // namespace rmx
