// Instantiations of the column-pass kernels with 32 elements per thread (n = 32, 512, 1024).
#include "rmx_dispatch.h"

namespace rmx {

template <int LOGN, int MODE>
static KernelEntry col32_entry() {
    using GEO = TileGeom<LOGN, 5, true>;
    return KernelEntry{(PassKernel)k_col<LOGN, 5, MODE>, GEO::SMEM_BYTES > size_t(GEO::TILE) * 2 ? GEO::SMEM_BYTES : size_t(GEO::TILE) * 2, GEO::LOGG};
}

template <int MODE>
static KernelEntry col32_by_logn(int logn) {
    switch (logn) {
        case 5: return col32_entry<5, MODE>();
        case 9: return col32_entry<9, MODE>();
        case 10: return col32_entry<10, MODE>();
        default: return KernelEntry{nullptr, 0, 0};
    }
}

KernelEntry get_col_kernel32(int logn, int mode) {
    switch (mode) {
        case K_FWD_CU8: return col32_by_logn<K_FWD_CU8>(logn);
        case K_FWD: return col32_by_logn<K_FWD>(logn);
        case K_INV: return col32_by_logn<K_INV>(logn);
        case K_INV_ARGMAX: return col32_by_logn<K_INV_ARGMAX>(logn);
        case K_INV_PRE: return col32_by_logn<K_INV_PRE>(logn);
        case K_INV_ARGMAX_PRE: return col32_by_logn<K_INV_ARGMAX_PRE>(logn);
        default: return KernelEntry{nullptr, 0, 0};
    }
}

template <int LOGN, bool PRE>
static TmaKernelEntry tma_entry() {
    using GEO = TileGeom<LOGN, 5, true>;
    return TmaKernelEntry{(ArgmaxTmaKernel)k_col_argmax_tma<LOGN, 5, PRE>, GEO::SMEM_BYTES + 128, GEO::LOGG,
                          GEO::N < 256 ? GEO::N : 256, argmax_tma_ctas(5)};
}

template <int LOGN, int LOGE>
static TmaKernelEntry fwd_tma_entry() {
    using GEO = TileGeom<LOGN, LOGE, true>;
    const size_t stage = size_t(GEO::N) * 2 * GEO::G;
    return TmaKernelEntry{(ArgmaxTmaKernel)k_col_fwd_cu8_tma<LOGN, LOGE>, ((GEO::SMEM_BYTES + 127) & ~size_t(127)) + 128 + 2 * stage,
                          GEO::LOGG, GEO::N < 256 ? GEO::N : 256, min_ctas(LOGE)};
}

TmaKernelEntry get_fwd_tma_kernel(int logn, int loge) {
    if (loge == 5 && logn == 10) return fwd_tma_entry<10, 5>();
    if (loge == 5 && logn == 9) return fwd_tma_entry<9, 5>();
    if (loge == 4 && logn == 8) return fwd_tma_entry<8, 4>();
    if (loge == 4 && logn == 7) return fwd_tma_entry<7, 4>();
    return TmaKernelEntry{nullptr, 0, 0, 0, 0};
}

TmaKernelEntry get_argmax_tma_kernel16(int logn, bool pre);   // rmx_inst_col.cu

TmaKernelEntry get_argmax_tma_kernel(int logn, int loge, bool pre) {
    if (loge == 5 && logn == 10) return pre ? tma_entry<10, true>() : tma_entry<10, false>();
    if (loge == 5 && logn == 9) return pre ? tma_entry<9, true>() : tma_entry<9, false>();
    if (loge == 4) return get_argmax_tma_kernel16(logn, pre);
    return TmaKernelEntry{nullptr, 0, 0, 0, 0};
}

}  // namespace rmx
