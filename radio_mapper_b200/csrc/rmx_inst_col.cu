// Instantiations of the column-pass kernels.
#include "rmx_dispatch.h"

namespace rmx {

template <int LOGN, int LOGE, int MODE>
static KernelEntry col_entry() {
    using GEO = TileGeom<LOGN, LOGE, true>;
    return KernelEntry{(PassKernel)k_col<LOGN, LOGE, MODE>, GEO::SMEM_BYTES > size_t(GEO::TILE) * 2 ? GEO::SMEM_BYTES : size_t(GEO::TILE) * 2, GEO::LOGG};
}

template <int LOGE, int MODE>
static KernelEntry col_by_logn(int logn) {
    switch (logn - LOGE) {
        case 0: return col_entry<LOGE + 0, LOGE, MODE>();
        case 1: return col_entry<LOGE + 1, LOGE, MODE>();
        case 2: return col_entry<LOGE + 2, LOGE, MODE>();
        case 3: return col_entry<LOGE + 3, LOGE, MODE>();
        case 4: return col_entry<LOGE + 4, LOGE, MODE>();
        case 5: return col_entry<LOGE + 5, LOGE, MODE>();
        default: return KernelEntry{nullptr, 0, 0};
    }
}

KernelEntry get_col_kernel32(int logn, int mode);   // rmx_inst_col32.cu

KernelEntry get_col_kernel(int logn, int loge, int mode) {
    if (loge == 5) return get_col_kernel32(logn, mode);
    if (loge != 4) return KernelEntry{nullptr, 0, 0};
    switch (mode) {
        case K_FWD_CU8: return col_by_logn<4, K_FWD_CU8>(logn);
        case K_FWD: return col_by_logn<4, K_FWD>(logn);
        case K_INV: return col_by_logn<4, K_INV>(logn);
        case K_INV_ARGMAX: return col_by_logn<4, K_INV_ARGMAX>(logn);
        case K_INV_PRE: return col_by_logn<4, K_INV_PRE>(logn);
        case K_INV_ARGMAX_PRE: return col_by_logn<4, K_INV_ARGMAX_PRE>(logn);
        default: return KernelEntry{nullptr, 0, 0};
    }
}

}  // namespace rmx
