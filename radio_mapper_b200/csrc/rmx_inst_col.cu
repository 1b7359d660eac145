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

// TMA-fed arg-max pass with 16 values per thread: n = 64, 128, 256 (two-stage transforms whose
// G = 4096/n columns make 512 / 256 / 128-byte rows)
template <int LOGN, bool PRE>
static TmaKernelEntry tma_entry16() {
    using GEO = TileGeom<LOGN, 4, true>;
    return TmaKernelEntry{(ArgmaxTmaKernel)k_col_argmax_tma<LOGN, 4, PRE>, GEO::SMEM_BYTES + 128, GEO::LOGG,
                          GEO::N < 256 ? GEO::N : 256, argmax_tma_ctas(4)};
}

TmaKernelEntry get_argmax_tma_kernel16(int logn, bool pre) {
    // pre-twiddled input only (two-pass plans): with its own twiddle array the kernel no longer fits the
    // 64 registers of 4 CTAs per SM and measured slower than the strided-load kernel (cfg5, n = 128)
    if (!pre) return TmaKernelEntry{nullptr, 0, 0, 0, 0};
    switch (logn) {
        case 6: return tma_entry16<6, true>();
        case 7: return tma_entry16<7, true>();
        case 8: return tma_entry16<8, true>();
        default: return TmaKernelEntry{nullptr, 0, 0, 0, 0};
    }
}

template <int LOGN1, int LOGN0>
static FusedOuterEntry fused_entry() {
    using GEO1 = TileGeom<LOGN1, 4, true>;
    using GEO0 = TileGeom<LOGN0, 4, true>;
    return FusedOuterEntry{(FusedOuterKernel)k_outer_fused<LOGN1, LOGN0>,
                           GEO1::SMEM_BYTES > GEO0::SMEM_BYTES ? GEO1::SMEM_BYTES : GEO0::SMEM_BYTES, GEO1::LOGG, GEO0::LOGG};
}

FusedOuterEntry get_fused_outer_kernel(int logn1, int logn0) {
    if (logn1 == 8 && logn0 == 7) return fused_entry<8, 7>();      // L = 2^27 (cfg5)
    if (logn1 == 7 && logn0 == 7) return fused_entry<7, 7>();      // L = 2^26
    if (logn1 == 7 && logn0 == 6) return fused_entry<7, 6>();      // L = 2^25
    if (logn1 == 6 && logn0 == 6) return fused_entry<6, 6>();      // L = 2^24
    if (logn1 == 8 && logn0 == 8) return fused_entry<8, 8>();      // L = 2^28
    return FusedOuterEntry{nullptr, 0, 0, 0};
}

}  // namespace rmx
