// rmx_dispatch.h — lookup of the template-instantiated pass kernels (internal).
#pragma once
#include "rmx_kernels.cuh"
#include "rmx_fused_outer.cuh"

namespace rmx {

typedef void (*PassKernel)(const PassParams);

struct KernelEntry {
    PassKernel fn;      // nullptr if this (logn, loge, mode) is not instantiated
    size_t smem_bytes;  // dynamic shared memory
    int logG;           // log2(FFTs per tile)
};

typedef void (*ArgmaxTmaKernel)(const PassParams, const CUtensorMap, const unsigned);
struct TmaKernelEntry {
    ArgmaxTmaKernel fn;  // nullptr if not instantiated
    size_t smem_bytes;
    int logG;
    int box_rows;
    int ctas_per_sm;     // resident CTAs the kernel was compiled for (grid = this * SM count)
};
// persistent TMA-fed arg-max pass (32 values per thread; n = 512 or 1024)
TmaKernelEntry get_argmax_tma_kernel(int logn, int loge, bool pre_twiddled);

// persistent TMA-fed forward pass 0 from cu8 (n = 512, 1024 with 32 values per thread; n = 128, 256 with 16)
TmaKernelEntry get_fwd_tma_kernel(int logn, int loge);

// three-pass plans: middle + outer inverse pass + arg-max fused through an L2-resident scratch ring
typedef void (*FusedOuterKernel)(const FusedOuterParams);
struct FusedOuterEntry {
    FusedOuterKernel fn;     // nullptr if (logn1, logn0) is not instantiated
    size_t smem_bytes;
    int logG1, logG0;
};
FusedOuterEntry get_fused_outer_kernel(int logn1, int logn0);

// X_i-stationary pair pass (one row per tile, 16 values per thread: n = 4096); run = pairs per CTA
struct PairRunEntry {
    PassKernel fn;       // nullptr if not instantiated
    size_t smem_bytes;
    int run;
};
// (ctas: resident CTAs per SM the 2048-point variant is compiled for: 4, 5 or 6)
PairRunEntry get_pair_run_kernel(int logn, int loge, int run, int mem, int ctas = 0);
// the same pass with `groups` (2 or 3) warp groups per CTA that hand the FP32 pipe round (rmx_pair_pp.cuh);
// launch with groups * 256 threads, smem_bytes covers all groups
PairRunEntry get_pair_run_pp_kernel(int logn, int loge, int run, int groups);

typedef void (*WelchClusterKernel)(const WelchClusterParams);
struct WelchClusterEntry {
    WelchClusterKernel fn;   // nullptr unless log2(cluster size) is 1, 2 or 3
    size_t smem_bytes;
    int cluster;
};
// one-kernel Welch PSD for nperseg = cluster * 8192
WelchClusterEntry get_welch_cluster_kernel(int logc);

KernelEntry get_contig_kernel(int logn, int loge, int mode);
KernelEntry get_col_kernel(int logn, int loge, int mode);

// supported ranges for register tile size 2^loge
inline int max_contig_logn(int loge) { return kLogThreads + loge; }
inline int max_col_logn(int loge) { return kLogThreads + loge - 3; }   // >= 8 columns per tile
inline int min_logn(int loge) { return loge; }

}  // namespace rmx
