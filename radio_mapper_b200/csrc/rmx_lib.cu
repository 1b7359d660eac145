// rmx_lib.cu — plan management, pass scheduling, small kernels and the C ABI (include/rmx.h).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/rmx.h"
#include "rmx_dispatch.h"

using namespace rmx;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(RMX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK(name)                                                                  \
    do {                                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                \
        if (e_ != cudaSuccess)                                                              \
            return fail(RMX_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char* rmx_last_error(void) { return g_err; }
extern "C" int rmx_version(void) { return RMX_VERSION; }

// ---------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------
struct ProfRecord {
    const char* name;
    cudaEvent_t start, stop;
};

struct rmx_plan {
    bool prof_enabled = false;
    std::vector<ProfRecord> prof_records;      // one per launch since rmx_profile_enable
    std::vector<cudaEvent_t> prof_pool;        // recycled events
    int n_signals = 0;
    long long n_samples = 0;
    int logL = 0;
    int n_passes = 0;
    int loge[kMaxStages + 2] = {0};   // pass t register-tile size (elements per thread, log2)
    int logn[kMaxStages + 2] = {0};   // pass t transform length
    int logs[kMaxStages + 2] = {0};   // pass t column stride (0 for the contiguous pass)
    StageTables tabs[kMaxStages + 2];
    std::vector<void*> dev_allocs;
    float* d_hann = nullptr;
    double hann_sumsq = 0.0;
    long long lag_pos_max = 0, lag_neg_max = 0;
    bool force_full_search = false;   // true: never take the one-pass windowed path
    StageTables welch_row_tabs;       // 8192-point row transform of the cluster Welch kernel (built on first use)
    bool welch_row_tabs_ready = false;
    unsigned flags = 0;               // RMX_PLAN_* given to rmx_plan_create
    // tuning knobs (rmx_plan_set_option)
    int pair_run = 8;                 // pairs walked by one CTA of the X_i-stationary row pass (8 or 16)
    int pair_prefetch = 1;            // next X_j row by bulk copy into shared memory
    int pair_ctas = 5;                // resident CTAs per SM the 2048-point row pass (RMX_PLAN_ROW_E8) is compiled for: 4 | 5 | 6
    int pair_groups = 0;              // 2 | 3: warp groups per CTA handing the FP32 pipe round (rmx_pair_pp.cuh); 0 = independent CTAs
    int pair_xi_early = 1;            // a new X_i row is loaded one pair ahead of its first use (prefetch path)
    int pair_xi_smem = 0;             // stationary X_i row in shared memory instead of registers (takes the prefetch buffer's place)
    int pair_store = 0;               // finished rows through a staging buffer + bulk copy (takes the prefetch buffer's place)
    long long fwd_group_bytes = 0;    // forward passes run over groups of signals whose spectra fit this many bytes (0 = all at once)
    int welch_clusters = 0;           // resident clusters of the Welch kernel (0 = occupancy query)
    int fwd_tma = 1;                  // forward pass 0 through the persistent TMA-fed kernel (measured -10..-15 % on that pass)
    int fuse_outer = 0;               // three-pass plans: middle + outer inverse pass fused through an L2-resident scratch ring
                                      // (off: measured 17.0 ms against 16.0 ms for the separate launches at cfg5 -- the scratch
                                      // stays in L2 as intended, but both passes are FMA/L1-bound at the same cost per tile)
};

static int build_stage_tables(rmx_plan* pl, int logn, int loge, StageTables* out) {
    memset(out, 0, sizeof(*out));
    const int nstages = (logn + loge - 1) / loge;
    for (int s = 1; s < nstages; ++s) {
        const int logp = s * loge;
        const int logr = std::min(loge, logn - logp);
        const int P = 1 << logp, R = 1 << logr;
        std::vector<float2> h((size_t)(R - 1) * P);
        const double m = (double)P * R;
        for (int q = 1; q < R; ++q)
            for (int k = 0; k < P; ++k) {
                // exact angle reduction in integers
                const long long e = ((long long)q * k) % (long long)(P * R);
                const double a = -2.0 * M_PI * (double)e / m;
                h[(size_t)(q - 1) * P + k] = make_float2((float)cos(a), (float)sin(a));
            }
        float2* d = nullptr;
        CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(float2)));
        pl->dev_allocs.push_back(d);
        CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
        out->tw[s] = d;
    }
    return RMX_OK;
}

// register-tile size (log2 elements per thread) for a pass of length 2^logn: fewest Stockham
// stages first, then the smaller register footprint
static int choose_col_loge(int logn) {
    const bool ok16 = logn >= 4 && logn <= max_col_logn(4);
    const bool ok32 = logn >= 5 && logn <= max_col_logn(5) && get_col_kernel(logn, 5, K_FWD).fn != nullptr;
    if (ok16 && ok32) return ((logn + 4) / 5 < (logn + 3) / 4) ? 5 : 4;
    if (ok16) return 4;
    if (ok32) return 5;
    return 0;
}
static int choose_contig_loge(int logn, unsigned flags = 0) {
    if ((flags & RMX_PLAN_ROW_E8) && logn == max_contig_logn(3)) return 3;
    if (logn >= 4 && logn <= max_contig_logn(4)) return 4;
    if (logn == max_contig_logn(5)) return 5;
    return 0;
}

static int choose_passes(rmx_plan* pl) {
    const int logL = pl->logL;
    int maxc = max_contig_logn(5);
    const int maxk = max_col_logn(5), minn = 4;
    // Two-pass plans whose column pass still fits one tile with 4096-point rows use those: the 16-values-
    // per-thread row kernel keeps 3 CTAs per SM and measured ~12 % faster per element than the 8192-point
    // one (B200, L = 2^22), and the column pass of length 512/1024 takes the TMA-fed kernel.
    if (logL - 12 >= 5 && logL - 12 <= maxk) maxc = 12;
    if (logL - 13 > maxk) maxc = 12;        // three passes: the 4096-point row kernels are ~15 % faster per element (cfg5)
    { const int e = (int)((pl->flags >> 8) & 0x1fu); if (e >= 8 && e <= max_contig_logn(5)) maxc = e; }   // RMX_PLAN_ROW_LOGN: developer override
    if ((pl->flags & RMX_PLAN_ROW_E8) && logL - max_contig_logn(3) >= 5 && logL - max_contig_logn(3) <= maxk) maxc = max_contig_logn(3);
    if (logL < minn) return fail(RMX_ERR_UNSUPPORTED, "fft_len 2^%d is below the minimum 2^%d", logL, minn);
    if (logL <= maxc) {
        pl->n_passes = 1;
        pl->logn[0] = logL;
        pl->logs[0] = 0;
        pl->loge[0] = choose_contig_loge(logL);
        return RMX_OK;
    }
    // (Three-pass plans keep short outer passes: a 1024-point outermost pass at a 1 MB row stride touches
    // 1024 distinct 2 MB pages per tile and measured 1.75x slower than 128 rows -- TLB reach is 256 MB.)
    const int contig = std::min(maxc, logL - minn);
    int rem = logL - contig;
    const int ncol = (rem + maxk - 1) / maxk;
    if (ncol + 1 > kMaxStages + 1) return fail(RMX_ERR_UNSUPPORTED, "fft_len 2^%d needs too many passes", logL);
    int stride = logL;
    for (int t = 0; t < ncol; ++t) {
        // smaller transforms first: the outermost pass then has the most columns per tile
        const int a = rem / (ncol - t);
        pl->logn[t] = a;
        pl->loge[t] = choose_col_loge(a);
        stride -= a;
        pl->logs[t] = stride;
        rem -= a;
        if (pl->loge[t] == 0) return fail(RMX_ERR_UNSUPPORTED, "no pass split for fft_len 2^%d", logL);
    }
    pl->logn[ncol] = contig;
    pl->loge[ncol] = choose_contig_loge(contig, pl->flags);
    pl->logs[ncol] = 0;
    pl->n_passes = ncol + 1;
    return RMX_OK;
}

extern "C" int rmx_plan_create(rmx_plan** out, int n_signals, size_t n_samples, size_t fft_len, unsigned flags) {
    if (!out) return fail(RMX_ERR_ARG, "plan pointer is null");
    *out = nullptr;
    if (n_signals <= 0) return fail(RMX_ERR_ARG, "n_signals must be positive (got %d)", n_signals);
    if (n_samples == 0 || n_samples > fft_len) return fail(RMX_ERR_ARG, "need 0 < n_samples <= fft_len");
    if (fft_len & (fft_len - 1)) return fail(RMX_ERR_UNSUPPORTED, "fft_len must be a power of two (got %zu)", fft_len);
    if (fft_len > (size_t(1) << 30)) return fail(RMX_ERR_UNSUPPORTED, "fft_len above 2^30 is not supported");
    rmx_plan* pl = new rmx_plan();
    pl->flags = flags;
    pl->n_signals = n_signals;
    pl->n_samples = (long long)n_samples;
    pl->logL = 0;
    while ((size_t(1) << pl->logL) < fft_len) ++pl->logL;
    int rc = choose_passes(pl);
    for (int t = 0; rc == RMX_OK && t < pl->n_passes; ++t) rc = build_stage_tables(pl, pl->logn[t], pl->loge[t], &pl->tabs[t]);
    if (rc != RMX_OK) {
        rmx_plan_destroy(pl);
        return rc;
    }
    rmx_plan_set_max_lag(pl, -1);
    *out = pl;
    return RMX_OK;
}

extern "C" int rmx_plan_destroy(rmx_plan* pl) {
    if (!pl) return RMX_OK;
    // launches that still read the plan's twiddle tables / window may be queued on any stream
    if (!pl->dev_allocs.empty()) cudaDeviceSynchronize();
    for (void* p : pl->dev_allocs) cudaFree(p);
    for (auto& r : pl->prof_records) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    for (auto e : pl->prof_pool) cudaEventDestroy(e);
    delete pl;
    return RMX_OK;
}

extern "C" int rmx_plan_layout(const rmx_plan* pl, int32_t* pass_lengths, int cap) {
    if (!pl) return fail(RMX_ERR_ARG, "plan is null");
    for (int t = 0; t < pl->n_passes && t < cap; ++t) pass_lengths[t] = 1 << pl->logn[t];
    return pl->n_passes;
}

extern "C" int rmx_plan_set_max_lag(rmx_plan* pl, long long max_lag) {
    if (!pl) return fail(RMX_ERR_ARG, "plan is null");
    const long long L = 1LL << pl->logL;
    // lags representable without aliasing: |lag| <= min(N-1, L-N)
    long long full = std::min(pl->n_samples - 1, L - pl->n_samples);
    if (pl->n_samples * 2 - 1 <= L) full = pl->n_samples - 1;
    long long m = (max_lag < 0) ? full : std::min(max_lag, full);
    pl->lag_pos_max = m;
    pl->lag_neg_max = m;
    return RMX_OK;
}

static int tiles_per_item_pass0(const rmx_plan* pl) {
    // arg-max partials produced per pair by the outermost inverse pass
    if (pl->n_passes == 1) return 1;
    const int logG = kLogThreads + pl->loge[0] - pl->logn[0];
    return 1 << (pl->logs[0] - logG);
}

// One-pass windowed search: usable when the plan has >= 2 passes, the contiguous pass transforms
// one row per tile, and the lag window fits the first/last WU*NT outputs of a row.
struct WindowMode {
    int mode = -1;        // C_INV_PAIR_WIN2 / C_INV_PAIR_WIN8, or -1
    int slots = 0;        // 2*WU*NT complex partial sums per (pair, chunk)
    int rows_per_cta = 0;
    int n_chunks = 0;
};

static WindowMode window_mode(const rmx_plan* pl, int n_pairs) {
    WindowMode w;
    if (pl->force_full_search || pl->n_passes < 2 || pl->lag_pos_max != pl->lag_neg_max) return w;
    const int last = pl->n_passes - 1;
    if (pl->logn[last] != kLogThreads + pl->loge[last]) return w;
    const int nt = 1 << kLogThreads;                     // threads per row == rows-step NT
    const long long m = pl->lag_pos_max;
    int mode;
    if (m < 2LL * nt) mode = C_INV_PAIR_WIN2;
    else if (m < 4LL * nt) mode = C_INV_PAIR_WIN4;
    else if (m < 8LL * nt) mode = C_INV_PAIR_WIN8;
    else return w;
    if (!get_contig_kernel(pl->logn[last], pl->loge[last], mode).fn) return w;
    const int wu = window_wu(mode);
    if (2LL * wu * nt * 2 > (1LL << pl->logn[last])) return w;   // window wider than half a row: nothing to prune
    const long long rows = 1LL << (pl->logL - pl->logn[last]);
    w.mode = mode;
    w.slots = 2 * wu * nt;
    // rows per CTA: as many as keep >= ~8 CTAs per SM in flight (few pairs -> short chunks), at most 32
    long long rpc = 32;
    while (rpc > 1 && (rows / rpc) * (long long)std::max(1, n_pairs) < 148LL * 8) rpc >>= 1;
    w.rows_per_cta = (int)std::min<long long>(rpc, std::max<long long>(1, rows));
    w.n_chunks = (int)(rows / w.rows_per_cta);
    return w;
}

// three-pass plans whose two outer passes run fused (rmx_fused_outer.cuh): scratch ring + counters behind the partials
constexpr int kFusedRingSlots = 10, kFusedLag = 4;   // ~3.5 slabs in flight with 444 CTAs x 2 tiles: lag past them, ring past the lag
static FusedOuterEntry fused_outer_entry(const rmx_plan* pl) {
    if (pl->n_passes != 3 || !pl->fuse_outer || pl->loge[0] != 4 || pl->loge[1] != 4) return FusedOuterEntry{nullptr, 0, 0, 0};
    const FusedOuterEntry k = get_fused_outer_kernel(pl->logn[1], pl->logn[0]);
    if (k.fn && pl->logn[2] < k.logG1) return FusedOuterEntry{nullptr, 0, 0, 0};
    return k;
}
static size_t fused_outer_extra_bytes(const rmx_plan* pl) {
    const FusedOuterEntry k = fused_outer_entry(pl);
    if (!k.fn) return 0;
    const size_t slab = (size_t(1) << (pl->logn[0] + pl->logn[1] + k.logG1)) * sizeof(float2);
    return (size_t)kFusedRingSlots * slab + 1024;
}

extern "C" size_t rmx_plan_workspace_bytes(const rmx_plan* pl, int n_pairs) {
    if (!pl || n_pairs <= 0) return 0;
    const WindowMode w = window_mode(pl, n_pairs);
    if (w.mode >= 0) return (size_t)n_pairs * (w.n_chunks + 1) * w.slots * sizeof(float2) + 256;
    const size_t L = size_t(1) << pl->logL;
    return (size_t)n_pairs * (L * sizeof(float2) + (size_t)tiles_per_item_pass0(pl) * sizeof(Partial)) + 1024 +
           fused_outer_extra_bytes(pl);
}

extern "C" int rmx_plan_set_option(rmx_plan* pl, const char* name, long long value) {
    if (!pl || !name) return fail(RMX_ERR_ARG, "null argument to rmx_plan_set_option");
    if (!strcmp(name, "pair_run")) { if (value != 8 && value != 16) return fail(RMX_ERR_ARG, "pair_run must be 8 or 16"); pl->pair_run = (int)value; }
    else if (!strcmp(name, "pair_prefetch")) pl->pair_prefetch = value != 0;
    else if (!strcmp(name, "pair_xi_smem")) pl->pair_xi_smem = value != 0;
    else if (!strcmp(name, "pair_xi_early")) pl->pair_xi_early = value != 0;
    else if (!strcmp(name, "pair_store")) {
        if (value < 0 || value > 2) return fail(RMX_ERR_ARG, "pair_store must be 0, 1 or 2 (got %lld)", (long long)value);
        pl->pair_store = (int)value;
    }
    else if (!strcmp(name, "pair_ctas")) {
        if (value < 4 || value > 6) return fail(RMX_ERR_ARG, "pair_ctas must be 4, 5 or 6 (got %lld)", (long long)value);
        pl->pair_ctas = (int)value;
    }
    else if (!strcmp(name, "pair_groups")) {
        if (value != 0 && value != 2 && value != 3) return fail(RMX_ERR_ARG, "pair_groups must be 0, 2 or 3 (got %lld)", (long long)value);
        pl->pair_groups = (int)value;
    }
    else if (!strcmp(name, "fwd_group_bytes")) pl->fwd_group_bytes = value < 0 ? 0 : value;
    else if (!strcmp(name, "welch_clusters")) pl->welch_clusters = (int)std::max<long long>(0, value);
    else if (!strcmp(name, "fwd_tma")) pl->fwd_tma = value != 0;
    else if (!strcmp(name, "fuse_outer")) pl->fuse_outer = value != 0;
    else return fail(RMX_ERR_ARG, "unknown plan option '%s'", name);
    return RMX_OK;
}

extern "C" int rmx_plan_set_search_mode(rmx_plan* pl, int force_full) {
    if (!pl) return fail(RMX_ERR_ARG, "plan is null");
    pl->force_full_search = force_full != 0;
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------
// per-launch profiling (CUDA events on the launching stream; off by default)
// ---------------------------------------------------------------------------------------
struct ProfScope {
    rmx_plan* pl;
    cudaStream_t st;
    cudaEvent_t stop = nullptr;
    ProfScope(const rmx_plan* plan, const char* name, cudaStream_t stream) : pl(const_cast<rmx_plan*>(plan)), st(stream) {
        if (!pl || !pl->prof_enabled) { pl = nullptr; return; }
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!pl->prof_pool.empty()) { ev[i] = pl->prof_pool.back(); pl->prof_pool.pop_back(); }
            else if (cudaEventCreate(&ev[i]) != cudaSuccess) { pl = nullptr; return; }
        }
        pl->prof_records.push_back(ProfRecord{name, ev[0], ev[1]});
        stop = ev[1];
        cudaEventRecord(ev[0], st);
    }
    ~ProfScope() { if (pl) cudaEventRecord(stop, st); }
};

extern "C" int rmx_profile_enable(rmx_plan* pl, int enable) {
    if (!pl) return fail(RMX_ERR_ARG, "plan is null");
    for (auto& r : pl->prof_records) { pl->prof_pool.push_back(r.start); pl->prof_pool.push_back(r.stop); }
    pl->prof_records.clear();
    pl->prof_enabled = enable != 0;
    return RMX_OK;
}

extern "C" int rmx_profile_collect(rmx_plan* pl, rmx_prof_entry* out, int cap) {
    if (!pl || (!out && cap > 0)) return fail(RMX_ERR_ARG, "null argument to rmx_profile_collect");
    int n = 0;
    for (auto& r : pl->prof_records) {
        CUDA_TRY(cudaEventSynchronize(r.stop));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, r.start, r.stop));
        int k = 0;
        while (k < n && strncmp(out[k].name, r.name, sizeof(out[k].name)) != 0) ++k;
        if (k == n) {
            if (n >= cap) continue;
            memset(&out[n], 0, sizeof(out[n]));
            strncpy(out[n].name, r.name, sizeof(out[n].name) - 1);
            ++n;
        }
        out[k].launches += 1;
        out[k].total_ms += ms;
    }
    return n;
}

static unsigned grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 16));
}

// ---------------------------------------------------------------------------------------
// pass launchers
// ---------------------------------------------------------------------------------------
static int launch_pass(const rmx_plan* pl, const KernelEntry& k, const char* name, dim3 grid, const PassParams& pp,
                       cudaStream_t st) {
    if (!k.fn) return fail(RMX_ERR_UNSUPPORTED, "kernel %s is not instantiated for this size", name);
    if (k.smem_bytes > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem_bytes));
    {
        ProfScope prof(pl, name, st);
        k.fn<<<grid, kThreads, k.smem_bytes, st>>>(pp);
    }
    LAUNCH_CHECK(name);
    return RMX_OK;
}

static PassParams base_params(const rmx_plan* pl) {
    PassParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.logL = pl->logL;
    pp.n_samples = pl->n_samples;
    pp.cu8_stride = 2 * pl->n_samples;
    pp.src_item_stride = 1LL << pl->logL;
    pp.lag_pos_max = (int)pl->lag_pos_max;
    pp.lag_neg_max = (int)pl->lag_neg_max;
    pp.scale = 1.0f;
    return pp;
}

// Developer switches are plan flags (rmx_plan_create, include/rmx.h), fixed for the life of a plan: one process
// can A/B the kernel variants by creating two plans (tests/test_gpu_parity.py::test_kernel_variants_agree):
//   RMX_PLAN_TWIDDLE_IN_COL  inter-pass twiddles on the input of the column pass instead of the row pass output
//   RMX_PLAN_NO_TMA          arg-max pass through per-thread strided loads instead of the TMA-fed kernel
//   RMX_PLAN_NO_PAIR_RUN     one pair per CTA in the 4096-point row pass instead of the X_i-stationary walk
//   RMX_PLAN_ROW_LOGN(n)     force the row length of multi-pass plans
static bool twiddle_in_contig(const rmx_plan* pl) { return (pl->flags & RMX_PLAN_TWIDDLE_IN_COL) == 0; }

// The contiguous inverse pass (C_INV_PAIR) also applies the input twiddles -- and, for two-pass plans,
// the 1/L -- of the column pass that runs next (pass n_passes-2); that pass is launched pre_twiddled.
static void set_post_twiddle(const rmx_plan* pl, PassParams* pp) {
    const int np = pl->n_passes;
    if (np < 2 || !twiddle_in_contig(pl)) return;
    const int t = np - 2;
    pp->post_logm = pl->logn[t] + pl->logs[t];
    pp->post_logn = pl->logn[t];
    pp->post_scale = t == 0 ? 1.0f / (float)(size_t(1) << pl->logL) : 1.0f;
}

static unsigned tiles_of(const rmx_plan* pl, int pass, long long n_items) {
    const int logtile = kLogThreads + pl->loge[pass];
    const long long total = n_items << pl->logL;
    return (unsigned)((total + (1LL << logtile) - 1) >> logtile);
}


// ---------------------------------------------------------------------------------------
// TMA descriptors (cuTensorMapEncodeTiled resolved through the runtime; libcuda is not linked)
// ---------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }();
    return fn;
}

static int sm_count() {
    static int n = []() {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    return n;
}

// outermost inverse pass + arg-max through the persistent TMA-fed kernel; returns RMX_ERR_UNSUPPORTED
// (without setting an error) when this plan / pointer cannot take that path
static int launch_argmax_tma(const rmx_plan* pl, const PassParams& pp, bool pre, int cnt, cudaStream_t st, bool* taken) {
    *taken = false;
    if (pl->flags & RMX_PLAN_NO_TMA) return RMX_OK;
    const TmaKernelEntry k = get_argmax_tma_kernel(pl->logn[0], pl->loge[0], pre);
    auto enc = tensor_map_encoder();
    if (!k.fn || !enc) return RMX_OK;
    if (pl->logn[0] + pl->logs[0] != pl->logL) return RMX_OK;            // pass 0 spans the whole item
    if ((reinterpret_cast<uintptr_t>(pp.src) & 15) != 0) return RMX_OK;
    const unsigned long long s = 1ULL << pl->logs[0], n = 1ULL << pl->logn[0];
    if (2 * s > (1ULL << 32) - 1 || (unsigned long long)cnt * n > (1ULL << 31) - 1) return RMX_OK;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {2 * s, (cuuint64_t)cnt * n};
    const cuuint64_t gstride[1] = {s * sizeof(float2)};
    const cuuint32_t box[2] = {(cuuint32_t)(2u << k.logG), (cuuint32_t)k.box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(pp.src), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RMX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    const unsigned n_tiles = tiles_of(pl, 0, cnt);
    const unsigned grid = std::min<unsigned>(n_tiles, (unsigned)k.ctas_per_sm * (unsigned)sm_count());
    CUDA_TRY(cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem_bytes));
    {
        ProfScope prof(pl, "col_inv_argmax", st);
        k.fn<<<grid, kThreads, k.smem_bytes, st>>>(pp, tmap, n_tiles);
    }
    LAUNCH_CHECK("col_inv_argmax_tma");
    *taken = true;
    return RMX_OK;
}

// forward pass 0 from cu8 through the persistent TMA-fed kernel; *taken = false when this plan / input cannot take it
static int launch_fwd_tma(const rmx_plan* pl, const PassParams& pp, int cnt, cudaStream_t st, bool* taken,
                          const char* name = "col_fwd_cu8") {
    *taken = false;
    if (!pl->fwd_tma || pp.window != nullptr || pl->n_passes < 2) return RMX_OK;
    const TmaKernelEntry k = get_fwd_tma_kernel(pl->logn[0], pl->loge[0]);
    auto enc = tensor_map_encoder();
    if (!k.fn || !enc) return RMX_OK;
    if (pl->logn[0] + pl->logs[0] != pl->logL) return RMX_OK;            // pass 0 spans the whole item
    const unsigned long long s = 1ULL << pl->logs[0];
    if ((unsigned long long)pl->n_samples % s != 0) return RMX_OK;       // whole rows only (no read past the last signal)
    const unsigned long long rows = (unsigned long long)pl->n_samples >> pl->logs[0];
    if (rows == 0 || ((reinterpret_cast<uintptr_t>(pp.cu8) | (uintptr_t)pp.cu8_stride) & 15) != 0) return RMX_OK;
    if (2 * s > (1ULL << 32) - 1 || (1u << k.logG) > 256u) return RMX_OK;
    CUtensorMap tmap;
    // one UINT16 element = one (I, Q) byte pair, so a box row of G samples stays within the 256-element box limit
    const cuuint64_t gdim[3] = {s, rows, (cuuint64_t)cnt};
    const cuuint64_t gstride[2] = {2 * s, (cuuint64_t)pp.cu8_stride};
    const cuuint32_t box[3] = {(cuuint32_t)(1u << k.logG), (cuuint32_t)k.box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint8_t*>(pp.cu8), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RMX_ERR_CUDA, "cuTensorMapEncodeTiled (forward) failed (%d)", (int)r);
    const unsigned n_tiles = tiles_of(pl, 0, cnt);
    const unsigned grid = std::min<unsigned>(n_tiles, (unsigned)k.ctas_per_sm * (unsigned)sm_count());
    CUDA_TRY(cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem_bytes));
    {
        ProfScope prof(pl, name, st);
        k.fn<<<grid, kThreads, k.smem_bytes, st>>>(pp, tmap, n_tiles);
    }
    LAUNCH_CHECK("col_fwd_cu8_tma");
    *taken = true;
    return RMX_OK;
}

// innermost inverse pass over `cnt` pairs
static int launch_pair_pass(const rmx_plan* pl, const PassParams& pp_in, int cnt, cudaStream_t st) {
    PassParams pp = pp_in;
    pp.xi_early = pl->pair_xi_early;
    const int last = pl->n_passes - 1;
    // the bulk-copy prefetch needs 16-byte aligned spectrum rows (rows are multiples of 32 KB apart)
    const bool prefetch = pl->pair_prefetch != 0 && (reinterpret_cast<uintptr_t>(pp.spectra) & 15) == 0;
    const bool dst_ok = (reinterpret_cast<uintptr_t>(pp.dst) & 15) == 0;
    const bool staged = pl->pair_store == 1 && dst_ok;                       // dedicated staging buffer, no prefetch
    const bool xstaged = pl->pair_store == 2 && dst_ok && prefetch;           // staged in the exchange buffer, with prefetch
    const PairRunEntry kr = get_pair_run_kernel(pl->logn[last], pl->loge[last], pl->pair_run, pl->pair_xi_smem ? 4 : xstaged ? 3 : staged ? 2 : prefetch ? 1 : 0, pl->pair_ctas);
    if (kr.fn && pl->n_passes >= 2 && !(pl->flags & RMX_PLAN_NO_PAIR_RUN)) {
        const long long rows = 1LL << (pl->logL - pl->logn[last]);
        const long long blocks = (cnt + kr.run - 1) / kr.run;
        const PairRunEntry kg = (pl->pair_groups && prefetch && !staged && !xstaged && !pl->pair_xi_smem)
                                    ? get_pair_run_pp_kernel(pl->logn[last], pl->loge[last], pl->pair_run, pl->pair_groups)
                                    : PairRunEntry{nullptr, 0, 0};
        if (kg.fn) {
            const long long vblocks = rows * ((cnt + kg.run - 1) / kg.run);
            CUDA_TRY(cudaFuncSetAttribute((const void*)kg.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kg.smem_bytes));
            {
                ProfScope prof(pl, "contig_inv_pair", st);
                kg.fn<<<(unsigned)((vblocks + pl->pair_groups - 1) / pl->pair_groups), kThreads * pl->pair_groups, kg.smem_bytes, st>>>(pp);
            }
            LAUNCH_CHECK("contig_inv_pair_groups");
            return RMX_OK;
        }
        return launch_pass(pl, KernelEntry{kr.fn, kr.smem_bytes, 0}, "contig_inv_pair", dim3((unsigned)(rows * blocks)), pp, st);
    }
    return launch_pass(pl, get_contig_kernel(pl->logn[last], pl->loge[last], C_INV_PAIR), "contig_inv_pair",
                       dim3(tiles_of(pl, last, cnt)), pp, st);
}

// forward FFT of n_items signals from cu8 (optionally windowed) into `spectra`
static int forward_cu8(const rmx_plan* pl, const uint8_t* iq, long long stride_bytes, float2* spectra, int n_items,
                       const float* window, cudaStream_t st) {
    const int np = pl->n_passes;
    PassParams pp = base_params(pl);
    if (stride_bytes > 0) pp.cu8_stride = stride_bytes;
    pp.n_items = n_items;
    pp.cu8 = iq;
    pp.window = window;
    pp.dst = spectra;
    pp.src = spectra;
    if (np == 1) {
        if (window) return fail(RMX_ERR_UNSUPPORTED, "windowed single-pass forward is handled by the caller");
        pp.tabs = pl->tabs[0];
        return launch_pass(pl, get_contig_kernel(pl->logn[0], pl->loge[0], C_FWD_CU8), "contig_fwd_cu8",
                           dim3(tiles_of(pl, 0, n_items)), pp, st);
    }
    // Groups of signals whose spectra fit `fwd_group_bytes` run their passes back to back, so the output of one
    // pass is still in L2 when the next pass reads it (the passes are in place: the spectrum is written to HBM once).
    const long long L = 1LL << pl->logL;
    int group = n_items;
    if (pl->fwd_group_bytes > 0)
        group = (int)std::max<long long>(1, std::min<long long>(n_items, pl->fwd_group_bytes / (L * (long long)sizeof(float2))));
    const long long stride = pp.cu8_stride;
    for (int first = 0; first < n_items; first += group) {
        const int cnt = std::min(group, n_items - first);
        pp.n_items = cnt;
        pp.cu8 = iq + (long long)first * stride;
        pp.dst = spectra + (long long)first * L;
        pp.src = pp.dst;
        for (int t = 0; t < np - 1; ++t) {
            pp.tabs = pl->tabs[t];
            pp.logS = pl->logs[t];
            if (t == 0) {
                bool taken = false;
                int rc = launch_fwd_tma(pl, pp, cnt, st, &taken);
                if (rc) return rc;
                if (taken) continue;
            }
            int rc = launch_pass(pl, get_col_kernel(pl->logn[t], pl->loge[t], t == 0 ? K_FWD_CU8 : K_FWD),
                                 t == 0 ? "col_fwd_cu8" : "col_fwd", dim3(tiles_of(pl, t, cnt)), pp, st);
            if (rc) return rc;
        }
        pp.tabs = pl->tabs[np - 1];
        int rc = launch_pass(pl, get_contig_kernel(pl->logn[np - 1], pl->loge[np - 1], C_FWD), "contig_fwd",
                             dim3(tiles_of(pl, np - 1, cnt)), pp, st);
        if (rc) return rc;
    }
    return RMX_OK;
}

extern "C" int rmx_fft_forward_cu8(const rmx_plan* pl, const uint8_t* iq, size_t signal_stride_bytes,
                                   rmx_complex64* spectra, void* stream) {
    if (!pl || !iq || !spectra) return fail(RMX_ERR_ARG, "null argument to rmx_fft_forward_cu8");
    if (signal_stride_bytes != 0 && (signal_stride_bytes < (size_t)(2 * pl->n_samples) || (signal_stride_bytes & 1)))
        return fail(RMX_ERR_ARG, "signal_stride_bytes must be even and >= 2*n_samples (got %zu)", signal_stride_bytes);
    return forward_cu8(pl, iq, (long long)signal_stride_bytes, reinterpret_cast<float2*>(spectra), pl->n_signals, nullptr,
                       (cudaStream_t)stream);
}

// forward FFT of complex64 signals (zero-padded into `spectra`, then transformed in place)
extern "C" int rmx_fft_forward_c64(const rmx_plan* pl, const rmx_complex64* x, size_t signal_stride_elems,
                                   rmx_complex64* spectra, void* stream) {
    if (!pl || !x || !spectra) return fail(RMX_ERR_ARG, "null argument to rmx_fft_forward_c64");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t L = size_t(1) << pl->logL, N = (size_t)pl->n_samples;
    if (signal_stride_elems == 0) signal_stride_elems = N;
    if (signal_stride_elems < N) return fail(RMX_ERR_ARG, "signal_stride_elems must be >= n_samples");
    CUDA_TRY(cudaMemcpy2DAsync(spectra, L * sizeof(float2), x, signal_stride_elems * sizeof(float2), N * sizeof(float2),
                               (size_t)pl->n_signals, cudaMemcpyDeviceToDevice, st));
    if (N < L)
        CUDA_TRY(cudaMemset2DAsync(reinterpret_cast<float2*>(spectra) + N, L * sizeof(float2), 0, (L - N) * sizeof(float2),
                                   (size_t)pl->n_signals, st));
    const int np = pl->n_passes;
    PassParams pp = base_params(pl);
    pp.n_items = pl->n_signals;
    pp.src = reinterpret_cast<const float2*>(spectra);
    pp.dst = reinterpret_cast<float2*>(spectra);
    for (int t = 0; t < np - 1; ++t) {
        pp.tabs = pl->tabs[t];
        pp.logS = pl->logs[t];
        int rc = launch_pass(pl, get_col_kernel(pl->logn[t], pl->loge[t], K_FWD), "col_fwd", dim3(tiles_of(pl, t, pl->n_signals)), pp, st);
        if (rc) return rc;
    }
    pp.tabs = pl->tabs[np - 1];
    return launch_pass(pl, get_contig_kernel(pl->logn[np - 1], pl->loge[np - 1], C_FWD), "contig_fwd",
                       dim3(tiles_of(pl, np - 1, pl->n_signals)), pp, st);
}

// ---------------------------------------------------------------------------------------
// lag-search finalisation
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float2 block_sum3(float2 v, float2* sh) {
    // sum of v over the block (128 threads); result broadcast
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float2 t = make_float2(0.f, 0.f);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t.x += sh[w].x; t.y += sh[w].y; }
    return t;
}

__device__ __forceinline__ float parabolic(float ym, float y0, float yp) {
    const double den = (double)ym - 2.0 * (double)y0 + (double)yp;
    if (den == 0.0) return 0.f;
    return (float)(0.5 * ((double)ym - (double)yp) / den);
}

// multi-pass plans: reduce the tile partials, then evaluate c[m-1], c[m], c[m+1] by direct
// n_0-term sums over the input of the outermost inverse pass (still in the workspace).
__global__ void __launch_bounds__(128) k_finalize_sum(const Partial* __restrict__ partials, int tiles_per_item,
                                                      const float2* __restrict__ D, int logL, int logn0, int logs0,
                                                      int lag_pos_max, int lag_neg_max, float scale, int pre_twiddled,
                                                      rmx_peak* __restrict__ out) {
    __shared__ float s_v[4];
    __shared__ uint32_t s_l[4];
    __shared__ float2 s_sum[4];
    const int item = blockIdx.x;
    float bv = -1.f;
    uint32_t brank = 0xffffffffu;
    for (int t = threadIdx.x; t < tiles_per_item; t += blockDim.x) {
        const Partial p = partials[(long long)item * tiles_per_item + t];
        if (p.val >= 0.f && better(p.val, p.rank, bv, brank)) { bv = p.val; brank = p.rank; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, brank, off);
        if (better(ov, ol, bv, brank)) { bv = ov; brank = ol; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_l[threadIdx.x >> 5] = brank; }
    __syncthreads();
    bv = s_v[0]; brank = s_l[0];
    for (int w = 1; w < 4; ++w) if (better(s_v[w], s_l[w], bv, brank)) { bv = s_v[w]; brank = s_l[w]; }
    const int blag = (int)brank - lag_neg_max;

    const long long L = 1LL << logL;
    const int n0 = 1 << logn0;
    const float2* __restrict__ Dp = D + ((long long)item << logL);
    float y[3];
    for (int d = -1; d <= 1; ++d) {
        const long long lag = (long long)blag + d;
        float2 acc = make_float2(0.f, 0.f);
        const bool in_range = (lag >= -(long long)lag_neg_max) && (lag <= (long long)lag_pos_max);
        if (in_range) {
            const unsigned long long m = (unsigned long long)(lag < 0 ? lag + L : lag);
            const unsigned long long j = m & ((1ULL << logs0) - 1ULL);
            for (int k = threadIdx.x; k < n0; k += blockDim.x) {
                const float2 v = Dp[((long long)k << logs0) + (long long)j];
                // D[k][j] * w_L^{k*m}; a pre-twiddled workspace already carries w_L^{k*j}
                const unsigned long long e = ((unsigned long long)k * (pre_twiddled ? m - j : m)) & (unsigned long long)(L - 1);
                const float2 w = unit_root((uint32_t)e, logL, true);
                const float2 t = cmul(v, w);
                acc.x += t.x; acc.y += t.y;
            }
        }
        const float2 tot = block_sum3(acc, s_sum);
        // the workspace holds the unscaled inner passes; apply the 1/L of the inverse here
        y[d + 1] = in_range ? sqrtf(tot.x * tot.x + tot.y * tot.y) * scale : -1.f;
    }
    if (threadIdx.x == 0) {
        rmx_peak r;
        r.lag = blag;
        r.peak = y[1];
        r.frac = (y[0] >= 0.f && y[2] >= 0.f) ? parabolic(y[0], y[1], y[2]) : 0.f;
        r.pad = bv;
        out[item] = r;
    }
}

// three-pass plans with fused outer passes: the middle-pass output never reaches the workspace, so c[m-1], c[m],
// c[m+1] are evaluated from the ROW-pass output R (pre-twiddled for the middle pass):
//     c[m] = 1/L * sum_{k0 < n0} sum_{k1 < n1}  w_L^{k0*m} * w_{n1}^{k1*m1} * R[(k0*n1 + k1)*n2 + col],
//     m = m0*(n1*n2) + m1*n2 + col      (n0*n1 direct terms per lag: 32768 at cfg5, three lags per pair)
__global__ void __launch_bounds__(256) k_finalize_sum2(const Partial* __restrict__ partials, int tiles_per_item,
                                                       const float2* __restrict__ R, int logL, int logn0, int logn1, int logn2,
                                                       int lag_pos_max, int lag_neg_max, float scale, rmx_peak* __restrict__ out) {
    __shared__ float s_v[8];
    __shared__ uint32_t s_l[8];
    __shared__ float2 s_sum[8];
    const int item = blockIdx.x;
    float bv = -1.f;
    uint32_t brank = 0xffffffffu;
    for (int t = threadIdx.x; t < tiles_per_item; t += blockDim.x) {
        const Partial p = partials[(long long)item * tiles_per_item + t];
        if (p.val >= 0.f && better(p.val, p.rank, bv, brank)) { bv = p.val; brank = p.rank; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, brank, off);
        if (better(ov, ol, bv, brank)) { bv = ov; brank = ol; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_l[threadIdx.x >> 5] = brank; }
    __syncthreads();
    bv = s_v[0]; brank = s_l[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (better(s_v[w], s_l[w], bv, brank)) { bv = s_v[w]; brank = s_l[w]; }
    const int blag = (int)brank - lag_neg_max;

    const long long L = 1LL << logL;
    const int n_terms = 1 << (logn0 + logn1);
    const float2* __restrict__ Rp = R + ((long long)item << logL);
    float y[3];
    for (int d = -1; d <= 1; ++d) {
        const long long lag = (long long)blag + d;
        float2 acc = make_float2(0.f, 0.f);
        const bool in_range = (lag >= -(long long)lag_neg_max) && (lag <= (long long)lag_pos_max);
        if (in_range) {
            const unsigned long long m = (unsigned long long)(lag < 0 ? lag + L : lag);
            const unsigned long long col = m & ((1ULL << logn2) - 1ULL);
            const unsigned long long m1 = (m >> logn2) & ((1ULL << logn1) - 1ULL);
            for (int t = threadIdx.x; t < n_terms; t += blockDim.x) {
                const unsigned long long k0 = (unsigned long long)t >> logn1, k1 = (unsigned long long)t & ((1ULL << logn1) - 1ULL);
                const float2 v = Rp[((long long)t << logn2) + (long long)col];            // t = k0*n1 + k1
                const float2 w0 = unit_root((uint32_t)((k0 * m) & (unsigned long long)(L - 1)), logL, true);
                const float2 w1 = unit_root((uint32_t)((k1 * m1) & ((1ULL << logn1) - 1ULL)), logn1, true);
                const float2 tt = cmul(cmul(v, w1), w0);
                acc.x += tt.x; acc.y += tt.y;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
        __syncthreads();
        float2 tot = make_float2(0.f, 0.f);
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tot.x += s_sum[w].x; tot.y += s_sum[w].y; }
        y[d + 1] = in_range ? sqrtf(tot.x * tot.x + tot.y * tot.y) * scale : -1.f;
    }
    if (threadIdx.x == 0) {
        rmx_peak r;
        r.lag = blag;
        r.peak = y[1];
        r.frac = (y[0] >= 0.f && y[2] >= 0.f) ? parabolic(y[0], y[1], y[2]) : 0.f;
        r.pad = bv;
        out[item] = r;
    }
}

// single-pass plans: the workspace already holds c in natural order
__global__ void __launch_bounds__(128) k_finalize_direct(const float2* __restrict__ C, int logL, int lag_pos_max,
                                                         int lag_neg_max, rmx_peak* __restrict__ out) {
    __shared__ float s_v[4];
    __shared__ uint32_t s_l[4];
    const int item = blockIdx.x;
    const long long L = 1LL << logL;
    const float2* __restrict__ c = C + ((long long)item << logL);
    float bv = -1.f;
    uint32_t brank = 0xffffffffu;
    const uint32_t span = (uint32_t)lag_pos_max + (uint32_t)lag_neg_max;
    for (long long m = threadIdx.x; m < L; m += blockDim.x) {
        const uint32_t rank = (uint32_t)((m + lag_neg_max) & (L - 1));
        const float v = cnorm2(c[m]);
        if (rank <= span && better(v, rank, bv, brank)) { bv = v; brank = rank; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, brank, off);
        if (better(ov, ol, bv, brank)) { bv = ov; brank = ol; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_l[threadIdx.x >> 5] = brank; }
    __syncthreads();
    if (threadIdx.x == 0) {
        bv = s_v[0]; brank = s_l[0];
        for (int w = 1; w < 4; ++w) if (better(s_v[w], s_l[w], bv, brank)) { bv = s_v[w]; brank = s_l[w]; }
        const int blag = (int)brank - lag_neg_max;
        float y[3];
        for (int d = -1; d <= 1; ++d) {
            const long long lag = (long long)blag + d;
            if (lag < -(long long)lag_neg_max || lag > (long long)lag_pos_max) { y[d + 1] = -1.f; continue; }
            const float2 v = c[lag < 0 ? lag + L : lag];
            y[d + 1] = sqrtf(v.x * v.x + v.y * v.y);
        }
        rmx_peak r;
        r.lag = blag;
        r.peak = y[1];
        r.frac = (y[0] >= 0.f && y[2] >= 0.f) ? parabolic(y[0], y[1], y[2]) : 0.f;
        r.pad = bv;
        out[item] = r;
    }
}

// windowed mode, step 1: c[pair][slot] = scale * sum over row chunks of the partial vectors
__global__ void __launch_bounds__(256) k_window_reduce(const float2* __restrict__ partials, int n_chunks, int slots, float scale,
                                                       float2* __restrict__ c) {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= slots) return;
    const float2* __restrict__ base = partials + (long long)blockIdx.y * n_chunks * slots + sidx;
    float2 acc = make_float2(0.f, 0.f);
    for (int k = 0; k < n_chunks; ++k) {
        const float2 v = base[(long long)k * slots];
        acc.x += v.x; acc.y += v.y;
    }
    c[(long long)blockIdx.y * slots + sidx] = make_float2(acc.x * scale, acc.y * scale);
}

// step 2: arg-max of |c| over |lag| <= M (slot = lag mod slots), parabolic vertex
__global__ void __launch_bounds__(256) k_finalize_window(const float2* __restrict__ c, int slots, int lag_max,
                                                         rmx_peak* __restrict__ out) {
    __shared__ float s_v[8];
    __shared__ uint32_t s_l[8];
    const int item = blockIdx.x;
    const float2* __restrict__ ci = c + (long long)item * slots;
    float bv = -1.f;
    uint32_t brank = 0xffffffffu;
    for (int lag = -lag_max + (int)threadIdx.x; lag <= lag_max; lag += blockDim.x) {
        const float m2 = cnorm2(ci[lag >= 0 ? lag : lag + slots]);
        const uint32_t rank = (uint32_t)(lag + lag_max);
        if (better(m2, rank, bv, brank)) { bv = m2; brank = rank; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, brank, off);
        if (better(ov, ol, bv, brank)) { bv = ov; brank = ol; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = bv; s_l[threadIdx.x >> 5] = brank; }
    __syncthreads();
    if (threadIdx.x == 0) {
        bv = s_v[0]; brank = s_l[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (better(s_v[w], s_l[w], bv, brank)) { bv = s_v[w]; brank = s_l[w]; }
        const int blag = (int)brank - lag_max;
        float y[3];
        for (int d = -1; d <= 1; ++d) {
            const int lag = blag + d;
            if (lag < -lag_max || lag > lag_max) { y[d + 1] = -1.f; continue; }
            const float2 v = ci[lag >= 0 ? lag : lag + slots];
            y[d + 1] = sqrtf(cnorm2(v));
        }
        rmx_peak r;
        r.lag = blag;
        r.peak = y[1];
        r.frac = (y[0] >= 0.f && y[2] >= 0.f) ? parabolic(y[0], y[1], y[2]) : 0.f;
        r.pad = bv;
        out[item] = r;
    }
}

static int xcorr_windowed(const rmx_plan* pl, const WindowMode& w, const rmx_complex64* spectra, const rmx_pair* pairs,
                          int n_pairs, rmx_peak* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const size_t per_pair = (size_t)(w.n_chunks + 1) * w.slots * sizeof(float2);
    if (workspace_bytes < per_pair)
        return fail(RMX_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one pair (%zu needed)", workspace_bytes, per_pair);
    const int chunk = (int)std::min<size_t>((size_t)n_pairs, workspace_bytes / per_pair);
    const int last = pl->n_passes - 1;
    const KernelEntry k = get_contig_kernel(pl->logn[last], pl->loge[last], w.mode);
    for (int first = 0; first < n_pairs; first += chunk) {
        const int cnt = std::min(chunk, n_pairs - first);
        PassParams pp = base_params(pl);
        pp.n_items = cnt;
        pp.spectra = reinterpret_cast<const float2*>(spectra);
        pp.pairs = reinterpret_cast<const int2*>(pairs) + first;
        pp.tabs = pl->tabs[last];
        pp.items_per_cta = w.rows_per_cta;
        pp.win_partials = reinterpret_cast<float2*>(workspace);
        pp.win_lag_max = (int)pl->lag_pos_max;
        pp.row_npass = last;
        for (int t = 0; t < last; ++t) pp.row_logn[t] = pl->logn[t];
        int rc = launch_pass(pl, k, "contig_inv_pair_window", dim3((unsigned)cnt * (unsigned)w.n_chunks), pp, st);
        if (rc) return rc;
        float2* cvec = pp.win_partials + (size_t)chunk * w.n_chunks * w.slots;      // after the partials
        {
            ProfScope prof(pl, "finalize_window", st);
            k_window_reduce<<<dim3((w.slots + 255) / 256, cnt), 256, 0, st>>>(pp.win_partials, w.n_chunks, w.slots,
                                                                             1.0f / (float)(size_t(1) << pl->logL), cvec);
            k_finalize_window<<<cnt, 256, 0, st>>>(cvec, w.slots, (int)pl->lag_pos_max, out + first);
        }
        LAUNCH_CHECK("finalize_window");
    }
    return RMX_OK;
}

extern "C" int rmx_xcorr_pairs_peak(const rmx_plan* pl, const rmx_complex64* spectra, const rmx_pair* pairs,
                                    int n_pairs, rmx_peak* out, void* workspace, size_t workspace_bytes,
                                    void* stream) {
    if (!pl || !spectra || !pairs || !out || !workspace) return fail(RMX_ERR_ARG, "null argument to rmx_xcorr_pairs_peak");
    if (n_pairs <= 0) return RMX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        const WindowMode w = window_mode(pl, n_pairs);
        if (w.mode >= 0) return xcorr_windowed(pl, w, spectra, pairs, n_pairs, out, workspace, workspace_bytes, st);
    }
    const size_t L = size_t(1) << pl->logL;
    const int tpi = tiles_per_item_pass0(pl);
    const size_t per_pair = L * sizeof(float2) + (size_t)tpi * sizeof(Partial);
    // the fused outer passes need the inter-pass twiddles on the row-pass output (the default placement)
    FusedOuterEntry fused = twiddle_in_contig(pl) ? fused_outer_entry(pl) : FusedOuterEntry{nullptr, 0, 0, 0};
    size_t extra = fused.fn ? fused_outer_extra_bytes(pl) : 0;
    if (fused.fn && workspace_bytes < per_pair + 1024 + extra) { fused.fn = nullptr; extra = 0; }    // small caller workspace: unfused passes
    if (workspace_bytes < per_pair + 1024 + extra)
        return fail(RMX_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one pair (%zu needed)", workspace_bytes,
                    per_pair + 1024 + extra);
    const int chunk = (int)std::min<size_t>((size_t)n_pairs, (workspace_bytes - 1024 - extra) / per_pair);
    float2* D = reinterpret_cast<float2*>(workspace);
    Partial* partials = reinterpret_cast<Partial*>(reinterpret_cast<char*>(workspace) + (((size_t)chunk * L * sizeof(float2) + 255) & ~size_t(255)));
    char* after_partials = reinterpret_cast<char*>(partials) + (((size_t)chunk * tpi * sizeof(Partial) + 255) & ~size_t(255));
    unsigned* fused_counters = reinterpret_cast<unsigned*>(after_partials);
    float2* fused_scratch = reinterpret_cast<float2*>(after_partials + 1024);
    const int np = pl->n_passes;

    for (int first = 0; first < n_pairs; first += chunk) {
        const int cnt = std::min(chunk, n_pairs - first);
        PassParams pp = base_params(pl);
        pp.n_items = cnt;
        pp.spectra = reinterpret_cast<const float2*>(spectra);
        pp.pairs = reinterpret_cast<const int2*>(pairs) + first;
        pp.src = D;
        pp.dst = D;
        pp.partials = partials;
        const float inv_len = 1.0f / (float)L;
        // the 1/L of the inverse transform rides on the twiddles of the outermost column pass
        // (exact: a power of two); single-pass plans scale in the contiguous kernel
        pp.scale = np == 1 ? inv_len : 1.0f;
        // innermost pass first: rows of X_j * conj(X_i)
        pp.tabs = pl->tabs[np - 1];
        set_post_twiddle(pl, &pp);
        int rc = launch_pair_pass(pl, pp, cnt, st);
        if (rc) return rc;
        if (fused.fn) {
            // three-pass plan: middle + outer pass + arg-max in one persistent kernel (rmx_fused_outer.cuh)
            FusedOuterParams fp;
            memset(&fp, 0, sizeof(fp));
            fp.src = D;
            fp.scratch = fused_scratch;
            fp.partials = partials;
            fp.counters = fused_counters;
            fp.tabs1 = pl->tabs[1];
            fp.tabs0 = pl->tabs[0];
            fp.n_pairs = cnt;
            fp.logL = pl->logL;
            fp.logn2 = pl->logn[2];
            fp.lag_pos_max = (int)pl->lag_pos_max;
            fp.lag_neg_max = (int)pl->lag_neg_max;
            fp.ring_slots = kFusedRingSlots;
            fp.f_lag = kFusedLag;
            fp.scale = inv_len;
            CUDA_TRY(cudaMemsetAsync(fused_counters, 0, (1 + 2 * kFusedRingSlots) * sizeof(unsigned), st));
            if (fused.smem_bytes > 48 * 1024)
                CUDA_TRY(cudaFuncSetAttribute((const void*)fused.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused.smem_bytes));
            {
                ProfScope prof(pl, "outer_fused_argmax", st);
                fused.fn<<<3 * sm_count(), kThreads, fused.smem_bytes, st>>>(fp);
            }
            LAUNCH_CHECK("outer_fused_argmax");
            {
                ProfScope prof(pl, "finalize_sum", st);
                k_finalize_sum2<<<cnt, 256, 0, st>>>(partials, tpi, D, pl->logL, pl->logn[0], pl->logn[1], pl->logn[2],
                                                     (int)pl->lag_pos_max, (int)pl->lag_neg_max, inv_len, out + first);
            }
            LAUNCH_CHECK("finalize_sum2");
            continue;
        }
        for (int t = np - 2; t >= 0; --t) {
            pp.tabs = pl->tabs[t];
            pp.logS = pl->logs[t];
            pp.scale = t == 0 ? inv_len : 1.0f;
            const bool pre = (t == np - 2) && twiddle_in_contig(pl);      // twiddled by the contiguous pass
            if (t == 0) {
                bool taken = false;
                rc = launch_argmax_tma(pl, pp, pre, cnt, st, &taken);
                if (rc) return rc;
                if (taken) continue;
            }
            rc = launch_pass(pl, get_col_kernel(pl->logn[t], pl->loge[t], t == 0 ? (pre ? K_INV_ARGMAX_PRE : K_INV_ARGMAX) : (pre ? K_INV_PRE : K_INV)),
                             t == 0 ? "col_inv_argmax" : "col_inv", dim3(tiles_of(pl, t, cnt)), pp, st);
            if (rc) return rc;
        }
        if (np == 1) {
            {
                ProfScope prof(pl, "finalize_direct", st);
                k_finalize_direct<<<cnt, 128, 0, st>>>(D, pl->logL, (int)pl->lag_pos_max, (int)pl->lag_neg_max, out + first);
            }
            LAUNCH_CHECK("finalize_direct");
        } else {
            {
                ProfScope prof(pl, "finalize_sum", st);
                // two-pass plans: the workspace (input of pass 0) is pre-twiddled and already scaled
                k_finalize_sum<<<cnt, 128, 0, st>>>(partials, tpi, D, pl->logL, pl->logn[0], pl->logs[0],
                                                    (int)pl->lag_pos_max, (int)pl->lag_neg_max,
                                                    (np == 2 && twiddle_in_contig(pl)) ? 1.0f : inv_len,
                                                    (np == 2 && twiddle_in_contig(pl)) ? 1 : 0, out + first);
            }
            LAUNCH_CHECK("finalize_sum");
        }
    }
    return RMX_OK;
}

// full correlation output: out[p][m] = ifft(X_j conj X_i)[m], natural order, m = lag mod L
extern "C" int rmx_xcorr_full(const rmx_plan* pl, const rmx_complex64* spectra, const rmx_pair* pairs, int n_pairs,
                              rmx_complex64* out, void* stream) {
    if (!pl || !spectra || !pairs || !out) return fail(RMX_ERR_ARG, "null argument to rmx_xcorr_full");
    if (n_pairs <= 0) return RMX_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int np = pl->n_passes;
    PassParams pp = base_params(pl);
    pp.n_items = n_pairs;
    pp.spectra = reinterpret_cast<const float2*>(spectra);
    pp.pairs = reinterpret_cast<const int2*>(pairs);
    pp.src = reinterpret_cast<const float2*>(out);
    pp.dst = reinterpret_cast<float2*>(out);
    const float inv_len = 1.0f / (float)(size_t(1) << pl->logL);
    pp.scale = np == 1 ? inv_len : 1.0f;
    pp.tabs = pl->tabs[np - 1];
    set_post_twiddle(pl, &pp);
    int rc = launch_pair_pass(pl, pp, n_pairs, st);
    for (int t = np - 2; rc == RMX_OK && t >= 0; --t) {
        pp.tabs = pl->tabs[t];
        pp.logS = pl->logs[t];
        pp.scale = t == 0 ? inv_len : 1.0f;
        rc = launch_pass(pl, get_col_kernel(pl->logn[t], pl->loge[t], (t == np - 2 && twiddle_in_contig(pl)) ? K_INV_PRE : K_INV), "col_inv", dim3(tiles_of(pl, t, n_pairs)), pp, st);
    }
    return rc;
}

// ---------------------------------------------------------------------------------------
// Bluestein helpers (arbitrary-length DFT through the power-of-two engine)
// ---------------------------------------------------------------------------------------
// chirp[n] = exp(-i*pi*n^2/N); n^2 is reduced mod 2N in integers so the angle is exact
__device__ __forceinline__ float2 chirp_at(unsigned long long n, unsigned long long N) {
    const unsigned long long e = (n * n) % (2ULL * N);
    double s, c;
    sincospi((double)e / (double)N, &s, &c);
    return make_float2((float)c, (float)(-s));
}

// a[n] = x[n]*chirp[n] (n < N, else 0);  cc[m] = chirp[min(m, Lp-m)] for |m| < N (else 0)
__global__ void __launch_bounds__(256) k_bluestein_prepare(const float2* __restrict__ x, unsigned long long N,
                                                           unsigned long long Lp, float2* __restrict__ a,
                                                           float2* __restrict__ cc) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < Lp;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        float2 av = make_float2(0.f, 0.f), cv = make_float2(0.f, 0.f);
        if (i < N) {
            const float2 w = chirp_at(i, N);
            av = cmul(x[i], w);
            cv = w;
        } else if (Lp - i < N) {
            cv = chirp_at(Lp - i, N);
        }
        a[i] = av;
        cc[i] = cv;
    }
}

// X[k] = chirp[k] * conv[k]
__global__ void __launch_bounds__(256) k_bluestein_finish(const float2* __restrict__ conv, unsigned long long N,
                                                          float2* __restrict__ X) {
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < N;
         k += (unsigned long long)gridDim.x * blockDim.x)
        X[k] = cmul(conv[k], chirp_at(k, N));
}

extern "C" int rmx_bluestein_prepare(const rmx_complex64* x, size_t n, size_t padded_len, rmx_complex64* a,
                                     rmx_complex64* chirp_circ, void* stream) {
    if (!x || !a || !chirp_circ) return fail(RMX_ERR_ARG, "null argument to rmx_bluestein_prepare");
    if (n == 0 || padded_len < 2 * n - 1) return fail(RMX_ERR_ARG, "padded_len must be >= 2n-1");
    k_bluestein_prepare<<<grid_for((long long)padded_len, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(x), n, padded_len, reinterpret_cast<float2*>(a), reinterpret_cast<float2*>(chirp_circ));
    LAUNCH_CHECK("bluestein_prepare");
    return RMX_OK;
}

extern "C" int rmx_bluestein_finish(const rmx_complex64* conv, size_t n, rmx_complex64* out, void* stream) {
    if (!conv || !out) return fail(RMX_ERR_ARG, "null argument to rmx_bluestein_finish");
    if (n == 0) return RMX_OK;
    k_bluestein_finish<<<grid_for((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(conv), n, reinterpret_cast<float2*>(out));
    LAUNCH_CHECK("bluestein_finish");
    return RMX_OK;
}

// out[k] = 20*log10(|x[k]| + 1e-12) for a natural-order complex vector, optional fftshift
__global__ void __launch_bounds__(256) k_abs_db(const float2* __restrict__ x, size_t n, float* __restrict__ out, int shift) {
    const size_t half = n / 2;      // np.fft.fftshift moves index k to (k + n//2) % n
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const float2 v = x[k];
        size_t dst = k;
        if (shift) { dst = k + half; if (dst >= n) dst -= n; }
        out[dst] = 20.0f * log10f(hypotf(v.x, v.y) + 1e-12f);
    }
}

extern "C" int rmx_abs_db(const rmx_complex64* x, size_t n, float* out_db, int shift, void* stream) {
    if (!x || !out_db) return fail(RMX_ERR_ARG, "null argument to rmx_abs_db");
    if (n == 0) return RMX_OK;
    k_abs_db<<<grid_for((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(x), n, out_db, shift);
    LAUNCH_CHECK("abs_db");
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------
// layout helpers, dB spectrum
// ---------------------------------------------------------------------------------------
struct LayoutDesc {
    int n_passes;
    int logn[kMaxStages + 2];
    int logs[kMaxStages + 2];
};

static LayoutDesc layout_of(const rmx_plan* pl) {
    LayoutDesc d;
    d.n_passes = pl->n_passes;
    for (int t = 0; t < kMaxStages + 2; ++t) { d.logn[t] = pl->logn[t]; d.logs[t] = pl->logs[t]; }
    return d;
}

// frequency bin stored at position pos of the digit-transposed layout
__device__ __forceinline__ long long layout_freq(const LayoutDesc& d, long long pos) {
    long long f = 0;
    int weight = 0;
    for (int t = 0; t < d.n_passes; ++t) {
        const long long digit = (pos >> d.logs[t]) & ((1LL << d.logn[t]) - 1);
        f |= digit << weight;
        weight += d.logn[t];
    }
    return f;
}

__global__ void k_spectrum_natural(LayoutDesc d, const float2* __restrict__ in, float2* __restrict__ out, int logL,
                                   long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pos = i & ((1LL << logL) - 1);
        out[(i - pos) + layout_freq(d, pos)] = in[i];
    }
}

__global__ void k_spectrum_db(LayoutDesc d, const float2* __restrict__ in, float* __restrict__ out, int logL,
                              long long total, int shift) {
    const long long L = 1LL << logL;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pos = i & (L - 1);
        long long f = layout_freq(d, pos);
        if (shift) f = (f + (L >> 1)) & (L - 1);
        const float2 v = in[i];
        // 20*log10(|X| + 1e-12): buoy_node.py:405
        out[(i - pos) + f] = 20.0f * log10f(hypotf(v.x, v.y) + 1e-12f);
    }
}


extern "C" int rmx_spectrum_natural(const rmx_plan* pl, const rmx_complex64* spectra, rmx_complex64* out,
                                    int n_signals, void* stream) {
    if (!pl || !spectra || !out) return fail(RMX_ERR_ARG, "null argument to rmx_spectrum_natural");
    if ((const void*)spectra == (const void*)out) return fail(RMX_ERR_ARG, "rmx_spectrum_natural cannot run in place");
    const long long total = (long long)n_signals << pl->logL;
    k_spectrum_natural<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        layout_of(pl), reinterpret_cast<const float2*>(spectra), reinterpret_cast<float2*>(out), pl->logL, total);
    LAUNCH_CHECK("spectrum_natural");
    return RMX_OK;
}

extern "C" int rmx_spectrum_db(const rmx_plan* pl, const rmx_complex64* spectra, float* out_db, int n_signals,
                               int shift, void* stream) {
    if (!pl || !spectra || !out_db) return fail(RMX_ERR_ARG, "null argument to rmx_spectrum_db");
    const long long total = (long long)n_signals << pl->logL;
    k_spectrum_db<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        layout_of(pl), reinterpret_cast<const float2*>(spectra), out_db, pl->logL, total, shift);
    LAUNCH_CHECK("spectrum_db");
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------
// stage 1 stand-alone: cu8 -> complex64
// ---------------------------------------------------------------------------------------
// Each thread converts 16 input bytes (one 128-bit load, 8 samples) and writes four float4.
__global__ void __launch_bounds__(256) k_unpack_cu8(const uint8_t* __restrict__ in, float2* __restrict__ out, size_t n) {
    const size_t nvec = n / 8;
    const uint4* __restrict__ in4 = reinterpret_cast<const uint4*>(in);
    float4* __restrict__ out4 = reinterpret_cast<float4*>(out);
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldg(in4 + v);
        const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 o;
            o.x = (float)(ws[k] & 0xffu) - 127.5f;
            o.y = (float)((ws[k] >> 8) & 0xffu) - 127.5f;
            o.z = (float)((ws[k] >> 16) & 0xffu) - 127.5f;
            o.w = (float)(ws[k] >> 24) - 127.5f;
            out4[v * 4 + k] = o;
        }
    }
    // tail (n not a multiple of 8)
    for (size_t i = nvec * 8 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_float2((float)in[2 * i] - 127.5f, (float)in[2 * i + 1] - 127.5f);
}

__global__ void __launch_bounds__(256) k_unpack_cu8_scalar(const uint8_t* __restrict__ in, float2* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_float2((float)in[2 * i] - 127.5f, (float)in[2 * i + 1] - 127.5f);
}

extern "C" int rmx_unpack_cu8(const uint8_t* in, rmx_complex64* out, size_t n_samples, void* stream) {
    if (n_samples == 0) return RMX_OK;
    if (!in || !out) return fail(RMX_ERR_ARG, "null argument to rmx_unpack_cu8");
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (aligned) {
        k_unpack_cu8<<<grid_for((long long)(n_samples + 7) / 8, 256), 256, 0, st>>>(in, reinterpret_cast<float2*>(out), n_samples);
    } else {
        k_unpack_cu8_scalar<<<grid_for((long long)n_samples, 256), 256, 0, st>>>(in, reinterpret_cast<float2*>(out), n_samples);
    }
    LAUNCH_CHECK("unpack_cu8");
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------
// Welch PSD
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_window(const uint8_t* __restrict__ in, const float* __restrict__ w,
                                                       float2* __restrict__ out, int logL, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float ww = w[i & ((1LL << logL) - 1)];
        out[i] = make_float2(((float)in[2 * i] - 127.5f) * ww, ((float)in[2 * i + 1] - 127.5f) * ww);
    }
}

__global__ void __launch_bounds__(256) k_psd_finalize(LayoutDesc d, const float* __restrict__ accum, float* __restrict__ psd,
                                                      int logL, float scale) {
    const long long L = 1LL << logL;
    for (long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x; pos < L; pos += (long long)gridDim.x * blockDim.x)
        psd[layout_freq(d, pos)] = accum[pos] * scale;
}

static int ensure_hann(rmx_plan* pl) {
    if (pl->d_hann) return RMX_OK;
    const size_t L = size_t(1) << pl->logL;
    std::vector<float> h(L);
    double ss = 0.0;
    for (size_t n = 0; n < L; ++n) {
        // periodic Hann == scipy.signal.get_window('hann', L) (sym=False), rounded to float32 as scipy.welch does
        const double w = 0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)L);
        h[n] = (float)w;
        ss += (double)h[n] * (double)h[n];
    }
    CUDA_TRY(cudaMalloc(&pl->d_hann, L * sizeof(float)));
    pl->dev_allocs.push_back(pl->d_hann);
    CUDA_TRY(cudaMemcpy(pl->d_hann, h.data(), L * sizeof(float), cudaMemcpyHostToDevice));
    pl->hann_sumsq = ss;
    return RMX_OK;
}

extern "C" size_t rmx_welch_workspace_bytes(const rmx_plan* pl, int segments_in_flight) {
    if (!pl || segments_in_flight <= 0) return 0;
    const size_t L = size_t(1) << pl->logL;
    return L * sizeof(float) + (size_t)segments_in_flight * L * sizeof(float2) + 512;
}

extern "C" int rmx_welch_path(const rmx_plan* pl, const uint8_t* iq) {
    if (!pl) return fail(RMX_ERR_ARG, "plan is null");
    const WelchClusterEntry k = get_welch_cluster_kernel(pl->logL - 13);
    return (k.fn && !(pl->flags & RMX_PLAN_NO_WELCH_CLUSTER) && (reinterpret_cast<uintptr_t>(iq) & 7) == 0) ? 1 : 0;
}

extern "C" int rmx_welch_psd(rmx_plan* pl, const uint8_t* iq, float* psd, double sample_rate, void* workspace,
                             size_t workspace_bytes, void* stream) {
    if (!pl || !iq || !psd || !workspace) return fail(RMX_ERR_ARG, "null argument to rmx_welch_psd");
    if (pl->n_samples != (1LL << pl->logL)) return fail(RMX_ERR_ARG, "Welch needs n_samples == fft_len (nperseg)");
    if (!(sample_rate > 0)) return fail(RMX_ERR_ARG, "sample_rate must be positive");
    int rc = ensure_hann(pl);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t L = size_t(1) << pl->logL;
    const size_t accum_bytes = (L * sizeof(float) + 255) & ~size_t(255);
    if (workspace_bytes < accum_bytes + L * sizeof(float2))
        return fail(RMX_ERR_WORKSPACE, "Welch workspace too small: %zu bytes", workspace_bytes);
    float* accum = reinterpret_cast<float*>(workspace);
    float2* Y = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + accum_bytes);
    {
        // one-kernel path: nperseg = C * 8192 with a cluster of C = 2, 4 or 8 CTAs holding the segment on chip
        const WelchClusterEntry k = get_welch_cluster_kernel(pl->logL - 13);
        if (k.fn && rmx_welch_path(pl, iq) == 1) {
            if (!pl->welch_row_tabs_ready) {
                rc = build_stage_tables(pl, 13, 5, &pl->welch_row_tabs);
                if (rc) return rc;
                pl->welch_row_tabs_ready = true;
            }
            CUDA_TRY(cudaMemsetAsync(accum, 0, L * sizeof(float), st));
            WelchClusterParams wp;
            wp.cu8 = iq;
            wp.window = pl->d_hann;
            wp.accum = accum;
            wp.tabs = pl->welch_row_tabs;
            wp.n_segments = pl->n_signals;
            wp.logL = pl->logL;
            CUDA_TRY(cudaFuncSetAttribute((const void*)k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem_bytes));
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)k.cluster);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = k.smem_bytes;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)k.cluster;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            // persistent clusters: exactly as many as can be resident at once (clusters are placed inside
            // one GPC, so this is fewer than CTAs-per-SM * SMs / cluster size)
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, (const void*)k.fn, &cfg) != cudaSuccess || max_clusters < 1) {
                cudaGetLastError();
                max_clusters = std::max(1, sm_count() / k.cluster);
            }
            if (pl->welch_clusters > 0) max_clusters = pl->welch_clusters;
            const int n_clusters = std::min(pl->n_signals, max_clusters);
            cfg.gridDim = dim3((unsigned)(n_clusters * k.cluster));
            {
                ProfScope prof(pl, "welch_cluster", st);
                CUDA_TRY(cudaLaunchKernelEx(&cfg, k.fn, wp));
            }
            LAUNCH_CHECK("welch_cluster");
            LayoutDesc d;
            memset(&d, 0, sizeof(d));
            d.n_passes = 2;
            d.logn[0] = pl->logL - 13; d.logs[0] = 13;
            d.logn[1] = 13; d.logs[1] = 0;
            const float scale = (float)(1.0 / (sample_rate * pl->hann_sumsq * (double)pl->n_signals));
            k_psd_finalize<<<grid_for((long long)L, 256), 256, 0, st>>>(d, accum, psd, pl->logL, scale);
            LAUNCH_CHECK("psd_finalize");
            return RMX_OK;
        }
    }
    const int group = (int)std::min<size_t>((size_t)pl->n_signals, (workspace_bytes - accum_bytes) / (L * sizeof(float2)));
    CUDA_TRY(cudaMemsetAsync(accum, 0, L * sizeof(float), st));
    const int np = pl->n_passes;
    const int logtile = kLogThreads + pl->loge[np - 1];
    for (int first = 0; first < pl->n_signals; first += group) {
        const int cnt = std::min(group, pl->n_signals - first);
        const uint8_t* in = iq + (size_t)first * 2 * L;
        if (np == 1) {
            const long long total = (long long)cnt << pl->logL;
            k_unpack_window<<<grid_for(total, 256), 256, 0, st>>>(in, pl->d_hann, Y, pl->logL, total);
            LAUNCH_CHECK("unpack_window");
        } else {
            PassParams pp = base_params(pl);
            pp.n_items = cnt;
            pp.cu8 = in;
            pp.window = pl->d_hann;
            pp.src = Y;
            pp.dst = Y;
            for (int t = 0; t < np - 1; ++t) {
                pp.tabs = pl->tabs[t];
                pp.logS = pl->logs[t];
                // (the persistent TMA-fed pass 0 measured SLOWER here: 0.228 against 0.193 ms for 1000 x 64k -- 16-point
                // columns have no exchange to overlap the prefetch with, and 16000 small CTAs hide latency better)
                rc = launch_pass(pl, get_col_kernel(pl->logn[t], pl->loge[t], t == 0 ? K_FWD_CU8 : K_FWD), "welch_col_fwd",
                                 dim3(tiles_of(pl, t, cnt)), pp, st);
                if (rc) return rc;
            }
        }
        // last pass: FFT rows and accumulate |X|^2 over the group's segments
        PassParams pp = base_params(pl);
        pp.n_items = cnt;
        pp.src = Y;
        pp.accum = accum;
        pp.tabs = pl->tabs[np - 1];
        KernelEntry k = get_contig_kernel(pl->logn[np - 1], pl->loge[np - 1], C_FWD_PSD);
        // one row per tile: the next segment's row is prefetched by a bulk copy into a landing buffer behind the
        // exchange area (needs 16-byte aligned rows)
        if (pl->pair_prefetch && pl->logn[np - 1] == kLogThreads + pl->loge[np - 1] && np > 1 &&
            (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {
            pp.prefetch = 1;
            const size_t n = size_t(1) << pl->logn[np - 1];
            k.smem_bytes = ((k.smem_bytes / sizeof(float2) + 15) & ~size_t(15)) * sizeof(float2) + n * sizeof(float2);
        }
        const unsigned tiles_per_sig = (unsigned)std::max<long long>(1, (1LL << pl->logL) >> logtile);
        if ((1LL << pl->logL) < (1LL << logtile)) return fail(RMX_ERR_UNSUPPORTED, "Welch needs nperseg >= %d", 1 << logtile);
        // enough CTAs to fill the GPU twice over; each accumulates a chunk of segments in registers
        int chunks = std::max(1, std::min(cnt, (int)((148 * 4 + tiles_per_sig - 1) / tiles_per_sig)));
        pp.items_per_cta = (cnt + chunks - 1) / chunks;
        chunks = (cnt + pp.items_per_cta - 1) / pp.items_per_cta;
        rc = launch_pass(pl, k, "contig_fwd_psd", dim3(tiles_per_sig, chunks), pp, st);
        if (rc) return rc;
    }
    // density scaling: 1 / (fs * sum(w^2)), averaged over the segments
    const float scale = (float)(1.0 / (sample_rate * pl->hann_sumsq * (double)pl->n_signals));
    k_psd_finalize<<<grid_for((long long)L, 256), 256, 0, st>>>(layout_of(pl), accum, psd, pl->logL, scale);
    LAUNCH_CHECK("psd_finalize");
    return RMX_OK;
}

__global__ void k_power_db(const float* __restrict__ in, float* __restrict__ out, size_t n, float eps) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = 10.0f * log10f(in[i] + eps);
}

extern "C" int rmx_power_db(const float* in, float* out, size_t n, float eps, void* stream) {
    if (n == 0) return RMX_OK;
    if (!in || !out) return fail(RMX_ERR_ARG, "null argument to rmx_power_db");
    k_power_db<<<grid_for((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, eps);
    LAUNCH_CHECK("power_db");
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------
// peak candidates, mean/median, stats
// ---------------------------------------------------------------------------------------
// scipy.signal._peak_finding_utils._local_maxima_1d, one thread per rising edge.
__global__ void __launch_bounds__(256) k_threshold_peaks(const float* __restrict__ x, int n, float height,
                                                         int32_t* __restrict__ idx, int32_t* __restrict__ count, int cap) {
    const int i_max = n - 1;
    for (int i = 1 + blockIdx.x * blockDim.x + threadIdx.x; i < i_max; i += gridDim.x * blockDim.x) {
        const float v = x[i];
        if (!(x[i - 1] < v) || !(v >= height)) continue;
        int ahead = i + 1;
        while (ahead < i_max && x[ahead] == v) ++ahead;
        if (x[ahead] < v) {
            const int mid = (i + ahead - 1) / 2;
            const int slot = atomicAdd(count, 1);
            if (slot < cap) idx[slot] = mid;
        }
    }
}

extern "C" int rmx_threshold_peaks(const float* db, int n, float height, int32_t* idx, int32_t* count, int cap,
                                   void* stream) {
    if (!db || !idx || !count) return fail(RMX_ERR_ARG, "null argument to rmx_threshold_peaks");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    if (n < 3) return RMX_OK;
    k_threshold_peaks<<<grid_for(n, 256), 256, 0, st>>>(db, n, height, idx, count, cap);
    LAUNCH_CHECK("threshold_peaks");
    return RMX_OK;
}

extern "C" int rmx_select_by_distance_host(const int32_t* pos, const float* heights, int n, int distance, uint8_t* keep) {
    // find_peaks(distance=): visit peaks from the highest to the lowest; a kept peak removes every
    // not-yet-visited neighbour closer than `distance` samples.  Ties in height are broken by
    // position (later position = higher priority, as a stable argsort visited from the end does).
    if (n <= 0) return RMX_OK;
    if (!pos || !heights || !keep) return fail(RMX_ERR_ARG, "null argument to rmx_select_by_distance_host");
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return heights[a] < heights[b]; });
    std::fill(keep, keep + n, (uint8_t)1);
    for (int o = n - 1; o >= 0; --o) {
        const int j = order[o];
        if (!keep[j]) continue;
        for (int k = j - 1; k >= 0 && pos[j] - pos[k] < distance; --k) keep[k] = 0;
        for (int k = j + 1; k < n && pos[k] - pos[j] < distance; ++k) keep[k] = 0;
    }
    return RMX_OK;
}

__device__ __forceinline__ uint32_t float_order_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// single CTA: mean (double) + radix-select of the two middle order statistics
__global__ void __launch_bounds__(1024) k_mean_median(const float* __restrict__ x, int n, float* __restrict__ out) {
    __shared__ unsigned hist[256];
    __shared__ double s_sum[32];
    __shared__ uint32_t s_prefix;
    __shared__ unsigned s_rank;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)x[i];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_sum[w];
        out[0] = (float)(t / (double)n);
    }
    float med[2];
    for (int which = 0; which < 2; ++which) {
        const unsigned rank0 = which == 0 ? (unsigned)((n - 1) / 2) : (unsigned)(n / 2);
        if (threadIdx.x == 0) { s_prefix = 0; s_rank = rank0; }
        __syncthreads();
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t k = float_order_key(x[i]);
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xffu], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned r = s_rank, b = 0;
                while (b < 255 && r >= hist[b]) { r -= hist[b]; ++b; }
                s_rank = r;
                s_prefix = prefix | (b << shift);
            }
            __syncthreads();
        }
        med[which] = key_to_float(s_prefix);
        __syncthreads();
    }
    // np.median: mean of the two middle values (identical when n is odd)
    if (threadIdx.x == 0) out[1] = (n & 1) ? med[0] : 0.5f * (med[0] + med[1]);
}

extern "C" int rmx_mean_median(const float* db, int n, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    if (!db || !out) return fail(RMX_ERR_ARG, "null argument to rmx_mean_median");
    if (n <= 0) return fail(RMX_ERR_ARG, "rmx_mean_median needs n > 0");
    k_mean_median<<<1, 1024, 0, (cudaStream_t)stream>>>(db, n, out);
    LAUNCH_CHECK("mean_median");
    return RMX_OK;
}

// ---------------------------------------------------------------------------------------
// batched block detection: scipy.signal.find_peaks(height=, distance=) + mean/median per row
// ---------------------------------------------------------------------------------------
// One CTA (1024 threads) per dB spectrum.  (1) mean and median of the row (same arithmetic as
// k_mean_median); (2) candidates = strict local maxima with plateau mid-points (_local_maxima_1d),
// value >= threshold, compacted in ascending bin order; (3) the distance rule: candidates are ranked by
// (height, position), visited from the highest priority down, and a kept candidate removes every
// not-yet-visited neighbour closer than `distance` bins (_select_by_peak_distance; ties resolved like
// rmx_select_by_distance_host); (4) the kept peaks, ascending, with their heights.
constexpr int kPeakCap = 16384;         // candidates per row held in shared memory (a 32768-bin row has at most 16383)

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* s_warp, int* total) {
    // exclusive prefix sum of one int per thread over a 1024-thread block
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = s_warp[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += t;
        }
        s_warp[lane] = x;                 // inclusive over warps
    }
    __syncthreads();
    const int base = w == 0 ? 0 : s_warp[w - 1];
    *total = s_warp[31];
    return base + inc - v;
}

__global__ void __launch_bounds__(1024) k_find_peaks_batch(const float* __restrict__ db, int n, long long row_stride,
                                                           float height, int height_mode, int distance,
                                                           int32_t* __restrict__ idx_out, float* __restrict__ h_out,
                                                           int32_t* __restrict__ count_out, int cap,
                                                           float* __restrict__ stats_out, int smem_cap,
                                                           int gate_dc_bins, float gate_conf_min) {
    extern __shared__ __align__(8) unsigned char s_dyn[];
    // [smem_cap] each: (height key << 32 | candidate index) for the priority sort, candidate bin, keep flag
    unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(s_dyn);
    int32_t* s_pos = reinterpret_cast<int32_t*>(s_sort + smem_cap);
    uint8_t* s_keep = reinterpret_cast<uint8_t*>(s_pos + smem_cap);
    __shared__ unsigned hist[256];
    __shared__ double s_sum[32];
    __shared__ uint32_t s_prefix;
    __shared__ unsigned s_rank;
    __shared__ int s_warp[32];
    __shared__ float s_stats[2];

    const float* __restrict__ x = db + (long long)blockIdx.x * row_stride;
    // ---- (1) mean, median -------------------------------------------------------------------
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)x[i];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += s_sum[w];
        s_stats[0] = (float)(t / (double)n);
    }
    float med[2];
    for (int which = 0; which < 2; ++which) {
        const unsigned rank0 = which == 0 ? (unsigned)((n - 1) / 2) : (unsigned)(n / 2);
        if (threadIdx.x == 0) { s_prefix = 0; s_rank = rank0; }
        __syncthreads();
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t k = float_order_key(x[i]);
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xffu], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned r = s_rank, b = 0;
                while (b < 255 && r >= hist[b]) { r -= hist[b]; ++b; }
                s_rank = r;
                s_prefix = prefix | (b << shift);
            }
            __syncthreads();
        }
        med[which] = key_to_float(s_prefix);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s_stats[1] = (n & 1) ? med[0] : 0.5f * (med[0] + med[1]);
        stats_out[2 * blockIdx.x] = s_stats[0];
        stats_out[2 * blockIdx.x + 1] = s_stats[1];
    }
    __syncthreads();
    // threshold: absolute, or mean + height evaluated like the host does (double sum, rounded to float)
    const float thr = height_mode ? (float)((double)s_stats[0] + (double)height) : height;

    // ---- (2) candidates in ascending order: contiguous chunk of rising edges per thread -------
    const int i_max = n - 1;
    const int chunk = (max(i_max - 1, 0) + (int)blockDim.x - 1) / (int)blockDim.x;
    const int lo = 1 + (int)threadIdx.x * chunk, hi = min(lo + chunk, i_max);
    int mine = 0;
    for (int i = lo; i < hi; ++i) {
        const float v = x[i];
        if (!(x[i - 1] < v) || !(v >= thr)) continue;
        int ahead = i + 1;
        while (ahead < i_max && x[ahead] == v) ++ahead;
        if (x[ahead] < v) ++mine;
    }
    int n_cand;
    int at = block_exclusive_scan_1024(mine, s_warp, &n_cand);
    if (n_cand > smem_cap) {                      // overflow: report the candidate count, no peaks
        if (threadIdx.x == 0) count_out[blockIdx.x] = -n_cand;
        return;
    }
    for (int i = lo; i < hi; ++i) {
        const float v = x[i];
        if (!(x[i - 1] < v) || !(v >= thr)) continue;
        int ahead = i + 1;
        while (ahead < i_max && x[ahead] == v) ++ahead;
        if (x[ahead] < v) {
            s_pos[at] = (i + ahead - 1) / 2;
            s_sort[at] = ((unsigned long long)float_order_key(v) << 32) | (unsigned)at;
            s_keep[at] = 1;
            ++at;
        }
    }
    __syncthreads();
    // ---- (3) distance rule --------------------------------------------------------------------
    // Sequential definition: visit candidates from the highest priority (height, then position) down; a kept one
    // removes its not-yet-visited neighbours closer than `distance`.  Equivalent fixed point, evaluated in
    // parallel: a candidate is REMOVED once some higher-priority neighbour within `distance` is KEPT, and KEPT once
    // all of them are removed.  Rounds = longest chain of such dependencies (a handful on noise-like spectra);
    // if 64 rounds do not settle everything, the remainder is finished in priority order by one thread.
    if (distance > 1 && n_cand > 1) {
        constexpr uint8_t REMOVED = 0, KEPT = 1, UNDECIDED = 2;
        for (int k = threadIdx.x; k < n_cand; k += blockDim.x) s_keep[k] = UNDECIDED;
        __syncthreads();
        auto higher = [&](int m, float hm, int j, float hj) { return hm > hj || (hm == hj && m > j); };
        int rounds = 0;
        bool pending = true;
        while (pending && rounds < 64) {
            int undecided = 0;
            for (int j = threadIdx.x; j < n_cand; j += blockDim.x) {
                if (s_keep[j] != UNDECIDED) continue;
                const int pj = s_pos[j];
                const float hj = x[pj];
                bool removed = false, waiting = false;
                for (int m = j - 1; m >= 0 && pj - s_pos[m] < distance; --m) {
                    if (!higher(m, x[s_pos[m]], j, hj)) continue;
                    const uint8_t st = s_keep[m];
                    removed |= st == KEPT;
                    waiting |= st == UNDECIDED;
                }
                for (int m = j + 1; m < n_cand && s_pos[m] - pj < distance; ++m) {
                    if (!higher(m, x[s_pos[m]], j, hj)) continue;
                    const uint8_t st = s_keep[m];
                    removed |= st == KEPT;
                    waiting |= st == UNDECIDED;
                }
                if (removed) s_keep[j] = REMOVED;
                else if (!waiting) s_keep[j] = KEPT;
                else undecided = 1;
            }
            pending = __syncthreads_or(undecided) != 0;
            ++rounds;
        }
        if (pending) {
            // rare: long dependency chains (e.g. a monotone ramp of plateaus).  Sort the composites and let one
            // thread finish the undecided candidates in priority order.
            int m2 = 1;
            while (m2 < n_cand) m2 <<= 1;
            for (int k = threadIdx.x; k < m2; k += blockDim.x)
                s_sort[k] = k < n_cand ? (((unsigned long long)float_order_key(x[s_pos[k]]) << 32) | (unsigned)k) : ~0ULL;
            __syncthreads();
            for (int size = 2; size <= m2; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int t = threadIdx.x; t < (m2 >> 1); t += blockDim.x) {
                        const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                        const int j = i | stride;
                        const bool up = (i & size) == 0;
                        const unsigned long long a = s_sort[i], b = s_sort[j];
                        if ((a > b) == up) { s_sort[i] = b; s_sort[j] = a; }
                    }
                    __syncthreads();
                }
            }
            if (threadIdx.x == 0) {
                for (int o = n_cand - 1; o >= 0; --o) {
                    const int j = (int)(unsigned)(s_sort[o] & 0xffffffffULL);
                    if (s_keep[j] == REMOVED) continue;
                    if (s_keep[j] == UNDECIDED) {
                        // every higher-priority candidate is decided by now
                        bool removed = false;
                        const int pj = s_pos[j];
                        const float hj = x[pj];
                        for (int m = j - 1; m >= 0 && pj - s_pos[m] < distance; --m)
                            removed |= s_keep[m] == KEPT && higher(m, x[s_pos[m]], j, hj);
                        for (int m = j + 1; m < n_cand && s_pos[m] - pj < distance; ++m)
                            removed |= s_keep[m] == KEPT && higher(m, x[s_pos[m]], j, hj);
                        s_keep[j] = removed ? REMOVED : KEPT;
                    }
                }
            }
            __syncthreads();
        }
    }
    // optional gates of the buoy detector (buoy_node.py:423-433), applied AFTER the distance rule like the
    // reference's loop over find_peaks' output: drop bins closer than gate_dc_bins to DC (|f - fc| < 10 kHz) and
    // peaks whose confidence clip((p - median) / 20, 0, 1) is below gate_conf_min (float32 arithmetic, as numpy)
    if (gate_dc_bins > 0 || gate_conf_min > 0.f) {
        const float median = s_stats[1];
        for (int k = threadIdx.x; k < n_cand; k += blockDim.x) {
            if (!s_keep[k]) continue;
            const int pos = s_pos[k];
            const float conf = fminf(fmaxf((x[pos] - median) / 20.0f, 0.0f), 1.0f);
            if (min(pos, n - pos) < gate_dc_bins || conf < gate_conf_min) s_keep[k] = 0;
        }
        __syncthreads();
    }
    // ---- (4) kept peaks, ascending ---------------------------------------------------------------
    const int per = (n_cand + (int)blockDim.x - 1) / (int)blockDim.x;
    const int k0 = (int)threadIdx.x * per, k1 = min(k0 + per, n_cand);
    int kept = 0;
    for (int k = k0; k < k1; ++k) kept += s_keep[k];
    int total;
    int w = block_exclusive_scan_1024(kept, s_warp, &total);
    int32_t* __restrict__ io = idx_out + (long long)blockIdx.x * cap;
    float* __restrict__ ho = h_out + (long long)blockIdx.x * cap;
    for (int k = k0; k < k1; ++k)
        if (s_keep[k]) {
            if (w < cap) { io[w] = s_pos[k]; ho[w] = x[s_pos[k]]; }
            ++w;
        }
    if (threadIdx.x == 0) count_out[blockIdx.x] = total;
}

extern "C" int rmx_find_peaks_batch(const float* db, int n_rows, int n, size_t row_stride, float height, int height_mode,
                                    int distance, int gate_dc_bins, float gate_conf_min, int32_t* idx, float* heights,
                                    int32_t* count, int cap, float* stats, void* stream) {
    if (!db || !idx || !heights || !count || !stats) return fail(RMX_ERR_ARG, "null argument to rmx_find_peaks_batch");
    if (n_rows <= 0) return RMX_OK;
    if (n < 1 || cap < 1) return fail(RMX_ERR_ARG, "rmx_find_peaks_batch needs n >= 1 and cap >= 1");
    if (row_stride == 0) row_stride = (size_t)n;
    if (row_stride < (size_t)n) return fail(RMX_ERR_ARG, "row_stride must be >= n");
    // shared memory for as many candidates as a row can have (a strict local maximum every other bin), rounded
    // up to a power of two for the bitonic sort and bounded by what one SM offers
    int smem_cap = 1024;
    while (smem_cap < n / 2 + 1 && smem_cap < kPeakCap) smem_cap <<= 1;
    const size_t smem = (size_t)smem_cap * (sizeof(unsigned long long) + sizeof(int32_t) + 1);
    CUDA_TRY(cudaFuncSetAttribute((const void*)k_find_peaks_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_find_peaks_batch<<<n_rows, 1024, smem, (cudaStream_t)stream>>>(db, n, (long long)row_stride, height, height_mode, distance,
                                                                       idx, heights, count, cap, stats, smem_cap, gate_dc_bins, gate_conf_min);
    LAUNCH_CHECK("find_peaks_batch");
    return RMX_OK;
}

// -3 dB bandwidth walk of iq_stream_client.py:254-278 for every peak of every row: one thread per peak.
__global__ void __launch_bounds__(256) k_peak_bandwidth_batch(const float* __restrict__ db, int n, long long row_stride,
                                                              const int32_t* __restrict__ idx, const int32_t* __restrict__ count,
                                                              int cap, float drop_db, int32_t* __restrict__ width) {
    const int row = blockIdx.y;
    const int c = min(max(count[row], 0), cap);
    const float* __restrict__ x = db + (long long)row * row_stride;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < c; k += gridDim.x * blockDim.x) {
        const int peak = idx[(long long)row * cap + k];
        const float thr = x[peak] - drop_db;
        int left = peak, right = peak;
        while (left > 0 && x[left] > thr) --left;
        while (right < n - 1 && x[right] > thr) ++right;
        width[(long long)row * cap + k] = right - left;
    }
}

extern "C" int rmx_peak_bandwidth_batch(const float* db, int n_rows, int n, size_t row_stride, const int32_t* idx,
                                        const int32_t* count, int cap, float drop_db, int32_t* width, void* stream) {
    if (!db || !idx || !count || !width) return fail(RMX_ERR_ARG, "null argument to rmx_peak_bandwidth_batch");
    if (n_rows <= 0) return RMX_OK;
    if (n < 1 || cap < 1) return fail(RMX_ERR_ARG, "rmx_peak_bandwidth_batch needs n >= 1 and cap >= 1");
    if (row_stride == 0) row_stride = (size_t)n;
    k_peak_bandwidth_batch<<<dim3((unsigned)std::min(64, (cap + 255) / 256), (unsigned)n_rows), 256, 0, (cudaStream_t)stream>>>(
        db, n, (long long)row_stride, idx, count, cap, drop_db, width);
    LAUNCH_CHECK("peak_bandwidth_batch");
    return RMX_OK;
}

// exact integer statistics: |x|^2 = ((2I-255)^2 + (2Q-255)^2) / 4
__global__ void __launch_bounds__(256) k_signal_stats(const uint8_t* __restrict__ in, size_t n,
                                                      unsigned long long* __restrict__ acc) {
    unsigned long long sum = 0;
    unsigned mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uchar2 b = reinterpret_cast<const uchar2*>(in)[i];
        const int a = 2 * (int)b.x - 255, c = 2 * (int)b.y - 255;
        const unsigned p = (unsigned)(a * a + c * c);
        sum += p;
        mx = max(mx, p);
    }
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[0], sum);
        atomicMax(reinterpret_cast<unsigned*>(&acc[1]), mx);
    }
}

__global__ void k_signal_stats_final(const unsigned long long* __restrict__ acc, size_t n, rmx_stats* __restrict__ out) {
    rmx_stats s;
    s.mean_power = ((double)acc[0] / 4.0) / (double)n;
    const unsigned mx = *reinterpret_cast<const unsigned*>(&acc[1]);
    s.peak_amplitude = sqrtf((float)mx * 0.25f);
    s.pad = 0.f;
    *out = s;
}

// per-signal exact energy: sum over samples of (2I-255)^2 + (2Q-255)^2  ( = 4 * sum |x|^2 )
__global__ void __launch_bounds__(256) k_signal_energy(const uint8_t* __restrict__ in, size_t stride_bytes, size_t n,
                                                       unsigned long long* __restrict__ out) {
    const uint8_t* __restrict__ base = in + (size_t)blockIdx.y * stride_bytes;
    unsigned long long sum = 0;
    const bool vec = ((uintptr_t)base % 16 == 0);
    if (vec) {
        const uint4* __restrict__ b4 = reinterpret_cast<const uint4*>(base);
        const size_t nvec = n / 8;
        for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
            const uint4 w = __ldg(b4 + v);
            const unsigned ws[4] = {w.x, w.y, w.z, w.w};
            unsigned part = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int a = 2 * (int)((ws[k] >> (8 * b)) & 0xffu) - 255;
                    part += (unsigned)(a * a);
                }
            sum += part;
        }
        for (size_t i = nvec * 8 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const int a = 2 * (int)base[2 * i] - 255, c = 2 * (int)base[2 * i + 1] - 255;
            sum += (unsigned)(a * a + c * c);
        }
    } else {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const int a = 2 * (int)base[2 * i] - 255, c = 2 * (int)base[2 * i + 1] - 255;
            sum += (unsigned)(a * a + c * c);
        }
    }
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&out[blockIdx.y], sum);
}

extern "C" int rmx_signal_energy(const uint8_t* iq, size_t signal_stride_bytes, int n_signals, size_t n_samples,
                                 unsigned long long* out, void* stream) {
    if (!iq || !out) return fail(RMX_ERR_ARG, "null argument to rmx_signal_energy");
    if (n_signals <= 0 || n_samples == 0) return fail(RMX_ERR_ARG, "rmx_signal_energy needs positive sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (signal_stride_bytes == 0) signal_stride_bytes = 2 * n_samples;
    CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)n_signals * sizeof(unsigned long long), st));
    const unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((long long)(n_samples / 8 + 255) / 256, 296 / std::max(1, std::min(n_signals, 296)) + 1));
    k_signal_energy<<<dim3(bx, n_signals), 256, 0, st>>>(iq, signal_stride_bytes, n_samples, out);
    LAUNCH_CHECK("signal_energy");
    return RMX_OK;
}

// statistics of complex64 samples: double sum of |x|^2 and max |x|^2 (x^2 + y^2 is exact in double,
// so the final sqrt is the correctly rounded |x| that numpy's hypot-based np.abs returns)
__global__ void __launch_bounds__(256) k_signal_stats_c64(const float2* __restrict__ x, size_t n, double* __restrict__ sum_out,
                                                          unsigned long long* __restrict__ max_out) {
    double sum = 0.0, mx = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 v = x[i];
        sum += (double)(v.x * v.x + v.y * v.y);          // float32 |x|^2 like np.abs(x)**2, accumulated in double
        mx = fmax(mx, (double)v.x * (double)v.x + (double)v.y * (double)v.y);
    }
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sum_out, sum);
        atomicMax(max_out, (unsigned long long)__double_as_longlong(mx));   // non-negative doubles order like their bits
    }
}

__global__ void k_signal_stats_c64_final(const double* __restrict__ sum_in, const unsigned long long* __restrict__ max_in,
                                         size_t n, rmx_stats* __restrict__ out) {
    rmx_stats s;
    s.mean_power = *sum_in / (double)n;
    s.peak_amplitude = (float)sqrt(__longlong_as_double((long long)*max_in));
    s.pad = 0.f;
    *out = s;
}

extern "C" int rmx_signal_stats_c64(const rmx_complex64* x, size_t n_samples, rmx_stats* out, void* workspace, void* stream) {
    if (!x || !out || !workspace) return fail(RMX_ERR_ARG, "null argument to rmx_signal_stats_c64");
    if (n_samples == 0) return fail(RMX_ERR_ARG, "rmx_signal_stats_c64 needs n_samples > 0");
    cudaStream_t st = (cudaStream_t)stream;
    double* sum = reinterpret_cast<double*>(workspace);
    unsigned long long* mx = reinterpret_cast<unsigned long long*>(sum + 1);
    CUDA_TRY(cudaMemsetAsync(workspace, 0, 16, st));
    k_signal_stats_c64<<<grid_for((long long)n_samples, 256), 256, 0, st>>>(reinterpret_cast<const float2*>(x), n_samples, sum, mx);
    LAUNCH_CHECK("signal_stats_c64");
    k_signal_stats_c64_final<<<1, 1, 0, st>>>(sum, mx, n_samples, out);
    LAUNCH_CHECK("signal_stats_c64_final");
    return RMX_OK;
}

extern "C" int rmx_signal_stats(const uint8_t* iq, size_t n_samples, rmx_stats* out, void* stream) {
    if (!iq || !out) return fail(RMX_ERR_ARG, "null argument to rmx_signal_stats");
    if (n_samples == 0) return fail(RMX_ERR_ARG, "rmx_signal_stats needs n_samples > 0");
    cudaStream_t st = (cudaStream_t)stream;
    // the 16 accumulator bytes live in the output struct's storage until the final kernel overwrites it
    static_assert(sizeof(rmx_stats) >= 2 * sizeof(unsigned long long), "rmx_stats too small for the accumulators");
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(out);
    CUDA_TRY(cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), st));
    k_signal_stats<<<grid_for((long long)n_samples, 256), 256, 0, st>>>(iq, n_samples, acc);
    LAUNCH_CHECK("signal_stats");
    k_signal_stats_final<<<1, 1, 0, st>>>(acc, n_samples, out);
    LAUNCH_CHECK("signal_stats_final");
    return RMX_OK;
}
