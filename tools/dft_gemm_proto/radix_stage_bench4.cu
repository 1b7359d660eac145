// radix_stage_bench4.cu -- PROTOTYPE: from the resident-tile microbenchmark to the real 4096-point row pass, one
// ingredient at a time, to see which one costs the FP32-pipe efficiency (92 % resident, 63 % in k_contig_pair_run).
//   level 0  rmx::fft_tile on a resident register tile (radix_stage_bench3's "tile")
//   level 1  + pair product with a stationary X_i row in registers + inter-pass twiddles (arithmetic only)
//   level 2  + the X_j row read from a 32 KB shared-memory landing buffer every iteration
//   level 3  + the finished row stored to global memory (16 x 8-byte stores per thread, a fresh row every iteration)
//   level 4  + the next X_j row prefetched by a bulk copy (mbarrier) from an L2-resident source, as the product does
//   level 5  = level 4 with the split-phase exchange barriers (rmx::fft_tile_split), i.e. the product's loop body
//   level 6  = level 5 without the stores (level 4's loads, no level 3)
//   level 7  level 0 + pair product only            level 8  level 0 + inter-pass twiddles only
//   level 9  level 1 with the X_i row read from shared memory at use instead of held in registers
//   level 10 level 1 with two-level inter-pass twiddles (4 + 4 factors live instead of 16)
//   level 11 level 3 with streaming stores (st.global.cs)     level 12 level 3 into a 16 MB (L2-resident) destination
//   level 13 level 2 + the row through a shared-memory staging buffer + one bulk copy (TMA engine)
//   level 14 level 2 + the row written to the staging buffer only (no global store)
//   level 15 design B: X_i row in shared memory (read at use), X_j by per-thread loads from global, direct row stores
//   level 16 design B + the row staged in the EXCHANGE buffer + one bulk copy (the copy drains during the next tile's loads)
//   level 17 level 16 with two-level inter-pass twiddles
// 3 CTAs of 256 threads per SM, wall time per tile and SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../../radio_mapper_b200/csrc -o radix_stage_bench4 radix_stage_bench4.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "rmx_kernels.cuh"

using namespace rmx;
using GEO = TileGeom<12, 4, false>;
constexpr int E = GEO::E, NT = GEO::NT;
constexpr uint32_t ROW_BYTES = (uint32_t)(GEO::N * sizeof(float2));

struct P4 {
    StageTables tabs;
    const float2* src;      // L2-resident spectrum rows
    float2* dst;            // large workspace
    unsigned src_rows, dst_rows;
    int iters;
    float2* sink;
};

template <int LEVEL>
__global__ void __launch_bounds__(kThreads, 3) k_level(const P4 p) {
    extern __shared__ float2 smem[];
    __shared__ float2 s_pw[8];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ __align__(8) unsigned long long s_split[2];
    float2* land = smem + ((GEO::NP + 15) & ~15);
    const int i0 = threadIdx.x, g = 0;
    constexpr bool PAIR = (LEVEL >= 1 && LEVEL <= 6) || LEVEL == 7 || LEVEL >= 9, POST = (LEVEL >= 1 && LEVEL <= 6) || LEVEL >= 8;
    constexpr bool LAND = (LEVEL >= 2 && LEVEL <= 6) || (LEVEL >= 11 && LEVEL <= 14), STORE = (LEVEL >= 3 && LEVEL <= 5) || LEVEL == 11 || LEVEL == 12 || LEVEL == 15;
    constexpr bool PREFETCH = LEVEL >= 4 && LEVEL <= 6, SPLIT = LEVEL == 5 || LEVEL == 6;
    constexpr bool DESIGN_B = LEVEL >= 15;
    constexpr bool XI_SMEM = LEVEL == 9 || DESIGN_B, POST2 = LEVEL == 10 || LEVEL == 17, STAGE = LEVEL == 13 || LEVEL == 14;
    constexpr bool XSTAGE = LEVEL == 16 || LEVEL == 17;
    if (threadIdx.x < 4) s_pw[threadIdx.x] = unit_root((blockIdx.x * (NT << threadIdx.x)) & 0x1fffffu, 21, true);
    const float2 tw_base = row_twiddle_base<E>(blockIdx.x & 511u, (uint32_t)i0, 21, true, 1.0f / 4096.0f);
    float2 a[E], r[E];
#pragma unroll
    for (int u = 0; u < E; ++u) {
        a[u] = make_float2((float)((i0 * 13 + u * 5) % 255) - 127.5f, (float)((i0 * 19 + u * 11) % 255) - 127.5f);
        r[u] = make_float2((float)((i0 * 31 + u * 7) % 255) - 127.5f, (float)((i0 * 17 + u * 3) % 255) - 127.5f);
        if (LAND || XI_SMEM) land[i0 + u * NT] = XI_SMEM ? a[u] : r[u];
    }
    SplitBarriers sb;
    if (SPLIT) split_init(sb, s_split);
    unsigned srow = (blockIdx.x * 7u) % p.src_rows;
    if (PREFETCH && threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        mbar_expect_tx(&mbar, ROW_BYTES);
        bulk_load_1d(land, p.src + (size_t)srow * GEO::N, ROW_BYTES, &mbar);
    }
    __syncthreads();
    uint32_t parity = 0;
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < p.iters; ++it) {
        if constexpr (PREFETCH) {
            mbar_wait(&mbar, parity);
            parity ^= 1u;
        }
        if constexpr (LAND) {
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = land[i0 + u * NT];
        }
        if constexpr (DESIGN_B) {
            srow = (srow + 13u) % p.src_rows;
            const float2* __restrict__ xj = p.src + (size_t)srow * GEO::N;
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = __ldg(xj + i0 + u * NT);
            if constexpr (XSTAGE) {                      // the previous row has left the exchange buffer
                if (it > 0) { if (threadIdx.x == 0) bulk_store_wait_read(); __syncthreads(); }
            }
        }
        if constexpr (PAIR) {
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = cmul_conj(r[u], XI_SMEM ? land[i0 + u * NT] : a[u]);
        }
        if constexpr (PREFETCH) {
            __syncthreads();
            if (threadIdx.x == 0 && it + 1 < p.iters) {
                srow = (srow + 13u) % p.src_rows;
                fence_proxy_async();
                mbar_expect_tx(&mbar, ROW_BYTES);
                bulk_load_1d(land, p.src + (size_t)srow * GEO::N, ROW_BYTES, &mbar);
            }
        }
        if constexpr (SPLIT) fft_tile_split<GEO, true>(r, smem, g, i0, p.tabs, sb);
        else fft_tile<GEO, true, true>(r, smem, g, i0, p.tabs);
        if constexpr (POST2) {
            // tw[u] = base * step^u as hi[u >> 2] * lo[u & 3]: 8 factors live instead of 16
            const float2 s1 = s_pw[0], s2 = s_pw[1], s4 = s_pw[2], s8 = s_pw[3];
            const float2 lo1 = s1, lo2 = s2, lo3 = cmul(s1, s2);
            const float2 hi0 = tw_base, hi1 = cmul(tw_base, s4), hi2 = cmul(tw_base, s8), hi3 = cmul(hi1, s8);
            const float2 hi[4] = {hi0, hi1, hi2, hi3};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                r[4 * m] = cmul(r[4 * m], hi[m]);
                r[4 * m + 1] = cmul(r[4 * m + 1], cmul(hi[m], lo1));
                r[4 * m + 2] = cmul(r[4 * m + 2], cmul(hi[m], lo2));
                r[4 * m + 3] = cmul(r[4 * m + 3], cmul(hi[m], lo3));
            }
        } else if constexpr (POST) {
            float2 tw[E];
            row_twiddles_from_base<E>(tw, tw_base, s_pw);
#pragma unroll
            for (int u = 0; u < E; ++u) r[u] = cmul(r[u], tw[u]);
        } else {
#pragma unroll
            for (int u = 0; u < E; ++u) { r[u].x *= 0.015625f; r[u].y *= 0.015625f; }
        }
        if constexpr (STORE) {
            const unsigned rows = LEVEL == 12 ? 512u : p.dst_rows;
            float2* __restrict__ out = p.dst + (size_t)(((unsigned long long)blockIdx.x * p.iters + it) % rows) * GEO::N;
#pragma unroll
            for (int u = 0; u < E; ++u) {
                if constexpr (LEVEL == 11) __stcs(out + i0 + u * NT, r[u]);
                else out[i0 + u * NT] = r[u];
            }
        } else if constexpr (XSTAGE) {
            float2* __restrict__ out = p.dst + (size_t)(((unsigned long long)blockIdx.x * p.iters + it) % p.dst_rows) * GEO::N;
            __syncthreads();                             // every thread is past its last exchange read
#pragma unroll
            for (int u = 0; u < E; ++u) smem[i0 + u * NT] = r[u];
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) bulk_store_1d(out, smem, ROW_BYTES);
        } else if constexpr (STAGE) {
            float2* __restrict__ out = p.dst + (size_t)(((unsigned long long)blockIdx.x * p.iters + it) % p.dst_rows) * GEO::N;
            if (LEVEL == 13 && it > 0) { if (threadIdx.x == 0) bulk_store_wait_read(); __syncthreads(); }
#pragma unroll
            for (int u = 0; u < E; ++u) land[i0 + u * NT] = r[u];
            if constexpr (LEVEL == 13) {
                fence_proxy_async();
                __syncthreads();
                if (threadIdx.x == 0) bulk_store_1d(out, land, ROW_BYTES);
            }
        } else if constexpr (LAND) {
#pragma unroll
            for (int u = 0; u < E; ++u) { acc.x += r[u].x; acc.y += r[u].y; }
        }
        if constexpr (!SPLIT && !XSTAGE) __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < E; ++u) { acc.x += r[u].x; acc.y += r[u].y; }
    if ((LEVEL == 13 || XSTAGE) && threadIdx.x == 0) bulk_store_wait_read();
    p.sink[blockIdx.x * kThreads + threadIdx.x] = acc;
}

template <int LEVEL>
static float run(const P4& p, int grid, size_t smem) {
    cudaFuncSetAttribute((const void*)k_level<LEVEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_level<LEVEL><<<grid, kThreads, smem>>>(p);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "level %d failed: %s\n", LEVEL, cudaGetErrorString(cudaGetLastError())); exit(1); }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    return ms;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    P4 p{};
    for (int s = 1; s < 3; ++s) {
        const int logp = s * 4, P = 1 << logp, R = 16;
        std::vector<float2> h((size_t)(R - 1) * P);
        for (int q = 1; q < R; ++q)
            for (int k = 0; k < P; ++k) {
                const double a = -2.0 * M_PI * (double)(((long long)q * k) % ((long long)P * R)) / ((double)P * R);
                h[(size_t)(q - 1) * P + k] = make_float2((float)cos(a), (float)sin(a));
            }
        float2* d;
        cudaMalloc(&d, h.size() * sizeof(float2));
        cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
        p.tabs.tw[s] = d;
    }
    p.src_rows = 2048;                                  // 64 MB: L2-resident like the spectra of a window
    p.dst_rows = 1u << 18;                              // 8 GB workspace
    float2 *src, *dst;
    cudaMalloc(&src, (size_t)p.src_rows * ROW_BYTES);
    cudaMalloc(&dst, (size_t)p.dst_rows * ROW_BYTES);
    cudaMemset(src, 0x3c, (size_t)p.src_rows * ROW_BYTES);
    p.src = src; p.dst = dst; p.iters = iters;
    const int grid = 3 * sms;
    cudaMalloc(&p.sink, (size_t)grid * kThreads * sizeof(float2));
    const size_t smem = (size_t((GEO::NP + 15) & ~15) + GEO::N) * sizeof(float2);
    const char* what[18] = {"fft_tile, resident", "+ pair product, inter-pass twiddles", "+ X_j from the landing buffer (LDS)",
                            "+ row stores to global", "+ bulk-copy prefetch of the next X_j row", "= with split-phase barriers (product loop body)",
                            "level 5 without the row stores", "level 0 + pair product only", "level 0 + inter-pass twiddles only",
                            "level 1, X_i row from shared memory at use", "level 1, two-level inter-pass twiddles",
                            "level 3 with st.global.cs", "level 3 into a 16 MB destination", "level 2 + staged row + bulk store",
                            "level 2 + staged row only (no global store)", "design B: X_i in smem, X_j by LDG, direct stores",
                            "design B + row staged in the exchange buffer + bulk store", "level 16 with two-level inter-pass twiddles"};
    float ms[18];
    ms[0] = run<0>(p, grid, smem); ms[1] = run<1>(p, grid, smem); ms[2] = run<2>(p, grid, smem); ms[3] = run<3>(p, grid, smem);
    ms[4] = run<4>(p, grid, smem); ms[5] = run<5>(p, grid, smem); ms[6] = run<6>(p, grid, smem); ms[7] = run<7>(p, grid, smem);
    ms[8] = run<8>(p, grid, smem); ms[9] = run<9>(p, grid, smem); ms[10] = run<10>(p, grid, smem); ms[11] = run<11>(p, grid, smem);
    ms[12] = run<12>(p, grid, smem); ms[13] = run<13>(p, grid, smem); ms[14] = run<14>(p, grid, smem);
    ms[15] = run<15>(p, grid, smem); ms[16] = run<16>(p, grid, smem); ms[17] = run<17>(p, grid, smem);
    for (int l = 0; l < 18; ++l)
        printf("{\"level\": %d, \"what\": \"%s\", \"ctas_per_sm\": 3, \"iters\": %d, \"ms\": %.3f, \"ns_per_tile_per_sm\": %.1f, \"cycles_at_1965MHz\": %.0f}\n",
               l, what[l], iters, ms[l], ms[l] * 1e6 / iters / 3.0, ms[l] * 1e6 / iters / 3.0 * 1.965);
    return 0;
}
