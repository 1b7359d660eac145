// radix_stage_bench.cu -- PROTOTYPE companion of dft_gemm_proto.cu: the CUDA-core radix kernel under the same conditions
// (operands resident on chip, no global traffic in the timed loop).  One CTA = 256 threads transforms 4096-point tiles
// (three radix-16 Stockham stages, two shared-memory exchanges, stage twiddles from the plan tables) with the product's
// own building block rmx::fft_tile; the loop re-transforms the register tile ITERS times.  Reports cycles per tile, from
// which the cost of a radix-64-equivalent stage (1.5 radix-16 stages) over the Welch workload follows.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../../radio_mapper_b200/csrc -o radix_stage_bench radix_stage_bench.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "rmx_fft_core.cuh"

using namespace rmx;
using GEO = TileGeom<12, 4, false>;

__global__ void __launch_bounds__(kThreads, 3) k_radix_tile(StageTables tabs, int iters, float2* sink, long long* cycles) {
    extern __shared__ float2 smem[];
    const int i0 = threadIdx.x, g = 0;
    float2 r[GEO::E];
#pragma unroll
    for (int u = 0; u < GEO::E; ++u) r[u] = make_float2((float)((i0 * 31 + u * 7) % 255) - 127.5f, (float)((i0 * 17 + u * 3) % 255) - 127.5f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        fft_tile<GEO, false, true>(r, smem, g, i0, tabs);
#pragma unroll
        for (int u = 0; u < GEO::E; ++u) { r[u].x *= 0.015625f; r[u].y *= 0.015625f; }      // keep magnitudes bounded
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < GEO::E; ++u) { acc.x += r[u].x; acc.y += r[u].y; }
    sink[blockIdx.x * kThreads + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    StageTables tabs{};
    const int logn = 12, loge = 4;
    for (int s = 1; s < 3; ++s) {                      // same tables as rmx_lib.cu:build_stage_tables
        const int logp = s * loge, logr = loge, P = 1 << logp, R = 1 << logr;
        std::vector<float2> h((size_t)(R - 1) * P);
        for (int q = 1; q < R; ++q)
            for (int k = 0; k < P; ++k) {
                const double a = -2.0 * M_PI * (double)(((long long)q * k) % ((long long)P * R)) / ((double)P * R);
                h[(size_t)(q - 1) * P + k] = make_float2((float)cos(a), (float)sin(a));
            }
        float2* d;
        cudaMalloc(&d, h.size() * sizeof(float2));
        cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
        tabs.tw[s] = d;
    }
    (void)logn;
    const int grid = 3 * sms;
    float2* sink;
    long long* cyc;
    cudaMalloc(&sink, (size_t)grid * kThreads * sizeof(float2));
    cudaMalloc(&cyc, grid * 8);
    cudaFuncSetAttribute((const void*)k_radix_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::SMEM_BYTES);
    k_radix_tile<<<grid, kThreads, GEO::SMEM_BYTES>>>(tabs, iters, sink, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_radix_tile<<<grid, kThreads, GEO::SMEM_BYTES>>>(tabs, iters, sink, cyc);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // tiles per SM-second: 3 resident CTAs share an SM
    const double tiles_total = (double)grid * iters;                       // 4096-point tiles, three radix-16 stages each
    const double us_per_tile_chip = ms * 1e3 / tiles_total;                // chip-wide time per tile
    const double welch_tiles = 1000.0 * 65536.0 / 4096.0;
    // a 64k FFT = 4 radix-16 stages; this tile has 3 -> scale by 4/3; a radix-64-equivalent stage = 1.5 radix-16 stages -> scale by 0.5
    printf("{\"proto\": \"CUDA-core radix kernel (rmx::fft_tile, 4096-point tiles = 3 radix-16 stages + 2 exchanges, twiddle tree), operands resident on chip\", "
           "\"sms\": %d, \"ctas_per_sm\": 3, \"iters_per_cta\": %d, \"ms_total\": %.4f, "
           "\"us_per_radix64_equivalent_stage_of_welch_cfg2\": %.1f, \"us_per_full_64k_fft_of_welch_cfg2\": %.1f}\n",
           sms, iters, ms, us_per_tile_chip * welch_tiles * 0.5, us_per_tile_chip * welch_tiles * 4.0 / 3.0);
    return 0;
}
