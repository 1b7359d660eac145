// radix_stage_bench2.cu -- companion of radix_stage_bench.cu for the two-transforms-per-thread packed fp32x2 tile
// (rmx_fft_soa.cuh): one CTA of 256 threads re-transforms TWO resident 4096-point tiles ITERS times.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../../radio_mapper_b200/csrc -o radix_stage_bench2 radix_stage_bench2.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "rmx_fft_soa.cuh"

using namespace rmx;
using GEO = TileGeom<12, 4, false>;

#ifndef CTAS
#define CTAS 2
#endif

__global__ void __launch_bounds__(kThreads, CTAS) k_radix_tile2(StageTables tabs, int iters, float2* sink, long long* cycles) {
    extern __shared__ float2 smem[];
    float2* pre = smem;
    float2* pim = smem + GEO::NP;
    const int i0 = threadIdx.x, g = 0;
    C2 r[GEO::E];
#pragma unroll
    for (int u = 0; u < GEO::E; ++u) {
        r[u].re = make_float2((float)((i0 * 31 + u * 7) % 255) - 127.5f, (float)((i0 * 13 + u * 5) % 255) - 127.5f);
        r[u].im = make_float2((float)((i0 * 17 + u * 3) % 255) - 127.5f, (float)((i0 * 19 + u * 11) % 255) - 127.5f);
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        fft_tile2<GEO, false>(r, pre, pim, g, i0, tabs);
        const float2 sc = make_float2(0.015625f, 0.015625f);
#pragma unroll
        for (int u = 0; u < GEO::E; ++u) { r[u].re = __fmul2_rn(r[u].re, sc); r[u].im = __fmul2_rn(r[u].im, sc); }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < GEO::E; ++u) { acc = __fadd2_rn(acc, r[u].re); acc = __fadd2_rn(acc, r[u].im); }
    sink[blockIdx.x * kThreads + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    StageTables tabs{};
    const int loge = 4;
    for (int s = 1; s < 3; ++s) {
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
    const int grid = CTAS * sms;
    const size_t smem = soa_smem_bytes<GEO>();
    float2* sink;
    long long* cyc;
    cudaMalloc(&sink, (size_t)grid * kThreads * sizeof(float2));
    cudaMalloc(&cyc, grid * 8);
    cudaFuncSetAttribute((const void*)k_radix_tile2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_radix_tile2<<<grid, kThreads, smem>>>(tabs, iters, sink, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_radix_tile2<<<grid, kThreads, smem>>>(tabs, iters, sink, cyc);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tiles_total = 2.0 * (double)grid * iters;
    const double us_per_tile_chip = ms * 1e3 / tiles_total;
    const double welch_tiles = 1000.0 * 65536.0 / 4096.0;
    printf("{\"proto\": \"CUDA-core radix kernel, TWO transforms per thread in packed fp32x2 registers (rmx::fft_tile2), operands resident on chip\", "
           "\"sms\": %d, \"ctas_per_sm\": %d, \"iters_per_cta\": %d, \"ms_total\": %.4f, \"us_per_tile_chip\": %.5f, "
           "\"us_per_radix64_equivalent_stage_of_welch_cfg2\": %.1f, \"us_per_full_64k_fft_of_welch_cfg2\": %.1f}\n",
           sms, CTAS, iters, ms, us_per_tile_chip, us_per_tile_chip * welch_tiles * 0.5, us_per_tile_chip * welch_tiles * 4.0 / 3.0);
    return 0;
}
