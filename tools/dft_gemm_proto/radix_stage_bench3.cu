// radix_stage_bench3.cu -- PROTOTYPE: how many warps does the radix-16 butterfly code need to fill the FP32 pipe?
// Kernel A (fp): per iteration 15 twiddle multiplies + one 16-point register DFT, nothing else (no shared memory, no
// barrier).  Kernel B (tile): rmx::fft_tile on a resident 4096-point tile (3 butterfly stages + 2 exchanges).
// Both with 1, 2 and 3 resident CTAs of 256 threads per SM (same binary: 80 registers).  Reported: wall time per
// iteration per SM (CUDA events), and with ONE CTA per SM the cycles per iteration (clock64; with several CTAs per SM
// the scheduler serves them unevenly -- mean CTA lifetime 3/4 resp. 2/3 of the kernel's -- so only wall time compares).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../../radio_mapper_b200/csrc -o radix_stage_bench3 radix_stage_bench3.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "rmx_fft_core.cuh"

using namespace rmx;
using GEO = TileGeom<12, 4, false>;

__global__ void __launch_bounds__(kThreads, 3) k_fp(int iters, float2* sink, long long* cycles) {
    const int i0 = threadIdx.x;
    float2 r[16], w[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        r[u] = make_float2((float)((i0 * 31 + u * 7) % 255) - 127.5f, (float)((i0 * 17 + u * 3) % 255) - 127.5f);
        float s, c;
        sincospif((float)((i0 * u) & 4095) * (1.0f / 2048.0f), &s, &c);
        w[u] = make_float2(c, s);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 1; u < 16; ++u) r[u] = cmul(r[u], w[u]);
        dft_regs<16, true>(r);
#pragma unroll
        for (int u = 0; u < 16; ++u) { r[u].x *= 0.25f; r[u].y *= 0.25f; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 16; ++u) { acc.x += r[u].x; acc.y += r[u].y; }
    sink[blockIdx.x * kThreads + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(kThreads, 3) k_tile(StageTables tabs, int iters, float2* sink, long long* cycles) {
    extern __shared__ float2 smem[];
    const int i0 = threadIdx.x, g = 0;
    float2 r[GEO::E];
#pragma unroll
    for (int u = 0; u < GEO::E; ++u) r[u] = make_float2((float)((i0 * 31 + u * 7) % 255) - 127.5f, (float)((i0 * 17 + u * 3) % 255) - 127.5f);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        fft_tile<GEO, false, true>(r, smem, g, i0, tabs);
#pragma unroll
        for (int u = 0; u < GEO::E; ++u) { r[u].x *= 0.015625f; r[u].y *= 0.015625f; }
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
        tabs.tw[s] = d;
    }
    float2* sink;
    long long* cyc;
    cudaMalloc(&sink, (size_t)3 * sms * kThreads * sizeof(float2));
    cudaMalloc(&cyc, 3 * sms * 8);
    cudaFuncSetAttribute((const void*)k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::SMEM_BYTES);
    std::vector<long long> h(3 * sms);
    for (int which = 0; which < 2; ++which)
        for (int k = 1; k <= 3; ++k) {
            const int grid = k * sms;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_fp<<<grid, kThreads>>>(iters, sink, cyc);
                else k_tile<<<grid, kThreads, GEO::SMEM_BYTES>>>(tabs, iters, sink, cyc);
                cudaEventRecord(e1);
                if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
            double mean = 0;
            for (int i = 0; i < grid; ++i) mean += (double)h[i];
            mean /= grid * (double)iters;
            printf("{\"kernel\": \"%s\", \"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"ns_per_sm_iteration\": %.1f, \"ms\": %.3f, \"mean_cta_lifetime_cycles_per_iteration\": %.1f}\n",
                   which == 0 ? "fp: 15 cmul + dft16 + scale, registers only" : "tile: rmx::fft_tile 4096 points", k, 8 * k, ms * 1e6 / iters / k, ms, mean);
        }
    return 0;
}
