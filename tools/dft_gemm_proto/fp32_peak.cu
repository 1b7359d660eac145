// fp32_peak.cu -- what FP32 rate does this B200 sustain?  (developer aid; DESIGN.md section 3, "the FP32 rate limit")
// Register-only instruction streams (no memory, no barriers), 256-thread CTAs, 1..8 CTAs per SM:
//   ffma  : 16 independent FFMA chains per thread            (2 flop per lane-op)
//   ffma2 : 8 independent packed FFMA2 chains per thread     (4 flop per lane-instruction)
//   fadd  : 16 independent FADD chains per thread            (1 flop per lane-op)
// Reports, per occupancy: wall time (CUDA events) -> TFLOP/s and the share of the issue slots used at the nominal SM
// clock; with ONE CTA per SM also the SM clock seen by the CTA (clock64 / wall time; with several CTAs per SM the
// scheduler serves them unevenly, so a CTA's lifetime is not the kernel's).  The finding it documents: 8 warps per SM
// with 16 independent chains each already run the FP32 pipe at its peak (70-73 TFLOP/s of the nominal 74.4).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp32_peak fp32_peak.cu   (or __graft_entry__.build())
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int kThreads = 256;

template <int MODE>
__global__ void __launch_bounds__(kThreads) k_stream(int iters, float seed, float* sink, long long* cycles) {
    float a[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) a[u] = seed + (float)(threadIdx.x * 16 + u) * 1e-6f;
    const float m = 0.999999f, c = 1e-7f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            if constexpr (MODE == 0) {
#pragma unroll
                for (int u = 0; u < 16; ++u) a[u] = fmaf(a[u], m, c);
            } else if constexpr (MODE == 1) {
#pragma unroll
                for (int u = 0; u < 16; u += 2) {
                    float2 v = make_float2(a[u], a[u + 1]);
                    v = __ffma2_rn(v, make_float2(m, m), make_float2(c, c));
                    a[u] = v.x; a[u + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int u = 0; u < 16; ++u) a[u] = a[u] + c;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 16; ++u) s += a[u];
    sink[blockIdx.x * kThreads + threadIdx.x] = s;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    int sms = 148, mhz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    float* sink;
    long long* cyc;
    cudaMalloc(&sink, (size_t)8 * sms * kThreads * sizeof(float));
    cudaMalloc(&cyc, 8 * sms * 8);
    std::vector<long long> h(8 * sms);
    const char* names[3] = {"ffma", "ffma2", "fadd"};
    const double flop_per_lane_instr[3] = {2.0, 4.0, 1.0};
    const double lane_instr_per_iter[3] = {128.0, 64.0, 128.0};
    // optional 2nd / 3rd argument: only this stream (0 ffma, 1 ffma2, 2 fadd) / only this many CTAs per SM -- for runs long
    // enough to sit at the board power limit
    const int only_mode = argc > 2 ? atoi(argv[2]) : -1, only_k = argc > 3 ? atoi(argv[3]) : -1;
    for (int mode = 0; mode < 3; ++mode)
        for (int k : {1, 2, 3, 4, 6, 8}) {
            if ((only_mode >= 0 && mode != only_mode) || (only_k >= 0 && k != only_k)) continue;
            const int grid = k * sms;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k_stream<0><<<grid, kThreads>>>(iters, 1.0f, sink, cyc);
                else if (mode == 1) k_stream<1><<<grid, kThreads>>>(iters, 1.0f, sink, cyc);
                else k_stream<2><<<grid, kThreads>>>(iters, 1.0f, sink, cyc);
                cudaEventRecord(e1);
                if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
            double mean = 0;
            for (int i = 0; i < grid; ++i) mean += (double)h[i];
            mean /= grid;
            const double instr = (double)grid * kThreads * iters * lane_instr_per_iter[mode];      // thread-instructions
            const double tflops = instr * flop_per_lane_instr[mode] / (ms * 1e-3) / 1e12;
            const double issue_util = instr / 32.0 / (sms * 4.0) / (ms * 1e-3 * mhz * 1e3);          // warp-instr per scheduler-cycle at the nominal clock
            printf("{\"stream\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"issue_slots_used_at_nominal_clock\": %.3f, "
                   "\"tflops\": %.2f, \"nominal_mhz\": %d", names[mode], 8 * k, ms, issue_util, tflops, mhz / 1000);
            if (k == 1) printf(", \"sm_mhz_seen_by_cta\": %.0f", mean / (ms * 1e3));
            printf("}\n");
        }
    return 0;
}
