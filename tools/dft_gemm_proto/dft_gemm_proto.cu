// dft_gemm_proto.cu -- PROTOTYPE (not on the product path): one radix-64 DFT stage as a GEMM on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), with the 3xTF32 split that holds float32 accuracy.
//
// Why: BASELINE.json north_star -- "tensor cores are used only if a DFT-as-GEMM FFT stage beats the SMEM radix kernel
// in measured ncu counters".  This program measures that stage on B200 (time per tile with operands resident in shared
// memory, i.e. the best case for the tensor path) and its error against float64, so the decision rests on numbers.
//
// The stage:  Y[64 x C] = F64 . X[64 x C]  (complex), written as the real GEMM
//     [Yr]   [ C  S] [Xr]
//     [Yi] = [-S  C] [Xi]        C[k][n] = cos(2 pi k n / 64),  S[k][n] = sin(2 pi k n / 64)
// i.e. D[128 x N] = A[128 x 128] . B[128 x N] with N = 64 columns per tile (4096 complex points).
// 3xTF32: A = Ah + Al, B = Bh + Bl (h = top 19 bits, what kind::tf32 reads; l = remainder): D = Ah.Bh + Ah.Bl + Al.Bh.
// Operands sit in shared memory in the canonical K-major, no-swizzle UMMA layout (8-row x 16-byte core matrices;
// LBO = 128 B between K-adjacent cores, SBO = 4096 B between 8-row groups); 16 MMAs of K = 8 per product.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o dft_gemm_proto dft_gemm_proto.cu
// run  : ./dft_gemm_proto [iterations per CTA]      (prints one JSON line)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

constexpr int M = 128, K = 128, N = 64;          // GEMM tile: D[M x N] += A[M x K] B[K x N]
constexpr int UMMA_K = 8;                          // tf32: 32 bytes of K per instruction
constexpr uint32_t LBO = 128, SBO = (K / 4) * 128; // bytes
constexpr int A_BYTES = M * K * 4, B_BYTES = N * K * 4;
constexpr int TMEM_COLS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, k) in the canonical K-major no-swizzle layout
__host__ __device__ inline uint32_t canon(int r, int k) { return (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, 16-byte units
    d |= (uint64_t)(LBO >> 4) << 16;                     // leading byte offset (K direction)
    d |= (uint64_t)(SBO >> 4) << 32;                     // stride byte offset (M/N direction)
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    return d;                                            // layout_type = 0: SWIZZLE_NONE
}

// instruction descriptor: c = f32, a = b = tf32, K-major both, M = 128, N = 64
__host__ __device__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct Params {
    const float* a_hi;   // canonical layout, A_BYTES
    const float* a_lo;
    const float* b_hi;   // canonical layout, B_BYTES
    const float* b_lo;
    float* d_out;        // [M][N] row-major result of the first iteration (CTA 0)
    long long* cycles;   // per-CTA cycles of the timed loop
    int iters;
    int epilogue;        // 0: MMA only; 1: TMEM read-out + twiddle + hi/lo split + re-store of the B tile every iteration
};

__global__ void __launch_bounds__(128, 1) k_dft_gemm(const Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sAh = smem;
    unsigned char* sAl = sAh + A_BYTES;
    unsigned char* sBh = sAl + A_BYTES;
    unsigned char* sBl = sBh + B_BYTES;
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    // operands -> shared memory (already in canonical order: plain copies)
    for (int i = tid; i < A_BYTES / 16; i += 128) {
        reinterpret_cast<uint4*>(sAh)[i] = reinterpret_cast<const uint4*>(p.a_hi)[i];
        reinterpret_cast<uint4*>(sAl)[i] = reinterpret_cast<const uint4*>(p.a_lo)[i];
    }
    for (int i = tid; i < B_BYTES / 16; i += 128) {
        reinterpret_cast<uint4*>(sBh)[i] = reinterpret_cast<const uint4*>(p.b_hi)[i];
        reinterpret_cast<uint4*>(sBl)[i] = reinterpret_cast<const uint4*>(p.b_lo)[i];
    }
    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy operand stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;
    const uint32_t idesc = make_idesc();
    uint32_t parity = 0;

    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
        if (tid == 0) {
            // D = Ah.Bh + Ah.Bl + Al.Bh : 3 x 16 MMAs of K = 8 (two K-cores = 256 bytes per step)
            const uint32_t a_addr[3] = {smem_u32(sAh), smem_u32(sAh), smem_u32(sAl)};
            const uint32_t b_addr[3] = {smem_u32(sBh), smem_u32(sBl), smem_u32(sBh)};
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int k = 0; k < K / UMMA_K; ++k)
                    mma_tf32(tmem, make_desc(a_addr[s] + k * 2 * LBO), make_desc(b_addr[s] + k * 2 * LBO), idesc, (s | k) != 0);
            umma_commit(&mbar);
        }
        mbar_wait(&mbar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.epilogue || it == 0) {
            // row m = tid of D lives in TMEM lane m: this warp reads its own 32 lanes, 64 columns in two chunks
            const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
            float v[2][32];
            tmem_ld32(lane_addr + 0, v[0]);
            tmem_ld32(lane_addr + 32, v[1]);
            if (it == 0 && blockIdx.x == 0 && p.d_out) {
#pragma unroll
                for (int c = 0; c < 64; ++c) p.d_out[tid * N + c] = v[c >> 5][c & 31];
            }
            if (p.epilogue) {
                // what a chained FFT stage would do with D: twiddle (complex multiply needs the partner row m^64,
                // emulated here by a rotation with constants), split into the tf32 high part and the remainder, and
                // store both as the next stage's B tile (row n, k = tid) in the canonical layout
                const float cw = 0.99518472667f, sw = 0.09801714033f;
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const float x = v[c >> 5][c & 31];
                    const float y = v[(c ^ 1) >> 5][(c ^ 1) & 31];
                    const float w = x * cw - y * sw;
                    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
                    const float lo = w - hi;
                    const uint32_t off = canon(c, tid);
                    *reinterpret_cast<float*>(sBh + off) = hi * 0.015625f;     // keep magnitudes bounded over the iterations
                    *reinterpret_cast<float*>(sBl + off) = lo * 0.015625f;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                              // TMEM drained and B re-written before the next MMAs
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long t1 = clock64();
    if (tid == 0) p.cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    // A: real form of the 64-point DFT matrix; B: one tile of data, 64 complex columns x 64 rows (cu8-like values)
    std::vector<double> A((size_t)M * K), B((size_t)K * N);
    for (int k = 0; k < 64; ++k)
        for (int n = 0; n < 64; ++n) {
            const double th = 2.0 * M_PI * (double)((k * n) % 64) / 64.0, c = cos(th), s = sin(th);
            A[(size_t)k * K + n] = c;            A[(size_t)k * K + 64 + n] = s;
            A[(size_t)(64 + k) * K + n] = -s;    A[(size_t)(64 + k) * K + 64 + n] = c;
        }
    srand(7);
    for (size_t i = 0; i < B.size(); ++i) B[i] = (double)(rand() % 256) - 127.5 + 0.37 * (double)(rand() % 1000) / 1000.0;
    std::vector<float> ah(M * K), al(M * K), bh(N * K), bl(N * K);
    auto split = [](double x, float& hi, float& lo) {
        const float f = (float)x;
        uint32_t u;
        memcpy(&u, &f, 4);
        u &= 0xFFFFE000u;
        memcpy(&hi, &u, 4);
        lo = f - hi;
    };
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) split(A[(size_t)m * K + k], ah[canon(m, k) / 4], al[canon(m, k) / 4]);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) split(B[(size_t)k * N + n], bh[canon(n, k) / 4], bl[canon(n, k) / 4]);

    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d_ah, *d_al, *d_bh, *d_bl, *d_out;
    long long* d_cyc;
    CK(cudaMalloc(&d_ah, A_BYTES)); CK(cudaMalloc(&d_al, A_BYTES)); CK(cudaMalloc(&d_bh, B_BYTES)); CK(cudaMalloc(&d_bl, B_BYTES));
    CK(cudaMalloc(&d_out, M * N * 4)); CK(cudaMalloc(&d_cyc, sms * 8));
    CK(cudaMemcpy(d_ah, ah.data(), A_BYTES, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_al, al.data(), A_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bh, bh.data(), B_BYTES, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_bl, bl.data(), B_BYTES, cudaMemcpyHostToDevice));
    const size_t smem = 2 * A_BYTES + 2 * B_BYTES + 1024;
    CK(cudaFuncSetAttribute((const void*)k_dft_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    Params p{d_ah, d_al, d_bh, d_bl, d_out, d_cyc, 1, 0};
    k_dft_gemm<<<1, 128, smem>>>(p);                       // numerics: one tile, one iteration
    CK(cudaDeviceSynchronize());
    std::vector<float> out(M * N);
    CK(cudaMemcpy(out.data(), d_out, M * N * 4, cudaMemcpyDeviceToHost));
    double err2 = 0, ref2 = 0, errmax = 0, err2_1x = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0, acc1 = 0;
            for (int k = 0; k < K; ++k) {
                acc += A[(size_t)m * K + k] * B[(size_t)k * N + n];
                acc1 += (double)ah[canon(m, k) / 4] * (double)bh[canon(n, k) / 4];      // what a single TF32 product computes
            }
            const double e = (double)out[m * N + n] - acc;
            err2 += e * e; ref2 += acc * acc; errmax = fmax(errmax, fabs(e));
            err2_1x += (acc1 - acc) * (acc1 - acc);
        }
    const double rel = sqrt(err2 / ref2), rel1 = sqrt(err2_1x / ref2);

    double ms[2] = {0, 0}, cyc[2] = {0, 0};
    for (int epi = 0; epi < 2; ++epi) {
        Params q{d_ah, d_al, d_bh, d_bl, nullptr, d_cyc, iters, epi};
        k_dft_gemm<<<sms, 128, smem>>>(q);                 // warm-up
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_dft_gemm<<<sms, 128, smem>>>(q);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        ms[epi] = t;
        std::vector<long long> c(sms);
        CK(cudaMemcpy(c.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost));
        double s = 0;
        for (int i = 0; i < sms; ++i) s += (double)c[i];
        cyc[epi] = s / sms / iters;
    }
    // one radix-64 stage over the Welch workload (1000 segments x 65536 points): points / (4096 points per tile)
    const double tiles = 1000.0 * 65536.0 / 4096.0;
    const double us_stage[2] = {ms[0] * 1e3 / ((double)iters * sms) * tiles, ms[1] * 1e3 / ((double)iters * sms) * tiles};
    printf("{\"proto\": \"tcgen05 kind::tf32 radix-64 DFT stage as GEMM, 3xTF32, M=128 N=64 K=128, operands resident in shared memory\", "
           "\"sms\": %d, \"iters_per_cta\": %d, \"rel_l2_error_3xtf32_vs_f64\": %.3e, \"max_abs_error\": %.3e, "
           "\"rel_l2_error_single_tf32_product\": %.3e, "
           "\"cycles_per_tile_mma_only\": %.1f, \"cycles_per_tile_with_epilogue\": %.1f, "
           "\"ms_total_mma_only\": %.4f, \"ms_total_with_epilogue\": %.4f, "
           "\"us_per_radix64_stage_of_welch_cfg2_mma_only\": %.1f, \"us_per_radix64_stage_of_welch_cfg2_with_epilogue\": %.1f, "
           "\"stages_for_64k\": 2.667}\n",
           sms, iters, rel, errmax, rel1, cyc[0], cyc[1], ms[0], ms[1], us_stage[0], us_stage[1]);
    return 0;
}
