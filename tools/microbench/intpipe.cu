// tools/microbench/intpipe.cu -- measures the integer-pipe ceilings the MSM/NTT rooflines are quoted
// against (32-bit IMAD, IMAD.HI, IMAD.WIDE, IADD3) and the throughput of this repo's Montgomery
// multiplier, on whatever GPU it runs on.  Build + run (on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I zk-saas_b200/csrc -o /tmp/intpipe tools/microbench/intpipe.cu
//   /tmp/intpipe > gpurun_out/intpipe.json
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace zkg;

#define CHAINS 8
template <int MODE>
__global__ void k_int(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[CHAINS];
    uint64_t w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; w[c] = threadIdx.x * 3 + c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
                if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
                // data-dependent forms (a loop-invariant product or addend is hoisted / strength-reduced by ptxas:
                // the first version of this file measured 64-bit adds here and called them IMAD.WIDE)
                if (MODE == 2) asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, %1; mov.b64 {%0, %1}, t; }" : "+r"(x[c]), "+r"(x[(c + 1) % CHAINS]));
                if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]));
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
                // 3-input adds that cannot be strength-reduced (each chain mixes its two neighbours)
                if (MODE == 6) asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t; }" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]), "r"(x[(c + 3) % CHAINS]));
                // 32x32+64 with carry-in/out (what ptxas turns into IMAD.WIDE.U32.X): 2 chains of 4 per step
                if (MODE == 5 && (c & 3) == 0) {
                    uint32_t* lo = reinterpret_cast<uint32_t*>(&w[c]);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(lo[1]) : "r"(a), "r"(x[c]));
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[2]), "+r"(lo[3]) : "r"(a), "r"(x[c]));
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[4]), "+r"(lo[5]) : "r"(a), "r"(x[c]));
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[6]), "+r"(lo[7]) : "r"(a), "r"(x[c]));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + (uint32_t)w[c] + (uint32_t)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F, int ILP, int VARIANT>
__global__ void k_modmul(F* io, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    F x[ILP], y = io[tid];
#pragma unroll
    for (int c = 0; c < ILP; ++c) { x[c] = y; x[c].v[0] ^= c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < ILP; ++c) x[c] = VARIANT == 0 ? fp_mul_r29(x[c], y) : VARIANT == 2 ? fp_sqr(x[c]) : fp_mul(x[c], y);
    }
    F s = x[0];
#pragma unroll
    for (int c = 1; c < ILP; ++c) s = fp_add(s, x[c]);
    io[tid] = s;
}

template <class F>
__global__ void k_modadd(F* io, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    F x = io[tid], y = x;
    y.v[1] ^= 5;
    for (int it = 0; it < iters; ++it) { x = fp_add(x, y); y = fp_sub(y, x); }
    io[tid] = fp_add(x, y);
}

template <class K>
static float time_ms(K launch, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int blocks = sms * 8, threads = 256;
    uint32_t* d; cudaMalloc(&d, (size_t)blocks * threads * 64);
    cudaMemset(d, 1, (size_t)blocks * threads * 64);
    int iters = 2000;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
    const char* names[7] = {"imad_lo", "imad_hi", "imad_wide", "iadd", "lop3", "imad_wide_carry", "iadd3_mixed"};
    double ops = (double)blocks * threads * iters * 8.0 * CHAINS;
    float ms;
    ms = time_ms([&] { k_int<0><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"%s_Tops\": %.3f", names[0], ops / ms / 1e9);
    ms = time_ms([&] { k_int<1><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"%s_Tops\": %.3f", names[1], ops / ms / 1e9);
    ms = time_ms([&] { k_int<2><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"%s_Tops\": %.3f", names[2], ops / ms / 1e9);
    ms = time_ms([&] { k_int<3><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"%s_Tops\": %.3f", names[3], ops / ms / 1e9);
    ms = time_ms([&] { k_int<4><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"%s_Tops\": %.3f", names[4], ops / ms / 1e9);
    ms = time_ms([&] { k_int<5><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"%s_Tops\": %.3f", names[5], ops / ms / 1e9);
    ms = time_ms([&] { k_int<6><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"%s_Tops\": %.3f", names[6], ops / ms / 1e9);
    int mi = 400;
    for (int th : {128, 256, 512}) {
        int bl = sms * (2048 / th);
        double mm = (double)bl * th * mi;
        ms = time_ms([&] { k_modmul<Fq, 1, 0><<<bl, th>>>((Fq*)d, mi); }); printf(", \"modmul29_ilp1_t%d_G\": %.2f", th, mm / ms / 1e6);
        ms = time_ms([&] { k_modmul<Fq, 2, 0><<<bl, th>>>((Fq*)d, mi); }); printf(", \"modmul29_ilp2_t%d_G\": %.2f", th, 2 * mm / ms / 1e6);
        ms = time_ms([&] { k_modmul<Fq, 1, 1><<<bl, th>>>((Fq*)d, mi); }); printf(", \"modmul_cios_ilp1_t%d_G\": %.2f", th, mm / ms / 1e6);
        ms = time_ms([&] { k_modmul<Fq, 1, 2><<<bl, th>>>((Fq*)d, mi); }); printf(", \"modsqr_ilp1_t%d_G\": %.2f", th, mm / ms / 1e6);
    }
    {
        int th = 256, bl = sms * 8;
        ms = time_ms([&] { k_modadd<Fq><<<bl, th>>>((Fq*)d, 2000); });
        printf(", \"modaddsub_G\": %.2f", (double)bl * th * 2000 * 2 / ms / 1e6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf(", \"cuda_status\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
