// tools/microbench/dfma.cu -- is the FP64 pipe of B200 (sm_100a) a usable second multiplier for 254-bit
// Montgomery arithmetic?  Every field kernel of this library is pinned at the IMAD.WIDE issue rate
// (tools/microbench/widemad.cu: 9.27e12/s).  A double-precision FMA delivers a 53x53-bit product in two
// instructions (Emmart's hi/lo split), so IF the DFMA pipe (a) issues at or above the wide-MAD rate and
// (b) co-issues with the integer pipe, a 52-bit-limb multiplier could share the work.  Measured here:
//   mode 0  DFMA alone, 8 independent data-dependent chains per thread
//   mode 1  IMAD.WIDE.U32 alone (same harness; cross-check against widemad.cu)
//   mode 2  DFMA and IMAD.WIDE interleaved 1:1 (co-issue: time vs the slower of the two alone)
//   mode 3  the Emmart product step: hi = fma(a,b,c1) (round-to-zero), lo = fma(a,b,c2 - hi), then both
//           halves moved to the integer side and accumulated with 64-bit adds (what a 52-bit-limb
//           multiplier executes per limb product: 2 DFMA + 1 DADD + 2 x 64-bit integer add)
//   mode 4  DMUL alone; mode 5 DADD alone (do the three FP64 ops share one pipe at one rate?)
//   k_mix<NW, NL>  the hybrid question: per step NW carry-chained wide MADs (a CIOS row is 8 of them) next to NL Emmart limb
//           products, reported as 254-bit Montgomery products per second if a product were made of 128 wide MADs (the integer
//           multiplier of fp.cuh) or 55 limb products (5 x 5 + 5 x (5 + 1) of a 52-bit-limb multiplier):
//           equivalent = wide_mads / 128 + limb_products / 55.  NL = 0 is the integer multiplier alone.
// The SASS of every mode is checked with cuobjdump (DFMA / IMAD.WIDE.U32 counts per loop body).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma tools/microbench/dfma.cu && /tmp/dfma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 8

template <int MODE>
__global__ void k_dfma(double* out, double a, double b, uint32_t ia, int iters) {
    double x[CHAINS];
    uint32_t p[CHAINS], q[CHAINS];
    unsigned long long acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        x[c] = a + (double)(threadIdx.x * 8 + c) * 1e-9;
        p[c] = threadIdx.x * 7 + c + ia; q[c] = threadIdx.x * 3 + c + 11;
        acc[c] = c;
    }
    const double c1 = 4503599627370496.0 * 4503599627370496.0;        // 2^104
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0 || MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(b), "d"(a));
                if (MODE == 1 || MODE == 2)
                    asm volatile("{ .reg .u64 w; mul.wide.u32 w, %0, %1; mov.b64 {%0, %1}, w; }" : "+r"(p[c]), "+r"(q[c]));
                if (MODE == 3) {
                    double hi, lo, sub;
                    asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(x[c]), "d"(b), "d"(c1));
                    asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(sub) : "d"(c1), "d"(hi));
                    asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(lo) : "d"(x[c]), "d"(b), "d"(sub));
                    acc[c] += (unsigned long long)__double_as_longlong(hi);
                    acc[c] += (unsigned long long)__double_as_longlong(lo);
                    x[c] = __longlong_as_double((long long)((acc[c] & 0x000fffffffffffffull) | 0x4330000000000000ull));   // keep a dependence, stay finite
                }
                if (MODE == 4) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(b));
                if (MODE == 5) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + (double)(p[c] ^ q[c]) + (double)acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NW, int NL>
__global__ void k_mix(double* out, double a, double b, uint32_t ia, int iters) {
    uint32_t lo[8], hi[8], m = ia | 1u;
    double x[4];
    unsigned long long acc[4];
#pragma unroll
    for (int c = 0; c < 8; ++c) { lo[c] = threadIdx.x * 7 + c + ia; hi[c] = threadIdx.x * 3 + c + 11; }
#pragma unroll
    for (int c = 0; c < 4; ++c) { x[c] = a + (double)(threadIdx.x * 4 + c) * 1e-9; acc[c] = c; }
    const double c1 = 4503599627370496.0 * 4503599627370496.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // NW wide MADs as carry chains of 8 (one CIOS row each)
#pragma unroll
            for (int r = 0; r < NW / 8; ++r) {
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(m), "r"(hi[7]));
#pragma unroll
                for (int c = 1; c < 8; ++c)
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(m), "r"(hi[c - 1]));
                m = lo[7] ^ hi[3];
            }
#pragma unroll
            for (int q = 0; q < NL; ++q) {
                const int c = q & 3;
                double h, l, sub;
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(h) : "d"(x[c]), "d"(b), "d"(c1));
                asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(sub) : "d"(c1), "d"(h));
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(l) : "d"(x[c]), "d"(b), "d"(sub));
                acc[c] += (unsigned long long)__double_as_longlong(h);
                acc[c] += (unsigned long long)__double_as_longlong(l);
                x[c] = __longlong_as_double((long long)((acc[c] & 0x000fffffffffffffull) | 0x4330000000000000ull));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += (double)(lo[c] ^ hi[c]);
#pragma unroll
    for (int c = 0; c < 4; ++c) s += x[c] + (double)acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)m;
}

template <class K>
static float time_ms(K launch, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 500;
    double* d; cudaMalloc(&d, (size_t)blocks * threads * 8);
    double base = (double)blocks * threads * iters * (double)UNROLL * CHAINS;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
    float ms;
    ms = time_ms([&] { k_dfma<0><<<blocks, threads>>>(d, 1.0000001, 0.9999999, 3, iters); }); printf(", \"dfma_alone_Tops\": %.3f", base / ms / 1e9);
    float ms_d = ms;
    ms = time_ms([&] { k_dfma<1><<<blocks, threads>>>(d, 1.0000001, 0.9999999, 3, iters); }); printf(", \"imad_wide_alone_Tops\": %.3f", base / ms / 1e9);
    float ms_i = ms;
    ms = time_ms([&] { k_dfma<2><<<blocks, threads>>>(d, 1.0000001, 0.9999999, 3, iters); });
    printf(", \"interleaved_pairs_Tops\": %.3f, \"interleaved_ms\": %.4f, \"dfma_alone_ms\": %.4f, \"imad_wide_alone_ms\": %.4f, \"coissue_overlap\": %.3f",
           base / ms / 1e9, ms, ms_d, ms_i, (ms_d + ms_i - ms) / (ms_d < ms_i ? ms_d : ms_i));
    ms = time_ms([&] { k_dfma<3><<<blocks, threads>>>(d, 1.0000001, 0.9999999, 3, iters); }); printf(", \"emmart_limb_products_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_dfma<4><<<blocks, threads>>>(d, 1.0000001, 0.9999999, 3, iters); }); printf(", \"dmul_alone_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_dfma<5><<<blocks, threads>>>(d, 1.0000001, 0.9999999, 3, iters); }); printf(", \"dadd_alone_Tops\": %.3f", base / ms / 1e9);
    {
        const double steps = (double)blocks * threads * iters * 4.0;
#define MIX(NW, NL)                                                                                              \
        ms = time_ms([&] { k_mix<NW, NL><<<blocks, threads>>>(d, 1.0000001, 0.9999999, 3, iters); });               \
        printf(", \"mix_%dw_%dl_Gmodmul_equiv\": %.2f", NW, NL, steps * ((NW) / 128.0 + (NL) / 55.0) / ms / 1e6);
        MIX(32, 0) MIX(32, 1) MIX(32, 2) MIX(32, 4) MIX(32, 8) MIX(16, 8) MIX(0, 8)
#undef MIX
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf(", \"cuda_status\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
