// tools/microbench/widemad.cu -- issue rate of the 32x32->64 multiply forms on sm_100a, measured with
// data-dependent chains that ptxas cannot hoist or strength-reduce (the SASS of every mode is checked
// with cuobjdump before the numbers are trusted; see profiles/).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/widemad tools/microbench/widemad.cu && /tmp/widemad
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 8

__device__ __forceinline__ void mulw(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mul.wide.u32 {%0, %1}, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

template <int MODE>
__global__ void k_wide(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[CHAINS], y[CHAINS], z[CHAINS], t[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x * 7 + c + a; y[c] = threadIdx.x * 3 + c + b; z[c] = x[c] ^ y[c]; t[c] = z[c] + 1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (MODE == 0) {          // IMAD.WIDE.U32 Rd, Ra, Rb, RZ   (both halves feed the next product)
                    asm volatile("{ .reg .u64 w; mul.wide.u32 w, %0, %1; mov.b64 {%0, %1}, w; }" : "+r"(x[c]), "+r"(y[c]));
                }
                if (MODE == 1) {          // mad.wide with a live 64-bit addend: what does ptxas emit?
                    asm volatile("{ .reg .u64 w; mov.b64 w, {%0, %1}; mad.wide.u32 w, %0, %2, w; mov.b64 {%0, %1}, w; }" : "+r"(x[c]), "+r"(y[c]) : "r"(a));
                }
                if (MODE == 2) {          // immediate multiplicand, RZ addend (+ one LOP3 on the other pipe to keep both halves live)
                    x[c] ^= y[c];
                    asm volatile("{ .reg .u64 w; mul.wide.u32 w, %0, 0x10460b6; mov.b64 {%0, %1}, w; }" : "+r"(x[c]), "+r"(y[c]));
                }
                if (MODE == 3) {          // product + 3-input carry adds on the ALU pipe: acc(z,t) += x*y ; x,y <- rotate
                    uint32_t lo, hi;
                    asm volatile("{ .reg .u64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(lo), "=r"(hi) : "r"(x[c]), "r"(y[c]));
                    asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(z[c]), "+r"(t[c]) : "r"(lo), "r"(hi));
                    x[c] = t[c]; y[c] = z[c];       // register renaming only
                }
                if (MODE == 4 && (c & 3) == 0) {     // the CIOS row: one carry chain of 4 wide MADs (IMAD.WIDE.U32.X)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[c]), "+r"(y[c]) : "r"(a), "r"(z[c]));
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[c + 1]), "+r"(y[c + 1]) : "r"(a), "r"(z[c]));
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[c + 2]), "+r"(y[c + 2]) : "r"(a), "r"(z[c]));
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[c + 3]), "+r"(y[c + 3]) : "r"(a), "r"(z[c]));
                    z[c] = y[c + 3];
                }
                if (MODE == 5) {          // 32-bit IMAD (lo) with data dependence, for the full-rate reference
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y[c]), "r"(a));
                }
                if (MODE == 6) {          // IMAD.HI with data dependence
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y[c]), "r"(a));
                }
                if (MODE == 7) {          // two products folded by ONE 3-input carry add pair (the r29 inner pattern)
                    uint32_t l0, h0, l1, h1;
                    asm volatile("{ .reg .u64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(l0), "=r"(h0) : "r"(x[c]), "r"(y[c]));
                    asm volatile("{ .reg .u64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(l1), "=r"(h1) : "r"(y[c]), "r"(a));
                    asm volatile("{ .reg .u64 p, q, r; mov.b64 p, {%2, %3}; mov.b64 q, {%4, %5}; mov.b64 r, {%0, %1}; add.u64 r, r, p; add.u64 r, r, q; mov.b64 {%0, %1}, r; }"
                                 : "+r"(z[c]), "+r"(t[c]) : "r"(l0), "r"(h0), "r"(l1), "r"(h1));
                    x[c] = t[c]; y[c] = z[c];
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + y[c] + z[c] + t[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 2000;
    uint32_t* d; cudaMalloc(&d, (size_t)blocks * threads * 4);
    double base = (double)blocks * threads * iters * (double)UNROLL * CHAINS;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
    float ms;
    ms = time_ms([&] { k_wide<0><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"imad_wide_rz_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_wide<1><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"mad_wide_addend_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_wide<2><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"imad_wide_imm_plus_lop3_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_wide<3><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"wide_rz_plus_add64_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_wide<4><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"imad_wide_x_chain_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_wide<5><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"imad_lo_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_wide<6><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"imad_hi_Tops\": %.3f", base / ms / 1e9);
    ms = time_ms([&] { k_wide<7><<<blocks, threads>>>(d, 3, 5, iters); }); printf(", \"two_wide_rz_plus_iadd3_pair_Tprod\": %.3f", 2 * base / ms / 1e9);
    cudaError_t e = cudaDeviceSynchronize();
    printf(", \"cuda_status\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
