// tools/microbench/coissue.cu -- can the integer MAD pipe (IMAD / IMAD.WIDE) and the ALU pipe
// (IADD3 / LOP3 / SHF) be fed in the same cycle window on sm_100?  Decides how the big-integer
// multiplier should split work between multiplies and carry handling.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[8], y[8];
    uint64_t w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { x[c] = threadIdx.x + c; y[c] = threadIdx.x * 7 + c; w[c] = threadIdx.x * 3 + c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (MODE == 0 || MODE == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
                if (MODE == 1 || MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[c]) : "r"(a), "r"(x[(c + 1) & 7]));
                if (MODE == 3 || MODE == 4) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a), "r"(y[c]));
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[c]) : "r"(a), "r"(b));
                if (MODE == 5) {   // 64-bit add: IADD3 + IADD3.X
                    uint32_t* p = reinterpret_cast<uint32_t*>(&w[c]);
                    asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(p[0]), "+r"(p[1]) : "r"(x[c]), "r"(y[c]));
                }
                if (MODE == 6) {   // mul.wide (no addend) + 64-bit add: the shape ptxas prefers
                    uint64_t t;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(y[c]));
                    uint32_t* p = reinterpret_cast<uint32_t*>(&w[c]);
                    uint32_t* q = reinterpret_cast<uint32_t*>(&t);
                    asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(p[0]), "+r"(p[1]) : "r"(q[0]), "r"(q[1]));
                }
                if (MODE == 7) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(y[c]) : "r"(x[c]));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c] + y[c] + (uint32_t)w[c] + (uint32_t)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class L>
static float best_ms(L launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2000;
    uint32_t* d; cudaMalloc(&d, (size_t)blocks * threads * 4);
    double slots = (double)blocks * threads * iters * 64.0;     // 64 (groups of) instructions per iteration per thread
    const char* names[8] = {"imad_only", "lop3_only", "imad_plus_lop3", "imadwide_only", "imadwide_plus_lop3", "add64 (iadd3+iadd3.x)", "mulwide_plus_add64", "shf_only"};
    printf("{");
#define RUN(M) { float ms = best_ms([&] { k<M><<<blocks, threads>>>(d, 3, 5, iters); }); printf("%s\"%s_Tgroups_per_s\": %.3f", M ? ", " : "", names[M], slots / ms / 1e9); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    printf(", \"status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
