// tests/host_emu/emu.cpp -- compiles the product's device arithmetic headers for the HOST
// (ZKG_HOST_EMU: PTX carry-chain primitives emulated in C++), so the exact limb schedules used
// by the CUDA kernels can be checked against the oracle on a machine without a GPU.
// Test infrastructure only.
#define ZKG_HOST_EMU 1
#include <cstring>
#include "../../zk-saas_b200/csrc/fp.cuh"

using namespace zkg;

template <class F>
static F ld(const uint64_t* p) { F r; memcpy(r.v, p, 32); return r; }
template <class F>
static void st(uint64_t* p, const F& a) { memcpy(p, a.v, 32); }

#define FIELD_API(NAME, F)                                                                         \
    extern "C" void emu_##NAME##_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_mul(ld<F>(a + 4 * i), ld<F>(b + 4 * i)));  \
    }                                                                                              \
    extern "C" void emu_##NAME##_add(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_add(ld<F>(a + 4 * i), ld<F>(b + 4 * i)));  \
    }                                                                                              \
    extern "C" void emu_##NAME##_sub(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_sub(ld<F>(a + 4 * i), ld<F>(b + 4 * i)));  \
    }                                                                                              \
    extern "C" void emu_##NAME##_neg(const uint64_t* a, uint64_t* o, size_t n) {                   \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_neg(ld<F>(a + 4 * i)));                    \
    }                                                                                              \
    extern "C" void emu_##NAME##_inv(const uint64_t* a, uint64_t* o, size_t n) {                   \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_inv(ld<F>(a + 4 * i)));                    \
    }                                                                                              \
    extern "C" void emu_##NAME##_from_mont(const uint64_t* a, uint64_t* o, size_t n) {             \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_from_mont(ld<F>(a + 4 * i)));              \
    }                                                                                              \
    extern "C" void emu_##NAME##_to_mont(const uint64_t* a, uint64_t* o, size_t n) {               \
        for (size_t i = 0; i < n; ++i) st(o + 4 * i, fp_to_mont(ld<F>(a + 4 * i)));                \
    }

FIELD_API(fr, Fr)
FIELD_API(fq, Fq)
