// ntt.cu -- BN254 Fr transforms of the d_fft / d_ifft path and the PackedSharingParams maps.
//
// Replaces (reference file:line)
//   fft1_in_place                     dist-primitives/src/dfft/mod.rs:178-208   (clients)
//   king closure of fft2_with_rearrange dist-primitives/src/dfft/mod.rs:264-304 (king)
//     = transpose + unpack_missing_shares (:265-274) + fft2_in_place (:210-237) + distribute_powers
//       (:278-280) + fft_in_place_rearrange (:322-335) + pack / pack_vec (:287-302, utils/pack.rs:8-35)
//   king closure of deg_red           dist-primitives/src/utils/deg_red.rs:103-111
//   pack / det_pack / unpack / unpack2 secret-sharing/src/pss.rs:69-166
//
// fft1 closed form (checked against the literal loops in oracle/): with N = m/l, y = the share
// vector read in bit-reversed order and w_N = gen^l,
//      fft1(px)[k] = X[(k+1) mod N],   X = DFT_N(y; w_N).
// The reference's "+1" twiddle convention is exactly that cyclic shift, so the kernel is an
// ordinary bit-reversed-input / natural-output NTT whose final store is shifted by one slot.
// Large N runs as a multi-pass four-step: every pass loads a tile into shared memory (limb-major,
// bank-conflict free), multiplies by one on-the-fly inter-pass twiddle w^(klow * j) composed from a
// two-level power table, and runs up to 10 radix-2 stages out of shared memory.
//
// fft2 closed form: the log2(l) stages only ever combine the l secrets of one share column k, and
// send them to positions (k + q*m/l + 1) mod m; so the whole king pipeline is column-local up to
// one permutation, which the first kernel applies while storing its results in pack order.
#include "common.cuh"
#include "fp.cuh"
#include "host_fr.hpp"
#include <deque>
#include <string>
#include <thread>
#include <vector>

namespace zkg {

using host::HFr;

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    Fr r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(s), b = __ldg(s + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr ld_fr_rw(const Fr* p) {
    Fr r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4 a = s[0], b = s[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& r) {
    uint4* d = reinterpret_cast<uint4*>(p);
    d[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    d[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

static int env_int_ntt(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

struct FrArg { uint32_t v[8]; };       // Fr passed by value as a kernel argument
static FrArg to_arg(const HFr& h) { FrArg a; memcpy(a.v, h.v, 32); return a; }
__device__ __forceinline__ Fr from_arg(const FrArg& a) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = a.v[i];
    return r;
}

// out[i] = base^i, i < count  (square-and-multiply per thread; count is a few thousand)
__global__ void k_pow_table(FrArg base_, uint32_t count, Fr* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr base = from_arg(base_), acc = Fr::one();
    for (int b = 31 - __clz(i | 1); b >= 0; --b) {
        acc = fp_sqr(acc);
        if ((i >> b) & 1) acc = fp_mul(acc, base);
    }
    st_fr(out + i, acc);
}
// t[i] *= mult  (turns a power table w^i into c * w^i)
__global__ void k_scale_table(Fr* __restrict__ t, uint32_t count, FrArg mult_) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    st_fr(t + i, fp_mul(ld_fr_rw(t + i), from_arg(mult_)));
}

// Two-level power table of a root w: w^e = lo[e & (LO-1)] * hi[e >> LO_BITS]
static constexpr int TW_LO_BITS = 12;
static constexpr uint32_t TW_LO = 1u << TW_LO_BITS;
struct PowTable { const Fr* lo; const Fr* hi; };
__device__ __forceinline__ Fr pow_lookup(const PowTable& t, uint64_t e) {
    Fr a = ld_fr(t.lo + (uint32_t)(e & (TW_LO - 1)));
    uint64_t h = e >> TW_LO_BITS;
    if (h) a = fp_mul(a, ld_fr(t.hi + h));
    return a;
}

__device__ __forceinline__ uint32_t bitrev32(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// ------------------------------------------------------------------------------------------
// One NTT pass.  The transform index is split as  p = upper * 2^(s+b) + row * 2^s + klow; a block
// owns one `upper`, CW consecutive klow ("columns") and all 2^b rows, and performs CW independent
// size-2^b DFTs (bit-reversed rows in, natural rows out) after multiplying element (row, klow) by
// rho^(klow * bitrev_b(row)), rho = w_N^(N / 2^(s+b)).
// ------------------------------------------------------------------------------------------
// Inter-pass twiddles of one pass as a table in tile order: out[(row << s) + klow] = rho^(klow * bitrev_b(row)),
// rho = w_N^(N / 2^(s+b)).  Built once per (root, pass shape) and cached with the other parameter tables.
__global__ void k_ntt_twiddle_table(Fr* __restrict__ out, int s, int b, int rho_shift, PowTable tw) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ((size_t)1 << (s + b))) return;
    uint32_t row = (uint32_t)(idx >> s), klow = (uint32_t)(idx & (((size_t)1 << s) - 1));
    uint64_t e = (uint64_t)klow * bitrev32(row, b);
    st_fr(out + idx, e ? pow_lookup(tw, e << rho_shift) : Fr::one());
}

struct NttPass {
    int logN, s, b, cw_log;
    int first, last, shift;          // shift: store X[k] at (k-1) mod N (fft1 convention)
    int has_scale;
    FrArg scale;                     // pre-multiplier applied on the first pass (d_ifft's size_inv)
};

__global__ void __launch_bounds__(512, 2)
k_ntt_pass(const Fr* __restrict__ in, Fr* __restrict__ out, NttPass P, const Fr* __restrict__ tw_small, PowTable tw,
           const Fr* __restrict__ tw_full, const Fr* __restrict__ mask) {
    extern __shared__ uint32_t smem[];
    const uint32_t T = 1u << P.b, CW = 1u << P.cw_log, E = T * CW;   // elements per block
    uint32_t* S = smem;                       // 8 limb planes of E words
    uint32_t* TWS = smem + 8 * E;             // 8 limb planes of T/2 words (small twiddles w_T^i)
    const uint32_t cols_per_upper = (1u << P.s) >> P.cw_log;
    const uint32_t upper = blockIdx.x / cols_per_upper;
    const uint32_t klow0 = (blockIdx.x % cols_per_upper) << P.cw_log;
    const size_t base = ((size_t)upper << (P.s + P.b)) + klow0;
    const int rho_shift = P.logN - P.s - P.b;   // rho^e = w_N^(e << rho_shift)

    for (uint32_t i = threadIdx.x; i < T / 2; i += blockDim.x) {
        Fr t = ld_fr(tw_small + i);
#pragma unroll
        for (int l = 0; l < 8; ++l) TWS[l * (T / 2) + i] = t.v[l];
    }
    Fr scale = from_arg(P.scale);
    for (uint32_t idx = threadIdx.x; idx < E; idx += blockDim.x) {
        uint32_t col = idx & (CW - 1), row = idx >> P.cw_log;
        Fr x = ld_fr(in + base + ((size_t)row << P.s) + col);
        if (P.first && P.has_scale) x = fp_mul(x, scale);
        if (!P.first) {
            if (tw_full) {
                // inter-pass twiddle straight from the per-pass table, laid out like the tile ([row][klow]):
                // one coalesced 32-byte read instead of a second product on the binding (integer MAD) pipe
                if (row) x = fp_mul(x, ld_fr(tw_full + ((size_t)row << P.s) + klow0 + col));
            } else {
                uint64_t e = (uint64_t)(klow0 + col) * bitrev32(row, P.b);
                if (e) x = fp_mul(x, pow_lookup(tw, e << rho_shift));
            }
        }
#pragma unroll
        for (int l = 0; l < 8; ++l) S[l * E + idx] = x.v[l];
    }
    __syncthreads();

    for (int sg = 0; sg < P.b; ++sg) {
        const uint32_t half = 1u << sg;
        for (uint32_t u = threadIdx.x; u < E / 2; u += blockDim.x) {
            uint32_t col = u & (CW - 1), t = u >> P.cw_log;
            uint32_t kin = t & (half - 1);
            uint32_t i0 = (((t >> sg) << (sg + 1)) | kin) * CW + col;
            uint32_t i1 = i0 + half * CW;
            Fr x, y;
#pragma unroll
            for (int l = 0; l < 8; ++l) { x.v[l] = S[l * E + i0]; y.v[l] = S[l * E + i1]; }
            if (kin) {
                Fr w;
                uint32_t ti = kin << (P.b - 1 - sg);
#pragma unroll
                for (int l = 0; l < 8; ++l) w.v[l] = TWS[l * (T / 2) + ti];
                y = fp_mul(y, w);
            }
            Fr a = fp_add(x, y), d = fp_sub(x, y);
#pragma unroll
            for (int l = 0; l < 8; ++l) { S[l * E + i0] = a.v[l]; S[l * E + i1] = d.v[l]; }
        }
        __syncthreads();
    }

    const size_t N = (size_t)1 << P.logN;
    for (uint32_t idx = threadIdx.x; idx < E; idx += blockDim.x) {
        uint32_t col = idx & (CW - 1), row = idx >> P.cw_log;
        Fr x;
#pragma unroll
        for (int l = 0; l < 8; ++l) x.v[l] = S[l * E + idx];
        size_t k = base + ((size_t)row << P.s) + col;
        if (P.last && P.shift) k = (k + N - 1) & (N - 1);
        if (P.last && mask) x = fp_add(x, ld_fr(mask + k));
        st_fr(out + k, x);
    }
}

// ------------------------------------------------------------------------------------------
// The same pass with the radix-2 stages grouped three at a time in REGISTERS (radix-8 units): a
// thread owns 8 elements, runs up to three stages on them without touching shared memory, and the
// tile is exchanged through shared memory only between groups (two barriers per pass instead of
// b).  The first group loads straight from global memory (scale / inter-pass twiddle applied on the
// way in) and the last one stores straight to global memory (shift / mask applied on the way out).
// Shared layout: two planes of 16-byte half elements, one padding slot per 8 (stride-8 row access
// of the first group stays conflict free), then the small twiddles.  Units are numbered column-fastest,
// so a quarter warp always touches 8 consecutive slots.
//   sub_mode (pass 0, s = 0): the block owns CW consecutive sub-transforms of T contiguous elements;
//   "column" c is then the sub-transform: global index = base + c*T + row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ntt_slot(uint32_t idx) { return idx + (idx >> 3); }

template <int LE>          // elements per thread = 2^LE: 8 (radix-8 groups, 16 warps/SM) or 4 (radix-4 groups, 24 warps/SM)
__global__ void __launch_bounds__(256, LE == 3 ? 2 : 3)
k_ntt_pass8(const Fr* __restrict__ in, Fr* __restrict__ out, NttPass P, const Fr* __restrict__ tw_small, PowTable tw,
            const Fr* __restrict__ tw_full, const Fr* __restrict__ mask, int sub_mode) {
    extern __shared__ uint4 sm4[];
    const uint32_t T = 1u << P.b, CW = 1u << P.cw_log, E = T * CW;
    const uint32_t PL = E + (E >> 3) + 1;
    uint4* S0 = sm4;
    uint4* S1 = sm4 + PL;
    uint4* TW = sm4 + 2 * PL;                                 // entry i: TW[2i] (limbs 0..3), TW[2i+1] (limbs 4..7)
    size_t base;
    uint32_t klow0 = 0;
    if (sub_mode) {
        base = (size_t)blockIdx.x * E;
    } else {
        const uint32_t cols_per_upper = (1u << P.s) >> P.cw_log;
        const uint32_t upper = blockIdx.x / cols_per_upper;
        klow0 = (blockIdx.x % cols_per_upper) << P.cw_log;
        base = ((size_t)upper << (P.s + P.b)) + klow0;
    }
    const int rho_shift = P.logN - P.s - P.b;
    const size_t N = (size_t)1 << P.logN;
    constexpr int EPT = 1 << LE;
    const uint32_t units = E >> LE ? E >> LE : 1;             // threads that own elements
    const uint32_t uidx = threadIdx.x;

    for (uint32_t i = threadIdx.x; i < T / 2; i += blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(tw_small + i);
        TW[2 * i] = __ldg(src);
        TW[2 * i + 1] = __ldg(src + 1);
    }
    const Fr scale = from_arg(P.scale);

    // groups of stages: as even as possible, at most 3 each
    const int ngroups = (P.b + LE - 1) / LE ? (P.b + LE - 1) / LE : 1;
    int sg0 = 0;
    for (int gi = 0; gi < ngroups; ++gi) {
        const int G = (P.b - sg0 + (ngroups - gi) - 1) / (ngroups - gi);    // stages in this group (0 when b == 0)
        const bool from_global = gi == 0, to_global = gi == ngroups - 1;
        Fr v[EPT];
        uint32_t rowv[EPT], colv[EPT];
        const bool active = uidx < units && (E >= (uint32_t)EPT || uidx == 0);
        // element e of this thread: sub-butterfly q = e >> G, member j = e & (2^G - 1)
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const uint32_t q = (uint32_t)e >> G, j = (uint32_t)e & ((1u << G) - 1);
            const uint32_t beta = q * units + uidx;
            const uint32_t col = beta & (CW - 1), rest = beta >> P.cw_log;
            const uint32_t low = rest & ((1u << sg0) - 1), high = rest >> sg0;
            rowv[e] = (high << (sg0 + G)) | (j << sg0) | low;
            colv[e] = col;
        }
        if (active) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (E < (uint32_t)EPT && (uint32_t)e >= E) continue;
                const uint32_t row = rowv[e], col = colv[e];
                if (from_global) {
                    const size_t g = sub_mode ? base + (size_t)col * T + row : base + ((size_t)row << P.s) + col;
                    Fr x = ld_fr(in + g);
                    if (P.first && P.has_scale) x = fp_mul(x, scale);
                    if (!P.first && row) {
                        if (tw_full) x = fp_mul(x, ld_fr(tw_full + ((size_t)row << P.s) + klow0 + col));
                        else {
                            uint64_t ex = (uint64_t)(klow0 + col) * bitrev32(row, P.b);
                            if (ex) x = fp_mul(x, pow_lookup(tw, ex << rho_shift));
                        }
                    }
                    v[e] = x;
                } else {
                    const uint32_t sl = ntt_slot(row * CW + col);
                    uint4 a = S0[sl], b4 = S1[sl];
                    v[e].v[0] = a.x; v[e].v[1] = a.y; v[e].v[2] = a.z; v[e].v[3] = a.w;
                    v[e].v[4] = b4.x; v[e].v[5] = b4.y; v[e].v[6] = b4.z; v[e].v[7] = b4.w;
                }
            }
        }
        if (from_global) __syncthreads();                      // small twiddles are in place
        if (active) {
#pragma unroll
            for (int sp = 0; sp < LE; ++sp) {
                if (sp >= G) break;
                const int sigma = sg0 + sp;                      // global stage
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    if ((e >> sp) & 1) continue;
                    const int f = e + (1 << sp);
                    if (E < (uint32_t)EPT && (uint32_t)f >= E) continue;
                    const uint32_t kin = rowv[e] & ((1u << sigma) - 1);
                    Fr y = v[f];
                    if (kin) {
                        const uint32_t ti = kin << (P.b - 1 - sigma);
                        uint4 a = TW[2 * ti], b4 = TW[2 * ti + 1];
                        Fr w;
                        w.v[0] = a.x; w.v[1] = a.y; w.v[2] = a.z; w.v[3] = a.w;
                        w.v[4] = b4.x; w.v[5] = b4.y; w.v[6] = b4.z; w.v[7] = b4.w;
                        y = fp_mul(y, w);
                    }
                    const Fr x = v[e];
                    v[e] = fp_add(x, y);
                    v[f] = fp_sub(x, y);
                }
            }
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (E < (uint32_t)EPT && (uint32_t)e >= E) continue;
                const uint32_t row = rowv[e], col = colv[e];
                if (to_global) {
                    size_t k = sub_mode ? base + (size_t)col * T + row : base + ((size_t)row << P.s) + col;
                    if (P.last && P.shift) k = (k + N - 1) & (N - 1);
                    Fr x = v[e];
                    if (P.last && mask) x = fp_add(x, ld_fr(mask + k));
                    st_fr(out + k, x);
                } else {
                    const uint32_t sl = ntt_slot(row * CW + col);
                    S0[sl] = make_uint4(v[e].v[0], v[e].v[1], v[e].v[2], v[e].v[3]);
                    S1[sl] = make_uint4(v[e].v[4], v[e].v[5], v[e].v[6], v[e].v[7]);
                }
            }
        }
        if (!to_global) __syncthreads();
        sg0 += G;
    }
}

// ------------------------------------------------------------------------------------------
// Destination of the king's stage-1 scatter.  Unsharded: one buffer of m elements.  Sharded over G GPUs: the
// pack-order buffer is cut into G segments of 2^log_seg elements, segment r living in the memory of GPU r (its own
// output columns); p[r] is that segment as seen from THIS GPU -- a peer mapping over NVLink (cudaDeviceEnablePeerAccess
// in one process, a CUDA IPC mapping across processes).  Every slot is written by exactly one thread of one GPU, so the
// rotate / bit-reverse / stride permutation of dfft/mod.rs:284-300 IS the exchange: stage 1 stores straight into the
// owner's memory and no collective (and no zero-filled full-size buffer) is needed.
struct SDest {
    Fr* p[16];
    int log_seg;                  // 63 when unsharded (every index maps to p[0])
};
__device__ __forceinline__ Fr* sdest(const SDest& s, size_t dst) {
    return s.p[dst >> s.log_seg] + (dst & (((size_t)1 << s.log_seg) - 1));
}
static SDest sdest_single(Fr* S) {
    SDest d;
    for (int i = 0; i < 16; ++i) d.p[i] = S;
    d.log_seg = 63;
    return d;
}

// King, kernel 1: per share column k: secrets = U * shares (unpack2 or Lagrange matrix), the
// column-local fft2 butterflies, g^pos powers, and the store in *pack order*:
//   S[c*l + j] = j-th secret of output column c.
// mode 0: consecutive packing (pack_vec)          c = pos / l,          j = pos % l
// mode 1: rearrange (bit-reverse + stride m/l)    p = bitrev_m(pos):    c = p % (m/l), j = p / (m/l)
// mode 2: no fft2 at all (deg_red): S[k*l + j] = secrets[j]
// mode 3: fft2 only, natural positions, input read directly from `direct` ([k*l + j])
// ------------------------------------------------------------------------------------------
template <int LL>
__global__ void __launch_bounds__(256, 3)
k_king_stage1(const Fr* __restrict__ shares, uint32_t n_recv, const Fr* __restrict__ U, const Fr* __restrict__ direct,
              size_t mbyl, size_t col0, size_t cols, int log_m, int mode, PowTable gen_tw, int has_g, PowTable g_tw,
              SDest S) {
    // this launch owns the share columns [col0, col0 + cols) of the mbyl columns (cols == mbyl unsharded);
    // inputs are indexed locally (kk), twiddles and destinations by the global column k
    size_t kk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= cols) return;
    const size_t k = col0 + kk;
    constexpr int LOGL = LL == 2 ? 1 : LL == 4 ? 2 : 3;
    Fr v[LL];
    if (direct) {
#pragma unroll
        for (int j = 0; j < LL; ++j) v[j] = ld_fr(direct + kk * LL + j);
    } else {
#pragma unroll
        for (int j = 0; j < LL; ++j) v[j] = Fr::zero();
        // rows of U times the share column, four terms per Montgomery reduction (fp_dot)
        uint32_t r = 0;
        for (; r + 4 <= n_recv; r += 4) {
            Fr x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = ld_fr(shares + (size_t)(r + q) * cols + kk);
            // the matrix row is the `b` operand: one limb per CIOS row, read from the (L1-resident) table
            // as it is needed instead of being held in 32 registers
#pragma unroll
            for (int j = 0; j < LL; ++j) v[j] = fp_add(v[j], fp_dot<FrParams, 4>(x, U + (size_t)j * n_recv + r));
        }
        for (; r < n_recv; ++r) {
            Fr x = ld_fr(shares + (size_t)r * cols + kk);
#pragma unroll
            for (int j = 0; j < LL; ++j) v[j] = fp_add(v[j], fp_mul(ld_fr(U + (size_t)j * n_recv + r), x));
        }
    }
    if (mode == 2) {
#pragma unroll
        for (int j = 0; j < LL; ++j) st_fr(sdest(S, k * LL + j), v[j]);
        return;
    }
    const size_t m = (size_t)1 << log_m;
    // fft2: stage i = LOGL..1; before the stage there are C = m/2^i columns of E = 2^i entries.
    // The entries descending from share column k live in columns kap_q = k + q*mbyl; we keep them as
    // v[q*E + j] (q-th descendant column, j-th entry).
    size_t C = mbyl;
#pragma unroll
    for (int i = LOGL; i >= 1; --i) {
        const int E = 1 << i, Q = LL / E;            // Q descendant columns so far
        Fr nv[LL];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            size_t kap = k + (size_t)q * mbyl;
            Fr tw = pow_lookup(gen_tw, (((uint64_t)(kap + 1)) << (i - 1)) & (m - 1));
#pragma unroll
            for (int j = 0; j < E / 2; ++j) {
                Fr x = v[q * E + 2 * j];
                Fr y = fp_mul(v[q * E + 2 * j + 1], tw);
                // sums stay in column kap, differences go to column kap + C = k + (q + Q)*mbyl
                nv[q * (E / 2) + j] = fp_add(x, y);
                nv[(q + Q) * (E / 2) + j] = fp_sub(x, y);
            }
        }
#pragma unroll
        for (int j = 0; j < LL; ++j) v[j] = nv[j];
        C <<= 1;
    }
    // now v[q] is the single entry of column k + q*mbyl; rotate_right(1): position = column + 1 mod m
#pragma unroll
    for (int q = 0; q < LL; ++q) {
        size_t pos = (k + (size_t)q * mbyl + 1) & (m - 1);
        Fr x = v[q];
        if (has_g && pos) x = fp_mul(x, pow_lookup(g_tw, pos));
        size_t dst;
        if (mode == 1) {
            size_t p = (size_t)(__brevll((unsigned long long)pos) >> (64 - log_m));
            dst = (p & (mbyl - 1)) * LL + (p / mbyl);
        } else {
            dst = pos;        // mode 0: S[c*l + j] with c = pos / l, j = pos % l is S[pos]; mode 3 likewise
        }
        st_fr(sdest(S, dst), x);
    }
}

// ------------------------------------------------------------------------------------------
// King, kernel 1 specialised for l = 2 with all n = 8 shares present (the reference's one real
// configuration; anything else takes the generic kernel above).  Same results; fewer products:
// a thread owns T = k + 1 (the exponent of the fft2 twiddle AND the position of the column's first
// output), blocks are aligned on 256 values of T, so the HIGH factor of every two-level power lookup
// (exponent >> 12) is uniform over the block.  Those factors are folded into per-`hi` copies of the
// 2 x 8 unpack matrix (Us0 = U0 * Ghi, Us1 = U1 * Whi * Ghi; a cached table of m/2^13 x 16 elements built
// by k_king_scaled_matrices), after which a column needs only the LOW table entries:
//     v0 = <Us0, shares>, v1 = <Us1, shares>            (4 x fp_dot<4>)
//     y  = v1 * gen_lo[T & 4095]
//     out(pos = T)        = (v0 + y) * g_lo [T & 4095]
//     out(pos = T + m/2)  = (v0 - y) * g_lo2[T & 4095],   g_lo2[i] = g^(i + m/2)
// i.e. 3 products after the dots instead of 6 (1 instead of 2 when g = 1).  The one column whose second
// position wraps to 0 (T = m/2, factor g^0) recomputes the plain way.
// ------------------------------------------------------------------------------------------
// v0 = <M0, column>, v1 = <M1, column> for the 8 shares of column kk (four terms per Montgomery reduction)
__device__ __forceinline__ void king_l2_dots(const Fr* __restrict__ shares, size_t cols, size_t kk, const Fr* M0, const Fr* M1,
                                             Fr& v0, Fr& v1) {
    v0 = Fr::zero(); v1 = Fr::zero();
#pragma unroll
    for (int r = 0; r < 8; r += 4) {
        Fr x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = ld_fr(shares + (size_t)(r + q) * cols + kk);
        v0 = fp_add(v0, fp_dot<FrParams, 4>(x, M0 + r));
        v1 = fp_add(v1, fp_dot<FrParams, 4>(x, M1 + r));
    }
}

// per value of hi = T >> 12: the 2 x 8 unpack matrix with the high power-table factors folded in
//   Us[hi][0..7] = U0 * Ghi[hi],  Us[hi][8..15] = U1 * Whi[hi] * Ghi[hi]      (factors of hi = 0 are 1)
__global__ void k_king_scaled_matrices(const Fr* __restrict__ U, uint32_t hi_n, PowTable gen_tw, int has_g, PowTable g_tw,
                                       Fr* __restrict__ Us) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= hi_n * 16) return;
    const uint32_t hi = t >> 4, i = t & 15;
    Fr u = ld_fr(U + i);
    if (hi) {
        if (i >= 8) u = fp_mul(u, ld_fr(gen_tw.hi + hi));
        if (has_g) u = fp_mul(u, ld_fr(g_tw.hi + hi));
    }
    st_fr(Us + t, u);
}

__global__ void __launch_bounds__(256, 3)
k_king_stage1_l2(const Fr* __restrict__ shares, const Fr* __restrict__ U, const Fr* __restrict__ UsTab, size_t mbyl, size_t col0,
                 size_t cols, int log_m, int mode, PowTable gen_tw, int has_g, PowTable g_tw, const Fr* __restrict__ g_lo2,
                 SDest S) {
    const size_t Tlo = col0 + 1;
    const size_t T = (Tlo & ~(size_t)255) + (size_t)blockIdx.x * 256 + threadIdx.x;
    const Fr* Us = UsTab + (size_t)(T >> TW_LO_BITS) * 16;             // uniform over the block (L1-resident)
    if (T < Tlo || T > col0 + cols) return;
    const size_t k = T - 1, kk = k - col0;
    const size_t m = (size_t)1 << log_m;
    const uint32_t lo = (uint32_t)(T & (TW_LO - 1));
    // the one column whose second position wraps to 0 (T = m/2: factor g^0 there) takes the plain matrix and
    // composes its own factors
    const bool wrap = T == mbyl;
    const Fr* M = wrap ? U : Us;
    Fr v0, v1;
    king_l2_dots(shares, cols, kk, M, M + 8, v0, v1);
    Fr y = fp_mul(v1, wrap ? pow_lookup(gen_tw, T) : ld_fr(gen_tw.lo + lo));
    Fr a = fp_add(v0, y), b = fp_sub(v0, y);
    if (has_g) {
        a = fp_mul(a, wrap ? pow_lookup(g_tw, T) : ld_fr(g_tw.lo + lo));
        if (!wrap) b = fp_mul(b, ld_fr(g_lo2 + lo));
    }
    const size_t pos0 = T, pos1 = (T + mbyl) & (m - 1);
    size_t d0 = pos0, d1 = pos1;
    if (mode == 1) {
        size_t p0 = (size_t)(__brevll((unsigned long long)pos0) >> (64 - log_m));
        size_t p1 = (size_t)(__brevll((unsigned long long)pos1) >> (64 - log_m));
        d0 = (p0 & (mbyl - 1)) * 2 + (p0 / mbyl);
        d1 = (p1 & (mbyl - 1)) * 2 + (p1 / mbyl);
    }
    st_fr(sdest(S, d0), a);
    st_fr(sdest(S, d1), b);
}

// The same for latency-bound sizes (<= 2^13 columns: 17.6 -> 13.8 us; no gain from 2^15 up): TWO threads per column,
// lane j of a pair computing secret j (two 4-term inner
// products), one exchange through warp shuffles (lane 1 hands over y = v1 * twiddle, lane 0 hands over v0), and each
// lane finishing and storing one of the two outputs.  ~900 wide MADs on the critical path instead of 1664.
__global__ void __launch_bounds__(256)
k_king_stage1_l2_split(const Fr* __restrict__ shares, const Fr* __restrict__ U, const Fr* __restrict__ UsTab, size_t mbyl, size_t col0,
                       size_t cols, int log_m, int mode, PowTable gen_tw, int has_g, PowTable g_tw, const Fr* __restrict__ g_lo2,
                       SDest S) {
    const size_t Tlo = col0 + 1;
    const uint32_t j = threadIdx.x & 1;
    const size_t T = (Tlo & ~(size_t)127) + (size_t)blockIdx.x * 128 + (threadIdx.x >> 1);
    const bool valid = T >= Tlo && T <= col0 + cols;                 // the same for both lanes of a pair
    const bool wrap = T == mbyl;
    const size_t m = (size_t)1 << log_m;
    const uint32_t lo = (uint32_t)(T & (TW_LO - 1));
    Fr v = Fr::zero(), y = Fr::zero();
    if (valid) {
        const size_t kk = T - 1 - col0;
        const Fr* M = (wrap ? U : UsTab + (size_t)(T >> TW_LO_BITS) * 16) + 8 * j;
#pragma unroll
        for (int r = 0; r < 8; r += 4) {
            Fr x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = ld_fr(shares + (size_t)(r + q) * cols + kk);
            v = fp_add(v, fp_dot<FrParams, 4>(x, M + r));
        }
        if (j) y = fp_mul(v, wrap ? pow_lookup(gen_tw, T) : ld_fr(gen_tw.lo + lo));
    }
    // lane 0 receives y, lane 1 receives v0 (every lane of the warp takes part in the shuffles)
    Fr other;
#pragma unroll
    for (int i = 0; i < 8; ++i) other.v[i] = __shfl_xor_sync(0xffffffffu, j ? y.v[i] : v.v[i], 1);
    if (!valid) return;
    Fr o = j ? fp_sub(other, y) : fp_add(v, other);
    if (has_g) {
        if (!j) o = fp_mul(o, wrap ? pow_lookup(g_tw, T) : ld_fr(g_tw.lo + lo));
        else if (!wrap) o = fp_mul(o, ld_fr(g_lo2 + lo));
    }
    const size_t pos = j ? (T + mbyl) & (m - 1) : T;
    size_t d = pos;
    if (mode == 1) {
        size_t pb = (size_t)(__brevll((unsigned long long)pos) >> (64 - log_m));
        d = (pb & (mbyl - 1)) * 2 + (pb / mbyl);
    }
    st_fr(sdest(S, d), o);
}

// ------------------------------------------------------------------------------------------
// Dense map with the INPUT vector in registers (few inputs, many outputs): pack / det_pack.
//   out(c, i) = sum_{j<K1} M[i][j] * in1(c, j) + sum_{j<K2} M[i][K1+j] * in2(c, j)
// element (c, j) of an operand sits at ptr[c * cs + j * rs].
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
k_map_in_regs(const Fr* __restrict__ M, int rows, int k1, const Fr* __restrict__ in1, size_t in1_cs, size_t in1_rs,
              const Fr* __restrict__ in2, size_t in2_cs, size_t in2_rs, int k2, Fr* __restrict__ out, size_t out_cs,
              size_t out_rs, size_t cols) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    Fr x[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < k1) x[j] = ld_fr(in1 + c * in1_cs + (size_t)j * in1_rs);
        else if (j < k1 + k2) x[j] = ld_fr(in2 + c * in2_cs + (size_t)(j - k1) * in2_rs);
        else x[j] = Fr::zero();
    }
    // M is zero-padded to K columns and x to K entries, so whole groups of four can always be used
    for (int i = 0; i < rows; ++i) {
        Fr acc = Fr::zero();
#pragma unroll
        for (int j = 0; j < K; j += 4) {
            Fr d = fp_dot<FrParams, 4>(x + j, M + (size_t)i * K + j);      // matrix limbs read as needed
            acc = j == 0 ? d : fp_add(acc, d);
        }
        st_fr(out + c * out_cs + (size_t)i * out_rs, acc);
    }
}

// pack for the reference's one real configuration (l = t = 2, n = 8; pss.rs:14 "(2, 2, 8) - currently
// implemented") as the transforms it is defined by -- size-4 inverse DFT on the coset g<zeta_4>, size-8
// DFT on <zeta_8> of the zero-padded coefficients -- instead of the dense 8 x 4 matrix: 10 field
// products per column instead of 32.  Same canonical results (the map is the same linear map).
// cst[0] = zeta_4^-1, cst[1..4] = g^-d / 4 (d = 0..3), cst[5] = zeta_4, cst[6..8] = zeta_8^1..3
struct Pack2Consts { FrArg c[9]; FrArg oa[4], ob[4]; };   // oa/ob: the odd half as 2-term inner products (see k_pack_l2)
__global__ void __launch_bounds__(256)
k_pack_l2(Pack2Consts K, const Fr* __restrict__ secrets, size_t s_cs, size_t s_rs, const Fr* __restrict__ rand, size_t r_cs,
          size_t r_rs, int has_rand, Fr* __restrict__ out, size_t out_cs, size_t out_rs, size_t cols) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    Fr v0 = ld_fr(secrets + c * s_cs), v1 = ld_fr(secrets + c * s_cs + s_rs);
    Fr v2 = has_rand ? ld_fr(rand + c * r_cs) : Fr::zero(), v3 = has_rand ? ld_fr(rand + c * r_cs + r_rs) : Fr::zero();
    // inverse DFT_4 (root zeta_4^-1), then coefficient d scaled by g^-d / 4
    Fr a0 = fp_add(v0, v2), a1 = fp_sub(v0, v2), b0 = fp_add(v1, v3);
    Fr c0 = fp_mul(fp_add(a0, b0), from_arg(K.c[1]));
    Fr c2 = fp_mul(fp_sub(a0, b0), from_arg(K.c[3]));
    // DFT_8 of (c0, c1, c2, c3, 0, 0, 0, 0): s_j = E_(j mod 4) + zeta_8^j O_(j mod 4)
    Fr t2 = fp_mul(c2, from_arg(K.c[5]));
    Fr E[4] = {fp_add(c0, c2), fp_add(c0, t2), fp_sub(c0, c2), fp_sub(c0, t2)};
    // the odd half never materialises c1, c3: zeta_8^k O_k = oa[k] * a1 + ob[k] * (v1 - v3), one 2-term inner product
    // each (4 x 192 wide MADs instead of the 7 x 128 of b1, c1, c3, t3 and the three zeta_8^k factors)
    Fr x[2] = {a1, fp_sub(v1, v3)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        Fr y[2] = {from_arg(K.oa[k]), from_arg(K.ob[k])};
        Fr o = fp_dot<FrParams, 2>(x, y);
        st_fr(out + c * out_cs + (size_t)k * out_rs, fp_add(E[k], o));
        st_fr(out + c * out_cs + (size_t)(k + 4) * out_rs, fp_sub(E[k], o));
    }
}

// Small batches (<= 2^13 columns; measured: 13.3 -> 11.3 us at 2^13, but 13.2 -> 15.4 us at 2^15, where the duplicated
// even half costs more than the shorter chain saves) are latency-bound: the same pack with TWO threads per column,
// lane h of a pair producing outputs h, h+2, h+4, h+6 (E_h, E_(h+2) and the two inner products that go with them; the
// three products of the even half are computed by both lanes).  768 wide MADs on the critical path instead of 1152.
__global__ void __launch_bounds__(256)
k_pack_l2_split(Pack2Consts K, const Fr* __restrict__ secrets, size_t s_cs, size_t s_rs, const Fr* __restrict__ rand, size_t r_cs,
                size_t r_rs, int has_rand, Fr* __restrict__ out, size_t out_cs, size_t out_rs, size_t cols) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t c = t >> 1;
    const int h = (int)(t & 1);
    if (c >= cols) return;
    Fr v0 = ld_fr(secrets + c * s_cs), v1 = ld_fr(secrets + c * s_cs + s_rs);
    Fr v2 = has_rand ? ld_fr(rand + c * r_cs) : Fr::zero(), v3 = has_rand ? ld_fr(rand + c * r_cs + r_rs) : Fr::zero();
    Fr a0 = fp_add(v0, v2), a1 = fp_sub(v0, v2), b0 = fp_add(v1, v3);
    Fr c0 = fp_mul(fp_add(a0, b0), from_arg(K.c[1]));
    Fr c2 = fp_mul(fp_sub(a0, b0), from_arg(K.c[3]));
    if (h) c2 = fp_mul(c2, from_arg(K.c[5]));            // E_1, E_3 use zeta_8^2 c2
    Fr E0 = fp_add(c0, c2), E2 = fp_sub(c0, c2);          // E_h, E_(h+2)
    Fr x[2] = {a1, fp_sub(v1, v3)};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = h + 2 * q;
        Fr y[2] = {from_arg(K.oa[k]), from_arg(K.ob[k])};
        Fr o = fp_dot<FrParams, 2>(x, y);
        const Fr& E = q ? E2 : E0;
        st_fr(out + c * out_cs + (size_t)k * out_rs, fp_add(E, o));
        st_fr(out + c * out_cs + (size_t)(k + 4) * out_rs, fp_sub(E, o));
    }
}

// Dense map with the OUTPUT accumulators in registers (many inputs, few outputs): unpack / unpack2.
template <int ROWS>
__global__ void __launch_bounds__(256)
k_map_acc_regs(const Fr* __restrict__ M, int k, const Fr* __restrict__ in, size_t in_cs, size_t in_rs,
               Fr* __restrict__ out, size_t out_cs, size_t out_rs, size_t cols) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    Fr acc[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) acc[i] = Fr::zero();
    int j = 0;
    for (; j + 4 <= k; j += 4) {
        Fr x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = ld_fr(in + c * in_cs + (size_t)(j + q) * in_rs);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) acc[i] = fp_add(acc[i], fp_dot<FrParams, 4>(x, M + (size_t)i * k + j));
    }
    for (; j < k; ++j) {
        Fr x = ld_fr(in + c * in_cs + (size_t)j * in_rs);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) acc[i] = fp_add(acc[i], fp_mul(ld_fr(M + (size_t)i * k + j), x));
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) st_fr(out + c * out_cs + (size_t)i * out_rs, acc[i]);
}

// v[i] *= g^i
__global__ void k_distribute_powers(Fr* __restrict__ v, size_t n, PowTable g_tw) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    st_fr(v + i, fp_mul(ld_fr_rw(v + i), pow_lookup(g_tw, i)));
}
// v[i] *= s * g^i  (ifft tail: size_inv * offset^-i)
__global__ void k_scale_powers(Fr* __restrict__ v, size_t n, FrArg s_, int has_g, PowTable g_tw) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr f = from_arg(s_);
    if (has_g && i) f = fp_mul(f, pow_lookup(g_tw, i));
    st_fr(v + i, fp_mul(ld_fr_rw(v + i), f));
}
// Outer step of the rank-sharded fft1 (four-step with G = number of ranks as the short dimension):
// recv[g][j] = w^(i1*k2) * Inner_i1[k2] from rank g (i1 = bitrev_G(g), k2 = first_col + j);
// out[k1][j] = X[k2 + N2*k1] = sum_i1 wG^(i1*k1) * recv[bitrev_G(i1)][j],  wG = w^(N2) of order G.
__global__ void __launch_bounds__(256)
k_fft1_shard_outer(const Fr* __restrict__ recv, size_t cnt, int G, int logG, const Fr* __restrict__ wG_pows,
                   Fr* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    for (int k1 = 0; k1 < G; ++k1) {
        Fr acc = Fr::zero();
        for (int g = 0; g < G; ++g) {
            int i1 = (int)bitrev32((uint32_t)g, logG);
            int e = (i1 * k1) & (G - 1);
            Fr v = ld_fr(recv + (size_t)g * cnt + j);
            if (e) v = fp_mul(v, ld_fr(wG_pows + e));
            acc = fp_add(acc, v);
        }
        st_fr(out + (size_t)k1 * cnt + j, acc);
    }
}
// The same G-point transform as log2 G radix-2 stages in registers (G <= 8): the rows arrive in bit-reversed i1 order (row g
// holds i1 = bitrev_G(g)), which is exactly the input order of a decimation-in-time transform with natural-order output --
// G/2 log2 G products per column instead of G^2 (12 instead of 64 at G = 8: 0.11 ms -> 0.02 ms for a 2^20-element block).
template <int LG>
__global__ void __launch_bounds__(256)
k_fft1_shard_outer_r2(const Fr* __restrict__ recv, size_t cnt, const Fr* __restrict__ wG_pows, Fr* __restrict__ out) {
    constexpr int G = 1 << LG;
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    Fr v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) v[g] = ld_fr(recv + (size_t)g * cnt + j);
#pragma unroll
    for (int s_ = 1; s_ <= LG; ++s_) {
        const int len = 1 << s_, half = len >> 1;
#pragma unroll
        for (int i = 0; i < G; i += len) {
#pragma unroll
            for (int k = 0; k < half; ++k) {
                Fr u = v[i + k], w = v[i + k + half];
                const int e = k * (G / len);                 // twiddle wG^e
                if (e) w = fp_mul(w, ld_fr(wG_pows + e));
                v[i + k] = fp_add(u, w);
                v[i + k + half] = fp_sub(u, w);
            }
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < G; ++k1) st_fr(out + (size_t)k1 * cnt + j, v[k1]);
}

// fft1 sharded, local step's last pass fused with the exchange: element k2 of this rank's twiddled inner transform
// belongs to rank k2 / cnt (cnt = N2 / G columns per rank) and is stored straight into that rank's receive buffer
// (chunk `rank` of it) -- over NVLink peer memory for the other ranks.  The twiddle product w^(i1 k2) that the local
// step needs anyway is applied on the way, so the all-to-all costs no extra pass over the data.
__global__ void __launch_bounds__(256)
k_twiddle_scatter(const Fr* __restrict__ in, size_t N2, int has_tw, PowTable tw, SDest recv_by_rank, int log_cnt, uint32_t rank) {
    size_t k2 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k2 >= N2) return;
    Fr x = ld_fr_rw(in + k2);
    if (has_tw && k2) x = fp_mul(x, pow_lookup(tw, k2));
    const size_t cnt = (size_t)1 << log_cnt;
    st_fr(recv_by_rank.p[k2 >> log_cnt] + ((size_t)rank << log_cnt) + (k2 & (cnt - 1)), x);
}

__global__ void k_vec_add(Fr* __restrict__ v, const Fr* __restrict__ a, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(v + i, fp_add(ld_fr_rw(v + i), ld_fr(a + i)));
}

// out[bitrev(i)] = in[i]
__global__ void k_bitrev(const Fr* __restrict__ in, Fr* __restrict__ out, int log_n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_n)) return;
    size_t r = log_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
    st_fr(out + r, ld_fr(in + i));
}

// ------------------------------------------------------------------------------------------
// d_pp king closure pieces (dist-primitives/src/dpp/mod.rs:55-66): q[i] = num[i] / den[i] with a
// chunked Montgomery batch inversion, then the inclusive prefix PRODUCT of q as a three-kernel scan.
// ------------------------------------------------------------------------------------------
static constexpr uint32_t DPP_CHUNK = 32;     // elements per thread

// q[i] = num[i] * den[i]^-1 over one chunk per thread; scratch holds the running products.  A zero
// denominator (the reference's `.inverse().unwrap()` panics) raises *err.
__global__ void __launch_bounds__(128)
k_dpp_divide(Fr* __restrict__ num, const Fr* __restrict__ den, Fr* __restrict__ scratch, size_t m, int* __restrict__ err) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * DPP_CHUNK, hi = lo + DPP_CHUNK < m ? lo + DPP_CHUNK : m;
    if (lo >= m) return;
    Fr run = Fr::one();
    for (size_t i = lo; i < hi; ++i) {
        Fr d = ld_fr(den + i);
        if (d.is_zero()) { atomicExch(err, 1); d = Fr::one(); }
        st_fr(scratch + i, run);                 // product of den[lo..i)
        run = fp_mul(run, d);
    }
    Fr inv = fp_inv(run);                        // one inversion per chunk
    for (size_t i = hi; i-- > lo;) {
        Fr d = ld_fr(den + i);
        if (d.is_zero()) d = Fr::one();
        Fr di = fp_mul(inv, ld_fr_rw(scratch + i));      // den[i]^-1
        inv = fp_mul(inv, d);
        st_fr(num + i, fp_mul(ld_fr_rw(num + i), di));
    }
}
// phase 1: in-place inclusive product inside each chunk, chunk total to tot[t]
__global__ void __launch_bounds__(128)
k_dpp_scan_local(Fr* __restrict__ q, size_t m, Fr* __restrict__ tot) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * DPP_CHUNK, hi = lo + DPP_CHUNK < m ? lo + DPP_CHUNK : m;
    if (lo >= m) return;
    Fr run = ld_fr_rw(q + lo);
    for (size_t i = lo + 1; i < hi; ++i) { run = fp_mul(run, ld_fr_rw(q + i)); st_fr(q + i, run); }
    st_fr(tot + t, run);
}
// phase 2: exclusive prefix product of the chunk totals, one block (sequential over tiles of 1024)
__global__ void __launch_bounds__(1024)
k_dpp_scan_totals(Fr* __restrict__ tot, size_t n) {
    __shared__ Fr sh[1024];
    Fr carry = Fr::one();
    for (size_t base = 0; base < n; base += 1024) {
        size_t i = base + threadIdx.x;
        Fr x = i < n ? ld_fr_rw(tot + i) : Fr::one();
        sh[threadIdx.x] = x;
        __syncthreads();
        for (uint32_t off = 1; off < 1024; off <<= 1) {
            Fr a = threadIdx.x >= off ? sh[threadIdx.x - off] : Fr::one();
            __syncthreads();
            if (threadIdx.x >= off) sh[threadIdx.x] = fp_mul(sh[threadIdx.x], a);
            __syncthreads();
        }
        Fr excl = threadIdx.x ? sh[threadIdx.x - 1] : Fr::one();
        if (i < n) st_fr(tot + i, fp_mul(carry, excl));
        carry = fp_mul(carry, sh[1023]);
        __syncthreads();
    }
}
// phase 3: multiply every element of chunk t by the product of all earlier chunks
__global__ void __launch_bounds__(256)
k_dpp_scan_apply(Fr* __restrict__ q, size_t m, const Fr* __restrict__ tot) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || i < DPP_CHUNK) return;
    st_fr(q + i, fp_mul(ld_fr_rw(q + i), ld_fr(tot + i / DPP_CHUNK)));
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------

__global__ void k_negate(Fr* __restrict__ v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(v + i, fp_neg(ld_fr(v + i)));
}

static int ilog2(size_t x) { int r = 0; while (((size_t)1 << r) < x) ++r; return r; }
static bool is_pow2(size_t x) { return x && !(x & (x - 1)); }

// Parameter tables live in the context's persistent cache (keyed by their defining values), so a
// prover that keeps calling d_fft / d_ifft with the same domain builds them once.
struct PowKey { char tag[8]; uint64_t w[4]; uint64_t count; };

// out[i] = w^i, i < count, cached
static int32_t cached_pow_seq(zkg_ctx* ctx, const char* tag, const HFr& w, size_t count, const Fr** out) {
    PowKey k;
    memset(&k, 0, sizeof k);
    strncpy(k.tag, tag, sizeof k.tag - 1);
    memcpy(k.w, w.v, 32);
    k.count = count;
    void* p; bool fresh;
    ZKG_TRY(ctx_cache_get(ctx, &k, sizeof k, (count ? count : 1) * sizeof(Fr), &p, &fresh));
    if (fresh && count) {
        k_pow_table<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(to_arg(w), (uint32_t)count, (Fr*)p);
        ctx->launches += 1;
        ZKG_CUDA(cudaGetLastError());
    }
    *out = (const Fr*)p;
    return ZKG_OK;
}

// out[i] = c * w^i, i < count, cached (keyed by w, c and count)
static int32_t cached_pow_seq_scaled(zkg_ctx* ctx, const char* tag, const HFr& w, const HFr& c, size_t count, const Fr** out) {
    struct { PowKey k; uint64_t c[4]; } key;
    memset(&key, 0, sizeof key);
    strncpy(key.k.tag, tag, sizeof key.k.tag - 1);
    memcpy(key.k.w, w.v, 32);
    key.k.count = count;
    memcpy(key.c, c.v, 32);
    void* p; bool fresh;
    ZKG_TRY(ctx_cache_get(ctx, &key, sizeof key, (count ? count : 1) * sizeof(Fr), &p, &fresh));
    if (fresh && count) {
        k_pow_table<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(to_arg(w), (uint32_t)count, (Fr*)p);
        k_scale_table<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>((Fr*)p, (uint32_t)count, to_arg(c));
        ctx->launches += 2;
        ZKG_CUDA(cudaGetLastError());
    }
    *out = (const Fr*)p;
    return ZKG_OK;
}

static int32_t build_pow_table(zkg_ctx* ctx, const HFr& w, size_t max_exp_excl, PowTable* out) {
    size_t hi_n = (max_exp_excl + TW_LO - 1) >> TW_LO_BITS;
    if (hi_n == 0) hi_n = 1;
    ZKG_TRY(cached_pow_seq(ctx, "pow_lo", w, TW_LO, &out->lo));
    ZKG_TRY(cached_pow_seq(ctx, "pow_hi", host::h_pow(w, TW_LO), hi_n, &out->hi));
    return ZKG_OK;
}

// In-order-output NTT of d_in (bit-reversed input order) with root w_N, into d_out.
// shift = 1 stores X[k] at (k-1) mod N.  d_tmp: N-element scratch (used when > 2 passes or in == out).
static int32_t ntt_bitrev_in(zkg_ctx* ctx, const Fr* d_in, Fr* d_out, Fr* d_tmp, size_t N, const HFr& wN,
                             int shift, const HFr* scale, const Fr* d_mask) {
    const int logN = ilog2(N);
    ZKG_REQUIRE(is_pow2(N) && logN <= 24 + 3, "ntt: size %zu unsupported", N);
    int npass = (logN + 9) / 10;
    if (npass == 0) npass = 1;
    phase_mark(ctx, 0);
    PowTable tw{nullptr, nullptr};
    if (npass > 1) ZKG_TRY(build_pow_table(ctx, wN, N, &tw));
    int s = 0;
    const Fr* src = d_in;
    for (int q = 0; q < npass; ++q) {
        int b = (logN - s + (npass - q) - 1) / (npass - q);      // spread the remaining bits evenly
        // radix-8 register-blocked kernel for large transforms; small ones keep the radix-2 kernel, whose four
        // times as many threads hide latency better when there is less than one tile per SM
        const bool r8 = logN >= env_int_ntt("ZKG_NTT_R8_MIN", 18);
        // elements per thread of the register-blocked kernel: 8 (radix-8 groups, 122 registers, 16 warps/SM), or 4
        // (radix-4 groups, 76 registers, 24 warps/SM) for the largest transforms, where the extra resident warps win
        // 3-5 % (N = 2^21: 0.472 -> 0.450 ms, 2^23: 1.894 -> 1.841 ms) and lose nothing below
        const int ept_env = env_int_ntt("ZKG_NTT_EPT_LOG", 0);
        const int ept_log = ept_env == 2 || ept_env == 3 ? ept_env : (logN >= 21 ? 2 : 3);
        int elog = env_int_ntt("ZKG_NTT_ELOG", ept_log == 3 ? 11 : 10);
        if (elog > 8 + ept_log) elog = 8 + ept_log;                          // at most 256 threads per block
        if (elog < 5) elog = 5;
        int cw_log = 0;
        if (q > 0 || r8) {
            cw_log = (r8 ? elog : 11) - b; if (cw_log > (q > 0 ? s : logN - b)) cw_log = q > 0 ? s : logN - b; if (cw_log < 0) cw_log = 0;
            // small transforms: prefer more, narrower tiles (down to 2 columns = 64-byte rows; measured: m = 2^17
            // 33.7 -> 25.5 us against a floor of 4 columns) so that the pass spreads over the 148 SMs instead of a
            // dozen fat blocks
            const int cw_floor = env_int_ntt("ZKG_NTT_CW_FLOOR", 1);
            while (cw_log > cw_floor && (N >> (b + cw_log)) < 296) --cw_log;
        }
        NttPass P;
        P.logN = logN; P.s = s; P.b = b; P.cw_log = cw_log;
        P.first = q == 0; P.last = q == npass - 1; P.shift = shift;
        P.has_scale = scale != nullptr;
        P.scale = to_arg(scale ? *scale : host::h_one());
        // small twiddles: w_T^i, i < T/2, w_T = wN^(N/T)
        size_t T = (size_t)1 << b;
        const Fr* tws;
        HFr wT = host::h_pow(wN, N >> b);
        ZKG_TRY(cached_pow_seq(ctx, "tw_small", wT, T / 2, &tws));
        // inter-pass twiddle table (passes after the first); beyond 2^24 entries the pass composes them on the fly
        const Fr* tw_full = nullptr;
        if (q > 0 && s + b <= 24 && !getenv("ZKG_NTT_ONTHEFLY")) {
            struct { PowKey k; int s, b, logN, pad; } key;
            memset(&key, 0, sizeof key);
            strncpy(key.k.tag, "ntt_tw", sizeof key.k.tag - 1);
            memcpy(key.k.w, wN.v, 32);
            key.s = s; key.b = b; key.logN = logN;
            const size_t cnt = (size_t)1 << (s + b);
            void* tp; bool fresh;
            ZKG_TRY(ctx_cache_get(ctx, &key, sizeof key, cnt * sizeof(Fr), &tp, &fresh));
            if (fresh) {
                k_ntt_twiddle_table<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>((Fr*)tp, s, b, logN - s - b, tw);
                ctx->launches += 1;
                ZKG_CUDA(cudaGetLastError());
            }
            tw_full = (const Fr*)tp;
        }
        // destination: last pass -> d_out; otherwise the scratch (in place on scratch is safe: a
        // block reads and writes the same index set when no shift is applied)
        Fr* dst = P.last ? d_out : d_tmp;
        size_t E = T << cw_log;
        unsigned blocks = (unsigned)(N / E);
        if (q == 0) phase_mark(ctx, 1);
        if (r8) {
            size_t PL = E + (E >> 3) + 1;
            size_t shmem = (2 * PL + T) * sizeof(uint4);
            if (ept_log == 3) {
                unsigned threads = (unsigned)(E / 8 < 32 ? 32 : E / 8);
                ZKG_CUDA(cudaFuncSetAttribute(k_ntt_pass8<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                k_ntt_pass8<3><<<blocks, threads, shmem, ctx->stream>>>(src, dst, P, tws, tw, tw_full, P.last ? d_mask : nullptr, q == 0 ? 1 : 0);
            } else {
                unsigned threads = (unsigned)(E / 4 < 32 ? 32 : E / 4);
                ZKG_CUDA(cudaFuncSetAttribute(k_ntt_pass8<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                k_ntt_pass8<2><<<blocks, threads, shmem, ctx->stream>>>(src, dst, P, tws, tw, tw_full, P.last ? d_mask : nullptr, q == 0 ? 1 : 0);
            }
        } else {
            size_t shmem = (8 * E + 8 * (T / 2 ? T / 2 : 1)) * sizeof(uint32_t);
            unsigned threads = (unsigned)(E / 2 < 32 ? 32 : (E / 2 > 512 ? 512 : E / 2));
            ZKG_CUDA(cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            k_ntt_pass<<<blocks, threads, shmem, ctx->stream>>>(src, dst, P, tws, tw, tw_full, P.last ? d_mask : nullptr);
        }
        ctx->launches += 1;
        if (P.last) phase_mark(ctx, 2);
        ZKG_CUDA(cudaGetLastError());
        src = dst;
        s += b;
    }
    return ZKG_OK;
}


// cached PSS matrices (host) per packing factor
static std::mutex g_pss_mu;
static host::PssMatrices g_pss[4];      // l = 2, 4, 8
static const host::PssMatrices* pss_get(uint32_t l) {
    int slot = l == 2 ? 0 : l == 4 ? 1 : l == 8 ? 2 : -1;
    if (slot < 0) return nullptr;
    std::lock_guard<std::mutex> lk(g_pss_mu);
    if (g_pss[slot].l != l) host::pss_matrices(l, &g_pss[slot]);
    return &g_pss[slot];
}

static int32_t upload(zkg_ctx* ctx, const std::vector<HFr>& m, const Fr** out) {
    void* p; bool fresh;
    ZKG_TRY(ctx_cache_get(ctx, m.data(), m.size() * sizeof(HFr), m.size() * sizeof(Fr), &p, &fresh));
    // blocking copy, once per distinct matrix: the source may be a temporary
    if (fresh && !m.empty()) ZKG_CUDA(cudaMemcpy(p, m.data(), m.size() * sizeof(Fr), cudaMemcpyHostToDevice));
    *out = (const Fr*)p;
    return ZKG_OK;
}

// pad matrix rows from kk to K columns (k_map_in_regs indexes rows with stride K)
static std::vector<HFr> pad_rows(const std::vector<HFr>& m, int rows, int kk, int K) {
    std::vector<HFr> o((size_t)rows * K, host::h_zero());
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < kk; ++j) o[(size_t)i * K + j] = m[(size_t)i * kk + j];
    return o;
}

static Pack2Consts pack2_consts() {
    using namespace host;
    HFr g = h_load(BN254_FR_GENERATOR_MONT_64), z4 = h_root_of_unity(4), z8 = h_root_of_unity(8);
    HFr ginv = h_inv(g), quarter = h_inv(h_from_u64(4));
    Pack2Consts K;
    K.c[0] = to_arg(h_inv(z4));
    HFr sc = quarter;
    for (int d = 0; d < 4; ++d) { K.c[1 + d] = to_arg(sc); sc = h_mul(sc, ginv); }
    K.c[5] = to_arg(z4);                 // == zeta_8^2
    K.c[6] = to_arg(z8);
    K.c[7] = to_arg(z4);
    K.c[8] = to_arg(h_mul(z8, z4));
    // odd half: c1 = (a1 + b1) K2, c3 = (a1 - b1) K4, b1 = d K0 (d = v1 - v3);  O_k = c1 +- c3 [* K5];  o_k = z_k O_k
    HFr K0 = h_inv(z4), K2 = h_mul(quarter, ginv), K4 = h_mul(h_mul(K2, ginv), ginv), K5 = z4;
    HFr zk[4] = {h_one(), z8, z4, h_mul(z8, z4)};
    for (int k = 0; k < 4; ++k) {
        HFr w = (k & 1) ? h_mul(K4, K5) : K4;                       // the c3 coefficient of O_k before its sign
        HFr pa = k < 2 ? h_add(K2, w) : h_sub(K2, w);               // a1 coefficient
        HFr pb = k < 2 ? h_sub(K2, w) : h_add(K2, w);               // b1 coefficient (c3 carries -b1)
        K.oa[k] = to_arg(h_mul(zk[k], pa));
        K.ob[k] = to_arg(h_mul(zk[k], h_mul(K0, pb)));
    }
    return K;
}

static int32_t launch_pack(zkg_ctx* ctx, const Fr* dM, int K, int rows, int l, int t_used, const Fr* secrets, size_t s_cs,
                           size_t s_rs, const Fr* rand, size_t r_cs, size_t r_rs, Fr* out, size_t o_cs, size_t o_rs,
                           size_t cols) {
    if (cols == 0) return ZKG_OK;
    unsigned blocks = (unsigned)((cols + 255) / 256);
    if (l == 2 && rows == 8 && !getenv("ZKG_PACK_DENSE")) {
        static const Pack2Consts P2 = pack2_consts();
        if (cols <= (size_t)env_int_ntt("ZKG_KING_SPLIT_MAX", 1 << 13))      // latency-bound sizes: two threads per column
            k_pack_l2_split<<<(unsigned)((2 * cols + 255) / 256), 256, 0, ctx->stream>>>(P2, secrets, s_cs, s_rs, rand, r_cs, r_rs, t_used ? 1 : 0, out, o_cs, o_rs, cols);
        else
            k_pack_l2<<<blocks, 256, 0, ctx->stream>>>(P2, secrets, s_cs, s_rs, rand, r_cs, r_rs, t_used ? 1 : 0, out, o_cs, o_rs, cols);
        ctx->launches += 1;
        ZKG_CUDA(cudaGetLastError());
        return ZKG_OK;
    }
#define LP(KK) k_map_in_regs<KK><<<blocks, 256, 0, ctx->stream>>>(dM, rows, l, secrets, s_cs, s_rs, rand, r_cs, r_rs, t_used, out, o_cs, o_rs, cols)
    if (K == 4) LP(4); else if (K == 8) LP(8); else LP(16);
#undef LP
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

static int32_t launch_unpack(zkg_ctx* ctx, const Fr* dM, int l, int k, const Fr* in, size_t i_cs, size_t i_rs, Fr* out,
                             size_t o_cs, size_t o_rs, size_t cols) {
    if (cols == 0) return ZKG_OK;
    unsigned blocks = (unsigned)((cols + 255) / 256);
    if (l == 2) k_map_acc_regs<2><<<blocks, 256, 0, ctx->stream>>>(dM, k, in, i_cs, i_rs, out, o_cs, o_rs, cols);
    else if (l == 4) k_map_acc_regs<4><<<blocks, 256, 0, ctx->stream>>>(dM, k, in, i_cs, i_rs, out, o_cs, o_rs, cols);
    else k_map_acc_regs<8><<<blocks, 256, 0, ctx->stream>>>(dM, k, in, i_cs, i_rs, out, o_cs, o_rs, cols);
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

// keeps host-side matrices alive until the stream has consumed the async upload
struct HostKeep { std::deque<std::vector<HFr>> v; };   // deque: references stay valid across push_back

// unpack matrix for the received party set (unpack2 if all present, Lagrange otherwise)
static int32_t recv_matrix(uint32_t l, const uint32_t* parties, uint32_t n_recv, HostKeep& keep, const std::vector<HFr>** out) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    if (n_recv == pm->n) { *out = &pm->unpack2; return ZKG_OK; }
    ZKG_REQUIRE(parties, "parties list required when shares are missing");
    keep.v.emplace_back();
    if (!host::pss_lagrange_matrix(l, parties, n_recv, &keep.v.back())) {
        set_error("not enough shares to reconstruct: got %u of n = %u (need > %u distinct parties)", n_recv, pm->n,
                  2 * (pm->t + pm->l - 1));
        return ZKG_ERR_BAD_ARG;
    }
    *out = &keep.v.back();
    return ZKG_OK;
}

// ---- king pipeline on device buffers ------------------------------------------------------
// mode_fft: 1 = fft2 + powers + (re)packing (d_fft/d_ifft king), 0 = deg_red king
// stage 1 on the share columns [col0, col0+cols): unpack (+ fft2 + powers) and scatter into S (pack order)
static int32_t king_stage1(zkg_ctx* ctx, const Fr* d_shares, const uint32_t* parties, uint32_t n_recv, size_t mbyl,
                           size_t col0, size_t cols, uint32_t l, const HFr* gen, const HFr* g, int rearrange,
                           int mode_fft, const SDest& S, HostKeep& keep) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    ZKG_REQUIRE(is_pow2(mbyl) || !mode_fft, "king: m/l = %zu is not a power of two", mbyl);
    ZKG_REQUIRE(col0 + cols <= mbyl, "king: column range [%zu, %zu) outside 0..%zu", col0, col0 + cols, mbyl);
    if (cols == 0) return ZKG_OK;
    const size_t m = mbyl * l;
    const int log_m = ilog2(m);
    ZKG_REQUIRE(!mode_fft || log_m <= 28, "king: m = %zu exceeds the 2-adicity of Fr", m);
    const std::vector<HFr>* U;
    ZKG_TRY(recv_matrix(l, parties, n_recv, keep, &U));
    const Fr* dU;
    phase_mark(ctx, 0);
    ZKG_TRY(upload(ctx, *U, &dU));
    PowTable gen_tw{nullptr, nullptr}, g_tw{nullptr, nullptr};
    int has_g = 0;
    if (mode_fft) {
        ZKG_TRY(build_pow_table(ctx, *gen, m, &gen_tw));
        has_g = !(*g == host::h_one());
        if (has_g) ZKG_TRY(build_pow_table(ctx, *g, m, &g_tw));
    }
    unsigned blocks = (unsigned)((cols + 255) / 256);
    int mode = mode_fft ? (rearrange ? 1 : 0) : 2;
    if (l == 2 && n_recv == 8 && mode_fft && !getenv("ZKG_KING_GENERIC")) {
        // specialised kernel: thread <-> T = k + 1, blocks aligned on 256 values of T
        const Fr* g_lo2 = nullptr;
        if (has_g) ZKG_TRY(cached_pow_seq_scaled(ctx, "pow_l2", *g, host::h_pow(*g, mbyl), TW_LO, &g_lo2));
        const uint32_t hi_n = (uint32_t)((mbyl >> TW_LO_BITS) + 1);      // T = 1 .. m/2
        struct { char tag[8]; uint64_t gen[4], g[4]; uint64_t hi_n; } key;
        memset(&key, 0, sizeof key);
        strncpy(key.tag, "kingUs", sizeof key.tag - 1);
        memcpy(key.gen, gen->v, 32);
        memcpy(key.g, g->v, 32);
        key.hi_n = hi_n;
        void* up; bool fresh;
        ZKG_TRY(ctx_cache_get(ctx, &key, sizeof key, (size_t)hi_n * 16 * sizeof(Fr), &up, &fresh));
        if (fresh) {
            k_king_scaled_matrices<<<(hi_n * 16 + 127) / 128, 128, 0, ctx->stream>>>(dU, hi_n, gen_tw, has_g, g_tw, (Fr*)up);
            ctx->launches += 1;
            ZKG_CUDA(cudaGetLastError());
        }
        const size_t Tlo = col0 + 1, Thi = col0 + cols, A = Tlo & ~(size_t)255;
        unsigned blocks2 = (unsigned)((Thi - A) / 256 + 1);
        phase_mark(ctx, 1);
        if (cols <= (size_t)env_int_ntt("ZKG_KING_SPLIT_MAX", 1 << 13)) {     // latency-bound sizes: two threads per column
            const size_t A2 = Tlo & ~(size_t)127;
            k_king_stage1_l2_split<<<(unsigned)((Thi - A2) / 128 + 1), 256, 0, ctx->stream>>>(d_shares, dU, (const Fr*)up, mbyl, col0, cols,
                                                                                          log_m, mode, gen_tw, has_g, g_tw, g_lo2, S);
        } else
        k_king_stage1_l2<<<blocks2, 256, 0, ctx->stream>>>(d_shares, dU, (const Fr*)up, mbyl, col0, cols, log_m, mode, gen_tw, has_g,
                                                          g_tw, g_lo2, S);
        ctx->launches += 1;
        ZKG_CUDA(cudaGetLastError());
        phase_mark(ctx, 2);
        return ZKG_OK;
    }
    phase_mark(ctx, 1);
#define KS(LLv) k_king_stage1<LLv><<<blocks, 256, 0, ctx->stream>>>(d_shares, n_recv, dU, nullptr, mbyl, col0, cols, log_m, mode, gen_tw, has_g, g_tw, S)
    if (l == 2) KS(2); else if (l == 4) KS(4); else KS(8);
#undef KS
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    phase_mark(ctx, 2);
    return ZKG_OK;
}

// stage 2 on `cols` output columns: secrets S[c*l..] + rand[c*t..] -> party-major shares out[p*cols + c]
static int32_t king_stage2(zkg_ctx* ctx, const Fr* S, const Fr* d_rand, size_t cols, uint32_t l, Fr* d_out, HostKeep& keep) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    if (cols == 0) return ZKG_OK;
    const int K = (int)(pm->l + pm->t) <= 4 ? 4 : (int)(pm->l + pm->t) <= 8 ? 8 : 16;
    keep.v.push_back(pad_rows(pm->pack, pm->n, pm->l + pm->t, K));
    const Fr* dP;
    ZKG_TRY(upload(ctx, keep.v.back(), &dP));
    ZKG_TRY(launch_pack(ctx, dP, K, pm->n, pm->l, pm->t, S, pm->l, 1, d_rand, pm->t, 1, d_out, 1, cols, cols));
    phase_mark(ctx, 3);
    return ZKG_OK;
}

// ---- king pipeline on device buffers ------------------------------------------------------
// mode_fft: 1 = fft2 + powers + (re)packing (d_fft/d_ifft king), 0 = deg_red king
static int32_t king_stage1(zkg_ctx* ctx, const Fr* d_shares, const uint32_t* parties, uint32_t n_recv, size_t mbyl,
                           size_t col0, size_t cols, uint32_t l, const HFr* gen, const HFr* g, int rearrange,
                           int mode_fft, Fr* S, HostKeep& keep) {
    return king_stage1(ctx, d_shares, parties, n_recv, mbyl, col0, cols, l, gen, g, rearrange, mode_fft, sdest_single(S), keep);
}

static int32_t king_dev(zkg_ctx* ctx, const Fr* d_shares, const uint32_t* parties, uint32_t n_recv, size_t mbyl,
                        uint32_t l, const HFr* gen, const HFr* g, int rearrange, const Fr* d_rand, Fr* d_out,
                        int mode_fft, HostKeep& keep) {
    if (mbyl == 0) return ZKG_OK;
    ZKG_TRY(ctx->ws.reserve(mbyl * l * sizeof(Fr)));
    Fr* S = (Fr*)ctx->ws.p;
    ZKG_TRY(king_stage1(ctx, d_shares, parties, n_recv, mbyl, 0, mbyl, l, gen, g, rearrange, mode_fft, S, keep));
    return king_stage2(ctx, S, d_rand, mbyl, l, d_out, keep);
}

// gather host vectors (one per party) into a party-major device buffer
static int32_t h2d_party_major(zkg_ctx* ctx, const uint64_t* const* by_party, uint32_t np, size_t len, Fr* d) {
    for (uint32_t r = 0; r < np; ++r) {
        ZKG_REQUIRE(by_party[r], "NULL share vector for index %u", r);
        ZKG_TRY(copy_h2d(d + (size_t)r * len, by_party[r], len * 32, ctx->stream));
    }
    return ZKG_OK;
}

static int32_t king_host(int device, const uint64_t* const* shares_by_party, const uint32_t* parties, uint32_t n_recv,
                         size_t mbyl, uint32_t l, const uint64_t* gen, const uint64_t* g, int rearrange,
                         const uint64_t* rand, uint64_t* const* out_by_party, int mode_fft) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    ZKG_REQUIRE(shares_by_party && out_by_party && (mbyl == 0 || rand), "king: NULL argument");
    ZKG_REQUIRE(n_recv >= 1 && n_recv <= pm->n, "king: n_recv = %u out of range", n_recv);
    if (mbyl == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    size_t in_b = align_up((size_t)n_recv * mbyl * 32, 256), rand_b = align_up(mbyl * pm->t * 32, 256);
    size_t out_b = (size_t)pm->n * mbyl * 32;
    ZKG_TRY(ctx->io.reserve(in_b + rand_b + out_b));
    Fr* d_in = (Fr*)ctx->io.p;
    Fr* d_rand = (Fr*)((uint8_t*)ctx->io.p + in_b);
    Fr* d_out = (Fr*)((uint8_t*)ctx->io.p + in_b + rand_b);
    ZKG_TRY(h2d_party_major(ctx, shares_by_party, n_recv, mbyl, d_in));
    ZKG_TRY(copy_h2d(d_rand, rand, mbyl * pm->t * 32, ctx->stream));
    HostKeep keep;
    HFr hgen = mode_fft ? host::h_load(gen) : host::h_one(), hg = mode_fft ? host::h_load(g) : host::h_one();
    ZKG_TRY(king_dev(ctx, d_in, parties, n_recv, mbyl, l, &hgen, &hg, rearrange, d_rand, d_out, mode_fft, keep));
    for (uint32_t p = 0; p < pm->n; ++p) {
        ZKG_REQUIRE(out_by_party[p], "NULL output vector for party %u", p);
        ZKG_TRY(copy_d2h(out_by_party[p], d_out + (size_t)p * mbyl, mbyl * 32, ctx->stream));
    }
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

// ---- fft1 on a device buffer ----------------------------------------------------------------
static int32_t fft1_dev(zkg_ctx* ctx, Fr* d_px, size_t mbyl, uint32_t l, const HFr& gen, const HFr* pre_scale,
                        const Fr* d_mask) {
    ZKG_REQUIRE(l >= 1 && is_pow2(l) && is_pow2(mbyl), "fft1: m/l = %zu and l = %u must be powers of two", mbyl, l);
    ZKG_REQUIRE(ilog2(mbyl * l) <= 28, "fft1: m exceeds the 2-adicity of Fr");
    ZKG_TRY(ctx->ws.reserve(mbyl * sizeof(Fr)));
    Fr* tmp = (Fr*)ctx->ws.p;
    HFr wN = host::h_pow(gen, l);
    // Writing the result back into d_px is safe: with one pass a single block reads the whole
    // vector before it stores; with several passes the last one reads the scratch, and pass 0 has
    // already consumed d_px (stream order), so the shifted stores cannot race with any load.
    return ntt_bitrev_in(ctx, d_px, d_px, tmp, mbyl, wN, 1, pre_scale, d_mask);
}

// ---- fft1 of one lane sharded over G ranks (four-step with one all-to-all) ----------
// X = DFT_N(bitrev_N(px)) with root w = gen^l (fft1(px)[k] = X[(k+1) mod N]).  Write the transform
// index as i = i1 + G*i2: bitrev_N(i) = bitrev_G(i1)*N2 + bitrev_N2(i2), so rank g's contiguous block
// IS the bit-reversed input of the inner size-N2 transform number i1 = bitrev_G(g):
//   local :  Inner_i1 = DFT_N2(bitrev(block), root w^G);  T_i1[k2] = w^(i1*k2) * Inner_i1[k2]
//   all-to-all over k2 ranges (rank d receives columns [d*N2/G, (d+1)*N2/G) of every T_i1)
//   outer :  X[k2 + N2*k1] = sum_i1 (w^N2)^(i1*k1) * T_i1[k2]          (G-point DFT per column)
static int32_t fft1_shard_local(zkg_ctx* ctx, Fr* d_block, size_t N2, uint32_t l, uint32_t G, uint32_t rank,
                                const HFr& gen, const HFr* pre_scale) {
    ZKG_REQUIRE(is_pow2(G) && is_pow2(N2) && is_pow2(l) && rank < G, "fft1_shard: block %zu, l %u, ranks %u must be powers of two", N2, l, G);
    ZKG_REQUIRE(ilog2(N2 * G * l) <= 28, "fft1_shard: m exceeds the 2-adicity of Fr");
    ZKG_TRY(ctx->ws.reserve(N2 * sizeof(Fr)));
    HFr w = host::h_pow(gen, l);
    ZKG_TRY(ntt_bitrev_in(ctx, d_block, d_block, (Fr*)ctx->ws.p, N2, host::h_pow(w, G), 0, pre_scale, nullptr));
    uint32_t i1 = 0;
    for (int b = 0, lg = ilog2(G); b < lg; ++b) i1 |= ((rank >> b) & 1u) << (lg - 1 - b);
    if (i1 && N2 > 1) {
        PowTable tw;
        ZKG_TRY(build_pow_table(ctx, host::h_pow(w, i1), N2, &tw));
        k_distribute_powers<<<(unsigned)((N2 + 255) / 256), 256, 0, ctx->stream>>>(d_block, N2, tw);
        ctx->launches += 1;
        ZKG_CUDA(cudaGetLastError());
    }
    return ZKG_OK;
}

static int32_t fft1_shard_outer(zkg_ctx* ctx, const Fr* d_recv, size_t cnt, size_t N2, uint32_t l, uint32_t G,
                                const HFr& gen, Fr* d_out) {
    ZKG_REQUIRE(is_pow2(G) && G <= 64 && is_pow2(N2) && is_pow2(l), "fft1_shard: ranks %u / block %zu / l %u must be powers of two", G, N2, l);
    const Fr* wG;
    ZKG_TRY(cached_pow_seq(ctx, "shardG", host::h_pow(gen, (uint64_t)l * N2), G, &wG));
    const unsigned blocks = (unsigned)((cnt + 255) / 256);
    if (G == 2) k_fft1_shard_outer_r2<1><<<blocks, 256, 0, ctx->stream>>>(d_recv, cnt, wG, d_out);
    else if (G == 4) k_fft1_shard_outer_r2<2><<<blocks, 256, 0, ctx->stream>>>(d_recv, cnt, wG, d_out);
    else if (G == 8) k_fft1_shard_outer_r2<3><<<blocks, 256, 0, ctx->stream>>>(d_recv, cnt, wG, d_out);
    else k_fft1_shard_outer<<<blocks, 256, 0, ctx->stream>>>(d_recv, cnt, (int)G, ilog2(G), wG, d_out);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}


// local step + exchange in one: inner transform in place, then k_twiddle_scatter into the ranks' receive buffers
static int32_t fft1_shard_local_scatter(zkg_ctx* ctx, Fr* d_block, size_t N2, uint32_t l, uint32_t G, uint32_t rank,
                                        const HFr& gen, const HFr* pre_scale, Fr* const* recv_by_rank) {
    ZKG_REQUIRE(is_pow2(G) && G <= 16 && is_pow2(N2) && N2 >= G && is_pow2(l) && rank < G,
                "fft1_shard: block %zu, l %u, ranks %u must be powers of two (ranks <= 16, block >= ranks)", N2, l, G);
    ZKG_REQUIRE(ilog2(N2 * G * l) <= 28, "fft1_shard: m exceeds the 2-adicity of Fr");
    ZKG_TRY(ctx->ws.reserve(N2 * sizeof(Fr)));
    HFr w = host::h_pow(gen, l);
    ZKG_TRY(ntt_bitrev_in(ctx, d_block, d_block, (Fr*)ctx->ws.p, N2, host::h_pow(w, G), 0, pre_scale, nullptr));
    uint32_t i1 = 0;
    for (int b = 0, lg = ilog2(G); b < lg; ++b) i1 |= ((rank >> b) & 1u) << (lg - 1 - b);
    PowTable tw{nullptr, nullptr};
    if (i1) ZKG_TRY(build_pow_table(ctx, host::h_pow(w, i1), N2, &tw));
    SDest dst;
    for (uint32_t r = 0; r < 16; ++r) dst.p[r] = recv_by_rank[r < G ? r : 0];
    dst.log_seg = 0;
    k_twiddle_scatter<<<(unsigned)((N2 + 255) / 256), 256, 0, ctx->stream>>>(d_block, N2, i1 ? 1 : 0, tw, dst, ilog2(N2 / G), rank);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Single-process multi-GPU entry points (SURVEY.md 8b last line, 8e): one call from the unchanged Rust caller drives
// several GPUs of the box.  The exchange step of each pipeline is done by the kernels themselves, storing into peer
// memory over NVLink (cudaDeviceEnablePeerAccess); without peer access the call fails with ZKG_ERR_NCCL.
// ------------------------------------------------------------------------------------------------------------------
static int32_t enable_peers(const int32_t* devices, int n_dev) {
    for (int a = 0; a < n_dev; ++a) {
        DeviceGuard dg(devices[a]);
        for (int b = 0; b < n_dev; ++b) {
            if (a == b) continue;
            int can = 0;
            ZKG_CUDA(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
            if (!can) {
                set_error("GPU %d cannot map the memory of GPU %d (no NVLink / PCIe peer access); the sharded entry points need it",
                          devices[a], devices[b]);
                return ZKG_ERR_NCCL;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[a], devices[b], cudaGetErrorString(e)); return ZKG_ERR_NCCL; }
        }
    }
    return ZKG_OK;
}

static int32_t check_device_list(const int32_t* devices, int32_t n_dev, const char* who) {
    ZKG_REQUIRE(devices && n_dev >= 1 && n_dev <= 16 && is_pow2((size_t)n_dev), "%s: the device list must hold 1, 2, 4, 8 or 16 GPUs", who);
    for (int a = 0; a < n_dev; ++a)
        for (int b = 0; b < a; ++b) ZKG_REQUIRE(devices[a] != devices[b], "%s: device %d listed twice", who, devices[a]);
    return ZKG_OK;
}

// runs fn(d) for every device on its own host thread (device 0 on the caller's); first failure wins, message carried over
template <class Fn>
static int32_t for_each_device(int n_dev, Fn fn) {
    std::vector<int32_t> rc(n_dev, ZKG_OK);
    std::vector<std::string> msg(n_dev);
    auto run = [&](int d) { rc[d] = fn(d); if (rc[d] != ZKG_OK) msg[d] = zkg_last_error(); };
    std::vector<std::thread> th;
    for (int d = 1; d < n_dev; ++d) th.emplace_back(run, d);
    run(0);
    for (auto& t : th) t.join();
    for (int d = 0; d < n_dev; ++d)
        if (rc[d] != ZKG_OK) { set_error("%s", msg[d].c_str()); return rc[d]; }
    return ZKG_OK;
}

// King closure of fft2_with_rearrange (mode_fft = 1) or deg_red (0) sharded by share columns over the listed GPUs.
static int32_t king_host_sharded(const int32_t* devices, int32_t n_dev, const uint64_t* const* shares_by_party, const uint32_t* parties,
                                 uint32_t n_recv, size_t mbyl, uint32_t l, const uint64_t* gen, const uint64_t* g, int rearrange,
                                 const uint64_t* rand, uint64_t* const* out_by_party, int mode_fft) {
    ZKG_TRY(check_device_list(devices, n_dev, "king_sharded"));
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    if (n_dev == 1 || mbyl < (size_t)n_dev * 256 || !is_pow2(mbyl))
        return king_host(devices[0], shares_by_party, parties, n_recv, mbyl, l, gen, g, rearrange, rand, out_by_party, mode_fft);
    ZKG_REQUIRE(shares_by_party && out_by_party && rand, "king: NULL argument");
    ZKG_REQUIRE(n_recv >= 1 && n_recv <= pm->n, "king: n_recv = %u out of range", n_recv);
    ZKG_TRY(enable_peers(devices, n_dev));
    const size_t cols = mbyl / n_dev, seg = cols * l;                 // columns and S elements per GPU
    std::vector<PooledCtx> pcs(n_dev);
    for (int d = 0; d < n_dev; ++d) ZKG_TRY(pcs[d].acquire(devices[d]));
    std::vector<Fr*> S(n_dev), d_in(n_dev), d_rand(n_dev), d_out(n_dev);
    std::vector<cudaEvent_t> ev(n_dev, nullptr);
    const size_t in_b = align_up((size_t)n_recv * cols * 32, 256), rand_b = align_up(cols * pm->t * 32, 256), out_b = (size_t)pm->n * cols * 32;
    for (int d = 0; d < n_dev; ++d) {                                // carve every GPU's buffers first: stage 1 needs all S pointers
        zkg_ctx* ctx = pcs[d].ctx;
        DeviceGuard dg(ctx->device);
        ZKG_TRY(ctx->ws.reserve(seg * sizeof(Fr)));
        ZKG_TRY(ctx->io.reserve(in_b + rand_b + out_b));
        S[d] = (Fr*)ctx->ws.p;
        d_in[d] = (Fr*)ctx->io.p;
        d_rand[d] = (Fr*)((uint8_t*)ctx->io.p + in_b);
        d_out[d] = (Fr*)((uint8_t*)ctx->io.p + in_b + rand_b);
        ZKG_CUDA(cudaEventCreateWithFlags(&ev[d], cudaEventDisableTiming));
    }
    SDest dst;
    for (int r = 0; r < 16; ++r) dst.p[r] = S[r < n_dev ? r : 0];
    dst.log_seg = ilog2(seg);
    HFr hgen = mode_fft ? host::h_load(gen) : host::h_one(), hg = mode_fft ? host::h_load(g) : host::h_one();
    std::vector<HostKeep> keep(n_dev);
    int32_t rc = for_each_device(n_dev, [&](int d) -> int32_t {     // upload the column slice, unpack + fft2 + powers, scatter to the owners
        zkg_ctx* ctx = pcs[d].ctx;
        DeviceGuard dg(ctx->device);
        const size_t lo = (size_t)d * cols;
        for (uint32_t r = 0; r < n_recv; ++r) {
            ZKG_REQUIRE(shares_by_party[r], "NULL share vector for index %u", r);
            ZKG_TRY(copy_h2d(d_in[d] + (size_t)r * cols, shares_by_party[r] + lo * 4, cols * 32, ctx->stream));
        }
        ZKG_TRY(copy_h2d(d_rand[d], rand + lo * pm->t * 4, cols * pm->t * 32, ctx->stream));
        ZKG_TRY(king_stage1(ctx, d_in[d], parties, n_recv, mbyl, lo, cols, l, &hgen, &hg, rearrange, mode_fft, dst, keep[d]));
        ZKG_CUDA(cudaEventRecord(ev[d], ctx->stream));
        return ZKG_OK;
    });
    if (rc == ZKG_OK)
        rc = for_each_device(n_dev, [&](int d) -> int32_t {         // once EVERY GPU has scattered: pack the own columns, download
            zkg_ctx* ctx = pcs[d].ctx;
            DeviceGuard dg(ctx->device);
            for (int q = 0; q < n_dev; ++q) ZKG_CUDA(cudaStreamWaitEvent(ctx->stream, ev[q], 0));
            ZKG_TRY(king_stage2(ctx, S[d], d_rand[d], cols, l, d_out[d], keep[d]));
            const size_t lo = (size_t)d * cols;
            for (uint32_t p = 0; p < pm->n; ++p) {
                ZKG_REQUIRE(out_by_party[p], "NULL output vector for party %u", p);
                ZKG_TRY(copy_d2h(out_by_party[p] + lo * 4, d_out[d] + (size_t)p * cols, cols * 32, ctx->stream));
            }
            return ZKG_OK;
        });
    for (int d = 0; d < n_dev; ++d) {
        DeviceGuard dg(pcs[d].ctx->device);
        cudaStreamSynchronize(pcs[d].ctx->stream);
        if (ev[d]) cudaEventDestroy(ev[d]);
    }
    return rc;
}

// fft1_in_place of ONE lane over the listed GPUs: block d of the lane goes to GPU d, inner transforms + twiddles, exchange by
// peer stores, G-point outer transforms, and the result is laid back into px in fft1 order (px[k] = X[(k+1) mod N]).
static int32_t fft1_host_sharded(const int32_t* devices, int32_t n_dev, uint64_t* px, size_t mbyl, uint32_t l, const uint64_t* gen,
                                 const uint64_t* pre_scale, const uint64_t* in_mask) {
    ZKG_TRY(check_device_list(devices, n_dev, "fft1_sharded"));
    ZKG_REQUIRE(gen && (mbyl == 0 || px), "fft1_sharded: NULL argument");
    if (mbyl == 0) return ZKG_OK;
    const size_t G = (size_t)n_dev;
    if (n_dev == 1 || !is_pow2(mbyl) || mbyl < G * G * 256) return zkg_fft1_bn254(devices[0], px, mbyl, l, gen, pre_scale, in_mask);
    ZKG_TRY(enable_peers(devices, n_dev));
    const size_t N2 = mbyl / G, cnt = N2 / G;
    std::vector<PooledCtx> pcs(n_dev);
    for (int d = 0; d < n_dev; ++d) ZKG_TRY(pcs[d].acquire(devices[d]));
    std::vector<Fr*> blk(n_dev), recv(n_dev), out(n_dev), msk(n_dev);
    std::vector<cudaEvent_t> ev(n_dev, nullptr);
    const size_t b = align_up(N2 * 32, 256);
    for (int d = 0; d < n_dev; ++d) {
        zkg_ctx* ctx = pcs[d].ctx;
        DeviceGuard dg(ctx->device);
        ZKG_TRY(ctx->io.reserve(4 * b));
        blk[d] = (Fr*)ctx->io.p;
        recv[d] = (Fr*)((uint8_t*)ctx->io.p + b);
        out[d] = (Fr*)((uint8_t*)ctx->io.p + 2 * b);
        msk[d] = (Fr*)((uint8_t*)ctx->io.p + 3 * b);
        ZKG_CUDA(cudaEventCreateWithFlags(&ev[d], cudaEventDisableTiming));
    }
    HFr hgen = host::h_load(gen), hs;
    if (pre_scale) hs = host::h_load(pre_scale);
    // rank d's result row k1 (cnt elements) holds X[d*cnt + j + N2*k1], i.e. px positions starting at (d*cnt + N2*k1 - 1) mod N
    auto for_rows = [&](int d, auto&& fn) -> int32_t {
        for (size_t k1 = 0; k1 < G; ++k1) {
            const size_t x0 = (size_t)d * cnt + N2 * k1;
            Fr* row = nullptr; (void)row;
            if (x0 == 0) {                                           // X[0] wraps to px[N-1]; the rest of the row starts at px[0]
                ZKG_TRY(fn(k1 * cnt, mbyl - 1, (size_t)1));
                ZKG_TRY(fn(k1 * cnt + 1, (size_t)0, cnt - 1));
            } else {
                ZKG_TRY(fn(k1 * cnt, x0 - 1, cnt));
            }
        }
        return ZKG_OK;
    };
    int32_t rc = for_each_device(n_dev, [&](int d) -> int32_t {
        zkg_ctx* ctx = pcs[d].ctx;
        DeviceGuard dg(ctx->device);
        ZKG_TRY(copy_h2d(blk[d], px + (size_t)d * N2 * 4, N2 * 32, ctx->stream));
        if (in_mask)
            ZKG_TRY(for_rows(d, [&](size_t o, size_t pos, size_t len) { return copy_h2d(msk[d] + o, in_mask + pos * 4, len * 32, ctx->stream); }));
        ZKG_TRY(fft1_shard_local_scatter(ctx, blk[d], N2, l, (uint32_t)G, (uint32_t)d, hgen, pre_scale ? &hs : nullptr, recv.data()));
        ZKG_CUDA(cudaEventRecord(ev[d], ctx->stream));
        return ZKG_OK;
    });
    if (rc == ZKG_OK)
        rc = for_each_device(n_dev, [&](int d) -> int32_t {
            zkg_ctx* ctx = pcs[d].ctx;
            DeviceGuard dg(ctx->device);
            for (int q = 0; q < n_dev; ++q) ZKG_CUDA(cudaStreamWaitEvent(ctx->stream, ev[q], 0));
            ZKG_TRY(fft1_shard_outer(ctx, recv[d], cnt, N2, l, (uint32_t)G, hgen, out[d]));
            if (in_mask) {
                k_vec_add<<<(unsigned)((N2 + 255) / 256), 256, 0, ctx->stream>>>(out[d], msk[d], N2);
                ctx->launches += 1;
                ZKG_CUDA(cudaGetLastError());
            }
            return for_rows(d, [&](size_t o, size_t pos, size_t len) { return copy_d2h(px + pos * 4, out[d] + o, len * 32, ctx->stream); });
        });
    for (int d = 0; d < n_dev; ++d) {
        DeviceGuard dg(pcs[d].ctx->device);
        cudaStreamSynchronize(pcs[d].ctx->stream);
        if (ev[d]) cudaEventDestroy(ev[d]);
    }
    return rc;
}

}  // namespace zkg

using namespace zkg;

extern "C" {


int32_t zkg_king_fft2_bn254_sharded(const int32_t* devices, int32_t n_devices, const uint64_t* const* shares_by_party,
                                    const uint32_t* parties, uint32_t n_recv, size_t mbyl, uint32_t l, const uint64_t gen[4],
                                    const uint64_t g[4], int32_t rearrange, const uint64_t* rand, uint64_t* const* out_by_party) {
    ZKG_REQUIRE(gen && g, "king: gen / g are NULL");
    return king_host_sharded(devices, n_devices, shares_by_party, parties, n_recv, mbyl, l, gen, g, rearrange, rand, out_by_party, 1);
}
int32_t zkg_deg_red_king_bn254_sharded(const int32_t* devices, int32_t n_devices, const uint64_t* const* shares_by_party,
                                       const uint32_t* parties, uint32_t n_recv, size_t cols, uint32_t l, const uint64_t* rand,
                                       uint64_t* const* out_by_party) {
    return king_host_sharded(devices, n_devices, shares_by_party, parties, n_recv, cols, l, nullptr, nullptr, 0, rand, out_by_party, 0);
}
int32_t zkg_fft1_bn254_sharded(const int32_t* devices, int32_t n_devices, uint64_t* px, size_t mbyl, uint32_t l, const uint64_t gen[4],
                               const uint64_t* pre_scale, const uint64_t* in_mask) {
    return fft1_host_sharded(devices, n_devices, px, mbyl, l, gen, pre_scale, in_mask);
}

int32_t zkg_fft1_shard_local_scatter_bn254_dev(zkg_ctx* ctx, uint64_t* d_block, size_t block_len, uint32_t l, uint32_t n_ranks,
                                               uint32_t rank, const uint64_t gen[4], const uint64_t* pre_scale,
                                               void* const* d_recv_by_rank) {
    ZKG_REQUIRE(ctx && gen && d_recv_by_rank && (block_len == 0 || d_block), "fft1_shard_local_scatter: NULL argument");
    if (block_len == 0) return ZKG_OK;
    ZKG_REQUIRE(n_ranks >= 1 && n_ranks <= 16, "fft1_shard_local_scatter: %u ranks unsupported (<= 16)", n_ranks);
    for (uint32_t r = 0; r < n_ranks; ++r) ZKG_REQUIRE(d_recv_by_rank[r], "fft1_shard_local_scatter: NULL receive buffer for rank %u", r);
    DeviceGuard dg(ctx->device);
    HFr hs;
    if (pre_scale) hs = host::h_load(pre_scale);
    return fft1_shard_local_scatter(ctx, (Fr*)d_block, block_len, l, n_ranks, rank, host::h_load(gen), pre_scale ? &hs : nullptr,
                                    (Fr* const*)d_recv_by_rank);
}

int32_t zkg_king_stage1_scatter_bn254_dev(zkg_ctx* ctx, const uint64_t* d_shares_local, const uint32_t* parties, uint32_t n_recv,
                                          size_t col0, size_t cols, size_t mbyl, uint32_t l, const uint64_t gen[4],
                                          const uint64_t g[4], int32_t rearrange, void* const* d_S_by_rank, uint32_t n_ranks) {
    ZKG_REQUIRE(ctx && gen && g && d_S_by_rank && (cols == 0 || d_shares_local), "king_stage1_scatter: NULL argument");
    ZKG_REQUIRE(n_ranks >= 1 && n_ranks <= 16 && is_pow2(n_ranks) && is_pow2(mbyl) && mbyl % n_ranks == 0,
                "king_stage1_scatter: m/l = %zu and %u ranks must be powers of two", mbyl, n_ranks);
    DeviceGuard dg(ctx->device);
    SDest dst;
    for (uint32_t r = 0; r < 16; ++r) {
        dst.p[r] = (Fr*)d_S_by_rank[r < n_ranks ? r : 0];
        ZKG_REQUIRE(dst.p[r], "king_stage1_scatter: NULL segment for rank %u", r);
    }
    dst.log_seg = n_ranks == 1 ? 63 : ilog2(mbyl / n_ranks * l);
    HFr hgen = host::h_load(gen), hg = host::h_load(g);
    HostKeep keep;
    ZKG_TRY(king_stage1(ctx, (const Fr*)d_shares_local, parties, n_recv, mbyl, col0, cols, l, &hgen, &hg, rearrange, 1, dst, keep));
    if (!keep.v.empty()) ZKG_CUDA(cudaStreamSynchronize(ctx->stream));     // a Lagrange matrix uploaded from host memory owned by this call
    return ZKG_OK;
}

int32_t zkg_fft1_bn254_dev(zkg_ctx* ctx, uint64_t* d_px, size_t mbyl, uint32_t l, const uint64_t gen[4],
                           const uint64_t* pre_scale, const uint64_t* d_in_mask) {
    ZKG_REQUIRE(ctx && gen && (mbyl == 0 || d_px), "fft1: NULL argument");
    if (mbyl == 0) return ZKG_OK;
    DeviceGuard dg(ctx->device);
    HFr hs;
    if (pre_scale) hs = host::h_load(pre_scale);
    return fft1_dev(ctx, (Fr*)d_px, mbyl, l, host::h_load(gen), pre_scale ? &hs : nullptr, (const Fr*)d_in_mask);
}

int32_t zkg_fft1_shard_local_bn254_dev(zkg_ctx* ctx, uint64_t* d_block, size_t block_len, uint32_t l, uint32_t n_ranks,
                                       uint32_t rank, const uint64_t gen[4], const uint64_t* pre_scale) {
    ZKG_REQUIRE(ctx && gen && (block_len == 0 || d_block), "fft1_shard_local: NULL argument");
    if (block_len == 0) return ZKG_OK;
    DeviceGuard dg(ctx->device);
    HFr hs;
    if (pre_scale) hs = host::h_load(pre_scale);
    return fft1_shard_local(ctx, (Fr*)d_block, block_len, l, n_ranks, rank, host::h_load(gen), pre_scale ? &hs : nullptr);
}

int32_t zkg_fft1_shard_outer_bn254_dev(zkg_ctx* ctx, const uint64_t* d_recv, size_t cols, size_t block_len, uint32_t l,
                                       uint32_t n_ranks, const uint64_t gen[4], uint64_t* d_out) {
    ZKG_REQUIRE(ctx && gen && (cols == 0 || (d_recv && d_out)), "fft1_shard_outer: NULL argument");
    if (cols == 0) return ZKG_OK;
    DeviceGuard dg(ctx->device);
    return fft1_shard_outer(ctx, (const Fr*)d_recv, cols, block_len, l, n_ranks, host::h_load(gen), (Fr*)d_out);
}

int32_t zkg_fft1_bn254(int32_t device, uint64_t* px, size_t mbyl, uint32_t l, const uint64_t gen[4],
                       const uint64_t* pre_scale, const uint64_t* in_mask) {
    ZKG_REQUIRE(gen && (mbyl == 0 || px), "fft1: NULL argument");
    if (mbyl == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    size_t bytes = align_up(mbyl * 32, 256);
    ZKG_TRY(ctx->io.reserve(2 * bytes));
    Fr* d_px = (Fr*)ctx->io.p;
    Fr* d_mask = (Fr*)((uint8_t*)ctx->io.p + bytes);
    ZKG_TRY(copy_h2d(d_px, px, mbyl * 32, ctx->stream));
    if (in_mask) ZKG_TRY(copy_h2d(d_mask, in_mask, mbyl * 32, ctx->stream));
    HFr hs;
    if (pre_scale) hs = host::h_load(pre_scale);
    ZKG_TRY(fft1_dev(ctx, d_px, mbyl, l, host::h_load(gen), pre_scale ? &hs : nullptr, in_mask ? d_mask : nullptr));
    ZKG_TRY(copy_d2h(px, d_px, mbyl * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_king_fft2_bn254(int32_t device, const uint64_t* const* shares_by_party, const uint32_t* parties,
                            uint32_t n_recv, size_t mbyl, uint32_t l, const uint64_t gen[4], const uint64_t g[4],
                            int32_t rearrange, const uint64_t* rand, uint64_t* const* out_by_party) {
    ZKG_REQUIRE(gen && g, "king_fft2: NULL gen/g");
    return king_host(device, shares_by_party, parties, n_recv, mbyl, l, gen, g, rearrange, rand, out_by_party, 1);
}

int32_t zkg_king_fft2_bn254_dev(zkg_ctx* ctx, const uint64_t* d_shares, const uint32_t* parties, uint32_t n_recv,
                                size_t mbyl, uint32_t l, const uint64_t gen[4], const uint64_t g[4], int32_t rearrange,
                                const uint64_t* d_rand, uint64_t* d_out) {
    ZKG_REQUIRE(ctx && gen && g && (mbyl == 0 || (d_shares && d_rand && d_out)), "king_fft2: NULL argument");
    DeviceGuard dg(ctx->device);
    HostKeep keep;
    HFr hgen = host::h_load(gen), hg = host::h_load(g);
    ZKG_TRY(king_dev(ctx, (const Fr*)d_shares, parties, n_recv, mbyl, l, &hgen, &hg, rearrange, (const Fr*)d_rand, (Fr*)d_out, 1, keep));
    // the async uploads read host matrices owned by `keep`: wait for them before it goes away
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_king_stage1_bn254_dev(zkg_ctx* ctx, const uint64_t* d_shares_local, const uint32_t* parties, uint32_t n_recv,
                                  size_t col0, size_t cols, size_t mbyl, uint32_t l, const uint64_t gen[4],
                                  const uint64_t g[4], int32_t rearrange, uint64_t* d_S_full) {
    ZKG_REQUIRE(ctx && gen && g && (cols == 0 || (d_shares_local && d_S_full)), "king_stage1: NULL argument");
    DeviceGuard dg(ctx->device);
    HostKeep keep;
    HFr hgen = host::h_load(gen), hg = host::h_load(g);
    return king_stage1(ctx, (const Fr*)d_shares_local, parties, n_recv, mbyl, col0, cols, l, &hgen, &hg, rearrange, 1,
                       (Fr*)d_S_full, keep);
}

int32_t zkg_king_stage2_bn254_dev(zkg_ctx* ctx, const uint64_t* d_S_local, const uint64_t* d_rand_local, size_t cols,
                                  uint32_t l, uint64_t* d_out_local) {
    ZKG_REQUIRE(ctx && (cols == 0 || (d_S_local && d_rand_local && d_out_local)), "king_stage2: NULL argument");
    DeviceGuard dg(ctx->device);
    HostKeep keep;
    return king_stage2(ctx, (const Fr*)d_S_local, (const Fr*)d_rand_local, cols, l, (Fr*)d_out_local, keep);
}

int32_t zkg_dpp_king_bn254(int32_t device, const uint64_t* const* shares_by_party, const uint32_t* parties, uint32_t n_recv,
                           size_t cols, uint32_t l, const uint64_t* rand, uint64_t* const* out_by_party) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    ZKG_REQUIRE(shares_by_party && out_by_party && (cols == 0 || rand), "dpp_king: NULL argument");
    ZKG_REQUIRE(n_recv >= 1 && n_recv <= pm->n, "dpp_king: n_recv = %u out of range", n_recv);
    if (cols == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t m = cols * l, n_chunks = (m + DPP_CHUNK - 1) / DPP_CHUNK;
    size_t in_b = align_up((size_t)n_recv * 2 * cols * 32, 256), rand_b = align_up(cols * pm->t * 32, 256);
    size_t out_b = align_up((size_t)pm->n * cols * 32, 256);
    ZKG_TRY(ctx->io.reserve(in_b + rand_b + out_b + 256));
    Fr* d_in = (Fr*)ctx->io.p;
    Fr* d_rand = (Fr*)((uint8_t*)ctx->io.p + in_b);
    Fr* d_out = (Fr*)((uint8_t*)ctx->io.p + in_b + rand_b);
    int* d_err = (int*)((uint8_t*)ctx->io.p + in_b + rand_b + out_b);
    // workspace: numden secrets (2m) | scratch (m) | chunk totals
    ZKG_TRY(ctx->ws.reserve((3 * m + n_chunks + 8) * sizeof(Fr)));
    Fr* S = (Fr*)ctx->ws.p;
    Fr* scratch = S + 2 * m;
    Fr* tot = scratch + m;
    ZKG_TRY(h2d_party_major(ctx, shares_by_party, n_recv, 2 * cols, d_in));
    ZKG_TRY(copy_h2d(d_rand, rand, cols * pm->t * 32, ctx->stream));
    ZKG_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
    HostKeep keep;
    HFr one = host::h_one();
    // unpack all 2*cols columns (num columns, then den columns): dpp/mod.rs:44-53
    ZKG_TRY(king_stage1(ctx, d_in, parties, n_recv, 2 * cols, 0, 2 * cols, l, &one, &one, 0, 0, S, keep));
    unsigned cb = (unsigned)((n_chunks + 127) / 128);
    k_dpp_divide<<<cb, 128, 0, ctx->stream>>>(S, S + m, scratch, m, d_err);                 // :55-58
    k_dpp_scan_local<<<cb, 128, 0, ctx->stream>>>(S, m, tot);                               // :62-66
    k_dpp_scan_totals<<<1, 1024, 0, ctx->stream>>>(tot, n_chunks);
    k_dpp_scan_apply<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(S, m, tot);
    ctx->launches += 4;
    ZKG_CUDA(cudaGetLastError());
    ZKG_TRY(king_stage2(ctx, S, d_rand, cols, l, d_out, keep));                             // pack_vec, :71
    int h_err = 0;
    ZKG_TRY(copy_d2h(&h_err, d_err, sizeof(int), ctx->stream));
    for (uint32_t p = 0; p < pm->n; ++p) {
        ZKG_REQUIRE(out_by_party[p], "NULL output vector for party %u", p);
        ZKG_TRY(copy_d2h(out_by_party[p], d_out + (size_t)p * cols, cols * 32, ctx->stream));
    }
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    ZKG_REQUIRE(h_err == 0, "dpp_king: a denominator is zero (the reference's inverse().unwrap() would panic)");
    return ZKG_OK;
}

int32_t zkg_deg_red_king_bn254(int32_t device, const uint64_t* const* shares_by_party, const uint32_t* parties,
                               uint32_t n_recv, size_t cols, uint32_t l, const uint64_t* rand,
                               uint64_t* const* out_by_party) {
    return king_host(device, shares_by_party, parties, n_recv, cols, l, nullptr, nullptr, 0, rand, out_by_party, 0);
}

int32_t zkg_deg_red_king_bn254_dev(zkg_ctx* ctx, const uint64_t* d_shares, const uint32_t* parties, uint32_t n_recv,
                                   size_t cols, uint32_t l, const uint64_t* d_rand, uint64_t* d_out) {
    ZKG_REQUIRE(ctx && (cols == 0 || (d_shares && d_rand && d_out)), "deg_red_king: NULL argument");
    DeviceGuard dg(ctx->device);
    HostKeep keep;
    HFr one = host::h_one();
    ZKG_TRY(king_dev(ctx, (const Fr*)d_shares, parties, n_recv, cols, l, &one, &one, 0, (const Fr*)d_rand, (Fr*)d_out, 0, keep));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

// ---- PSS, column-major batches ---------------------------------------------------------------
static int32_t pss_host(int device, uint32_t l, int which, const uint64_t* in, const uint64_t* rand, uint64_t* out,
                        size_t cols) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    ZKG_REQUIRE(cols == 0 || (in && out), "pss: NULL argument");
    if (cols == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t n = pm->n, t = pm->t;
    size_t in_elems = which == 0 ? cols * l : cols * n, out_elems = which == 0 ? cols * n : cols * l;
    size_t in_b = align_up(in_elems * 32, 256), rand_b = align_up(cols * t * 32, 256);
    ZKG_TRY(ctx->io.reserve(in_b + rand_b + out_elems * 32));
    Fr* d_in = (Fr*)ctx->io.p;
    Fr* d_rand = (Fr*)((uint8_t*)ctx->io.p + in_b);
    Fr* d_out = (Fr*)((uint8_t*)ctx->io.p + in_b + rand_b);
    ZKG_TRY(copy_h2d(d_in, in, in_elems * 32, ctx->stream));
    if (which == 0 && rand) ZKG_TRY(copy_h2d(d_rand, rand, cols * t * 32, ctx->stream));
    HostKeep keep;
    const Fr* dM;
    if (which == 0) {
        const int K = (int)(l + t) <= 4 ? 4 : (int)(l + t) <= 8 ? 8 : 16;
        keep.v.push_back(pad_rows(pm->pack, (int)n, (int)(l + t), K));
        ZKG_TRY(upload(ctx, keep.v.back(), &dM));
        // det_pack (rand == NULL): the t padding entries are zero, so only the first l columns contribute
        ZKG_TRY(launch_pack(ctx, dM, K, (int)n, (int)l, rand ? (int)t : 0, d_in, l, 1, d_rand, t, 1, d_out, n, 1, cols));
    } else {
        ZKG_TRY(upload(ctx, which == 1 ? pm->unpack : pm->unpack2, &dM));
        ZKG_TRY(launch_unpack(ctx, dM, (int)l, (int)n, d_in, n, 1, d_out, l, 1, cols));
    }
    ZKG_TRY(copy_d2h(out, d_out, out_elems * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

// Vector packing into per-party share vectors (party-major output), the two layouts the reference uses:
//   layout 0  chunks of l consecutive values, the last one zero-padded   (pack_from_witness, sha256.rs:131-156;
//             pack_vec + transpose, utils/pack.rs:8-20)
//   layout 1  bit-reverse x, then column i packs (x'[i], x'[i + m/l], ...)   (QAP::pss, groth16/src/qap.rs:99-112)
static int32_t pack_vec_host(int device, uint32_t l, int layout, const uint64_t* x, size_t len, const uint64_t* rand,
                             uint64_t* const* out_by_party) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    ZKG_REQUIRE(layout == 0 || layout == 1, "pack_vec: layout %d unknown", layout);
    ZKG_REQUIRE(out_by_party && (len == 0 || (x && rand)), "pack_vec: NULL argument");
    if (len == 0) return ZKG_OK;
    ZKG_REQUIRE(layout == 0 || (is_pow2(len) && len >= l), "pack_vec: bit-reversed layout needs a power-of-two length >= l, got %zu", len);
    const size_t n = pm->n, t = pm->t;
    const size_t cols = (len + l - 1) / l, padded = cols * l;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t x_b = align_up(padded * 32, 256), rand_b = align_up(cols * t * 32, 256), out_b = align_up(cols * n * 32, 256);
    ZKG_TRY(ctx->io.reserve(2 * x_b + rand_b + out_b));
    Fr* d_x = (Fr*)ctx->io.p;
    Fr* d_rev = (Fr*)((uint8_t*)ctx->io.p + x_b);
    Fr* d_rand = (Fr*)((uint8_t*)ctx->io.p + 2 * x_b);
    Fr* d_out = (Fr*)((uint8_t*)ctx->io.p + 2 * x_b + rand_b);
    ZKG_TRY(copy_h2d(d_x, x, len * 32, ctx->stream));
    if (padded > len) ZKG_CUDA(cudaMemsetAsync(d_x + len, 0, (padded - len) * 32, ctx->stream));
    ZKG_TRY(copy_h2d(d_rand, rand, cols * t * 32, ctx->stream));
    HostKeep keep;
    const Fr* dM;
    const int K = (int)(l + t) <= 4 ? 4 : (int)(l + t) <= 8 ? 8 : 16;
    keep.v.push_back(pad_rows(pm->pack, (int)n, (int)(l + t), K));
    ZKG_TRY(upload(ctx, keep.v.back(), &dM));
    if (layout == 1) {
        k_bitrev<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(d_x, d_rev, ilog2(len));
        ctx->launches += 1;
        ZKG_CUDA(cudaGetLastError());
        ZKG_TRY(launch_pack(ctx, dM, K, (int)n, (int)l, (int)t, d_rev, 1, cols, d_rand, t, 1, d_out, 1, cols, cols));
    } else {
        ZKG_TRY(launch_pack(ctx, dM, K, (int)n, (int)l, (int)t, d_x, l, 1, d_rand, t, 1, d_out, 1, cols, cols));
    }
    for (size_t p = 0; p < n; ++p) {
        ZKG_REQUIRE(out_by_party[p], "NULL output vector for party %zu", p);
        ZKG_TRY(copy_d2h(out_by_party[p], d_out + p * cols, cols * 32, ctx->stream));
    }
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_pss_pack_vec_bn254_fr(int32_t device, uint32_t l, int32_t layout, const uint64_t* x, size_t len,
                                  const uint64_t* rand, uint64_t* const* out_by_party) {
    return pack_vec_host(device, l, layout, x, len, rand, out_by_party);
}

int32_t zkg_pss_pack_bn254_fr(int32_t device, uint32_t l, const uint64_t* secrets, const uint64_t* rand, uint64_t* shares, size_t cols) {
    return pss_host(device, l, 0, secrets, rand, shares, cols);
}
int32_t zkg_pss_unpack_bn254_fr(int32_t device, uint32_t l, const uint64_t* shares, uint64_t* secrets, size_t cols) {
    return pss_host(device, l, 1, shares, nullptr, secrets, cols);
}
int32_t zkg_pss_unpack2_bn254_fr(int32_t device, uint32_t l, const uint64_t* shares, uint64_t* secrets, size_t cols) {
    return pss_host(device, l, 2, shares, nullptr, secrets, cols);
}


// FftMask::sample (dist-primitives/src/dfft/mod.rs:30-85) with the random draws as inputs, every step on the device:
// one upload of the m mask values and the 2 x (m/l * t) packing draws, one download of the 2 x n share vectors (the
// Python mirror of round 1 crossed PCIe between each of its five steps).
//   in shares  = transpose(pack_vec(mask))                                            :41-42
//   S          = -(fft2_in_place(mask, gen) .* g^i)                                   :44-51
//   out shares = rearrange ? pack of the bit-reversed, stride-m/l columns : pack_vec  :55-74
// fft2 + powers + the rearrange permutation are one launch of the king's stage-1 kernel reading the mask directly.
// mode_fft == 0 is DegRedMask::sample over Fr with gen = 1 (utils/deg_red.rs:40-66): S = -mask, no transform.
static int32_t mask_sample_host(int device, int mode_fft, int rearrange, const uint64_t* g, const uint64_t* gen, size_t m, uint32_t l,
                                const uint64_t* mask_values, const uint64_t* rand_in, const uint64_t* rand_out,
                                uint64_t* const* in_by_party, uint64_t* const* out_by_party) {
    const host::PssMatrices* pm = pss_get(l);
    ZKG_REQUIRE(pm, "packing factor l = %u unsupported (2, 4, 8)", l);
    ZKG_REQUIRE(in_by_party && out_by_party && (m == 0 || (mask_values && rand_in && rand_out)), "mask_sample: NULL argument");
    ZKG_REQUIRE(!mode_fft || (g && gen), "fft_mask_sample: g / gen are NULL");
    if (m == 0) return ZKG_OK;
    ZKG_REQUIRE(m % l == 0, "mask_sample: %zu values is not a multiple of l = %u", m, l);
    ZKG_REQUIRE(!mode_fft || (is_pow2(m) && m >= l && ilog2(m) <= 28), "fft_mask_sample: m = %zu must be a power of two in [l, 2^28]", m);
    const size_t n = pm->n, t = pm->t, cols = m / l;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t m_b = align_up(m * 32, 256), r_b = align_up(cols * t * 32, 256), o_b = align_up(cols * n * 32, 256);
    ZKG_TRY(ctx->io.reserve(2 * m_b + 2 * r_b + 2 * o_b));
    uint8_t* d = (uint8_t*)ctx->io.p;
    Fr *d_mask = (Fr*)d, *d_S = (Fr*)(d + m_b), *d_ri = (Fr*)(d + 2 * m_b), *d_ro = (Fr*)(d + 2 * m_b + r_b);
    Fr *d_in = (Fr*)(d + 2 * m_b + 2 * r_b), *d_out = (Fr*)(d + 2 * m_b + 2 * r_b + o_b);
    ZKG_TRY(copy_h2d(d_mask, mask_values, m * 32, ctx->stream));
    ZKG_TRY(copy_h2d(d_ri, rand_in, cols * t * 32, ctx->stream));
    ZKG_TRY(copy_h2d(d_ro, rand_out, cols * t * 32, ctx->stream));
    HostKeep keep;
    ZKG_TRY(king_stage2(ctx, d_mask, d_ri, cols, l, d_in, keep));                       // in shares: consecutive chunks
    if (mode_fft) {
        PowTable gen_tw{nullptr, nullptr}, g_tw{nullptr, nullptr};
        HFr hgen = host::h_load(gen), hg = host::h_load(g);
        ZKG_TRY(build_pow_table(ctx, hgen, m, &gen_tw));
        const int has_g = !(hg == host::h_one());
        if (has_g) ZKG_TRY(build_pow_table(ctx, hg, m, &g_tw));
        const unsigned blocks = (unsigned)((cols + 255) / 256);
        const int log_m = ilog2(m), mode = rearrange ? 1 : 0;
#define KS(LLv) k_king_stage1<LLv><<<blocks, 256, 0, ctx->stream>>>(nullptr, 0, nullptr, d_mask, cols, 0, cols, log_m, mode, gen_tw, has_g, g_tw, sdest_single(d_S))
        if (l == 2) KS(2); else if (l == 4) KS(4); else KS(8);
#undef KS
        ctx->launches += 1;
    } else {
        ZKG_CUDA(cudaMemcpyAsync(d_S, d_mask, m * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    k_negate<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(d_S, m);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    ZKG_TRY(king_stage2(ctx, d_S, d_ro, cols, l, d_out, keep));
    for (size_t p = 0; p < n; ++p) {
        ZKG_REQUIRE(in_by_party[p] && out_by_party[p], "NULL output vector for party %zu", p);
        ZKG_TRY(copy_d2h(in_by_party[p], d_in + p * cols, cols * 32, ctx->stream));
        ZKG_TRY(copy_d2h(out_by_party[p], d_out + p * cols, cols * 32, ctx->stream));
    }
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_fft_mask_sample_bn254(int32_t device, int32_t rearrange, const uint64_t g[4], const uint64_t gen[4], size_t m, uint32_t l,
                                  const uint64_t* mask_values, const uint64_t* rand_in, const uint64_t* rand_out,
                                  uint64_t* const* in_by_party, uint64_t* const* out_by_party) {
    return mask_sample_host(device, 1, rearrange, g, gen, m, l, mask_values, rand_in, rand_out, in_by_party, out_by_party);
}
int32_t zkg_deg_red_mask_sample_bn254(int32_t device, size_t num, uint32_t l, const uint64_t* mask_values, const uint64_t* rand_in,
                                      const uint64_t* rand_out, uint64_t* const* in_by_party, uint64_t* const* out_by_party) {
    return mask_sample_host(device, 0, 0, nullptr, nullptr, num * l, l, mask_values, rand_in, rand_out, in_by_party, out_by_party);
}

// ---- stand-alone pieces ----------------------------------------------------------------------
int32_t zkg_fft2_bn254(int32_t device, uint64_t* s1, size_t m, uint32_t l, const uint64_t gen[4]) {
    ZKG_REQUIRE(gen && (m == 0 || s1), "fft2: NULL argument");
    ZKG_REQUIRE(l == 2 || l == 4 || l == 8, "packing factor l = %u unsupported (2, 4, 8)", l);
    if (m == 0) return ZKG_OK;
    ZKG_REQUIRE(is_pow2(m) && m >= l && ilog2(m) <= 28, "fft2: m = %zu must be a power of two in [l, 2^28]", m);
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    ZKG_TRY(ctx->io.reserve(2 * m * 32));
    Fr* d_in = (Fr*)ctx->io.p;
    Fr* d_out = d_in + m;
    ZKG_TRY(copy_h2d(d_in, s1, m * 32, ctx->stream));
    PowTable gen_tw, none{nullptr, nullptr};
    ZKG_TRY(build_pow_table(ctx, host::h_load(gen), m, &gen_tw));
    size_t mbyl = m / l;
    unsigned blocks = (unsigned)((mbyl + 255) / 256);
    int log_m = ilog2(m);
    if (l == 2) k_king_stage1<2><<<blocks, 256, 0, ctx->stream>>>(nullptr, 0, nullptr, d_in, mbyl, 0, mbyl, log_m, 3, gen_tw, 0, none, sdest_single(d_out));
    else if (l == 4) k_king_stage1<4><<<blocks, 256, 0, ctx->stream>>>(nullptr, 0, nullptr, d_in, mbyl, 0, mbyl, log_m, 3, gen_tw, 0, none, sdest_single(d_out));
    else k_king_stage1<8><<<blocks, 256, 0, ctx->stream>>>(nullptr, 0, nullptr, d_in, mbyl, 0, mbyl, log_m, 3, gen_tw, 0, none, sdest_single(d_out));
    ZKG_CUDA(cudaGetLastError());
    ZKG_TRY(copy_d2h(s1, d_out, m * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_distribute_powers_bn254(int32_t device, uint64_t* v, size_t n, const uint64_t g[4]) {
    ZKG_REQUIRE(g && (n == 0 || v), "distribute_powers: NULL argument");
    if (n == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    ZKG_TRY(ctx->io.reserve(n * 32));
    Fr* d = (Fr*)ctx->io.p;
    ZKG_TRY(copy_h2d(d, v, n * 32, ctx->stream));
    PowTable tw;
    ZKG_TRY(build_pow_table(ctx, host::h_load(g), n, &tw));
    k_distribute_powers<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d, n, tw);
    ZKG_CUDA(cudaGetLastError());
    ZKG_TRY(copy_d2h(v, d, n * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_bitrev_bn254(int32_t device, uint64_t* v, size_t n) {
    ZKG_REQUIRE(n == 0 || v, "bitrev: NULL argument");
    if (n <= 1) return ZKG_OK;
    ZKG_REQUIRE(is_pow2(n), "bitrev: n = %zu is not a power of two", n);
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    ZKG_TRY(ctx->io.reserve(2 * n * 32));
    Fr* d = (Fr*)ctx->io.p;
    ZKG_TRY(copy_h2d(d, v, n * 32, ctx->stream));
    k_bitrev<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d, d + n, ilog2(n));
    ZKG_CUDA(cudaGetLastError());
    ZKG_TRY(copy_d2h(v, d + n, n * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

int32_t zkg_fr_fft_bn254(int32_t device, uint64_t* v, size_t n, const uint64_t* offset, int32_t inverse) {
    ZKG_REQUIRE(n == 0 || v, "fr_fft: NULL argument");
    if (n == 0) return ZKG_OK;
    ZKG_REQUIRE(is_pow2(n) && ilog2(n) <= 27, "fr_fft: n = %zu must be a power of two <= 2^27", n);
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    ZKG_TRY(ctx->io.reserve(3 * n * 32));
    Fr* d = (Fr*)ctx->io.p;
    Fr* d_rev = d + n;
    Fr* d_tmp = d + 2 * n;
    ZKG_TRY(copy_h2d(d, v, n * 32, ctx->stream));
    HFr w = host::h_root_of_unity(n), one = host::h_one();
    HFr off = offset ? host::h_load(offset) : one;
    bool coset = !(off == one);
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (!inverse) {
        if (coset) {
            PowTable tw;
            ZKG_TRY(build_pow_table(ctx, off, n, &tw));
            k_distribute_powers<<<blocks, 256, 0, ctx->stream>>>(d, n, tw);
        }
        k_bitrev<<<blocks, 256, 0, ctx->stream>>>(d, d_rev, ilog2(n));
        ZKG_TRY(ntt_bitrev_in(ctx, d_rev, d, d_tmp, n, w, 0, nullptr, nullptr));
    } else {
        k_bitrev<<<blocks, 256, 0, ctx->stream>>>(d, d_rev, ilog2(n));
        ZKG_TRY(ntt_bitrev_in(ctx, d_rev, d, d_tmp, n, host::h_inv(w), 0, nullptr, nullptr));
        HFr n_inv = host::h_inv(host::h_from_u64(n));
        PowTable tw{nullptr, nullptr};
        if (coset) ZKG_TRY(build_pow_table(ctx, host::h_inv(off), n, &tw));
        k_scale_powers<<<blocks, 256, 0, ctx->stream>>>(d, n, to_arg(n_inv), coset ? 1 : 0, tw);
    }
    ZKG_CUDA(cudaGetLastError());
    ZKG_TRY(copy_d2h(v, d, n * 32, ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

}  // extern "C"
