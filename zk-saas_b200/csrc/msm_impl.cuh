// msm_impl.cuh -- BN254 G1/G2 multi-scalar multiplication on sm_100a.
//
// Replaces `G::msm(bases, scalars)` (ark-ec 0.4.2 VariableBaseMSM::msm -> msm_bigint_wnaf) at
// dist-primitives/src/dmsm/mod.rs:73 -- the hottest loop of the reference (callers:
// groth16/src/prove.rs:52,106,154,209,219).
//
// Pipeline (all on the context's stream, no host round-trips):
//   k_digits      scalars: Montgomery -> canonical, signed radix-2^c digits, per-(window,bucket) histogram
//   k_scan        exclusive scan of the histogram inside each window  -> bucket start offsets
//   k_scatter     counting-sort the point indices of every window by bucket
//   k_accumulate  one thread per (window,bucket): gather affine bases (128-bit loads), XYZZ mixed adds
//   k_reduce_lvl  multi-level running-sum reduction  sum_b (b+1)*B_b  per window (see below)
//   k_final       Horner over windows (c doublings each) + normalisation to affine
// The arithmetic is integer-pipe bound (IMAD.WIDE), not HBM bound: per point 96 B of traffic vs
// ~10 field multiplications per window.
#pragma once
#include "common.cuh"
#include "ec.cuh"
#include "msm_common.cuh"
#include <string>
#include <thread>

namespace zkg {

// ------------------------------------------------------------------------------------------
// loads / stores
// ------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T load_vec(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = __ldg(s + i);
    return r;
}
template <class T>
__device__ __forceinline__ T load_vec_rw(const T* p) {     // data written earlier by this launch sequence
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
    return r;
}
template <class T>
__device__ __forceinline__ void store_vec(T* p, const T& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
}

// ------------------------------------------------------------------------------------------
// arkworks Affine images (72 B / 136 B, `infinity` flag after the coordinates) -> packed (x,y)
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void k_pack_bases(const uint8_t* __restrict__ ark, size_t stride, size_t n, Affine<F>* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = ark + i * stride;
    Affine<F> a;
    uint32_t* w = reinterpret_cast<uint32_t*>(&a);
    constexpr int WORDS = sizeof(Affine<F>) / 4;
    if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
        const uint2* s = reinterpret_cast<const uint2*>(p);
#pragma unroll
        for (int k = 0; k < WORDS / 2; ++k) { uint2 v = s[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
    } else {
#pragma unroll
        for (int k = 0; k < WORDS; ++k)
            w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
    }
    if (p[sizeof(Affine<F>)] != 0) a = Affine<F>::inf();
    store_vec(out + i, a);
}

// ------------------------------------------------------------------------------------------
// scalars -> signed digits + histogram
// ------------------------------------------------------------------------------------------
// Emits the digits of windows [w_lo, w_hi) only (a chunk of the sort pipeline may cover a window range); the carry chain
// is recomputed from window 0, which costs a few shifts per skipped window.  Grid-stride: a sort that runs UNDER another
// chunk's accumulation is launched with a couple of blocks per SM, so that it takes a thin slice of every SM instead of
// every slot the accumulation frees (the side stream has the higher priority).
template <int THIN>      // THIN: at most 32 registers, so that a 128-thread block fits beside five resident k_accumulate blocks
static __global__ void __launch_bounds__(THIN ? 128 : 256, THIN ? 16 : 1) k_digits(const Fr* __restrict__ scalars, size_t n, int c, int W, int w_lo, int w_hi, uint32_t bstride,
                         uint32_t* __restrict__ digits, uint32_t* __restrict__ counts) {
    const uint32_t half = 1u << (c - 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        Fr s = fp_from_mont(load_vec(scalars + i));      // into_bigint()
        uint32_t carry = 0;
        for (int w = 0; w < w_hi; ++w) {
            uint32_t coef = msm_window_bits(s.v, w, c) + carry;
            uint32_t neg = 0;
            carry = 0;
            if (w != W - 1 && coef >= half) {
                if (coef == (1u << c)) coef = 0;
                else { coef = (1u << c) - coef; neg = 1; }
                carry = 1;
            }
            if (w < w_lo) continue;
            uint32_t d = MSM_DIGIT_NONE;
            if (coef != 0) {
                d = (coef - 1) | (neg << 31);
                atomicAdd(&counts[(size_t)(w - w_lo) * bstride + (coef - 1)], 1u);      // bstride = 0: windows share the buckets
            }
            digits[(size_t)(w - w_lo) * n + i] = d;
        }
    }
}

// Exclusive scan of the bucket counts inside each window, in three small launches so that a merged
// plan (ONE window of up to 2^22 buckets) does not serialise on a single block:
//   k_scan_partial  per (window, tile of SCAN_TILE counts): tile total
//   k_scan_tiles    per window: exclusive scan of its tile totals (one block; <= 2^22/4096 = 1024 tiles)
//   k_scan_final    per (window, tile): exclusive scan inside the tile + tile offset
static constexpr uint32_t SCAN_TILE = 4096;      // counts per tile; 256 threads x 16 consecutive counts
static constexpr uint32_t SCAN_PER_THREAD = 16;

static __global__ void k_scan_partial(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t tiles, uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t red[256];
    const uint32_t w = blockIdx.y, tile = blockIdx.x;
    const uint32_t* in = counts + (size_t)w * nb;
    uint32_t lo = tile * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t s = 0;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER_THREAD; ++k) s += lo + k < nb ? in[lo + k] : 0u;
    red[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sum[(size_t)w * tiles + tile] = red[0];
}
static __global__ void k_scan_tiles(uint32_t* __restrict__ tile_sum, uint32_t tiles) {
    __shared__ uint32_t part[1024];
    uint32_t* v = tile_sum + (size_t)blockIdx.x * tiles;
    uint32_t x = threadIdx.x < tiles ? v[threadIdx.x] : 0u;
    part[threadIdx.x] = x;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t a = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += a;
        __syncthreads();
    }
    if (threadIdx.x < tiles) v[threadIdx.x] = part[threadIdx.x] - x;      // exclusive
}
static __global__ void k_scan_final(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t tiles,
                                    const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ cursor) {
    __shared__ uint32_t part[256];
    const uint32_t w = blockIdx.y, tile = blockIdx.x;
    const uint32_t* in = counts + (size_t)w * nb;
    uint32_t* out = cursor + (size_t)w * nb;
    uint32_t lo = tile * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t c[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER_THREAD; ++k) { c[k] = lo + k < nb ? in[lo + k] : 0u; s += c[k]; }
    part[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t off = 1; off < 256; off <<= 1) {
        uint32_t a = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += a;
        __syncthreads();
    }
    uint32_t base = tile_off[(size_t)w * tiles + tile] + part[threadIdx.x] - s;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER_THREAD; ++k) {
        if (lo + k < nb) out[lo + k] = base;
        base += c[k];
    }
}

// bstride / sstride / ioff = (nb, n, 0) for per-window buckets, (0, 0, n_total) for merged windows, where the
// sorted entry indexes the window-shifted base table (w * n_total + point).  Windows are counted from w_lo.
// One thread takes a point through the chunk's windows, four at a time: the four returned atomics are in flight
// together, so a launch of a few hundred blocks (a sort hidden under an accumulation) still keeps the L2 busy.
template <int ILP>
static __global__ void __launch_bounds__(256, ILP > 4 ? 4 : 8) k_scatter(const uint32_t* __restrict__ digits, size_t n, int wg, uint32_t bstride, size_t sstride, size_t ioff,
                          size_t point0, int w_lo, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll 1
        for (int w0 = 0; w0 < wg; w0 += ILP) {
            uint32_t d[ILP], pos[ILP];
#pragma unroll
            for (int k = 0; k < ILP; ++k) d[k] = w0 + k < wg ? digits[(size_t)(w0 + k) * n + i] : MSM_DIGIT_NONE;
#pragma unroll
            for (int k = 0; k < ILP; ++k)
                if (d[k] != MSM_DIGIT_NONE) pos[k] = atomicAdd(&cursor[(size_t)(w0 + k) * bstride + (d[k] & 0x7fffffffu)], 1u);
#pragma unroll
            for (int k = 0; k < ILP; ++k)
                if (d[k] != MSM_DIGIT_NONE)
                    sorted[(size_t)(w0 + k) * sstride + pos[k]] = (uint32_t)(point0 + i + (size_t)(w0 + k + w_lo) * ioff) | (d[k] & 0x80000000u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Load balancing: order the (window,bucket) slots by descending population, so that the 32 lanes
// of a warp walk buckets of (nearly) equal length and the longest buckets start first.  Counting
// sort keyed by min(count, SIZE_KEYS-1); ~1 M slots, three tiny kernels.
// ------------------------------------------------------------------------------------------
static constexpr uint32_t SIZE_KEYS = 2048;

static __global__ void k_size_hist(const uint32_t* __restrict__ counts, size_t slots, uint32_t hkey, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[SIZE_KEYS];
    for (uint32_t i = threadIdx.x; i < SIZE_KEYS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < slots; t += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&h[min(counts[t], hkey)], 1u);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < SIZE_KEYS; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
}
// start[key] = number of slots with a larger key (descending order); one block of SIZE_KEYS/2 threads
static __global__ void k_size_scan(const uint32_t* __restrict__ hist, uint32_t* __restrict__ start) {
    __shared__ uint32_t v[SIZE_KEYS];
    for (uint32_t i = threadIdx.x; i < SIZE_KEYS; i += blockDim.x) v[i] = hist[SIZE_KEYS - 1 - i];   // reversed
    __syncthreads();
    for (uint32_t off = 1; off < SIZE_KEYS; off <<= 1) {
        uint32_t a[2];
        for (int r = 0; r < 2; ++r) { uint32_t i = threadIdx.x + r * blockDim.x; a[r] = i >= off ? v[i - off] : 0; }
        __syncthreads();
        for (int r = 0; r < 2; ++r) { uint32_t i = threadIdx.x + r * blockDim.x; v[i] += a[r]; }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < SIZE_KEYS; i += blockDim.x)
        start[SIZE_KEYS - 1 - i] = i ? v[i - 1] : 0;          // exclusive prefix in reversed (descending) order
}
// Block-aggregated: the populations of uniform scalars fall on a few dozen keys, so one global atomic per (warp, key)
// piles onto the same addresses (42 us for 2^19 slots).  A 1024-thread block ranks its slots in a shared histogram
// (warp-aggregated), reserves one contiguous range per key it saw, and writes.
static __global__ void __launch_bounds__(1024) k_size_scatter(const uint32_t* __restrict__ counts, size_t slots, uint32_t hkey, uint32_t* __restrict__ start,
                                      uint32_t* __restrict__ order) {
    __shared__ uint32_t sh[SIZE_KEYS];
    for (uint32_t i = threadIdx.x; i < SIZE_KEYS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = t < slots;
    uint32_t key = valid ? min(counts[t], hkey) : 0xffffffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, key);
    int leader = __ffs(peers) - 1;
    uint32_t lane = threadIdx.x & 31;
    uint32_t rank = __popc(peers & ((1u << lane) - 1));
    uint32_t base = 0;
    if (valid && (int)lane == leader) base = atomicAdd(&sh[key], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    rank += base;                                          // rank of this slot among the block's slots with its key
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < SIZE_KEYS; i += blockDim.x) {
        uint32_t c = sh[i];
        if (c) sh[i] = atomicAdd(&start[i], c);            // the block's range for key i
    }
    __syncthreads();
    if (valid) order[sh[key] + rank] = (uint32_t)t;
}

// ------------------------------------------------------------------------------------------
// bucket accumulation: one thread per (window, bucket)
// ------------------------------------------------------------------------------------------
// (64-thread blocks at ten per SM, for a finer-grained drain at the end of the launch, measured slower: 8.20 vs 7.90 ms)
template <class F>
__global__ void __launch_bounds__(128, sizeof(F) > 32 ? 3 : 5)     // G2: three blocks per SM (<= 168 registers); G1: five (<= 102)
k_accumulate(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ cursor_end, const uint32_t* __restrict__ counts,
             const uint32_t* __restrict__ order, size_t sstride, uint32_t nb, int W, int accumulate_into, uint32_t hkey, int pf,
             XYZZ<F>* __restrict__ buckets) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)W * nb) return;
    // slots in descending population order: neighbouring lanes get buckets of equal length
    size_t slot = order[t];
    uint32_t w = (uint32_t)(slot / nb);
    uint32_t end = cursor_end[slot], cnt = counts[slot];
    const uint32_t* idx = sorted + (size_t)w * sstride;
    if (accumulate_into && cnt == 0) return;                 // nothing new for this bucket in this chunk
    if (cnt >= hkey) {                                       // heavy bucket: left to k_accumulate_heavy
        if (!accumulate_into) store_vec(buckets + slot, XYZZ<F>::inf());
        return;
    }
    XYZZ<F> acc = accumulate_into ? load_vec_rw(buckets + slot) : XYZZ<F>::inf();
    if (pf && cnt) {
        // the next point's line is requested one addition (~13 us of this warp's time) ahead, so the gather finds it in cache
        // (prefetch.global.L2 and .L1 measure the same: accumulate 7.88 -> 7.83 ms at 2^22, G2 2^19 3.50 -> 3.47 ms)
        uint32_t e = idx[end - cnt];
        for (uint32_t k = end - cnt; k < end; ++k) {
            const uint32_t e_next = k + 1 < end ? idx[k + 1] : e;
            const void* nx = bases + (e_next & 0x7fffffffu);
            if (pf == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
            else asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
            Affine<F> p = load_vec(bases + (e & 0x7fffffffu));
            xyzz_madd(acc, p, (e >> 31) != 0);
            e = e_next;
        }
    } else {
        for (uint32_t k = end - cnt; k < end; ++k) {
            uint32_t e = idx[k];
            Affine<F> p = load_vec(bases + (e & 0x7fffffffu));
            xyzz_madd(acc, p, (e >> 31) != 0);
        }
    }
    store_vec(buckets + slot, acc);
}

// ------------------------------------------------------------------------------------------
// Batched-affine bucket accumulation.  An XYZZ mixed add costs 10 field products; an AFFINE add
// costs 3 plus an inversion, and Montgomery's trick turns k inversions into one plus 3(k-1)
// products -- 6 products per add once the one inversion is shared widely enough.  One thread still
// owns one bucket (slots ordered by population, so a block is homogeneous), but it sums its points
// as a pairwise tree instead of a chain:
//   round 1   the sorted entries are paired (2i, 2i+1) straight from the base table;
//             forward pass: denominators d_i and their running product (kept in local memory),
//             ONE inversion per BLOCK of the product of all threads' totals (prefix/suffix scans in
//             shared memory, binary-Euclid inverse by one thread while other blocks keep the SM busy),
//             backward pass: 1/d_i recovered with 2 products, the affine sum written to local memory;
//   round r   the same on the previous round's results (ping-pong between two local buffers);
// until at most BA_FINISH points are left, which join the bucket by the ordinary XYZZ chain (so
// the reduction kernels and the chunked / multi-GPU paths see the same bucket format).
// P = Q (doubling), P = -Q and infinities are handled per pair: the pair contributes d = 2y or
// d = 1 to the batch and takes the matching formula on the way back.
// ------------------------------------------------------------------------------------------
static constexpr int BA_THREADS = 256;
static constexpr int BA_MAXP = 96;            // pairs in round 1: buckets of up to 192 points
static constexpr uint32_t BA_FINISH = 32;     // points left to the XYZZ chain
static constexpr uint32_t BA_MIN = 40;        // blocks whose largest bucket is smaller keep the plain chain

template <class F>
struct BaPair { int kind; F d; };             // kind 0: generic add, 1: doubling, 2: result is p, 3: result is q, 4: result is infinity

template <class F>
__device__ __forceinline__ BaPair<F> ba_classify(const Affine<F>& p, const Affine<F>& q) {
    BaPair<F> r;
    r.d = F::one();
    if (p.is_inf()) { r.kind = 3; return r; }
    if (q.is_inf()) { r.kind = 2; return r; }
    F dx = f_sub(q.x, p.x);
    if (!dx.is_zero()) { r.kind = 0; r.d = dx; return r; }
    if (p.y == q.y && !p.y.is_zero()) { r.kind = 1; r.d = f_dbl(p.y); return r; }
    r.kind = 4;
    return r;
}
template <class F>
__device__ __forceinline__ Affine<F> ba_finish_pair(const Affine<F>& p, const Affine<F>& q, int kind, const F& dinv) {
    if (kind == 2) return p;
    if (kind == 3) return q;
    if (kind == 4) return Affine<F>::inf();
    F lam;
    if (kind == 0) lam = f_mul(f_sub(q.y, p.y), dinv);
    else { F xx = f_sqr(p.x); lam = f_mul(f_add(f_dbl(xx), xx), dinv); }
    Affine<F> r;
    r.x = f_sub(f_sub(f_sqr(lam), p.x), q.x);
    r.y = f_sub(f_mul(lam, f_sub(p.x, r.x)), p.y);
    return r;
}

// 1 / t for every thread of the block from ONE field inversion: inclusive prefix and suffix product
// scans in shared memory, inverse of the grand total by thread 0, 1/t = inv * prefix_excl * suffix_excl.
template <class F>
__device__ __forceinline__ F ba_block_inverse(const F& t, F* sh_pre, F* sh_suf, F* sh_inv) {
    const int tid = threadIdx.x;
    sh_pre[tid] = t;
    sh_suf[tid] = t;
    __syncthreads();
    for (int off = 1; off < BA_THREADS; off <<= 1) {
        F a, b;
        const bool ha = tid >= off, hb = tid + off < BA_THREADS;
        if (ha) a = sh_pre[tid - off];
        if (hb) b = sh_suf[tid + off];
        __syncthreads();
        if (ha) sh_pre[tid] = f_mul(sh_pre[tid], a);
        if (hb) sh_suf[tid] = f_mul(sh_suf[tid], b);
        __syncthreads();
    }
    if (tid == 0) *sh_inv = f_inv(sh_pre[BA_THREADS - 1]);
    __syncthreads();
    F r = *sh_inv;
    if (tid > 0) r = f_mul(r, sh_pre[tid - 1]);
    if (tid + 1 < BA_THREADS) r = f_mul(r, sh_suf[tid + 1]);
    __syncthreads();                       // the scratch is reused by the next round
    return r;
}

template <class F>
__global__ void __launch_bounds__(BA_THREADS, 3)
k_accumulate_ba(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
                const uint32_t* __restrict__ cursor_end, const uint32_t* __restrict__ counts,
                const uint32_t* __restrict__ order, size_t sstride, uint32_t nb, int W, int accumulate_into, uint32_t hkey,
                XYZZ<F>* __restrict__ buckets) {
    __shared__ __align__(16) unsigned char raw_pre[BA_THREADS * sizeof(F)];
    __shared__ __align__(16) unsigned char raw_suf[BA_THREADS * sizeof(F)];
    __shared__ __align__(16) unsigned char raw_inv[sizeof(F)];
    __shared__ uint32_t cnt0_sh;
    F* sh_pre = reinterpret_cast<F*>(raw_pre);
    F* sh_suf = reinterpret_cast<F*>(raw_suf);
    F* sh_inv = reinterpret_cast<F*>(raw_inv);

    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < (size_t)W * nb;
    size_t slot = 0;
    uint32_t end = 0, cnt = 0;
    const uint32_t* idx = sorted;
    if (valid) {
        slot = order[t];
        end = cursor_end[slot];
        cnt = counts[slot];
        idx = sorted + (size_t)(slot / nb) * sstride + (end - cnt);
    }
    if (threadIdx.x == 0) cnt0_sh = cnt;          // descending order: the block's largest bucket
    __syncthreads();
    const uint32_t cnt0 = cnt0_sh;
    if (!valid) cnt = 0;

    if (cnt0 > 2 * BA_MAXP || cnt0 < BA_MIN) {     // block-uniform: plain chain (heavy buckets are skipped as in k_accumulate)
        if (!valid || (accumulate_into && cnt == 0)) return;
        if (cnt >= hkey) {
            if (!accumulate_into) store_vec(buckets + slot, XYZZ<F>::inf());
            return;
        }
        XYZZ<F> acc = accumulate_into ? load_vec_rw(buckets + slot) : XYZZ<F>::inf();
        for (uint32_t k = 0; k < cnt; ++k) {
            uint32_t e = idx[k];
            xyzz_madd(acc, load_vec(bases + (e & 0x7fffffffu)), (e >> 31) != 0);
        }
        store_vec(buckets + slot, acc);
        return;
    }

    Affine<F> bufA[BA_MAXP];
    Affine<F> bufB[BA_MAXP / 2];
    F pre[BA_MAXP];
    auto load_signed = [&](uint32_t e) {
        Affine<F> p = load_vec(bases + (e & 0x7fffffffu));
        if ((e >> 31) && !p.is_inf()) p.y = f_neg(p.y);
        return p;
    };

    // ---- round 1: pairs straight from the base table ----
    uint32_t npts = cnt, n0 = cnt0;
    {
        const uint32_t pairs = npts >> 1;
        F run = F::one();
        // forward pass: the common case needs only the two x coordinates (one 32-byte sector per point);
        // equal or zero x sends the pair through the full classification
#pragma unroll 2
        for (uint32_t i = 0; i < pairs; ++i) {
            const uint32_t e0 = idx[2 * i], e1 = idx[2 * i + 1];
            F px = load_vec(&bases[e0 & 0x7fffffffu].x), qx = load_vec(&bases[e1 & 0x7fffffffu].x);
            F d = f_sub(qx, px);
            if (d.is_zero() || px.is_zero() || qx.is_zero()) d = ba_classify(load_signed(e0), load_signed(e1)).d;
            pre[i] = run;
            run = f_mul(run, d);
        }
        F inv = ba_block_inverse(run, sh_pre, sh_suf, sh_inv);
        for (uint32_t i = pairs; i-- > 0;) {
            Affine<F> p = load_signed(idx[2 * i]), q = load_signed(idx[2 * i + 1]);
            BaPair<F> c = ba_classify(p, q);
            F dinv = f_mul(inv, pre[i]);
            inv = f_mul(inv, c.d);
            bufA[i] = ba_finish_pair(p, q, c.kind, dinv);
        }
        if (npts & 1) bufA[pairs] = load_signed(idx[npts - 1]);
        npts = pairs + (npts & 1);
        n0 = (n0 >> 1) + (n0 & 1);
    }
    // ---- further rounds on the previous results (A -> B -> A ...) ----
    bool in_a = true;
    while (n0 > BA_FINISH) {                       // block-uniform
        const uint32_t pairs = npts >> 1;
        F run = F::one();
        for (uint32_t i = 0; i < pairs; ++i) {
            const Affine<F>& p = in_a ? bufA[2 * i] : bufB[2 * i];
            const Affine<F>& q = in_a ? bufA[2 * i + 1] : bufB[2 * i + 1];
            BaPair<F> c = ba_classify(p, q);
            pre[i] = run;
            run = f_mul(run, c.d);
        }
        F inv = ba_block_inverse(run, sh_pre, sh_suf, sh_inv);
        // ascending writes into the OTHER buffer; the inverse recovery runs descending, so recover
        // first into pre[] (1/d_i replaces the prefix product) and finish the pairs afterwards
        for (uint32_t i = pairs; i-- > 0;) {
            const Affine<F>& p = in_a ? bufA[2 * i] : bufB[2 * i];
            const Affine<F>& q = in_a ? bufA[2 * i + 1] : bufB[2 * i + 1];
            BaPair<F> c = ba_classify(p, q);
            F dinv = f_mul(inv, pre[i]);
            inv = f_mul(inv, c.d);
            Affine<F> r = ba_finish_pair(p, q, c.kind, dinv);
            if (in_a) bufB[i] = r; else bufA[i] = r;
        }
        if (npts & 1) {
            if (in_a) bufB[pairs] = bufA[npts - 1]; else bufA[pairs] = bufB[npts - 1];
        }
        npts = pairs + (npts & 1);
        n0 = (n0 >> 1) + (n0 & 1);
        in_a = !in_a;
    }
    // ---- the few remaining points join the bucket through the XYZZ chain ----
    if (!valid || (accumulate_into && cnt == 0)) return;
    XYZZ<F> acc = accumulate_into ? load_vec_rw(buckets + slot) : XYZZ<F>::inf();
    for (uint32_t i = 0; i < npts; ++i) xyzz_madd(acc, in_a ? bufA[i] : bufB[i], false);
    store_vec(buckets + slot, acc);
}


// ------------------------------------------------------------------------------------------
// G2 bucket accumulation on LANE PAIRS (round 2).  One thread per G2 bucket holds 4 x Fq2 of accumulator + an Fq2 point =
// 96 registers of data before any temporary: k_accumulate<Fq2> needs 168 registers, three blocks per SM, a 1.1 KB stack,
// and runs its Fq2 routines out of line.  Here two adjacent lanes share a bucket and lane r holds
// component c_r of every Fq2 value, so the per-lane state is that of the G1 kernel and everything is inlined Fq code:
//   (a0 + a1 u)(b0 + b1 u):  lane 0 forms a0 b0 - a1 b1, lane 1 forms a0 b1 + a1 b0, each as ONE two-term Montgomery
//   inner product (fp_dot<2>: 192 wide MADs per lane -- the 384 of a Karatsuba product, split in two) after fetching
//   the partner's halves with 16 shuffles;  squaring: lane 0 (a0+a1)(a0-a1), lane 1 2 a0 a1 (128 each);
//   a b - c d: one four-term inner product per lane (320).
// Same wide-MAD count as the single-thread formulas, half the registers per thread, twice the threads.
// Control flow is kept WARP-UNIFORM: the bucket loop runs to the longest bucket of the warp (slots are ordered by population, so
// the 16 buckets of a warp have nearly equal lengths) with finished pairs predicated off, and the special cases of the addition
// (identity operands, P = Q, P = -Q) are resolved by selects -- a pair that diverged around a shuffle would split the warp for
// the rest of the loop (measured: 3x slower).  The rare P = +-Q case takes a warp-uniform branch (__any_sync) in which every
// lane evaluates the doubling and only the affected pair keeps it.
struct PairLane {
    bool r;             // false: holds c0, true: holds c1
};
static constexpr uint32_t PAIR_FULL = 0xffffffffu;
__device__ __forceinline__ Fq pair_swap(const Fq& a) {
    Fq o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = __shfl_xor_sync(PAIR_FULL, a.v[i], 1);
    return o;
}
__device__ __forceinline__ Fq pair_sel(bool c, const Fq& a, const Fq& b) {
    Fq o;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = c ? a.v[i] : b.v[i];
    return o;
}
// both lanes: is the Fq2 value zero?  (the shuffle is executed unconditionally)
__device__ __forceinline__ bool pair_is_zero(const Fq& a) {
    const uint32_t own = a.is_zero() ? 1u : 0u;
    const uint32_t other = __shfl_xor_sync(PAIR_FULL, own, 1);
    return (own & other) != 0;
}
__device__ __forceinline__ Fq pair_mul(const PairLane& L, const Fq& ao, const Fq& bo) {
    const Fq ap = pair_swap(ao), bp = pair_swap(bo);
    Fq x[2] = {ao, pair_sel(L.r, ap, fp_neg(ap))};
    Fq y[2] = {pair_sel(L.r, bp, bo), pair_sel(L.r, bo, bp)};
    return fp_dot<FqParams, 2>(x, y);
}
__device__ __forceinline__ Fq pair_sqr(const PairLane& L, const Fq& ao) {
    const Fq ap = pair_swap(ao);
    Fq m = fp_mul(pair_sel(L.r, ao, fp_add(ao, ap)), pair_sel(L.r, ap, fp_sub(ao, ap)));
    return pair_sel(L.r, fp_dbl(m), m);
}
// a b - c d
__device__ __forceinline__ Fq pair_mulsub(const PairLane& L, const Fq& ao, const Fq& bo, const Fq& co, const Fq& d_o) {
    const Fq ap = pair_swap(ao), bp = pair_swap(bo), cp = pair_swap(co), dp = pair_swap(d_o);
    const Fq nco = fp_neg(co);
    // lane 0:  a0 b0 - a1 b1 - c0 d0 + c1 d1      lane 1 (own = index 1):  a0 b1 + a1 b0 - c0 d1 - c1 d0
    Fq x[4] = {pair_sel(L.r, ap, ao), pair_sel(L.r, ao, fp_neg(ap)), pair_sel(L.r, fp_neg(cp), nco), pair_sel(L.r, nco, cp)};
    Fq y[4] = {bo, bp, d_o, dp};
    return fp_dot<FqParams, 4>(x, y);
}
// one() of Fq2 as seen by this lane: (R, 0)
__device__ __forceinline__ Fq pair_one(const PairLane& L) { return pair_sel(L.r, Fq::zero(), Fq::one()); }

struct PairXYZZ { Fq x, y, zz, zzz; };       // this lane's halves
__device__ __forceinline__ PairXYZZ pair_sel(bool c, const PairXYZZ& a, const PairXYZZ& b) {
    PairXYZZ o;
    o.x = pair_sel(c, a.x, b.x); o.y = pair_sel(c, a.y, b.y); o.zz = pair_sel(c, a.zz, b.zz); o.zzz = pair_sel(c, a.zzz, b.zzz);
    return o;
}

// 2 * (affine p)     (mdbl-2008-s-1, a = 0), pair form of xyzz_dbl_affine; out of line: rare
static __device__ __noinline__ PairXYZZ pair_dbl_affine(bool r, const Fq& px, const Fq& py) {
    PairLane L; L.r = r;
    PairXYZZ o;
    Fq u = fp_dbl(py);
    Fq v = pair_sqr(L, u);
    Fq w = pair_mul(L, u, v);
    Fq s_ = pair_mul(L, px, v);
    Fq xx = pair_sqr(L, px);
    Fq m = fp_add(fp_dbl(xx), xx);
    Fq x3 = fp_sub(fp_sub(pair_sqr(L, m), s_), s_);
    o.y = pair_mulsub(L, m, fp_sub(s_, x3), w, py);
    o.x = x3;
    o.zz = v;
    o.zzz = w;
    return o;
}

// acc += (neg ? -p : p) when `act`      (madd-2008-s), pair form of xyzz_madd; every lane of the warp executes it
__device__ __forceinline__ void pair_madd(const PairLane& L, PairXYZZ& acc, const Fq& px, const Fq& py_in, bool neg, bool act) {
    const bool p_inf = pair_is_zero(px) & pair_is_zero(py_in);
    const Fq py = pair_sel(neg, fp_neg(py_in), py_in);
    const bool acc_inf = pair_is_zero(acc.zz);
    Fq u2 = pair_mul(L, px, acc.zz);
    Fq s2 = pair_mul(L, py, acc.zzz);
    Fq pp_ = fp_sub(u2, acc.x);
    Fq rr = fp_sub(s2, acc.y);
    const bool generic = act && !p_inf && !acc_inf;
    const bool x_eq = pair_is_zero(pp_);                          // (no short-circuit: every lane executes the shuffles)
    const bool same_y = pair_is_zero(rr);
    const bool same_x = generic && x_eq;
    PairXYZZ res;
    Fq pp = pair_sqr(L, pp_);
    Fq ppp = pair_mul(L, pp_, pp);
    Fq q = pair_mul(L, acc.x, pp);
    res.x = fp_sub(fp_sub(fp_sub(pair_sqr(L, rr), ppp), q), q);
    res.y = pair_mulsub(L, rr, fp_sub(q, res.x), acc.y, ppp);
    res.zz = pair_mul(L, acc.zz, pp);
    res.zzz = pair_mul(L, acc.zzz, ppp);
    if (__any_sync(PAIR_FULL, same_x)) {                         // P == Q or P == -Q somewhere in the warp (rare)
        PairXYZZ d = pair_dbl_affine(L.r, px, py);
        PairXYZZ inf;
        inf.x = Fq::zero(); inf.y = Fq::zero(); inf.zz = Fq::zero(); inf.zzz = Fq::zero();
        res = pair_sel(same_x, pair_sel(same_y, d, inf), res);
    }
    PairXYZZ first;                                              // identity + p
    first.x = px; first.y = py; first.zz = pair_one(L); first.zzz = first.zz;
    acc = pair_sel(act && !p_inf, pair_sel(acc_inf, first, res), acc);
}

template <int BLOCKS>
__global__ void __launch_bounds__(128, BLOCKS)
k_accumulate_g2pair(const Affine<Fq2>* __restrict__ bases, const uint32_t* __restrict__ sorted,
                    const uint32_t* __restrict__ cursor_end, const uint32_t* __restrict__ counts,
                    const uint32_t* __restrict__ order, size_t sstride, uint32_t nb, int W, int accumulate_into, uint32_t hkey,
                    XYZZ<Fq2>* __restrict__ buckets) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t pair = t >> 1;
    const bool valid = pair < (size_t)W * nb;                // no early exit: every lane takes part in the warp's shuffles
    PairLane L;
    L.r = (t & 1) != 0;
    size_t slot = 0;
    uint32_t end = 0, cnt = 0;
    if (valid) { slot = order[pair]; end = cursor_end[slot]; cnt = counts[slot]; }
    const uint32_t* idx = sorted + (size_t)(slot / nb) * sstride + (end - cnt);
    const bool heavy = cnt >= hkey;                          // left to k_accumulate_heavy
    const bool work = valid && !heavy && !(accumulate_into && cnt == 0);
    const uint32_t my_cnt = work ? cnt : 0u;
    const uint32_t max_cnt = __reduce_max_sync(PAIR_FULL, my_cnt);
    Fq* bk = reinterpret_cast<Fq*>(buckets + slot) + (L.r ? 1 : 0);      // x.c_r; y, zz, zzz follow at strides of two Fq
    PairXYZZ acc;
    acc.x = Fq::zero(); acc.y = Fq::zero(); acc.zz = Fq::zero(); acc.zzz = Fq::zero();
    if (work && accumulate_into) { acc.x = load_vec_rw(bk); acc.y = load_vec_rw(bk + 2); acc.zz = load_vec_rw(bk + 4); acc.zzz = load_vec_rw(bk + 6); }
    for (uint32_t k = 0; k < max_cnt; ++k) {
        const bool act = k < my_cnt;
        const uint32_t e = act ? idx[k] : 0u;                // finished pairs re-read base 0 and discard the result
        const Fq* pb = reinterpret_cast<const Fq*>(bases + (e & 0x7fffffffu)) + (L.r ? 1 : 0);
        const Fq px = load_vec(pb), py = load_vec(pb + 2);
        pair_madd(L, acc, px, py, (e >> 31) != 0, act);
    }
    if (work || (valid && heavy && !accumulate_into)) {      // heavy buckets start from the identity
        store_vec(bk, acc.x); store_vec(bk + 2, acc.y); store_vec(bk + 4, acc.zz); store_vec(bk + 6, acc.zzz);
    }
}

// Skewed scalar distributions (many equal scalars, boolean witnesses, structured inputs) put thousands
// of points into one bucket; one thread per bucket would serialise them.  Buckets with at least
// SIZE_KEYS-1 points are the first hist[SIZE_KEYS-1] entries of `order`: blocks (x = heavy bucket,
// y = part) stride over a slice of the bucket, fold their 128 partial sums in shared memory and add
// the block sum into the bucket under a per-bucket spin lock (the holder never waits on anything,
// so the lock cannot deadlock).  Uniform scalars never reach the threshold (the window rule keeps
// the mean population near 2^7), so this launch is normally a few thousand empty blocks.
static constexpr uint32_t HEAVY_LOCKS = 1024;      // lock words, hashed by slot
static constexpr uint32_t HEAVY_PARTS = 32;        // gridDim.y
static constexpr uint32_t HEAVY_PART_MIN = 4096;   // points per part before a bucket is split further

template <class F>
__global__ void __launch_bounds__(128)
k_accumulate_heavy(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
                   const uint32_t* __restrict__ cursor_end, const uint32_t* __restrict__ counts,
                   const uint32_t* __restrict__ order, const uint32_t* __restrict__ size_hist,
                   uint32_t* __restrict__ locks, size_t sstride, uint32_t nb, uint32_t hkey, XYZZ<F>* buckets) {
    __shared__ __align__(16) unsigned char raw[128 * sizeof(XYZZ<F>)];
    XYZZ<F>* sh = reinterpret_cast<XYZZ<F>*>(raw);
    const uint32_t n_heavy = size_hist[hkey];
    const uint32_t tid = threadIdx.x, part = blockIdx.y;
    for (uint32_t h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        size_t slot = order[h];
        uint32_t w = (uint32_t)(slot / nb);
        uint32_t end = cursor_end[slot], cnt = counts[slot];
        uint32_t parts = (cnt + HEAVY_PART_MIN - 1) / HEAVY_PART_MIN;
        if (parts > HEAVY_PARTS) parts = HEAVY_PARTS;
        if (part >= parts) continue;                                   // uniform over the block
        const uint32_t lo = end - cnt + (uint32_t)(((uint64_t)cnt * part) / parts);
        const uint32_t hi = end - cnt + (uint32_t)(((uint64_t)cnt * (part + 1)) / parts);
        const uint32_t* idx = sorted + (size_t)w * sstride;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t k = lo + tid; k < hi; k += 128) {
            uint32_t e = idx[k];
            Affine<F> p = load_vec(bases + (e & 0x7fffffffu));
            xyzz_madd(acc, p, (e >> 31) != 0);
        }
        sh[tid] = acc;
        __syncthreads();
        for (uint32_t s = 64; s > 0; s >>= 1) {
            if (tid < s) { XYZZ<F> a = sh[tid]; xyzz_add(a, sh[tid + s]); sh[tid] = a; }
            __syncthreads();
        }
        if (tid == 0) {
            uint32_t* lock = locks + (slot & (HEAVY_LOCKS - 1));
            while (atomicCAS(lock, 0u, 1u) != 0u) __nanosleep(200);
            __threadfence();
            XYZZ<F> b;
            {
                const uint4* src = reinterpret_cast<const uint4*>(buckets + slot);
                uint4* dst = reinterpret_cast<uint4*>(&b);
#pragma unroll
                for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 16); ++i) dst[i] = __ldcg(src + i);     // L2: other SMs update it
            }
            xyzz_add(b, sh[0]);
            {
                const uint4* src = reinterpret_cast<const uint4*>(&b);
                uint4* dst = reinterpret_cast<uint4*>(buckets + slot);
#pragma unroll
                for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 16); ++i) __stcg(dst + i, src[i]);
            }
            __threadfence();
            atomicExch(lock, 0u);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Bucket reduction.  Per window we need  S = sum_b (b+1) B_b = sum_b b*B_b + sum_b B_b.
// Invariant at level k (M_k = L^k):   sum_b b*B_b = sum_idx [ Cs_k[idx] + M_k * idx * R_k[idx] ]
// with R_0 = B, Cs_0 = 0.  One level folds L consecutive entries (segment u):
//   R_{k+1}[u]  = sum_i R_k[uL+i]
//   Cs_{k+1}[u] = sum_i Cs_k[uL+i] + M_k * sum_i i*R_k[uL+i]        (running-sum trick for the i* term)
// When one entry is left, S = Cs + R.  Depth is ~3L adds per level, fully parallel inside a level.
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_reduce_lvl(const XYZZ<F>* __restrict__ R_in, const XYZZ<F>* __restrict__ Cs_in, uint32_t n_in, uint32_t L,
             int log2_M, int W, XYZZ<F>* __restrict__ R_out, XYZZ<F>* __restrict__ Cs_out, uint32_t n_out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)W * n_out) return;
    uint32_t w = (uint32_t)(t / n_out), u = (uint32_t)(t % n_out);
    uint32_t lo = u * L, hi = min(lo + L, n_in);
    const XYZZ<F>* Rw = R_in + (size_t)w * n_in;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf(), cs = XYZZ<F>::inf();
    for (uint32_t k = hi; k-- > lo;) {
        XYZZ<F> e = load_vec_rw(Rw + k);
        if (k != lo) {                 // weight (k - lo) >= 1
            xyzz_add(run, e);
            xyzz_add(acc, run);
        } else {
            xyzz_add(run, e);          // weight 0: only joins the plain sum
        }
        if (Cs_in) {
            XYZZ<F> cc = load_vec_rw(Cs_in + (size_t)w * n_in + k);
            xyzz_add(cs, cc);
        }
    }
    xyzz_dbl_k(acc, log2_M);
    xyzz_add(cs, acc);
    store_vec(R_out + t, run);
    store_vec(Cs_out + t, cs);
}

// ------------------------------------------------------------------------------------------
// Log-depth continuation of the bucket reduction.  After level 0 every window holds n1 = nb/8 pairs
// (R_u, A_u) and needs   sum_u A_u + 8 * sum_u u * R_u   (+ sum_u R_u).  Writing u in binary,
//   sum_u u R_u = sum_j 2^j m_j,   m_j = sum of the R_u whose index has bit j set,
// and the masked sums obey a butterfly: merging two sibling nodes of 2^(b-1) entries into one of 2^b,
//   total = total_lo + total_hi,  sumA = sumA_lo + sumA_hi,  m_j = m_j_lo + m_j_hi (j < b-1),  m_(b-1) = total_hi.
// All (b+1) additions of all merges of a level are independent: k_merge_lvl runs them one per thread,
// so a level costs ONE group-addition latency instead of the 24 of a running-sum level, and the total
// work stays ~2 additions per entry.  k_bits_final then scales m_j by 2^(j+3) in parallel lanes and
// tree-sums them.  (7 levels x ~0.2 ms -> ~0.35 ms at 2^19 buckets.)
// Node layout at level b >= 1: (b+2) consecutive XYZZ values [total, sumA, m_0 .. m_(b-1)].
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_merge_lvl(const XYZZ<F>* __restrict__ in, const XYZZ<F>* __restrict__ in_R0, const XYZZ<F>* __restrict__ in_A0,
            XYZZ<F>* __restrict__ out, int b, uint32_t n_prev, int Wb) {
    const uint32_t n_new = n_prev >> 1, slots = (uint32_t)b + 2;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)Wb * n_new * slots) return;
    uint32_t v = (uint32_t)(t % slots);
    size_t node = t / slots;                               // w * n_new + q
    uint32_t w = (uint32_t)(node / n_new), q = (uint32_t)(node % n_new);
    size_t lo = (size_t)w * n_prev + 2 * (size_t)q, hi = lo + 1;
    XYZZ<F> r;
    if (b == 1) {                                          // children are level-0 entries (R_u, A_u)
        if (v == 0) { r = load_vec_rw(in_R0 + lo); xyzz_add(r, load_vec_rw(in_R0 + hi)); }
        else if (v == 1) {
            if (in_A0) { r = load_vec_rw(in_A0 + lo); xyzz_add(r, load_vec_rw(in_A0 + hi)); }
            else r = XYZZ<F>::inf();                       // level 0 skipped: the entries are the buckets themselves
        }
        else r = load_vec_rw(in_R0 + hi);                  // m_0 = total of the odd child
    } else {
        const uint32_t ps = (uint32_t)b + 1;               // slots per child node
        if (v == slots - 1) r = load_vec_rw(in + hi * ps); // m_(b-1) = total_hi
        else { r = load_vec_rw(in + lo * ps + v); xyzz_add(r, load_vec_rw(in + hi * ps + v)); }
    }
    store_vec(out + t, r);
}

// One block of 32 threads per window: S_w = sumA + 8 * sum_j 2^j m_j + total.  Lane j owns m_j.
template <class F>
__global__ void __launch_bounds__(32)
k_bits_final(const XYZZ<F>* __restrict__ nodes, const XYZZ<F>* __restrict__ R0, const XYZZ<F>* __restrict__ A0, int B, int lshift,
             XYZZ<F>* __restrict__ S_out, XYZZ<F>* __restrict__ Z_out) {
    __shared__ XYZZ<F> sh[32];
    const uint32_t w = blockIdx.x, lane = threadIdx.x;
    XYZZ<F> x = XYZZ<F>::inf();
    if (B == 0) {                                          // a single level-0 entry: u = 0 has weight 0
        if (lane == 0) { x = load_vec_rw(R0 + w); if (A0) xyzz_add(x, load_vec_rw(A0 + w)); }
    } else {
        const XYZZ<F>* nd = nodes + (size_t)w * (B + 2);
        if ((int)lane < B) {
            x = load_vec_rw(nd + 2 + lane);
            xyzz_dbl_k(x, (int)lane + lshift);                             // L * 2^j
        } else if ((int)lane == B) {
            x = load_vec_rw(nd);                           // total
            xyzz_add(x, load_vec_rw(nd + 1));              // + sumA
        }
    }
    sh[lane] = x;
    __syncthreads();
    for (uint32_t off = 16; off > 0; off >>= 1) {
        if (lane < off) { XYZZ<F> a = sh[lane]; xyzz_add(a, sh[lane + off]); sh[lane] = a; }
        __syncthreads();
    }
    if (lane == 0) { store_vec(S_out + w, sh[0]); store_vec(Z_out + w, XYZZ<F>::inf()); }
}

// ------------------------------------------------------------------------------------------
// The butterfly levels are a LATENCY chain: level b has only n1/2^b * (b+1) additions, and one thread needs ~11 us for
// an addition (14 dependent products on a lone warp, ~6 cycles per instruction).  The independent products of an
// addition can overlap across the four sub-partitions of an SM (the trick of k_final_coop), and a warp instruction costs
// the same for one lane as for 32 -- so a block of FOUR WARPS evaluates 32 additions at once: lane l of warp w computes
// product slot w of addition l, the operands and the eight temporaries live in shared memory (word-major, one value per
// lane: conflict-free), one barrier per formula level (4 for add-2008-s, 3 for dbl-2008-s-1).  ~3.5 us per addition level
// instead of ~11, same total work.  The exceptional cases (an identity operand, equal or opposite points) are detected
// per lane by warp 0 at level 2, evaluated there with the plain routine while the operands are still intact, and
// written over the lane's result after the last level.
// ------------------------------------------------------------------------------------------
template <class F> struct ShF { uint32_t w[sizeof(F) / 4][32]; };
template <class F> struct ShPt { ShF<F> x, y, zz, zzz; };

__device__ __forceinline__ Fq sh_ld(const ShF<Fq>& s, int lane) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = s.w[k][lane];
    return r;
}
__device__ __forceinline__ void sh_st(ShF<Fq>& s, int lane, const Fq& a) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s.w[k][lane] = a.v[k];
}
__device__ __forceinline__ Fq2 sh_ld(const ShF<Fq2>& s, int lane) {
    Fq2 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) { r.c0.v[k] = s.w[k][lane]; r.c1.v[k] = s.w[8 + k][lane]; }
    return r;
}
__device__ __forceinline__ void sh_st(ShF<Fq2>& s, int lane, const Fq2& a) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { s.w[k][lane] = a.c0.v[k]; s.w[8 + k][lane] = a.c1.v[k]; }
}
template <class F>
__device__ __forceinline__ XYZZ<F> sh_ld_pt(const ShPt<F>& p, int lane) {
    XYZZ<F> r;
    r.x = sh_ld(p.x, lane); r.y = sh_ld(p.y, lane); r.zz = sh_ld(p.zz, lane); r.zzz = sh_ld(p.zzz, lane);
    return r;
}
template <class F>
__device__ __forceinline__ void sh_st_pt(ShPt<F>& p, int lane, const XYZZ<F>& a) {
    sh_st(p.x, lane, a.x); sh_st(p.y, lane, a.y); sh_st(p.zz, lane, a.zz); sh_st(p.zzz, lane, a.zzz);
}
// component `comp` (0..3 = x, y, zz, zzz) of a point
template <class F>
__device__ __forceinline__ ShF<F>& sh_comp(ShPt<F>& p, int comp) { return comp == 0 ? p.x : comp == 1 ? p.y : comp == 2 ? p.zz : p.zzz; }

// A[lane] += B[lane] for the lanes with `active` (the same value in all four warps); every thread of the 128-thread
// block calls it.  Inactive lanes touch nothing.
template <class F>
__device__ __forceinline__ void coop32_add(ShPt<F>& A, const ShPt<F>& B, ShF<F>* t, int warp, int lane, bool active) {
    F pp_ = F::zero(), rr_ = F::zero();
    XYZZ<F> fix = XYZZ<F>::inf();
    bool special = false;
    if (active) {
        if (warp == 0) sh_st(t[0], lane, f_mul(sh_ld(A.x, lane), sh_ld(B.zz, lane)));      // U1
        if (warp == 1) sh_st(t[1], lane, f_mul(sh_ld(B.x, lane), sh_ld(A.zz, lane)));      // U2
        if (warp == 2) sh_st(t[2], lane, f_mul(sh_ld(A.y, lane), sh_ld(B.zzz, lane)));     // S1
        if (warp == 3) sh_st(t[3], lane, f_mul(sh_ld(B.y, lane), sh_ld(A.zzz, lane)));     // S2
    }
    __syncthreads();
    if (active) {
        if (warp == 0) {
            pp_ = f_sub(sh_ld(t[1], lane), sh_ld(t[0], lane));
            special = sh_ld(A.zz, lane).is_zero() || sh_ld(B.zz, lane).is_zero() || pp_.is_zero();
            if (special) { fix = sh_ld_pt(A, lane); xyzz_add(fix, sh_ld_pt(B, lane)); }
            sh_st(t[4], lane, f_sqr(pp_));                                                  // PP
        }
        if (warp == 1) { rr_ = f_sub(sh_ld(t[3], lane), sh_ld(t[2], lane)); sh_st(t[5], lane, f_sqr(rr_)); }   // RR
        if (warp == 2) sh_st(t[6], lane, f_mul(sh_ld(A.zz, lane), sh_ld(B.zz, lane)));
        if (warp == 3) sh_st(t[7], lane, f_mul(sh_ld(A.zzz, lane), sh_ld(B.zzz, lane)));
    }
    __syncthreads();
    if (active) {
        if (warp == 0) sh_st(t[1], lane, f_mul(pp_, sh_ld(t[4], lane)));                    // PPP (U2 is dead)
        if (warp == 2) sh_st(t[0], lane, f_mul(sh_ld(t[0], lane), sh_ld(t[4], lane)));      // Q = U1 PP (only this warp reads U1 here)
        if (warp == 3) sh_st(A.zz, lane, f_mul(sh_ld(t[6], lane), sh_ld(t[4], lane)));      // ZZ3
    }
    __syncthreads();
    if (active) {
        if (warp == 1) {
            F ppp = sh_ld(t[1], lane), q = sh_ld(t[0], lane);
            F x3 = f_sub(f_sub(f_sub(sh_ld(t[5], lane), ppp), q), q);
            sh_st(A.y, lane, f_mulsub(rr_, f_sub(q, x3), sh_ld(t[2], lane), ppp));
            sh_st(A.x, lane, x3);
        }
        if (warp == 3) sh_st(A.zzz, lane, f_mul(sh_ld(t[7], lane), sh_ld(t[1], lane)));     // ZZZ3
    }
    __syncthreads();
    if (active && warp == 0 && special) sh_st_pt(A, lane, fix);
    __syncthreads();
}

// A[lane] = 2 A[lane] for the active lanes   (dbl-2008-s-1, a = 0; three levels)
template <class F>
__device__ __forceinline__ void coop32_dbl(ShPt<F>& A, ShF<F>* t, int warp, int lane, bool active) {
    active = active && !sh_ld(A.zz, lane).is_zero();                                        // 2 * identity: nothing to do
    if (active) {
        if (warp == 0) { F u = f_dbl(sh_ld(A.y, lane)); sh_st(t[0], lane, u); sh_st(t[1], lane, f_sqr(u)); }     // U, V
        if (warp == 1) { F xx = f_sqr(sh_ld(A.x, lane)); sh_st(t[2], lane, f_add(f_dbl(xx), xx)); }              // M
    }
    __syncthreads();
    if (active) {
        if (warp == 0) sh_st(t[3], lane, f_mul(sh_ld(t[0], lane), sh_ld(t[1], lane)));      // W
        if (warp == 1) sh_st(t[4], lane, f_mul(sh_ld(A.x, lane), sh_ld(t[1], lane)));       // S
        if (warp == 2) sh_st(t[5], lane, f_sqr(sh_ld(t[2], lane)));                         // M^2
        if (warp == 3) sh_st(A.zz, lane, f_mul(sh_ld(t[1], lane), sh_ld(A.zz, lane)));      // ZZ3
    }
    __syncthreads();
    if (active) {
        if (warp == 0) {
            F s_ = sh_ld(t[4], lane);
            F x3 = f_sub(f_sub(sh_ld(t[5], lane), s_), s_);
            sh_st(A.y, lane, f_mulsub(sh_ld(t[2], lane), f_sub(s_, x3), sh_ld(t[3], lane), sh_ld(A.y, lane)));
            sh_st(A.x, lane, x3);
        }
        if (warp == 1) sh_st(A.zzz, lane, f_mul(sh_ld(t[3], lane), sh_ld(A.zzz, lane)));    // ZZZ3
    }
    __syncthreads();
}

// k_merge_lvl on four-warp blocks: 32 node entries per block.  Warp w moves component w of the operands and results.
template <class F>
__global__ void __launch_bounds__(128)
k_merge_lvl_coop(const XYZZ<F>* __restrict__ in, const XYZZ<F>* __restrict__ in_R0, const XYZZ<F>* __restrict__ in_A0,
                 XYZZ<F>* __restrict__ out, int b, uint32_t n_prev, int Wb) {
    __shared__ ShPt<F> A, Bp;
    __shared__ ShF<F> t[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n_new = n_prev >> 1, slots = (uint32_t)b + 2;
    const size_t e = (size_t)blockIdx.x * 32 + lane;
    const bool valid = e < (size_t)Wb * n_new * slots;
    const XYZZ<F>* pa = nullptr;
    const XYZZ<F>* pb = nullptr;
    if (valid) {
        uint32_t v = (uint32_t)(e % slots);
        size_t node = e / slots;
        uint32_t w = (uint32_t)(node / n_new), q = (uint32_t)(node % n_new);
        size_t lo = (size_t)w * n_prev + 2 * (size_t)q, hi = lo + 1;
        if (b == 1) {
            if (v == 0) { pa = in_R0 + lo; pb = in_R0 + hi; }
            else if (v == 1) { if (in_A0) { pa = in_A0 + lo; pb = in_A0 + hi; } }      // no level 0: sumA starts as the identity
            else pa = in_R0 + hi;
        } else {
            const uint32_t ps = (uint32_t)b + 1;
            if (v == slots - 1) pa = in + hi * ps;
            else { pa = in + lo * ps + v; pb = in + hi * ps + v; }
        }
    }
    sh_st(sh_comp(A, warp), lane, pa ? load_vec_rw(reinterpret_cast<const F*>(pa) + warp) : F::zero());
    if (pb) sh_st(sh_comp(Bp, warp), lane, load_vec_rw(reinterpret_cast<const F*>(pb) + warp));
    __syncthreads();
    coop32_add(A, Bp, t, warp, lane, pb != nullptr);
    if (valid) store_vec(reinterpret_cast<F*>(out + e) + warp, sh_ld(sh_comp(A, warp), lane));
}

// k_bits_final on a four-warp block per window: lane j scales m_j by 2^(j+lshift), L = 2^lshift the level-0 segment (the lanes' doubling chains run together,
// masked by their lengths), lane B holds total + sumA, then a five-level tree of 32-wide additions.
template <class F>
__global__ void __launch_bounds__(128)
k_bits_final_coop(const XYZZ<F>* __restrict__ nodes, const XYZZ<F>* __restrict__ R0, const XYZZ<F>* __restrict__ A0, int B, int lshift,
                  XYZZ<F>* __restrict__ S_out, XYZZ<F>* __restrict__ Z_out) {
    __shared__ ShPt<F> A, Bp;
    __shared__ ShF<F> t[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x;
    // A[lane]: m_lane (lane < B), total (lane == B; sumA is added below), identity otherwise
    {
        F a = F::zero(), bb = F::zero();
        if (B == 0) {
            if (lane == 0) { a = load_vec_rw(reinterpret_cast<const F*>(R0 + w) + warp); if (A0) bb = load_vec_rw(reinterpret_cast<const F*>(A0 + w) + warp); }
        } else {
            const XYZZ<F>* nd = nodes + (size_t)w * (B + 2);
            if (lane < B) a = load_vec_rw(reinterpret_cast<const F*>(nd + 2 + lane) + warp);
            else if (lane == B) { a = load_vec_rw(reinterpret_cast<const F*>(nd) + warp); bb = load_vec_rw(reinterpret_cast<const F*>(nd + 1) + warp); }
        }
        sh_st(sh_comp(A, warp), lane, a);
        sh_st(sh_comp(Bp, warp), lane, bb);
    }
    __syncthreads();
    coop32_add(A, Bp, t, warp, lane, B == 0 ? lane == 0 : lane == B);
    for (int r = 0; r < B - 1 + lshift; ++r) coop32_dbl(A, t, warp, lane, lane < B && r < lane + lshift);
    for (int off = 16; off > 0; off >>= 1) {
        // B[lane] = A[lane + off]: every warp copies its component, then the lower half adds
        if (lane < off) sh_st(sh_comp(Bp, warp), lane, sh_ld(sh_comp(A, warp), lane + off));
        __syncthreads();
        coop32_add(A, Bp, t, warp, lane, lane < off);
    }
    if (lane == 0) {
        store_vec(reinterpret_cast<F*>(S_out + w) + warp, sh_ld(sh_comp(A, warp), 0));
        store_vec(reinterpret_cast<F*>(Z_out + w) + warp, F::zero());
    }
}

// Horner over the window sums + normalisation.  mode 0: Jacobian image (x, y, 1) / (1, 1, 0);
// mode 1: raw XYZZ partial (for multi-GPU combination).
template <class F>
__global__ void k_final(const XYZZ<F>* __restrict__ R, const XYZZ<F>* __restrict__ Cs, int c, int W, int mode,
                        F* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = W - 1; w >= 0; --w) {
        xyzz_dbl_k(acc, c);
        XYZZ<F> s = load_vec_rw(R + w);
        XYZZ<F> cs = load_vec_rw(Cs + w);
        xyzz_add(s, cs);
        xyzz_add(acc, s);
    }
    if (mode == 1) {
        out[0] = acc.x; out[1] = acc.y; out[2] = acc.zz; out[3] = acc.zzz;
        return;
    }
    if (acc.is_inf()) { out[0] = F::one(); out[1] = F::one(); out[2] = F::zero(); return; }
    Affine<F> a = xyzz_to_affine(acc);
    out[0] = a.x; out[1] = a.y; out[2] = F::one();
}

// identity result for n == 0
// ------------------------------------------------------------------------------------------
// The Horner over windows is a chain of c*(W-1) DEPENDENT doublings: one thread runs it at the speed of one
// SM sub-partition's multiplier (~3.5 us per doubling).  A formula's independent products cannot overlap inside a
// warp -- every warp instruction of a sub-partition goes through the same multiply pipe, four cycles each -- but
// they can across the FOUR sub-partitions of an SM: this kernel runs one block of four warps, lane 0 of each
// evaluating one product of a formula level, shared memory + __syncthreads between levels.
//   Jacobian doubling (dbl-2009-l, a = 0): 3 levels   {A = X^2, B = Y^2, YZ} {F = (3A)^2, C = B^2, t = (X+B)^2} {Y3}
//   XYZZ addition (add-2008-s):            4 levels   {U1, U2, S1, S2} {PP, RR, ZZ1 ZZ2, ZZZ1 ZZZ2} {PPP, Q, ZZ3} {Y3, ZZZ3}
// Same values as k_final (a group element has one normal form); the exceptional cases of the addition (equal or
// opposite points, infinity) are block-uniform and fall back to the plain routine on one thread.
// ------------------------------------------------------------------------------------------
// The two formulas on four warps.  `acc`, `b` and the scratch `t` (>= 8 values) live in shared memory; every thread
// of the 128-thread block calls them (the barriers are inside), only lane 0 of each warp computes.
template <class F>
__device__ __forceinline__ void coop_dbl(XYZZ<F>& acc, F* t, int warp, bool lead) {      // dbl-2008-s-1 (a = 0), 3 levels
    if (acc.is_inf()) return;                                                            // block-uniform
    if (lead && warp == 0) { F u = f_dbl(acc.y); t[0] = u; t[1] = f_sqr(u); }           // U, V
    if (lead && warp == 1) { F xx = f_sqr(acc.x); t[2] = f_add(f_dbl(xx), xx); }         // M
    __syncthreads();
    if (lead && warp == 0) t[3] = f_mul(t[0], t[1]);                                     // W
    if (lead && warp == 1) t[4] = f_mul(acc.x, t[1]);                                    // S
    if (lead && warp == 2) t[5] = f_sqr(t[2]);                                           // M^2
    if (lead && warp == 3) acc.zz = f_mul(t[1], acc.zz);                                 // ZZ3 (no other reader in this level)
    __syncthreads();
    if (lead && warp == 0) {
        F x3 = f_sub(f_sub(t[5], t[4]), t[4]);
        acc.y = f_mulsub(t[2], f_sub(t[4], x3), t[3], acc.y);
        acc.x = x3;
    }
    if (lead && warp == 1) acc.zzz = f_mul(t[3], acc.zzz);                               // ZZZ3
    __syncthreads();
}
template <class F>
__device__ __forceinline__ void coop_add(XYZZ<F>& acc, const XYZZ<F>& b, F* t, int warp, bool lead, int tid) {   // add-2008-s, 4 levels
    if (b.is_inf()) return;                                                              // block-uniform
    if (acc.is_inf()) { __syncthreads(); if (tid == 0) acc = b; __syncthreads(); return; }
    if (lead && warp == 0) t[0] = f_mul(acc.x, b.zz);                                    // U1
    if (lead && warp == 1) t[1] = f_mul(b.x, acc.zz);                                    // U2
    if (lead && warp == 2) t[2] = f_mul(acc.y, b.zzz);                                   // S1
    if (lead && warp == 3) t[3] = f_mul(b.y, acc.zzz);                                   // S2
    __syncthreads();
    if (t[0] == t[1]) {                                                                  // same x: doubling or cancellation (rare)
        __syncthreads();
        if (tid == 0) { XYZZ<F> a_ = acc; xyzz_add(a_, b); acc = a_; }
        __syncthreads();
        return;
    }
    F pp_, rr_;
    if (lead && warp == 0) { pp_ = f_sub(t[1], t[0]); t[4] = f_sqr(pp_); }               // PP
    if (lead && warp == 1) { rr_ = f_sub(t[3], t[2]); t[5] = f_sqr(rr_); }               // RR
    if (lead && warp == 2) t[6] = f_mul(acc.zz, b.zz);
    if (lead && warp == 3) t[7] = f_mul(acc.zzz, b.zzz);
    __syncthreads();
    if (lead && warp == 0) t[1] = f_mul(pp_, t[4]);                                      // PPP (U2 is dead)
    if (lead && warp == 2) t[0] = f_mul(t[0], t[4]);                                     // Q = U1 PP (only this thread reads U1 here)
    if (lead && warp == 3) acc.zz = f_mul(t[6], t[4]);                                   // ZZ3
    __syncthreads();
    if (lead && warp == 1) {                                                             // X3, Y3
        F x3 = f_sub(f_sub(f_sub(t[5], t[1]), t[0]), t[0]);
        acc.y = f_mulsub(rr_, f_sub(t[0], x3), t[2], t[1]);
        acc.x = x3;
    }
    if (lead && warp == 3) acc.zzz = f_mul(t[7], t[1]);                                  // ZZZ3
    __syncthreads();
}

template <class F>
__global__ void __launch_bounds__(128)
k_final_coop(const XYZZ<F>* __restrict__ R, const XYZZ<F>* __restrict__ Cs, int c, int W, int mode, F* __restrict__ out) {
    __shared__ XYZZ<F> Ssh[64];                  // S_w = R_w + Cs_w
    __shared__ XYZZ<F> acc;
    __shared__ F t[8];
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool lead = (tid & 31) == 0;
    if (tid < W) { XYZZ<F> s_ = load_vec_rw(R + tid); xyzz_add(s_, load_vec_rw(Cs + tid)); Ssh[tid] = s_; }
    if (tid == 0) acc = XYZZ<F>::inf();
    __syncthreads();
    for (int w = W - 1; w >= 0; --w) {
        for (int i = 0; i < c; ++i) coop_dbl(acc, t, warp, lead);
        coop_add(acc, Ssh[w], t, warp, lead, tid);
    }
    if (tid != 0) return;
    if (mode == 1) {
        out[0] = acc.x; out[1] = acc.y; out[2] = acc.zz; out[3] = acc.zzz;
        return;
    }
    if (acc.is_inf()) { out[0] = F::one(); out[1] = F::one(); out[2] = F::zero(); return; }
    Affine<F> a = xyzz_to_affine(acc);
    out[0] = a.x; out[1] = a.y; out[2] = F::one();
}

template <class F>
__global__ void k_identity(int mode, F* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (mode == 1) { out[0] = F::zero(); out[1] = F::zero(); out[2] = F::zero(); out[3] = F::zero(); }
    else { out[0] = F::one(); out[1] = F::one(); out[2] = F::zero(); }
}

// sum of XYZZ partials (multi-GPU combine) + normalisation
template <class F>
__global__ void k_combine(const XYZZ<F>* __restrict__ parts, size_t n, F* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (size_t i = 0; i < n; ++i) xyzz_add(acc, load_vec_rw(parts + i));
    if (acc.is_inf()) { out[0] = F::one(); out[1] = F::one(); out[2] = F::zero(); return; }
    Affine<F> a = xyzz_to_affine(acc);
    out[0] = a.x; out[1] = a.y; out[2] = F::one();
}

// ------------------------------------------------------------------------------------------
// fixed-base generation: out[i] = scalars[i] * G  (synthetic CRS; MsmMask::sample's gen * x)
// 4-bit fixed windows over a shared-memory table of (j * 16^w) * G would be faster; this is a
// test/bench data generator, a plain double-and-add with one inversion per point is enough.
// ------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ Affine<F> generator();
template <>
__device__ __forceinline__ Affine<Fq> generator<Fq>() {
    Affine<Fq> g;
#pragma unroll
    for (int i = 0; i < 8; ++i) { g.x.v[i] = BN254_G1_GEN_X_MONT_L(i); g.y.v[i] = BN254_G1_GEN_Y_MONT_L(i); }
    return g;
}
template <>
__device__ __forceinline__ Affine<Fq2> generator<Fq2>() {
    Affine<Fq2> g;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        g.x.c0.v[i] = BN254_G2_GEN_X_C0_MONT_L(i); g.x.c1.v[i] = BN254_G2_GEN_X_C1_MONT_L(i);
        g.y.c0.v[i] = BN254_G2_GEN_Y_C0_MONT_L(i); g.y.c1.v[i] = BN254_G2_GEN_Y_C1_MONT_L(i);
    }
    return g;
}

template <class F>
__global__ void __launch_bounds__(128) k_fixed_base(const Fr* __restrict__ scalars, size_t n, Affine<F>* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = fp_from_mont(load_vec(scalars + i));
    Affine<F> g = generator<F>();
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int b = 253; b >= 0; --b) {
        xyzz_dbl(acc);
        if ((s.v[b >> 5] >> (b & 31)) & 1) xyzz_madd(acc, g, false);
    }
    store_vec(out + i, xyzz_to_affine(acc));
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// The pipeline is split in host-side steps so that one MSM can be fed in CHUNKS -- a point range of the input (the next
// range still crossing PCIe) and/or a range of the scalar windows:
//   msm_plan              window size, workspace carve-up (two sets of digit/sort arrays sized for one chunk each)
//   msm_chunk_sort        digits -> counting sort -> size ordering of the chunk's bucket slots
//   msm_chunk_accumulate  bucket accumulation INTO the bucket array
//   msm_finish            bucket reduction + Horner + normalisation
// Sort pipeline (round 2): the counting sort is bound by L2 atomics (SMs 16 % busy) and the accumulation by the integer
// multiplier (L2 12 % busy), so when a call has more than one chunk every sort runs on the context's high-priority SIDE
// stream, into alternating sort sets, and the sort of chunk k+1 hides under the accumulation of chunk k; only the first
// chunk's sort is exposed.  A device-resident MSM is cut into two WINDOW groups for this (a short first group, so the
// exposed sort is short); the host-pointer paths pipeline their point-range chunks the same way.
struct MsmSortSet {
    uint32_t *digits = nullptr, *sorted = nullptr, *counts = nullptr, *cursor = nullptr, *order = nullptr, *shist = nullptr;
};
// ZKG_MSM_BOUNDS (experiment switch of the host-pointer paths): cumulative chunk ends in 64ths of the point range,
// e.g. "1,4,12,28,46,64"; at most 8 chunks.  Returns the number of chunks (0: not set) and fills bounds[0..K].
static inline int msm_env_bounds(size_t n, size_t* bounds) {
    const char* plan = getenv("ZKG_MSM_BOUNDS");
    if (!plan || !*plan || n < 4096) return 0;
    int K = 0;
    bounds[0] = 0;
    for (const char* q = plan; *q && K < 7;) {
        char* end = nullptr;
        long v = strtol(q, &end, 10);
        if (end == q) break;
        if (v > 64) v = 64;
        size_t b = v == 64 ? n : n / 64 * (size_t)v;
        if (b > bounds[K]) bounds[++K] = b;
        q = *end ? end + 1 : end;
    }
    if (K == 0 || bounds[K] != n) bounds[++K] = n;
    return K;
}

struct MsmChunkDesc {
    const Fr* d_scalars = nullptr;
    size_t n = 0, point0 = 0;      // point range [point0, point0 + n) of the call (d_scalars points at its first scalar)
    int w_lo = 0, w_hi = 0;        // scalar windows covered
    int into = 0;                  // 1: the chunk's buckets already hold earlier chunks' sums
};
template <class F>
struct MsmPlan {
    size_t n_total = 0, chunk_cap = 0;
    int c = 0, W = 0;          // scalar windows
    int Wb = 0;                // bucket sets: W (one per window) or 1 (merged: bases carry the window shifts)
    bool merged = false;
    uint32_t nb = 0, n1 = 0;
    uint32_t L = 8;            // level-0 segment of the bucket reduction (1: no level 0, the butterfly starts on the buckets)
    size_t slots = 0;          // Wb * nb
    MsmSortSet set[2];
    uint32_t hkey[2] = {SIZE_KEYS - 1, SIZE_KEYS - 1};      // heavy-bucket threshold of the chunk sorted into each set
    int n_sets = 1;
    bool side = false;         // sorts run on ctx->aux_stream (needs n_sets == 2)
    XYZZ<F>* buckets = nullptr;
    XYZZ<F>* Rb[2] = {nullptr, nullptr};
    XYZZ<F>* Cb[2] = {nullptr, nullptr};
    int sorts_issued = 0, chunks_done = 0;
};
static constexpr uint32_t MSM_REDUCE_L = 8;

// wg0 / wg1: the largest number of windows a chunk sorted into set 0 / set 1 covers (wg1 = 0: one set, no pipelining)
template <class F>
static int32_t msm_plan(zkg_ctx* ctx, size_t n_total, size_t chunk_cap, MsmPlan<F>* pl, int merged_c = 0, int c_forced = 0,
                        int wg0 = 0, int wg1 = 0) {
    ZKG_REQUIRE(n_total < ((size_t)1 << 31), "msm: n = %zu exceeds 2^31-1", n_total);
    pl->n_total = n_total;
    pl->chunk_cap = chunk_cap;
    int c = merged_c ? merged_c : c_forced;
    if (!c) {
        c = env_int("ZKG_MSM_C", 0);
        if (c < 2 || c > 22) c = msm_pick_c(n_total, sizeof(F) > 32);
    }
    pl->c = c;
    pl->W = msm_num_windows(c);
    pl->merged = merged_c != 0;
    pl->Wb = pl->merged ? 1 : pl->W;
    ZKG_REQUIRE(!pl->merged || n_total * (size_t)pl->W < ((size_t)1 << 31), "msm: n*W exceeds 2^31-1");
    pl->nb = 1u << (c - 1);
    pl->slots = (size_t)pl->Wb * pl->nb;
    // Level 0 (running sums over 8 buckets) is 15 DEPENDENT additions per thread, ~165 us on its own however few buckets
    // there are; small bucket sets skip it and run three more butterfly levels instead (~8 us each on four-warp blocks):
    // 2^13 G1 points 0.46 -> 0.41 ms, G2 0.92 -> 0.76 ms; from 2^15 buckets up the extra level work costs more than it saves.
    pl->L = (uint32_t)env_int("ZKG_MSM_REDUCE_L", pl->slots <= ((size_t)1 << 14) ? 1 : (int)MSM_REDUCE_L);
    if (pl->L != 1) pl->L = MSM_REDUCE_L;
    pl->n1 = (pl->nb + pl->L - 1) / pl->L;
    if (wg0 <= 0 || wg0 > pl->W) wg0 = pl->W;
    if (wg1 < 0 || wg1 > pl->W) wg1 = pl->W;
    pl->n_sets = wg1 > 0 ? 2 : 1;
    pl->side = false;
    const uint32_t scan_tiles = (pl->nb + SCAN_TILE - 1) / SCAN_TILE;
    ZKG_REQUIRE(scan_tiles <= 1024, "msm: window of %d bits too large", c);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_set[2][6] = {};
    for (int k = 0; k < pl->n_sets; ++k) {
        const size_t wg = (size_t)(k == 0 ? wg0 : wg1);
        const size_t set_slots = pl->merged ? (size_t)pl->nb : wg * pl->nb;
        o_set[k][0] = carve(sizeof(uint32_t) * wg * chunk_cap);        // digits
        o_set[k][1] = carve(sizeof(uint32_t) * wg * chunk_cap);        // sorted
        o_set[k][2] = carve(sizeof(uint32_t) * set_slots);             // counts
        o_set[k][3] = carve(sizeof(uint32_t) * set_slots);             // cursor
        o_set[k][4] = carve(sizeof(uint32_t) * set_slots);             // order
        o_set[k][5] = carve(sizeof(uint32_t) * (2 * SIZE_KEYS + HEAVY_LOCKS + (pl->merged ? 1 : wg) * scan_tiles));   // hist | locks | start | scan tiles
    }
    size_t o_buckets = carve(sizeof(XYZZ<F>) * pl->slots);
    size_t o_r0 = carve(sizeof(XYZZ<F>) * pl->Wb * pl->n1);
    size_t o_c0 = carve(sizeof(XYZZ<F>) * pl->Wb * pl->n1);
    size_t o_r1 = carve(sizeof(XYZZ<F>) * pl->Wb * pl->n1);
    size_t o_c1 = carve(sizeof(XYZZ<F>) * pl->Wb * pl->n1);
    ZKG_TRY(ctx->ws.reserve(off));
    uint8_t* ws = (uint8_t*)ctx->ws.p;
    for (int k = 0; k < pl->n_sets; ++k) {
        pl->set[k].digits = (uint32_t*)(ws + o_set[k][0]);
        pl->set[k].sorted = (uint32_t*)(ws + o_set[k][1]);
        pl->set[k].counts = (uint32_t*)(ws + o_set[k][2]);
        pl->set[k].cursor = (uint32_t*)(ws + o_set[k][3]);
        pl->set[k].order = (uint32_t*)(ws + o_set[k][4]);
        pl->set[k].shist = (uint32_t*)(ws + o_set[k][5]);
    }
    pl->buckets = (XYZZ<F>*)(ws + o_buckets);
    pl->Rb[0] = (XYZZ<F>*)(ws + o_r0); pl->Rb[1] = (XYZZ<F>*)(ws + o_r1);
    pl->Cb[0] = (XYZZ<F>*)(ws + o_c0); pl->Cb[1] = (XYZZ<F>*)(ws + o_c1);
    pl->sorts_issued = 0;
    pl->chunks_done = 0;
    return ZKG_OK;
}

// Turn the plan's sorts over to the side stream.  `ready` orders them after the call's inputs; the caller records it
// (msm_side_begin does, on the main stream, for device-resident inputs).
template <class F>
static int32_t msm_side_begin(zkg_ctx* ctx, MsmPlan<F>* pl) {
    if (pl->n_sets != 2 || !env_int("ZKG_MSM_SIDE", 1)) return ZKG_OK;
    ZKG_TRY(ctx_aux_stream(ctx));
    ZKG_CUDA(cudaEventRecord(ctx->aux_ev[0], ctx->stream));
    pl->side = true;
    return ZKG_OK;
}

// digits + counting sort + size ordering of one chunk.  ready: event the chunk's scalars wait for (nullptr: aux_ev[0],
// recorded by msm_side_begin at the start of the call); only used when the sorts run on the side stream.
template <class F>
static int32_t msm_chunk_sort(zkg_ctx* ctx, MsmPlan<F>* pl, const MsmChunkDesc& ch, cudaEvent_t ready = nullptr) {
    if (ch.n == 0) return ZKG_OK;
    const int k = pl->sorts_issued, si = pl->n_sets == 2 ? (k & 1) : 0;
    const MsmSortSet& ss = pl->set[si];
    cudaStream_t st = pl->side ? ctx->aux_stream : ctx->stream;
    const size_t n = ch.n;
    const int wg = ch.w_hi - ch.w_lo;
    const int wb = pl->merged ? 1 : wg;                        // bucket sets of this chunk
    const size_t slots = (size_t)wb * pl->nb;
    {
        const int hk = env_int("ZKG_MSM_HEAVY_KEY", 0);
        pl->hkey[si] = hk >= 2 && hk < (int)SIZE_KEYS ? (uint32_t)hk : msm_heavy_key((double)n * wg / (double)slots);
    }
    const uint32_t hkey = pl->hkey[si];
    if (k == 0) phase_mark(ctx, 0);
    if (pl->side) {
        ZKG_CUDA(cudaStreamWaitEvent(st, ready ? ready : ctx->aux_ev[0], 0));
        if (k >= 2) ZKG_CUDA(cudaStreamWaitEvent(st, ctx->aux_ev[3 + si], 0));      // the set's previous chunk has been accumulated
    }
    ZKG_CUDA(cudaMemsetAsync(ss.counts, 0, sizeof(uint32_t) * slots, st));
    ZKG_CUDA(cudaMemsetAsync(ss.shist, 0, sizeof(uint32_t) * (SIZE_KEYS + HEAVY_LOCKS), st));
    const uint32_t bstride = pl->merged ? 0u : pl->nb;
    const size_t sstride = pl->merged ? 0 : n, ioff = pl->merged ? pl->n_total : 0;
    // a sort that runs under an accumulation (every chunk but the first of a pipelined call) gets a thin grid
    const bool hidden = pl->side && k > 0;
    const int TB = hidden ? env_int("ZKG_MSM_SORT_TB_HIDDEN", 256) : 256;
    unsigned sort_grid = (unsigned)((n + TB - 1) / TB);
    {
        const unsigned cap = (unsigned)ctx->sm_count * (unsigned)env_int(hidden ? "ZKG_MSM_SORT_BPS_HIDDEN" : "ZKG_MSM_SORT_BPS", hidden ? (n * (size_t)wg >= ((size_t)1 << 27) ? 2 : 1) : 32);
        if (cap && sort_grid > cap) sort_grid = cap;
    }
    {
        const int dtb = hidden ? env_int("ZKG_MSM_DIGITS_TB_HIDDEN", TB) : TB;
        unsigned dgrid = (unsigned)((n + dtb - 1) / dtb);
        if (dgrid > sort_grid) dgrid = sort_grid;
        if (dtb <= 128) k_digits<1><<<dgrid, dtb, 0, st>>>(ch.d_scalars, n, pl->c, pl->W, ch.w_lo, ch.w_hi, bstride, ss.digits, ss.counts);
        else k_digits<0><<<dgrid, dtb, 0, st>>>(ch.d_scalars, n, pl->c, pl->W, ch.w_lo, ch.w_hi, bstride, ss.digits, ss.counts);
    }
    {
        const uint32_t tiles = (pl->nb + SCAN_TILE - 1) / SCAN_TILE;      // <= 1024 (nb <= 2^22)
        uint32_t* tile_sum = ss.shist + 2 * SIZE_KEYS + HEAVY_LOCKS;
        k_scan_partial<<<dim3(tiles, wb), 256, 0, st>>>(ss.counts, pl->nb, tiles, tile_sum);
        k_scan_tiles<<<wb, 1024, 0, st>>>(tile_sum, tiles);
        k_scan_final<<<dim3(tiles, wb), 256, 0, st>>>(ss.counts, pl->nb, tiles, tile_sum, ss.cursor);
    }
    // (measured and dropped: scattering one window at a time, to keep the destination region L2-sized, changes nothing --
    //  1.28 vs 1.20 ms at 2^22 -- so the counting sort is not bound by the footprint of its scattered 4-byte stores;
    //  a two-level variant -- block-aggregated scatter into <= 1024 partitions of consecutive buckets, then one block per
    //  partition with shared-memory cursors -- was parity-green and twice as slow, 2.2 vs 1.15 ms: its first pass still
    //  issues one isolated store per entry, now 8 bytes, and its second pays a shared-memory atomic per entry;
    //  letting the histogram atomic of k_digits return the entry's rank, so that this scatter needs no atomic, is slower
    //  too -- 2.2 ms on the prepared path: an atomic WITH a return value costs more than the fire-and-forget reduction
    //  the histogram compiles to now, +0.18 ms for k_digits alone;
    //  fixed-capacity bucket lists -- no histogram, no scan, one returned atomic and one store per entry -- bring the
    //  phase from 1.14 to 0.89 ms (5.2 -> 4.3 ms at 2^24) but need an overflow path for every non-uniform input;
    //  2.5 % of a step, not built)
    const int ilp = hidden ? env_int("ZKG_MSM_SCATTER_ILP_HIDDEN", 4) : env_int("ZKG_MSM_SCATTER_ILP", 4);
    if (ilp == 8)
        k_scatter<8><<<sort_grid, TB, 0, st>>>(ss.digits, n, wg, bstride, sstride, ioff, ch.point0, ch.w_lo, ss.cursor, ss.sorted);
    else if (ilp == 1)
        k_scatter<1><<<sort_grid, TB, 0, st>>>(ss.digits, n, wg, bstride, sstride, ioff, ch.point0, ch.w_lo, ss.cursor, ss.sorted);
    else
        k_scatter<4><<<sort_grid, TB, 0, st>>>(ss.digits, n, wg, bstride, sstride, ioff, ch.point0, ch.w_lo, ss.cursor, ss.sorted);
    unsigned hb = (unsigned)((slots + 1023) / 1024);
    if (hb > 592) hb = 592;
    uint32_t* sstart = ss.shist + SIZE_KEYS + HEAVY_LOCKS;
    k_size_hist<<<hb, 256, 0, st>>>(ss.counts, slots, hkey, ss.shist);
    k_size_scan<<<1, SIZE_KEYS / 2, 0, st>>>(ss.shist, sstart);
    k_size_scatter<<<(unsigned)((slots + 1023) / 1024), 1024, 0, st>>>(ss.counts, slots, hkey, sstart, ss.order);
    if (pl->side) ZKG_CUDA(cudaEventRecord(ctx->aux_ev[1 + si], st));
    pl->sorts_issued += 1;
    ctx->launches += 8;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

static inline void msm_launch_g2pair(const Affine<Fq2>* bases, const uint32_t* sorted, const uint32_t* cursor, const uint32_t* counts,
                                     const uint32_t* order, size_t sstride, uint32_t nb, int Wb, int into, uint32_t hkey, XYZZ<Fq2>* buckets, size_t slots,
                                     cudaStream_t st) {
    const unsigned grid = (unsigned)((2 * slots + 127) / 128);
    if (env_int("ZKG_MSM_G2_PAIR_BLOCKS", 4) == 3)
        k_accumulate_g2pair<3><<<grid, 128, 0, st>>>(bases, sorted, cursor, counts, order, sstride, nb, Wb, into, hkey, buckets);
    else
        k_accumulate_g2pair<4><<<grid, 128, 0, st>>>(bases, sorted, cursor, counts, order, sstride, nb, Wb, into, hkey, buckets);
}
static inline void msm_launch_g2pair(const Affine<Fq>*, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, size_t, uint32_t,
                                     int, int, uint32_t, XYZZ<Fq>*, size_t, cudaStream_t) {}      // never taken: sizeof(Fq) == 32

// second half of a chunk: bucket accumulation of the points sorted by msm_chunk_sort (needs the bases).  Chunks are
// accumulated in the order they were sorted.
template <class F>
static int32_t msm_chunk_accumulate(zkg_ctx* ctx, MsmPlan<F>* pl, const MsmChunkDesc& ch, const Affine<F>* d_bases) {
    if (ch.n == 0) return ZKG_OK;
    const int k = pl->chunks_done, si = pl->n_sets == 2 ? (k & 1) : 0;
    const MsmSortSet& ss = pl->set[si];
    cudaStream_t st = ctx->stream;
    const size_t n = ch.n;
    const int wg = ch.w_hi - ch.w_lo;
    const int wb = pl->merged ? 1 : wg;
    const size_t slots = (size_t)wb * pl->nb;
    XYZZ<F>* buckets = pl->buckets + (pl->merged ? 0 : (size_t)ch.w_lo * pl->nb);
    const uint32_t hkey = pl->hkey[si];
    if (pl->side) ZKG_CUDA(cudaStreamWaitEvent(st, ctx->aux_ev[1 + si], 0));
    if (k == 0) phase_mark(ctx, 1);
    const size_t sstride = pl->merged ? 0 : n;
    const int use_ba = env_int("ZKG_MSM_BA", 0);            // read per call, like ZKG_MSM_C (tests flip it inside one process)
    // G2: the lane-pair kernel wins while the launch is short of threads (2^16 points: accumulate 0.71 -> 0.60 ms) and loses
    // ~5 % to its shuffles and selects once the one-thread-per-bucket kernel fills the machine (2^19: 3.52 vs 3.70 ms)
    const int g2_pair = env_int("ZKG_MSM_G2_PAIR", pl->n_total <= ((size_t)1 << 17) ? 1 : 0);
    if (sizeof(F) > 32 && !use_ba && g2_pair)
        msm_launch_g2pair(d_bases, ss.sorted, ss.cursor, ss.counts, ss.order, sstride, pl->nb, wb, ch.into, hkey, buckets, slots, st);
    else if (use_ba)
        k_accumulate_ba<F><<<(unsigned)((slots + BA_THREADS - 1) / BA_THREADS), BA_THREADS, 0, st>>>(
            d_bases, ss.sorted, ss.cursor, ss.counts, ss.order, sstride, pl->nb, wb, ch.into, hkey, buckets);
    else
        k_accumulate<F><<<(unsigned)((slots + 127) / 128), 128, 0, st>>>(d_bases, ss.sorted, ss.cursor, ss.counts, ss.order,
                                                                       sstride, pl->nb, wb, ch.into, hkey, env_int("ZKG_MSM_PREFETCH", 1), buckets);
    {
        size_t max_heavy = (n * (size_t)wg) / hkey;
        if (max_heavy > slots) max_heavy = slots;
        if (max_heavy > 0) {
            unsigned hg = (unsigned)(max_heavy < 148 ? max_heavy : 148);
            k_accumulate_heavy<F><<<dim3(hg, HEAVY_PARTS), 128, 0, st>>>(d_bases, ss.sorted, ss.cursor, ss.counts, ss.order, ss.shist,
                                                                       ss.shist + SIZE_KEYS, sstride, pl->nb, hkey, buckets);
            ctx->launches += 1;
        }
    }
    if (pl->side) ZKG_CUDA(cudaEventRecord(ctx->aux_ev[3 + si], st));
    ctx->launches += 1;
    pl->chunks_done += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

// Device-resident inputs: the whole point range at once.  Large MSMs are cut into WINDOW groups so that each group's
// sort hides under the previous group's accumulation (merged plans: the groups share the buckets, later groups are
// accumulated into them; per-window plans: the groups own disjoint bucket sets).
// groups[]: window boundaries 0 = g[0] < g[1] < ... < g[ng] = W; ng = 1: no pipelining.
static inline int msm_window_groups(size_t n, int W, int* g) {
    int g0 = env_int("ZKG_MSM_GROUP0", -1);
    const int ng_env = env_int("ZKG_MSM_GROUPS", 2);
    // measured at 2^22 G1 points (W = 13): first group of 4 windows 9.76 -> 9.36 ms; below ~2^25 entries the second
    // sort no longer fits under the first group's accumulation and the split only costs (2^20: 2.96 -> 2.99 ms)
    if (g0 < 0) g0 = (n * (size_t)W >= ((size_t)1 << 25) && W >= 6) ? (W + 1) / 3 : 0;
    g[0] = 0;
    if (g0 <= 0 || g0 >= W) { g[1] = W; return 1; }
    if (ng_env >= 3 && W - g0 >= 2) { g[1] = g0; g[2] = g0 + (W - g0) / 2; g[3] = W; return 3; }
    g[1] = g0; g[2] = W;
    return 2;
}
template <class F>
static int32_t msm_chunks_device(zkg_ctx* ctx, MsmPlan<F>* pl, const Affine<F>* d_bases, const Fr* d_scalars, size_t n, const int* g, int ng) {
    MsmChunkDesc ch[4];
    for (int k = 0; k < ng; ++k) {
        ch[k].d_scalars = d_scalars; ch[k].n = n; ch[k].point0 = 0;
        ch[k].w_lo = g[k]; ch[k].w_hi = g[k + 1];
        ch[k].into = (k > 0 && pl->merged) ? 1 : 0;
    }
    if (ng > 1) ZKG_TRY(msm_side_begin<F>(ctx, pl));
    if (!pl->side) {
        for (int k = 0; k < ng; ++k) {
            ZKG_TRY(msm_chunk_sort<F>(ctx, pl, ch[k]));
            ZKG_TRY(msm_chunk_accumulate<F>(ctx, pl, ch[k], d_bases));
        }
        return ZKG_OK;
    }
    // two sort sets: the side stream runs at most two sorts ahead of the accumulations
    int sorted = 0;
    for (int k = 0; k < ng; ++k) {
        while (sorted < ng && sorted < k + 2) ZKG_TRY(msm_chunk_sort<F>(ctx, pl, ch[sorted++]));
        ZKG_TRY(msm_chunk_accumulate<F>(ctx, pl, ch[k], d_bases));
    }
    return ZKG_OK;
}
// largest window count of the groups sorted into set 0 (even groups) / set 1 (odd groups)
static inline void msm_group_caps(const int* g, int ng, int W, int* wg0, int* wg1) {
    *wg0 = 0; *wg1 = 0;
    for (int k = 0; k < ng; ++k) {
        int w = g[k + 1] - g[k];
        if (k & 1) { if (w > *wg1) *wg1 = w; } else { if (w > *wg0) *wg0 = w; }
    }
    if (ng == 1) { *wg0 = W; *wg1 = 0; }
}

template <class F>
static int32_t msm_finish(zkg_ctx* ctx, MsmPlan<F>* pl, F* d_out, int mode) {
    cudaStream_t st = ctx->stream;
    if (pl->chunks_done == 0) {
        k_identity<F><<<1, 32, 0, st>>>(mode, d_out);
        ctx->launches += 1;
        ZKG_CUDA(cudaGetLastError());
        return ZKG_OK;
    }
    phase_mark(ctx, 2);                                     // every chunk's accumulation is in the stream by now
    // level 0: running sums over segments of 8 buckets (throughput-bound: two adds per bucket)
    const uint32_t n1 = pl->n1;
    const XYZZ<F>* R0 = pl->Rb[0];
    const XYZZ<F>* A0 = pl->Cb[0];
    int lshift = 0;
    if (pl->L > 1) {
        size_t th = (size_t)pl->Wb * n1;
        k_reduce_lvl<F><<<(unsigned)((th + 127) / 128), 128, 0, st>>>(pl->buckets, nullptr, pl->nb, pl->L, 0, pl->Wb, pl->Rb[0], pl->Cb[0], n1);
        ctx->launches += 1;
        while (((uint32_t)1 << lshift) < pl->L) ++lshift;
    } else {
        R0 = pl->buckets;                                   // n1 = nb entries (R_u = B_u, A_u = identity)
        A0 = nullptr;
    }
    // log-depth butterfly over the n1 segment results (n1 is a power of two); ping-pong between the
    // two halves of the reduction scratch (each half = 2 * Wb * n1 values, enough for every level)
    int B = 0;
    while (((uint32_t)1 << B) < n1) ++B;
    const int coop_red = env_int("ZKG_MSM_COOP_REDUCE", 1);  // four-warp blocks for the latency-bound levels
    XYZZ<F>* buf[2] = {pl->Rb[1], pl->Rb[0]};              // level 1 -> second half, level 2 -> first half (level-0 data is dead by then)
    const XYZZ<F>* cur = nullptr;
    uint32_t n_prev = n1;
    for (int lvl = 1; lvl <= B; ++lvl) {
        XYZZ<F>* dst = buf[(lvl - 1) & 1];
        size_t th = (size_t)pl->Wb * (n_prev >> 1) * (lvl + 2);
        if (coop_red) k_merge_lvl_coop<F><<<(unsigned)((th + 31) / 32), 128, 0, st>>>(cur, R0, A0, dst, lvl, n_prev, pl->Wb);
        else k_merge_lvl<F><<<(unsigned)((th + 127) / 128), 128, 0, st>>>(cur, R0, A0, dst, lvl, n_prev, pl->Wb);
        ctx->launches += 1;
        cur = dst;
        n_prev >>= 1;
    }
    XYZZ<F>* S_arr = B == 0 ? pl->Rb[1] : buf[B & 1];       // the buffer the last level did not write
    XYZZ<F>* Z_arr = S_arr + pl->Wb;
    if (coop_red) k_bits_final_coop<F><<<pl->Wb, 128, 0, st>>>(cur, R0, A0, B, lshift, S_arr, Z_arr);
    else k_bits_final<F><<<pl->Wb, 32, 0, st>>>(cur, R0, A0, B, lshift, S_arr, Z_arr);
    ctx->launches += 1;
    const XYZZ<F>* Rin = S_arr;
    const XYZZ<F>* Cin = Z_arr;
    const int coop_tail = env_int("ZKG_MSM_COOP_TAIL", 1);
    // (the same Horner on the 32-wide word-major primitives of the merge levels measured no faster: 0.75 vs 0.74 ms at c = 5,
    //  W = 51, and 0.17 ms slower on G2 -- a lone chain is bound by the product latency, not by the operand traffic)
    if (!pl->merged && pl->Wb > 1 && pl->Wb <= 64 && coop_tail)
        k_final_coop<F><<<1, 128, 0, st>>>(Rin, Cin, pl->c, pl->Wb, mode, d_out);       // Horner over windows on four warps
    else
        k_final<F><<<1, 32, 0, st>>>(Rin, Cin, pl->merged ? 0 : pl->c, pl->Wb, mode, d_out);
    ctx->launches += 1;
    phase_mark(ctx, 3);
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

template <class F>
static int32_t msm_run(zkg_ctx* ctx, const Affine<F>* d_bases, const Fr* d_scalars, size_t n, F* d_out, int mode) {
    MsmPlan<F> pl;
    if (n) {
        int c = env_int("ZKG_MSM_C", 0);
        if (c < 2 || c > 22) c = msm_pick_c(n, sizeof(F) > 32);
        const int W = msm_num_windows(c);
        int g[5], wg0, wg1;
        const int ng = msm_window_groups(n, W, g);
        msm_group_caps(g, ng, W, &wg0, &wg1);
        ZKG_TRY(msm_plan<F>(ctx, n, n, &pl, 0, c, wg0, wg1));
        ZKG_TRY(msm_chunks_device<F>(ctx, &pl, d_bases, d_scalars, n, g, ng));
    }
    return msm_finish<F>(ctx, &pl, d_out, mode);
}

// ------------------------------------------------------------------------------------------
// Prepared bases (static CRS shares): table[w*n + i] = 2^(c*w) * P_i in packed affine form.
// One thread per point walks the windows: c doublings, one inversion per stored copy.  One-time
// cost (~7 k multiplications per point); it buys MSMs with a single bucket set and no Horner tail.
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_prepare_bases(const Affine<F>* __restrict__ bases, size_t n, int c, int W, Affine<F>* __restrict__ table) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = load_vec(bases + i);
    store_vec(table + i, p);
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    for (int w = 1; w < W; ++w) {
        xyzz_dbl_k(acc, c);
        store_vec(table + (size_t)w * n + i, xyzz_to_affine(acc));
    }
}

template <class F>
static int32_t msm_prepare(zkg_ctx* ctx, const Affine<F>* d_bases, size_t n, int c, Affine<F>* d_table) {
    if (n == 0) return ZKG_OK;
    k_prepare_bases<F><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_bases, n, c, msm_num_windows(c), d_table);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

// MSM against a prepared table (n points, window size c fixed at preparation time)
template <class F>
static int32_t msm_run_prepared(zkg_ctx* ctx, const Affine<F>* d_table, int c, const Fr* d_scalars, size_t n, F* d_out, int mode) {
    MsmPlan<F> pl;
    if (n) {
        const int W = msm_num_windows(c);
        int g[5], wg0, wg1;
        const int ng = msm_window_groups(n, W, g);
        msm_group_caps(g, ng, W, &wg0, &wg1);
        ZKG_TRY(msm_plan<F>(ctx, n, n, &pl, c, 0, wg0, wg1));
        ZKG_TRY(msm_chunks_device<F>(ctx, &pl, d_table, d_scalars, n, g, ng));
    }
    return msm_finish<F>(ctx, &pl, d_out, mode);
}

// Registered bases + HOST scalars: the scalars cross PCIe in quarters on the copy stream while the
// previous quarter is being sorted and accumulated (merged plans take any point range of the table);
// the sort of quarter j+1 (side stream) runs under the accumulation of quarter j.
template <class F>
static int32_t msm_run_prepared_host(zkg_ctx* ctx, const Affine<F>* d_table, int c, const uint64_t* h_scalars, size_t n,
                                     F* d_out, int mode = 0) {
    MsmPlan<F> pl;
    if (n) {
        // graded chunks (1/16, 3/16, 1/4, 1/4, 1/4): the first copy, which nothing hides, is short
        size_t bounds[10] = {0, n, n, n, n, n, n, n, n, n};
        int K = msm_env_bounds(n, bounds);
        if (K == 0) {
            K = 1;
            bounds[1] = n;
            if (n >= ((size_t)1 << 18)) { K = 5; bounds[1] = n / 16; bounds[2] = n / 4; bounds[3] = n / 2; bounds[4] = n / 4 * 3; bounds[5] = n; }
        }
        size_t chunk = 0;
        for (int j = 0; j < K; ++j) if (bounds[j + 1] - bounds[j] > chunk) chunk = bounds[j + 1] - bounds[j];
        ZKG_TRY(ctx->io.reserve(align_up(n * 32, 256) + 512));
        uint8_t* d_sc = (uint8_t*)ctx->io.p;
        const int W = msm_num_windows(c);
        ZKG_TRY(msm_plan<F>(ctx, n, chunk, &pl, c, 0, W, K > 1 ? W : 0));
        ZKG_TRY(ctx_copy_stream(ctx, K));
        ZKG_CUDA(cudaEventRecord(ctx->copy_ev[0], ctx->stream));
        ZKG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
        if (K > 1) ZKG_TRY(msm_side_begin<F>(ctx, &pl));
        MsmChunkDesc ch[8];
        // A chunk's copy is issued right before its sort is enqueued (the sort waits for it on the side stream), and the side
        // stream runs at most two sorts ahead of the accumulations (two sort sets).  With pinned scalars the copies are
        // asynchronous and simply run ahead; with PAGEABLE scalars copy_h2d stages through the pinned slots on this thread, so
        // issuing every copy up front would hold back the first accumulation until the whole vector has been staged
        // (2^22 scalars: 12.2 ms per call instead of 10.x).
        for (int j = 0; j < K; ++j) {
            size_t lo = bounds[j], hi = bounds[j + 1];
            ch[j].d_scalars = (const Fr*)d_sc + lo; ch[j].n = hi - lo; ch[j].point0 = lo;
            ch[j].w_lo = 0; ch[j].w_hi = W; ch[j].into = 0;
        }
        auto copy_chunk = [&](int j) -> int32_t {
            size_t lo = bounds[j], hi = bounds[j + 1];
            ZKG_TRY(copy_h2d(d_sc + lo * 32, (const uint8_t*)h_scalars + lo * 32, (hi - lo) * 32, ctx->copy_stream));
            ZKG_CUDA(cudaEventRecord(ctx->copy_ev[j], ctx->copy_stream));
            return ZKG_OK;
        };
        if (!pl.side) {
            for (int j = 0; j < K; ++j) {
                if (ch[j].n == 0) continue;
                ZKG_TRY(copy_chunk(j));
                ZKG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[j], 0));
                ch[j].into = pl.chunks_done > 0 ? 1 : 0;
                ZKG_TRY(msm_chunk_sort<F>(ctx, &pl, ch[j]));
                ZKG_TRY(msm_chunk_accumulate<F>(ctx, &pl, ch[j], d_table));
            }
        } else {
            int sorted = 0, accd = 0;
            auto next = [&](int j) { while (j < K && ch[j].n == 0) ++j; return j; };
            sorted = next(0); accd = next(0);
            int ahead = 0;
            while (accd < K) {
                while (sorted < K && ahead < 2) {
                    ZKG_TRY(copy_chunk(sorted));
                    ZKG_TRY(msm_chunk_sort<F>(ctx, &pl, ch[sorted], ctx->copy_ev[sorted]));
                    sorted = next(sorted + 1); ++ahead;
                }
                ch[accd].into = pl.chunks_done > 0 ? 1 : 0;
                ZKG_TRY(msm_chunk_accumulate<F>(ctx, &pl, ch[accd], d_table));
                accd = next(accd + 1); --ahead;
            }
        }
    }
    return msm_finish<F>(ctx, &pl, d_out, mode);
}

template <class F>
static int32_t pack_bases(zkg_ctx* ctx, const void* d_ark, size_t stride, size_t n, void* d_packed);

// ------------------------------------------------------------------------------------------
// CRS share pre-processing: PackedSharingParams::det_pack over GROUP elements,
// as pack_from_arkworks_proving_key does for every l-chunk of the proving key
// (groth16/src/proving_key.rs:72-104, secret-sharing/src/pss.rs:69-87).  det_pack is the linear map
// share_i = sum_{k<l} M[i][k] * P_k with M the first l columns of the pack matrix, so one thread per
// (chunk, party) runs an interleaved double-and-add over the l canonical 254-bit scalars and
// normalises -- the result is the same group element arkworks' FFT-over-points computes.
// scal: n_parties x l canonical scalars (8 x u32 each).  out[i * chunks + j] = party i's share of chunk j.
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_det_pack_group(const Affine<F>* __restrict__ bases, size_t chunks, int l, int n_parties, const uint32_t* __restrict__ scal,
                 Affine<F>* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * (size_t)n_parties) return;
    size_t j = t % chunks;
    int i = (int)(t / chunks);
    const uint32_t* sc = scal + (size_t)i * l * 8;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int b = 253; b >= 0; --b) {
        xyzz_dbl(acc);
        for (int k = 0; k < l; ++k)
            if ((sc[k * 8 + (b >> 5)] >> (b & 31)) & 1) xyzz_madd(acc, load_vec(bases + j * l + k), false);
    }
    store_vec(out + t, xyzz_to_affine(acc));
}

// packed (x, y) -> arkworks Affine image (coordinates, infinity flag, zero padding up to `stride`)
template <class F>
__global__ void k_unpack_bases(const Affine<F>* __restrict__ in, size_t n, uint8_t* __restrict__ ark, size_t stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> a = load_vec_rw(in + i);
    uint8_t* p = ark + i * stride;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(&a);
    for (size_t k = 0; k < sizeof(Affine<F>); ++k) p[k] = src[k];
    p[sizeof(Affine<F>)] = a.is_inf() ? 1 : 0;
    for (size_t k = sizeof(Affine<F>) + 1; k < stride; ++k) p[k] = 0;
}

// host-pointer entry: bases (n = chunks*l arkworks images) -> n_parties share vectors of `chunks` images
template <class F>
static int32_t crs_det_pack_host(int device, const void* bases, size_t stride, size_t n, int l, int n_parties,
                                 const uint32_t* h_scal, void* const* out_by_party, size_t out_stride) {
    ZKG_REQUIRE(stride >= sizeof(Affine<F>) + 1 && out_stride >= sizeof(Affine<F>) + 1, "crs_det_pack: stride too small");
    ZKG_REQUIRE(l > 0 && n % (size_t)l == 0, "crs_det_pack: %zu bases is not a multiple of l = %d", n, l);
    if (n == 0) return ZKG_OK;
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    const size_t chunks = n / l, outs = chunks * (size_t)n_parties;
    size_t ark_b = align_up(n * stride, 256), pk_b = align_up(n * sizeof(Affine<F>), 256), sc_b = align_up((size_t)n_parties * l * 32, 256);
    size_t o_pk = align_up(outs * sizeof(Affine<F>), 256), o_ark = align_up(outs * out_stride, 256);
    ZKG_TRY(ctx->io.reserve(ark_b + pk_b + sc_b + o_pk + o_ark));
    uint8_t* d = (uint8_t*)ctx->io.p;
    uint8_t *d_ark = d, *d_pk = d + ark_b, *d_sc = d_pk + pk_b, *d_opk = d_sc + sc_b, *d_oark = d_opk + o_pk;
    ZKG_TRY(copy_h2d(d_ark, bases, n * stride, ctx->stream));
    ZKG_TRY(copy_h2d(d_sc, h_scal, (size_t)n_parties * l * 32, ctx->stream));
    ZKG_TRY(pack_bases<F>(ctx, d_ark, stride, n, d_pk));
    k_det_pack_group<F><<<(unsigned)((outs + 127) / 128), 128, 0, ctx->stream>>>((const Affine<F>*)d_pk, chunks, l, n_parties,
                                                                              (const uint32_t*)d_sc, (Affine<F>*)d_opk);
    k_unpack_bases<F><<<(unsigned)((outs + 255) / 256), 256, 0, ctx->stream>>>((const Affine<F>*)d_opk, outs, d_oark, out_stride);
    ctx->launches += 2;
    ZKG_CUDA(cudaGetLastError());
    for (int i = 0; i < n_parties; ++i) {
        ZKG_REQUIRE(out_by_party[i], "crs_det_pack: NULL output for party %d", i);
        ZKG_TRY(copy_d2h(out_by_party[i], d_oark + (size_t)i * chunks * out_stride, chunks * out_stride, ctx->stream));
    }
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

template <class F>
static int32_t pack_bases(zkg_ctx* ctx, const void* d_ark, size_t stride, size_t n, void* d_packed) {
    if (n == 0) return ZKG_OK;
    ZKG_REQUIRE(stride >= sizeof(Affine<F>) + 1, "base stride %zu too small", stride);
    k_pack_bases<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)d_ark, stride, n, (Affine<F>*)d_packed);
    ctx->launches += 1;
    ZKG_CUDA(cudaGetLastError());
    return ZKG_OK;
}

// blocking host-pointer MSM (both curves).  Large inputs are fed in point-range chunks: the H2D
// copy of chunk j+1 (copy stream) overlaps digits/sort/accumulate of chunk j (compute stream), so
// the PCIe transfer of the 72+32 B/point inputs hides behind the bucket accumulation.
// Enqueues the whole host-pointer MSM of `n` points on `ctx` (copies, sort, accumulation, reduction) and leaves the result in
// the context's staging area: mode 0 = normalised Jacobian image (3 F), mode 1 = XYZZ partial (4 F) for multi-GPU combination.
// Does not synchronise; *d_result stays valid until the next call on the context.
template <class F>
static int32_t msm_host_enqueue(zkg_ctx* ctx, const void* bases, size_t stride, size_t n, const uint64_t* scalars, int mode,
                                F** d_result) {
    ZKG_REQUIRE(n == 0 || (bases && scalars), "msm: NULL input");
    ZKG_REQUIRE(n == 0 || stride >= sizeof(Affine<F>) + 1, "base stride %zu too small", stride);
    // Chunk boundaries (2^20 points: 1/16, 3/16, 1/4, 1/4, 1/4; from 2^21: eight graded chunks): the first copy,
    // which nothing can hide, is short; every later copy is covered by the previous chunk's digits/sort/accumulate.
    // ZKG_MSM_CHUNKS=k forces k equal chunks, ZKG_MSM_BOUNDS an explicit plan.
    size_t bounds[18];
    int K = env_int("ZKG_MSM_CHUNKS", 0);
    if (K > 8) K = 8;
    const int Kenv = msm_env_bounds(n, bounds);
    if (Kenv > 0) {
        K = Kenv;
    } else if (K > 0) {
        for (int j = 0; j <= K; ++j) bounds[j] = n * (size_t)j / (size_t)K;
    } else if (n >= ((size_t)1 << 21)) {
        // the accumulation (2.25 ns/point) is slower than the copy (1.9 ns/point): a tiny first chunk starts the compute
        // stream early and growing chunks keep it fed -- measured at 2^22: 12.13 ms against 12.53 ms with five chunks
        static const int ends[9] = {0, 1, 3, 8, 16, 28, 40, 52, 64};
        K = 8;
        for (int j = 0; j <= K; ++j) bounds[j] = ends[j] == 64 ? n : n / 64 * (size_t)ends[j];
    } else if (n >= ((size_t)1 << 20)) {
        K = 5;
        bounds[0] = 0; bounds[1] = n / 16; bounds[2] = n / 4; bounds[3] = n / 2; bounds[4] = n / 4 * 3; bounds[5] = n;
    } else if (n >= ((size_t)1 << 17)) {
        K = 2;
        bounds[0] = 0; bounds[1] = n / 4; bounds[2] = n;
    } else {
        K = 1;
        bounds[0] = 0; bounds[1] = n;
    }
    size_t chunk_cap = 0;
    for (int j = 0; j < K; ++j) if (bounds[j + 1] - bounds[j] > chunk_cap) chunk_cap = bounds[j + 1] - bounds[j];
    size_t ark_bytes = align_up(n * stride, 256), sc_bytes = align_up(n * 32, 256), pk_bytes = align_up(n * sizeof(Affine<F>), 256);
    ZKG_TRY(ctx->io.reserve(ark_bytes + sc_bytes + pk_bytes + 256));
    uint8_t* d_ark = (uint8_t*)ctx->io.p;
    uint8_t* d_sc = d_ark + ark_bytes;
    uint8_t* d_pk = d_sc + sc_bytes;
    F* d_out = (F*)(d_pk + pk_bytes);
    MsmPlan<F> pl;
    if (n) {
        int c = env_int("ZKG_MSM_C", 0);
        if (c < 2 || c > 22) c = msm_pick_c(n, sizeof(F) > 32);
        const int W = msm_num_windows(c);
        ZKG_TRY(msm_plan<F>(ctx, n, chunk_cap, &pl, 0, c, W, K > 1 ? W : 0));
        ZKG_TRY(ctx_copy_stream(ctx, 2 * K));
        // order the copy stream after whatever the compute stream last did with these buffers
        ZKG_CUDA(cudaEventRecord(ctx->copy_ev[0], ctx->stream));
        ZKG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
        if (K > 1) ZKG_TRY(msm_side_begin<F>(ctx, &pl));
        // copy chunk j, then launch its compute, then copy chunk j+1 ...: with pinned sources the copies
        // simply run ahead on the copy stream; with pageable sources the host thread stages chunk j+1
        // while the device works on chunk j.  The sorts go to the side stream: chunk j+1's scalars arrive
        // right after chunk j's bases, so its sort runs under chunk j's accumulation.
        for (int j = 0; j < K; ++j) {
            size_t lo = bounds[j], hi = bounds[j + 1];
            if (lo >= hi) continue;
            MsmChunkDesc ch;
            ch.d_scalars = (const Fr*)d_sc + lo; ch.n = hi - lo; ch.point0 = 0;
            ch.w_lo = 0; ch.w_hi = W; ch.into = pl.chunks_done > 0 ? 1 : 0;
            // the scalars go first: digits and the counting sort need nothing else, so the chunk's bases cross
            // PCIe while its own sort runs
            ZKG_TRY(copy_h2d(d_sc + lo * 32, (const uint8_t*)scalars + lo * 32, (hi - lo) * 32, ctx->copy_stream));
            ZKG_CUDA(cudaEventRecord(ctx->copy_ev[2 * j], ctx->copy_stream));
            if (!pl.side) ZKG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[2 * j], 0));
            ZKG_TRY(msm_chunk_sort<F>(ctx, &pl, ch, ctx->copy_ev[2 * j]));      // launched before the (possibly host-staged) bases copy
            ZKG_TRY(copy_h2d(d_ark + lo * stride, (const uint8_t*)bases + lo * stride, (hi - lo) * stride, ctx->copy_stream));
            ZKG_CUDA(cudaEventRecord(ctx->copy_ev[2 * j + 1], ctx->copy_stream));
            ZKG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[2 * j + 1], 0));
            ZKG_TRY(pack_bases<F>(ctx, d_ark + lo * stride, stride, hi - lo, d_pk + lo * sizeof(Affine<F>)));
            ZKG_TRY(msm_chunk_accumulate<F>(ctx, &pl, ch, (const Affine<F>*)d_pk + lo));
        }
    }
    ZKG_TRY(msm_finish<F>(ctx, &pl, d_out, mode));
    *d_result = d_out;
    return ZKG_OK;
}

template <class F>
static int32_t msm_host(int device, const void* bases, size_t stride, size_t n_bases, const uint64_t* scalars,
                        size_t n_scalars, uint64_t* out_xyz) {
    if (n_bases != n_scalars) {
        set_error("msm: bases.len() = %zu, scalars.len() = %zu", n_bases, n_scalars);
        return ZKG_ERR_LEN_MISMATCH;
    }
    ZKG_REQUIRE(out_xyz != nullptr, "msm: out is NULL");
    PooledCtx pc;
    ZKG_TRY(pc.acquire(device));
    zkg_ctx* ctx = pc.ctx;
    DeviceGuard dg(ctx->device);
    F* d_out = nullptr;
    ZKG_TRY(msm_host_enqueue<F>(ctx, bases, stride, n_bases, scalars, 0, &d_out));
    ZKG_TRY(copy_d2h(out_xyz, d_out, 3 * sizeof(F), ctx->stream));
    ZKG_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZKG_OK;
}

// ------------------------------------------------------------------------------------------
// ONE MSM over several GPUs of the box from a single call (BASELINE configs[3]): the point range is
// split evenly, device d runs the full host-pointer pipeline on its slice (one host thread per device drives the
// chunked PCIe copies of that slice; the slices cross different PCIe links concurrently), the XYZZ partials travel
// to the first device over NVLink (cudaMemcpyPeerAsync, 128 / 256 B each) behind cross-device events, and one tiny
// kernel adds and normalises them.  MSM is linear, so the result is the single-GPU group element bit for bit.
// ------------------------------------------------------------------------------------------
template <class F>
static int32_t msm_host_sharded(const int32_t* devices, int32_t n_dev, const void* bases, size_t stride, size_t n_bases,
                                const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz) {
    if (n_bases != n_scalars) {
        set_error("msm: bases.len() = %zu, scalars.len() = %zu", n_bases, n_scalars);
        return ZKG_ERR_LEN_MISMATCH;
    }
    ZKG_REQUIRE(devices && n_dev >= 1 && n_dev <= 16 && out_xyz, "msm_sharded: bad device list");
    for (int a = 0; a < n_dev; ++a)
        for (int b = 0; b < a; ++b) ZKG_REQUIRE(devices[a] != devices[b], "msm_sharded: device %d listed twice", devices[a]);
    if (n_dev == 1 || n_bases < (size_t)n_dev * 1024) return msm_host<F>(devices[0], bases, stride, n_bases, scalars, n_scalars, out_xyz);
    std::vector<PooledCtx> pcs(n_dev);
    for (int d = 0; d < n_dev; ++d) ZKG_TRY(pcs[d].acquire(devices[d]));
    std::vector<int32_t> rc(n_dev, ZKG_OK);
    std::vector<std::string> msg(n_dev);
    std::vector<F*> d_part(n_dev, nullptr);
    std::vector<cudaEvent_t> done(n_dev, nullptr);
    auto work = [&](int d) {
        zkg_ctx* ctx = pcs[d].ctx;
        DeviceGuard dg(ctx->device);
        const size_t lo = n_bases * (size_t)d / (size_t)n_dev, hi = n_bases * (size_t)(d + 1) / (size_t)n_dev;
        rc[d] = msm_host_enqueue<F>(ctx, (const uint8_t*)bases + lo * stride, stride, hi - lo, scalars + lo * 4, 1, &d_part[d]);
        if (rc[d] == ZKG_OK) {
            if (cudaEventCreateWithFlags(&done[d], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(done[d], ctx->stream) != cudaSuccess) {
                set_error("msm_sharded: event on device %d failed", ctx->device);
                rc[d] = ZKG_ERR_CUDA;
            }
        }
        if (rc[d] != ZKG_OK) msg[d] = zkg_last_error();          // the message is thread-local: carry it to the caller
    };
    std::vector<std::thread> th;
    for (int d = 1; d < n_dev; ++d) th.emplace_back(work, d);
    work(0);
    for (auto& t : th) t.join();
    int32_t first = ZKG_OK;
    for (int d = 0; d < n_dev && first == ZKG_OK; ++d)
        if (rc[d] != ZKG_OK) { first = rc[d]; set_error("%s", msg[d].c_str()); }
    zkg_ctx* c0 = pcs[0].ctx;
    DeviceGuard dg(c0->device);
    if (first == ZKG_OK) {
        // gather area: a small persistent block of the first device's parameter cache
        void* gather = nullptr; bool fresh = false;
        const char key[] = "msm_sharded_gather";
        first = ctx_cache_get(c0, key, sizeof key, 16 * sizeof(XYZZ<F>) + 3 * sizeof(F), &gather, &fresh);
        if (first == ZKG_OK) {
            XYZZ<F>* g = (XYZZ<F>*)gather;
            F* d_out = (F*)((uint8_t*)gather + 16 * sizeof(XYZZ<F>));
            for (int d = 0; d < n_dev && first == ZKG_OK; ++d) {
                cudaError_t e = cudaStreamWaitEvent(c0->stream, done[d], 0);
                if (e == cudaSuccess)
                    e = d == 0 ? cudaMemcpyAsync(g, d_part[0], sizeof(XYZZ<F>), cudaMemcpyDeviceToDevice, c0->stream)
                               : cudaMemcpyPeerAsync(g + d, c0->device, d_part[d], pcs[d].ctx->device, sizeof(XYZZ<F>), c0->stream);
                if (e != cudaSuccess) { set_error("msm_sharded: gathering the partial of device %d failed: %s", pcs[d].ctx->device, cudaGetErrorString(e)); first = ZKG_ERR_CUDA; }
            }
            if (first == ZKG_OK) {
                k_combine<F><<<1, 32, 0, c0->stream>>>(g, (size_t)n_dev, d_out);
                c0->launches += 1;
                first = copy_d2h(out_xyz, d_out, 3 * sizeof(F), c0->stream);
            }
        }
    }
    // every context is drained before it returns to the pool (also on the error paths)
    for (int d = 0; d < n_dev; ++d) {
        DeviceGuard dgd(pcs[d].ctx->device);
        cudaStreamSynchronize(pcs[d].ctx->stream);
        if (done[d]) cudaEventDestroy(done[d]);
    }
    return first;
}


// per-curve entry points, instantiated in msm_g1.cu (F = Fq) and msm_g2.cu (F = Fq2)
#define ZKG_MSM_DECLARE(G)                                                                                          \
    int32_t msm_run_##G(zkg_ctx* ctx, const void* d_bases, const uint64_t* d_scalars, size_t n, void* d_out, int mode); \
    int32_t pack_bases_##G(zkg_ctx* ctx, const void* d_ark, size_t stride, size_t n, void* d_packed);               \
    int32_t msm_host_##G(int device, const void* bases, size_t stride, size_t n_bases, const uint64_t* scalars,     \
                         size_t n_scalars, uint64_t* out_xyz);                                                      \
    int32_t msm_host_sharded_##G(const int32_t* devices, int32_t n_dev, const void* bases, size_t stride, size_t n_bases, \
                                 const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz);                     \
    int32_t combine_##G(zkg_ctx* ctx, const uint64_t* d_parts, size_t n, uint64_t* d_out);                          \
    int32_t fixed_base_##G(zkg_ctx* ctx, const uint64_t* d_scalars, size_t n, void* d_packed);                      \
    int32_t prepare_##G(zkg_ctx* ctx, const void* d_bases, size_t n, int c, void* d_table);                         \
    int32_t msm_run_prepared_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* d_scalars, size_t n, void* d_out, int mode); \
    int32_t msm_run_prepared_host_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* h_scalars, size_t n, void* d_out, int mode); \
    int32_t crs_det_pack_##G(int device, const void* bases, size_t stride, size_t n, int l, int n_parties,          \
                             const uint32_t* h_scal, void* const* out_by_party, size_t out_stride);
ZKG_MSM_DECLARE(g1)
ZKG_MSM_DECLARE(g2)

#define ZKG_MSM_DEFINE(G, F)                                                                                        \
    int32_t msm_run_##G(zkg_ctx* ctx, const void* d_bases, const uint64_t* d_scalars, size_t n, void* d_out, int mode) { \
        return msm_run<F>(ctx, (const Affine<F>*)d_bases, (const Fr*)d_scalars, n, (F*)d_out, mode);                \
    }                                                                                                               \
    int32_t pack_bases_##G(zkg_ctx* ctx, const void* d_ark, size_t stride, size_t n, void* d_packed) {              \
        return pack_bases<F>(ctx, d_ark, stride, n, d_packed);                                                      \
    }                                                                                                               \
    int32_t msm_host_##G(int device, const void* bases, size_t stride, size_t n_bases, const uint64_t* scalars,     \
                         size_t n_scalars, uint64_t* out_xyz) {                                                     \
        return msm_host<F>(device, bases, stride, n_bases, scalars, n_scalars, out_xyz);                            \
    }                                                                                                               \
    int32_t msm_host_sharded_##G(const int32_t* devices, int32_t n_dev, const void* bases, size_t stride, size_t n_bases, \
                                 const uint64_t* scalars, size_t n_scalars, uint64_t* out_xyz) {                    \
        return msm_host_sharded<F>(devices, n_dev, bases, stride, n_bases, scalars, n_scalars, out_xyz);            \
    }                                                                                                               \
    int32_t combine_##G(zkg_ctx* ctx, const uint64_t* d_parts, size_t n, uint64_t* d_out) {                         \
        k_combine<F><<<1, 32, 0, ctx->stream>>>((const XYZZ<F>*)d_parts, n, (F*)d_out);                             \
        ctx->launches += 1;                                                                                         \
        ZKG_CUDA(cudaGetLastError());                                                                               \
        return ZKG_OK;                                                                                              \
    }                                                                                                               \
    int32_t prepare_##G(zkg_ctx* ctx, const void* d_bases, size_t n, int c, void* d_table) {                        \
        return msm_prepare<F>(ctx, (const Affine<F>*)d_bases, n, c, (Affine<F>*)d_table);                           \
    }                                                                                                               \
    int32_t msm_run_prepared_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* d_scalars, size_t n, void* d_out, int mode) { \
        return msm_run_prepared<F>(ctx, (const Affine<F>*)d_table, c, (const Fr*)d_scalars, n, (F*)d_out, mode);    \
    }                                                                                                               \
    int32_t msm_run_prepared_host_##G(zkg_ctx* ctx, const void* d_table, int c, const uint64_t* h_scalars, size_t n, void* d_out, int mode) { \
        return msm_run_prepared_host<F>(ctx, (const Affine<F>*)d_table, c, h_scalars, n, (F*)d_out, mode);          \
    }                                                                                                               \
    int32_t crs_det_pack_##G(int device, const void* bases, size_t stride, size_t n, int l, int n_parties,          \
                             const uint32_t* h_scal, void* const* out_by_party, size_t out_stride) {                \
        return crs_det_pack_host<F>(device, bases, stride, n, l, n_parties, h_scal, out_by_party, out_stride);      \
    }                                                                                                               \
    int32_t fixed_base_##G(zkg_ctx* ctx, const uint64_t* d_scalars, size_t n, void* d_packed) {                     \
        if (n == 0) return ZKG_OK;                                                                                  \
        k_fixed_base<F><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const Fr*)d_scalars, n, (Affine<F>*)d_packed); \
        ctx->launches += 1;                                                                                         \
        ZKG_CUDA(cudaGetLastError());                                                                               \
        return ZKG_OK;                                                                                              \
    }

}  // namespace zkg
