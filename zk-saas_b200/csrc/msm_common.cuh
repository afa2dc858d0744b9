// msm_common.cuh -- window/digit logic shared by the MSM kernels (device) and their host-emulated
// tests.  Restates the signed-digit recoding of ark-ec 0.4.2 `make_digits` (the algorithm behind
// `VariableBaseMSM::msm`, reference call site dist-primitives/src/dmsm/mod.rs:73) with our own
// window-size rule: a group element has one normalised affine form, so the window size changes
// the schedule, not the result.
#pragma once
#include "fp.cuh"

namespace zkg {

static constexpr int MSM_SCALAR_BITS = 254;          // BN254 Fr
static constexpr uint32_t MSM_DIGIT_NONE = 0xffffffffu;

// number of windows such that W*c >= 255: the top window then always has a spare bit, so the
// final carry of the signed recoding is absorbed without an extra window.
ZKG_HD int msm_num_windows(int c) { return MSM_SCALAR_BITS / c + 1; }

// Window size for n points.  Cost model (field multiplications): n*W mixed adds of 10 plus
// W*2^(c-1) buckets reduced with two 14-multiplication adds each.  The TOP window only holds the
// r = 254 - (W-1)c leftover bits, so its points fall into just 2^r buckets; one thread walks one
// bucket, so a small r serialises n/2^r additions on a handful of threads (measured: c = 18 at
// n = 2^22 has r = 2 and takes 7 s instead of 10 ms).  Window sizes whose top buckets would hold
// more than 1024 points are therefore excluded.
ZKG_HD int msm_pick_c(size_t n) {
    int lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    int best = 0, fallback = 5, fallback_used = -1;
    double best_cost = 0;
    for (int c = 5; c <= 20; ++c) {
        if (c > lg + 1 && c > 5) break;
        int W = msm_num_windows(c);
        int r = MSM_SCALAR_BITS - (W - 1) * c;             // bits in the top window (1..c-1)
        int used = r < c - 1 ? r : c - 1;                  // log2 of the top window's populated buckets
        if (used > fallback_used) { fallback = c; fallback_used = used; }
        if ((n >> used) > 1024) continue;
        double cost = (double)n * W * 10.0 + (double)W * (double)((size_t)1 << (c - 1)) * 28.0 + (double)((c - 1 + 2) / 3) * 13e6;
        if (best == 0 || cost < best_cost) { best = c; best_cost = cost; }
    }
    return best ? best : fallback;      // huge n: no window meets the bound, take the best-filled top window
}

// Window size when the bases come with precomputed window shifts 2^(cw) * P (static CRS shares):
// all windows then share ONE set of 2^(c-1) buckets, the bucket reduction shrinks W-fold and the
// Horner over windows disappears, so much larger windows pay off.  The top window's 2^r buckets
// still receive n extra points, hence the same fill constraint.
ZKG_HD int msm_pick_c_merged(size_t n) {
    int lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    int best = 0, fallback = 8, fallback_used = -1;
    double best_cost = 0;
    for (int c = 8; c <= 23; ++c) {
        if (c > lg + 4 && c > 8) break;
        int W = msm_num_windows(c);
        int r = MSM_SCALAR_BITS - (W - 1) * c;
        int used = r < c - 1 ? r : c - 1;
        if (used > fallback_used) { fallback = c; fallback_used = used; }
        if ((n >> used) > 1024) continue;
        // the top window's buckets are walked by one thread each, ~4 us (27.6 k addition units) per point: keep that
        // chain under a quarter of the accumulation (measured: c = 22 at 2^22 points loses 3.5 ms to it)
        {
            double chain = (double)(n >> used) * 27.6e3, quarter = 0.25 * (double)n * W;
            if (chain > (quarter > 2.0e6 ? quarter : 2.0e6)) continue;
        }
        // Calibrated on B200 (tools/scratch/msm_reg.py sweeps), in units of one bucket addition (0.145 ns):
        //   accumulate  n*W, divided by an occupancy factor when there are fewer buckets (= threads) than the
        //               ~75 k the 148 SMs want, plus 15 % for digits + counting sort;
        //   reduction   2.8 per bucket (two full additions) + the latency chain: ~1.7 M for the serial
        //               head and tail, 70 k per log-depth merge level.
        double threads = (double)((size_t)1 << (c - 1));
        double occ = threads >= 75000.0 ? 1.0 : threads / 75000.0;
        double occ_sqrt = 1.0;                       // sqrt(occ) by Newton (no <cmath> in device builds of this header)
        if (occ < 1.0) { occ_sqrt = 0.5 * (1.0 + occ); for (int it = 0; it < 6; ++it) occ_sqrt = 0.5 * (occ_sqrt + occ / occ_sqrt); }
        double cost = (double)n * W * (1.0 / occ_sqrt + 0.15) + threads * 2.8 + 1.7e6 + 7.0e4 * (c - 4);
        if (best == 0 || cost < best_cost) { best = c; best_cost = cost; }
    }
    return best ? best : fallback;
}

// c-bit window `w` of a canonical 256-bit scalar held as 8 x u32
ZKG_HD uint32_t msm_window_bits(const uint32_t* s, int w, int c) {
    int bit = w * c;
    int limb = bit >> 5, off = bit & 31;
    uint64_t lo = limb < 8 ? s[limb] : 0u;
    uint64_t hi = limb + 1 < 8 ? s[limb + 1] : 0u;
    uint64_t v = (lo | (hi << 32)) >> off;
    return (uint32_t)(v & (((uint64_t)1 << c) - 1));
}

// Signed radix-2^c digits d_w in [-2^(c-1), 2^(c-1)] with sum d_w 2^(cw) = s.
// Encoding: MSM_DIGIT_NONE for 0, else (|d|-1) | (d<0 ? 1<<31 : 0); |d|-1 is the bucket index.
ZKG_HD void msm_signed_digits(const uint32_t* s, int c, int W, uint32_t* out) {
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    for (int w = 0; w < W; ++w) {
        uint32_t coef = msm_window_bits(s, w, c) + carry;
        uint32_t neg = 0;
        carry = 0;
        if (w != W - 1 && coef >= half && coef != 0) {
            // coef in [2^(c-1), 2^c]: use coef - 2^c (<= 0) and carry one into the next window
            if (coef == (1u << c)) { coef = 0; }
            else { coef = (1u << c) - coef; neg = 1; }
            carry = 1;
        }
        out[w] = coef == 0 ? MSM_DIGIT_NONE : ((coef - 1) | (neg << 31));
    }
}

}  // namespace zkg
