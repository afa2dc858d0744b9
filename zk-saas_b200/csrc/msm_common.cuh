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

// Bucket populations from which a bucket counts as HEAVY (split over a block instead of walked by one thread): well
// above anything a uniform digit distribution produces, low enough to catch the under-filled top window of small MSMs.
ZKG_HD uint32_t msm_heavy_key(double mean) {
    double k = 8.0 * mean + 64.0;
    return k < 128.0 ? 128u : k > 2047.0 ? 2047u : (uint32_t)k;
}

// Window size for n points on the per-window path (no prepared table).  Cost model in MICROSECONDS, calibrated on B200
// window sweeps (tools/scratch/sweep_gen.py, round 2; it reproduces the measured totals within ~10 % from 2^10 to 2^22):
//   sort        50 + 0.02 ns per (point, window)
//   accumulate  one thread walks one bucket: (n / buckets) additions at max(7 us, 13 us * threads / 95 k) each -- the
//               latency of a lone thread, or the share of a full machine; the TOP window only holds r = 254 - (W-1)c
//               bits, so its 2^r buckets get n / 2^r points each: a serial chain unless they reach the heavy-bucket
//               threshold and are split over blocks (k_accumulate_heavy: ~110 us of tree and lock per bucket)
//   reduction   level 0 (two additions per bucket, >= 165 us) or none for small bucket sets, ~8 us per butterfly
//               level, 70 us for the bit scaling, and the Horner over windows: c (W-1) DEPENDENT doublings at ~2 us
// (round 1's model counted multiplications only and ignored both latency terms: it chose c = 8 at n = 2^16, where 64
//  top-window buckets of 1024 points each made the MSM take 7.5 ms; c = 15 takes 1.15 ms.)
// g2: the same with the measured Fq2 ratios.
ZKG_HD int msm_pick_c(size_t n, bool g2 = false) {
    int lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    const double lat_min = g2 ? 15.0 : 7.0, lat_full = g2 ? 29.0 : 13.0, thr_full = g2 ? 56800.0 : 95000.0;
    const double add_ns = g2 ? 0.75 : 0.25, lvl_us = g2 ? 16.0 : 8.0, bits_us = g2 ? 190.0 : 70.0, dbl_us = g2 ? 6.5 : 2.0;
    int best = 5;
    double best_cost = -1.0;
    for (int c = 4; c <= 20; ++c) {
        if (c > lg + 2 && c > 4) break;
        const int W = msm_num_windows(c);
        const int r = MSM_SCALAR_BITS - (W - 1) * c;             // bits in the top window (1..c)
        const int used = r < c - 1 ? r : c - 1;                  // log2 of the top window's populated buckets
        const double nb = (double)((size_t)1 << (c - 1)), slots = nb * W, mean = (double)n / nb;
        const double lat = lat_full * slots / thr_full > lat_min ? lat_full * slots / thr_full : lat_min;
        double acc = mean * lat;
        const double top = (double)n / (double)((size_t)1 << used);
        if (top >= (double)msm_heavy_key(mean)) {
            double parts = top / 4096.0;
            parts = parts < 1.0 ? 1.0 : parts > 32.0 ? 32.0 : parts;
            double hb = (double)((size_t)1 << used);                 // heavy buckets; the launch walks them 148 at a time
            double rounds = hb / 148.0;
            rounds = rounds < 1.0 ? 1.0 : rounds;
            const double hthreads = (hb < 148.0 ? hb : 148.0) * parts * 128.0;
            const double lat_h = lat_full * hthreads / thr_full > lat_min ? lat_full * hthreads / thr_full : lat_min;
            acc += rounds * (top / (parts * 128.0) * lat_h + (g2 ? 300.0 : 110.0));
        } else if (top * lat > acc) acc = top * lat;
        const double sort = 50.0 + (double)n * W * 0.02e-3;
        double red;
        if (slots <= 16384.0) red = (c - 1) * lvl_us;
        else {
            double l0 = slots * 2.0 * add_ns * 1e-3;
            red = (l0 > 165.0 ? l0 : 165.0) + (c - 4) * lvl_us + slots / 8.0 * 3.0 * add_ns * 1e-3;
        }
        red += bits_us + (double)c * (W - 1) * dbl_us + W * 1.5 + 60.0;
        const double cost = sort + acc + red;
        if (best_cost < 0 || cost < best_cost) { best = c; best_cost = cost; }
    }
    return best;
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
