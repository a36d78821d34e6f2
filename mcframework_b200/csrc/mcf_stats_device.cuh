// mcf_stats_device.cuh -- per-thread building blocks of the StatsEngine kernels.
//
// `__host__ __device__` like mcf_device.cuh, so tests/hostemu can exercise the same logic with g++.
//
// Restated algorithms:
//   moments    : shifted power sums per thread -> (n, mean, M2, M3, M4), merged with the pairwise
//                update formulas of Chan/Golub/LeVeque and Pebay (reference: np.mean / np.std /
//                scipy.stats.skew / kurtosis at stats_engine.py:796,833,902,937)
//   select     : order-preserving u64 keys of doubles + MSD radix digits (np.percentile,
//                stats_engine.py:865,1197,1286)
//   bootstrap  : NumPy's buffered 32-bit Lemire bounded integers (Generator.integers,
//                stats_engine.py:1039-1041).  Every 32-bit slot of the stream is accepted or
//                rejected on its own value, so "the k-th index" == "the k-th accepted slot".
#pragma once
#include "mcf_device.cuh"

namespace mcf {

// ----------------------------------------------------------------------------- moments
struct Moments {
    double n, mean, m2, m3, m4;
};

MCF_HD Moments moments_empty() {
    Moments m;
    m.n = 0.0;
    m.mean = 0.0;
    m.m2 = m.m3 = m.m4 = 0.0;
    return m;
}

// from power sums of d = x - c over n elements
MCF_HD Moments moments_from_shifted(double n, double c, double s1, double s2, double s3, double s4) {
    Moments m = moments_empty();
    if (n <= 0.0) return m;
    const double mu = s1 / n;  // mean of d
    m.n = n;
    m.mean = c + mu;
    m.m2 = s2 - s1 * mu;
    m.m3 = s3 - 3.0 * mu * s2 + 2.0 * mu * mu * s1;
    m.m4 = s4 - 4.0 * mu * s3 + 6.0 * mu * mu * s2 - 3.0 * mu * mu * mu * s1;
    return m;
}

MCF_HD Moments moments_merge(const Moments &a, const Moments &b) {
    if (a.n == 0.0) return b;
    if (b.n == 0.0) return a;
    Moments r;
    const double n = a.n + b.n, d = b.mean - a.mean, dn = d / n, dn2 = dn * dn;
    r.n = n;
    r.mean = a.mean + dn * b.n;
    r.m2 = a.m2 + b.m2 + d * dn * a.n * b.n;
    r.m3 = a.m3 + b.m3 + d * dn2 * a.n * b.n * (a.n - b.n) + 3.0 * dn * (a.n * b.m2 - b.n * a.m2);
    r.m4 = a.m4 + b.m4 + d * dn * dn2 * a.n * b.n * (a.n * a.n - a.n * b.n + b.n * b.n) +
           6.0 * dn2 * (a.n * a.n * b.m2 + b.n * b.n * a.m2) + 4.0 * dn * (a.n * b.m3 - b.n * a.m3);
    return r;
}

// ----------------------------------------------------------------------------- select keys
MCF_HD uint64_t double_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
MCF_HD bool bits_nonfinite(uint64_t b) { return (b & 0x7ff0000000000000ULL) == 0x7ff0000000000000ULL; }
MCF_HD bool bits_nan(uint64_t b) { return bits_nonfinite(b) && (b & 0x000fffffffffffffULL) != 0; }

#define MCF_KEY_EXCLUDED 0xFFFFFFFFFFFFFFFFULL
// Monotone map double -> u64 (total order, -0.0 < +0.0 adjacent).  NaNs (either sign) map to the maximum key,
// which is where np.percentile's sort puts them; with omit_nonfinite also +-inf do.
MCF_HD uint64_t select_key(double d, int omit_nonfinite) {
    const uint64_t b = double_bits(d);
    if (bits_nonfinite(b) && (omit_nonfinite || bits_nan(b))) return MCF_KEY_EXCLUDED;
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
MCF_HD double select_unkey(uint64_t k) {
    const uint64_t b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
    return bits_to_double(b);
}

// ----------------------------------------------------------------------------- bootstrap slots
// Slot s of a stream: the s-th 32-bit value next_uint32() would return.  If the generator holds a
// buffered upper half (has_uint32), that is slot 0; afterwards draw j yields (low, high).
struct SlotRule {
    uint32_t n;          // population size (2 <= n < 2^32)
    uint32_t threshold;  // 2^32 mod n : slot accepted iff low32(u*n) >= threshold
};
MCF_HD SlotRule slot_rule(uint32_t n) {
    SlotRule r;
    r.n = n;
    r.threshold = (uint32_t)((0xFFFFFFFFu - (n - 1u)) % n);
    return r;
}
MCF_HD bool slot_accept(const SlotRule &r, uint32_t u, uint32_t &idx) {
    const uint64_t m = (uint64_t)u * (uint64_t)r.n;
    idx = (uint32_t)(m >> 32);
    return (uint32_t)m >= r.threshold;
}

}  // namespace mcf
