// mcf_device.cuh -- per-thread building blocks of the B200 replication kernels.
//
// Everything here is `__host__ __device__` so that tests/hostemu can compile the SAME source
// with g++ and check the stream logic on a CPU-only box (the product never runs that build).
// The kernels in mcf_kernels.cu are thin grid/block wrappers around these functions.
//
// Algorithms restated (NumPy 2.3.x, the reference's pinned dependency):
//   PCG64 XSL-RR 128/64      numpy/random/src/pcg64/pcg64.h      (reference: core.py:315,364)
//   Philox4x64-10            numpy/random/src/philox/philox.h    (reference: core.py:102-103,659)
//   next_double              (u64 >> 11) * 2^-53
//   random_standard_normal   256-layer ziggurat, distributions.c (reference: black_scholes.py:103,239)
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#include "../../include/mcf_b200.h"

#if defined(__CUDACC__)
#define MCF_HD __host__ __device__ __forceinline__
#else
#define MCF_HD inline
#endif

namespace mcf {

// ---------------------------------------------------------------------------------------
// exact-rounding helpers: the reference's results come from separately rounded mul/add
// (no FMA in NumPy's x86-64 wheels), so nothing on the parity path may be contracted.
// ---------------------------------------------------------------------------------------
MCF_HD double xmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;  // host build also uses -ffp-contract=off
    return r;
#endif
}
MCF_HD double xadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
MCF_HD double xsub(double a, double b) { return xadd(a, -b); }

MCF_HD double bits_to_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}

// exact conversion of an integer < 2^52 (one FP64 add instead of a 64-bit I2F)
MCF_HD double u52_to_double(uint64_t v) { return bits_to_double(0x4330000000000000ULL | v) - 4503599627370496.0; }

// next_double: 53-bit integer * 2^-53 (exact)
MCF_HD double u64_to_unit_double(uint64_t v) {
    uint64_t k = v >> 11;  // < 2^53
    // split so both halves fit the 2^52 trick; every step is exact
    double hi = u52_to_double(k >> 32);          // < 2^21
    double lo = u52_to_double(k & 0xFFFFFFFFULL);  // < 2^32
    return (hi * 4294967296.0 + lo) * (1.0 / 9007199254740992.0);
}

MCF_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// Full 64x64 -> 128 product with a compile-time-constant multiplier (the Philox round multipliers).
// CHAIN = false: mul.hi.u64 + mul.lo.u64 as ptxas expands them -- four IMAD.WIDE.U32 plus IMAD.X, IMAD.MOV (both on the
// multiplier pipe, the busiest one of the RNG kernels) and one IADD3.
// CHAIN = true (device only): the same four wide multiplications written as a carry chain, so that the carry of the
// middle sum is materialised by SEL on the integer ALU instead of IMAD.X: B = xl*ml, A = xh*ml, C = xl*mh + A (carry cy),
// w1 = B.hi + C.lo (carry k), (w3:w2) = xh*mh + (cy:C.hi) + k.  Measured in the planning scan: 204 -> 192 ms per two
// plans of 1e8 x 252.  Kernels at the register limit with more live state (the 8-set replay) spill with it and keep
// CHAIN = false.
template <bool CHAIN = false>
MCF_HD void mul64_full(uint64_t m, uint64_t x, uint64_t &hi, uint64_t &lo) {
#if defined(__CUDA_ARCH__)
    if (CHAIN) {
        const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32), ml = (uint32_t)m, mh = (uint32_t)(m >> 32);
        uint32_t w0, w1, w2, w3;
        asm("{\n\t"
            ".reg .u32 b1, a0, a1, c0, c1, cy;\n\t"
            ".reg .u64 tb, ta;\n\t"
            "mul.wide.u32 tb, %4, %6;\n\t"
            "mul.wide.u32 ta, %5, %6;\n\t"
            "mov.b64 {%0, b1}, tb;\n\t"
            "mov.b64 {a0, a1}, ta;\n\t"
            "mad.lo.cc.u32 c0, %4, %7, a0;\n\t"
            "madc.hi.cc.u32 c1, %4, %7, a1;\n\t"
            "addc.u32 cy, 0, 0;\n\t"
            "add.cc.u32 %1, b1, c0;\n\t"
            "madc.lo.cc.u32 %2, %5, %7, c1;\n\t"
            "madc.hi.u32 %3, %5, %7, cy;\n\t"
            "}"
            : "=&r"(w0), "=&r"(w1), "=&r"(w2), "=&r"(w3)
            : "r"(xl), "r"(xh), "r"(ml), "r"(mh));
        lo = ((uint64_t)w1 << 32) | w0;
        hi = ((uint64_t)w3 << 32) | w2;
        return;
    }
#endif
    hi = mulhi64(m, x);
    lo = m * x;
}

MCF_HD int popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
MCF_HD int ctz64(uint64_t v) {  // v != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
// position of the k-th (0-based) set bit of m; requires k < popc(m)
MCF_HD int select_bit64(uint64_t m, int k) {
    int pos = 0;
#pragma unroll
    for (int w = 32; w >= 1; w >>= 1) {
        uint64_t low = m & ((1ULL << w) - 1ULL);
        int c = popc64(low);
        if (k >= c) {
            k -= c;
            m >>= w;
            pos += w;
        } else {
            m = low;
        }
    }
    return pos;
}
MCF_HD uint64_t mask_from(unsigned e) { return e >= 64 ? 0ULL : ~((1ULL << e) - 1ULL); }  // bits >= e

// ---------------------------------------------------------------------------------------
// 128-bit helpers for the PCG64 LCG
// ---------------------------------------------------------------------------------------
struct u128 {
    uint64_t hi, lo;
};
MCF_HD u128 mk128(uint64_t hi, uint64_t lo) {
    u128 r;
    r.hi = hi;
    r.lo = lo;
    return r;
}
MCF_HD u128 mul128(u128 a, u128 b) {  // low 128 bits of a*b
    u128 r;
    r.lo = a.lo * b.lo;
    r.hi = mulhi64(a.lo, b.lo) + a.lo * b.hi + a.hi * b.lo;
    return r;
}
MCF_HD u128 add128(u128 a, u128 b) {
    u128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ULL : 0ULL);
    return r;
}

#define MCF_PCG_MULT_HI 0x2360ED051FC65DA4ULL
#define MCF_PCG_MULT_LO 0x4385DF649FCCF645ULL

MCF_HD u128 pcg_step(u128 s, u128 inc) { return add128(mul128(s, mk128(MCF_PCG_MULT_HI, MCF_PCG_MULT_LO)), inc); }
MCF_HD uint64_t pcg_output(u128 s) {
    uint64_t v = s.hi ^ s.lo;
    unsigned rot = (unsigned)(s.hi >> 58);
    return (v >> rot) | (v << ((64u - rot) & 63u));
}
// state after `delta` steps (pcg_advance_lcg_128, O(log delta))
MCF_HD u128 pcg_advance(u128 s, u128 inc, uint64_t delta) {
    u128 cur_mult = mk128(MCF_PCG_MULT_HI, MCF_PCG_MULT_LO), cur_plus = inc;
    u128 acc_mult = mk128(0, 1), acc_plus = mk128(0, 0);
    while (delta > 0) {
        if (delta & 1ULL) {
            acc_mult = mul128(acc_mult, cur_mult);
            acc_plus = add128(mul128(acc_plus, cur_mult), cur_plus);
        }
        cur_plus = mul128(add128(cur_mult, mk128(0, 1)), cur_plus);
        cur_mult = mul128(cur_mult, cur_mult);
        delta >>= 1;
    }
    return add128(mul128(acc_mult, s), acc_plus);
}

// ---------------------------------------------------------------------------------------
// Philox4x64-10
// ---------------------------------------------------------------------------------------
#define MCF_PHILOX_M0 0xD2E7470EE14C6C93ULL
#define MCF_PHILOX_M1 0xCA5A826395121157ULL
#define MCF_PHILOX_W0 0x9E3779B97F4A7C15ULL
#define MCF_PHILOX_W1 0xBB67AE8584CAA73BULL

MCF_HD void philox4x64_10(uint64_t c0, uint64_t c1, uint64_t c2, uint64_t c3, uint64_t k0, uint64_t k1,
                          uint64_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t hi0, lo0, hi1, lo1;
        mul64_full(MCF_PHILOX_M0, c0, hi0, lo0);
        mul64_full(MCF_PHILOX_M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += MCF_PHILOX_W0;
        k1 += MCF_PHILOX_W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// The same block function with the ten round keys precomputed (k + r*W): when the key is a launch constant (one
// launch per stream, keys in the kernel-parameter constant bank) the key schedule costs no instruction at all.
struct PhiloxKeys {
    uint64_t rk[20];          // rk[2r], rk[2r+1]: keys of round r
    uint64_t b0, b1, b2, b3;  // base counter of the stream
    uint64_t off;             // buffer offset of the stream (buf_pos; multiple of 4 for the fast paths)
    // Rounds 0 and 1 with the constant halves folded (see philox_block_folded): three of the four counter words of a
    // stream never change, so one of the two multiplications of round 0 and of round 1 has launch-constant operands.
    uint64_t f0_c2x;  // b3 ^ rk[1]:                              round-0 c2' = mulhi(M0, c0) ^ f0_c2x
    uint64_t f1_c0x;  // M1*b2 ^ rk[2]:                           round-1 c0'' = mulhi(M1, c2') ^ f1_c0x
    uint64_t f1_c2x;  // mulhi(M0, A0) ^ rk[3]:                   round-1 c2'' = c3' ^ f1_c2x
    uint64_t f1_c3;   // M0 * A0,  A0 = mulhi(M1, b2) ^ b1 ^ rk[0]: round-1 c3''
    uint32_t fold_ok;  // b0 + block never carries into b1 for any block this stream can reach
};
MCF_HD PhiloxKeys philox_keys_of(const mcf_stream &s) {
    PhiloxKeys K;
    uint64_t k0 = s.s[0], k1 = s.s[1];
    for (int r = 0; r < 10; r++) {
        K.rk[2 * r] = k0;
        K.rk[2 * r + 1] = k1;
        k0 += MCF_PHILOX_W0;
        k1 += MCF_PHILOX_W1;
    }
    K.b0 = s.s[2];
    K.b1 = s.s[3];
    K.b2 = s.s[4];
    K.b3 = s.s[5];
    K.off = (uint64_t)s.buf_pos;
    const uint64_t A0 = mulhi64(MCF_PHILOX_M1, K.b2) ^ K.b1 ^ K.rk[0];
    K.f0_c2x = K.b3 ^ K.rk[1];
    K.f1_c0x = (MCF_PHILOX_M1 * K.b2) ^ K.rk[2];
    K.f1_c2x = mulhi64(MCF_PHILOX_M0, A0) ^ K.rk[3];
    K.f1_c3 = MCF_PHILOX_M0 * A0;
    K.fold_ok = K.b0 < (1ULL << 62) ? 1u : 0u;  // block indices stay far below 2^62
    return K;
}
MCF_HD void philox_block_keys(const PhiloxKeys &K, uint64_t blk, uint64_t &o0, uint64_t &o1, uint64_t &o2, uint64_t &o3) {
    uint64_t c0 = K.b0 + blk;
    const uint64_t carry = c0 < K.b0 ? 1ULL : 0ULL;
    uint64_t c1 = K.b1 + carry;
    uint64_t c2 = K.b2 + ((carry && c1 == 0) ? 1ULL : 0ULL);
    uint64_t c3 = K.b3;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t hi0, lo0, hi1, lo1;
        mul64_full(MCF_PHILOX_M0, c0, hi0, lo0);
        mul64_full(MCF_PHILOX_M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ K.rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ K.rk[2 * r + 1];
        c3 = lo0;
    }
    o0 = c0;
    o1 = c1;
    o2 = c2;
    o3 = c3;
}

// Block `blk` of a stream whose low counter word never carries (K.fold_ok): 18 wide multiplications instead of 20.
//   round 0: (x, b1, b2, b3)      -> (A0,              M1*b2,      mulhi(M0,x) ^ b3 ^ rk1, M0*x)     A0 constant
//   round 1: (A0, M1*b2, c2, c3)  -> (mulhi(M1,c2) ^ M1*b2 ^ rk2, M1*c2, mulhi(M0,A0) ^ c3 ^ rk3, M0*A0)
// `r0hi:r0lo` = M0 * (b0 + blk), the one product of round 0 that depends on the block index.  A thread that walks
// consecutive blocks keeps it as a running 128-bit value (philox_r0_next: + M0 per block) instead of multiplying.
template <bool CHAIN = false>
MCF_HD void philox_block_folded_r0(const PhiloxKeys &K, uint64_t r0hi, uint64_t r0lo, uint64_t &o0, uint64_t &o1, uint64_t &o2,
                                   uint64_t &o3) {
    uint64_t c0, c1, c2 = r0hi ^ K.f0_c2x, c3 = r0lo;
    mul64_full<CHAIN>(MCF_PHILOX_M1, c2, c0, c1);  // round 1
    c0 ^= K.f1_c0x;
    c2 = c3 ^ K.f1_c2x;
    c3 = K.f1_c3;
#pragma unroll
    for (int r = 2; r < 10; r++) {
        uint64_t hi0, lo0, hi1, lo1;
        mul64_full<CHAIN>(MCF_PHILOX_M0, c0, hi0, lo0);
        mul64_full<CHAIN>(MCF_PHILOX_M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ K.rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ K.rk[2 * r + 1];
        c3 = lo0;
    }
    o0 = c0;
    o1 = c1;
    o2 = c2;
    o3 = c3;
}
MCF_HD void philox_r0_next(uint64_t &r0hi, uint64_t &r0lo) {  // (hi:lo) += M0
#if defined(__CUDA_ARCH__)
    asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, 0;" : "+l"(r0lo), "+l"(r0hi) : "l"(MCF_PHILOX_M0));
#else
    r0lo += MCF_PHILOX_M0;
    r0hi += r0lo < MCF_PHILOX_M0 ? 1ULL : 0ULL;
#endif
}

template <bool CHAIN = false>
MCF_HD void philox_block_folded(const PhiloxKeys &K, uint64_t blk, uint64_t &o0, uint64_t &o1, uint64_t &o2, uint64_t &o3) {
    const uint64_t x = K.b0 + blk;
    uint64_t c0, c1, c2, c3;
    mul64_full<CHAIN>(MCF_PHILOX_M0, x, c2, c3);  // round 0
    c2 ^= K.f0_c2x;
    mul64_full<CHAIN>(MCF_PHILOX_M1, c2, c0, c1);  // round 1
    c0 ^= K.f1_c0x;
    c2 = c3 ^ K.f1_c2x;
    c3 = K.f1_c3;
#pragma unroll
    for (int r = 2; r < 10; r++) {
        uint64_t hi0, lo0, hi1, lo1;
        mul64_full<CHAIN>(MCF_PHILOX_M0, c0, hi0, lo0);
        mul64_full<CHAIN>(MCF_PHILOX_M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ K.rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ K.rk[2 * r + 1];
        c3 = lo0;
    }
    o0 = c0;
    o1 = c1;
    o2 = c2;
    o3 = c3;
}

// ---------------------------------------------------------------------------------------
// Cursors: random access into a reference stream.  `pos` counts 64-bit draws from the
// state captured in the mcf_stream record (pos 0 = the next value NumPy would return).
// ---------------------------------------------------------------------------------------
template <int KIND>
struct Cursor;

template <>
struct Cursor<MCF_GEN_PCG64> {
    u128 state, inc;
    uint64_t pos;
    MCF_HD void init(const mcf_stream &s, uint64_t p) {
        inc = mk128(s.s[2], s.s[3]);
        state = pcg_advance(mk128(s.s[0], s.s[1]), inc, p);
        pos = p;
    }
    MCF_HD uint64_t next64() {
        state = pcg_step(state, inc);  // NumPy steps first, then outputs from the new state
        pos++;
        return pcg_output(state);
    }
};

template <>
struct Cursor<MCF_GEN_PHILOX> {
    uint64_t k0, k1, b0, b1, b2, b3;  // key and base counter
    uint64_t q;                       // buf_pos + pos: block = base + (q >> 2), lane = q & 3
    uint64_t pos;
    uint64_t v0, v1, v2, v3;  // current block
    uint64_t have;            // block index held in v*, ~0 = none
    MCF_HD void init(const mcf_stream &s, uint64_t p) {
        k0 = s.s[0];
        k1 = s.s[1];
        b0 = s.s[2];
        b1 = s.s[3];
        b2 = s.s[4];
        b3 = s.s[5];
        q = (uint64_t)s.buf_pos + p;
        pos = p;
        have = ~0ULL;
        v0 = v1 = v2 = v3 = 0;
    }
    MCF_HD void load_block(uint64_t blk) {
        uint64_t c0 = b0 + blk;
        uint64_t carry = c0 < b0 ? 1ULL : 0ULL;
        uint64_t c1 = b1 + carry;
        uint64_t c2 = b2 + ((carry && c1 == 0) ? 1ULL : 0ULL);
        uint64_t out[4];
        philox4x64_10(c0, c1, c2, b3, k0, k1, out);
        v0 = out[0];
        v1 = out[1];
        v2 = out[2];
        v3 = out[3];
        have = blk;
    }
    MCF_HD uint64_t next64() {
        uint64_t blk = q >> 2;
        if (blk != have) load_block(blk);
        unsigned lane = (unsigned)(q & 3ULL);
        uint64_t v = lane == 0 ? v0 : (lane == 1 ? v1 : (lane == 2 ? v2 : v3));
        q++;
        pos++;
        return v;
    }
};

// ---------------------------------------------------------------------------------------
// Draw delivery.  `consume_draws(cur, f)` hands consecutive 64-bit draws to `f` until it returns false.
// For Philox the draws come in blocks of four that are generated ONCE per loop trip by every lane of the
// warp together: lanes whose ziggurat chains have drifted apart still execute the (expensive) block function
// converged, instead of each lane regenerating blocks at its own iteration (4x divergence waste).
// ---------------------------------------------------------------------------------------
template <class F>
MCF_HD void consume_draws(Cursor<MCF_GEN_PCG64> &cur, F f) {
    while (f(cur.next64())) {
    }
}

template <class F>
MCF_HD void consume_draws(Cursor<MCF_GEN_PHILOX> &cur, F f) {
    // first block: the stream position may sit anywhere inside it
    {
        const unsigned first = (unsigned)(cur.q & 3ULL);
        cur.load_block(cur.q >> 2);
        unsigned n = 0;
        bool go = true;
        if (first == 0) {
            n++;
            go = f(cur.v0);
        }
        if (go && first <= 1) {
            n++;
            go = f(cur.v1);
        }
        if (go && first <= 2) {
            n++;
            go = f(cur.v2);
        }
        if (go) {
            n++;
            go = f(cur.v3);
        }
        cur.q += n;
        cur.pos += n;
        if (!go) return;
    }
    // whole blocks: position bookkeeping once per block, not per draw
    for (;;) {
        cur.load_block(cur.q >> 2);
        unsigned n = 1;
        bool go = f(cur.v0);
        if (go) {
            n = 2;
            go = f(cur.v1);
        }
        if (go) {
            n = 3;
            go = f(cur.v2);
        }
        if (go) {
            n = 4;
            go = f(cur.v3);
        }
        cur.q += n;
        cur.pos += n;
        if (!go) return;
    }
}

// ---------------------------------------------------------------------------------------
// Ziggurat
// ---------------------------------------------------------------------------------------
struct ZigKW {  // ki and wi side by side: one 16-byte shared-memory load per draw in the fast walk
    uint64_t ki;
    double wi;
};
struct ZigTables {
    const uint64_t *ki;
    const double *wi;
    const double *fi;
    // 512 entries indexed by the low 9 bits of a draw (layer, sign): {ki[layer] | 2^52-exponent bits, sign ? -wi : wi}.
    // `rabs < ki` becomes a compare of the biased words and the signed normal is one add and one multiply.
    const ZigKW *kws;
};
#define MCF_U52_BIAS 0x4330000000000000ULL
// The scan reads the 52 bits of `rabs` as the bit pattern of a SUBNORMAL double d = rabs * 2^-1074 (no exponent bias to
// OR in) and keeps the table weight as wi' = wi * 2^974, so d * wi' = x * 2^-100 -- a normal number, rounded exactly like
// rabs * wi -- and the sub-chunk sums are accumulated with ONE fused multiply-add per fast draw in that scaled range
// (rescaled by 2^100, exactly, after the loop): two FP64 operations per draw (compare + FMA) instead of four.
// Measured: 97.4 -> 93.8 ms per plan of 1e8 x 252 (profiles/r2_scan_variants.jsonl).
#define MCF_SCAN_BIAS 0ULL
#define MCF_SCAN_WI_SCALE 0x1p974
#define MCF_SCAN_VALUE(d, wi) xmul(xmul((d), (wi)), 0x1p100)
MCF_HD void zig_fill_kws(ZigKW *kws, unsigned e, uint64_t ki, double wi) {
    kws[e].ki = ki | MCF_SCAN_BIAS;
    kws[e].wi = ((e & 0x100u) ? -wi : wi) * MCF_SCAN_WI_SCALE;
}
#define MCF_ZIG_R 3.6541528853610088
#define MCF_ZIG_INV_R 0.27366123732975828

// NumPy's random_standard_normal as a state machine over single draws (distributions.c):
//   state 0: the draw starts an attempt (1 draw when rabs < ki[idx]: 98.5 %)
//   state 1: wedge -- the draw is the uniform of the acceptance test; accepted or not, back to state 0
//   state 2/3: tail (idx == 0) -- draws come in (xx, yy) pairs until 2*yy > xx*xx
struct ZigFsm {
    uint32_t state, idx;
    uint64_t rabs;
    double x, xx;
};
MCF_HD void zig_reset(ZigFsm &s) {
    s.state = 0;
    s.idx = 0;
    s.rabs = 0;
    s.x = 0.0;
    s.xx = 0.0;
}

// Feed one draw.  Returns true when a normal is produced (in `out`).
MCF_HD bool zig_feed(ZigFsm &s, uint64_t r, const ZigTables &t, double &out) {
    if (s.state == 0) {
        const unsigned idx = (unsigned)(r & 0xffULL);
        r >>= 8;
        const unsigned sign = (unsigned)(r & 1ULL);
        const uint64_t rabs = (r >> 1) & 0x000fffffffffffffULL;
        const double v = xmul(u52_to_double(rabs), t.wi[idx]);
        const double x = sign ? -v : v;
        if (rabs < t.ki[idx]) {
            out = x;
            return true;
        }
        s.idx = idx;
        s.rabs = rabs;
        s.x = x;
        s.state = idx == 0 ? 2u : 1u;
        return false;
    }
    if (s.state == 1) {
        s.state = 0;
        const double f1 = t.fi[s.idx - 1], f0 = t.fi[s.idx];
        const double lhs = xadd(xmul(xsub(f1, f0), u64_to_unit_double(r)), f0);
        // NumPy accepts iff lhs < exp(-x*x/2).  Inside layer idx, exp(-x*x/2) = f0*exp(d) with
        // d = (xb^2 - x^2)/2 in [0, 0.73] (xb = right edge of the layer, f0 = f(xb)), and
        //   1 + d + d^2/2 + d^3/6 + d^4/24  <=  exp(d)  <=  that + d^5/57      (0 <= d <= 0.75).
        // The two certified bounds (with a 1e-9 relative guard band, 7 orders above the rounding of either side)
        // settle 99.998 % of the wedge tests; only the sliver in between evaluates exp().  The DECISION is NumPy's.
        const double xb = t.wi[s.idx] * 4503599627370496.0, ax = fabs(s.x);
        const double d = 0.5 * (xb - ax) * (xb + ax);
        const double pl = 1.0 + d * (1.0 + d * (0.5 + d * (1.0 / 6.0 + d * (1.0 / 24.0))));
        const double d2 = d * d;
        const double ph = pl + d2 * d2 * d * (1.0 / 57.0);
        bool accept;
        if (d >= 0.0 && d <= 0.75 && lhs < f0 * pl * (1.0 - 1e-9))
            accept = true;
        else if (d >= 0.0 && d <= 0.75 && lhs > f0 * ph * (1.0 + 1e-9))
            accept = false;
        else
            accept = lhs < exp(xmul(xmul(-0.5, s.x), s.x));
        if (accept) {
            out = s.x;
            return true;
        }
        return false;
    }
    if (s.state == 2) {
        s.xx = xmul(-MCF_ZIG_INV_R, log1p(-u64_to_unit_double(r)));
        s.state = 3;
        return false;
    }
    const double yy = -log1p(-u64_to_unit_double(r));
    if (xadd(yy, yy) > xmul(s.xx, s.xx)) {
        const double m = xadd(MCF_ZIG_R, s.xx);
        out = ((s.rabs >> 8) & 1ULL) ? -m : m;
        s.state = 0;
        return true;
    }
    s.state = 2;
    return false;
}

// the next `n` standard normals of the stream, in order: f(z) for each
template <class Cur, class F>
MCF_HD void for_each_normal(Cur &cur, const ZigTables &t, int64_t n, F f) {
    if (n <= 0) return;
    ZigFsm s;
    zig_reset(s);
    int64_t left = n;
    consume_draws(cur, [&](uint64_t r) {
        double z;
        if (zig_feed(s, r, t, z)) {
            f(z);
            left--;
        }
        return left > 0;
    });
}

// One pass through NumPy's `for (;;)` body (cursor-driven form; tests and single draws).
template <class Cur>
MCF_HD bool zig_attempt(Cur &cur, const ZigTables &t, double &x) {
    ZigFsm s;
    zig_reset(s);
    bool emitted = false;
    do {
        emitted = zig_feed(s, cur.next64(), t, x);
    } while (s.state != 0);
    return emitted;
}

template <class Cur>
MCF_HD double zig_normal(Cur &cur, const ZigTables &t) {
    double x;
    while (!zig_attempt(cur, t, x)) {
    }
    return x;
}

// ---------------------------------------------------------------------------------------
// Stream lookup: replication index -> stream record (streams sorted by first_sim)
// ---------------------------------------------------------------------------------------
MCF_HD int find_stream(const mcf_stream *streams, int n_streams, int64_t sim) {
    int lo = 0, hi = n_streams - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (streams[mid].first_sim <= sim)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------
// Pi: hits among `npts` consecutive points starting at the cursor (sims/pi.py:71-72)
// ---------------------------------------------------------------------------------------
// x = -1 + 2*((v >> 11) * 2^-53) = k*2^-52 - 1 with k = v >> 11: exactly representable, so any exact evaluation equals
// NumPy's two roundings.  With k = m + top*2^52 (m the low 52 bits, top the 53rd): x = (k - 2^52)*2^-52 = d*2^-52 - (top ? 1 : 2) where
// d = 2^52 + m is built from the bits -- one logic operation and ONE exact FMA per coordinate (the product lies in [1, 2),
// the difference is a multiple of 2^-52 below 1 in magnitude, so nothing is rounded anywhere).
MCF_HD double pi_coord(uint64_t v) {
    const double d = bits_to_double(0x4330000000000000ULL | ((v >> 11) & 0x000fffffffffffffULL));
    return fma(d, 1.0 / 4503599627370496.0, (v >> 63) ? -1.0 : -2.0);
}
template <class Cur>
MCF_HD int64_t pi_count_hits(Cur &cur, int64_t npts) {
    if (npts <= 0) return 0;
    int64_t hits = 0, left = 2 * npts;
    double x = 0.0;
    consume_draws(cur, [&](uint64_t r) {
        const double c = pi_coord(r);  // == -1 + 2*next_double exactly
        if (left & 1) {  // second coordinate of the point
            hits += (xadd(xmul(x, x), xmul(c, c)) <= 1.0) ? 1 : 0;
        } else {
            x = c;
        }
        left--;
        return left > 0;
    });
    return hits;
}

// Block form for block-aligned Philox streams whose lanes start on block boundaries: two points per Philox block,
// no per-draw bookkeeping; one coordinate costs a shift, one logic operation and one FMA (pi_coord).
MCF_HD int pi_inside(uint64_t vx, uint64_t vy) {
    const double x = pi_coord(vx), y = pi_coord(vy);
    return (xadd(xmul(x, x), xmul(y, y)) <= 1.0) ? 1 : 0;
}
// hits among `npts` points whose first draw is the first draw of block `blk0`
template <class Load>
MCF_HD int64_t pi_count_hits_blocks(Load load, uint64_t blk0, int64_t npts) {
    int64_t hits = 0;
    uint64_t blk = blk0, a, b, c, d;
    for (; npts >= 2; npts -= 2, blk++) {
        load(blk, a, b, c, d);
        hits += pi_inside(a, b) + pi_inside(c, d);
    }
    if (npts > 0) {
        load(blk, a, b, c, d);
        hits += pi_inside(a, b);
    }
    return hits;
}

// ---------------------------------------------------------------------------------------
// Planning pass over a stream window.  A chunk is 64 consecutive draw positions.
//   startmask bit j : a ziggurat attempt starts at chunk_base + j
//   emitmask  bit j : ... and that attempt produces a normal
//   exit            : first attempt start at or beyond the chunk end, minus the chunk end
//   gen_entry       : offset the masks were generated from (0 = speculative)
// ---------------------------------------------------------------------------------------
#define MCF_CHUNK 64
#define MCF_SUBS 4       // sub-chunks per chunk (for the stored partial sums)
#define MCF_SUB_SHIFT 4  // log2(draws per sub-chunk)
#define MCF_EXIT_OVERFLOW 255
#define MCF_GEN_ENTRY_INVALID 254  // "generated with no usable entry": forces the general re-walk

struct ChunkInfo {
    uint64_t startmask, emitmask;
    uint32_t exit, gen_entry;
};

// cursor must be positioned at chunk_base + entry; zsub_out[MCF_SUBS] receives the sub-chunk sums
template <class Cur>
MCF_HD ChunkInfo scan_chunk(Cur &cur, uint64_t chunk_base, uint32_t entry, const ZigTables &t, double *zsub_out) {
    ChunkInfo ci;
    ci.startmask = 0;
    ci.emitmask = 0;
    ci.gen_entry = entry;
    for (int k = 0; k < MCF_SUBS; k++) zsub_out[k] = 0.0;
    const uint64_t end = chunk_base + MCF_CHUNK;
    if (cur.pos < end) {
        ZigFsm s;
        zig_reset(s);
        unsigned j = (unsigned)(cur.pos - chunk_base);  // offset of the next draw inside the chunk (may pass 64)
        unsigned js = 0;                                // offset where the attempt in progress started
        // Attempt starts are monotone and an attempt emits before the next one starts, so one running sum that is
        // flushed whenever a start enters a new sub-chunk yields the per-sub-chunk sums in stream order.
        unsigned cursub = j >> MCF_SUB_SHIFT;
        double zc = 0.0;
        consume_draws(cur, [&](uint64_t r) {
            if (s.state == 0) {
                js = j;
                ci.startmask |= 1ULL << js;
                if (js < MCF_CHUNK && (js >> MCF_SUB_SHIFT) != cursub) {
                    zsub_out[cursub] = zc;
                    cursub = js >> MCF_SUB_SHIFT;
                    zc = 0.0;
                }
            }
            double x;
            if (zig_feed(s, r, t, x)) {
                ci.emitmask |= 1ULL << js;
                zc = xadd(zc, x);
            }
            j++;
            return !(s.state == 0 && j >= MCF_CHUNK);  // go on inside the chunk, or until the straddling attempt ends
        });
        if (cursub < MCF_SUBS) zsub_out[cursub] = zc;
    }
    const uint64_t over = cur.pos - end;
    ci.exit = over >= MCF_EXIT_OVERFLOW ? MCF_EXIT_OVERFLOW : (uint32_t)over;
    return ci;
}

// ---------------------------------------------------------------------------------------
// Speculative walk of one chunk of a block-aligned Philox stream (entry offset 0), branch-light form.
// Same result as scan_chunk(cur, chunk_base, 0, ...) up to the rounding of the sub-chunk sums.
// `load(blk, a, b, c, d)` yields the four draws of Philox block `blk` (relative to the stream's base counter) and is
// called for CONSECUTIVE blocks starting at `blk0`, the block holding the chunk's first draw (the device keeps round 0's
// product as a running value, philox_r0_next); one block of look-ahead is always loaded.
//
// Phase 1 touches every draw once and has no dependence between draws: the draw is looked at as if an attempt started
// there (one 16-byte table load, one FP64 compare); when it is "fast-looking" (rabs < ki) its value joins the sub-chunk
// sum with one fused multiply-add (see MCF_SCAN_VALUE), and a draw that is not (1.57 %) is parked in a per-thread slot
// together with its two successors.  Phase 2 resolves the chain from the parked draws alone, in order:
//   * what a parked draw yields -- accepted or not, its normal, how many draws it consumes, the fast values of the two
//     draws behind it -- does not depend on the chain, so `eval_parked` evaluates every parked draw EAGERLY (on the
//     device the warp's parked draws are compacted and spread over the lanes: ~1.5 dense rounds of the wedge
//     arithmetic instead of 3.3-3.7 sparse ones);
//   * the chain walk then only reads results: a parked draw that an earlier attempt consumed is not a start -- drop it;
//     otherwise the draws it consumes stop being starts (where they were fast-looking their value leaves the sum
//     again) and its flags give its emit bit and normal.
// Fast-looking draws that nobody consumed are exactly the fast attempts, so
//     startmask = ~consumed,  emitmask = (fast-looking & ~consumed) | accepted.
// More parked draws than slots (4e-3 of the chunks) or a tail whose first pair is rejected mark the chunk invalid
// for the general re-walk.
// ---------------------------------------------------------------------------------------
#define MCF_PARK_CAP 4
#ifndef MCF_SCAN_UNROLL
#define MCF_SCAN_UNROLL 1  // sub-chunk iterations of the phase-1 loop per loop body
#endif
#define MCF_PRAGMA(x) _Pragma(#x)
#define MCF_UNROLL(n) MCF_PRAGMA(unroll n)
struct ParkSlots {  // slot k of this thread: a[k * stride], b[k * stride], c[k * stride], f[k * stride]
    uint64_t *a, *b, *c;
    unsigned stride;
    uint8_t *f;  // flags of the evaluated slot
};
MCF_HD double zig_fast_value(uint64_t w, const ZigTables &t) {  // signed value of a fast-looking draw
    const uint64_t dbits = ((w >> 9) & 0x000fffffffffffffULL) | MCF_SCAN_BIAS;
    return MCF_SCAN_VALUE(bits_to_double(dbits), t.kws[(unsigned)w & 0x1ffu].wi);
}

#define MCF_PE_ACC 1u   // the attempt emits a normal
#define MCF_PE_LEN2 2u  // tail attempt: consumes the next two draws (else wedge: the next one)
#define MCF_PE_BAD 4u   // tail whose first pair is rejected: the chunk goes to the general walk
#define MCF_PE_LEN0 8u  // a fast attempt after all (never set by the FP64 compare of phase 1; kept for exactness)
struct ParkEval {
    double x, fv1, fv2;
    unsigned flags;
};
MCF_HD ParkEval park_eval(uint64_t a, uint64_t u, uint64_t u2, const ZigTables &t) {
    ParkEval r;
    r.x = 0.0;
    r.flags = 0;
    ZigFsm f;
    zig_reset(f);
    double x = 0.0;
    if (zig_feed(f, a, t, x)) {
        r.flags = MCF_PE_ACC | MCF_PE_LEN0;
        r.x = x;
    } else if ((a & 0xffULL) != 0) {  // wedge: same arithmetic and decision as zig_feed, state 1
        if (zig_feed(f, u, t, x)) {
            r.flags = MCF_PE_ACC;
            r.x = x;
        }
    } else {  // tail, first pair
        const uint64_t rabs = (a >> 9) & 0x000fffffffffffffULL;
        const double xx = xmul(-MCF_ZIG_INV_R, log1p(-u64_to_unit_double(u)));
        const double yy = -log1p(-u64_to_unit_double(u2));
        const double m = xadd(MCF_ZIG_R, xx);
        r.x = ((rabs >> 8) & 1ULL) ? -m : m;
        r.flags = MCF_PE_LEN2 | ((xadd(yy, yy) > xmul(xx, xx)) ? MCF_PE_ACC : MCF_PE_BAD);
    }
    r.fv1 = zig_fast_value(u, t);
    r.fv2 = zig_fast_value(u2, t);
    return r;
}
MCF_HD void park_eval_store(const ParkSlots &park, unsigned slot, const ParkEval &r) {
    union {
        double d;
        uint64_t u;
    } cv;
    cv.d = r.x;
    park.a[slot] = cv.u;
    cv.d = r.fv1;
    park.b[slot] = cv.u;
    cv.d = r.fv2;
    park.c[slot] = cv.u;
    park.f[slot] = (uint8_t)r.flags;
}
// every thread evaluates its own parked draws (host emulation, kernels whose warps may be partial)
MCF_HD void eval_parked_local(const ParkSlots &park, unsigned n_parked, const ZigTables &t) {
    for (unsigned k = 0; k < n_parked; k++) {
        const unsigned slot = k * park.stride;
        park_eval_store(park, slot, park_eval(park.a[slot], park.b[slot], park.c[slot], t));
    }
}

//
// Entry patch.  `entry_of(exit)` is called once, with the chunk's speculative exit, and returns the exit of the
// PREVIOUS chunk of the stream where the caller can know it cheaply (the neighbouring lane's; 0 otherwise).  A
// non-zero entry e (1 or 2 draws consumed by the previous chunk's last attempt) whose skipped draws were plain fast
// attempts leaves the chain untouched from e on: their values leave the first sub-chunk sum, their start bits are
// cleared, and the chunk is recorded as generated with entry e -- no re-walk (97 % of the chunks that would be
// re-walked).  Anything else marks the chunk invalid for the general walk.
//
// Look-ahead.  The last two draws of a chunk may need the first two draws of the NEXT chunk (wedge uniform, tail pair).
// `share_first(w0, w1)` is called once with the chunk's own first two draws and `lookahead(a, b)` must later return
// the next chunk's: on the device the neighbouring thread's (through shared memory), so the 17th Philox block per
// chunk is generated by one thread per CTA only; the host build simply generates it.
// `eval_parked(n)`: afterwards slot k < n of this thread holds {x, fv1, fv2} (bit patterns in a/b/c) and its flags
template <class Load, class EntryOf, class ShareFirst, class LookAhead, class EvalParked>
MCF_HD ChunkInfo scan_chunk_fast2(Load load, uint64_t blk0, const ZigTables &t, const ParkSlots &park, double *zsub_out,
                                  bool *valid, EntryOf entry_of, ShareFirst share_first, LookAhead lookahead,
                                  EvalParked eval_parked) {
    uint64_t F = 0;   // fast-looking draws
    unsigned np = 0;  // parked draws
    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
    uint64_t w0, w1, w2, w3, n0, n1, n2, n3;
    load(blk0, w0, w1, w2, w3);
    share_first(w0, w1);
    const double vk0 = zig_fast_value(w0, t), vk1 = zig_fast_value(w1, t);  // the first two draws (entry patch)
#define MCF_FAST2_DRAW(J, WA, WB, WC)                                                       \
    {                                                                                       \
        const uint64_t dbits = (((WA) >> 9) & 0x000fffffffffffffULL) | MCF_SCAN_BIAS;       \
        const ZigKW kw = t.kws[(unsigned)(WA)&0x1ffu];                                      \
        if (bits_to_double(dbits) < bits_to_double(kw.ki)) {                                \
            f16 |= 1u << (J);                                                               \
            zc = fma(bits_to_double(dbits), kw.wi, zc); /* x * 2^-100, one rounding */      \
        } else {                                                                            \
            if (np < MCF_PARK_CAP) {                                                        \
                park.a[np * park.stride] = (WA);                                            \
                park.b[np * park.stride] = (WB);                                            \
                park.c[np * park.stride] = (WC);                                            \
            }                                                                               \
            np++;                                                                           \
        }                                                                                   \
    }
MCF_UNROLL(MCF_SCAN_UNROLL)
    for (unsigned sub = 0; sub < MCF_SUBS; sub++) {
        double zc = 0.0;
        unsigned f16 = 0;
        const uint64_t b = blk0 + sub * 4;
        // blocks b, b+2 live in w0..w3, blocks b+1, b+3 in n0..n3; the block after the current one is always loaded
        load(b + 1, n0, n1, n2, n3);
        MCF_FAST2_DRAW(0, w0, w1, w2)
        MCF_FAST2_DRAW(1, w1, w2, w3)
        MCF_FAST2_DRAW(2, w2, w3, n0)
        MCF_FAST2_DRAW(3, w3, n0, n1)
        load(b + 2, w0, w1, w2, w3);
        MCF_FAST2_DRAW(4, n0, n1, n2)
        MCF_FAST2_DRAW(5, n1, n2, n3)
        MCF_FAST2_DRAW(6, n2, n3, w0)
        MCF_FAST2_DRAW(7, n3, w0, w1)
        load(b + 3, n0, n1, n2, n3);
        MCF_FAST2_DRAW(8, w0, w1, w2)
        MCF_FAST2_DRAW(9, w1, w2, w3)
        MCF_FAST2_DRAW(10, w2, w3, n0)
        MCF_FAST2_DRAW(11, w3, n0, n1)
        if (sub + 1 < MCF_SUBS)
            load(b + 4, w0, w1, w2, w3);
        else
            lookahead(w0, w1);  // first two draws of the next chunk
        MCF_FAST2_DRAW(12, n0, n1, n2)
        MCF_FAST2_DRAW(13, n1, n2, n3)
        MCF_FAST2_DRAW(14, n2, n3, w0)
        MCF_FAST2_DRAW(15, n3, w0, w1)
        F |= (uint64_t)f16 << (16 * sub);
        z0 = z1;
        z1 = z2;
        z2 = z3;
        z3 = zc;
    }
#undef MCF_FAST2_DRAW
    z0 = xmul(z0, 0x1p100);  // exact: back from the scaled accumulation range
    z1 = xmul(z1, 0x1p100);
    z2 = xmul(z2, 0x1p100);
    z3 = xmul(z3, 0x1p100);
    // phase 2: the chain, from the parked draws
    bool ok = np <= MCF_PARK_CAP;
    uint64_t S = ~0ULL, A = 0, N = ~F;
    unsigned exitv = 0, k = 0;
    auto add_to_sub = [&](unsigned sub, double x) {
        z0 = sub == 0 ? xadd(z0, x) : z0;
        z1 = sub == 1 ? xadd(z1, x) : z1;
        z2 = sub == 2 ? xadd(z2, x) : z2;
        z3 = sub == 3 ? xadd(z3, x) : z3;
    };
    eval_parked(np < MCF_PARK_CAP ? np : (unsigned)MCF_PARK_CAP);
    while (N != 0 && k < MCF_PARK_CAP) {
        const unsigned j = (unsigned)ctz64(N);
        N &= N - 1;
        const unsigned slot = k * park.stride;
        k++;
        if (!((S >> j) & 1ULL)) continue;  // consumed by an earlier attempt
        const unsigned fl = park.f[slot];
        const unsigned len = (fl & MCF_PE_LEN0) ? 0u : ((fl & MCF_PE_LEN2) ? 2u : 1u);
        if (fl & MCF_PE_BAD) ok = false;  // longer tail: general re-walk
        if (fl & MCF_PE_ACC) {
            A |= 1ULL << j;
            add_to_sub(j >> MCF_SUB_SHIFT, bits_to_double(park.a[slot]));
        }
        for (unsigned q = 1; q <= len; q++) {
            const unsigned f = j + q;
            if (f >= MCF_CHUNK) {
                exitv++;
                continue;
            }
            S &= ~(1ULL << f);
            if ((F >> f) & 1ULL) add_to_sub(f >> MCF_SUB_SHIFT, -bits_to_double(q == 1 ? park.b[slot] : park.c[slot]));
        }
    }
    const unsigned e = entry_of(exitv);
    unsigned gen_entry = 0;
    if (e != 0) {
        const uint64_t low = (1ULL << (e & 63u)) - 1ULL;
        if (ok && e <= 2 && (F & S & low) == low) {
            z0 = xsub(z0, vk0);
            if (e == 2) z0 = xsub(z0, vk1);
            S &= ~low;
            gen_entry = e;
        } else {
            ok = false;
        }
    }
    zsub_out[0] = z0;
    zsub_out[1] = z1;
    zsub_out[2] = z2;
    zsub_out[3] = z3;
    ChunkInfo ci;
    ci.gen_entry = gen_entry;
    ci.startmask = S;
    ci.emitmask = (S & F) | A;
    ci.exit = exitv;
    *valid = ok;
    return ci;
}

// For a chunk whose emissions have stream-local ordinals [E, E+count): find every replication
// boundary inside.  Replication m (m >= 1) starts right after the attempt that emitted normal
// number m*nps-1.  `visit(m, position)` is called for m in [1, n_sims]; m == n_sims marks the
// end of the stream's last replication (total consumption).
template <class F>
MCF_HD void locate_in_chunk(uint64_t chunk_base, uint64_t eff_start, uint64_t eff_emit, uint32_t exit_, uint64_t E,
                            uint64_t nps, uint64_t n_sims, F visit) {
    uint64_t cnt = (uint64_t)popc64(eff_emit);
    if (cnt == 0) return;
    uint64_t m = (E + nps) / nps;  // smallest m with m*nps - 1 >= E
    for (; m <= n_sims; m++) {
        uint64_t k = m * nps - 1;
        if (k >= E + cnt) break;
        int j = select_bit64(eff_emit, (int)(k - E));
        uint64_t later = (j >= 63) ? 0ULL : (eff_start & ~((2ULL << j) - 1ULL));
        uint64_t endpos = later ? (uint64_t)ctz64(later) : (uint64_t)MCF_CHUNK + exit_;
        visit(m, chunk_base + endpos);
    }
}

// ---------------------------------------------------------------------------------------
// Mask-driven replay.  The planning pass has already decided, for every draw position of a stream, whether an
// attempt that STARTS there emits a normal (emitmask, valid from the chunk's true entry offset).  A normal that
// is emitted by the attempt starting at draw v is always the fast-path value +-rabs(v)*wi[idx(v)] -- also when it
// was accepted by the wedge test -- except in the tail (idx == 0, rabs >= ki[0]).  So the replay needs no state
// machine, no acceptance tests and no ki/fi tables: for each set bit, convert one draw.  The stream range of
// replication i is [sim_start[i], sim_start[i+1]) (or the stream's end), so there is no emit counter either.
// ---------------------------------------------------------------------------------------
struct PlanView {
    const uint64_t *emit;  // [n_streams * cps]
    const uint8_t *entry;  // [n_streams * cps] true entry offset of every chunk
    const double *zsum;    // [n_streams * cps * MCF_SUBS] sums of the normals of the attempts starting in each sub-chunk
    int64_t cps;           // chunks per stream
};

MCF_HD uint64_t plan_emit_word(const PlanView &pv, int64_t si, uint64_t chunk) {
    const int64_t id = si * pv.cps + (int64_t)chunk;
    return pv.emit[id] & mask_from(pv.entry[id]);
}

// the normal emitted by the tail attempt whose first draw (at position p) carried `rabs`
template <int KIND>
MCF_HD double zig_tail_value(const mcf_stream &st, uint64_t p, uint64_t rabs) {
    Cursor<KIND> c;
    c.init(st, p + 1);
    for (;;) {
        const double xx = xmul(-MCF_ZIG_INV_R, log1p(-u64_to_unit_double(c.next64())));
        const double yy = -log1p(-u64_to_unit_double(c.next64()));
        if (xadd(yy, yy) > xmul(xx, xx)) {
            const double m = xadd(MCF_ZIG_R, xx);
            return ((rabs >> 8) & 1ULL) ? -m : m;
        }
    }
}

// `tail(p, rabs)` yields the normal of the tail attempt whose first draw sits at position p (3e-4 of the emits)
template <class Tail>
MCF_HD double zig_value_of_start_t(Tail tail, uint64_t p, uint64_t r, const double *wi, uint64_t ki0) {
    const unsigned idx = (unsigned)(r & 0xffULL);
    const unsigned sign = (unsigned)((r >> 8) & 1ULL);
    const uint64_t rabs = (r >> 9) & 0x000fffffffffffffULL;
    if (idx == 0 && rabs >= ki0) return tail(p, rabs);
    const double v = xmul(u52_to_double(rabs), wi[idx]);
    return sign ? -v : v;
}
template <int KIND>
MCF_HD double zig_value_of_start(const mcf_stream &st, uint64_t p, uint64_t r, const double *wi, uint64_t ki0) {
    const unsigned idx = (unsigned)(r & 0xffULL);
    const unsigned sign = (unsigned)((r >> 8) & 1ULL);
    const uint64_t rabs = (r >> 9) & 0x000fffffffffffffULL;
    if (idx == 0 && rabs >= ki0) return zig_tail_value<KIND>(st, p, rabs);  // 3e-4 of the emits
    const double v = xmul(u52_to_double(rabs), wi[idx]);
    return sign ? -v : v;
}

// Block-aligned Philox streams (off % 4 == 0): f(z) for every normal emitted by attempts starting in [start, end), whole
// blocks of four draws, every lane generating its block in the same loop trip.  `load(blk, a, b, c, d)` yields block
// `blk` counted from the stream's base counter.
template <class Load, class Tail, class F>
MCF_HD void replay_emits_blocks(Load load, Tail tail, const PlanView &pv, int64_t si, uint64_t off, uint64_t start,
                                uint64_t end, const double *wi, uint64_t ki0, F f) {
    if (end <= start) return;
    const uint64_t b_first = (start + off) >> 2, b_last = (end - 1 + off) >> 2;
    uint64_t chunk = ~0ULL, word = 0;
    for (uint64_t blk = b_first; blk <= b_last; blk++) {
        const uint64_t p0 = (blk << 2) - off;  // stream position of the block's first draw (may precede `start`)
        const uint64_t c = (p0 + 3) >> 6;      // blocks never straddle chunks when off % 4 == 0 (p0 % 4 == 0)
        if (c != chunk) {
            chunk = c;
            word = plan_emit_word(pv, si, c);
        }
        unsigned nib = (unsigned)(word >> (p0 & 63ULL)) & 0xFu;
        if (blk == b_first) nib &= (0xFu << (unsigned)(start - p0)) & 0xFu;
        if (blk == b_last) nib &= 0xFu >> (unsigned)(p0 + 4 - end);
        uint64_t v0, v1, v2, v3;
        load(blk, v0, v1, v2, v3);
        if (nib & 1u) f(zig_value_of_start_t(tail, p0, v0, wi, ki0));
        if (nib & 2u) f(zig_value_of_start_t(tail, p0 + 1, v1, wi, ki0));
        if (nib & 4u) f(zig_value_of_start_t(tail, p0 + 2, v2, wi, ki0));
        if (nib & 8u) f(zig_value_of_start_t(tail, p0 + 3, v3, wi, ki0));
    }
}

// Sum of the normals emitted by attempts starting in [start, end) -- the same additions in the same order as
// replay_emits_blocks with an accumulating f, but branch-free per draw (converted always, added under the emit bit):
// lanes of a warp sit at different offsets inside their blocks, and four data-dependent branches per block would
// serialise them.  Only the tail (3e-4 of the emits) branches.
template <class Load, class Tail>
MCF_HD double sum_emits_blocks(Load load, Tail tail, const PlanView &pv, int64_t si, uint64_t off, uint64_t start,
                               uint64_t end, const double *wi, uint64_t ki0) {
    double acc = 0.0;
    if (end <= start) return acc;
    const uint64_t b_first = (start + off) >> 2, b_last = (end - 1 + off) >> 2;
    uint64_t chunk = ~0ULL, word = 0;
    for (uint64_t blk = b_first; blk <= b_last; blk++) {
        const uint64_t p0 = (blk << 2) - off;
        const uint64_t c = (p0 + 3) >> 6;
        if (c != chunk) {
            chunk = c;
            word = plan_emit_word(pv, si, c);
        }
        unsigned nib = (unsigned)(word >> (p0 & 63ULL)) & 0xFu;
        if (blk == b_first) nib &= (0xFu << (unsigned)(start - p0)) & 0xFu;
        if (blk == b_last) nib &= 0xFu >> (unsigned)(p0 + 4 - end);
        uint64_t v[4];
        load(blk, v[0], v[1], v[2], v[3]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint64_t r = v[k];
            const unsigned idx = (unsigned)(r & 0xffULL);
            const uint64_t rabs = (r >> 9) & 0x000fffffffffffffULL;
            const double m = xmul(u52_to_double(rabs), wi[idx]);
            double z = ((r >> 8) & 1ULL) ? -m : m;
            const bool on = (nib >> k) & 1u;
            if (on && idx == 0 && rabs >= ki0) z = tail(p0 + (uint64_t)k, rabs);
            acc = on ? xadd(acc, z) : acc;
        }
    }
    return acc;
}

// A replication boundary at stream position p cuts its 16-draw sub-chunk in two.  Only the SHORTER side is regenerated
// (at most two Philox blocks, one loop for both cases so that a warp does not run it twice); the other side is the
// stored sub-chunk sum minus it.  `before` belongs to the replication that ends at p, `after` to the one that starts
// there.  An aligned boundary (offset 0) regenerates nothing: before = 0, after = the whole sum.
struct BoundarySums {
    double before, after;
};
template <class Load, class Tail>
MCF_HD BoundarySums boundary_sums(Load load, Tail tail, const PlanView &pv, int64_t si, uint64_t off, uint64_t p,
                                  const double *wi, uint64_t ki0) {
    const uint64_t s = p >> MCF_SUB_SHIFT, o = p & ((1ULL << MCF_SUB_SHIFT) - 1ULL);
    const double S = pv.zsum[si * pv.cps * MCF_SUBS + (int64_t)s];
    const bool head = o <= 8;  // regenerate [sub-chunk start, p), else [p, sub-chunk end)
    const double part = sum_emits_blocks(load, tail, pv, si, off, head ? (s << MCF_SUB_SHIFT) : p,
                                         head ? p : ((s + 1) << MCF_SUB_SHIFT), wi, ki0);
    const double rest = xsub(S, part);
    BoundarySums r;
    r.before = head ? part : rest;
    r.after = head ? rest : part;
    return r;
}
// sum of the normals of the replication [start, end) given the two boundary cuts (requires end - start >= 16):
// after(start), the whole sub-chunks in between in order, before(end)
MCF_HD double zsum_from_boundaries(const PlanView &pv, int64_t si, uint64_t start, uint64_t end, double after_start,
                                   double before_end) {
    const double *sums = pv.zsum + si * pv.cps * MCF_SUBS;
    double acc = after_start;
    for (uint64_t k = (start >> MCF_SUB_SHIFT) + 1; k < (end >> MCF_SUB_SHIFT); k++) acc = xadd(acc, sums[k]);
    return xadd(acc, before_end);
}

// f(z) for every normal emitted by attempts starting in [start, end) of stream `si`, in stream order
template <class F>
MCF_HD void replay_emits(Cursor<MCF_GEN_PCG64> &cur, const mcf_stream &st, const PlanView &pv, int64_t si, uint64_t start,
                         uint64_t end, const double *wi, uint64_t ki0, F f) {
    if (end <= start) return;
    cur.init(st, start);
    uint64_t word = plan_emit_word(pv, si, start >> 6);
    for (uint64_t p = start; p < end; p++) {
        if ((p & 63ULL) == 0 && p != start) word = plan_emit_word(pv, si, p >> 6);
        const uint64_t r = cur.next64();
        if ((word >> (p & 63ULL)) & 1ULL) f(zig_value_of_start<MCF_GEN_PCG64>(st, p, r, wi, ki0));
    }
}

// Philox: whole blocks of four draws, every lane generating its block in the same loop trip.
// Requires a block-aligned stream (buf_pos % 4 == 0: every stream of a parallel run); see replay_emits_any.
template <class F>
MCF_HD void replay_emits(Cursor<MCF_GEN_PHILOX> &cur, const mcf_stream &st, const PlanView &pv, int64_t si, uint64_t start,
                         uint64_t end, const double *wi, uint64_t ki0, F f) {
    if (end <= start) return;
    cur.init(st, start);
    const uint64_t off = (uint64_t)st.buf_pos;
    if (off & 3ULL) {  // generator captured mid-buffer (user-supplied Philox): draw by draw
        uint64_t w = plan_emit_word(pv, si, start >> 6);
        for (uint64_t p = start; p < end; p++) {
            if ((p & 63ULL) == 0 && p != start) w = plan_emit_word(pv, si, p >> 6);
            const uint64_t r = cur.next64();
            if ((w >> (p & 63ULL)) & 1ULL) f(zig_value_of_start<MCF_GEN_PHILOX>(st, p, r, wi, ki0));
        }
        return;
    }
    replay_emits_blocks(
        [&](uint64_t blk, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
            cur.load_block(blk);
            a = cur.v0;
            b = cur.v1;
            c = cur.v2;
            d = cur.v3;
        },
        [&](uint64_t p, uint64_t rabs) { return zig_tail_value<MCF_GEN_PHILOX>(st, p, rabs); }, pv, si, off, start, end,
        wi, ki0, f);
}

// Sum of the normals of the attempts starting in [start, end): whole 16-draw sub-chunks come from the sums the
// planning pass stored, only the two partial sub-chunks at the ends are regenerated (about 6 % of the draws for
// 252 steps, 0.6 % for 2520).  Summation order: head partial, stored sums in order, tail partial (deterministic).
template <int KIND>
MCF_HD double replay_zsum(Cursor<KIND> &cur, const mcf_stream &st, const PlanView &pv, int64_t si, uint64_t start,
                          uint64_t end, const double *wi, uint64_t ki0) {
    double acc = 0.0;
    if (end <= start) return acc;
    const uint64_t s0 = start >> MCF_SUB_SHIFT, s1 = (end - 1) >> MCF_SUB_SHIFT;
    auto add = [&](double z) { acc = xadd(acc, z); };
    if (s0 == s1) {
        replay_emits(cur, st, pv, si, start, end, wi, ki0, add);
        return acc;
    }
    const double *sums = pv.zsum + si * pv.cps * MCF_SUBS;
    if ((start & ((1ULL << MCF_SUB_SHIFT) - 1ULL)) == 0)
        acc = sums[s0];  // aligned start: nothing of this sub-chunk belongs to the previous replication
    else
        replay_emits(cur, st, pv, si, start, (s0 + 1) << MCF_SUB_SHIFT, wi, ki0, add);
    for (uint64_t k = s0 + 1; k < s1; k++) acc = xadd(acc, sums[k]);
    replay_emits(cur, st, pv, si, s1 << MCF_SUB_SHIFT, end, wi, ki0, add);
    return acc;
}

// ---------------------------------------------------------------------------------------
// Replication bodies (pass 2): the cursor sits where the replication's first draw starts.
// ---------------------------------------------------------------------------------------
struct GbmDerived {  // host-precomputed exactly as the reference computes them in Python floats
    int32_t mode;
    double a;      // per-step drift      (r - 0.5*sigma^2)*dt   | loc of rng.normal
    double b;      // per-step vol scale  sigma*sqrt(dt)          | scale of rng.normal
    double s0;     // S0 | V0
    double logs0;  // log(S0)
    double K;
    double disc;   // exp(-r*T)
    double rdt;    // r*dt   (discount exponent per step, "american")
};

// European / portfolio terminal value from a stream of normals; sum in fp64, sequential order
// (the reference sums pairwise; the difference is covered by the stated tolerance).
// `each(f)` must call f(z) for every normal of the replication, in stream order.
template <class Each>
MCF_HD double terminal_value(Each each, const GbmDerived &g) {
    switch (g.mode) {
    case MCF_NORMAL_SUM: {
        double acc = 0.0;
        each([&](double z) { acc = xadd(acc, z); });
        return acc;
    }
    case MCF_NORMAL_LOC_SCALE: {  // rng.normal(loc, scale) = loc + scale*z, summed over the replication's draws
        double acc = 0.0;
        const double a = g.a, b = g.b;
        each([&](double z) { acc = xadd(acc, xadd(a, xmul(b, z))); });
        return acc;
    }
    case MCF_BS_EURO_CALL:
    case MCF_BS_EURO_PUT:
    case MCF_PORTFOLIO_GBM:
    case MCF_PATH_TERMINAL: {
        double acc = 0.0;
        const double a = g.a, b = g.b;
        each([&](double z) { acc = xadd(acc, xadd(a, xmul(b, z))); });
        if (g.mode == MCF_PATH_TERMINAL) return exp(xadd(g.logs0, acc));  // exp(log(S0) + cumsum[-1])
        const double st = xmul(g.s0, exp(acc));
        if (g.mode == MCF_PORTFOLIO_GBM) return st;
        const double pay = g.mode == MCF_BS_EURO_CALL ? fmax(xsub(st, g.K), 0.0) : fmax(xsub(g.K, st), 0.0);
        return xmul(pay, g.disc);
    }
    case MCF_PORTFOLIO_LOG1P: {
        double acc = 0.0;
        const double a = g.a, b = g.b;
        each([&](double z) { acc = xadd(acc, log1p(xadd(a, xmul(b, z)))); });
        return xmul(g.s0, exp(acc));
    }
    case MCF_BS_AMER_CALL:
    case MCF_BS_AMER_PUT: {
        const bool call = g.mode == MCF_BS_AMER_CALL;
        const double a = g.a, b = g.b, K = g.K, logs0 = g.logs0, nrdt = -g.rdt;
        const double s0 = exp(logs0);
        double best = call ? fmax(xsub(s0, K), 0.0) : fmax(xsub(K, s0), 0.0);  // k = 0, discount 1
        double acc = 0.0, kf = 0.0;
        each([&](double z) {
            kf += 1.0;  // exact for every representable step count
            acc = xadd(acc, xadd(a, xmul(b, z)));
            const double s = exp(xadd(logs0, acc));
            const double intr = call ? fmax(xsub(s, K), 0.0) : fmax(xsub(K, s), 0.0);
            best = fmax(best, xmul(intr, exp(xmul(nrdt, kf))));
        });
        return best;
    }
    default:
        return bits_to_double(0x7ff8000000000000ULL);
    }
}

// state-machine driven form (no plan masks needed): tests, single replications
template <class Cur>
MCF_HD double run_terminal(Cur &cur, const ZigTables &t, int64_t n_steps, const GbmDerived &g) {
    return terminal_value([&](auto f) { for_each_normal(cur, t, n_steps, f); }, g);
}

}  // namespace mcf
