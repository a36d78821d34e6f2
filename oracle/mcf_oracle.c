/*
 * mcf_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the arithmetic on the hot path of the reference
 * (milanfusco/mcFramework): the replication loop behind
 * MonteCarloSimulation.run() and the draws its built-in simulations make.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.  The product
 * (mcframework_b200/) never links, imports or calls it.
 *
 * The reference itself is pure Python; all arithmetic lives in its pinned
 * third-party dependency NumPy (pyproject.toml:18 "numpy>=1.24"; the oracle is
 * pinned to NumPy 2.3.5 as installed in this image).  What is restated here is
 * therefore NumPy's published algorithms, anchored on the reference call sites:
 *
 *   SeedSequence(seed).spawn / generate_state   core.py:363,637 (numpy/random/bit_generator.pyx)
 *   PCG64 (default_rng)                         core.py:315,328,364; stats_engine.py:341
 *   Philox4x64-10                               core.py:102-103,659
 *   Generator.uniform / random                  sims/pi.py:71,76,80
 *   Generator.standard_normal / normal          sims/black_scholes.py:103,239; sims/portfolio.py:81,83
 *   Generator.integers (32-bit Lemire)          stats_engine.py:1040; core.py:25 (docstring dice)
 *   np.sum pairwise summation                   sims/black_scholes.py:241; sims/portfolio.py:82,84
 *
 * Pinning: tests/test_oracle_cpu.py checks every function below against
 * (1) NumPy's own known-answer files numpy/random/tests/data/{pcg64,philox}-testset-{1,2}.csv,
 * (2) live NumPy draws, and (3) the JSON fixtures under tests/golden/ produced by running the
 * unmodified reference in the build container (tools/gen_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "zig_tables_oracle.h"
static const unsigned long long MCF_ZIG_KI[256] = {MCF_ZIG_KI_LIST};
static const unsigned long long MCF_ZIG_WI_BITS[256] = {MCF_ZIG_WI_BITS_LIST};
static const unsigned long long MCF_ZIG_FI_BITS[256] = {MCF_ZIG_FI_BITS_LIST};

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------ */
/* SeedSequence (numpy/random/bit_generator.pyx: SeedSequence)               */
/* ------------------------------------------------------------------------ */
#define SS_XSHIFT 16
#define SS_INIT_A 0x43b0d7e5u
#define SS_MULT_A 0x931e8875u
#define SS_INIT_B 0x8b51f9ddu
#define SS_MULT_B 0x58f38dedu
#define SS_MIX_L 0xca01f9ddu
#define SS_MIX_R 0x4973f715u
#define SS_POOL 4

static uint32_t ss_hashmix(uint32_t value, uint32_t *hash_const) {
    value ^= *hash_const;
    *hash_const *= SS_MULT_A;
    value *= *hash_const;
    value ^= value >> SS_XSHIFT;
    return value;
}

static uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = SS_MIX_L * x - SS_MIX_R * y;
    r ^= r >> SS_XSHIFT;
    return r;
}

/* entropy: little-endian uint32 words of the integer seed (at least one word);
 * spawn_key: the child's spawn key tuple.  out32 receives n_words32 words. */
void mcfo_seedseq_state(const uint32_t *entropy, int n_ent, const uint32_t *spawn_key, int n_spawn,
                        int n_words32, uint32_t *out32) {
    uint32_t assembled[64];
    int n = 0;
    for (int i = 0; i < n_ent; i++) assembled[n++] = entropy[i];
    if (n_spawn > 0) {
        while (n < SS_POOL) assembled[n++] = 0u; /* pad run entropy before appending spawn key */
        for (int i = 0; i < n_spawn; i++) assembled[n++] = spawn_key[i];
    }
    uint32_t pool[SS_POOL];
    uint32_t hc = SS_INIT_A;
    for (int i = 0; i < SS_POOL; i++) pool[i] = ss_hashmix(i < n ? assembled[i] : 0u, &hc);
    for (int s = 0; s < SS_POOL; s++)
        for (int d = 0; d < SS_POOL; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], &hc));
    for (int s = SS_POOL; s < n; s++)
        for (int d = 0; d < SS_POOL; d++) pool[d] = ss_mix(pool[d], ss_hashmix(assembled[s], &hc));
    hc = SS_INIT_B;
    for (int i = 0; i < n_words32; i++) {
        uint32_t v = pool[i % SS_POOL];
        v ^= hc;
        hc *= SS_MULT_B;
        v *= hc;
        v ^= v >> SS_XSHIFT;
        out32[i] = v;
    }
}

/* ------------------------------------------------------------------------ */
/* Bit generators                                                            */
/* ------------------------------------------------------------------------ */
enum { MCFO_PCG64 = 0, MCFO_PHILOX = 1 };

typedef struct {
    int kind;
    /* PCG64 XSL-RR 128/64 (numpy/random/src/pcg64/pcg64.h) */
    u128 state, inc;
    /* Philox4x64-10 (numpy/random/src/philox/philox.h) */
    uint64_t ctr[4], key[2], buf[4];
    int buf_pos;
    /* next_uint32 half-word buffer shared by both */
    int has_uint32;
    uint32_t uinteger;
    /* bookkeeping: raw 64-bit outputs consumed since init (test aid only) */
    uint64_t n_raw;
} mcfo_gen;

#define PCG_MULT ((((u128)0x2360ED051FC65DA4ULL) << 64) | (u128)0x4385DF649FCCF645ULL)

int mcfo_gen_sizeof(void) { return (int)sizeof(mcfo_gen); }

/* default_rng(seed_seq): w = seed_seq.generate_state(4, uint64) */
void mcfo_init_pcg64(mcfo_gen *g, const uint64_t w[4]) {
    memset(g, 0, sizeof(*g));
    g->kind = MCFO_PCG64;
    u128 initstate = ((u128)w[0] << 64) | w[1];
    u128 initseq = ((u128)w[2] << 64) | w[3];
    g->inc = (initseq << 1) | 1u;
    g->state = 0;
    g->state = g->state * PCG_MULT + g->inc;
    g->state += initstate;
    g->state = g->state * PCG_MULT + g->inc;
}

/* from bit_generator.state["state"]: explicit 128-bit state and inc */
void mcfo_init_pcg64_raw(mcfo_gen *g, uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo,
                         int has_uint32, uint32_t uinteger) {
    memset(g, 0, sizeof(*g));
    g->kind = MCFO_PCG64;
    g->state = ((u128)state_hi << 64) | state_lo;
    g->inc = ((u128)inc_hi << 64) | inc_lo;
    g->has_uint32 = has_uint32;
    g->uinteger = uinteger;
}

/* Philox(seed_seq): key = seed_seq.generate_state(2, uint64); counter = 0; empty buffer */
void mcfo_init_philox(mcfo_gen *g, const uint64_t key[2]) {
    memset(g, 0, sizeof(*g));
    g->kind = MCFO_PHILOX;
    g->key[0] = key[0];
    g->key[1] = key[1];
    g->buf_pos = 4;
}

/* O(log n) LCG jump (PCG reference: pcg_advance_lcg_128) */
void mcfo_pcg64_advance(mcfo_gen *g, uint64_t delta_hi, uint64_t delta_lo) {
    u128 delta = ((u128)delta_hi << 64) | delta_lo;
    u128 cur_mult = PCG_MULT, cur_plus = g->inc, acc_mult = 1, acc_plus = 0;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    g->state = acc_mult * g->state + acc_plus;
}

static inline uint64_t rotr64(uint64_t v, unsigned r) { return (v >> r) | (v << ((-r) & 63)); }

#define PHILOX_M0 0xD2E7470EE14C6C93ULL
#define PHILOX_M1 0xCA5A826395121157ULL
#define PHILOX_W0 0x9E3779B97F4A7C15ULL
#define PHILOX_W1 0xBB67AE8584CAA73BULL

static void philox_block(const uint64_t ctr_in[4], const uint64_t key_in[2], uint64_t out[4]) {
    uint64_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint64_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        if (r) {
            k0 += PHILOX_W0;
            k1 += PHILOX_W1;
        }
        u128 p0 = (u128)PHILOX_M0 * c0, p1 = (u128)PHILOX_M1 * c2;
        uint64_t hi0 = (uint64_t)(p0 >> 64), lo0 = (uint64_t)p0;
        uint64_t hi1 = (uint64_t)(p1 >> 64), lo1 = (uint64_t)p1;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

static inline uint64_t gen_next64(mcfo_gen *g) {
    g->n_raw++;
    if (g->kind == MCFO_PCG64) {
        g->state = g->state * PCG_MULT + g->inc; /* step, then output from the new state */
        uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state;
        return rotr64(hi ^ lo, (unsigned)(hi >> 58));
    }
    if (g->buf_pos < 4) return g->buf[g->buf_pos++];
    /* counter is incremented BEFORE the block is generated */
    if (++g->ctr[0] == 0)
        if (++g->ctr[1] == 0)
            if (++g->ctr[2] == 0) ++g->ctr[3];
    philox_block(g->ctr, g->key, g->buf);
    g->buf_pos = 1;
    return g->buf[0];
}

static inline uint32_t gen_next32(mcfo_gen *g) {
    if (g->has_uint32) {
        g->has_uint32 = 0;
        return g->uinteger;
    }
    uint64_t v = gen_next64(g);
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

static inline double gen_double(mcfo_gen *g) { return (double)(gen_next64(g) >> 11) * (1.0 / 9007199254740992.0); }

uint64_t mcfo_raw_consumed(const mcfo_gen *g) { return g->n_raw; }
void mcfo_get_u32_buffer(const mcfo_gen *g, int *has_uint32, uint32_t *uinteger) {
    *has_uint32 = g->has_uint32;
    *uinteger = g->uinteger;
}

void mcfo_raw64(mcfo_gen *g, int64_t n, uint64_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = gen_next64(g);
}
void mcfo_raw32(mcfo_gen *g, int64_t n, uint32_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = gen_next32(g);
}
void mcfo_doubles(mcfo_gen *g, int64_t n, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = gen_double(g);
}
/* Generator.uniform(low, high): low + (high-low)*next_double  (distributions.c random_uniform) */
void mcfo_uniform(mcfo_gen *g, double low, double high, int64_t n, double *out) {
    double range = high - low;
    for (int64_t i = 0; i < n; i++) out[i] = low + range * gen_double(g);
}

/* ------------------------------------------------------------------------ */
/* Ziggurat standard normal (distributions.c random_standard_normal)         */
/* ------------------------------------------------------------------------ */
static inline double bits2d(uint64_t b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}

static double gen_normal(mcfo_gen *g) {
    const double zig_r = bits2d(MCF_ZIG_NOR_R_BITS), zig_inv_r = bits2d(MCF_ZIG_NOR_INV_R_BITS);
    for (;;) {
        uint64_t r = gen_next64(g);
        int idx = (int)(r & 0xff);
        r >>= 8;
        int sign = (int)(r & 1);
        uint64_t rabs = (r >> 1) & 0x000fffffffffffffULL;
        double x = (double)rabs * bits2d(MCF_ZIG_WI_BITS[idx]);
        if (sign) x = -x;
        if (rabs < MCF_ZIG_KI[idx]) return x; /* ~98.5 % */
        if (idx == 0) {
            for (;;) {
                double xx = -zig_inv_r * log1p(-gen_double(g));
                double yy = -log1p(-gen_double(g));
                if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -(zig_r + xx) : zig_r + xx;
            }
        } else {
            double f1 = bits2d(MCF_ZIG_FI_BITS[idx - 1]), f0 = bits2d(MCF_ZIG_FI_BITS[idx]);
            if ((f1 - f0) * gen_double(g) + f0 < exp(-0.5 * x * x)) return x;
        }
    }
}

void mcfo_normals(mcfo_gen *g, int64_t n, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = gen_normal(g);
}

/* ------------------------------------------------------------------------ */
/* Bounded integers: Generator.integers(low, high) for high-low-1 < 2^32      */
/* (distributions.c buffered_bounded_lemire_uint32 via random_bounded_uint64_fill) */
/* ------------------------------------------------------------------------ */
static inline uint32_t gen_lemire32(mcfo_gen *g, uint32_t rng) {
    const uint32_t rng_excl = rng + 1u;
    uint64_t m = (uint64_t)gen_next32(g) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        const uint32_t threshold = (uint32_t)((0xFFFFFFFFu - rng) % rng_excl);
        while (leftover < threshold) {
            m = (uint64_t)gen_next32(g) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

/* integers(low, high, size=n) with 0 < high-low <= 2^32; out is int64 like NumPy's default dtype */
void mcfo_integers32(mcfo_gen *g, int64_t low, int64_t high, int64_t n, int64_t *out) {
    uint64_t rng = (uint64_t)(high - low) - 1u;
    for (int64_t i = 0; i < n; i++) {
        if (rng == 0)
            out[i] = low;
        else if (rng == 0xFFFFFFFFull)
            out[i] = low + (int64_t)gen_next32(g);
        else
            out[i] = low + (int64_t)gen_lemire32(g, (uint32_t)rng);
    }
}

/* ------------------------------------------------------------------------ */
/* np.sum over a contiguous float64 vector: pairwise summation                */
/* (numpy/_core/src/umath/loops_utils.h.src  @TYPE@_pairwise_sum)             */
/* ------------------------------------------------------------------------ */
static double np_pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}
double mcfo_np_sum(const double *a, int64_t n) { return np_pairwise_sum(a, n); }

/* ------------------------------------------------------------------------ */
/* Simulations: one call = n_sims consecutive single_simulation() calls on g  */
/* ------------------------------------------------------------------------ */

/* sims/pi.py:69-82.  hits_out may be NULL. */
void mcfo_pi(mcfo_gen *g, int64_t n_sims, int64_t n_points, int antithetic, double *out, int64_t *hits_out) {
    for (int64_t s = 0; s < n_sims; s++) {
        int64_t hits = 0;
        int64_t m = antithetic ? n_points / 2 : n_points;
        for (int64_t i = 0; i < m; i++) {
            double x = -1.0 + 2.0 * gen_double(g);
            double y = -1.0 + 2.0 * gen_double(g);
            double xx = x * x, yy = y * y;
            hits += (xx + yy <= 1.0);
        }
        if (antithetic) {
            hits *= 2; /* (-x,-y) lands inside iff (x,y) does */
            if (2 * m < n_points) {
                double x = -1.0 + 2.0 * gen_double(g);
                double y = -1.0 + 2.0 * gen_double(g);
                double xx = x * x, yy = y * y;
                hits += (xx + yy <= 1.0);
            }
        }
        if (hits_out) hits_out[s] = hits;
        out[s] = 4.0 * (double)hits / (double)n_points;
    }
}

/* Terminal-value modes.  p[] layout is per mode (see comments). */
enum {
    MCFO_BS_EURO_CALL = 0, /* p = {S0, K, T, r, sigma}  black_scholes.py:237-243 */
    MCFO_BS_EURO_PUT = 1,
    MCFO_BS_AMER_CALL = 2, /* max discounted intrinsic along the path  black_scholes.py:245-255 */
    MCFO_BS_AMER_PUT = 3,
    MCFO_PATH_TERMINAL = 4, /* p = {S0, r, sigma, T}  black_scholes.py:399-401 */
    MCFO_PORTFOLIO_GBM = 5, /* p = {V0, mu, sigma}; n_steps = int(years/dt)  portfolio.py:79-82 */
    MCFO_PORTFOLIO_LOG1P = 6 /* portfolio.py:83-84 */
};

/* _simulate_gbm_path (black_scholes.py:102-106): path has n_steps+1 entries */
static void gbm_path(mcfo_gen *g, double S0, double r, double sigma, double T, int64_t n_steps, double *path,
                     double *scratch) {
    double dt = T / (double)n_steps;
    double drift = (r - 0.5 * sigma * sigma) * dt, vol = sigma * sqrt(dt);
    double l0 = log(S0), acc = 0.0;
    for (int64_t k = 0; k < n_steps; k++) scratch[k] = gen_normal(g);
    path[0] = exp(l0);
    for (int64_t k = 0; k < n_steps; k++) {
        acc += drift + vol * scratch[k]; /* np.cumsum is a running sum */
        path[k + 1] = exp(l0 + acc);
    }
}

void mcfo_gbm_terminal(mcfo_gen *g, int64_t n_sims, int64_t n_steps, int mode, const double *p, double *out) {
    double *lr = (double *)malloc(sizeof(double) * (size_t)(n_steps + 1));
    double *path = (double *)malloc(sizeof(double) * (size_t)(n_steps + 1));
    for (int64_t s = 0; s < n_sims; s++) {
        switch (mode) {
        case MCFO_BS_EURO_CALL:
        case MCFO_BS_EURO_PUT: {
            double S0 = p[0], K = p[1], T = p[2], r = p[3], sigma = p[4];
            double dt = T / (double)n_steps;
            double drift = (r - 0.5 * sigma * sigma) * dt, vol = sigma * sqrt(dt);
            for (int64_t k = 0; k < n_steps; k++) lr[k] = drift + vol * gen_normal(g);
            double ST = S0 * exp(np_pairwise_sum(lr, n_steps));
            double pay = mode == MCFO_BS_EURO_CALL ? fmax(ST - K, 0.0) : fmax(K - ST, 0.0);
            out[s] = pay * exp(-r * T);
            break;
        }
        case MCFO_BS_AMER_CALL:
        case MCFO_BS_AMER_PUT: {
            double S0 = p[0], K = p[1], T = p[2], r = p[3], sigma = p[4];
            gbm_path(g, S0, r, sigma, T, n_steps, path, lr);
            double dt = T / (double)n_steps, best = -INFINITY;
            for (int64_t k = 0; k <= n_steps; k++) {
                double intr = mode == MCFO_BS_AMER_CALL ? fmax(path[k] - K, 0.0) : fmax(K - path[k], 0.0);
                double v = intr * exp(-r * dt * (double)k);
                if (v > best) best = v;
            }
            out[s] = best;
            break;
        }
        case MCFO_PATH_TERMINAL: {
            gbm_path(g, p[0], p[1], p[2], p[3], n_steps, path, lr);
            out[s] = path[n_steps];
            break;
        }
        case MCFO_PORTFOLIO_GBM: {
            double V0 = p[0], mu = p[1], sigma = p[2], dt = 1.0 / 252.0;
            double loc = (mu - 0.5 * sigma * sigma) * dt, scale = sigma * sqrt(dt);
            for (int64_t k = 0; k < n_steps; k++) lr[k] = loc + scale * gen_normal(g); /* random_normal */
            out[s] = V0 * exp(np_pairwise_sum(lr, n_steps));
            break;
        }
        case MCFO_PORTFOLIO_LOG1P: {
            double V0 = p[0], mu = p[1], sigma = p[2], dt = 1.0 / 252.0;
            double loc = mu * dt, scale = sigma * sqrt(dt);
            for (int64_t k = 0; k < n_steps; k++) lr[k] = log1p(loc + scale * gen_normal(g));
            out[s] = V0 * exp(np_pairwise_sum(lr, n_steps));
            break;
        }
        default:
            out[s] = NAN;
        }
    }
    free(lr);
    free(path);
}

/* BlackScholesPathSimulation.simulate_paths (black_scholes.py:415-418): row-major (n_paths, n_steps+1) */
void mcfo_gbm_paths(mcfo_gen *g, int64_t n_paths, int64_t n_steps, double S0, double r, double sigma, double T,
                    double *paths) {
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(n_steps + 1));
    for (int64_t i = 0; i < n_paths; i++) gbm_path(g, S0, r, sigma, T, n_steps, paths + i * (n_steps + 1), scratch);
    free(scratch);
}

/* stats_engine.py:1039-1041: idx = rng.integers(0, n, size=(B, n)); means = x[idx].mean(axis=1).
 * idx_out (optional) receives the first idx_keep indices of the stream. */
void mcfo_bootstrap_means(mcfo_gen *g, const double *x, int64_t n, int64_t B, double *means, int64_t *idx_out,
                          int64_t idx_keep) {
    double *row = (double *)malloc(sizeof(double) * (size_t)n);
    int64_t kept = 0;
    for (int64_t b = 0; b < B; b++) {
        for (int64_t i = 0; i < n; i++) {
            int64_t j = (n == 1) ? 0 : (int64_t)gen_lemire32(g, (uint32_t)(n - 1));
            if (idx_out && kept < idx_keep) idx_out[kept++] = j;
            row[i] = x[j];
        }
        means[b] = np_pairwise_sum(row, n) / (double)n; /* ndarray.mean(axis=1) on C-contiguous rows */
    }
    free(row);
}

/* user-defined dice of core.py:22-25: float(rng.integers(1, 7, size=2).sum()) */
void mcfo_dice_sums(mcfo_gen *g, int64_t n_sims, int64_t n_dice, int64_t low, int64_t high, double *out) {
    for (int64_t s = 0; s < n_sims; s++) {
        int64_t acc = 0, v;
        for (int64_t d = 0; d < n_dice; d++) {
            mcfo_integers32(g, low, high, 1, &v);
            acc += v;
        }
        out[s] = (double)acc;
    }
}
