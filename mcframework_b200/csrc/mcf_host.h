// mcf_host.h -- host-side geometry and parameter derivation shared by the C ABI (mcf_kernels.cu)
// and by the CPU emulation harness in tests/hostemu (which mirrors the kernels' decomposition).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "mcf_device.cuh"

#define MCF_TILE 1024  // chunks per tile
#define MCF_HDR_WORDS 16

struct PlanGeom {
    int64_t cps;        // chunks per stream (multiple of MCF_TILE)
    int64_t tps;        // tiles per stream
    int64_t n_chunks;   // n_streams * cps
    int64_t n_tiles;
    int64_t off_start, off_emit, off_exit, off_gen, off_entry, off_tsum, off_tpre, off_send, off_zsum, off_queue, q_cap,
        total;
};

// Device-resident description of a finished plan (at byte MCF_LAYOUT_OFFSET of the workspace): lets the replay
// kernels find the emit masks without the host having to remember the geometry.
#define MCF_QCOUNT_OFFSET 128  // int counters[16] inside the first 256 bytes of the workspace, one per fix-up iteration
#define MCF_LAYOUT_OFFSET 64
#define MCF_LAYOUT_MAGIC 0x6d63665f706c616eLL
struct PlanLayout {
    int64_t magic, cps, off_emit, off_entry, off_send, off_zsum, n_streams;
};

static inline PlanGeom plan_geom(int32_t n_streams, int64_t max_sims, int64_t nps, double slack) {
    PlanGeom g;
    double need = (double)max_sims * (double)nps;
    int64_t window = (int64_t)(need * (1.0 + slack)) + 4096;
    g.cps = ((window + MCF_CHUNK - 1) / MCF_CHUNK + MCF_TILE - 1) / MCF_TILE * MCF_TILE;
    g.tps = g.cps / MCF_TILE;
    g.n_chunks = (int64_t)n_streams * g.cps;
    g.n_tiles = (int64_t)n_streams * g.tps;
    g.q_cap = g.n_chunks / 8 + 4096;  // chunks whose entry differs from the speculative one: ~2 % expected
    int64_t o = 256;                  // header: status words, PlanLayout, queue counters
    auto region = [&o](int64_t bytes) {  // every region starts 256-byte aligned (vector loads/stores on the byte arrays)
        const int64_t at = o;
        o = (o + bytes + 255) / 256 * 256;
        return at;
    };
    g.off_start = region(g.n_chunks * 8);
    g.off_emit = region(g.n_chunks * 8);
    g.off_tsum = region(g.n_tiles * 8);
    g.off_tpre = region(g.n_tiles * 8);
    g.off_exit = region(g.n_chunks);
    g.off_gen = region(g.n_chunks);
    g.off_entry = region(g.n_chunks);
    g.off_send = region((int64_t)n_streams * 8);
    g.off_zsum = region(g.n_chunks * 8 * MCF_SUBS);
    g.off_queue = region(2 * g.q_cap * 8);  // two queues, used alternately by the fix-up rounds
    g.total = o;
    return g;
}

#define MCF_MAX_SPECS 8
struct SpecPack {
    int n;
    int affine;  // every set is affine in sum(Z): terminal values come from chunk sums + the two partial chunks
    mcf::GbmDerived g[MCF_MAX_SPECS];
};

// Host-side derivation of the per-step constants, in the same floating-point order as the reference's
// Python expressions (so a, b, disc are the very doubles NumPy would use).
static inline int derive_spec(const mcf_gbm_spec &s, int64_t n_steps, mcf::GbmDerived *g) {
    memset(g, 0, sizeof(*g));
    g->mode = s.mode;
    switch (s.mode) {
    case MCF_BS_EURO_CALL:
    case MCF_BS_EURO_PUT:
    case MCF_BS_AMER_CALL:
    case MCF_BS_AMER_PUT: {
        const double S0 = s.p[0], K = s.p[1], T = s.p[2], r = s.p[3], sigma = s.p[4];
        const double dt = T / (double)n_steps;
        g->a = (r - 0.5 * sigma * sigma) * dt;
        g->b = sigma * sqrt(dt);
        g->s0 = S0;
        g->logs0 = log(S0);
        g->K = K;
        g->disc = exp(-r * T);
        g->rdt = r * dt;
        return MCF_OK;
    }
    case MCF_PATH_TERMINAL: {
        const double S0 = s.p[0], r = s.p[1], sigma = s.p[2], T = s.p[3];
        const double dt = T / (double)n_steps;
        g->a = (r - 0.5 * sigma * sigma) * dt;
        g->b = sigma * sqrt(dt);
        g->s0 = S0;
        g->logs0 = log(S0);
        return MCF_OK;
    }
    case MCF_PORTFOLIO_GBM:
    case MCF_PORTFOLIO_LOG1P: {
        const double V0 = s.p[0], mu = s.p[1], sigma = s.p[2], dt = 1.0 / 252.0;
        g->a = s.mode == MCF_PORTFOLIO_GBM ? (mu - 0.5 * sigma * sigma) * dt : mu * dt;
        g->b = sigma * sqrt(dt);
        g->s0 = V0;
        return MCF_OK;
    }
    case MCF_NORMAL_SUM:
        return MCF_OK;
    case MCF_NORMAL_LOC_SCALE:
        g->a = s.p[0];
        g->b = s.p[1];
        return MCF_OK;
    default:
        return MCF_ERR_BAD_ARG;
    }
}


// Pi decomposition: every replication is cut into `parts` slices of 32 lanes x `ppt` points.
struct PiGeom {
    int64_t m, draws_per_sim, ppt, parts;
    int extra, weight;
};
static inline PiGeom pi_geom(int64_t n_points, int antithetic) {
    PiGeom g;
    g.m = antithetic ? n_points / 2 : n_points;
    g.extra = antithetic && (2 * g.m < n_points);
    g.weight = antithetic ? 2 : 1;
    g.draws_per_sim = 2 * g.m + (g.extra ? 2 : 0);
    g.ppt = (g.m + 31) / 32;
    g.ppt += g.ppt & 1;  // even: every lane then starts on the same Philox block phase (2 draws per point, 4 per block)
    if (g.ppt < 2) g.ppt = 2;
    if (g.ppt > 256) g.ppt = 256;  // long replications are cut into slices of 32*256 points
    g.parts = (g.m + 32 * g.ppt - 1) / (32 * g.ppt);
    if (g.parts < 1) g.parts = 1;
    return g;
}
