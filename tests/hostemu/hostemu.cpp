// hostemu.cpp -- TEST HARNESS ONLY.  Compiles the product's per-thread device functions
// (mcframework_b200/csrc/mcf_device.cuh, mcf_host.h) with g++ and replays each kernel's
// grid/thread decomposition serially, so the stream logic (jump-ahead, ziggurat chain scan,
// entry fix-up, locate, replay) can be checked against the oracle on a box without a GPU.
// The product never loads this library; on a GPU box the same checks run against the real
// kernels through the C ABI (tests/test_gpu_parity.py).
#include <stdint.h>
#include <stdlib.h>
#include <vector>

#include "../../mcframework_b200/csrc/mcf_host.h"
#include "../../mcframework_b200/csrc/zig_tables.h"

using namespace mcf;

static const unsigned long long h_ki[256] = {MCF_ZIG_KI_LIST};
static const unsigned long long h_wi[256] = {MCF_ZIG_WI_BITS_LIST};
static const unsigned long long h_fi[256] = {MCF_ZIG_FI_BITS_LIST};

static ZigTables host_tables() {
    static double wi[256], fi[256];
    static uint64_t ki[256];
    static ZigKW kws[512];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; i++) {
            ki[i] = h_ki[i];
            wi[i] = bits_to_double(h_wi[i]);
            fi[i] = bits_to_double(h_fi[i]);
        }
        for (unsigned e = 0; e < 512; e++) zig_fill_kws(kws, e, ki[e & 255], wi[e & 255]);
        init = true;
    }
    ZigTables t;
    t.ki = ki;
    t.wi = wi;
    t.fi = fi;
    t.kws = kws;
    return t;
}

static int stream_of(const mcf_stream *streams, int n_streams, int64_t sim, int64_t hint) {
    int64_t g = hint > 0 ? sim / hint : 0;
    if (g < n_streams) {
        int64_t f = streams[g].first_sim;
        if (f <= sim && sim < f + streams[g].n_sims) return (int)g;
    }
    return find_stream(streams, n_streams, sim);
}

template <int KIND>
static void pi_impl(const mcf_stream *streams, int n_streams, int64_t n_total, int64_t n_points, int antithetic,
                    int64_t hint, int64_t *hits) {
    const PiGeom pg = pi_geom(n_points, antithetic);
    for (int64_t i = 0; i < n_total; i++) hits[i] = 0;
    for (int64_t unit = 0; unit < n_total * pg.parts; unit++) {
        const int64_t sim = unit / pg.parts, part = unit - sim * pg.parts;
        const mcf_stream st = streams[stream_of(streams, n_streams, sim, hint)];
        const uint64_t base = (uint64_t)(sim - st.first_sim) * (uint64_t)pg.draws_per_sim;
        if (KIND == MCF_GEN_PHILOX && (st.buf_pos & 3u) == 0 && (pg.draws_per_sim % 4) == 0) {  // as pi_philox_stream_kernel
            const PhiloxKeys K = philox_keys_of(st);
            auto load = [&](uint64_t blk, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) { philox_block_keys(K, blk, a, b, c, d); };
            const uint64_t qbase = K.off + base;
            for (int lane = 0; lane < 32; lane++) {
                const int64_t p0 = (part * 32 + lane) * pg.ppt;
                int64_t cnt = pg.m - p0;
                cnt = cnt < 0 ? 0 : (cnt > pg.ppt ? pg.ppt : cnt);
                int64_t h = 0;
                if (cnt > 0) h = pi_count_hits_blocks(load, (qbase + 2ULL * (uint64_t)p0) >> 2, cnt) * pg.weight;
                if (pg.extra && part == 0 && lane == 0) {
                    const uint64_t q = qbase + 2ULL * (uint64_t)pg.m;
                    uint64_t a, b, c, d;
                    load(q >> 2, a, b, c, d);
                    h += (q & 2ULL) ? pi_inside(c, d) : pi_inside(a, b);
                }
                hits[sim] += h;
            }
            continue;
        }
        for (int lane = 0; lane < 32; lane++) {
            const int64_t p0 = (part * 32 + lane) * pg.ppt;
            int64_t cnt = pg.m - p0;
            cnt = cnt < 0 ? 0 : (cnt > pg.ppt ? pg.ppt : cnt);
            int64_t h = 0;
            if (cnt > 0) {
                Cursor<KIND> cur;
                cur.init(st, base + 2ULL * (uint64_t)p0);
                h = pi_count_hits(cur, cnt) * pg.weight;
            }
            if (pg.extra && part == 0 && lane == 0) {
                Cursor<KIND> cur;
                cur.init(st, base + 2ULL * (uint64_t)pg.m);
                h += pi_count_hits(cur, 1);
            }
            hits[sim] += h;
        }
    }
}

template <int KIND>
static int plan_impl(const mcf_stream *streams, int n_streams, int64_t max_sims, int64_t nps, double slack, int iters,
                     uint64_t *sim_start, uint64_t *stream_end, int64_t *n_regenerated, uint64_t *em_out,
                     uint8_t *ent_out, double *zs_out) {
    const ZigTables t = host_tables();
    const PlanGeom g = plan_geom(n_streams, max_sims, nps, slack);
    std::vector<uint64_t> sm(g.n_chunks), em(g.n_chunks), tsum(g.n_tiles), tpre(g.n_tiles);
    std::vector<uint8_t> ex(g.n_chunks), gen(g.n_chunks), ent(g.n_chunks);
    std::vector<double> zs(g.n_chunks * MCF_SUBS);
    int status = 0;
    int64_t n_fast_invalid = 0, n_fast_mismatch = 0;
    unsigned prev_exit_spec = 0;
    constexpr int RUN = KIND == MCF_GEN_PCG64 ? 8 : 1;
    // scan
    for (int64_t run = 0; run < g.n_chunks / RUN; run++) {
        const int64_t rps = g.cps / RUN, si = run / rps, c0 = (run - si * rps) * RUN;
        Cursor<KIND> cur;
        cur.init(streams[si], (uint64_t)c0 * MCF_CHUNK);
        uint32_t entry = 0;
        if (KIND == MCF_GEN_PHILOX && RUN == 1 && (streams[si].buf_pos & 3u) == 0) {  // as plan_scan_kernel
            const int64_t id = si * g.cps + c0;
            bool valid;
            const PhiloxKeys K = philox_keys_of(streams[si]);  // as plan_scan_philox_stream_kernel
            const uint64_t blk0 = (K.off + (uint64_t)c0 * MCF_CHUNK) >> 2;
            uint64_t slots[3 * MCF_PARK_CAP];
            ParkSlots park;
            park.a = slots;
            park.b = slots + MCF_PARK_CAP;
            park.c = slots + 2 * MCF_PARK_CAP;
            park.stride = 1;
            uint8_t slot_flags[MCF_PARK_CAP];
            park.f = slot_flags;
            uint64_t r0hi, r0lo, next_blk = blk0;
            mul64_full(MCF_PHILOX_M0, K.b0 + blk0, r0hi, r0lo);
            ChunkInfo ci = scan_chunk_fast2(
                [&](uint64_t blk, uint64_t &a, uint64_t &b, uint64_t &cc, uint64_t &d) {
                    if (K.fold_ok) {  // as the device kernel: consecutive blocks, running round-0 product
                        if (blk != next_blk) status = -1001;  // harness-only: the walk must ask for consecutive blocks
                        next_blk++;
                        philox_block_folded_r0(K, r0hi, r0lo, a, b, cc, d);
                        philox_r0_next(r0hi, r0lo);
                        return;
                    }
                    if (K.fold_ok)
                        philox_block_folded(K, blk, a, b, cc, d);
                    else
                        philox_block_keys(K, blk, a, b, cc, d);
                },
                blk0, t, park, &zs[id * MCF_SUBS], &valid, [&](unsigned exit_spec) {
                    // as plan_scan_philox_stream_kernel: the neighbouring lane's speculative exit, none for lane 0
                    const unsigned e = (c0 & 31) == 0 ? 0u : prev_exit_spec;
                    prev_exit_spec = exit_spec;
                    return e;
                },
                [](uint64_t, uint64_t) {},
                [&](uint64_t &a, uint64_t &b) {  // the next chunk's first two draws (the device reads its neighbour's)
                    uint64_t c2, d2;
                    if (K.fold_ok)
                        philox_block_folded(K, blk0 + 16, a, b, c2, d2);
                    else
                        philox_block_keys(K, blk0 + 16, a, b, c2, d2);
                }
                , [&](unsigned n_parked) { eval_parked_local(park, n_parked, t); }
                );
            if (valid) {  // the speculative walk must agree with the general one from the same entry
                double zref[MCF_SUBS];
                Cursor<KIND> c2;
                c2.init(streams[si], (uint64_t)c0 * MCF_CHUNK + ci.gen_entry);
                const ChunkInfo cr = scan_chunk(c2, (uint64_t)c0 * MCF_CHUNK, ci.gen_entry, t, zref);
                const uint64_t m = mask_from(ci.gen_entry);
                bool same = (cr.startmask & m) == (ci.startmask & m) && (cr.emitmask & m) == (ci.emitmask & m) &&
                            cr.exit == ci.exit;
                for (int q = 0; q < MCF_SUBS; q++) {
                    const double d = zref[q] - zs[id * MCF_SUBS + q];
                    same = same && (d < 1e-13 && d > -1e-13);
                }
                if (!same) n_fast_mismatch++;
            }
            sm[id] = ci.startmask;
            em[id] = ci.emitmask;
            ex[id] = (uint8_t)ci.exit;
            gen[id] = valid ? (uint8_t)ci.gen_entry : (uint8_t)MCF_GEN_ENTRY_INVALID;
            n_fast_invalid += valid ? 0 : 1;
            continue;
        }
        for (int c = 0; c < RUN; c++) {
            const int64_t id = si * g.cps + c0 + c;
            ChunkInfo ci = scan_chunk(cur, (uint64_t)(c0 + c) * MCF_CHUNK, entry, t, &zs[id * MCF_SUBS]);
            sm[id] = ci.startmask;
            em[id] = ci.emitmask;
            ex[id] = (uint8_t)ci.exit;
            gen[id] = (uint8_t)ci.gen_entry;
            if (ci.exit == MCF_EXIT_OVERFLOW) status = MCF_ERR_NOT_CONVERGED;
            entry = ci.exit;
        }
    }
    // resolve (threads of one launch see a consistent snapshot here; the GPU version may see a mix, which the
    // fix-point iteration tolerates)
    int64_t regen = 0;
    int changed_last = 0;
    for (int it = 0; it < iters; it++) {
        changed_last = 0;
        std::vector<uint8_t> ex_snapshot(ex);
        for (int64_t id = 0; id < g.n_chunks; id++) {
            const int64_t si = id / g.cps, c = id - si * g.cps;
            const uint32_t e = c == 0 ? 0u : ex_snapshot[id - 1];
            if (e != (uint32_t)gen[id]) {  // same rule as plan_resolve_kernel
                regen++;
                Cursor<KIND> cur;
                cur.init(streams[si], (uint64_t)c * MCF_CHUNK + e);
                ChunkInfo ci = scan_chunk(cur, (uint64_t)c * MCF_CHUNK, e, t, &zs[id * MCF_SUBS]);
                sm[id] = ci.startmask;
                em[id] = ci.emitmask;
                gen[id] = (uint8_t)ci.gen_entry;
                if (ci.exit != ex[id]) {
                    ex[id] = (uint8_t)ci.exit;
                    changed_last = 1;
                }
            }
            ent[id] = (uint8_t)e;
        }
    }
    if (changed_last && !status) status = MCF_ERR_NOT_CONVERGED;
    if (n_regenerated) *n_regenerated = regen;
    if (n_fast_mismatch) status = -1000;  // harness-only code: scan_chunk_fast2 disagreed with scan_chunk
    // tile count + per-stream scan
    for (int64_t tile = 0; tile < g.n_tiles; tile++) {
        uint64_t s = 0;
        for (int k = 0; k < MCF_TILE; k++) s += popc64(em[tile * MCF_TILE + k] & mask_from(ent[tile * MCF_TILE + k]));
        tsum[tile] = s;
    }
    for (int si = 0; si < n_streams; si++) {
        uint64_t acc = 0;
        for (int64_t k = 0; k < g.tps; k++) {
            tpre[si * g.tps + k] = acc;
            acc += tsum[si * g.tps + k];
        }
        if (acc < (uint64_t)streams[si].n_sims * (uint64_t)nps && !status) status = MCF_ERR_WINDOW_TOO_SMALL;
    }
    // locate
    for (int64_t tile = 0; tile < g.n_tiles; tile++) {
        const int64_t si = tile / g.tps;
        const mcf_stream st = streams[si];
        uint64_t E = tpre[tile];
        for (int k = 0; k < MCF_TILE; k++) {
            const int64_t id = tile * MCF_TILE + k;
            const int64_t chunk = id - si * g.cps;
            if (chunk == 0) sim_start[st.first_sim] = 0;
            const uint64_t m = mask_from(ent[id]);
            locate_in_chunk((uint64_t)chunk * MCF_CHUNK, sm[id] & m, em[id] & m, ex[id], E, (uint64_t)nps,
                            (uint64_t)st.n_sims, [&](uint64_t mm, uint64_t pos) {
                                if (mm < (uint64_t)st.n_sims)
                                    sim_start[st.first_sim + (int64_t)mm] = pos;
                                else if (stream_end)
                                    stream_end[si] = pos;
                            });
            E += popc64(em[id] & m);
        }
    }
    if (em_out)
        for (int64_t id = 0; id < g.n_chunks; id++) {
            em_out[id] = em[id];
            ent_out[id] = ent[id];
            for (int q = 0; q < MCF_SUBS; q++) zs_out[id * MCF_SUBS + q] = zs[id * MCF_SUBS + q];
        }
    return status;
}

template <int KIND>
static int terminal_impl(const mcf_stream *streams, int n_streams, int64_t n_total, int64_t n_steps, int64_t hint,
                         const uint64_t *sim_start, const mcf_gbm_spec *specs, int n_specs, double *out,
                         const uint64_t *em, const uint8_t *ent, const double *zs, int64_t cps,
                         const uint64_t *stream_end) {
    const ZigTables t = host_tables();
    PlanView pv;
    pv.emit = em;
    pv.entry = ent;
    pv.zsum = zs;
    pv.cps = cps;
    bool affine = true;
    for (int s = 0; s < n_specs; s++) {
        const int m = specs[s].mode;
        if (m == MCF_BS_AMER_CALL || m == MCF_BS_AMER_PUT || m == MCF_PORTFOLIO_LOG1P) affine = false;
    }
    SpecPack pack;
    pack.n = n_specs;
    for (int s = 0; s < n_specs; s++)
        if (derive_spec(specs[s], n_steps, &pack.g[s])) return MCF_ERR_BAD_ARG;
    for (int64_t i = 0; i < n_total; i++) {
        const int64_t si = stream_of(streams, n_streams, i, hint);
        const mcf_stream st = streams[si];
        Cursor<KIND> cur;
        cur.init(st, sim_start[i]);
        if (!em) {  // state-machine replay (no masks)
            if (n_specs != 1) return MCF_ERR_BAD_ARG;
            out[i] = run_terminal(cur, t, n_steps, pack.g[0]);
            continue;
        }
        const uint64_t start = sim_start[i];
        const uint64_t end = (i + 1 == st.first_sim + st.n_sims) ? stream_end[si] : sim_start[i + 1];
        if (!affine) {
            if (n_specs != 1) return MCF_ERR_BAD_ARG;
            out[i] = terminal_value([&](auto f) { replay_emits(cur, st, pv, si, start, end, t.wi, t.ki[0], f); }, pack.g[0]);
            continue;
        }
        double zsum;
        if (KIND == MCF_GEN_PHILOX && (st.buf_pos & 3u) == 0 && st.s[2] < (1ULL << 62) && n_steps >= (1 << MCF_SUB_SHIFT)) {
            // as gbm_terminal_affine_stream_kernel: the boundary at the end is cut by this replication, the one at the
            // start by its predecessor (same function, same arithmetic)
            const PhiloxKeys K = philox_keys_of(st);
            auto load = [&](uint64_t blk, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) { philox_block_folded(K, blk, a, b, c, d); };
            auto tail = [&](uint64_t p, uint64_t rabs) { return zig_tail_value<MCF_GEN_PHILOX>(st, p, rabs); };
            const BoundarySums be = boundary_sums(load, tail, pv, si, K.off, end, t.wi, t.ki[0]);
            const BoundarySums bs = boundary_sums(load, tail, pv, si, K.off, start, t.wi, t.ki[0]);
            zsum = zsum_from_boundaries(pv, si, start, end, bs.after, be.before);
        } else {
            zsum = replay_zsum(cur, st, pv, si, start, end, t.wi, t.ki[0]);
        }
        for (int s = 0; s < n_specs; s++) {
            const GbmDerived &g = pack.g[s];
            const double acc = xadd(xmul((double)n_steps, g.a), xmul(g.b, zsum));
            double v;
            if (g.mode == MCF_NORMAL_SUM)
                v = zsum;
            else if (g.mode == MCF_NORMAL_LOC_SCALE)
                v = acc;
            else if (g.mode == MCF_PATH_TERMINAL)
                v = exp(xadd(g.logs0, acc));
            else {
                const double stv = xmul(g.s0, exp(acc));
                if (g.mode == MCF_PORTFOLIO_GBM)
                    v = stv;
                else
                    v = xmul(g.mode == MCF_BS_EURO_CALL ? fmax(xsub(stv, g.K), 0.0) : fmax(xsub(g.K, stv), 0.0),
                             g.disc);
            }
            out[(int64_t)s * n_total + i] = v;
        }
    }
    return 0;
}

template <int KIND>
static void paths_impl(const mcf_stream *streams, int n_streams, int64_t n_total, int64_t n_steps, int64_t hint,
                       const uint64_t *sim_start, const double *p4, int time_major, double *paths) {
    const ZigTables t = host_tables();
    mcf_gbm_spec s;
    memset(&s, 0, sizeof(s));
    s.mode = MCF_PATH_TERMINAL;
    for (int k = 0; k < 4; k++) s.p[k] = p4[k];
    GbmDerived g;
    derive_spec(s, n_steps, &g);
    for (int64_t i = 0; i < n_total; i++) {
        const mcf_stream st = streams[stream_of(streams, n_streams, i, hint)];
        Cursor<KIND> cur;
        cur.init(st, sim_start[i]);
        const int64_t stride = time_major ? n_total : 1;
        double *p = paths + (time_major ? i : i * (n_steps + 1));
        double acc = 0.0;
        p[0] = exp(g.logs0);
        for (int64_t k = 1; k <= n_steps; k++) {
            acc = xadd(acc, xadd(g.a, xmul(g.b, zig_normal(cur, t))));
            p[k * stride] = exp(xadd(g.logs0, acc));
        }
    }
}

extern "C" {

void emu_pi_hits(uint32_t kind, const mcf_stream *streams, int n_streams, int64_t n_total, int64_t n_points,
                 int antithetic, int64_t hint, int64_t *hits) {
    if (kind == MCF_GEN_PCG64)
        pi_impl<MCF_GEN_PCG64>(streams, n_streams, n_total, n_points, antithetic, hint, hits);
    else
        pi_impl<MCF_GEN_PHILOX>(streams, n_streams, n_total, n_points, antithetic, hint, hits);
}

// raw draws of a block-aligned Philox stream through the folded block function (4*n_blocks values from block blk0)
int emu_philox_folded(const mcf_stream *st, uint64_t blk0, int64_t n_blocks, uint64_t *out) {
    const PhiloxKeys K = philox_keys_of(*st);
    if (!K.fold_ok) return -1;
    for (int64_t b = 0; b < n_blocks; b++)
        philox_block_folded(K, blk0 + (uint64_t)b, out[4 * b], out[4 * b + 1], out[4 * b + 2], out[4 * b + 3]);
    return 0;
}

int64_t emu_plan_chunks(int n_streams, int64_t max_sims, int64_t nps, double slack, int64_t *cps) {
    const PlanGeom g = plan_geom(n_streams, max_sims, nps, slack);
    *cps = g.cps;
    return g.n_chunks;
}

int emu_normal_plan(uint32_t kind, const mcf_stream *streams, int n_streams, int64_t max_sims, int64_t nps,
                    double slack, int iters, uint64_t *sim_start, uint64_t *stream_end, int64_t *n_regenerated,
                    uint64_t *em_out, uint8_t *ent_out, double *zs_out) {
    if (kind == MCF_GEN_PCG64)
        return plan_impl<MCF_GEN_PCG64>(streams, n_streams, max_sims, nps, slack, iters, sim_start, stream_end,
                                        n_regenerated, em_out, ent_out, zs_out);
    return plan_impl<MCF_GEN_PHILOX>(streams, n_streams, max_sims, nps, slack, iters, sim_start, stream_end,
                                     n_regenerated, em_out, ent_out, zs_out);
}

int emu_gbm_terminal(uint32_t kind, const mcf_stream *streams, int n_streams, int64_t n_total, int64_t n_steps,
                     int64_t hint, const uint64_t *sim_start, const mcf_gbm_spec *specs, int n_specs, double *out,
                     const uint64_t *em, const uint8_t *ent, const double *zs, int64_t cps,
                     const uint64_t *stream_end) {
    if (kind == MCF_GEN_PCG64)
        return terminal_impl<MCF_GEN_PCG64>(streams, n_streams, n_total, n_steps, hint, sim_start, specs, n_specs, out,
                                            em, ent, zs, cps, stream_end);
    return terminal_impl<MCF_GEN_PHILOX>(streams, n_streams, n_total, n_steps, hint, sim_start, specs, n_specs, out, em,
                                         ent, zs, cps, stream_end);
}

void emu_gbm_paths(uint32_t kind, const mcf_stream *streams, int n_streams, int64_t n_total, int64_t n_steps,
                   int64_t hint, const uint64_t *sim_start, const double *p4, int time_major, double *paths) {
    if (kind == MCF_GEN_PCG64)
        paths_impl<MCF_GEN_PCG64>(streams, n_streams, n_total, n_steps, hint, sim_start, p4, time_major, paths);
    else
        paths_impl<MCF_GEN_PHILOX>(streams, n_streams, n_total, n_steps, hint, sim_start, p4, time_major, paths);
}

// raw draws at an arbitrary stream position (checks Cursor::init jump-ahead)
void emu_raw64(uint32_t kind, const mcf_stream *st, uint64_t pos, int64_t n, uint64_t *out) {
    if (kind == MCF_GEN_PCG64) {
        Cursor<MCF_GEN_PCG64> cur;
        cur.init(*st, pos);
        for (int64_t i = 0; i < n; i++) out[i] = cur.next64();
    } else {
        Cursor<MCF_GEN_PHILOX> cur;
        cur.init(*st, pos);
        for (int64_t i = 0; i < n; i++) out[i] = cur.next64();
    }
}

void emu_normals(uint32_t kind, const mcf_stream *st, uint64_t pos, int64_t n, double *out, uint64_t *end_pos) {
    const ZigTables t = host_tables();
    if (kind == MCF_GEN_PCG64) {
        Cursor<MCF_GEN_PCG64> cur;
        cur.init(*st, pos);
        for (int64_t i = 0; i < n; i++) out[i] = zig_normal(cur, t);
        *end_pos = cur.pos;
    } else {
        Cursor<MCF_GEN_PHILOX> cur;
        cur.init(*st, pos);
        for (int64_t i = 0; i < n; i++) out[i] = zig_normal(cur, t);
        *end_pos = cur.pos;
    }
}

int emu_sizeof_stream(void) { return (int)sizeof(mcf_stream); }
}
