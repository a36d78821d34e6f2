// mcf_kernels.cu -- sm_100a kernels + C ABI (include/mcf_b200.h) for the replication hot path.
//
// Layout of the work (DESIGN.md has the long form):
//   * Pi: one warp per (replication, slice); every lane jumps to its own position in the
//     reference stream (PCG64 advance / Philox counter), counts hits in registers, the warp
//     reduces with shuffles and lane 0 adds one integer per slice.
//   * Normal-consuming replications: the ziggurat makes stream consumption data dependent,
//     so a planning pass finds where each replication starts (scan 64-draw chunks
//     speculatively -> fix chunk entries -> tile sums -> per-stream scan -> locate), then one
//     thread per replication replays its draws with all path state in registers.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>

#include "mcf_device.cuh"
#include "mcf_host.h"
#include "zig_tables.h"

using namespace mcf;

// ------------------------------------------------------------------ ziggurat tables in HBM
__device__ const unsigned long long g_zig_ki[256] = {MCF_ZIG_KI_LIST};
__device__ const unsigned long long g_zig_wi[256] = {MCF_ZIG_WI_BITS_LIST};
__device__ const unsigned long long g_zig_fi[256] = {MCF_ZIG_FI_BITS_LIST};

struct ZigShared {
    uint64_t ki[256];
    double wi[256];
    double fi[256];
};
// The scan's shared-memory tables as ONE 14 KB image in HBM, built at compile time from the generated lists: ki, wi,
// fi (what ZigShared holds) followed by the 512 (layer, sign) entries {ki, +-wi * 2^974} of ZigTables::kws.  A CTA
// copies it with 896 16-byte loads (3.5 per thread) instead of deriving its 512 entries from scratch -- the table
// set-up was 3 % of the scan's instructions.
struct ScanTableImage {
    unsigned long long w[768 + 1024];
};
constexpr ScanTableImage make_scan_table_image() {
    constexpr unsigned long long ki[256] = {MCF_ZIG_KI_LIST}, wi[256] = {MCF_ZIG_WI_BITS_LIST}, fi[256] = {MCF_ZIG_FI_BITS_LIST};
    ScanTableImage t{};
    for (int i = 0; i < 256; i++) {
        t.w[i] = ki[i];
        t.w[256 + i] = wi[i];
        t.w[512 + i] = fi[i];
    }
    for (int e = 0; e < 512; e++) {  // zig_fill_kws: ki | MCF_SCAN_BIAS (= 0), (e & 256 ? -wi : wi) * 2^974 (exponent + 974)
        t.w[768 + 2 * e] = ki[e & 255];
        t.w[768 + 2 * e + 1] = (wi[e & 255] + (974ULL << 52)) | ((e & 256) ? (1ULL << 63) : 0ULL);
    }
    return t;
}
__device__ const ScanTableImage g_scan_table_image = make_scan_table_image();
static_assert(MCF_SCAN_BIAS == 0ULL, "the table image assumes the subnormal (unbiased) form of rabs");
// the speculative scan adds the (layer, sign)-indexed table and the per-thread parking slots (256 threads per CTA)
struct alignas(16) ScanShared {
    ZigShared z;
    ZigKW kws[512];
    uint64_t park[3 * MCF_PARK_CAP * 256];
    uint64_t first[2 * 256];  // every thread's first two draws: the previous thread's look-ahead
    uint8_t park_flags[MCF_PARK_CAP * 256];  // flags of the evaluated parked draws
    uint8_t items[MCF_PARK_CAP * 256];       // per warp: (lane << 2 | slot) of its parked draws, compacted
};

__device__ __forceinline__ ZigTables load_zig_tables(ZigShared &sh) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sh.ki[i] = g_zig_ki[i];
        sh.wi[i] = __longlong_as_double((long long)g_zig_wi[i]);
        sh.fi[i] = __longlong_as_double((long long)g_zig_fi[i]);
    }
    __syncthreads();
    ZigTables t;
    t.ki = sh.ki;
    t.wi = sh.wi;
    t.fi = sh.fi;
    t.kws = nullptr;
    return t;
}

__device__ __forceinline__ ZigTables load_scan_tables(ScanShared &sh, ParkSlots &park) {
    static_assert(offsetof(ScanShared, kws) == sizeof(ZigShared) && sizeof(ZigShared) + sizeof(sh.kws) == sizeof(ScanTableImage),
                  "z and kws must mirror the table image");
    const uint4 *src = reinterpret_cast<const uint4 *>(g_scan_table_image.w);
    uint4 *dst = reinterpret_cast<uint4 *>(&sh);
#pragma unroll
    for (int i = 0; i < 4; i++) {  // 896 vectors, 256 threads per CTA (every scan kernel's launch bound)
        const int j = (int)threadIdx.x + 256 * i;
        if (j < (int)(sizeof(ScanTableImage) / 16)) dst[j] = __ldg(&src[j]);
    }
    __syncthreads();
    ZigTables t;
    t.ki = sh.z.ki;
    t.wi = sh.z.wi;
    t.fi = sh.z.fi;
    t.kws = sh.kws;
    park.a = sh.park + threadIdx.x;
    park.b = park.a + MCF_PARK_CAP * 256;
    park.c = park.b + MCF_PARK_CAP * 256;
    park.stride = 256;
    park.f = sh.park_flags + threadIdx.x;
    return t;
}

__device__ __forceinline__ int stream_of(const mcf_stream *streams, int n_streams, int64_t sim, int64_t hint) {
    int64_t g = hint > 0 ? sim / hint : 0;
    if (g < n_streams) {
        int64_t f = streams[g].first_sim;
        if (f <= sim && sim < f + streams[g].n_sims) return (int)g;
    }
    return find_stream(streams, n_streams, sim);
}

// code-generation choices of the Philox kernels (tools/ab_variants.sh measures the alternatives)
#ifndef MCF_PI_CHAIN
#define MCF_PI_CHAIN false      // carry-chain form of the round multiplications (mul64_full<true>) in the Pi kernel
#endif
#ifndef MCF_REPLAY_CHAIN
#define MCF_REPLAY_CHAIN false  // ... in the boundary-form replay
#endif
#ifndef MCF_SCAN_CTAS
#define MCF_SCAN_CTAS 4         // resident CTAs per SM of the planning scan (64 registers per thread)
#endif

// Block-aligned Philox streams are replayed by kernels whose round keys are LAUNCH CONSTANTS (kernel parameters, i.e.
// operands in the constant bank).  One launch takes a whole batch of streams: blockIdx.y (or a uniform loop index)
// selects the stream, so its keys still reach the instructions as uniform-register operands, and a parallel run of any
// width is a handful of launches instead of one per stream.  112 x (240 + 16) B stays below the 32 764-byte parameter limit.
// Streams long enough to fill the GPU by themselves keep one launch each (NB = 1: the keys are direct constant-bank
// operands; with an indexed batch the compiler re-loads part of them inside the hot loop, +3 % instructions).
#define MCF_KEY_BATCH 112
#define MCF_BATCH_BELOW_CTAS 2048  // streams with fewer CTAs of work than this are launched in batches
template <int NB>
struct PhiloxKeyBatch {
    PhiloxKeys k[NB];
};
template <int NB>
struct StreamSpanBatch {  // first replication and replication count of every stream of the batch
    int64_t first_sim[NB];
    int64_t n_sims[NB];
};
// `launch(tag, first_stream, count)` is called with tag.value == 1 (count == 1) or MCF_KEY_BATCH
template <int NB>
struct BatchTag {
    static constexpr int value = NB;
};
template <class Launch>
static void for_each_stream_batch(int64_t n_streams, bool batched, Launch launch) {
    if (!batched) {
        for (int64_t s = 0; s < n_streams; s++) launch(BatchTag<1>(), s, 1);
        return;
    }
    for (int64_t s0 = 0; s0 < n_streams; s0 += MCF_KEY_BATCH)
        launch(BatchTag<MCF_KEY_BATCH>(), s0, (int)(n_streams - s0 < MCF_KEY_BATCH ? n_streams - s0 : MCF_KEY_BATCH));
}
template <int NB>
static void fill_batch(const mcf_stream *h_streams, int64_t s0, int cnt, PhiloxKeyBatch<NB> &KB, StreamSpanBatch<NB> *SB,
                       int64_t *longest) {
    int64_t big = 0;
    for (int k = 0; k < cnt; k++) {
        KB.k[k] = philox_keys_of(h_streams[s0 + k]);
        if (SB) {
            SB->first_sim[k] = h_streams[s0 + k].first_sim;
            SB->n_sims[k] = h_streams[s0 + k].n_sims;
        }
        if (h_streams[s0 + k].n_sims > big) big = h_streams[s0 + k].n_sims;
    }
    if (longest) *longest = big;
}

// ===================================================================================== Pi
template <int KIND>
__global__ void __launch_bounds__(256) pi_kernel(const mcf_stream *__restrict__ streams, int n_streams, int64_t n_total,
                                                 int64_t m, int64_t draws_per_sim, int64_t ppt, int64_t parts,
                                                 int weight, int extra_point, int64_t hint,
                                                 unsigned long long *__restrict__ hits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_units = n_total * parts;
    for (int64_t unit = warp0; unit < n_units; unit += n_warps) {
        const int64_t sim = unit / parts, part = unit - sim * parts;
        const int si = stream_of(streams, n_streams, sim, hint);
        const mcf_stream st = streams[si];
        const uint64_t base = (uint64_t)(sim - st.first_sim) * (uint64_t)draws_per_sim;
        const int64_t p0 = (part * 32 + lane) * ppt;
        int64_t cnt = m - p0;
        cnt = cnt < 0 ? 0 : (cnt > ppt ? ppt : cnt);
        int64_t h = 0;
        if (cnt > 0) {
            Cursor<KIND> cur;
            cur.init(st, base + 2ULL * (uint64_t)p0);
            h = pi_count_hits(cur, cnt) * weight;
        }
        if (extra_point && part == 0 && lane == 0) {  // odd n_points with antithetic pairs (sims/pi.py:79-80)
            Cursor<KIND> cur;
            cur.init(st, base + 2ULL * (uint64_t)m);
            h += pi_count_hits(cur, 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
        if (lane == 0) atomicAdd(&hits[sim], (unsigned long long)h);
    }
}

// One launch per block-aligned Philox stream (round keys = launch constants), lanes on Philox block boundaries:
// requires draws_per_sim % 4 == 0 (ppt is even by construction).
template <int NB>
__global__ void __launch_bounds__(256) pi_philox_stream_kernel(const __grid_constant__ PhiloxKeyBatch<NB> KB,
                                                               const __grid_constant__ StreamSpanBatch<NB> SB, int64_t m,
                                                               int64_t draws_per_sim, int64_t ppt, int64_t parts, int weight,
                                                               int extra_point, unsigned long long *__restrict__ hits) {
    const unsigned sb = NB == 1 ? 0u : blockIdx.y;
    const PhiloxKeys &K = KB.k[sb];
    const int64_t first_sim = SB.first_sim[sb], n_sims = SB.n_sims[sb];
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_units = n_sims * parts;
    auto load = [&](uint64_t blk, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) { philox_block_folded<MCF_PI_CHAIN>(K, blk, a, b, c, d); };
    for (int64_t unit = warp0; unit < n_units; unit += n_warps) {
        const int64_t sim = unit / parts, part = unit - sim * parts;
        const uint64_t base = K.off + (uint64_t)sim * (uint64_t)draws_per_sim;  // multiple of 4
        const int64_t p0 = (part * 32 + lane) * ppt;
        int64_t cnt = m - p0;
        cnt = cnt < 0 ? 0 : (cnt > ppt ? ppt : cnt);
        int64_t h = 0;
        if (cnt > 0) {
            // a lane walks consecutive blocks: round 0's product M0 * (b0 + blk) is carried along (+ M0 per block);
            // measured 1.78e11 -> 1.85e11 samples/s
            const uint64_t blk0 = (base + 2ULL * (uint64_t)p0) >> 2;
            uint64_t r0hi, r0lo;
            mul64_full<MCF_PI_CHAIN>(MCF_PHILOX_M0, K.b0 + blk0, r0hi, r0lo);
            h = pi_count_hits_blocks(
                    [&](uint64_t, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
                        philox_block_folded_r0<MCF_PI_CHAIN>(K, r0hi, r0lo, a, b, c, d);
                        philox_r0_next(r0hi, r0lo);
                    },
                    blk0, cnt) * weight;
        }
        if (extra_point && part == 0 && lane == 0) {  // odd n_points with antithetic pairs (sims/pi.py:79-80)
            const uint64_t q = base + 2ULL * (uint64_t)m;  // even; the point is (v0, v1) or (v2, v3) of its block
            uint64_t a, b, c, d;
            load(q >> 2, a, b, c, d);
            h += (q & 2ULL) ? pi_inside(c, d) : pi_inside(a, b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
        if (lane == 0) atomicAdd(&hits[first_sim + sim], (unsigned long long)h);
    }
}

__global__ void pi_finalize_kernel(const unsigned long long *__restrict__ hits, int64_t n_total, double n_points,
                                   double *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) out[i] = __ddiv_rn(__dmul_rn(4.0, (double)(long long)hits[i]), n_points);
}

// ========================================================================== planning pass
__global__ void plan_init_kernel(int *hdr, int iters, PlanLayout layout) {
    if (threadIdx.x < MCF_HDR_WORDS) {
        hdr[threadIdx.x] = 0;
        hdr[MCF_QCOUNT_OFFSET / 4 + threadIdx.x] = 0;  // per-iteration work-queue counters
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        hdr[MCF_HDR_WORDS - 1] = iters;
        *reinterpret_cast<PlanLayout *>(reinterpret_cast<char *>(hdr) + MCF_LAYOUT_OFFSET) = layout;
    }
}

// one thread scans RUN consecutive chunks of one stream, carrying the ziggurat chain through
template <int KIND, int RUN>
__global__ void __launch_bounds__(256, 3) plan_scan_kernel(const mcf_stream *__restrict__ streams, int64_t cps,
                                                        int64_t n_runs, uint64_t *__restrict__ startmask,
                                                        uint64_t *__restrict__ emitmask, uint8_t *__restrict__ exit_,
                                                        uint8_t *__restrict__ gen, double *__restrict__ zsum, int *hdr) {
    __shared__ ScanShared sh;
    ParkSlots park;
    const ZigTables t = load_scan_tables(sh, park);
    const int64_t run = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (run >= n_runs) return;
    const int64_t runs_per_stream = cps / RUN;
    const int64_t si = run / runs_per_stream;
    const int64_t c0 = (run - si * runs_per_stream) * RUN;
    const mcf_stream st = streams[si];
    Cursor<KIND> cur;
    cur.init(st, (uint64_t)c0 * MCF_CHUNK);
    uint32_t entry = 0;
    if (KIND == MCF_GEN_PHILOX && RUN == 1 && (st.buf_pos & 3u) == 0) {
        // block-aligned Philox stream (every stream of a parallel run): SIMT-friendly walk; chunks it cannot finish
        // are tagged with an impossible generation entry so that the fix-up pass queues them for the general walk
        const int64_t id = si * cps + c0;
        bool valid;
        Cursor<MCF_GEN_PHILOX> &pc = *reinterpret_cast<Cursor<MCF_GEN_PHILOX> *>(&cur);
        const uint64_t blk0 = ((uint64_t)st.buf_pos + (uint64_t)c0 * MCF_CHUNK) >> 2;
        ChunkInfo ci = scan_chunk_fast2(
            [&](uint64_t blk, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
                pc.load_block(blk);
                a = pc.v0;
                b = pc.v1;
                c = pc.v2;
                d = pc.v3;
            },
            blk0, t, park, zsum + id * MCF_SUBS, &valid, [](unsigned) { return 0u; }, [](uint64_t, uint64_t) {},
            [&](uint64_t &a, uint64_t &b) {
                pc.load_block(blk0 + 16);
                a = pc.v0;
                b = pc.v1;
            }
            , [&](unsigned n_parked) { eval_parked_local(park, n_parked, t); }  // warps may be partial here
            );
        startmask[id] = ci.startmask;
        emitmask[id] = ci.emitmask;
        exit_[id] = (uint8_t)ci.exit;
        gen[id] = valid ? (uint8_t)0 : (uint8_t)MCF_GEN_ENTRY_INVALID;
        return;
    }
#pragma unroll 1
    for (int c = 0; c < RUN; c++) {
        const int64_t id = si * cps + c0 + c;
        ChunkInfo ci = scan_chunk(cur, (uint64_t)(c0 + c) * MCF_CHUNK, entry, t, zsum + id * MCF_SUBS);
        startmask[id] = ci.startmask;
        emitmask[id] = ci.emitmask;
        exit_[id] = (uint8_t)ci.exit;
        gen[id] = (uint8_t)ci.gen_entry;
        if (ci.exit == MCF_EXIT_OVERFLOW) atomicExch(&hdr[0], MCF_ERR_NOT_CONVERGED);
        entry = ci.exit;
    }
}

// One launch per stream: the stream's ten Philox round keys and base counter are kernel parameters, i.e. operands
// in the constant bank -- the key schedule costs no instruction and no register.
template <int NB>
__global__ void __launch_bounds__(256, MCF_SCAN_CTAS) plan_scan_philox_stream_kernel(const __grid_constant__ PhiloxKeyBatch<NB> KB, int64_t stream0,
                                                                         int64_t cps, uint64_t *__restrict__ startmask,
                                                                         uint64_t *__restrict__ emitmask,
                                                                         uint8_t *__restrict__ exit_,
                                                                         uint8_t *__restrict__ gen,
                                                                         double *__restrict__ zsum) {
    __shared__ ScanShared sh;
    ParkSlots park;
    const ZigTables t = load_scan_tables(sh, park);
    const unsigned sb = NB == 1 ? 0u : blockIdx.y;
    const PhiloxKeys &K = KB.k[sb];
    // (resident CTAs walking the tiles with a grid stride, tables set up once per CTA, measured 1 % SLOWER: 90.2 vs 89.3 ms)
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // cps is a multiple of 1024: whole CTAs are live
    if (c >= cps) return;
    const int64_t id = (stream0 + sb) * cps + c;
    bool valid;
    const uint64_t blk0 = (K.off + (uint64_t)c * MCF_CHUNK) >> 2;
    // the walk asks for consecutive blocks: round 0's product M0 * (b0 + blk) is carried along (+ M0 per block)
    uint64_t r0hi, r0lo;
    mul64_full<true>(MCF_PHILOX_M0, K.b0 + blk0, r0hi, r0lo);
    const ChunkInfo ci = scan_chunk_fast2(
        [&](uint64_t, uint64_t &a, uint64_t &b, uint64_t &cc, uint64_t &d) {
            philox_block_folded_r0<true>(K, r0hi, r0lo, a, b, cc, d);
            philox_r0_next(r0hi, r0lo);
        },
        blk0, t, park, zsum + id * MCF_SUBS, &valid,
        [&](unsigned exit_spec) {
            // the previous chunk of the stream is the neighbouring lane's; lane 0 cannot know it (entry 0, fixed up later)
            const unsigned prev = __shfl_up_sync(0xffffffffu, exit_spec, 1);
            return (threadIdx.x & 31) == 0 ? 0u : prev;
        },
        [&](uint64_t a, uint64_t b) {
            sh.first[threadIdx.x] = a;
            sh.first[256 + threadIdx.x] = b;
            __syncthreads();  // whole CTAs are live (cps is a multiple of 1024)
        },
        [&](uint64_t &a, uint64_t &b) {
            if (threadIdx.x == 255) {  // the CTA's last chunk: its successor belongs to another CTA
                uint64_t c2, d2;
                philox_block_folded<true>(K, blk0 + 16, a, b, c2, d2);
            } else {
                a = sh.first[threadIdx.x + 1];
                b = sh.first[256 + threadIdx.x + 1];
            }
        }
        // Warp-cooperative evaluation of the parked draws (whole warps are live here): an exclusive prefix of the lanes'
        // counts compacts (lane, slot) pairs into the warp's item list; item i is evaluated by lane i mod 32 from the
        // owner's shared-memory slot and written back in place.
        , [&](unsigned n_parked) {
            const unsigned lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
            unsigned incl = n_parked;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (unsigned)d) incl += v;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            uint8_t *items = sh.items + wbase * MCF_PARK_CAP;
            for (unsigned k = 0; k < n_parked; k++) items[incl - n_parked + k] = (uint8_t)((lane << 2) | k);
            __syncwarp();
            ParkSlots all;  // slot index = k * 256 + thread
            all.a = sh.park;
            all.b = all.a + MCF_PARK_CAP * 256;
            all.c = all.b + MCF_PARK_CAP * 256;
            all.f = sh.park_flags;
            all.stride = 256;
            for (unsigned i = lane; i < total; i += 32) {
                const unsigned it = items[i], slot = (it & 3u) * 256u + wbase + (it >> 2);
                park_eval_store(all, slot, park_eval(all.a[slot], all.b[slot], all.c[slot], t));
            }
            __syncwarp();
        }
        );
    startmask[id] = ci.startmask;
    emitmask[id] = ci.emitmask;
    exit_[id] = (uint8_t)ci.exit;
    gen[id] = valid ? (uint8_t)ci.gen_entry : (uint8_t)MCF_GEN_ENTRY_INVALID;
}

// Fix chunk entries: a chunk's true entry offset is its predecessor's exit.  Two kernels per iteration so that the
// (rare, 2 %) re-walks run DENSELY: with one thread per chunk, half of all warps would drag a 64-draw re-walk along
// for one or two lanes.
//   plan_entries_kernel : entry_true[id] = exit[id-1]; chunks generated with another entry go to a work queue
//   plan_rescan_kernel  : one thread per queued chunk re-walks it from its true entry
__device__ __forceinline__ void plan_enqueue(int64_t id, int64_t *__restrict__ queue, int64_t q_cap, int *hdr, int iter) {
    int *qcount = reinterpret_cast<int *>(reinterpret_cast<char *>(hdr) + MCF_QCOUNT_OFFSET) + iter;
    const int k = atomicAdd(qcount, 1);
    if (k < q_cap)
        queue[k] = id;
    else
        atomicExch(&hdr[0], MCF_ERR_NOT_CONVERGED);  // queue overflow: the caller re-plans with more iterations/slack
}

// First round: every chunk.  16 chunks per thread through 16-byte loads/stores of the byte arrays
// (cps is a multiple of 1024, so a group of 16 never straddles two streams).
__global__ void __launch_bounds__(256) plan_entries_kernel(int64_t cps, int64_t n_chunks, const uint8_t *__restrict__ exit_,
                                                           const uint8_t *__restrict__ gen,
                                                           uint8_t *__restrict__ entry_true, int64_t *__restrict__ queue,
                                                           int64_t q_cap, int *hdr) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t id0 = g * 16;
    const bool live = id0 < n_chunks;
    uint32_t diffw[4] = {0u, 0u, 0u, 0u};
    int mine = 0;
    if (live) {
        const uint4 ex4 = *reinterpret_cast<const uint4 *>(exit_ + id0);
        const uint4 gn4 = *reinterpret_cast<const uint4 *>(gen + id0);
        const uint8_t prev = (id0 % cps) == 0 ? (uint8_t)0 : exit_[id0 - 1];
        const uint32_t exw[4] = {ex4.x, ex4.y, ex4.z, ex4.w}, gnw[4] = {gn4.x, gn4.y, gn4.z, gn4.w};
        uint32_t enw[4];
        uint32_t carry = prev;  // exit of the chunk before the current one
#pragma unroll
        for (int w = 0; w < 4; w++) {
            enw[w] = (exw[w] << 8) | carry;  // entry bytes of this word = exit bytes shifted by one chunk
            carry = exw[w] >> 24;
            // Re-walk whenever the true entry differs from the one the chunk was generated with.  (The masks alone
            // would stay usable when the entry happens to be an attempt start of the speculative walk -- chains
            // coincide from there -- but the stored sums must not contain the attempts started before it.)
            const uint32_t d = enw[w] ^ gnw[w];
            // one flag bit per differing byte (bit 8*k)
            diffw[w] = ((d | (d >> 1) | (d >> 2) | (d >> 3) | (d >> 4) | (d >> 5) | (d >> 6) | (d >> 7)) & 0x01010101u);
            mine += __popc(diffw[w]);
        }
        *reinterpret_cast<uint4 *>(entry_true + id0) = make_uint4(enw[0], enw[1], enw[2], enw[3]);
    }
    // one atomic per warp: exclusive scan of the per-thread counts
    const int lane = threadIdx.x & 31;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = atomicAdd(reinterpret_cast<int *>(reinterpret_cast<char *>(hdr) + MCF_QCOUNT_OFFSET), total);
    base = __shfl_sync(0xffffffffu, base, 31);
    int k = base + incl - mine;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint32_t d = diffw[w];
        while (d) {
            const int byte = (__ffs(d) - 1) >> 3;
            if (k < q_cap)
                queue[k] = id0 + w * 4 + byte;
            else
                atomicExch(&hdr[0], MCF_ERR_NOT_CONVERGED);  // queue overflow: the caller re-plans
            k++;
            d &= d - 1;
        }
    }
}

// Later rounds: only the successors of the chunks re-walked in the previous round can have a new entry.
__global__ void __launch_bounds__(256) plan_entries_from_queue_kernel(int64_t cps, const uint8_t *__restrict__ exit_,
                                                                      const uint8_t *__restrict__ gen,
                                                                      uint8_t *__restrict__ entry_true,
                                                                      const int64_t *__restrict__ prev_queue,
                                                                      int64_t *__restrict__ queue, int64_t q_cap, int *hdr,
                                                                      int iter) {
    int count = *(reinterpret_cast<const int *>(reinterpret_cast<const char *>(hdr) + MCF_QCOUNT_OFFSET) + iter - 1);
    if (count > q_cap) count = (int)q_cap;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = prev_queue[k] + 1;
        if (id % cps == 0) continue;  // the re-walked chunk was the last of its stream
        const uint8_t e = exit_[id - 1];
        entry_true[id] = e;
        if (e != gen[id]) plan_enqueue(id, queue, q_cap, hdr, iter);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) plan_rescan_kernel(const mcf_stream *__restrict__ streams, int64_t cps,
                                                          const int64_t *__restrict__ queue, int64_t q_cap,
                                                          uint64_t *__restrict__ startmask, uint64_t *__restrict__ emitmask,
                                                          uint8_t *__restrict__ exit_, uint8_t *__restrict__ gen,
                                                          const uint8_t *__restrict__ entry_true,
                                                          double *__restrict__ zsum, int *hdr, int iter) {
    __shared__ ZigShared sh;
    const ZigTables t = load_zig_tables(sh);
    int count = *(reinterpret_cast<const int *>(reinterpret_cast<const char *>(hdr) + MCF_QCOUNT_OFFSET) + iter);
    if (count > q_cap) count = (int)q_cap;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = queue[k];
        const int64_t si = id / cps, c = id - si * cps;
        const uint32_t e = entry_true[id];
        const mcf_stream st = streams[si];
        Cursor<KIND> cur;
        cur.init(st, (uint64_t)c * MCF_CHUNK + e);
        const ChunkInfo ci = scan_chunk(cur, (uint64_t)c * MCF_CHUNK, e, t, zsum + id * MCF_SUBS);
        startmask[id] = ci.startmask;
        emitmask[id] = ci.emitmask;
        gen[id] = (uint8_t)ci.gen_entry;
        if (ci.exit != (uint32_t)exit_[id]) {
            exit_[id] = (uint8_t)ci.exit;
            atomicExch(&hdr[1 + iter], 1);
        }
        if (ci.exit == MCF_EXIT_OVERFLOW) atomicExch(&hdr[0], MCF_ERR_NOT_CONVERGED);
    }
}

__device__ __forceinline__ uint32_t block_reduce_sum(uint32_t v, uint32_t *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;  // valid in thread 0
}

// 256 threads x 4 chunks = one tile
__global__ void __launch_bounds__(256) plan_tile_count_kernel(const uint64_t *__restrict__ emitmask,
                                                              const uint8_t *__restrict__ entry_true,
                                                              uint64_t *__restrict__ tile_sum) {
    __shared__ uint32_t sh[8];
    const int64_t base = (int64_t)blockIdx.x * MCF_TILE + threadIdx.x * 4;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) c += (uint32_t)popc64(emitmask[base + k] & mask_from(entry_true[base + k]));
    uint32_t s = block_reduce_sum(c, sh);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = s;
}

// one CTA per stream: exclusive scan of that stream's tile sums; also checks the window was large enough
__global__ void __launch_bounds__(1024) plan_tile_scan_kernel(const mcf_stream *__restrict__ streams, int64_t tps,
                                                              int64_t nps, const uint64_t *__restrict__ tile_sum,
                                                              uint64_t *__restrict__ tile_pre, int *hdr) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry_sh;
    const int64_t si = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < tps; t0 += blockDim.x) {
        const int64_t i = t0 + threadIdx.x;
        uint64_t v = i < tps ? tile_sum[si * tps + i] : 0ULL, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint64_t w = warp_tot[lane], z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t y = __shfl_up_sync(0xffffffffu, z, o);
                if (lane >= o) z += y;
            }
            warp_tot[lane] = z - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        const uint64_t carry = carry_sh;
        const uint64_t incl = carry + warp_tot[wid] + x;
        if (i < tps) tile_pre[si * tps + i] = incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_sh = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint64_t need = (uint64_t)streams[si].n_sims * (uint64_t)nps;
        if (carry_sh < need) atomicExch(&hdr[0], MCF_ERR_WINDOW_TOO_SMALL);
    }
}

// one CTA per tile: exclusive scan of the chunk counts, then find replication boundaries
__global__ void __launch_bounds__(256) plan_locate_kernel(const mcf_stream *__restrict__ streams, int64_t cps,
                                                          int64_t tps, int64_t nps,
                                                          const uint64_t *__restrict__ startmask,
                                                          const uint64_t *__restrict__ emitmask,
                                                          const uint8_t *__restrict__ exit_,
                                                          const uint8_t *__restrict__ entry_true,
                                                          const uint64_t *__restrict__ tile_pre,
                                                          uint64_t *__restrict__ sim_start,
                                                          uint64_t *__restrict__ stream_end,
                                                          uint64_t *__restrict__ stream_end_ws) {
    __shared__ uint32_t warp_tot[8];
    const int64_t tile = blockIdx.x;
    const int64_t si = tile / tps;
    const int64_t base = tile * MCF_TILE + threadIdx.x * 4;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t es[4], ee[4];
    uint32_t cnt[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint64_t m = mask_from(entry_true[base + k]);
        es[k] = startmask[base + k] & m;
        ee[k] = emitmask[base + k] & m;
        cnt[k] = (uint32_t)popc64(ee[k]);
        tot += cnt[k];
    }
    uint32_t x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; w++) wbase += warp_tot[w];
    uint64_t E = tile_pre[tile] + wbase + (x - tot);
    const mcf_stream st = streams[si];
    const int64_t first_chunk = base - si * cps;
    if (first_chunk == 0) sim_start[st.first_sim] = 0ULL;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint64_t chunk_base = (uint64_t)(first_chunk + k) * MCF_CHUNK;
        locate_in_chunk(chunk_base, es[k], ee[k], exit_[base + k], E, (uint64_t)nps, (uint64_t)st.n_sims,
                        [&](uint64_t m, uint64_t pos) {
                            if (m < (uint64_t)st.n_sims) {
                                sim_start[st.first_sim + (int64_t)m] = pos;
                            } else {
                                stream_end_ws[si] = pos;
                                if (stream_end) stream_end[si] = pos;
                            }
                        });
        E += cnt[k];
    }
}

// ====================================================================== replication pass
// Only the wi table (2 KB) and ki[0] are needed: acceptance decisions come from the plan's emit masks.
struct WiShared {
    double wi[256];
};
__device__ __forceinline__ const double *load_wi_table(WiShared &sh) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh.wi[i] = __longlong_as_double((long long)g_zig_wi[i]);
    __syncthreads();
    return sh.wi;
}

// The host no longer waits for the plan's status before it queues the replay (one round trip per plan saved): a replay
// launched on a plan that FAILED (window too small, queue overflow: header word 0) must not follow its unusable
// positions -- every replay kernel returns at once and the caller, who reads the status after the replay, re-plans.
__device__ __forceinline__ bool plan_failed(const char *__restrict__ plan_ws) {
    return *reinterpret_cast<const volatile int *>(plan_ws) != 0;
}

struct ReplaySetup {
    mcf_stream st;
    PlanView pv;
    int64_t si;
    uint64_t start, end;
};

// where replication i lives: its stream, its draw range [start, end), the plan's masks
__device__ __forceinline__ ReplaySetup replay_setup(const mcf_stream *__restrict__ streams, int n_streams, int64_t hint,
                                                    const uint64_t *__restrict__ sim_start, const char *__restrict__ plan_ws,
                                                    int64_t i) {
    ReplaySetup r;
    const PlanLayout *lay = reinterpret_cast<const PlanLayout *>(plan_ws + MCF_LAYOUT_OFFSET);
    r.si = stream_of(streams, n_streams, i, hint);
    r.st = streams[r.si];
    r.pv.emit = reinterpret_cast<const uint64_t *>(plan_ws + lay->off_emit);
    r.pv.entry = reinterpret_cast<const uint8_t *>(plan_ws + lay->off_entry);
    r.pv.zsum = reinterpret_cast<const double *>(plan_ws + lay->off_zsum);
    r.pv.cps = lay->cps;
    r.start = sim_start[i];
    const bool last = (i + 1 == r.st.first_sim + r.st.n_sims);
    r.end = last ? reinterpret_cast<const uint64_t *>(plan_ws + lay->off_send)[r.si] : sim_start[i + 1];
    return r;
}

template <int KIND>
__global__ void __launch_bounds__(256, 3) gbm_terminal_kernel(const mcf_stream *__restrict__ streams, int n_streams,
                                                              int64_t n_total, int64_t hint,
                                                              const uint64_t *__restrict__ sim_start,
                                                              const char *__restrict__ plan_ws, SpecPack specs,
                                                              double n_steps_f, double *__restrict__ out) {
    __shared__ WiShared sh;
    if (plan_failed(plan_ws)) return;
    const double *wi = load_wi_table(sh);
    const uint64_t ki0 = g_zig_ki[0];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const ReplaySetup rs = replay_setup(streams, n_streams, hint, sim_start, plan_ws, i);
    Cursor<KIND> cur;
    if (!specs.affine) {  // path-dependent body (one parameter set): every normal, in order
        out[i] = terminal_value([&](auto f) { replay_emits(cur, rs.st, rs.pv, rs.si, rs.start, rs.end, wi, ki0, f); },
                                specs.g[0]);
        return;
    }
    // Terminal log-price is affine in sum(Z): n*a + b*sum(Z).  One pass serves every parameter set (Greeks on common
    // random numbers), and whole chunks of the stream contribute through the sums the planning pass stored.
    const double zsum = replay_zsum(cur, rs.st, rs.pv, rs.si, rs.start, rs.end, wi, ki0);
    for (int s = 0; s < specs.n; s++) {
        const GbmDerived &g = specs.g[s];
        const double acc = xadd(xmul(n_steps_f, g.a), xmul(g.b, zsum));
        double v;
        if (g.mode == MCF_NORMAL_SUM) {
            v = zsum;
        } else if (g.mode == MCF_NORMAL_LOC_SCALE) {
            v = acc;
        } else if (g.mode == MCF_PATH_TERMINAL) {
            v = exp(xadd(g.logs0, acc));
        } else {
            const double stv = xmul(g.s0, exp(acc));
            if (g.mode == MCF_PORTFOLIO_GBM)
                v = stv;
            else
                v = xmul(g.mode == MCF_BS_EURO_CALL ? fmax(xsub(stv, g.K), 0.0) : fmax(xsub(g.K, stv), 0.0), g.disc);
        }
        out[(int64_t)s * n_total + i] = v;
    }
}

// Affine bodies on block-aligned Philox streams, one launch per stream (round keys and base counter in the constant
// bank, rounds 0/1 folded): every thread regenerates only the shorter side of the boundary at the END of its
// replication (at most two Philox blocks) and takes the other side of the boundary at its START from its neighbour.
template <int NB>
__global__ void __launch_bounds__(256, 4)
    gbm_terminal_affine_stream_kernel(const __grid_constant__ PhiloxKeyBatch<NB> KB, const __grid_constant__ StreamSpanBatch<NB> SB,
                                      const mcf_stream *__restrict__ streams, int64_t stream0, int64_t n_total,
                                      const uint64_t *__restrict__ sim_start, const char *__restrict__ plan_ws,
                                      const __grid_constant__ SpecPack specs, double n_steps_f, double *__restrict__ out) {
    __shared__ WiShared sh;
    __shared__ double sh_after[8];
    const unsigned sb = NB == 1 ? 0u : blockIdx.y;
    const PhiloxKeys &K = KB.k[sb];
    const int64_t si = stream0 + sb, first_sim = SB.first_sim[sb], n_sims = SB.n_sims[sb];
    if ((int64_t)blockIdx.x * blockDim.x >= n_sims) return;  // the grid is sized for the longest stream of the batch
    if (plan_failed(plan_ws)) return;
    const double *wi = load_wi_table(sh);
    const uint64_t ki0 = g_zig_ki[0];
    const PlanLayout *lay = reinterpret_cast<const PlanLayout *>(plan_ws + MCF_LAYOUT_OFFSET);
    PlanView pv;
    pv.emit = reinterpret_cast<const uint64_t *>(plan_ws + lay->off_emit);
    pv.entry = reinterpret_cast<const uint8_t *>(plan_ws + lay->off_entry);
    pv.zsum = reinterpret_cast<const double *>(plan_ws + lay->off_zsum);
    pv.cps = lay->cps;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < n_sims;
    const int64_t i = first_sim + j;
    auto load = [&](uint64_t blk, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) { philox_block_folded<MCF_REPLAY_CHAIN>(K, blk, a, b, c, d); };
    auto tail = [&](uint64_t p, uint64_t rabs) { return zig_tail_value<MCF_GEN_PHILOX>(streams[si], p, rabs); };
    uint64_t start = 0, end = 0;
    BoundarySums be;
    be.before = 0.0;
    be.after = 0.0;
    if (active) {
        start = sim_start[i];
        end = (j + 1 == n_sims) ? reinterpret_cast<const uint64_t *>(plan_ws + lay->off_send)[si] : sim_start[i + 1];
        be = boundary_sums(load, tail, pv, si, K.off, end, wi, ki0);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double h = __shfl_up_sync(0xffffffffu, be.after, 1);
    if (lane == 31) sh_after[warp] = be.after;
    __syncthreads();
    if (lane == 0 && warp > 0) h = sh_after[warp - 1];
    if (!active) return;
    if (threadIdx.x == 0 || j == 0) h = boundary_sums(load, tail, pv, si, K.off, start, wi, ki0).after;
    const double zsum = zsum_from_boundaries(pv, si, start, end, h, be.before);
    for (int s = 0; s < specs.n; s++) {
        const GbmDerived &g = specs.g[s];
        const double acc = xadd(xmul(n_steps_f, g.a), xmul(g.b, zsum));
        double v;
        if (g.mode == MCF_NORMAL_SUM) {
            v = zsum;
        } else if (g.mode == MCF_NORMAL_LOC_SCALE) {
            v = acc;
        } else if (g.mode == MCF_PATH_TERMINAL) {
            v = exp(xadd(g.logs0, acc));
        } else {
            const double stv = xmul(g.s0, exp(acc));
            if (g.mode == MCF_PORTFOLIO_GBM)
                v = stv;
            else
                v = xmul(g.mode == MCF_BS_EURO_CALL ? fmax(xsub(stv, g.K), 0.0) : fmax(xsub(g.K, stv), 0.0), g.disc);
        }
        out[(int64_t)s * n_total + i] = v;
    }
}

template <int KIND>
__global__ void __launch_bounds__(256, 3) gbm_paths_kernel(const mcf_stream *__restrict__ streams, int n_streams,
                                                           int64_t n_total, int64_t n_steps, int64_t hint,
                                                           const uint64_t *__restrict__ sim_start,
                                                           const char *__restrict__ plan_ws, GbmDerived g, int time_major,
                                                           double *__restrict__ paths) {
    __shared__ WiShared sh;
    if (plan_failed(plan_ws)) return;
    const double *wi = load_wi_table(sh);
    const uint64_t ki0 = g_zig_ki[0];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const ReplaySetup rs = replay_setup(streams, n_streams, hint, sim_start, plan_ws, i);
    Cursor<KIND> cur;
    const int64_t stride = time_major ? n_total : 1;
    double *p = paths + (time_major ? i : i * (n_steps + 1));
    double acc = 0.0;
    const double a = g.a, b = g.b, logs0 = g.logs0;
    p[0] = exp(logs0);
    replay_emits(cur, rs.st, rs.pv, rs.si, rs.start, rs.end, wi, ki0, [&](double z) {
        acc = xadd(acc, xadd(a, xmul(b, z)));
        p += stride;
        *p = exp(xadd(logs0, acc));
    });
}

// Slices from segment sums: seg[i*k + j] = sum of the normals of steps (j*stride, (j+1)*stride] of path i (what the
// boundary-form replay returns when every path is cut into k pseudo-replications of `stride` steps).  The sums are
// path-major, the slices time-major: a warp moves 32*PPL paths x 16 segments through a shared-memory tile, so that the
// reads are 128-byte rows of one path and every store is ONE 128-bit vector per lane (PPL = 2 adjacent paths in fp64,
// 4 in fp32) -- 512 contiguous bytes per warp and slice.  Accumulation and exp() stay in fp64; only the store narrows.
template <typename OutT>
struct SliceVec;
template <>
struct SliceVec<double> {
    static constexpr int PPL = 2;
    static __device__ __forceinline__ void store(double *p, const double *v) { *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]); }
};
template <>
struct SliceVec<float> {
    static constexpr int PPL = 4;
    static __device__ __forceinline__ void store(float *p, const double *v) {
        *reinterpret_cast<float4 *>(p) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    }
};
#define MCF_SLICE_COLS 16
template <typename OutT>
__global__ void __launch_bounds__(64) gbm_slices_from_sums_kernel(const double *__restrict__ seg, int64_t n_total, int64_t k,
                                                                 GbmDerived g, double stride_f, int time_major,
                                                                 OutT *__restrict__ out) {
    constexpr int PPL = SliceVec<OutT>::PPL, ROWS = 32 * PPL;
    __shared__ double tile_all[2][ROWS][MCF_SLICE_COLS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double(*tile)[MCF_SLICE_COLS + 1] = tile_all[warp];
    const int64_t p0 = ((int64_t)blockIdx.x * 2 + warp) * ROWS;  // first path of this warp
    if (p0 >= n_total) return;
    const bool portfolio = g.mode == MCF_PORTFOLIO_GBM;
    const double drift = xmul(stride_f, g.a);
    const int64_t mine = p0 + (int64_t)lane * PPL;  // this lane's PPL adjacent paths
    // vector stores need PPL live, aligned paths: n_total % PPL == 0 is required for the time-major vector path
    const bool vec_ok = time_major && (n_total % PPL) == 0;
    double acc[PPL], v[PPL];
#pragma unroll
    for (int q = 0; q < PPL; q++) {
        acc[q] = 0.0;
        v[q] = portfolio ? g.s0 : exp(g.logs0);
    }
    auto put = [&](int64_t slice) {  // slice values of this lane's paths
        if (vec_ok) {
            if (mine < n_total) SliceVec<OutT>::store(out + slice * n_total + mine, v);
        } else {
#pragma unroll
            for (int q = 0; q < PPL; q++)
                if (mine + q < n_total) out[time_major ? slice * n_total + mine + q : (mine + q) * (k + 1) + slice] = (OutT)v[q];
        }
    };
    put(0);
    const int half = lane >> 4, col = lane & 15;
    for (int64_t j0 = 0; j0 < k; j0 += MCF_SLICE_COLS) {
        // rows of 16 consecutive segments of one path: two paths per load instruction, 128 contiguous bytes each
#pragma unroll 4
        for (int r = 0; r < ROWS; r += 2) {
            const int64_t path = p0 + r + half, j = j0 + col;
            tile[r + half][col] = (path < n_total && j < k) ? __ldg(&seg[path * k + j]) : 0.0;
        }
        __syncwarp();
        const int cols = (int)(k - j0 < MCF_SLICE_COLS ? k - j0 : MCF_SLICE_COLS);
        for (int t = 0; t < cols; t++) {
#pragma unroll
            for (int q = 0; q < PPL; q++) {
                acc[q] = xadd(acc[q], xadd(drift, xmul(g.b, tile[lane * PPL + q][t])));
                v[q] = portfolio ? xmul(g.s0, exp(acc[q])) : exp(xadd(g.logs0, acc[q]));
            }
            put(j0 + t + 1);
        }
        __syncwarp();
    }
}

// Stored path slices (BASELINE cfg 3): value at steps 0, stride, 2*stride, ... <= n_steps.  Time-major output
// (n_slices, n_total) gives fully coalesced 8-byte stores per warp; row-major (n_total, n_slices) is what a
// host caller reshapes most easily.  mode: MCF_PORTFOLIO_GBM -> V0*exp(sum), MCF_PATH_TERMINAL -> exp(log S0 + sum).
template <int KIND>
__global__ void __launch_bounds__(256, 3) gbm_slices_kernel(const mcf_stream *__restrict__ streams, int n_streams,
                                                            int64_t n_total, int64_t hint,
                                                            const uint64_t *__restrict__ sim_start,
                                                            const char *__restrict__ plan_ws, GbmDerived g,
                                                            int64_t slice_stride, int64_t n_slices, int time_major,
                                                            double *__restrict__ out) {
    __shared__ WiShared sh;
    if (plan_failed(plan_ws)) return;
    const double *wi = load_wi_table(sh);
    const uint64_t ki0 = g_zig_ki[0];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const ReplaySetup rs = replay_setup(streams, n_streams, hint, sim_start, plan_ws, i);
    Cursor<KIND> cur;
    const int64_t stride = time_major ? n_total : 1;
    double *p = out + (time_major ? i : i * n_slices);
    const bool path_mode = g.mode == MCF_PATH_TERMINAL;
    double acc = 0.0;
    const double a = g.a, b = g.b, logs0 = g.logs0, s0 = g.s0;
    p[0] = path_mode ? exp(logs0) : s0;
    int64_t until_store = slice_stride;
    replay_emits(cur, rs.st, rs.pv, rs.si, rs.start, rs.end, wi, ki0, [&](double z) {
        acc = xadd(acc, xadd(a, xmul(b, z)));
        if (--until_store == 0) {
            p += stride;
            *p = path_mode ? exp(xadd(logs0, acc)) : xmul(s0, exp(acc));
            until_store = slice_stride;
        }
    });
}

// ================================================================================= C ABI
static int g_last_cuda_error = 0;
#define MCF_STR2(x) #x
#define MCF_STR(x) MCF_STR2(x)
#define MCF_CUDA_OK(x)                                   \
    do {                                                 \
        cudaError_t e__ = (x);                           \
        if (e__ != cudaSuccess) {                        \
            g_last_cuda_error = (int)e__;                \
            return MCF_ERR_CUDA;                         \
        }                                                \
    } while (0)

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// One side stream and one fork/join event pair per device, created on first use and kept for the life of the process
// (creating and destroying them per plan showed up as occasional multi-millisecond stalls).  A plan holds them only
// between its fork and its join, both stream-ordered, so concurrent host threads on one device merely share the lane.
#include <mutex>
static bool side_stream_for_current_device(cudaStream_t *side, cudaEvent_t *fork, cudaEvent_t *join) {
    struct Lane {
        cudaStream_t s = nullptr;
        cudaEvent_t f = nullptr, j = nullptr;
        bool tried = false, ok = false;
    };
    static Lane lanes[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    std::lock_guard<std::mutex> lock(mu);
    Lane &l = lanes[dev];
    if (!l.tried) {
        l.tried = true;
        l.ok = cudaStreamCreateWithFlags(&l.s, cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&l.f, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&l.j, cudaEventDisableTiming) == cudaSuccess;
        if (!l.ok) (void)cudaGetLastError();
    }
    *side = l.s;
    *fork = l.f;
    *join = l.j;
    return l.ok;
}

template <int KIND>
static int normal_plan_impl(const mcf_stream *d_streams, const mcf_stream *h_streams, const PlanGeom &g, int64_t nps,
                            int32_t iters, char *ws, uint64_t *d_sim_start, uint64_t *d_stream_end, cudaStream_t cs) {
    int *hdr = (int *)ws;
    uint64_t *sm = (uint64_t *)(ws + g.off_start), *em = (uint64_t *)(ws + g.off_emit);
    uint64_t *tsum = (uint64_t *)(ws + g.off_tsum), *tpre = (uint64_t *)(ws + g.off_tpre);
    uint8_t *ex = (uint8_t *)(ws + g.off_exit), *gen = (uint8_t *)(ws + g.off_gen), *ent = (uint8_t *)(ws + g.off_entry);
    double *zs = (double *)(ws + g.off_zsum);
    PlanLayout layout;
    layout.magic = MCF_LAYOUT_MAGIC;
    layout.cps = g.cps;
    layout.off_emit = g.off_emit;
    layout.off_entry = g.off_entry;
    layout.off_send = g.off_send;
    layout.off_zsum = g.off_zsum;
    layout.n_streams = g.n_chunks / g.cps;
    plan_init_kernel<<<1, 32, 0, cs>>>(hdr, iters, layout);
    constexpr int RUN = KIND == MCF_GEN_PCG64 ? 8 : 1;  // PCG64 pays an O(log n) jump per thread
    const int64_t n_runs = g.n_chunks / RUN;
    const int64_t n_streams = g.n_chunks / g.cps;
    bool per_stream = KIND == MCF_GEN_PHILOX && h_streams != nullptr;
    for (int64_t si = 0; per_stream && si < n_streams; si++)
        per_stream = (h_streams[si].buf_pos & 3u) == 0 && h_streams[si].s[2] < (1ULL << 62);  // aligned, foldable counter
    if (per_stream) {
        // Streams that get a launch of their own alternate between the caller's stream and a side stream, so that the tail
        // of one launch (the last partial wave of its ~42) overlaps the head of the next (-0.8 % per plan of 1e8 x 252).
        const bool batched = g.cps / 256 < MCF_BATCH_BELOW_CTAS;
        cudaStream_t side = nullptr;
        cudaEvent_t fork = nullptr, join = nullptr;
        const bool two = !batched && n_streams > 1 && side_stream_for_current_device(&side, &fork, &join);
        if (two) {
            cudaEventRecord(fork, cs);  // plan_init_kernel precedes every scan
            cudaStreamWaitEvent(side, fork, 0);
        }
        int nb = 0;
        for_each_stream_batch(n_streams, batched, [&](auto tag, int64_t s0, int cnt) {
            constexpr int NB = decltype(tag)::value;
            PhiloxKeyBatch<NB> KB;  // host staging of the parameter block (the launch copies it)
            fill_batch<NB>(h_streams, s0, cnt, KB, nullptr, nullptr);
            const unsigned gx = grid_for(g.cps, 256);
            plan_scan_philox_stream_kernel<NB><<<dim3(gx, cnt), 256, 0, (two && (nb++ & 1)) ? side : cs>>>(
                KB, s0, g.cps, sm, em, ex, gen, zs);
        });
        if (two) {
            cudaEventRecord(join, side);
            cudaStreamWaitEvent(cs, join, 0);
        }
    } else {
        plan_scan_kernel<KIND, RUN><<<grid_for(n_runs, 256), 256, 0, cs>>>(d_streams, g.cps, n_runs, sm, em, ex, gen, zs,
                                                                          hdr);
    }
    MCF_CUDA_OK(cudaGetLastError());
    // two work queues, used alternately: round k reads the chunks re-walked in round k-1
    int64_t *queues[2] = {(int64_t *)(ws + g.off_queue), (int64_t *)(ws + g.off_queue) + g.q_cap};
    for (int it = 0; it < iters; it++) {
        int64_t *queue = queues[it & 1];
        // the first round carries ~2 % of the chunks, later ones almost nothing: size the grids accordingly
        const int64_t expect = it == 0 ? g.q_cap : g.q_cap / 64 + 256;
        if (it == 0)
            plan_entries_kernel<<<grid_for(g.n_chunks / 16, 256), 256, 0, cs>>>(g.cps, g.n_chunks, ex, gen, ent, queue,
                                                                                g.q_cap, hdr);
        else
            plan_entries_from_queue_kernel<<<grid_for(expect, 256), 256, 0, cs>>>(g.cps, ex, gen, ent, queues[(it - 1) & 1],
                                                                                  queue, g.q_cap, hdr, it);
        MCF_CUDA_OK(cudaGetLastError());
        plan_rescan_kernel<KIND><<<grid_for(expect, 256), 256, 0, cs>>>(d_streams, g.cps, queue, g.q_cap, sm, em, ex, gen,
                                                                        ent, zs, hdr, it);
        MCF_CUDA_OK(cudaGetLastError());
    }
    plan_tile_count_kernel<<<(unsigned)g.n_tiles, 256, 0, cs>>>(em, ent, tsum);
    MCF_CUDA_OK(cudaGetLastError());
    plan_tile_scan_kernel<<<(unsigned)(g.n_chunks / g.cps), 1024, 0, cs>>>(d_streams, g.tps, nps, tsum, tpre, hdr);
    MCF_CUDA_OK(cudaGetLastError());
    plan_locate_kernel<<<(unsigned)g.n_tiles, 256, 0, cs>>>(d_streams, g.cps, g.tps, nps, sm, em, ex, ent, tpre,
                                                           d_sim_start, d_stream_end, (uint64_t *)(ws + g.off_send));
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

// Issue-rate probe (the roofline denominator of the RNG kernels, measured on the device the benchmark runs on): eight
// independent chains of one instruction per thread, 4 CTAs x 256 threads per SM, so the loop is bound by the issue rate
// of that instruction, not by its latency.  kind 0: IMAD.WIDE.U32 (32x32 -> 64 with a 64-bit addend), 1: IMAD (low 32
// bits), 2: LOP3.  The same loop as tools/probe_pipes.cu.
template <int KIND>
__global__ void __launch_bounds__(256) issue_rate_probe_kernel(uint64_t *out, uint32_t m0, uint32_t m1, int iters) {
    uint64_t a[8];
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = threadIdx.x * 977u + k;
        x[k] = threadIdx.x * 31u + k;
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (KIND == 0)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(x[k]), "r"(m0));
            else if (KIND == 1)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(m0), "r"(m1));
            else
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m0), "r"(m1));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k] + x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int mcf_probe_issue_rate(int32_t kind, double *warp_inst_per_s, void *cuda_stream) {
    if (!warp_inst_per_s || kind < 0 || kind > 2) return MCF_ERR_BAD_ARG;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    int dev = 0, sms = 0;
    MCF_CUDA_OK(cudaGetDevice(&dev));
    MCF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 4, iters = 16384;
    uint64_t *out = nullptr;
    MCF_CUDA_OK(cudaMalloc(&out, sizeof(uint64_t) * (size_t)blocks * 256));
    cudaEvent_t e0, e1;
    MCF_CUDA_OK(cudaEventCreate(&e0));
    MCF_CUDA_OK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 6; r++) {  // the first repetition warms up
        cudaEventRecord(e0, cs);
        if (kind == 0)
            issue_rate_probe_kernel<0><<<blocks, 256, 0, cs>>>(out, 0xE14C6C93u, 0xD2E7470Eu, iters);
        else if (kind == 1)
            issue_rate_probe_kernel<1><<<blocks, 256, 0, cs>>>(out, 0xE14C6C93u, 0xD2E7470Eu, iters);
        else
            issue_rate_probe_kernel<2><<<blocks, 256, 0, cs>>>(out, 0xE14C6C93u, 0xD2E7470Eu, iters);
        cudaEventRecord(e1, cs);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    MCF_CUDA_OK(cudaGetLastError());
    *warp_inst_per_s = (double)blocks * 8.0 * (double)iters * 8.0 / ((double)best * 1e-3);  // 8 warps per CTA, 8 chains
    return MCF_OK;
}

extern "C" {

int mcf_abi_version(void) { return MCF_ABI_VERSION; }

const char *mcf_build_info(void) {
    return "mcf_b200: sm_100a, CUDA " MCF_STR(CUDART_VERSION) ", built " __DATE__;
}

int mcf_last_cuda_error(void) { return g_last_cuda_error; }

int mcf_pi_hits(uint32_t gen_kind, const mcf_stream *d_streams, const mcf_stream *h_streams, int32_t n_streams,
                  int64_t n_total, int64_t n_points, int32_t antithetic, int64_t sims_per_stream_hint, int64_t *d_hits,
                  double *d_out, void *cuda_stream) {
    if (!d_streams || n_streams <= 0 || n_total <= 0 || n_points <= 0 || !d_hits) return MCF_ERR_BAD_ARG;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    const PiGeom pg = pi_geom(n_points, antithetic);
    MCF_CUDA_OK(cudaMemsetAsync(d_hits, 0, sizeof(int64_t) * (size_t)n_total, cs));
    const int64_t n_units = n_total * pg.parts;
    int64_t blocks = (n_units + 7) / 8;  // 8 warps per CTA
    const int64_t max_blocks = 148LL * 8 * 64;
    if (blocks > max_blocks) blocks = max_blocks;
    bool per_stream = gen_kind == MCF_GEN_PHILOX && h_streams != nullptr && (pg.draws_per_sim % 4) == 0;
    for (int si = 0; per_stream && si < n_streams; si++)
        per_stream = (h_streams[si].buf_pos & 3u) == 0 && h_streams[si].s[2] < (1ULL << 62);  // aligned, foldable counter
    if (per_stream) {
        int64_t biggest = 0;
        for (int si = 0; si < n_streams; si++) biggest = h_streams[si].n_sims > biggest ? h_streams[si].n_sims : biggest;
        for_each_stream_batch(n_streams, (biggest * pg.parts + 7) / 8 < MCF_BATCH_BELOW_CTAS, [&](auto tag, int64_t s0, int cnt) {
            constexpr int NB = decltype(tag)::value;
            PhiloxKeyBatch<NB> KB;
            StreamSpanBatch<NB> SB;
            int64_t longest = 0;
            fill_batch<NB>(h_streams, s0, cnt, KB, &SB, &longest);
            int64_t b = (longest * pg.parts + 7) / 8;
            if (b > max_blocks / cnt + 1) b = max_blocks / cnt + 1;  // grid-stride over the (replication, slice) units
            if (b < 1) return;
            pi_philox_stream_kernel<NB><<<dim3((unsigned)b, cnt), 256, 0, cs>>>(KB, SB, pg.m, pg.draws_per_sim, pg.ppt, pg.parts,
                                                                                pg.weight, pg.extra, (unsigned long long *)d_hits);
        });
    } else if (gen_kind == MCF_GEN_PCG64)
        pi_kernel<MCF_GEN_PCG64><<<(unsigned)blocks, 256, 0, cs>>>(
            d_streams, n_streams, n_total, pg.m, pg.draws_per_sim, pg.ppt, pg.parts, pg.weight, pg.extra,
            sims_per_stream_hint, (unsigned long long *)d_hits);
    else
        pi_kernel<MCF_GEN_PHILOX><<<(unsigned)blocks, 256, 0, cs>>>(
            d_streams, n_streams, n_total, pg.m, pg.draws_per_sim, pg.ppt, pg.parts, pg.weight, pg.extra,
            sims_per_stream_hint, (unsigned long long *)d_hits);
    MCF_CUDA_OK(cudaGetLastError());
    if (d_out) {
        pi_finalize_kernel<<<grid_for(n_total, 256), 256, 0, cs>>>((const unsigned long long *)d_hits, n_total,
                                                                   (double)n_points, d_out);
        MCF_CUDA_OK(cudaGetLastError());
    }
    return MCF_OK;
}


int64_t mcf_normal_plan_workspace_bytes(int32_t n_streams, int64_t max_sims_per_stream, int64_t normals_per_sim,
                                        double slack) {
    if (n_streams <= 0 || max_sims_per_stream <= 0 || normals_per_sim <= 0 || slack < 0) return MCF_ERR_BAD_ARG;
    return plan_geom(n_streams, max_sims_per_stream, normals_per_sim, slack).total;
}

int mcf_normal_plan(uint32_t gen_kind, const mcf_stream *d_streams, const mcf_stream *h_streams, int32_t n_streams,
                      int64_t n_total, int64_t max_sims_per_stream, int64_t normals_per_sim, double slack,
                      int32_t resolve_iters, void *d_workspace, int64_t workspace_bytes, uint64_t *d_sim_start,
                      uint64_t *d_stream_end, void *cuda_stream) {
    if (!d_streams || n_streams <= 0 || n_total <= 0 || max_sims_per_stream <= 0 || normals_per_sim <= 0 ||
        !d_workspace || !d_sim_start || resolve_iters < 1 || resolve_iters > MCF_HDR_WORDS - 3)
        return MCF_ERR_BAD_ARG;
    const PlanGeom g = plan_geom(n_streams, max_sims_per_stream, normals_per_sim, slack);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    if (gen_kind == MCF_GEN_PCG64)
        return normal_plan_impl<MCF_GEN_PCG64>(d_streams, h_streams, g, normals_per_sim, resolve_iters, (char *)d_workspace,
                                               d_sim_start, d_stream_end, cs);
    return normal_plan_impl<MCF_GEN_PHILOX>(d_streams, h_streams, g, normals_per_sim, resolve_iters, (char *)d_workspace,
                                            d_sim_start, d_stream_end, cs);
}


int mcf_workspace_status(const void *d_workspace, void *cuda_stream) {
    int hdr[MCF_HDR_WORDS];
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    MCF_CUDA_OK(cudaMemcpyAsync(hdr, d_workspace, sizeof(hdr), cudaMemcpyDeviceToHost, cs));
    MCF_CUDA_OK(cudaStreamSynchronize(cs));
    if (hdr[0] != 0) return hdr[0];
    const int iters = hdr[MCF_HDR_WORDS - 1];
    if (iters >= 1 && hdr[1 + iters - 1] != 0) return MCF_ERR_NOT_CONVERGED;
    return MCF_OK;
}

}  // extern "C"

extern "C" {

int mcf_gbm_terminal(uint32_t gen_kind, const mcf_stream *d_streams, const mcf_stream *h_streams, int32_t n_streams,
                     int64_t n_total, int64_t n_steps, int64_t sims_per_stream_hint, const uint64_t *d_sim_start,
                     const void *d_plan_workspace, const mcf_gbm_spec *h_specs, int32_t n_specs, double *d_out,
                     void *cuda_stream) {
    if (!d_streams || n_streams <= 0 || n_total <= 0 || n_steps <= 0 || !d_sim_start || !d_plan_workspace || !h_specs ||
        n_specs < 1 || n_specs > MCF_MAX_SPECS || !d_out)
        return MCF_ERR_BAD_ARG;
    SpecPack pack;
    pack.n = n_specs;
    pack.affine = 1;
    for (int s = 0; s < n_specs; s++) {
        int rc = derive_spec(h_specs[s], n_steps, &pack.g[s]);
        if (rc) return rc;
        const int m = h_specs[s].mode;
        const bool path_dependent = (m == MCF_BS_AMER_CALL || m == MCF_BS_AMER_PUT || m == MCF_PORTFOLIO_LOG1P);
        if (path_dependent) pack.affine = 0;
        if (n_specs > 1 && path_dependent) return MCF_ERR_BAD_ARG;  // path-dependent bodies take one set per launch
    }
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    // affine bodies of at least one sub-chunk per replication on aligned, foldable Philox streams: boundary form
    bool per_stream = gen_kind == MCF_GEN_PHILOX && pack.affine && n_steps >= (1 << MCF_SUB_SHIFT) && h_streams != nullptr;
    for (int si = 0; per_stream && si < n_streams; si++)
        per_stream = (h_streams[si].buf_pos & 3u) == 0 && h_streams[si].s[2] < (1ULL << 62);
    if (per_stream) {
        int64_t biggest = 0;
        for (int si = 0; si < n_streams; si++) biggest = h_streams[si].n_sims > biggest ? h_streams[si].n_sims : biggest;
        for_each_stream_batch(n_streams, grid_for(biggest, 256) < MCF_BATCH_BELOW_CTAS, [&](auto tag, int64_t s0, int cnt) {
            constexpr int NB = decltype(tag)::value;
            PhiloxKeyBatch<NB> KB;
            StreamSpanBatch<NB> SB;
            int64_t longest = 0;
            fill_batch<NB>(h_streams, s0, cnt, KB, &SB, &longest);
            if (longest <= 0) return;
            gbm_terminal_affine_stream_kernel<NB><<<dim3(grid_for(longest, 256), cnt), 256, 0, cs>>>(
                KB, SB, d_streams, s0, n_total, d_sim_start, (const char *)d_plan_workspace, pack, (double)n_steps, d_out);
        });
        MCF_CUDA_OK(cudaGetLastError());
        return MCF_OK;
    }
    if (gen_kind == MCF_GEN_PCG64)
        gbm_terminal_kernel<MCF_GEN_PCG64><<<grid_for(n_total, 256), 256, 0, cs>>>(
            d_streams, n_streams, n_total, sims_per_stream_hint, d_sim_start, (const char *)d_plan_workspace, pack,
            (double)n_steps, d_out);
    else
        gbm_terminal_kernel<MCF_GEN_PHILOX><<<grid_for(n_total, 256), 256, 0, cs>>>(
            d_streams, n_streams, n_total, sims_per_stream_hint, d_sim_start, (const char *)d_plan_workspace, pack,
            (double)n_steps, d_out);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}


int mcf_gbm_paths(uint32_t gen_kind, const mcf_stream *d_streams, int32_t n_streams, int64_t n_total, int64_t n_steps,
                    int64_t sims_per_stream_hint, const uint64_t *d_sim_start, const void *d_plan_workspace,
                    const double *h_p4, int32_t time_major, double *d_paths, void *cuda_stream) {
    if (!d_streams || n_streams <= 0 || n_total <= 0 || n_steps <= 0 || !d_sim_start || !d_plan_workspace || !h_p4 ||
        !d_paths)
        return MCF_ERR_BAD_ARG;
    mcf_gbm_spec s;
    memset(&s, 0, sizeof(s));
    s.mode = MCF_PATH_TERMINAL;
    for (int k = 0; k < 4; k++) s.p[k] = h_p4[k];
    GbmDerived g;
    const int rc = derive_spec(s, n_steps, &g);
    if (rc) return rc;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    if (gen_kind == MCF_GEN_PCG64)
        gbm_paths_kernel<MCF_GEN_PCG64><<<grid_for(n_total, 256), 256, 0, cs>>>(
            d_streams, n_streams, n_total, n_steps, sims_per_stream_hint, d_sim_start, (const char *)d_plan_workspace, g,
            time_major, d_paths);
    else
        gbm_paths_kernel<MCF_GEN_PHILOX><<<grid_for(n_total, 256), 256, 0, cs>>>(
            d_streams, n_streams, n_total, n_steps, sims_per_stream_hint, d_sim_start, (const char *)d_plan_workspace, g,
            time_major, d_paths);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}


}  // extern "C"

// Re-run only the locate step of a finished plan for another normals-per-replication count: `d_streams_new` describes the
// same streams cut into n_sims*old_nps/new_nps pseudo-replications (first_sim/n_sims scaled by the caller).  The masks,
// entries and tile prefix sums of the workspace are reused; only sim_start is recomputed (one pass over the masks).
extern "C" int mcf_normal_relocate(const mcf_stream *d_streams_new, int32_t n_streams, int64_t max_sims_per_stream,
                                   int64_t normals_per_sim, double slack, int64_t new_normals_per_sim, void *d_workspace,
                                   uint64_t *d_sim_start_new, void *cuda_stream) {
    if (!d_streams_new || n_streams <= 0 || max_sims_per_stream <= 0 || normals_per_sim <= 0 || new_normals_per_sim <= 0 ||
        !d_workspace || !d_sim_start_new)
        return MCF_ERR_BAD_ARG;
    const PlanGeom g = plan_geom(n_streams, max_sims_per_stream, normals_per_sim, slack);
    char *ws = (char *)d_workspace;
    plan_locate_kernel<<<(unsigned)g.n_tiles, 256, 0, (cudaStream_t)cuda_stream>>>(
        d_streams_new, g.cps, g.tps, new_normals_per_sim, (const uint64_t *)(ws + g.off_start),
        (const uint64_t *)(ws + g.off_emit), (const uint8_t *)(ws + g.off_exit), (const uint8_t *)(ws + g.off_entry),
        (const uint64_t *)(ws + g.off_tpre), d_sim_start_new, nullptr, (uint64_t *)(ws + g.off_send));
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

extern "C" int mcf_gbm_slices_from_sums(const double *d_segment_sums, int64_t n_total, int64_t n_segments, int64_t segment_steps,
                                        int64_t n_steps, const mcf_gbm_spec *h_spec, int32_t time_major, int32_t out_f32,
                                        void *d_out, void *cuda_stream) {
    if (!d_segment_sums || n_total <= 0 || n_segments <= 0 || segment_steps <= 0 || n_steps != n_segments * segment_steps ||
        !h_spec || !d_out || (h_spec->mode != MCF_PORTFOLIO_GBM && h_spec->mode != MCF_PATH_TERMINAL))
        return MCF_ERR_BAD_ARG;
    GbmDerived g;
    int rc = derive_spec(*h_spec, n_steps, &g);
    if (rc) return rc;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    if (out_f32)
        gbm_slices_from_sums_kernel<float><<<grid_for(n_total, 2 * 32 * SliceVec<float>::PPL), 64, 0, cs>>>(
            d_segment_sums, n_total, n_segments, g, (double)segment_steps, time_major, (float *)d_out);
    else
        gbm_slices_from_sums_kernel<double><<<grid_for(n_total, 2 * 32 * SliceVec<double>::PPL), 64, 0, cs>>>(
            d_segment_sums, n_total, n_segments, g, (double)segment_steps, time_major, (double *)d_out);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

extern "C" int mcf_gbm_slices(uint32_t gen_kind, const mcf_stream *d_streams, int32_t n_streams, int64_t n_total,
                              int64_t n_steps, int64_t sims_per_stream_hint, const uint64_t *d_sim_start,
                              const void *d_plan_workspace, const mcf_gbm_spec *h_spec, int64_t slice_stride,
                              int32_t time_major, double *d_out, void *cuda_stream) {
    if (!d_streams || n_streams <= 0 || n_total <= 0 || n_steps <= 0 || !d_sim_start || !d_plan_workspace || !h_spec ||
        slice_stride <= 0 ||
        !d_out || (h_spec->mode != MCF_PORTFOLIO_GBM && h_spec->mode != MCF_PATH_TERMINAL))
        return MCF_ERR_BAD_ARG;
    GbmDerived g;
    int rc = derive_spec(*h_spec, n_steps, &g);
    if (rc) return rc;
    const int64_t n_slices = n_steps / slice_stride + 1;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    if (gen_kind == MCF_GEN_PCG64)
        gbm_slices_kernel<MCF_GEN_PCG64><<<grid_for(n_total, 256), 256, 0, cs>>>(
            d_streams, n_streams, n_total, sims_per_stream_hint, d_sim_start, (const char *)d_plan_workspace, g,
            slice_stride, n_slices, time_major, d_out);
    else
        gbm_slices_kernel<MCF_GEN_PHILOX><<<grid_for(n_total, 256), 256, 0, cs>>>(
            d_streams, n_streams, n_total, sims_per_stream_hint, d_sim_start, (const char *)d_plan_workspace, g,
            slice_stride, n_slices, time_major, d_out);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}
