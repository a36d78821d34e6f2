// mcf_stats.cu -- sm_100a kernels + C ABI for the StatsEngine reductions and the Longstaff-Schwartz sweep.
//
//   moments   : one HBM pass (8 B/element): per-thread shifted power sums -> Chan/Pebay merges in
//               warp shuffles, shared memory and a one-CTA final merge (deterministic order).
//   select    : k simultaneous order statistics by MSD radix selection on order-preserving keys;
//               8 passes x 8-bit digits, histograms in shared memory with warp-aggregated atomics,
//               rank/prefix bookkeeping stays on the device (no host round trip between passes).
//   bootstrap : NumPy's (B, n) Lemire index matrix is "the accepted 32-bit slots of one stream, in
//               order"; pass A counts accepted slots per chunk, a scan gives every chunk its first
//               ordinal, pass B regenerates the slots, gathers x[idx] (8 independent loads in
//               flight per thread) and adds into the resample the ordinal belongs to.  Indices
//               never touch HBM.
//   lsm       : per time step one reduction sweep (8 sums of the normal equations in a shifted
//               basis) + a 3x3 pseudo-inverse solve + one decision sweep; state per path is
//               (exercise step, cash flow at exercise).
#include <cuda_runtime.h>
#include <type_traits>
#include <vector>
#include <stdio.h>
#include <stdlib.h>

#include "mcf_stats_device.cuh"

using namespace mcf;

static int g_stats_last_cuda_error = 0;
extern "C" int mcf_stats_last_cuda_error(void) { return g_stats_last_cuda_error; }
#define MCF_CUDA_OK(x)                          \
    do {                                        \
        cudaError_t e__ = (x);                  \
        if (e__ != cudaSuccess) {               \
            g_stats_last_cuda_error = (int)e__; \
            return MCF_ERR_CUDA;                \
        }                                       \
    } while (0)

static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
static inline int grid_cap(int64_t work_items, int per_block, int cap) {
    int64_t g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

#define MCF_SM_COUNT 148

// ======================================================================================= moments
struct MomentsPartial {
    Moments m;
    double sum, n_nan, n_pinf, n_ninf;
};

__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

__device__ __forceinline__ MomentsPartial partial_shfl_xor(const MomentsPartial &p, int o) {
    MomentsPartial q;
    q.m.n = shfl_xor_d(p.m.n, o);
    q.m.mean = shfl_xor_d(p.m.mean, o);
    q.m.m2 = shfl_xor_d(p.m.m2, o);
    q.m.m3 = shfl_xor_d(p.m.m3, o);
    q.m.m4 = shfl_xor_d(p.m.m4, o);
    q.sum = shfl_xor_d(p.sum, o);
    q.n_nan = shfl_xor_d(p.n_nan, o);
    q.n_pinf = shfl_xor_d(p.n_pinf, o);
    q.n_ninf = shfl_xor_d(p.n_ninf, o);
    return q;
}

// merge(lower-indexed, higher-indexed): the same operand order on every lane keeps the result
// bit-identical across lanes (the update formulas are not symmetric in rounding).
__device__ __forceinline__ MomentsPartial partial_merge(const MomentsPartial &a, const MomentsPartial &b) {
    MomentsPartial r;
    r.m = moments_merge(a.m, b.m);
    r.sum = a.sum + b.sum;
    r.n_nan = a.n_nan + b.n_nan;
    r.n_pinf = a.n_pinf + b.n_pinf;
    r.n_ninf = a.n_ninf + b.n_ninf;
    return r;
}

__device__ __forceinline__ MomentsPartial partial_warp_reduce(MomentsPartial p) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        MomentsPartial q = partial_shfl_xor(p, o);
        p = (lane & o) ? partial_merge(q, p) : partial_merge(p, q);
    }
    return p;
}

// Lean per-element work (the kernel must stay under the HBM roofline, not the FP64 pipe): 6 FP64 operations
// and one exponent test per element.  Everything else is derived afterwards:
//   n = elements seen - non-finite ones,  sum x = n*c + s1,  sum (x-target)^2 = M2 + n*(mean-target)^2.
struct MomAcc {
    double c, s1, s2, s3, s4;
    unsigned long long seen, n_nan, n_pinf, n_ninf;
};

__device__ __forceinline__ void mom_add(MomAcc &a, double x) {
    const uint64_t b = double_bits(x);
    a.seen++;
    if (bits_nonfinite(b)) {  // rare
        if (b & 0x000fffffffffffffULL)
            a.n_nan++;
        else if (b >> 63)
            a.n_ninf++;
        else
            a.n_pinf++;
        return;
    }
    const double d = x - a.c, d2 = d * d;
    a.s1 += d;
    a.s2 += d2;
    a.s3 = fma(d2, d, a.s3);
    a.s4 = fma(d2, d2, a.s4);
}

__global__ void __launch_bounds__(256, 3) moments_kernel(const double *__restrict__ x, int64_t n, double target,
                                                         MomentsPartial *__restrict__ partials) {
    MomAcc a;
    a.s1 = a.s2 = a.s3 = a.s4 = 0.0;
    a.seen = a.n_nan = a.n_pinf = a.n_ninf = 0ULL;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    // shift = a value from this thread's own neighbourhood (keeps |x - c| within the sample's spread)
    {
        const double probe = x[tid < n ? tid : 0];
        a.c = bits_nonfinite(double_bits(probe)) ? 0.0 : probe;
    }
    if ((((uintptr_t)x) & 15) == 0) {  // 16-byte vector loads when the base pointer allows it
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        const int64_t n2 = n >> 1;
        int64_t i = tid;
        for (; i + 7 * nthr < n2; i += 8 * nthr) {  // 8 independent 16 B loads in flight per thread
            double2 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = __ldg(&x2[i + k * nthr]);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                mom_add(a, v[k].x);
                mom_add(a, v[k].y);
            }
        }
        for (; i < n2; i += nthr) {
            const double2 v = __ldg(&x2[i]);
            mom_add(a, v.x);
            mom_add(a, v.y);
        }
        if ((n & 1) && tid == 0) mom_add(a, x[n - 1]);
    } else {
        for (int64_t i = tid; i < n; i += nthr) mom_add(a, x[i]);
    }
    MomentsPartial p;
    const double cnt = (double)(a.seen - a.n_nan - a.n_pinf - a.n_ninf);
    p.m = moments_from_shifted(cnt, a.c, a.s1, a.s2, a.s3, a.s4);
    p.sum = cnt > 0.0 ? fma(cnt, a.c, a.s1) : 0.0;
    p.n_nan = (double)a.n_nan;
    p.n_pinf = (double)a.n_pinf;
    p.n_ninf = (double)a.n_ninf;
    p = partial_warp_reduce(p);
    __shared__ MomentsPartial sh[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        MomentsPartial r = sh[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = partial_merge(r, sh[w]);
        partials[blockIdx.x] = r;
    }
    (void)target;
}

__global__ void __launch_bounds__(256) moments_final_kernel(const MomentsPartial *__restrict__ partials, int n_partials,
                                                            double target, double *__restrict__ out) {
    // thread t merges partials t, t+256, ... in order, then the warp trees, then warp 0 merges the 8 warp results:
    // a fixed order, so the result is deterministic
    MomentsPartial p;
    p.m = moments_empty();
    p.sum = p.n_nan = p.n_pinf = p.n_ninf = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) p = partial_merge(p, partials[i]);
    p = partial_warp_reduce(p);
    __shared__ MomentsPartial sh[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        p = sh[0];
        for (int w = 1; w < 8; w++) p = partial_merge(p, sh[w]);
        const double dt = p.m.mean - target;
        out[0] = p.m.n;
        out[1] = p.n_nan;
        out[2] = p.n_pinf;
        out[3] = p.n_ninf;
        out[4] = p.m.mean;
        out[5] = p.m.m2;
        out[6] = p.m.m3;
        out[7] = p.m.m4;
        out[8] = p.m.n > 0.0 ? fma(p.m.n * dt, dt, p.m.m2) : 0.0;  // sum (x - target)^2
        out[9] = p.sum;
        out[10] = 0.0;  // reserved
        out[11] = 0.0;
    }
}

// ======================================================================================== select
#define SEL_BITS 8
#define SEL_BINS (1 << SEL_BITS)
#define SEL_PASSES (64 / SEL_BITS)
#define MCF_SELECT_SAMPLE (1 << 20)  // strided sub-sample that brackets the targets of mcf_select_bracketed

struct SelectState {
    int32_t k, n_uniq, pass, pad;
    int64_t rank[MCF_SELECT_MAX_RANKS];      // remaining rank inside the current prefix
    uint64_t prefix[MCF_SELECT_MAX_RANKS];   // bits decided so far (right aligned)
    int32_t uniq_of[MCF_SELECT_MAX_RANKS];   // rank -> index into uniq[]
    uint64_t uniq[MCF_SELECT_MAX_RANKS];     // sorted distinct prefixes
    unsigned long long compact_n;            // candidates appended by select_compact_kernel (may exceed its capacity)
    int32_t compact_done, pad2;              // 1 once the compaction pass ran
};

struct SelectGeom {
    int64_t off_state, off_hist, hist_bytes, total;
};
static inline SelectGeom select_geom() {
    SelectGeom g;
    g.off_state = 0;
    g.off_hist = align256((int64_t)sizeof(SelectState));
    g.hist_bytes = (int64_t)MCF_SELECT_MAX_RANKS * SEL_BINS * 8;
    g.total = align256(g.off_hist + g.hist_bytes);
    return g;
}

__global__ void select_init_kernel(SelectState *st, unsigned long long *hist, SelectState init) {
    for (int i = threadIdx.x; i < MCF_SELECT_MAX_RANKS * SEL_BINS; i += blockDim.x) hist[i] = 0ULL;
    if (threadIdx.x == 0) *st = init;
}

__device__ __forceinline__ int select_slot_key(uint64_t key, int pass, int shift, int n_uniq,
                                               const uint64_t *__restrict__ sh_uniq, const unsigned *__restrict__ sh_bloom) {
    if (pass == 0) return (int)(key >> 56);
    const uint64_t pre = key >> (shift + SEL_BITS);
    // cheap reject: the last decided digit of the element must be the last digit of SOME tracked prefix
    const unsigned last = (unsigned)pre & (SEL_BINS - 1);
    if (!((sh_bloom[last >> 5] >> (last & 31)) & 1u)) return -1;
    int lo = 0, hi = n_uniq - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sh_uniq[mid] < pre)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (sh_uniq[lo] != pre) return -1;
    return lo * SEL_BINS + (int)((unsigned)(key >> shift) & (SEL_BINS - 1));
}
__device__ __forceinline__ int select_slot(double v, int omit, int pass, int shift, int n_uniq,
                                           const uint64_t *__restrict__ sh_uniq, const unsigned *__restrict__ sh_bloom) {
    return select_slot_key(select_key(v, omit), pass, shift, n_uniq, sh_uniq, sh_bloom);
}

// `keys` (may be NULL): candidates compacted after the second pass (select_compact_kernel); when that pass ran and its
// buffer did not overflow, the remaining passes read those few keys instead of all of x.
__global__ void __launch_bounds__(512) select_hist_kernel(const double *__restrict__ x, int64_t n, int omit,
                                                          const SelectState *__restrict__ st,
                                                          unsigned long long *__restrict__ hist,
                                                          const uint64_t *__restrict__ keys, int64_t keys_cap) {
    __shared__ unsigned int sh_hist[MCF_SELECT_MAX_RANKS * SEL_BINS];
    __shared__ uint64_t sh_uniq[MCF_SELECT_MAX_RANKS];
    __shared__ unsigned sh_bloom[SEL_BINS / 32];
    const int n_uniq = st->n_uniq, pass = st->pass;
    for (int i = threadIdx.x; i < n_uniq * SEL_BINS; i += blockDim.x) sh_hist[i] = 0u;
    if (threadIdx.x < SEL_BINS / 32) sh_bloom[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x < n_uniq) {
        const uint64_t u = st->uniq[threadIdx.x];
        sh_uniq[threadIdx.x] = u;
        atomicOr(&sh_bloom[((unsigned)u & (SEL_BINS - 1)) >> 5], 1u << ((unsigned)u & 31u));
    }
    __syncthreads();
    const int shift = 64 - SEL_BITS * (pass + 1);  // digit position of this pass
    const int lane = threadIdx.x & 31;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    const bool from_keys = keys != nullptr && st->compact_done && (int64_t)st->compact_n <= keys_cap;
    if (from_keys) n = (int64_t)st->compact_n;
    constexpr int UNROLL = 4;
    for (int64_t i0 = tid - lane; i0 < n; i0 += UNROLL * nthr) {  // warp-uniform trip count, UNROLL loads in flight
        uint64_t key[UNROLL];
        int slot[UNROLL];
        bool any = false;
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const int64_t i = i0 + lane + k * nthr;
            key[k] = i < n ? (from_keys ? __ldg(&keys[i]) : select_key(__ldg(&x[i]), omit)) : 0ULL;
        }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const int64_t i = i0 + lane + k * nthr;
            slot[k] = i < n ? select_slot_key(key[k], pass, shift, n_uniq, sh_uniq, sh_bloom) : -1;
            any |= slot[k] >= 0;
        }
        if (__ballot_sync(0xffffffffu, any) == 0u) continue;  // later passes: nothing of interest in 128 elements
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            // warp-aggregated shared-memory atomics: one add per distinct bin per warp
            if (__ballot_sync(0xffffffffu, slot[k] >= 0) == 0u) continue;
            const unsigned peers = __match_any_sync(0xffffffffu, slot[k]);
            if (slot[k] >= 0 && (int)(__ffs(peers) - 1) == lane) atomicAdd(&sh_hist[slot[k]], __popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_uniq * SEL_BINS; i += blockDim.x) {
        const unsigned v = sh_hist[i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
}

// After the second pass the tracked prefixes (16 bits: sign, exponent, 4 mantissa bits) hold a few per cent of a
// smooth sample.  Append the keys of those candidates to `keys` (warp-aggregated: one global atomic per warp and trip)
// so that the six remaining passes stream a few per cent of the data.  More candidates than `cap` (constant or heavily
// tied data): nothing is lost, the later passes simply keep reading x.
__global__ void __launch_bounds__(512) select_compact_kernel(const double *__restrict__ x, int64_t n, int omit,
                                                             SelectState *__restrict__ st, uint64_t *__restrict__ keys,
                                                             int64_t cap) {
    __shared__ uint64_t sh_uniq[MCF_SELECT_MAX_RANKS];
    __shared__ unsigned sh_bloom[SEL_BINS / 32];
    const int n_uniq = st->n_uniq, pass = st->pass;  // pass = number of digits already decided
    if (threadIdx.x < SEL_BINS / 32) sh_bloom[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x < n_uniq) {
        const uint64_t u = st->uniq[threadIdx.x];
        sh_uniq[threadIdx.x] = u;
        atomicOr(&sh_bloom[((unsigned)u & (SEL_BINS - 1)) >> 5], 1u << ((unsigned)u & 31u));
    }
    __syncthreads();
    const int shift = 64 - SEL_BITS * (pass + 1);
    const int lane = threadIdx.x & 31;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    constexpr int UNROLL = 4;
    // per-CTA staging: the hits of one round (UNROLL * 512 elements) are collected in shared memory and appended with ONE
    // global atomic (a per-warp atomic on the single counter costs 2 ms at 1e8 elements)
    __shared__ uint64_t stage[UNROLL * 512];
    __shared__ unsigned sh_cnt;
    __shared__ unsigned long long sh_base;
    __shared__ unsigned sh_map[2048];
    const bool use_map = pass == 2;
    if (use_map) {
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh_map[i] = 0u;
        __syncthreads();
        if (threadIdx.x < n_uniq) {
            const unsigned u = (unsigned)sh_uniq[threadIdx.x] & 0xFFFFu;
            atomicOr(&sh_map[u >> 5], 1u << (u & 31u));
        }
        __syncthreads();
    }
    for (int64_t b0 = (int64_t)blockIdx.x * blockDim.x; b0 < n; b0 += UNROLL * nthr) {  // CTA-uniform trip count
        if (threadIdx.x == 0) sh_cnt = 0u;
        __syncthreads();
        uint64_t key[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const int64_t i = b0 + threadIdx.x + k * nthr;
            key[k] = i < n ? select_key(__ldg(&x[i]), omit) : 0ULL;
        }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const int64_t i = b0 + threadIdx.x + k * nthr;
            bool hit;
            if (use_map) {  // two decided digits: exact membership from a 65536-bit map, one shared-memory load
                const unsigned pre = (unsigned)(key[k] >> 48);
                hit = i < n && ((sh_map[pre >> 5] >> (pre & 31u)) & 1u);
            } else {
                hit = i < n && select_slot_key(key[k], pass, shift, n_uniq, sh_uniq, sh_bloom) >= 0;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m == 0u) continue;
            unsigned base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(&sh_cnt, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (hit) stage[base + (unsigned)__popc(m & ((1u << lane) - 1u))] = key[k];
        }
        __syncthreads();
        const unsigned cnt = sh_cnt;
        if (threadIdx.x == 0 && cnt) sh_base = atomicAdd(&st->compact_n, (unsigned long long)cnt);
        __syncthreads();
        for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x)
            if ((int64_t)(sh_base + j) < cap) keys[sh_base + j] = stage[j];
    }
}
__global__ void select_compact_done_kernel(SelectState *st) { st->compact_done = 1; }

// ------------------------------------------------------------------------------------------- bracketed selection
// Large samples: instead of two full radix passes (the second one bound by shared-memory atomics: every element is still
// a candidate) plus a compaction pass, ONE pass over x suffices.  A strided sub-sample of 2^20 values brackets every
// requested order statistic: for target rank r the sample's order statistics at r*m/n -+ delta (delta = 6 standard
// deviations of the sampling error + slack) enclose it with overwhelming probability.  The pass over x counts, for each
// bracket [L, H] (overlapping ones merged; at most 8), how many keys lie below L and inside it, and appends the inside
// keys -- a few per cent of the data -- to `keys`.  The radix passes then run on those keys alone, with ranks rebased
// to the candidate set.  The counts are exact, so the result is the exact order statistic or -- target outside its
// bracket, scratch overflow (heavy ties) -- a flag that sends the caller to the plain radix path.  Nothing is
// approximated.
#define MCF_BRACKET_MAX 8
struct BracketSet {
    int32_t n_int, k, invalid, pad;
    uint64_t lo[MCF_BRACKET_MAX], hi[MCF_BRACKET_MAX];   // merged key intervals, ascending, disjoint
    int32_t interval_of[MCF_SELECT_MAX_RANKS];           // target rank -> interval
    unsigned long long below[MCF_BRACKET_MAX], inside[MCF_BRACKET_MAX];
};

__global__ void select_sample_kernel(const double *__restrict__ x, int64_t stride, int64_t m, double *__restrict__ sample) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) sample[i] = __ldg(&x[i * stride]);
}

// one thread: bracket values (lo[0..k), hi[0..k) in target order, targets ascending) -> merged key intervals
__global__ void select_bracket_build_kernel(const double *__restrict__ bvals, int k, const int32_t *__restrict__ open_lo,
                                            const int32_t *__restrict__ open_hi, BracketSet *__restrict__ bs) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    BracketSet b;
    memset(&b, 0, sizeof(b));
    b.k = k;
    int ni = 0;
    for (int r = 0; r < k; r++) {
        const uint64_t lo = open_lo[r] ? 0ULL : select_key(bvals[r], 0);
        const uint64_t hi = open_hi[r] ? MCF_KEY_EXCLUDED : select_key(bvals[k + r], 0);
        if (ni > 0 && lo <= b.hi[ni - 1]) {  // overlaps (or touches) the previous interval: merge
            if (hi > b.hi[ni - 1]) b.hi[ni - 1] = hi;
            if (lo < b.lo[ni - 1]) b.lo[ni - 1] = lo;
        } else if (ni < MCF_BRACKET_MAX - 1) {  // the three-probe search of the pass kernel resolves 7 intervals
            b.lo[ni] = lo;
            b.hi[ni] = hi;
            ni++;
        } else {
            b.invalid = 1;  // more than 7 disjoint brackets: the caller takes the radix path
        }
        b.interval_of[r] = ni - 1;
    }
    b.n_int = ni;
    *bs = b;
}

__global__ void __launch_bounds__(512) select_bracket_pass_kernel(const double *__restrict__ x, int64_t n,
                                                                  BracketSet *__restrict__ bs, SelectState *__restrict__ st,
                                                                  uint64_t *__restrict__ keys, int64_t cap) {
    // interval search: p = #{j : lo[j] <= key} by three probes of a sorted 8-entry table (at most 7 live intervals, dead
    // entries hold the maximum key), then one probe of hi[p-1]; per-thread counters live in shared memory, one column per
    // thread (no atomics, no bank conflicts): hist[p] counts keys with exactly p interval starts at or below them.
    __shared__ uint64_t sh_lo[MCF_BRACKET_MAX], sh_hi[MCF_BRACKET_MAX];
    __shared__ unsigned short sh_hist[MCF_BRACKET_MAX][512], sh_in[MCF_BRACKET_MAX][512];  // < 2^16 elements per thread
    constexpr int UNROLL = 4;
    __shared__ uint64_t stage[UNROLL * 512];
    __shared__ unsigned sh_cnt;
    __shared__ unsigned long long sh_base;
    if (threadIdx.x < MCF_BRACKET_MAX) {
        const bool live = (int)threadIdx.x < bs->n_int;
        sh_lo[threadIdx.x] = live ? bs->lo[threadIdx.x] : MCF_KEY_EXCLUDED;
        sh_hi[threadIdx.x] = live ? bs->hi[threadIdx.x] : 0ULL;
    }
#pragma unroll
    for (int j = 0; j < MCF_BRACKET_MAX; j++) {
        sh_hist[j][threadIdx.x] = 0;
        sh_in[j][threadIdx.x] = 0;
    }
    __syncthreads();
    const bool top_open = bs->n_int > 0 && bs->hi[bs->n_int - 1] == MCF_KEY_EXCLUDED;  // a bracket open towards +inf
    const int lane = threadIdx.x & 31;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b0 = (int64_t)blockIdx.x * blockDim.x; b0 < n; b0 += UNROLL * nthr) {  // CTA-uniform trip count
        if (threadIdx.x == 0) sh_cnt = 0u;
        __syncthreads();
        uint64_t key[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; q++) {
            const int64_t i = b0 + threadIdx.x + q * nthr;
            key[q] = i < n ? select_key(__ldg(&x[i]), 0) : MCF_KEY_EXCLUDED;
        }
#pragma unroll
        for (int q = 0; q < UNROLL; q++) {
            const int64_t i = b0 + threadIdx.x + q * nthr;
            bool hit = false;
            if (i < n) {
                // dead entries compare as "above every key" except for the maximum key itself (NaN), which an open
                // top bracket must still catch: strict < on dead entries is what `lo <= key` gives for lo = MAX only
                // when key = MAX, and then hi = 0 rejects it unless the live top bracket is open (handled below)
                int p = sh_lo[3] <= key[q] ? 4 : 0;
                p += sh_lo[p + 1] <= key[q] ? 2 : 0;
                p += sh_lo[p] <= key[q] ? 1 : 0;
                if (key[q] == MCF_KEY_EXCLUDED) p = top_open ? bs->n_int : p;  // NaN keys: the last live interval or none
                hit = p > 0 && key[q] <= sh_hi[p - 1];
                sh_hist[p][threadIdx.x] += 1;
                if (hit) sh_in[p - 1][threadIdx.x] += 1;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m == 0u) continue;
            unsigned base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(&sh_cnt, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (hit) stage[base + (unsigned)__popc(m & ((1u << lane) - 1u))] = key[q];
        }
        __syncthreads();
        const unsigned cnt = sh_cnt;
        if (threadIdx.x == 0 && cnt) sh_base = atomicAdd(&st->compact_n, (unsigned long long)cnt);
        __syncthreads();
        for (unsigned j = threadIdx.x; j < cnt; j += blockDim.x)
            if ((int64_t)(sh_base + j) < cap) keys[sh_base + j] = stage[j];
    }
    __syncthreads();
    // column sums: warp w adds rows w and w+... over the 512 columns
    const int wid = threadIdx.x >> 5;
    for (int row = wid; row < 2 * MCF_BRACKET_MAX; row += 16) {
        const unsigned short *src = row < MCF_BRACKET_MAX ? sh_hist[row] : sh_in[row - MCF_BRACKET_MAX];
        unsigned long long v = 0ULL;
        for (int c = lane; c < 512; c += 32) v += src[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) {
            if (row < MCF_BRACKET_MAX) {
                // hist[p]: keys with p starts at or below them are BELOW the starts of intervals p, p+1, ...
                for (int j = row; j < MCF_BRACKET_MAX; j++) atomicAdd(&bs->below[j], v);
            } else {
                atomicAdd(&bs->inside[row - MCF_BRACKET_MAX], v);
            }
        }
    }
}

// one thread: ranks relative to x -> ranks relative to the candidate set; anything that does not fit raises `invalid`
__global__ void select_bracket_rebase_kernel(BracketSet *__restrict__ bs, SelectState *__restrict__ st, int64_t cap,
                                             int32_t *__restrict__ invalid_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int bad = bs->invalid;
    if ((int64_t)st->compact_n > cap) bad = 1;
    unsigned long long before[MCF_BRACKET_MAX];
    unsigned long long acc = 0ULL;
    for (int j = 0; j < MCF_BRACKET_MAX; j++) {
        before[j] = acc;
        acc += bs->inside[j];
    }
    for (int r = 0; r < st->k && !bad; r++) {
        const int j = bs->interval_of[r];
        const unsigned long long rank = (unsigned long long)st->rank[r];
        if (j < 0 || j >= bs->n_int || rank < bs->below[j] || rank >= bs->below[j] + bs->inside[j]) {
            bad = 1;
            break;
        }
        st->rank[r] = (int64_t)(rank - bs->below[j] + before[j]);
    }
    st->compact_done = 1;
    *invalid_out = bad;
}

// one CTA: every rank walks the histogram of its prefix, picks its digit, then prefixes are re-uniqued
__global__ void __launch_bounds__(1024) select_choose_kernel(SelectState *st, unsigned long long *hist) {
    __shared__ SelectState s;
    static_assert(sizeof(SelectState) % 8 == 0, "copied as 64-bit words");
    for (int i = threadIdx.x; i < (int)(sizeof(SelectState) / 8); i += blockDim.x)  // (one thread copying 1.3 KB was a third of this kernel)
        reinterpret_cast<uint64_t *>(&s)[i] = reinterpret_cast<const uint64_t *>(st)[i];
    __syncthreads();
    const int n_rows = s.n_uniq;
    // one warp per rank: 8 coalesced 32-bin loads + shuffle scans instead of 256 dependent global loads
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    if (r < s.k) {
        const unsigned long long *h = hist + (size_t)s.uniq_of[r] * SEL_BINS;
        int64_t rem = s.rank[r];
        int d = SEL_BINS - 1;
        bool found = false;
        for (int c0 = 0; c0 < SEL_BINS && !found; c0 += 32) {
            const int64_t c = (int64_t)h[c0 + lane];
            int64_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, rem < incl);
            if (hit) {
                const int first = __ffs(hit) - 1;
                const int64_t before = __shfl_sync(0xffffffffu, incl - c, first);
                d = c0 + first;
                rem -= before;
                found = true;
            } else {
                rem -= __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (!found) {  // rank beyond the counted elements (cannot happen for valid ranks): clamp to the last bin
            d = SEL_BINS - 1;
        }
        if (lane == 0) {
            s.rank[r] = rem;
            s.prefix[r] = (s.prefix[r] << SEL_BITS) | (uint64_t)d;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // insertion sort of the distinct prefixes (k <= 32)
        int nu = 0;
        for (int r = 0; r < s.k; r++) {
            const uint64_t p = s.prefix[r];
            int pos = 0;
            bool found = false;
            for (; pos < nu; pos++) {
                if (s.uniq[pos] == p) {
                    found = true;
                    break;
                }
                if (s.uniq[pos] > p) break;
            }
            if (!found) {
                for (int j = nu; j > pos; j--) s.uniq[j] = s.uniq[j - 1];
                s.uniq[pos] = p;
                nu++;
            }
        }
        for (int r = 0; r < s.k; r++)
            for (int u = 0; u < nu; u++)
                if (s.uniq[u] == s.prefix[r]) s.uniq_of[r] = u;
        s.n_uniq = nu;
        s.pass += 1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)(sizeof(SelectState) / 8); i += blockDim.x)
        reinterpret_cast<uint64_t *>(st)[i] = reinterpret_cast<const uint64_t *>(&s)[i];
    for (int i = threadIdx.x; i < n_rows * SEL_BINS; i += blockDim.x) hist[i] = 0ULL;  // only the rows that were used
}

__global__ void select_finish_kernel(const SelectState *st, double *out) {
    if (threadIdx.x < st->k) out[threadIdx.x] = select_unkey(st->prefix[threadIdx.x]);
}

// ===================================================================================== bootstrap
struct BootGeom {
    int64_t n_chunks, n_blocks;
    int64_t off_counts, off_bsum, off_bbase, off_total, off_list, total;
};
#define BOOT_THREADS 256
#define BOOT_FINE 16  // the exact pass of the single-pass form cuts a CTA's slots into up to 16 finer CTAs
static inline BootGeom boot_geom(int64_t n_slots, int32_t chunk_slots) {
    BootGeom g;
    g.n_chunks = (n_slots + chunk_slots - 1) / chunk_slots;
    g.n_blocks = (g.n_chunks + BOOT_THREADS - 1) / BOOT_THREADS;
    int64_t o = 0;
    g.off_total = o;
    o += 256;
    g.off_counts = o;
    o = align256(o + g.n_blocks * BOOT_THREADS * 4);
    // (x BOOT_FINE: the single-pass form prefix-sums and lists FINE blocks -- see mcf_bootstrap_fused_phase1)
    g.off_bsum = o;
    o = align256(o + g.n_blocks * 8 * BOOT_FINE);
    g.off_bbase = o;
    o = align256(o + g.n_blocks * 8 * BOOT_FINE);
    g.off_list = o;  // block list of the single-pass form (CTAs that need exact ordinals)
    o = align256(o + g.n_blocks * 8);
    g.total = o;
    return g;
}

// Walks the 32-bit slots [s0, s1) of one stream in order.
template <int KIND>
struct SlotWalker {
    Cursor<KIND> cur;
    uint64_t v;
    uint32_t has, uinteger;
    bool have_v, aligned;
    __device__ __forceinline__ void init(const mcf_stream &st, uint64_t s0) {
        has = st.has_uint32 ? 1u : 0u;
        uinteger = st.uinteger;
        const uint64_t t0 = s0 >= has ? s0 - has : 0;
        cur.init(st, t0 >> 1);
        have_v = false;
        v = 0;
        aligned = has == 0 && (s0 & 1ULL) == 0;  // chunk lengths are multiples of 8: every group starts on a draw
    }
    // the 8 slots [sb, sb+8) -- four whole draws when the walker is aligned (no per-slot parity bookkeeping)
    __device__ __forceinline__ void get8(uint64_t sb, uint32_t u[8]) {
        if (aligned) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t d = cur.next64();
                u[2 * k] = (uint32_t)d;
                u[2 * k + 1] = (uint32_t)(d >> 32);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) u[j] = get(sb + j);
        }
    }
    __device__ __forceinline__ uint32_t get(uint64_t s) {
        if (has && s == 0) return uinteger;
        const uint64_t t = s - has;
        if (!(t & 1ULL) || !have_v) {
            v = cur.next64();
            have_v = true;
        }
        return (t & 1ULL) ? (uint32_t)(v >> 32) : (uint32_t)v;
    }
};

template <int KIND>
__global__ void __launch_bounds__(BOOT_THREADS) boot_count_kernel(const mcf_stream *__restrict__ stream, uint32_t n,
                                                                  uint64_t slot_lo, uint64_t slot_hi, int chunk_slots,
                                                                  int64_t n_chunks, uint32_t *__restrict__ counts,
                                                                  unsigned long long *__restrict__ block_sum,
                                                                  const int64_t *__restrict__ block_list = nullptr) {
    // block_list (may be NULL): the CTAs of this launch count the listed blocks instead of blocks 0 .. gridDim.x-1
    const int64_t blk = block_list ? block_list[blockIdx.x] : (int64_t)blockIdx.x;
    const int64_t chunk = blk * BOOT_THREADS + threadIdx.x;
    uint32_t cnt = 0;
    if (chunk < n_chunks) {
        const SlotRule rule = slot_rule(n);
        const uint64_t s0 = slot_lo + (uint64_t)chunk * (uint64_t)chunk_slots;
        uint64_t s1 = s0 + (uint64_t)chunk_slots;
        if (s1 > slot_hi) s1 = slot_hi;
        if (rule.threshold == 0) {
            cnt = (uint32_t)(s1 - s0);
        } else {
            const mcf_stream st = *stream;
            SlotWalker<KIND> w;
            w.init(st, s0);
            for (uint64_t sb = s0; sb < s1; sb += 8) {
                uint32_t u[8];
                w.get8(sb, u);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t idx;
                    cnt += (sb + j < s1 && slot_accept(rule, u[j], idx)) ? 1u : 0u;
                }
            }
        }
    }
    // counts[] is padded to n_blocks * BOOT_THREADS; with a block list it is COMPACT (indexed by list position)
    counts[block_list ? (int64_t)blockIdx.x * BOOT_THREADS + threadIdx.x : chunk] = cnt;
    __shared__ uint32_t sh[BOOT_THREADS / 32];
    uint32_t v = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w2 = 0; w2 < BOOT_THREADS / 32; w2++) t += sh[w2];
        block_sum[blk] = t;
    }
}

// one CTA: exclusive scan of u64 values; total written to *total
__global__ void __launch_bounds__(1024) scan_u64_kernel(const unsigned long long *__restrict__ in, int64_t n,
                                                        unsigned long long *__restrict__ out,
                                                        unsigned long long *__restrict__ total) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_sh;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < n; t0 += blockDim.x) {
        const int64_t i = t0 + threadIdx.x;
        const unsigned long long v = i < n ? in[i] : 0ULL;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long w = warp_tot[lane];
            unsigned long long z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, z, o);
                if (lane >= o) z += y;
            }
            warp_tot[lane] = z - w;
        }
        __syncthreads();
        const unsigned long long incl = carry_sh + warp_tot[wid] + x;
        if (i < n) out[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_sh = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_sh;
}

template <int KIND>
__global__ void __launch_bounds__(BOOT_THREADS) boot_accumulate_kernel(
    const mcf_stream *__restrict__ stream, const double *__restrict__ x, uint32_t n, uint64_t n_resamples,
    uint64_t slot_lo, uint64_t slot_hi, int chunk_slots, int64_t n_chunks, uint64_t ordinal_base,
    const uint32_t *__restrict__ counts, const unsigned long long *__restrict__ block_base, double *__restrict__ sums,
    unsigned long long *__restrict__ end_slot) {
    __shared__ uint32_t warp_tot[BOOT_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chunk = (int64_t)blockIdx.x * BOOT_THREADS + threadIdx.x;
    // block-exclusive scan of the per-chunk accepted counts
    const uint32_t cnt = counts[chunk];
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w2 = 0; w2 < wid; w2++) wbase += warp_tot[w2];
    uint64_t ord = ordinal_base + block_base[blockIdx.x] + wbase + (incl - cnt);
    const uint64_t total_needed = n_resamples * (uint64_t)n;

    double acc = 0.0;
    uint64_t b = 0;
    bool touched = false, flushed = false;
    if (chunk < n_chunks && cnt > 0 && ord < total_needed) {
        const SlotRule rule = slot_rule(n);
        const mcf_stream st = *stream;
        const uint64_t s0 = slot_lo + (uint64_t)chunk * (uint64_t)chunk_slots;
        uint64_t s1 = s0 + (uint64_t)chunk_slots;
        if (s1 > slot_hi) s1 = slot_hi;
        b = ord / n;
        uint64_t next_boundary = (b + 1) * (uint64_t)n;
        SlotWalker<KIND> w;
        w.init(st, s0);
        bool done = false;
        for (uint64_t sb = s0; sb < s1 && !done; sb += 8) {
            uint32_t idx[8], u[8];
            bool ok[8];
            double v[8];
            w.get8(sb, u);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                idx[j] = 0;
                ok[j] = slot_accept(rule, u[j], idx[j]) && (sb + j < s1);
            }
#pragma unroll
            // ld.global.cg: random 8-byte gathers must not drag whole 128-byte lines through L1 (measured: 113 B of DRAM
            // traffic per gather with ld.global.nc) -- L2-only caching requests single 32-byte sectors
            for (int j = 0; j < 8; j++) v[j] = ok[j] ? __ldcg(&x[idx[j]]) : 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (ok[j] && !done) {
                    if (ord == next_boundary) {
                        atomicAdd(&sums[b], acc);
                        acc = 0.0;
                        flushed = true;
                        b++;
                        next_boundary += n;
                    }
                    acc += v[j];
                    touched = true;
                    ord++;
                    if (ord == total_needed) {
                        *end_slot = sb + j + 1;  // slots consumed by the whole (B, n) draw
                        done = true;
                    }
                }
            }
        }
    }
    // common case: the whole warp stayed inside one resample -> one atomic per warp
    const uint64_t b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool simple = __all_sync(0xffffffffu, !touched || (b == b0 && !flushed));
    if (simple) {
        double t = touched ? acc : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        // lanes that did not touch anything may carry a stale b; use the first touched lane's resample
        const unsigned tm = __ballot_sync(0xffffffffu, touched);
        if (tm) {
            const uint64_t bt = __shfl_sync(0xffffffffu, b, __ffs(tm) - 1);
            if (lane == 0) atomicAdd(&sums[bt], t);
        }
    } else if (touched) {
        atomicAdd(&sums[b], acc);
    }
}

// ----------------------------------------------------------------------------- binned bootstrap (large populations)
// A random 8-byte gather from an HBM-resident vector moves a whole 128-byte line (measured: 113 B of DRAM traffic
// per gather, 5.8e10 gathers/s = the HBM limit at that granularity), while the same gather from an L2-resident
// table runs at 1.7e11/s.  For populations larger than L2 the work is therefore split:
//   boot_bin_kernel    : regenerate the slots of a batch of resamples and append every accepted index to the queue
//                        of (its resample, its 2^shift-element REGION of x) -- 4 bytes written per index, through
//                        per-CTA shared-memory staging (one global atomic per queue per flush, coalesced stores);
//   boot_gather_kernel : region by region (so that the region's part of x stays in L2 while all queues that point
//                        into it are streamed once), sum x[region + local index] per (resample, region) queue.
// The index stream is unchanged (bit-exact); only the order in which a resample's terms are added differs.
#define BIN_MAX_REGIONS 32
#define BIN_STAGE_CAP 352   // entries per staging buffer; flushed every BIN_FLUSH_EVERY groups of 8 slots per thread
#define BIN_FLUSH_EVERY 3   // 3 * 2048 entries per CTA over >= 24 regions: 256 +- 16 per buffer

struct BinParams {
    uint32_t n;               // population size
    uint32_t region_shift;    // region = idx >> region_shift
    uint32_t n_regions;       // <= BIN_MAX_REGIONS
    uint32_t n_res_window;    // resamples that can receive entries in this batch (queues per region)
    uint64_t b_first;         // first resample of the batch window
    uint64_t cap;             // entries per queue
    uint64_t n_resamples;     // B
};

template <int KIND>
__global__ void __launch_bounds__(BOOT_THREADS) boot_bin_kernel(const mcf_stream *__restrict__ stream, BinParams bp,
                                                                uint64_t slot_lo, uint64_t slot_hi, int chunk_slots,
                                                                int64_t n_chunks, int64_t block0, uint64_t ordinal_base,
                                                                const uint32_t *__restrict__ counts,
                                                                const unsigned long long *__restrict__ block_base,
                                                                uint32_t *__restrict__ queues,
                                                                unsigned int *__restrict__ tails,
                                                                unsigned long long *__restrict__ end_slot, int *overflow,
                                                                const int64_t *__restrict__ block_list = nullptr) {
    __shared__ uint32_t warp_tot[BOOT_THREADS / 32];
    // staging for the CTA's FIRST resample only; the few CTAs that straddle a resample boundary append the entries of
    // the second one directly (1 CTA in n / (256 * chunk_slots))
    __shared__ uint32_t sh_stage[BIN_MAX_REGIONS][BIN_STAGE_CAP];
    __shared__ unsigned int sh_cnt[BIN_MAX_REGIONS], sh_n[BIN_MAX_REGIONS], sh_base[BIN_MAX_REGIONS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t blk = block_list ? block_list[block0 + blockIdx.x] : block0 + blockIdx.x;
    const int64_t chunk = blk * BOOT_THREADS + threadIdx.x;
    const uint32_t R = bp.n_regions;
    if (threadIdx.x < BIN_MAX_REGIONS) sh_cnt[threadIdx.x] = 0;
    // block-exclusive scan of the per-chunk accepted counts (as in boot_accumulate_kernel); compact with a block list
    const uint32_t cnt = counts[block_list ? (block0 + (int64_t)blockIdx.x) * BOOT_THREADS + threadIdx.x : chunk];
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w2 = 0; w2 < wid; w2++) wbase += warp_tot[w2];
    const uint64_t blk_ord = ordinal_base + block_base[blk];  // first ordinal of this CTA's slots
    uint64_t ord = blk_ord + wbase + (incl - cnt);
    const uint64_t total_needed = bp.n_resamples * (uint64_t)bp.n;
    const uint64_t b_blk = blk_ord / bp.n;  // the CTA's slots (256 * chunk_slots <= n) touch resamples b_blk, b_blk + 1

    const SlotRule rule = slot_rule(bp.n);
    const mcf_stream st = *stream;
    const uint64_t s0 = slot_lo + (uint64_t)chunk * (uint64_t)chunk_slots;
    uint64_t s1 = s0 + (uint64_t)chunk_slots;
    if (s1 > slot_hi) s1 = slot_hi;
    const bool live = chunk < n_chunks && s0 < slot_hi;
    uint64_t b = live ? ord / bp.n : 0, next_boundary = (b + 1) * (uint64_t)bp.n;
    SlotWalker<KIND> w;
    if (live) w.init(st, s0);
    bool done = !live || ord >= total_needed;
    const uint32_t mask = (1u << bp.region_shift) - 1u;
    const int groups = chunk_slots / 8;
    const uint64_t q_row = (b_blk - bp.b_first) * R;  // queues of the CTA's first resample

    // CTA-wide flush of the staging buffers: one global atomic per region, coalesced stores
    auto flush = [&]() {
        __syncthreads();
        if (threadIdx.x < R) {
            unsigned c = sh_cnt[threadIdx.x];
            if (c > BIN_STAGE_CAP) c = BIN_STAGE_CAP;
            sh_n[threadIdx.x] = c;
            sh_base[threadIdx.x] = c ? atomicAdd(&tails[q_row + threadIdx.x], c) : 0u;
            sh_cnt[threadIdx.x] = 0;
        }
        __syncthreads();
        for (unsigned q = wid; q < R; q += BOOT_THREADS / 32) {
            const unsigned c = sh_n[q], base = sh_base[q];
            const uint64_t gq = q_row + q;
            if (c && (uint64_t)base + c > bp.cap) atomicExch(overflow, 1);
            for (unsigned j = lane; j < c; j += 32)
                if ((uint64_t)base + j < bp.cap) queues[gq * bp.cap + base + j] = sh_stage[q][j];
        }
        __syncthreads();
    };
    auto append_direct = [&](uint64_t bb, uint32_t idx) {  // staging full, or the second resample of a straddling CTA
        const uint64_t gq = (bb - bp.b_first) * R + (idx >> bp.region_shift);
        const unsigned gp = atomicAdd(&tails[gq], 1u);
        if (gp < bp.cap)
            queues[gq * bp.cap + gp] = idx & mask;
        else
            atomicExch(overflow, 1);
    };

    // Nearly every CTA lies inside ONE resample, before the end of the (B, n) draw, with all its slots in range:
    // no ordinal bookkeeping in the loop, just accept -> stage.
    uint64_t cta_total = 0;
    for (int w2 = 0; w2 < BOOT_THREADS / 32; w2++) cta_total += warp_tot[w2];
    const uint64_t cta_end = blk_ord + cta_total;
    const bool simple = cta_end <= (b_blk + 1) * (uint64_t)bp.n && cta_end < total_needed &&
                        (blk + 1) * BOOT_THREADS <= n_chunks &&
                        slot_lo + (uint64_t)(blk + 1) * BOOT_THREADS * (uint64_t)chunk_slots <= slot_hi;
    if (simple) {
        // 32-bit shared-window addresses, computed once (the generic-pointer form recomputes them per slot)
        unsigned cnt_addr = (unsigned)__cvta_generic_to_shared(sh_cnt);
        unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&sh_stage[0][0]);
        unsigned shift = bp.region_shift, lowmask = mask;
        asm volatile("" : "+r"(cnt_addr), "+r"(stage_addr), "+r"(shift), "+r"(lowmask));  // keep them in registers
        for (int g = 0; g < groups; g++) {
            uint32_t u[8];
            w.get8(s0 + (uint64_t)g * 8, u);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t idx;
                if (slot_accept(rule, u[j], idx)) {
                    const unsigned q = idx >> shift;
                    unsigned pos;
                    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(cnt_addr + q * 4u) : "memory");
                    if (pos < BIN_STAGE_CAP)
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage_addr + q * (BIN_STAGE_CAP * 4u) + pos * 4u),
                                     "r"(idx & lowmask)
                                     : "memory");
                    else
                        append_direct(b_blk, idx);
                }
            }
            if ((g % BIN_FLUSH_EVERY) == BIN_FLUSH_EVERY - 1 || g == groups - 1) flush();
        }
        return;
    }
    for (int g = 0; g < groups; g++) {  // uniform trip count: the flushes are CTA-wide
        const uint64_t sb = s0 + (uint64_t)g * 8;
        if (!done && sb < s1) {
            uint32_t u[8];
            w.get8(sb, u);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint64_t s = sb + j;
                uint32_t idx;
                if (s < s1 && !done && slot_accept(rule, u[j], idx)) {
                    if (ord == next_boundary) {
                        b++;
                        next_boundary += bp.n;
                    }
                    const unsigned q = idx >> bp.region_shift;
                    const unsigned pos = b == b_blk ? atomicAdd(&sh_cnt[q], 1u) : 0xFFFFFFFFu;
                    if (pos < BIN_STAGE_CAP)
                        sh_stage[q][pos] = idx & mask;
                    else
                        append_direct(b, idx);
                    ord++;
                    if (ord == total_needed) {
                        *end_slot = s + 1;  // slots consumed by the whole (B, n) draw
                        done = true;
                    }
                }
            }
        }
        if ((g % BIN_FLUSH_EVERY) == BIN_FLUSH_EVERY - 1 || g == groups - 1) flush();
    }
}

// ------------------------------------------------------------------------------------ single-pass ("safe zone") form
// The count pass exists only to give every CTA the ordinal of its first accepted slot -- i.e. the RESAMPLE its indices
// belong to.  But the ordinal at slot s is s * (1 - p) to within a few standard deviations of a binomial count
// (p = rejection probability): at 1e12 slots that is +-1.2e6 of 1e8 per resample.  A CTA whose estimated ordinal range,
// widened by that margin, lies strictly inside ONE resample (96 % of them) needs no exact ordinal at all: it bins into
// that resample in the first and only pass over its slots and counts what it accepted on the way
// (boot_bin_safe_kernel).  Only the CTAs near resample boundaries and near the end of the draw are counted
// (boot_count_kernel on a block list), prefix-summed together with the safe CTAs' counts, and binned with exact
// ordinals by the general kernel.  A verification pass checks every safe CTA against its exact ordinal range
// afterwards; a violation (an 8-sigma event) raises a flag and the caller recomputes with the two-pass form.
struct FuseGeom {
    uint64_t q;             // 2^32 - threshold: accepted values of a 32-bit slot
    uint64_t margin;        // ordinals
    uint64_t n, total_needed;
    uint64_t S;             // slots per CTA
    uint64_t slot_lo, slot_hi;
};
MCF_HD uint64_t fuse_est_ordinal(uint64_t s, uint64_t q) {  // floor(s * q / 2^32), exact
    const uint64_t hi = mulhi64(s, q), lo = s * q;
    return (hi << 32) | (lo >> 32);
}
// 1: safe (resample *b), 0: needs exact ordinals, -1: wholly beyond the end of the draw (nothing to do)
MCF_HD int fuse_block_class(const FuseGeom &g, int64_t blk, uint64_t *b, uint64_t *b_min, uint64_t *b_max) {
    const uint64_t s0 = g.slot_lo + (uint64_t)blk * g.S;
    const uint64_t s1 = s0 + g.S < g.slot_hi ? s0 + g.S : g.slot_hi;
    const uint64_t e0 = fuse_est_ordinal(s0, g.q), e1 = fuse_est_ordinal(s1, g.q);
    const uint64_t lo = e0 > g.margin ? e0 - g.margin : 0ULL, hi = e1 + g.margin;
    if (lo >= g.total_needed) return -1;
    *b_min = lo / g.n;
    *b_max = (hi < g.total_needed ? hi : g.total_needed - 1ULL) / g.n;
    if (s1 - s0 < g.S || e0 < g.margin || hi >= g.total_needed || lo / g.n != hi / g.n) return 0;
    *b = lo / g.n;
    return 1;
}

template <int KIND>
__global__ void __launch_bounds__(BOOT_THREADS) boot_bin_safe_kernel(const mcf_stream *__restrict__ stream, BinParams bp,
                                                                     FuseGeom fg, int chunk_slots, int64_t block0, int fine,
                                                                     unsigned long long *__restrict__ block_sum,
                                                                     uint32_t *__restrict__ queues,
                                                                     unsigned int *__restrict__ tails, int *overflow) {
    __shared__ uint32_t sh_stage[BIN_MAX_REGIONS][BIN_STAGE_CAP];
    __shared__ unsigned int sh_cnt[BIN_MAX_REGIONS], sh_n[BIN_MAX_REGIONS], sh_base[BIN_MAX_REGIONS];
    __shared__ unsigned int sh_acc[BOOT_THREADS / 32];
    const int64_t blk = block0 + blockIdx.x;
    uint64_t b = 0, bmin = 0, bmax = 0;
    const int cls = fuse_block_class(fg, blk, &b, &bmin, &bmax);
    if (cls != 1 || b < bp.b_first || b - bp.b_first >= (uint64_t)bp.n_res_window) {
        // not a safe CTA of this batch: its count comes from the exact pass (or it lies beyond the draw)
        return;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t R = bp.n_regions;
    if (threadIdx.x < BIN_MAX_REGIONS) sh_cnt[threadIdx.x] = 0;
    __syncthreads();
    const SlotRule rule = slot_rule(bp.n);
    const mcf_stream st = *stream;
    const int64_t chunk = blk * BOOT_THREADS + threadIdx.x;
    const uint64_t s0 = fg.slot_lo + (uint64_t)chunk * (uint64_t)chunk_slots;
    SlotWalker<KIND> w;
    w.init(st, s0);
    const uint32_t mask = (1u << bp.region_shift) - 1u;
    const int groups = chunk_slots / 8;
    const uint64_t q_row = (b - bp.b_first) * R;
    auto flush = [&]() {
        __syncthreads();
        if (threadIdx.x < R) {
            unsigned c = sh_cnt[threadIdx.x];
            if (c > BIN_STAGE_CAP) c = BIN_STAGE_CAP;
            sh_n[threadIdx.x] = c;
            sh_base[threadIdx.x] = c ? atomicAdd(&tails[q_row + threadIdx.x], c) : 0u;
            sh_cnt[threadIdx.x] = 0;
        }
        __syncthreads();
        for (unsigned q = wid; q < R; q += BOOT_THREADS / 32) {
            const unsigned c = sh_n[q], base = sh_base[q];
            const uint64_t gq = q_row + q;
            if (c && (uint64_t)base + c > bp.cap) atomicExch(overflow, 1);
            for (unsigned j = lane; j < c; j += 32)
                if ((uint64_t)base + j < bp.cap) queues[gq * bp.cap + base + j] = sh_stage[q][j];
        }
        __syncthreads();
    };
    unsigned cnt_addr = (unsigned)__cvta_generic_to_shared(sh_cnt);
    unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&sh_stage[0][0]);
    unsigned shift = bp.region_shift, lowmask = mask;
    asm volatile("" : "+r"(cnt_addr), "+r"(stage_addr), "+r"(shift), "+r"(lowmask));
    unsigned accepted = 0;
    for (int g = 0; g < groups; g++) {
        uint32_t u[8];
        w.get8(s0 + (uint64_t)g * 8, u);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t idx;
            if (slot_accept(rule, u[j], idx)) {
                accepted++;
                const unsigned q = idx >> shift;
                unsigned pos;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(cnt_addr + q * 4u) : "memory");
                if (pos < BIN_STAGE_CAP) {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage_addr + q * (BIN_STAGE_CAP * 4u) + pos * 4u),
                                 "r"(idx & lowmask)
                                 : "memory");
                } else {  // staging full: append directly
                    const uint64_t gq = q_row + q;
                    const unsigned gp = atomicAdd(&tails[gq], 1u);
                    if (gp < bp.cap)
                        queues[gq * bp.cap + gp] = idx & lowmask;
                    else
                        atomicExch(overflow, 1);
                }
            }
        }
        if ((g % BIN_FLUSH_EVERY) == BIN_FLUSH_EVERY - 1 || g == groups - 1) flush();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accepted += __shfl_xor_sync(0xffffffffu, accepted, o);
    if (lane == 0) sh_acc[wid] = accepted;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w2 = 0; w2 < BOOT_THREADS / 32; w2++) t += sh_acc[w2];
        block_sum[blk * fine] = t;  // the sums are kept per FINE block: a safe CTA owns the first of its `fine` entries
    }
}

// every safe CTA against its exact ordinal range
__global__ void boot_fuse_verify_kernel(FuseGeom fg, int64_t n_blocks, int fine, uint64_t ordinal_base,
                                        const unsigned long long *__restrict__ block_base,
                                        const unsigned long long *__restrict__ block_sum, int *__restrict__ flag) {
    for (int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; blk < n_blocks; blk += (int64_t)gridDim.x * blockDim.x) {
        uint64_t b = 0, bmin = 0, bmax = 0;
        if (fuse_block_class(fg, blk, &b, &bmin, &bmax) != 1) continue;
        const uint64_t ord0 = ordinal_base + block_base[blk * fine], ord1 = ord0 + block_sum[blk * fine];
        if (ord0 < b * fg.n || ord1 > (b + 1ULL) * fg.n || ord1 > fg.total_needed) atomicExch(flag, 1);
    }
}

// grid (G, n_res_window): one (resample, region) queue per blockIdx.y, `region` fixed per launch
__global__ void __launch_bounds__(256) boot_gather_kernel(const double *__restrict__ x, BinParams bp, uint32_t region,
                                                          const uint32_t *__restrict__ queues,
                                                          const unsigned int *__restrict__ tails,
                                                          double *__restrict__ sums) {
    const uint64_t gq = (uint64_t)blockIdx.y * bp.n_regions + region;
    uint64_t cnt = tails[gq];
    if (cnt > bp.cap) cnt = bp.cap;
    const uint32_t *seg = queues + gq * bp.cap;
    const double *xr = x + ((uint64_t)region << bp.region_shift);
    double acc = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 15 * stride < cnt; i += 16 * stride) {  // 16 independent gathers in flight: the kernel is L2-latency bound
        uint32_t e[16];
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) e[k] = __ldcs(&seg[i + k * stride]);  // streamed once: evict-first
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = __ldcg(&xr[e[k]]);
#pragma unroll
        for (int k = 0; k < 16; k++) acc += v[k];
    }
    for (; i < cnt; i += stride) acc += __ldcg(&xr[__ldcs(&seg[i])]);
    __shared__ double sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w2 = 0; w2 < 8; w2++) t += sh[w2];
        if (cnt) atomicAdd(&sums[bp.b_first + blockIdx.y], t);
    }
}

__global__ void boot_finalize_kernel(double *__restrict__ sums, int64_t B, double n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) sums[i] = __ddiv_rn(sums[i], n);
}

// ============================================================================ declared integer draws
// Replication i of a stream reads the 32-bit slots [i*k, (i+1)*k) of its stream (jump-ahead, no state carried from
// one replication to the next) through NumPy's 32-bit Lemire rule.  A rejected slot would shift all later ones: it is
// only flagged here (the caller then runs the reference's own loop), never papered over.
template <int KIND>
__global__ void __launch_bounds__(256) draw_integers_kernel(const mcf_stream *__restrict__ streams, int n_streams,
                                                            int64_t n_total, int64_t hint, uint32_t span, int64_t low,
                                                            int32_t k, int32_t op, int64_t match, double *__restrict__ out,
                                                            int *__restrict__ rejected) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    int64_t g = hint > 0 ? i / hint : 0;
    int si;
    if (g < n_streams && streams[g].first_sim <= i && i < streams[g].first_sim + streams[g].n_sims)
        si = (int)g;
    else
        si = find_stream(streams, n_streams, i);
    const mcf_stream st = streams[si];
    const uint64_t s0 = (uint64_t)(i - st.first_sim) * (uint64_t)k;
    SlotWalker<KIND> w;
    w.init(st, s0);
    const SlotRule rule = slot_rule(span);
    int64_t sum = 0, run = 0, best = 0;
    bool bad = false;
    for (int32_t t = 0; t < k; t++) {
        uint32_t idx;
        bad |= !slot_accept(rule, w.get(s0 + (uint64_t)t), idx);
        const int64_t v = low + (int64_t)idx;
        sum += v;
        run = v == match ? run + 1 : 0;
        best = run > best ? run : best;
    }
    if (bad) atomicExch(rejected, 1);
    out[i] = (double)(op == 0 ? sum : best);
}

// =========================================================================================== lsm
struct LsmState {
    double sums[8];    // n, Su, Su2, Su3, Su4, Sy, Suy, Su2y   (u = (S-K)/K)
    double coef[3];    // fitted = coef0 + coef1*u + coef2*u^2
    int32_t have_fit;  // 0: no in-the-money path at this step (or fit skipped)
    int32_t pad;
    double price_sum;
};

struct LsmGeom {
    int64_t off_state, off_counter, off_partials, off_ex, off_cf, total;
    int n_blocks;
};
#define LSM_THREADS 256
static inline LsmGeom lsm_geom(int64_t n_paths) {
    LsmGeom g;
    g.n_blocks = grid_cap(n_paths, LSM_THREADS * 4, MCF_SM_COUNT * 8);
    int64_t o = 0;
    g.off_state = o;
    o = align256(o + 2 * (int64_t)sizeof(LsmState));  // fused path: coefficients double-buffered by step parity
    g.off_counter = o;
    o = align256(o + 8);
    g.off_partials = o;
    o = align256(o + (int64_t)g.n_blocks * 8 * 8);
    g.off_ex = o;
    o = align256(o + n_paths * 4);
    g.off_cf = o;
    o = align256(o + n_paths * 8);
    g.total = o;
    return g;
}

__device__ __forceinline__ double lsm_intrinsic(double s, double K, int is_put) {
    return is_put ? fmax(K - s, 0.0) : fmax(s - K, 0.0);
}

// cf[i] holds the cash flow DISCOUNTED TO TIME 0: intrinsic(exercise step) * exp(-r*dt*exercise step).  The
// regression target of step t is then cf[i]*exp(+r*dt*t) -- one scalar per step instead of one exp per path.
__global__ void __launch_bounds__(LSM_THREADS) lsm_init_kernel(const double *__restrict__ s_T, int64_t n_paths,
                                                               int32_t n_steps, double K, int is_put, double disc_T,
                                                               int32_t *__restrict__ ex, double *__restrict__ cf) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_paths; i += (int64_t)gridDim.x * blockDim.x) {
        ex[i] = n_steps;
        cf[i] = lsm_intrinsic(s_T[i], K, is_put) * disc_T;
    }
}

// sweep 1 of step t: sums of the normal equations over in-the-money paths (16 B per path: S_t and cf)
__global__ void __launch_bounds__(LSM_THREADS) lsm_reduce_kernel(const double *__restrict__ s_t, int64_t n_paths,
                                                                 double K, double grow_t, int is_put,
                                                                 const double *__restrict__ cf,
                                                                 double *__restrict__ partials) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 0.0;
    const double invK = 1.0 / K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_paths; i += (int64_t)gridDim.x * blockDim.x) {
        const double s = __ldg(&s_t[i]);
        const double c = __ldg(&cf[i]);
        if (lsm_intrinsic(s, K, is_put) > 0.0) {
            const double y = c * grow_t;
            const double u = (s - K) * invK, u2 = u * u;
            a[0] += 1.0;
            a[1] += u;
            a[2] += u2;
            a[3] = fma(u2, u, a[3]);
            a[4] = fma(u2, u2, a[4]);
            a[5] += y;
            a[6] = fma(u, y, a[6]);
            a[7] = fma(u2, y, a[7]);
        }
    }
    __shared__ double sh[LSM_THREADS / 32][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double v = a[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        for (int w2 = 0; w2 < LSM_THREADS / 32; w2++) v += sh[w2][threadIdx.x];
        partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) lsm_sum_partials_kernel(const double *__restrict__ partials, int n_blocks,
                                                               LsmState *__restrict__ st) {
    // 8 warps, warp k sums column k in a fixed order
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    double v = 0.0;
    for (int b = lane; b < n_blocks; b += 32) v += partials[(size_t)b * 8 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) st->sums[k] = v;
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi) -> pseudo-inverse solve with lstsq's rcond=eps*max(M,N) spirit
__device__ void lsm_solve3(const double G[3][3], const double rhs[3], double n_rows, double coef[3]) {
    // Fast path (every step of an at-the-money sweep): the rescaled Gram matrix is safely positive definite, so an
    // LDL^T factorisation -- three divisions, no square root, ~0.3 us -- gives the unique least-squares solution.
    // The eigen-decomposition below (a serial ~7 us on the one thread that runs it: half of the fixed cost of a step)
    // is only needed to reproduce lstsq's min-norm answer when the fit is rank deficient or nearly so; a pivot below
    // 1e-4 of the largest diagonal entry (condition number above ~1e4) sends the step there.
    {
        const double a00 = G[0][0], a01 = G[0][1], a02 = G[0][2], a11 = G[1][1], a12 = G[1][2], a22 = G[2][2];
        const double floor_ = 1e-4 * fmax(a00, fmax(a11, a22));
        if (a00 > floor_) {
            const double l10 = a01 / a00, l20 = a02 / a00;
            const double d1 = a11 - l10 * a01;
            if (d1 > floor_) {
                const double l21 = (a12 - l20 * a01) / d1;
                const double d2 = a22 - l20 * a02 - l21 * (a12 - l20 * a01);
                if (d2 > floor_) {
                    const double y0 = rhs[0], y1 = rhs[1] - l10 * y0, y2 = rhs[2] - l20 * y0 - l21 * y1;
                    const double z2 = y2 / d2;
                    const double z1 = y1 / d1 - l21 * z2;
                    const double z0 = y0 / a00 - l10 * z1 - l20 * z2;
                    coef[0] = z0;
                    coef[1] = z1;
                    coef[2] = z2;
                    return;
                }
            }
        }
    }
    double A[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            A[i][j] = G[i][j];
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; sweep++) {
        // converged when the off-diagonal mass is far below one ulp of the diagonal (further rotations are no-ops)
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off <= 1e-22 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]))) break;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = fmax(fmax(fabs(A[0][0]), fabs(A[1][1])), fabs(A[2][2]));
    // singular values of X are sqrt(eigenvalues of X^T X); lstsq drops s < eps*max(M,N)*s_max
    const double rc = 2.220446049250313e-16 * fmax(n_rows, 3.0);
    const double cutoff = lmax * rc * rc;
    coef[0] = coef[1] = coef[2] = 0.0;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const double lam = A[e][e];
        if (lam > cutoff && lam > 0.0) {
            const double proj = (V[0][e] * rhs[0] + V[1][e] * rhs[1] + V[2][e] * rhs[2]) / lam;
#pragma unroll
            for (int i = 0; i < 3; i++) coef[i] += V[i][e] * proj;
        }
    }
}

__device__ void lsm_solve_state(LsmState *st);
__global__ void lsm_solve_kernel(LsmState *st) {
    if (threadIdx.x != 0) return;
    lsm_solve_state(st);
}
__device__ void lsm_solve_state(LsmState *st) {
    const double *s = st->sums;
    if (s[0] <= 0.0) {
        st->have_fit = 0;
        return;
    }
    // rescale u by its rms so the Gram matrix is well conditioned; fitted values are basis independent
    const double scale = s[2] > 0.0 ? sqrt(s[2] / s[0]) : 1.0;
    const double i1 = 1.0 / scale, i2 = i1 * i1, i3 = i2 * i1, i4 = i2 * i2;
    const double G[3][3] = {{s[0], s[1] * i1, s[2] * i2}, {s[1] * i1, s[2] * i2, s[3] * i3}, {s[2] * i2, s[3] * i3, s[4] * i4}};
    const double rhs[3] = {s[5], s[6] * i1, s[7] * i2};
    double c[3];
    lsm_solve3(G, rhs, s[0], c);
    st->coef[0] = c[0];
    st->coef[1] = c[1] * i1;
    st->coef[2] = c[2] * i2;
    st->have_fit = 1;
}

// sweep 2 of step t: exercise where intrinsic > fitted continuation
__global__ void __launch_bounds__(LSM_THREADS) lsm_apply_kernel(const double *__restrict__ s_t, int64_t n_paths, int32_t t,
                                                                double K, int is_put, double disc_t,
                                                                const LsmState *__restrict__ st,
                                                                int32_t *__restrict__ ex, double *__restrict__ cf) {
    if (!st->have_fit) return;
    const double c0 = st->coef[0], c1 = st->coef[1], c2 = st->coef[2], invK = 1.0 / K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_paths; i += (int64_t)gridDim.x * blockDim.x) {
        const double s = __ldg(&s_t[i]);
        const double intr = lsm_intrinsic(s, K, is_put);
        if (intr > 0.0) {
            const double u = (s - K) * invK;
            const double fitted = fma(fma(c2, u, c1), u, c0);
            if (intr > fitted) {
                ex[i] = t;
                cf[i] = intr * disc_t;
            }
        }
    }
}

// ---------------------------------------------------------------------------------- fused single-GPU backward induction
// One launch per step t: every path applies the exercise decision of step t (coefficients fitted by the previous
// launch) and, with its updated cash flow, contributes to the normal equations of step t-1.  The CTA that finishes
// last adds the per-CTA partial sums in a fixed order (deterministic) and solves the 3x3 system for step t-1, so the
// serial part costs no launch.  Per path and step: S_t and S_{t-1} (8 B each), cf read (8 B), cf written only where
// the option is exercised.  t == n_steps is the initialisation (cf = discounted terminal intrinsic, no decision);
// t == 1 accumulates the price sum instead of regression sums.
// Paths sharded over the GPUs of one box: the 8 sums of a step are exchanged INSIDE the step kernel, over peer memory.
// Every rank owns a small exchange buffer that all ranks have mapped (CUDA IPC over NVLink): the CTA that finishes last
// stores its rank's 8 sums into slot [parity][rank] of EVERY rank's buffer (each 8-byte word tagged with the step's
// sequence number, see lsm_publish_word), waits until all slots of its own buffer carry this step's number, adds them in
// rank order (the same order on every rank: identical coefficients everywhere, deterministic) and solves.  No collective library call, no extra
// launch, no host: a step of the sharded sweep is the single-GPU step plus one NVLink round trip.  Ranks can be at most
// one step apart (nobody passes a step without everybody's sums), hence two slot sets by step parity.
#define MCF_LSM_MAX_PEERS 8
struct LsmPeers {
    double *slots[MCF_LSM_MAX_PEERS];  // exchange buffer of every rank, as mapped in THIS process
    int32_t rank, world;               // world == 1: no exchange
    unsigned long long seq_base;       // sequence numbers of this sweep are seq_base + 1 .. seq_base + n_steps
    double n_paths_global;
};
#define MCF_LSM_SLOT_DOUBLES 16  // 16 words of 8 bytes: half a double + the sequence number each (128 bytes)
__host__ __device__ inline size_t lsm_slot_index(int parity, int rank) {
    return ((size_t)parity * MCF_LSM_MAX_PEERS + (size_t)rank) * MCF_LSM_SLOT_DOUBLES;
}
// A slot is 16 words of 8 bytes, each carrying HALF a double and the low 32 bits of the step's sequence number (word 2k:
// low half of sum k, word 2k+1: high half).  An 8-byte store is atomic over NVLink, so a word whose upper half shows the
// expected sequence number carries valid data: no fence, no separate flag, one NVLink store latency per step (the form
// with 8 stores, a system-scope fence and a flag store cost a round trip more).  16 threads per peer store the 16 words
// at once; the same 16 x world threads then poll this rank's own buffer.
__device__ __forceinline__ void lsm_publish_word(const LsmPeers &px, int peer, int w, int parity, unsigned seq32,
                                                 const double *sums) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(sums[w >> 1]);
    const unsigned half = (w & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
    volatile unsigned long long *dst =
        reinterpret_cast<volatile unsigned long long *>(px.slots[peer] + lsm_slot_index(parity, px.rank)) + w;
    *dst = ((unsigned long long)seq32 << 32) | half;
}
// false: the word did not arrive within ~2 s (a peer died); the caller poisons the sums instead of hanging the GPU
__device__ __forceinline__ bool lsm_collect_word(const LsmPeers &px, int src, int w, int parity, unsigned seq32, unsigned *half) {
    const volatile unsigned long long *p =
        reinterpret_cast<const volatile unsigned long long *>(px.slots[px.rank] + lsm_slot_index(parity, src)) + w;
    const long long t0 = clock64();
    for (;;) {
        const unsigned long long v = *p;
        if ((unsigned)(v >> 32) == seq32) {
            *half = (unsigned)v;
            return true;
        }
        if (clock64() - t0 > (1LL << 32)) return false;
    }
}

// Sweep direction.  Column S_{t-1} is read here for the regression and again by the NEXT launch for its decision, and
// the cash flows are read by every launch: 16 of the 24 bytes per path and step were touched one launch earlier.  One
// launch's footprint (240 MB at 1e7 paths) exceeds the 126 MB L2, so a front-to-back sweep every step finds nothing left
// (LRU against a cyclic pattern).  Successive launches therefore sweep in OPPOSITE directions (`reverse`), in one
// co-resident wave of CTAs: what the previous launch touched last is what this one touches first, and S_t -- dead after
// this launch -- is loaded with the evict-first policy so that it does not push the reusable lines out.
#ifndef MCF_LSM_PPT
#define MCF_LSM_PPT 8  // paths per thread and loop trip (MCF_LSM_PPT / VEC loads of S_t, S_{t-1} and the cash flows each)
#endif
#ifndef MCF_LSM_CTAS
#define MCF_LSM_CTAS 2  // resident CTAs per SM (register budget 65536 / (256 * CTAS))
#endif
// A thread moves VEC paths per load: 2 (16-byte loads) when the path count is even and the columns are 16-byte aligned.
template <int VEC> struct LsmVec;
template <> struct LsmVec<1> {
    typedef double T;
    static __device__ __forceinline__ void unpack(const T &v, double *o) { o[0] = v; }
    static __device__ __forceinline__ T pack(const double *o) { return o[0]; }
};
template <> struct LsmVec<2> {
    typedef double2 T;
    static __device__ __forceinline__ void unpack(const T &v, double *o) {
        o[0] = v.x;
        o[1] = v.y;
    }
    static __device__ __forceinline__ T pack(const double *o) { return make_double2(o[0], o[1]); }
};
// L2 residency of the cash flows.  They are read by every step and a sixth of them is rewritten by every step (deep
// in-the-money paths exercise again and again on the way back), 80 MB at 1e7 paths: inside the 126 MB L2 they cost no HBM
// traffic at all, neither the reads nor the write-backs.  Plain LRU loses them to the 160 MB of path columns streaming
// through per step, so their loads and stores carry an evict-last policy for the first `keep_bytes` of the array (the
// size of the device's set-aside for persisting lines, lsm_l2_keep_bytes below) -- measured 10.36 -> 9.29 ms per sweep.
__device__ __forceinline__ unsigned long long lsm_keep_policy(const void *base, uint32_t keep_bytes, uint32_t total_bytes) {
    unsigned long long pol;
    asm volatile("createpolicy.range.global.L2::evict_last.L2::evict_unchanged.b64 %0, [%1], %2, %3;"
                 : "=l"(pol)
                 : "l"(base), "r"(keep_bytes), "r"(total_bytes));
    return pol;
}
__device__ __forceinline__ double lsm_ld_keep(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 lsm_ld_keep(const double2 *p, unsigned long long pol) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void lsm_st_keep(double *p, double v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void lsm_st_keep(double2 *p, double2 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
// The step kernel is issue-bound before it is HBM-bound (round-2 ncu: 107 thread instructions per path, issue slots 58 %
// busy at 52 % of DRAM peak), so everything that is the same for every path of a launch is a template parameter --
// MODE bit 0: t == n_steps (the cash flows are initialised, not read), bit 1: t == 1 (no regression, the cash flows are
// summed); HAS_EX: exercise steps are recorded -- the loop runs without bounds checks except for the one trip that holds
// the last (partial) band, and a path is straight-line code: the intrinsic value is +-(S - K) (a sign flip in the high
// word; "in the money" is `> 0`, no max()), out-of-the-money paths enter the regression sums as exact zeros, the count
// of in-the-money paths is an integer.  Values per path are those of the branchy form bit for bit.
template <int MODE, bool HAS_EX, int VEC, bool KEEP>
__global__ void __launch_bounds__(LSM_THREADS, MCF_LSM_CTAS)
    lsm_fused_step_kernel(const double *__restrict__ s_t, const double *__restrict__ s_prev, int64_t n_paths, int32_t t,
                          int32_t n_steps, double K, int is_put, double disc_t, double grow_prev,
                          const LsmState *__restrict__ st_cur, LsmState *__restrict__ st_next,
                          int32_t *__restrict__ ex, double *__restrict__ cf, double *__restrict__ partials,
                          unsigned int *__restrict__ counter, double *__restrict__ price_out, int reverse,
                          uint32_t keep_bytes, const __grid_constant__ LsmPeers peers) {
    constexpr bool init = (MODE & 1) != 0, last_step = (MODE & 2) != 0;
    typedef typename LsmVec<VEC>::T V;
    constexpr int TRIP = MCF_LSM_PPT / VEC;  // loads per array and trip
    // Programmatic dependent launch: this grid may start while the previous step's grid is still finishing (its last
    // CTA is adding the partial sums and solving).  The path columns do not depend on it, so the first trip's loads of
    // S_t / S_{t-1} are issued BEFORE `griddepcontrol.wait`; the coefficients and the cash flows are read after it.
#ifndef MCF_LSM_NO_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    bool waited = false;
    double c0 = __longlong_as_double(0x7ff0000000000000LL), c1 = 0.0, c2 = 0.0;  // no fit: nothing exceeds +inf
    auto wait_for_previous_step = [&]() {
        if (waited) return;
        waited = true;
#ifndef MCF_LSM_NO_PDL
        asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
        if (!init && st_cur->have_fit) {
            c0 = st_cur->coef[0];
            c1 = st_cur->coef[1];
            c2 = st_cur->coef[2];
        }
    };
    const double invK = 1.0 / K;
    const int flip = is_put ? (int)0x80000000 : 0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0;
    int n_itm = 0;
    // one path: decide step t (c = its cash flow discounted to time 0), then its terms of the regression of step t-1
    auto one = [&](double s, double sp, double &c) -> bool {
        const double d = s - K;
        const double gap = __hiloint2double(__double2hiint(d) ^ flip, __double2loint(d));  // K - S for a put
        bool exercised;
        if (init) {
            exercised = true;
            c = gap > 0.0 ? gap * disc_t : 0.0;
        } else {
            const double u = d * invK;
            const double fitted = fma(fma(c2, u, c1), u, c0);
            exercised = gap > 0.0 && gap > fitted;
            c = exercised ? gap * disc_t : c;
        }
        if (last_step) {
            a0 += c;
        } else {
            const double dp = sp - K;
            const double gap_p = __hiloint2double(__double2hiint(dp) ^ flip, __double2loint(dp));
            const bool itm = gap_p > 0.0;
            const double u = itm ? dp * invK : 0.0, y = itm ? c * grow_prev : 0.0, u2 = u * u;
            n_itm += itm ? 1 : 0;
            a1 += u;
            a2 += u2;
            a3 = fma(u2, u, a3);
            a4 = fma(u2, u2, a4);
            a5 += y;
            a6 = fma(u, y, a6);
            a7 = fma(u2, y, a7);
        }
        return exercised;
    };
    // bands of `stride` units (a unit = VEC neighbouring paths); a thread owns one unit of every band.  Forward: band 0,
    // 1, ...; reverse: last band first.
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t lane_i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_units = n_paths / VEC;
    const int64_t bands = (n_units + stride - 1) / stride;
    const V *sv_t = reinterpret_cast<const V *>(s_t), *sv_prev = reinterpret_cast<const V *>(s_prev);
    V *cfv = reinterpret_cast<V *>(cf);
    // (a policy-hinted access is never the default: with an evict-normal policy the sweep at 2e7 paths took 39 ms, not 21)
    const unsigned long long keep = KEEP ? lsm_keep_policy(cf, keep_bytes, (uint32_t)(n_paths * 8)) : 0ULL;
    auto trip = [&](auto checked_tag, int64_t lo_band) {
        constexpr bool CHECKED = decltype(checked_tag)::value;
        V s[TRIP], p[TRIP], c[TRIP];
        int64_t unit[TRIP];
        bool on[TRIP];
#pragma unroll
        for (int q = 0; q < TRIP; q++) {  // all loads of the trip are issued before the first use
            const int64_t band = lo_band + q;
            unit[q] = band * stride + lane_i;
            on[q] = !CHECKED || (band >= 0 && band < bands && unit[q] < n_units);
            if (on[q]) {
                s[q] = __ldcs(sv_t + unit[q]);  // S_t is dead after this launch: evict-first
                if (!last_step) p[q] = KEEP ? __ldcg(sv_prev + unit[q]) : __ldg(sv_prev + unit[q]);
            }
        }
        wait_for_previous_step();
        if (!init) {
#pragma unroll
            for (int q = 0; q < TRIP; q++)
                if (on[q]) c[q] = KEEP ? lsm_ld_keep(cfv + unit[q], keep) : cfv[unit[q]];
        }
#pragma unroll
        for (int q = 0; q < TRIP; q++) {
            if (!on[q]) continue;
            double se[VEC], pe[VEC], ce[VEC];
            LsmVec<VEC>::unpack(s[q], se);
            if (!last_step) LsmVec<VEC>::unpack(p[q], pe);
            if (!init) LsmVec<VEC>::unpack(c[q], ce);
            bool any = false;
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const bool exercised = one(se[e], last_step ? 0.0 : pe[e], ce[e]);
                any |= exercised;
                if (HAS_EX && exercised) ex[unit[q] * VEC + e] = t;
            }
            if (any) {  // (an unchanged neighbour is rewritten with its own value)
                if (KEEP)
                    lsm_st_keep(cfv + unit[q], LsmVec<VEC>::pack(ce), keep);
                else
                    cfv[unit[q]] = LsmVec<VEC>::pack(ce);
            }
        }
    };
    for (int64_t b0 = 0; b0 < bands; b0 += TRIP) {
        const int64_t lo_band = reverse ? bands - TRIP - b0 : b0;
        if (lo_band >= 0 && lo_band + TRIP < bands)  // every band of the trip is complete
            trip(std::false_type(), lo_band);
        else
            trip(std::true_type(), lo_band);
    }
    wait_for_previous_step();  // (a grid with no band still orders itself behind the previous step)
    double a[8] = {last_step ? a0 : (double)n_itm, a1, a2, a3, a4, a5, a6, a7};

    __shared__ double sh[LSM_THREADS / 32][8];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double v = a[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        for (int w2 = 0; w2 < LSM_THREADS / 32; w2++) v += sh[w2][threadIdx.x];
        partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    {  // 8 warps, warp k sums column k in a fixed order (as lsm_sum_partials_kernel); the loads of a group of 16 are
       // issued together -- this CTA is the serial tail of the step, a dependent load per addition cost ~4 us
        double v = 0.0;
        for (int b0 = lane; b0 < (int)gridDim.x; b0 += 32 * 16) {
            double part[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int b = b0 + 32 * j;
                part[j] = b < (int)gridDim.x ? __ldcg(&partials[(size_t)b * 8 + wid]) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 16; j++) v += part[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) st_next->sums[wid] = v;
    }
    __syncthreads();
    if (peers.world > 1) {  // the ranks' sums meet here: publish to every rank, collect from every rank, add in rank order
        __shared__ unsigned sh_half[MCF_LSM_MAX_PEERS][16];
        __shared__ int sh_lost;
        const unsigned seq32 = (unsigned)(peers.seq_base + (unsigned long long)(n_steps - t + 1));
        const int peer = (int)threadIdx.x >> 4, w = (int)threadIdx.x & 15;
        if (threadIdx.x == 0) sh_lost = 0;
        __syncthreads();
        if (peer < peers.world) {
            lsm_publish_word(peers, peer, w, t & 1, seq32, st_next->sums);
            unsigned half = 0;
            if (!lsm_collect_word(peers, peer, w, t & 1, seq32, &half)) sh_lost = 1;
            sh_half[peer][w] = half;
        }
        __syncthreads();
        if (threadIdx.x < 8) {  // the same order on every rank: identical coefficients everywhere
            double total = 0.0;
            for (int src = 0; src < peers.world; src++)
                total += __hiloint2double((int)sh_half[src][2 * threadIdx.x + 1], (int)sh_half[src][2 * threadIdx.x]);
            st_next->sums[threadIdx.x] = sh_lost ? __longlong_as_double(0x7ff8000000000000LL) : total;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *counter = 0;
        const double n_all = peers.world > 1 ? peers.n_paths_global : (double)n_paths;
        if (last_step) {
            if (price_out) *price_out = st_next->sums[0] / n_all;
        } else {
            lsm_solve_state(st_next);
        }
    }
}

// (Measured and rejected: the whole sweep as ONE cooperative launch -- grid-wide barrier per step, every CTA adding the
// partials and solving the 3x3 system for itself.  16.6 ms against 14.7 ms for one launch per step at 1e7 x 252: the
// barrier and the redundant serial solve cost more than the launch gaps they replace.)
// discounted cash flows: per-block partial sums (column 0 of partials)
__global__ void __launch_bounds__(LSM_THREADS) lsm_price_kernel(int64_t n_paths, const double *__restrict__ cf,
                                                                double *__restrict__ partials) {
    double a = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_paths; i += (int64_t)gridDim.x * blockDim.x)
        a += cf[i];
    __shared__ double sh[LSM_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        if (threadIdx.x == 0)
            for (int w2 = 0; w2 < LSM_THREADS / 32; w2++) v += sh[w2];
        partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
}

__global__ void lsm_price_final_kernel(const LsmState *st, double n_paths, double *price_sum_out, double *price_out) {
    if (threadIdx.x == 0) {
        if (price_sum_out) *price_sum_out = st->sums[0];
        if (price_out) *price_out = st->sums[0] / n_paths;
    }
}

// (n_rows, n_cols) row-major -> (n_cols, n_rows) row-major, 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ in, int64_t n_rows, int64_t n_cols,
                                                        int64_t tiles_x, double *__restrict__ out) {
    __shared__ double tile[32][33];
    const int64_t tile_id = blockIdx.x;  // 1-D grid: either dimension may exceed the 65535 limit of grid.y
    const int64_t r0 = (tile_id / tiles_x) * 32, c0 = (tile_id % tiles_x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = r0 + k, c = c0 + tx;
        if (r < n_rows && c < n_cols) tile[k][tx] = in[r * n_cols + c];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t c = c0 + k, r = r0 + tx;
        if (r < n_rows && c < n_cols) out[c * n_rows + r] = tile[tx][k];
    }
}


// ================================================================================ small helpers
// number of elements strictly below v (NaN compares false, like NumPy's `means < m_hat`)
__global__ void __launch_bounds__(256) count_less_kernel(const double *__restrict__ x, int64_t n, double v,
                                                         unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += (__ldg(&x[i]) < v) ? 1ULL : 0ULL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// order-preserving compaction of the finite values (nan_policy="omit"): tile = 1024 elements per CTA
#define COMPACT_TILE 1024
__global__ void __launch_bounds__(256) compact_count_kernel(const double *__restrict__ x, int64_t n,
                                                            unsigned long long *__restrict__ tile_count) {
    const int64_t base = (int64_t)blockIdx.x * COMPACT_TILE;
    unsigned c = 0;
    for (int k = 0; k < 4; k++) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < n) c += bits_nonfinite(double_bits(x[i])) ? 0u : 1u;
    }
    __shared__ unsigned sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < 8; w++) t += sh[w];
        tile_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) compact_scatter_kernel(const double *__restrict__ x, int64_t n,
                                                              const unsigned long long *__restrict__ tile_base,
                                                              double *__restrict__ out) {
    // each thread owns 4 CONSECUTIVE elements so that the in-tile order is preserved by a plain scan
    const int64_t base = (int64_t)blockIdx.x * COMPACT_TILE + (int64_t)threadIdx.x * 4;
    double v[4];
    bool keep[4];
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t i = base + k;
        keep[k] = false;
        v[k] = 0.0;
        if (i < n) {
            v[k] = x[i];
            keep[k] = !bits_nonfinite(double_bits(v[k]));
        }
        c += keep[k] ? 1u : 0u;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    __shared__ unsigned sh[8];
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < wid; w++) wbase += sh[w];
    unsigned long long pos = tile_base[blockIdx.x] + wbase + (incl - c);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (keep[k]) out[pos++] = v[k];
}

// ========================================================================================= C ABI
extern "C" {

int64_t mcf_moments_workspace_bytes(int64_t n) {
    if (n < 0) return MCF_ERR_BAD_ARG;
    return align256((int64_t)sizeof(MomentsPartial) * MCF_SM_COUNT * 8);
}

int mcf_moments(const double *d_x, int64_t n, double target, double *d_out, void *d_workspace, int64_t workspace_bytes,
                void *cuda_stream) {
    if (!d_x || n <= 0 || !d_out || !d_workspace) return MCF_ERR_BAD_ARG;
    if (workspace_bytes < mcf_moments_workspace_bytes(n)) return MCF_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    const int blocks = grid_cap(n, 256 * 8, MCF_SM_COUNT * 8);
    moments_kernel<<<blocks, 256, 0, cs>>>(d_x, n, target, (MomentsPartial *)d_workspace);
    MCF_CUDA_OK(cudaGetLastError());
    moments_final_kernel<<<1, 256, 0, cs>>>((const MomentsPartial *)d_workspace, blocks, target, d_out);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int64_t mcf_select_workspace_bytes(void) { return select_geom().total; }
int64_t mcf_select_hist_offset(void) { return select_geom().off_hist; }
int64_t mcf_select_hist_bytes(void) { return select_geom().hist_bytes; }
int32_t mcf_select_passes(void) { return SEL_PASSES; }

int mcf_select_begin(const int64_t *h_ranks, int32_t k, void *d_workspace, int64_t workspace_bytes, void *cuda_stream) {
    if (!h_ranks || k < 1 || k > MCF_SELECT_MAX_RANKS || !d_workspace) return MCF_ERR_BAD_ARG;
    const SelectGeom g = select_geom();
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    SelectState init;
    memset(&init, 0, sizeof(init));
    init.k = k;
    init.n_uniq = 1;
    for (int r = 0; r < k; r++) {
        if (h_ranks[r] < 0) return MCF_ERR_BAD_ARG;
        init.rank[r] = h_ranks[r];
    }
    char *ws = (char *)d_workspace;
    select_init_kernel<<<1, 256, 0, (cudaStream_t)cuda_stream>>>((SelectState *)(ws + g.off_state),
                                                                 (unsigned long long *)(ws + g.off_hist), init);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

static int select_hist_launch(const double *d_x, int64_t n, int32_t omit_nonfinite, void *d_workspace, const uint64_t *keys,
                              int64_t keys_cap, void *cuda_stream) {
    const SelectGeom g = select_geom();
    char *ws = (char *)d_workspace;
    const int blocks = grid_cap(n, 512 * 8, MCF_SM_COUNT * 4);
    select_hist_kernel<<<blocks, 512, 0, (cudaStream_t)cuda_stream>>>(d_x, n, omit_nonfinite,
                                                                      (const SelectState *)(ws + g.off_state),
                                                                      (unsigned long long *)(ws + g.off_hist), keys, keys_cap);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int mcf_select_hist(const double *d_x, int64_t n, int32_t omit_nonfinite, void *d_workspace, void *cuda_stream) {
    if (!d_x || n <= 0 || !d_workspace) return MCF_ERR_BAD_ARG;
    return select_hist_launch(d_x, n, omit_nonfinite, d_workspace, nullptr, 0, cuda_stream);
}

int mcf_select_choose(void *d_workspace, void *cuda_stream) {
    if (!d_workspace) return MCF_ERR_BAD_ARG;
    const SelectGeom g = select_geom();
    char *ws = (char *)d_workspace;
    select_choose_kernel<<<1, 1024, 0, (cudaStream_t)cuda_stream>>>((SelectState *)(ws + g.off_state),
                                                                   (unsigned long long *)(ws + g.off_hist));
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int mcf_select_finish(const void *d_workspace, double *d_out, void *cuda_stream) {
    if (!d_workspace || !d_out) return MCF_ERR_BAD_ARG;
    const SelectGeom g = select_geom();
    select_finish_kernel<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(
        (const SelectState *)((const char *)d_workspace + g.off_state), d_out);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int64_t mcf_select_compact_bytes(int64_t n) { return n <= 0 ? MCF_ERR_BAD_ARG : align256((n / 8 + 1024) * 8); }

// d_compact (may be NULL): scratch of mcf_select_compact_bytes(n) bytes.  With it the candidates left after the second
// pass are compacted and passes 3..8 read only those.
int mcf_select_compacting(const double *d_x, int64_t n, int32_t omit_nonfinite, const int64_t *h_ranks, int32_t k,
                          double *d_out, void *d_workspace, int64_t workspace_bytes, void *d_compact, int64_t compact_bytes,
                          void *cuda_stream) {
    if (!d_x || n <= 0) return MCF_ERR_BAD_ARG;
    int rc = mcf_select_begin(h_ranks, k, d_workspace, workspace_bytes, cuda_stream);
    if (rc) return rc;
    const SelectGeom g = select_geom();
    SelectState *st = (SelectState *)((char *)d_workspace + g.off_state);
    const uint64_t *keys = (const uint64_t *)d_compact;
    const int64_t cap = d_compact ? compact_bytes / 8 : 0;
    for (int p = 0; p < SEL_PASSES; p++) {
        if (p == 2 && cap > 0) {
            select_compact_kernel<<<grid_cap(n, 512 * 8, MCF_SM_COUNT * 4), 512, 0, (cudaStream_t)cuda_stream>>>(
                d_x, n, omit_nonfinite, st, (uint64_t *)d_compact, cap);
            select_compact_done_kernel<<<1, 1, 0, (cudaStream_t)cuda_stream>>>(st);
            MCF_CUDA_OK(cudaGetLastError());
        }
        rc = select_hist_launch(d_x, n, omit_nonfinite, d_workspace, p >= 2 ? keys : nullptr, cap, cuda_stream);
        if (rc) return rc;
        rc = mcf_select_choose(d_workspace, cuda_stream);
        if (rc) return rc;
    }
    return mcf_select_finish(d_workspace, d_out, cuda_stream);
}

int64_t mcf_select_bracket_bytes(int64_t n) {
    if (n <= 0) return MCF_ERR_BAD_ARG;
    const SelectGeom g = select_geom();
    // candidate keys | sample | second select workspace | bracket values | open flags | bracket set | flag
    return align256((n / 8 + 1024) * 8) + align256((int64_t)MCF_SELECT_SAMPLE * 8) + g.total + align256(2 * MCF_SELECT_MAX_RANKS * 8) +
           align256(2 * MCF_SELECT_MAX_RANKS * 4) + align256((int64_t)sizeof(BracketSet)) + 256;
}

// Exact order statistics of a LARGE sample without NaN handling (omit_nonfinite = 0 semantics; NaNs, if any, sort last):
// h_ranks ascending, k <= 16.  *d_invalid (device int) != 0 afterwards means "not decided here": the caller runs
// mcf_select_compacting instead (ties heavier than the scratch, more than 8 disjoint brackets, a 6-sigma sampling miss).
int mcf_select_bracketed(const double *d_x, int64_t n, const int64_t *h_ranks, int32_t k, double *d_out, void *d_workspace,
                         int64_t workspace_bytes, void *d_scratch, int64_t scratch_bytes, int32_t *d_invalid,
                         void *cuda_stream) {
    if (!d_x || !h_ranks || !d_out || !d_workspace || !d_scratch || !d_invalid || k < 1 || 2 * k > MCF_SELECT_MAX_RANKS ||
        n < 4 * (int64_t)MCF_SELECT_SAMPLE)
        return MCF_ERR_BAD_ARG;
    if (scratch_bytes < mcf_select_bracket_bytes(n)) return MCF_ERR_WORKSPACE_TOO_SMALL;
    if (n / ((int64_t)grid_cap(n, 512 * 8, MCF_SM_COUNT * 4) * 512) >= 65000) return MCF_ERR_BAD_ARG;  // 16-bit per-thread counters
    for (int r = 0; r < k; r++)
        if (h_ranks[r] < 0 || h_ranks[r] >= n || (r > 0 && h_ranks[r] < h_ranks[r - 1])) return MCF_ERR_BAD_ARG;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    const SelectGeom g = select_geom();
    char *sc = (char *)d_scratch;
    const int64_t cap = n / 8 + 1024;
    uint64_t *keys = (uint64_t *)sc;
    sc += align256(cap * 8);
    double *sample = (double *)sc;
    sc += align256((int64_t)MCF_SELECT_SAMPLE * 8);
    void *ws2 = sc;
    sc += g.total;
    double *bvals = (double *)sc;
    sc += align256(2 * MCF_SELECT_MAX_RANKS * 8);
    int32_t *open_flags = (int32_t *)sc;
    sc += align256(2 * MCF_SELECT_MAX_RANKS * 4);
    BracketSet *bs = (BracketSet *)sc;

    // 1. strided sample and the sample ranks that bracket every target
    const int64_t m = MCF_SELECT_SAMPLE, stride = n / m;
    select_sample_kernel<<<(unsigned)((m + 255) / 256), 256, 0, cs>>>(d_x, stride, m, sample);
    int64_t sranks[MCF_SELECT_MAX_RANKS];
    int32_t h_open[2 * MCF_SELECT_MAX_RANKS];
    for (int r = 0; r < k; r++) {
        const double q = (double)h_ranks[r] / (double)n;
        const double pos = q * (double)m;
        const double delta = 6.0 * sqrt((double)m * q * (1.0 - q)) + 8.0;
        int64_t lo = (int64_t)floor(pos - delta), hi = (int64_t)ceil(pos + delta);
        h_open[r] = lo <= 0;          // nothing of the sample below: the bracket is open towards -inf
        h_open[k + r] = hi >= m - 1;  // ... towards +inf (NaN keys included)
        sranks[r] = lo < 0 ? 0 : lo;
        sranks[k + r] = hi > m - 1 ? m - 1 : hi;
    }
    // the 2k sample order statistics, ascending per half (targets ascend, so do their brackets)
    int rc = mcf_select(sample, m, 0, sranks, 2 * k, bvals, ws2, g.total, cuda_stream);
    if (rc) return rc;
    MCF_CUDA_OK(cudaMemcpyAsync(open_flags, h_open, sizeof(int32_t) * 2 * (size_t)k, cudaMemcpyHostToDevice, cs));
    MCF_CUDA_OK(cudaStreamSynchronize(cs));  // h_open / sranks live on this stack frame
    select_bracket_build_kernel<<<1, 32, 0, cs>>>(bvals, k, open_flags, open_flags + k, bs);
    // 2. one pass over x: exact counts below / inside every bracket, inside keys appended
    rc = mcf_select_begin(h_ranks, k, d_workspace, workspace_bytes, cuda_stream);
    if (rc) return rc;
    SelectState *st = (SelectState *)((char *)d_workspace + g.off_state);
    select_bracket_pass_kernel<<<grid_cap(n, 512 * 8, MCF_SM_COUNT * 4), 512, 0, cs>>>(d_x, n, bs, st, keys, cap);
    select_bracket_rebase_kernel<<<1, 32, 0, cs>>>(bs, st, cap, d_invalid);
    MCF_CUDA_OK(cudaGetLastError());
    // 3. radix passes over the candidate keys only
    for (int p = 0; p < SEL_PASSES; p++) {
        rc = select_hist_launch(d_x, n, 0, d_workspace, keys, cap, cuda_stream);
        if (rc) return rc;
        rc = mcf_select_choose(d_workspace, cuda_stream);
        if (rc) return rc;
    }
    return mcf_select_finish(d_workspace, d_out, cuda_stream);
}

int mcf_select(const double *d_x, int64_t n, int32_t omit_nonfinite, const int64_t *h_ranks, int32_t k, double *d_out,
               void *d_workspace, int64_t workspace_bytes, void *cuda_stream) {
    return mcf_select_compacting(d_x, n, omit_nonfinite, h_ranks, k, d_out, d_workspace, workspace_bytes, nullptr, 0,
                                 cuda_stream);
}

// Device-wide hint: how many bytes L2 fetches from DRAM on a miss (32, 64 or 128).  Random gathers want 32.
int mcf_set_l2_fetch_granularity(int32_t bytes) {
    MCF_CUDA_OK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return MCF_OK;
}

int64_t mcf_bootstrap_workspace_bytes(int64_t n_slots, int32_t chunk_slots) {
    if (n_slots <= 0 || chunk_slots < 8) return MCF_ERR_BAD_ARG;
    return boot_geom(n_slots, chunk_slots).total;
}

int mcf_bootstrap_count(uint32_t gen_kind, const mcf_stream *d_stream, int64_t n, uint64_t slot_lo, uint64_t slot_hi,
                        int32_t chunk_slots, void *d_workspace, int64_t workspace_bytes, uint64_t *d_total,
                        void *cuda_stream) {
    if (!d_stream || n < 2 || n > 0xFFFFFFFFLL || slot_hi <= slot_lo || chunk_slots < 8 || !d_workspace)
        return MCF_ERR_BAD_ARG;
    const BootGeom g = boot_geom((int64_t)(slot_hi - slot_lo), chunk_slots);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    char *ws = (char *)d_workspace;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    uint32_t *counts = (uint32_t *)(ws + g.off_counts);
    unsigned long long *bsum = (unsigned long long *)(ws + g.off_bsum), *bbase = (unsigned long long *)(ws + g.off_bbase);
    unsigned long long *total = (unsigned long long *)(ws + g.off_total);
    if (gen_kind == MCF_GEN_PCG64)
        boot_count_kernel<MCF_GEN_PCG64><<<(unsigned)g.n_blocks, BOOT_THREADS, 0, cs>>>(
            d_stream, (uint32_t)n, slot_lo, slot_hi, chunk_slots, g.n_chunks, counts, bsum);
    else
        boot_count_kernel<MCF_GEN_PHILOX><<<(unsigned)g.n_blocks, BOOT_THREADS, 0, cs>>>(
            d_stream, (uint32_t)n, slot_lo, slot_hi, chunk_slots, g.n_chunks, counts, bsum);
    MCF_CUDA_OK(cudaGetLastError());
    scan_u64_kernel<<<1, 1024, 0, cs>>>(bsum, g.n_blocks, bbase, total);
    MCF_CUDA_OK(cudaGetLastError());
    if (d_total) MCF_CUDA_OK(cudaMemcpyAsync(d_total, total, 8, cudaMemcpyDeviceToDevice, cs));
    return MCF_OK;
}

int mcf_bootstrap_accumulate(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n, int64_t n_resamples,
                             uint64_t slot_lo, uint64_t slot_hi, int32_t chunk_slots, uint64_t ordinal_base,
                             const void *d_workspace, int64_t workspace_bytes, double *d_sums, uint64_t *d_end_slot,
                             void *cuda_stream) {
    if (!d_stream || !d_x || n < 2 || n > 0xFFFFFFFFLL || n_resamples <= 0 || slot_hi <= slot_lo || chunk_slots < 8 ||
        !d_workspace || !d_sums || !d_end_slot)
        return MCF_ERR_BAD_ARG;
    const BootGeom g = boot_geom((int64_t)(slot_hi - slot_lo), chunk_slots);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    const char *ws = (const char *)d_workspace;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    const uint32_t *counts = (const uint32_t *)(ws + g.off_counts);
    const unsigned long long *bbase = (const unsigned long long *)(ws + g.off_bbase);
    if (gen_kind == MCF_GEN_PCG64)
        boot_accumulate_kernel<MCF_GEN_PCG64><<<(unsigned)g.n_blocks, BOOT_THREADS, 0, cs>>>(
            d_stream, d_x, (uint32_t)n, (uint64_t)n_resamples, slot_lo, slot_hi, chunk_slots, g.n_chunks, ordinal_base,
            counts, bbase, d_sums, (unsigned long long *)d_end_slot);
    else
        boot_accumulate_kernel<MCF_GEN_PHILOX><<<(unsigned)g.n_blocks, BOOT_THREADS, 0, cs>>>(
            d_stream, d_x, (uint32_t)n, (uint64_t)n_resamples, slot_lo, slot_hi, chunk_slots, g.n_chunks, ordinal_base,
            counts, bbase, d_sums, (unsigned long long *)d_end_slot);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

// Binned variant of mcf_bootstrap_accumulate for populations that do not fit L2 (see boot_bin_kernel).
// h_block_base: HOST copy of the per-CTA ordinal prefix (mcf_bootstrap_block_base copies it out after the count pass).
// region_shift: log2(elements per region), 0 = choose (2^22 elements = 32 MB -- half of one L2 partition --, at most 32 regions).
int64_t mcf_bootstrap_blocks(int64_t n_slots, int32_t chunk_slots) {
    if (n_slots <= 0 || chunk_slots < 8) return MCF_ERR_BAD_ARG;
    return boot_geom(n_slots, chunk_slots).n_blocks;
}

int mcf_bootstrap_block_base(const void *d_workspace, int64_t n_slots, int32_t chunk_slots, uint64_t *h_block_base,
                             void *cuda_stream) {
    if (!d_workspace || !h_block_base || n_slots <= 0 || chunk_slots < 8) return MCF_ERR_BAD_ARG;
    const BootGeom g = boot_geom(n_slots, chunk_slots);
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    MCF_CUDA_OK(cudaMemcpyAsync(h_block_base, (const char *)d_workspace + g.off_bbase, (size_t)g.n_blocks * 8,
                                cudaMemcpyDeviceToHost, cs));
    MCF_CUDA_OK(cudaStreamSynchronize(cs));
    return MCF_OK;
}

static inline uint32_t bin_region_shift(int64_t n, int32_t requested) {
    if (requested > 0) return (uint32_t)requested;
    uint32_t s = 22;
    while (((n + (1LL << s) - 1) >> s) > BIN_MAX_REGIONS) s++;
    return s;
}

// Two queue sets: while the gather kernels of batch k drain one (L2-request bound, the SMs mostly idle), the bin kernel of
// batch k+1 fills the other on the caller's stream (instruction bound).
static inline int64_t bin_set_bytes(int64_t window, int64_t R, int64_t cap) {
    return align256(window * R * cap * 4) + align256(window * R * 4);
}
int64_t mcf_bootstrap_binned_workspace_bytes(int64_t n, int32_t resamples_per_batch, int32_t region_shift) {
    if (n < 2 || resamples_per_batch < 1) return MCF_ERR_BAD_ARG;
    const uint32_t sh = bin_region_shift(n, region_shift);
    const int64_t R = (n + (1LL << sh) - 1) >> sh;
    if (R > BIN_MAX_REGIONS) return MCF_ERR_BAD_ARG;
    const int64_t region_len = 1LL << sh;
    const int64_t cap = region_len + 8 * (int64_t)sqrt((double)region_len) + 4096;
    const int64_t window = resamples_per_batch + 2;
    return 2 * bin_set_bytes(window, R, cap) + 256;
}

// A high-priority side stream and the events that order it against the caller's stream; released on scope exit.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_bin[2] = {nullptr, nullptr}, ev_gat[2] = {nullptr, nullptr};
    bool open(cudaStream_t caller) {
        int lo_pri = 0, hi_pri = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri) != cudaSuccess) return false;
        if (cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, hi_pri) != cudaSuccess) return false;
        for (int b = 0; b < 2; b++)
            if (cudaEventCreateWithFlags(&ev_bin[b], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&ev_gat[b], cudaEventDisableTiming) != cudaSuccess)
                return false;
        // everything queued on the caller's stream so far (x, the zeroed sums) precedes the side stream's work
        return cudaEventRecord(ev_bin[0], caller) == cudaSuccess && cudaStreamWaitEvent(stream, ev_bin[0], 0) == cudaSuccess;
    }
    ~SideStream() {
        for (int b = 0; b < 2; b++) {
            if (ev_bin[b]) cudaEventDestroy(ev_bin[b]);
            if (ev_gat[b]) cudaEventDestroy(ev_gat[b]);
        }
        if (stream) cudaStreamDestroy(stream);
    }
};

int mcf_bootstrap_accumulate_binned(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n,
                                    int64_t n_resamples, uint64_t slot_lo, uint64_t slot_hi, int32_t chunk_slots,
                                    uint64_t ordinal_base, const void *d_workspace, int64_t workspace_bytes,
                                    const uint64_t *h_block_base, int32_t resamples_per_batch, int32_t region_shift,
                                    void *d_bin_workspace, int64_t bin_workspace_bytes, double *d_sums,
                                    uint64_t *d_end_slot, void *cuda_stream) {
    if (!d_stream || !d_x || n < 2 || n > 0xFFFFFFFFLL || n_resamples <= 0 || slot_hi <= slot_lo || chunk_slots < 8 ||
        (chunk_slots % 8) || !d_workspace || !h_block_base || resamples_per_batch < 1 || !d_bin_workspace || !d_sums ||
        !d_end_slot || (int64_t)BOOT_THREADS * chunk_slots > n)
        return MCF_ERR_BAD_ARG;
    const BootGeom g = boot_geom((int64_t)(slot_hi - slot_lo), chunk_slots);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    if (bin_workspace_bytes < mcf_bootstrap_binned_workspace_bytes(n, resamples_per_batch, region_shift))
        return MCF_ERR_WORKSPACE_TOO_SMALL;
    const uint32_t sh = bin_region_shift(n, region_shift);
    const int64_t R = (n + (1LL << sh) - 1) >> sh;
    const int64_t region_len = 1LL << sh;
    const int64_t cap = region_len + 8 * (int64_t)sqrt((double)region_len) + 4096;
    const int64_t window = resamples_per_batch + 2;
    char *bws = (char *)d_bin_workspace;
    const int64_t set_bytes = bin_set_bytes(window, R, cap);
    uint32_t *queues[2];
    unsigned int *tails[2];
    for (int b = 0; b < 2; b++) {
        queues[b] = (uint32_t *)(bws + b * set_bytes);
        tails[b] = (unsigned int *)(bws + b * set_bytes + align256(window * R * cap * 4));
    }
    int *overflow = (int *)(bws + 2 * set_bytes);
    const char *ws = (const char *)d_workspace;
    const uint32_t *counts = (const uint32_t *)(ws + g.off_counts);
    const unsigned long long *bbase = (const unsigned long long *)(ws + g.off_bbase);
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    MCF_CUDA_OK(cudaMemsetAsync(overflow, 0, 4, cs));
    const uint64_t total_needed = (uint64_t)n_resamples * (uint64_t)n;

    // The gathers run on a side stream of their own (created per call: the entry point stays re-entrant). Measured at
    // n = 1e8, 128 resamples: 111.6 ms serial, 103.7 ms overlapped; capping the bin kernel below 4 CTAs per SM to leave
    // room for gather CTAs made it slower (150-164 ms), so the bin kernel keeps its full occupancy.
    const bool overlap = true;
    SideStream ss;
    if (overlap && !ss.open(cs)) return MCF_ERR_CUDA;  // the destructor releases whatever was created
    const cudaStream_t side = overlap ? ss.stream : cs;
    cudaEvent_t *ev_bin = ss.ev_bin, *ev_gat = ss.ev_gat;
    // batches = runs of CTAs whose slots start inside `resamples_per_batch` consecutive resamples
    int64_t blk = 0;
    int k = 0;
    int rc = MCF_OK;
    while (blk < g.n_blocks) {
        const uint64_t ord0 = ordinal_base + h_block_base[blk];
        if (ord0 >= total_needed) break;
        const uint64_t b_first = ord0 / (uint64_t)n;
        int64_t end = blk + 1;
        while (end < g.n_blocks && (ordinal_base + h_block_base[end]) / (uint64_t)n < b_first + (uint64_t)resamples_per_batch &&
               ordinal_base + h_block_base[end] < total_needed)
            end++;
        BinParams bp;
        bp.n = (uint32_t)n;
        bp.region_shift = sh;
        bp.n_regions = (uint32_t)R;
        bp.n_res_window = (uint32_t)window;
        bp.b_first = b_first;
        bp.cap = (uint64_t)cap;
        bp.n_resamples = (uint64_t)n_resamples;
        const int buf = overlap ? (k & 1) : 0;
        if (overlap && k >= 2) cudaStreamWaitEvent(cs, ev_gat[buf], 0);  // this queue set has been drained
        cudaMemsetAsync(tails[buf], 0, (size_t)(window * R * 4), cs);
        if (gen_kind == MCF_GEN_PCG64)
            boot_bin_kernel<MCF_GEN_PCG64><<<(unsigned)(end - blk), BOOT_THREADS, 0, cs>>>(
                d_stream, bp, slot_lo, slot_hi, chunk_slots, g.n_chunks, blk, ordinal_base, counts, bbase, queues[buf],
                tails[buf], (unsigned long long *)d_end_slot, overflow);
        else
            boot_bin_kernel<MCF_GEN_PHILOX><<<(unsigned)(end - blk), BOOT_THREADS, 0, cs>>>(
                d_stream, bp, slot_lo, slot_hi, chunk_slots, g.n_chunks, blk, ordinal_base, counts, bbase, queues[buf],
                tails[buf], (unsigned long long *)d_end_slot, overflow);
        if (overlap) {
            cudaEventRecord(ev_bin[buf], cs);
            cudaStreamWaitEvent(side, ev_bin[buf], 0);
        }
        for (uint32_t r = 0; r < (uint32_t)R; r++) {
            dim3 grid(64, (unsigned)window);
            boot_gather_kernel<<<grid, 256, 0, side>>>(d_x, bp, r, queues[buf], tails[buf], d_sums);
        }
        if (overlap) cudaEventRecord(ev_gat[buf], side);
        if (cudaGetLastError() != cudaSuccess) {
            rc = MCF_ERR_CUDA;
            break;
        }
        blk = end;
        k++;
    }
    if (overlap) {  // join the side stream into the caller's, then release it
        for (int b = 0; b < 2; b++)
            if (k > b) cudaStreamWaitEvent(cs, ev_gat[b], 0);
    }
    int h_over = 0;
    cudaMemcpyAsync(&h_over, overflow, 4, cudaMemcpyDeviceToHost, cs);
    const cudaError_t e1 = cudaStreamSynchronize(cs);
    if (overlap) cudaStreamSynchronize(side);
    if (e1 != cudaSuccess || rc != MCF_OK) return MCF_ERR_CUDA;
    return h_over ? MCF_ERR_WINDOW_TOO_SMALL : MCF_OK;
}

// ---- single-pass form of the binned bootstrap (see boot_bin_safe_kernel) ----
static FuseGeom fuse_geom(int64_t n, int64_t n_resamples, uint64_t slot_lo, uint64_t slot_hi, uint64_t slot_total,
                          int32_t chunk_slots) {
    FuseGeom fg;
    const SlotRule rule = slot_rule((uint32_t)n);
    fg.q = (1ULL << 32) - (uint64_t)rule.threshold;
    const double p = (double)rule.threshold / 4294967296.0;
    // 8 standard deviations of the accepted count at the LAST slot of the whole draw (all ranks), + slack
    fg.margin = (uint64_t)ceil(8.0 * sqrt((double)slot_total * p * (1.0 - p))) + 64ULL;
    fg.n = (uint64_t)n;
    fg.total_needed = (uint64_t)n_resamples * (uint64_t)n;
    fg.S = (uint64_t)BOOT_THREADS * (uint64_t)chunk_slots;
    fg.slot_lo = slot_lo;
    fg.slot_hi = slot_hi;
    return fg;
}

// Finer CTAs of the exact pass: chunk_slots / fine stays a multiple of 8, and the listed fine blocks must fit the list and
// the compact per-chunk counts, both sized for the coarse geometry (n_blocks entries) -- small populations, where a third of
// the CTAs straddle a resample boundary, get a smaller factor instead of giving up on the single-pass form.
static int fuse_fine(const FuseGeom &fg, int64_t n_blocks, int32_t chunk_slots) {
    int fine = BOOT_FINE;
    while (fine > 1 && (chunk_slots % (8 * fine)) != 0) fine >>= 1;
    int64_t unsafe = 0;
    uint64_t b = 0, bmin = 0, bmax = 0;
    for (int64_t blk = 0; blk < n_blocks; blk++) unsafe += fuse_block_class(fg, blk, &b, &bmin, &bmax) == 0;
    while (fine > 1 && unsafe * fine > n_blocks) fine >>= 1;
    return fine;
}

struct FuseQueues {
    uint32_t *queues[2];
    unsigned int *tails[2];
    int *overflow;
    int64_t R, cap, window;
    uint32_t sh;
};
static FuseQueues fuse_queues(int64_t n, int32_t resamples_per_batch, int32_t region_shift, void *d_bin_workspace) {
    FuseQueues q;
    q.sh = bin_region_shift(n, region_shift);
    q.R = (n + (1LL << q.sh) - 1) >> q.sh;
    const int64_t region_len = 1LL << q.sh;
    q.cap = region_len + 8 * (int64_t)sqrt((double)region_len) + 4096;
    q.window = resamples_per_batch + 2;
    char *bws = (char *)d_bin_workspace;
    const int64_t set_bytes = bin_set_bytes(q.window, q.R, q.cap);
    for (int b = 0; b < 2; b++) {
        q.queues[b] = (uint32_t *)(bws + b * set_bytes);
        q.tails[b] = (unsigned int *)(bws + b * set_bytes + align256(q.window * q.R * q.cap * 4));
    }
    q.overflow = (int *)(bws + 2 * set_bytes);
    return q;
}

// Phase 1: every safe CTA bins (and counts) its slots; then the CTAs that need exact ordinals are counted and all block
// sums are prefix-summed.  *d_total = accepted slots of this rank's range before the end of the draw.
int mcf_bootstrap_fused_phase1(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n, int64_t n_resamples,
                               uint64_t slot_lo, uint64_t slot_hi, uint64_t slot_total, int32_t chunk_slots, void *d_workspace,
                               int64_t workspace_bytes, int32_t resamples_per_batch, int32_t region_shift,
                               void *d_bin_workspace, int64_t bin_workspace_bytes, double *d_sums, uint64_t *d_total,
                               void *cuda_stream) {
    if (!d_stream || !d_x || n < 2 || n > 0xFFFFFFFFLL || n_resamples <= 0 || slot_hi <= slot_lo || slot_total < slot_hi ||
        chunk_slots < 8 || (chunk_slots % 8) || !d_workspace || resamples_per_batch < 1 || !d_bin_workspace || !d_sums ||
        (int64_t)BOOT_THREADS * chunk_slots > n)
        return MCF_ERR_BAD_ARG;
    const BootGeom g = boot_geom((int64_t)(slot_hi - slot_lo), chunk_slots);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    if (bin_workspace_bytes < mcf_bootstrap_binned_workspace_bytes(n, resamples_per_batch, region_shift))
        return MCF_ERR_WORKSPACE_TOO_SMALL;
    const FuseGeom fg = fuse_geom(n, n_resamples, slot_lo, slot_hi, slot_total, chunk_slots);
    const FuseQueues Q = fuse_queues(n, resamples_per_batch, region_shift, d_bin_workspace);
    char *ws = (char *)d_workspace;
    uint32_t *counts = (uint32_t *)(ws + g.off_counts);
    unsigned long long *bsum = (unsigned long long *)(ws + g.off_bsum), *bbase = (unsigned long long *)(ws + g.off_bbase);
    unsigned long long *total = (unsigned long long *)(ws + g.off_total);
    int64_t *d_list = (int64_t *)(ws + g.off_list);
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    // The exact pass runs on FINER CTAs (chunk_slots / fine slots per thread): the few CTAs that straddle a resample
    // boundary take the slot-by-slot path of the general kernel, and alone in a launch a 2e6-slot straddler is a 7 ms
    // critical path; cut 16 ways it is 0.45 ms and the launch has 16 times the CTAs.  Block sums and their prefix are
    // kept per fine block (a safe CTA owns the first of its `fine` entries, the others stay zero).
    const int fine = fuse_fine(fg, g.n_blocks, chunk_slots);
    const int64_t n_fine = g.n_blocks * fine;
    MCF_CUDA_OK(cudaMemsetAsync(Q.overflow, 0, 4, cs));
    MCF_CUDA_OK(cudaMemsetAsync(bsum, 0, (size_t)n_fine * 8, cs));
    SideStream ss;
    if (!ss.open(cs)) return MCF_ERR_CUDA;
    // The gathers of a batch run on the CALLER's stream, after its bin kernel: overlapped with the next batch's bin
    // kernel (as the two-pass form does) they take 26 ms per 32 resamples instead of 7.4 + 12.5 -- the safe kernel
    // streams 13 GB of queue entries through L2 fast enough to evict the 32 MB region the gathers depend on.  Measured at
    // n = 1e8, B = 512: 0.43 s overlapped, 0.32 s back to back.
    // (Also measured: the same overlap with the region's lines loaded evict-last into the persisting set-aside of L2, so
    // that the queue traffic cannot evict them: 0.43 s again -- the two kernels contend for more than the region.)
    const cudaStream_t side = cs;
    // classify the CTAs once on the host (integer arithmetic shared with the kernels): safe ones by batch of resamples,
    // the others into the exact list
    std::vector<int64_t> exact;
    int64_t blk = 0;
    int k = 0;
    int rc = MCF_OK;
    while (blk < g.n_blocks) {
        uint64_t b = 0, bmin = 0, bmax = 0, b_first = 0;
        bool have_first = false;
        int64_t end = blk;
        for (; end < g.n_blocks; end++) {
            const int cls = fuse_block_class(fg, end, &b, &bmin, &bmax);
            if (cls == 0)
                for (int f = 0; f < fine; f++)  // its fine blocks that still hold slots
                    if (slot_lo + (uint64_t)(end * fine + f) * (fg.S / (uint64_t)fine) < slot_hi) exact.push_back(end * fine + f);
            if (cls != 1) continue;
            if (!have_first) {
                have_first = true;
                b_first = b;
            } else if (b >= b_first + (uint64_t)resamples_per_batch) {
                break;  // this CTA opens the next batch
            }
        }
        if (have_first) {
            BinParams bp;
            bp.n = (uint32_t)n;
            bp.region_shift = Q.sh;
            bp.n_regions = (uint32_t)Q.R;
            bp.n_res_window = (uint32_t)resamples_per_batch;  // safe CTAs never leave their batch
            bp.b_first = b_first;
            bp.cap = (uint64_t)Q.cap;
            bp.n_resamples = (uint64_t)n_resamples;
            const int buf = k & 1;
            if (k >= 2) cudaStreamWaitEvent(cs, ss.ev_gat[buf], 0);  // this queue set has been drained
            cudaMemsetAsync(Q.tails[buf], 0, (size_t)(Q.window * Q.R * 4), cs);
            if (gen_kind == MCF_GEN_PCG64)
                boot_bin_safe_kernel<MCF_GEN_PCG64><<<(unsigned)(end - blk), BOOT_THREADS, 0, cs>>>(
                    d_stream, bp, fg, chunk_slots, blk, fine, bsum, Q.queues[buf], Q.tails[buf], Q.overflow);
            else
                boot_bin_safe_kernel<MCF_GEN_PHILOX><<<(unsigned)(end - blk), BOOT_THREADS, 0, cs>>>(
                    d_stream, bp, fg, chunk_slots, blk, fine, bsum, Q.queues[buf], Q.tails[buf], Q.overflow);
            cudaEventRecord(ss.ev_bin[buf], cs);
            cudaStreamWaitEvent(side, ss.ev_bin[buf], 0);
            bp.n_res_window = (uint32_t)Q.window;
            for (uint32_t r = 0; r < (uint32_t)Q.R; r++) {
                dim3 grid(64, (unsigned)resamples_per_batch);
                boot_gather_kernel<<<grid, 256, 0, side>>>(d_x, bp, r, Q.queues[buf], Q.tails[buf], d_sums);
            }
            cudaEventRecord(ss.ev_gat[buf], side);
            k++;
        }
        if (cudaGetLastError() != cudaSuccess) {
            rc = MCF_ERR_CUDA;
            break;
        }
        blk = end;
    }
    for (int b2 = 0; b2 < 2; b2++)
        if (k > b2) cudaStreamWaitEvent(cs, ss.ev_gat[b2], 0);
    if (rc != MCF_OK) return rc;
    // the (fine) CTAs that need exact ordinals: count them, then prefix-sum ALL block sums.  The compact per-chunk counts
    // and the list fit the regions sized for the coarse geometry (fuse_fine chose the factor accordingly).
    if ((int64_t)exact.size() > g.n_blocks) return MCF_ERR_WORKSPACE_TOO_SMALL;
    const int cs_fine = chunk_slots / fine;
    const int64_t n_chunks_fine = ((int64_t)(slot_hi - slot_lo) + cs_fine - 1) / cs_fine;
    if (!exact.empty()) {
        MCF_CUDA_OK(cudaMemcpyAsync(d_list, exact.data(), exact.size() * 8, cudaMemcpyHostToDevice, cs));
        if (gen_kind == MCF_GEN_PCG64)
            boot_count_kernel<MCF_GEN_PCG64><<<(unsigned)exact.size(), BOOT_THREADS, 0, cs>>>(
                d_stream, (uint32_t)n, slot_lo, slot_hi, cs_fine, n_chunks_fine, counts, bsum, d_list);
        else
            boot_count_kernel<MCF_GEN_PHILOX><<<(unsigned)exact.size(), BOOT_THREADS, 0, cs>>>(
                d_stream, (uint32_t)n, slot_lo, slot_hi, cs_fine, n_chunks_fine, counts, bsum, d_list);
    }
    scan_u64_kernel<<<1, 1024, 0, cs>>>(bsum, n_fine, bbase, total);
    MCF_CUDA_OK(cudaGetLastError());
    if (d_total) MCF_CUDA_OK(cudaMemcpyAsync(d_total, total, 8, cudaMemcpyDeviceToDevice, cs));
    MCF_CUDA_OK(cudaStreamSynchronize(cs));  // `exact` is a host vector; the side stream is released on return
    cudaStreamSynchronize(side);
    return MCF_OK;
}

// Phase 2: the listed CTAs with exact ordinals (general kernel), batch by batch, then the verification of the safe ones.
// *d_flag != 0 afterwards: a safe CTA was not where the estimate put it, or a queue overflowed -- recompute with the
// two-pass form.  ordinal_base = accepted slots of all lower ranks.
int mcf_bootstrap_fused_phase2(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n, int64_t n_resamples,
                               uint64_t slot_lo, uint64_t slot_hi, uint64_t slot_total, int32_t chunk_slots,
                               uint64_t ordinal_base, void *d_workspace, int64_t workspace_bytes, int32_t resamples_per_batch,
                               int32_t region_shift, void *d_bin_workspace, int64_t bin_workspace_bytes, double *d_sums,
                               uint64_t *d_end_slot, int32_t *d_flag, void *cuda_stream) {
    if (!d_stream || !d_x || n < 2 || n_resamples <= 0 || slot_hi <= slot_lo || chunk_slots < 8 || !d_workspace ||
        !d_bin_workspace || !d_sums || !d_end_slot || !d_flag)
        return MCF_ERR_BAD_ARG;
    const BootGeom g = boot_geom((int64_t)(slot_hi - slot_lo), chunk_slots);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    if (bin_workspace_bytes < mcf_bootstrap_binned_workspace_bytes(n, resamples_per_batch, region_shift))
        return MCF_ERR_WORKSPACE_TOO_SMALL;
    const FuseGeom fg = fuse_geom(n, n_resamples, slot_lo, slot_hi, slot_total, chunk_slots);
    const FuseQueues Q = fuse_queues(n, resamples_per_batch, region_shift, d_bin_workspace);
    char *ws = (char *)d_workspace;
    const uint32_t *counts = (const uint32_t *)(ws + g.off_counts);
    const unsigned long long *bsum = (const unsigned long long *)(ws + g.off_bsum);
    const unsigned long long *bbase = (const unsigned long long *)(ws + g.off_bbase);
    const int64_t *d_list = (const int64_t *)(ws + g.off_list);
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    MCF_CUDA_OK(cudaMemsetAsync(d_flag, 0, 4, cs));
    const int fine = fuse_fine(fg, g.n_blocks, chunk_slots);
    const int cs_fine = chunk_slots / fine;
    const int64_t n_chunks_fine = ((int64_t)(slot_hi - slot_lo) + cs_fine - 1) / cs_fine;
    // the same classification as phase 1 (the list on the device is in this order)
    std::vector<uint64_t> lo_b, hi_b;
    for (int64_t blk = 0; blk < g.n_blocks; blk++) {
        uint64_t b = 0, bmin = 0, bmax = 0;
        if (fuse_block_class(fg, blk, &b, &bmin, &bmax) == 0) {
            for (int f = 0; f < fine; f++) {
                if (slot_lo + (uint64_t)(blk * fine + f) * (fg.S / (uint64_t)fine) >= slot_hi) continue;
                // exact ordinals lie within the margin of the estimate: one resample of slack on either side
                lo_b.push_back(bmin > 0 ? bmin - 1 : 0);
                hi_b.push_back(bmax + 1);
            }
        }
    }
    SideStream ss;
    if (!ss.open(cs)) return MCF_ERR_CUDA;
    const cudaStream_t side = ss.stream;
    size_t i = 0;
    int k = 0;
    int rc = MCF_OK;
    while (i < lo_b.size()) {
        const uint64_t b_first = lo_b[i];
        if (hi_b[i] >= b_first + (uint64_t)Q.window) return MCF_ERR_BAD_ARG;  // a CTA wider than the queue window
        size_t end = i + 1;
        while (end < lo_b.size() && hi_b[end] < b_first + (uint64_t)Q.window) end++;
        BinParams bp;
        bp.n = (uint32_t)n;
        bp.region_shift = Q.sh;
        bp.n_regions = (uint32_t)Q.R;
        bp.n_res_window = (uint32_t)Q.window;
        bp.b_first = b_first;
        bp.cap = (uint64_t)Q.cap;
        bp.n_resamples = (uint64_t)n_resamples;
        const int buf = k & 1;
        if (k >= 2) cudaStreamWaitEvent(cs, ss.ev_gat[buf], 0);
        cudaMemsetAsync(Q.tails[buf], 0, (size_t)(Q.window * Q.R * 4), cs);
        if (gen_kind == MCF_GEN_PCG64)
            boot_bin_kernel<MCF_GEN_PCG64><<<(unsigned)(end - i), BOOT_THREADS, 0, cs>>>(
                d_stream, bp, slot_lo, slot_hi, cs_fine, n_chunks_fine, (int64_t)i, ordinal_base, counts, bbase, Q.queues[buf],
                Q.tails[buf], (unsigned long long *)d_end_slot, Q.overflow, d_list);
        else
            boot_bin_kernel<MCF_GEN_PHILOX><<<(unsigned)(end - i), BOOT_THREADS, 0, cs>>>(
                d_stream, bp, slot_lo, slot_hi, cs_fine, n_chunks_fine, (int64_t)i, ordinal_base, counts, bbase, Q.queues[buf],
                Q.tails[buf], (unsigned long long *)d_end_slot, Q.overflow, d_list);
        cudaEventRecord(ss.ev_bin[buf], cs);
        cudaStreamWaitEvent(side, ss.ev_bin[buf], 0);
        for (uint32_t r = 0; r < (uint32_t)Q.R; r++) {
            dim3 grid(16, (unsigned)Q.window);  // these queues hold a few per cent of a resample
            boot_gather_kernel<<<grid, 256, 0, side>>>(d_x, bp, r, Q.queues[buf], Q.tails[buf], d_sums);
        }
        cudaEventRecord(ss.ev_gat[buf], side);
        if (cudaGetLastError() != cudaSuccess) {
            rc = MCF_ERR_CUDA;
            break;
        }
        i = end;
        k++;
    }
    for (int b2 = 0; b2 < 2; b2++)
        if (k > b2) cudaStreamWaitEvent(cs, ss.ev_gat[b2], 0);
    boot_fuse_verify_kernel<<<MCF_SM_COUNT * 2, 256, 0, cs>>>(fg, g.n_blocks, fine, ordinal_base, bbase, bsum, (int *)d_flag);
    int h_over = 0;
    cudaMemcpyAsync(&h_over, Q.overflow, 4, cudaMemcpyDeviceToHost, cs);
    const cudaError_t e1 = cudaStreamSynchronize(cs);
    cudaStreamSynchronize(side);
    if (e1 != cudaSuccess || rc != MCF_OK) return MCF_ERR_CUDA;
    return h_over ? MCF_ERR_WINDOW_TOO_SMALL : MCF_OK;
}

int mcf_bootstrap_finalize(double *d_sums, int64_t n_resamples, int64_t n, void *cuda_stream) {
    if (!d_sums || n_resamples <= 0 || n <= 0) return MCF_ERR_BAD_ARG;
    boot_finalize_kernel<<<(unsigned)((n_resamples + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(d_sums, n_resamples,
                                                                                                    (double)n);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int mcf_draw_integers(uint32_t gen_kind, const mcf_stream *d_streams, int32_t n_streams, int64_t n_total,
                      int64_t sims_per_stream_hint, int64_t low, int64_t high, int32_t size, int32_t op, int64_t match,
                      double *d_out, int32_t *d_rejected, void *cuda_stream) {
    if (!d_streams || n_streams <= 0 || n_total <= 0 || size <= 0 || high - low < 2 || high - low > 0xFFFFFFFFLL ||
        (op != 0 && op != 1) || !d_out || !d_rejected)
        return MCF_ERR_BAD_ARG;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    MCF_CUDA_OK(cudaMemsetAsync(d_rejected, 0, 4, cs));
    const unsigned grid = (unsigned)((n_total + 255) / 256);
    const uint32_t span = (uint32_t)(high - low);
    if (gen_kind == MCF_GEN_PCG64)
        draw_integers_kernel<MCF_GEN_PCG64><<<grid, 256, 0, cs>>>(d_streams, n_streams, n_total, sims_per_stream_hint, span,
                                                                  low, size, op, match, d_out, d_rejected);
    else
        draw_integers_kernel<MCF_GEN_PHILOX><<<grid, 256, 0, cs>>>(d_streams, n_streams, n_total, sims_per_stream_hint, span,
                                                                   low, size, op, match, d_out, d_rejected);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int64_t mcf_lsm_workspace_bytes(int64_t n_paths) {
    if (n_paths <= 0) return MCF_ERR_BAD_ARG;
    return lsm_geom(n_paths).total;
}

int mcf_lsm_begin(const double *d_paths_tm, int64_t n_paths, int64_t n_steps, double K, double r, double dt,
                  int32_t is_put, void *d_workspace, int64_t workspace_bytes, void *cuda_stream) {
    if (!d_paths_tm || n_paths <= 0 || n_steps < 1 || n_steps > 0x7fffffff || !d_workspace) return MCF_ERR_BAD_ARG;
    const LsmGeom g = lsm_geom(n_paths);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    char *ws = (char *)d_workspace;
    lsm_init_kernel<<<g.n_blocks, LSM_THREADS, 0, (cudaStream_t)cuda_stream>>>(
        d_paths_tm + (size_t)n_steps * (size_t)n_paths, n_paths, (int32_t)n_steps, K, is_put,
        exp(-r * dt * (double)n_steps), (int32_t *)(ws + g.off_ex), (double *)(ws + g.off_cf));
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

// sums of step t -> workspace[0..8) doubles (all-reduce them across ranks when paths are sharded)
int mcf_lsm_reduce(const double *d_paths_tm, int64_t n_paths, int64_t t, double K, double r, double dt, int32_t is_put,
                   void *d_workspace, void *cuda_stream) {
    if (!d_paths_tm || n_paths <= 0 || t < 0 || !d_workspace) return MCF_ERR_BAD_ARG;
    const LsmGeom g = lsm_geom(n_paths);
    char *ws = (char *)d_workspace;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    lsm_reduce_kernel<<<g.n_blocks, LSM_THREADS, 0, cs>>>(d_paths_tm + (size_t)t * (size_t)n_paths, n_paths, K,
                                                          exp(r * dt * (double)t), is_put, (const double *)(ws + g.off_cf),
                                                          (double *)(ws + g.off_partials));
    MCF_CUDA_OK(cudaGetLastError());
    lsm_sum_partials_kernel<<<1, 256, 0, cs>>>((const double *)(ws + g.off_partials), g.n_blocks,
                                               (LsmState *)(ws + g.off_state));
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int mcf_lsm_apply(const double *d_paths_tm, int64_t n_paths, int64_t t, double K, double r, double dt, int32_t is_put,
                  void *d_workspace, void *cuda_stream) {
    if (!d_paths_tm || n_paths <= 0 || t < 0 || !d_workspace) return MCF_ERR_BAD_ARG;
    const LsmGeom g = lsm_geom(n_paths);
    char *ws = (char *)d_workspace;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    lsm_solve_kernel<<<1, 32, 0, cs>>>((LsmState *)(ws + g.off_state));
    MCF_CUDA_OK(cudaGetLastError());
    lsm_apply_kernel<<<g.n_blocks, LSM_THREADS, 0, cs>>>(d_paths_tm + (size_t)t * (size_t)n_paths, n_paths, (int32_t)t, K,
                                                         is_put, exp(-r * dt * (double)t),
                                                         (const LsmState *)(ws + g.off_state),
                                                         (int32_t *)(ws + g.off_ex), (double *)(ws + g.off_cf));
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

// sum of discounted cash flows over this rank's paths -> d_price_sum[0]; price = sum / n_paths -> d_price[0]
int mcf_lsm_finish(int64_t n_paths, double r, double dt, void *d_workspace, double *d_price_sum, double *d_price,
                   int32_t *d_exercise_times, void *cuda_stream) {
    if (n_paths <= 0 || !d_workspace) return MCF_ERR_BAD_ARG;
    const LsmGeom g = lsm_geom(n_paths);
    char *ws = (char *)d_workspace;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    lsm_price_kernel<<<g.n_blocks, LSM_THREADS, 0, cs>>>(n_paths, (const double *)(ws + g.off_cf),
                                                         (double *)(ws + g.off_partials));
    (void)r;
    (void)dt;
    MCF_CUDA_OK(cudaGetLastError());
    lsm_sum_partials_kernel<<<1, 256, 0, cs>>>((const double *)(ws + g.off_partials), g.n_blocks,
                                               (LsmState *)(ws + g.off_state));
    MCF_CUDA_OK(cudaGetLastError());
    lsm_price_final_kernel<<<1, 32, 0, cs>>>((const LsmState *)(ws + g.off_state), (double)n_paths, d_price_sum, d_price);
    MCF_CUDA_OK(cudaGetLastError());
    if (d_exercise_times)
        MCF_CUDA_OK(cudaMemcpyAsync(d_exercise_times, ws + g.off_ex, (size_t)n_paths * 4, cudaMemcpyDeviceToDevice, cs));
    return MCF_OK;
}

#ifdef MCF_LSM_NOREV  // (development: every step sweeps front to back)
#define MCF_LSM_REVERSE(k) 0
#else
#define MCF_LSM_REVERSE(k) (int)((k) & 1)
#endif
typedef void (*LsmStepKernel)(const double *, const double *, int64_t, int32_t, int32_t, double, int, double, double,
                              const LsmState *, LsmState *, int32_t *, double *, double *, unsigned int *, double *, int,
                              uint32_t, const LsmPeers);
static LsmStepKernel lsm_step_kernel_for(int mode, bool has_ex, int vec, bool keep) {
#define MCF_LSM_PICK2(M, E, V) (keep ? lsm_fused_step_kernel<M, E, V, true> : lsm_fused_step_kernel<M, E, V, false>)
#define MCF_LSM_PICK(M) \
    case M:             \
        return vec == 2 ? (has_ex ? MCF_LSM_PICK2(M, true, 2) : MCF_LSM_PICK2(M, false, 2)) \
                        : (has_ex ? MCF_LSM_PICK2(M, true, 1) : MCF_LSM_PICK2(M, false, 1));
    switch (mode) {
        MCF_LSM_PICK(0)
        MCF_LSM_PICK(1)
        MCF_LSM_PICK(2)
    default:
        MCF_LSM_PICK(3)
    }
#undef MCF_LSM_PICK
#undef MCF_LSM_PICK2
}

// L2 set-aside for the cash flows, held for the duration of one sweep: all of the array or nothing.  A partial window is
// worse than none -- a 79 MB window over the 160 MB array of 2e7 paths: 21.9 ms against 20.9 ms, a 48 MB window at 1e7
// paths: 11.1 ms against 10.4 -- because it takes the capacity that the alternating sweep direction would have used for
// both arrays.  And the set-aside must not outlive the sweep: the limit is context-wide and an enabled set-aside costs
// every other kernel its share of L2 whether persisting lines exist or not (a 2e7-path sweep, which cannot use it, took
// 40 ms instead of 21 with the 79 MB set-aside of an earlier sweep still in place).  So: raise the limit to what the
// array needs (device maximum: 79 of 126 MB on a B200), sweep, wait for the stream, reset the persisting lines and put
// the limit back.  A sweep that uses the set-aside therefore returns when it has finished.  MCF_LSM_L2_PERSIST=0 leaves
// the device alone (and the call asynchronous).
struct LsmL2Reservation {
    size_t before = 0;
    uint32_t keep_bytes = 0;  // 0 = no residency
    void acquire(size_t cf_bytes) {
        const char *e = getenv("MCF_LSM_L2_PERSIST");
        if ((e && e[0] == '0') || cf_bytes >= (1ULL << 32)) return;
        int dev = 0, max_persist = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess ||
            max_persist <= 0 || cf_bytes > (size_t)max_persist)
            return;
        if (cudaDeviceGetLimit(&before, cudaLimitPersistingL2CacheSize) != cudaSuccess) return;
        size_t ask = cf_bytes + cf_bytes / 8, now = 0;  // (the driver rounds the request to its own granularity)
        if (ask > (size_t)max_persist) ask = (size_t)max_persist;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ask) != cudaSuccess ||
            cudaDeviceGetLimit(&now, cudaLimitPersistingL2CacheSize) != cudaSuccess || now < cf_bytes) {
            (void)cudaGetLastError();
            (void)cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, before);
            return;
        }
        keep_bytes = (uint32_t)cf_bytes;
    }
    int release(cudaStream_t cs) {
        if (!keep_bytes) return MCF_OK;
        keep_bytes = 0;
        const cudaError_t e1 = cudaStreamSynchronize(cs);
        (void)cudaCtxResetPersistingL2Cache();
        (void)cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, before);
        return e1 == cudaSuccess ? MCF_OK : MCF_ERR_CUDA;
    }
};

static int lsm_sweep(const double *d_paths_tm, int64_t n_paths, int64_t n_steps, double K, double r, double dt, int32_t is_put,
                     void *d_workspace, int64_t workspace_bytes, double *d_price, int32_t *d_exercise_times,
                     const LsmPeers &peers, void *cuda_stream) {
    if (!d_paths_tm || n_paths <= 0 || n_steps < 1 || n_steps > 0x7fffffff || !d_workspace) return MCF_ERR_BAD_ARG;
    const LsmGeom g = lsm_geom(n_paths);
    if (workspace_bytes < g.total) return MCF_ERR_WORKSPACE_TOO_SMALL;
    char *ws = (char *)d_workspace;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    LsmState *st = (LsmState *)(ws + g.off_state);
    unsigned int *counter = (unsigned int *)(ws + g.off_counter);
    double *cf = (double *)(ws + g.off_cf), *partials = (double *)(ws + g.off_partials);
    MCF_CUDA_OK(cudaMemsetAsync(ws + g.off_state, 0, (size_t)(g.off_partials - g.off_state), cs));
    // one co-resident wave of CTAs (the sweep order of a launch is then the band order), never more than the partial-sum
    // slots the workspace was sized for
    int grid = MCF_SM_COUNT * MCF_LSM_CTAS;
    if (grid > g.n_blocks) grid = g.n_blocks;
    LsmL2Reservation l2;
    l2.acquire((size_t)n_paths * 8);
    const uint32_t keep_bytes = l2.keep_bytes;
    // 16-byte loads (two paths per load) when every column S_t and the cash flows are 16-byte aligned
    const int vec = (n_paths % 2 == 0 && ((uintptr_t)d_paths_tm & 15) == 0 && ((uintptr_t)cf & 15) == 0) ? 2 : 1;
    // t = n_steps: initialise cash flows and fit step n_steps-1; t = n_steps-1 .. 1: decide step t, fit step t-1
    for (int64_t t = n_steps; t >= 1; t--) {
        const double *s_t = d_paths_tm + (size_t)t * (size_t)n_paths;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(LSM_THREADS);
        cfg.stream = cs;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
#ifndef MCF_LSM_NO_PDL
        cfg.attrs = attr;
        cfg.numAttrs = t == n_steps ? 0 : 1;  // the first step follows a memset, not a kernel of this chain
#endif
        const int mode = (t == n_steps ? 1 : 0) | (t == 1 ? 2 : 0);
        const LsmStepKernel kernel = lsm_step_kernel_for(mode, d_exercise_times != nullptr, vec, keep_bytes != 0);
        const cudaError_t le = cudaLaunchKernelEx(
            &cfg, kernel, s_t, s_t - n_paths, n_paths, (int32_t)t, (int32_t)n_steps, K, (int)is_put, exp(-r * dt * (double)t),
            exp(r * dt * (double)(t - 1)), (const LsmState *)(st + (t & 1)), st + ((t - 1) & 1), d_exercise_times, cf, partials,
            counter, d_price, MCF_LSM_REVERSE(n_steps - t), keep_bytes, peers);
        if (le != cudaSuccess) {
            (void)l2.release(cs);
            return MCF_ERR_CUDA;
        }
    }
    const cudaError_t launched = cudaGetLastError();
    const int released = l2.release(cs);
    if (launched != cudaSuccess) return MCF_ERR_CUDA;
    return released;
}

int mcf_lsm(const double *d_paths_tm, int64_t n_paths, int64_t n_steps, double K, double r, double dt, int32_t is_put,
            void *d_workspace, int64_t workspace_bytes, double *d_price, int32_t *d_exercise_times, void *cuda_stream) {
    LsmPeers solo;
    memset(&solo, 0, sizeof(solo));
    solo.world = 1;
    return lsm_sweep(d_paths_tm, n_paths, n_steps, K, r, dt, is_put, d_workspace, workspace_bytes, d_price, d_exercise_times,
                     solo, cuda_stream);
}

// ---- peer-mapped exchange buffers (one per rank and process group; CUDA IPC, the GPUs of one box) ----
int64_t mcf_lsm_exchange_bytes(void) { return (int64_t)(2 * MCF_LSM_MAX_PEERS * MCF_LSM_SLOT_DOUBLES * sizeof(double)); }

int mcf_peer_buffer_create(int64_t bytes, void **d_ptr, void *handle64) {
    if (bytes <= 0 || !d_ptr || !handle64) return MCF_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 opaque bytes");
    MCF_CUDA_OK(cudaMalloc(d_ptr, (size_t)bytes));
    MCF_CUDA_OK(cudaMemset(*d_ptr, 0, (size_t)bytes));
    MCF_CUDA_OK(cudaDeviceSynchronize());
    MCF_CUDA_OK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, *d_ptr));
    return MCF_OK;
}
int mcf_peer_buffer_open(const void *handle64, void **d_ptr) {
    if (!handle64 || !d_ptr) return MCF_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    MCF_CUDA_OK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MCF_OK;
}
int mcf_peer_buffer_close(void *d_ptr) {
    if (d_ptr) MCF_CUDA_OK(cudaIpcCloseMemHandle(d_ptr));
    return MCF_OK;
}
int mcf_peer_buffer_destroy(void *d_ptr) {
    if (d_ptr) MCF_CUDA_OK(cudaFree(d_ptr));
    return MCF_OK;
}

int mcf_lsm_sharded(const double *d_paths_tm, int64_t n_paths_local, int64_t n_paths_global, int64_t n_steps, double K, double r,
                    double dt, int32_t is_put, void *d_workspace, int64_t workspace_bytes, void *const *peer_buffers,
                    int32_t rank, int32_t world, uint64_t sequence_base, double *d_price, int32_t *d_exercise_times,
                    void *cuda_stream) {
    if (!peer_buffers || world < 1 || world > MCF_LSM_MAX_PEERS || rank < 0 || rank >= world || n_paths_global < n_paths_local)
        return MCF_ERR_BAD_ARG;
    LsmPeers px;
    memset(&px, 0, sizeof(px));
    for (int p = 0; p < world; p++) {
        if (!peer_buffers[p]) return MCF_ERR_BAD_ARG;
        px.slots[p] = (double *)peer_buffers[p];
    }
    px.rank = rank;
    px.world = world;
    px.seq_base = sequence_base;
    px.n_paths_global = (double)n_paths_global;
    return lsm_sweep(d_paths_tm, n_paths_local, n_steps, K, r, dt, is_put, d_workspace, workspace_bytes, d_price,
                     d_exercise_times, px, cuda_stream);
}

int mcf_transpose(const double *d_in, int64_t n_rows, int64_t n_cols, double *d_out, void *cuda_stream) {
    if (!d_in || !d_out || n_rows <= 0 || n_cols <= 0) return MCF_ERR_BAD_ARG;
    const int64_t gx = (n_cols + 31) / 32, gy = (n_rows + 31) / 32;
    if (gx * gy > 0x7fffffffLL) return MCF_ERR_BAD_ARG;
    transpose_kernel<<<(unsigned)(gx * gy), 256, 0, (cudaStream_t)cuda_stream>>>(d_in, n_rows, n_cols, gx, d_out);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int mcf_count_less(const double *d_x, int64_t n, double value, int64_t *d_count, void *cuda_stream) {
    if (!d_x || n <= 0 || !d_count) return MCF_ERR_BAD_ARG;
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    MCF_CUDA_OK(cudaMemsetAsync(d_count, 0, 8, cs));
    count_less_kernel<<<grid_cap(n, 256 * 8, MCF_SM_COUNT * 4), 256, 0, cs>>>(d_x, n, value, (unsigned long long *)d_count);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

int64_t mcf_compact_workspace_bytes(int64_t n) {
    if (n <= 0) return MCF_ERR_BAD_ARG;
    const int64_t tiles = (n + COMPACT_TILE - 1) / COMPACT_TILE;
    return align256(tiles * 8) * 2 + 256;
}

// d_out[0 .. *d_count) = the finite values of d_x in their original order (d_out needs room for n doubles)
int mcf_compact_finite(const double *d_x, int64_t n, double *d_out, int64_t *d_count, void *d_workspace,
                       int64_t workspace_bytes, void *cuda_stream) {
    if (!d_x || n <= 0 || !d_out || !d_count || !d_workspace) return MCF_ERR_BAD_ARG;
    if (workspace_bytes < mcf_compact_workspace_bytes(n)) return MCF_ERR_WORKSPACE_TOO_SMALL;
    const int64_t tiles = (n + COMPACT_TILE - 1) / COMPACT_TILE;
    char *ws = (char *)d_workspace;
    unsigned long long *cnt = (unsigned long long *)ws, *base = (unsigned long long *)(ws + align256(tiles * 8));
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    compact_count_kernel<<<(unsigned)tiles, 256, 0, cs>>>(d_x, n, cnt);
    MCF_CUDA_OK(cudaGetLastError());
    scan_u64_kernel<<<1, 1024, 0, cs>>>(cnt, tiles, base, (unsigned long long *)d_count);
    MCF_CUDA_OK(cudaGetLastError());
    compact_scatter_kernel<<<(unsigned)tiles, 256, 0, cs>>>(d_x, n, base, d_out);
    MCF_CUDA_OK(cudaGetLastError());
    return MCF_OK;
}

}  // extern "C"
