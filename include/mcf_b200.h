/*
 * mcf_b200.h -- C ABI of the B200-native replication + statistics hot path of mcframework.
 *
 * The reference (milanfusco/mcFramework) is pure Python and has NO FFI for this path; the
 * entry points below are what a binding for it would call.  Each one cites the reference
 * code it replaces (paths under the reference's src/mcframework/).  A ctypes stub a
 * maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no exceptions, no torch types.  Return value: MCF_OK (0) or a negative error.
 *   - every `d_*` pointer is DEVICE memory owned by the caller; `h_*` is HOST memory.
 *   - no allocation inside the library: scratch comes from a caller-provided device workspace
 *     whose size the matching `*_workspace_bytes` function returns.
 *   - all launches go to the caller's `cudaStream_t` (passed as void*), nothing synchronises
 *     unless stated.  Kernels report data-dependent problems (e.g. an under-provisioned
 *     stream window) through a status word inside the workspace: read it with
 *     mcf_workspace_status() AFTER synchronising the stream.
 *   - `gen_kind` (MCF_GEN_*) names the generator of ALL records of a call (a launch never mixes
 *     generators); `sims_per_stream_hint` is the uniform block size (0 = unknown) that lets a
 *     thread find its stream record without a search.
 *   - random streams are described by `mcf_stream` records produced on the host from NumPy's
 *     own SeedSequence / bit_generator.state, so the device walks exactly the streams the
 *     reference would (core.py:363-364 PCG64; core.py:102-103,637,659 Philox per block).
 */
#ifndef MCF_B200_H
#define MCF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCF_ABI_VERSION 4

enum {
    MCF_OK = 0,
    MCF_ERR_BAD_ARG = -1,
    MCF_ERR_CUDA = -2,
    MCF_ERR_WORKSPACE_TOO_SMALL = -3,
    MCF_ERR_WINDOW_TOO_SMALL = -4, /* stream window under-provisioned: re-plan with a larger slack */
    MCF_ERR_NOT_CONVERGED = -5     /* chunk-entry fix-up needs more iterations (practically never) */
};

enum { MCF_GEN_PCG64 = 0, MCF_GEN_PHILOX = 1 };

/* One reference random stream and the replications that consume it back to back.
 * PCG64  (np.random.default_rng, core.py:364): s = {state_hi, state_lo, inc_hi, inc_lo, 0, 0}
 * Philox (np.random.Philox(child), core.py:659): s = {key0, key1, ctr0, ctr1, ctr2, ctr3}
 * buf_pos: Philox output-buffer position (4 = empty, the state right after construction).
 * has_uint32/uinteger: NumPy's buffered upper half for next_uint32 (Generator.integers). */
typedef struct mcf_stream {
    uint32_t kind;
    uint32_t buf_pos;
    uint32_t has_uint32;
    uint32_t uinteger;
    uint64_t s[6];
    int64_t first_sim; /* index of this stream's first replication in the result vector */
    int64_t n_sims;    /* replications drawn from this stream, consecutively */
} mcf_stream;

/* Per-replication bodies that consume standard normals (ziggurat, variable stream consumption). */
enum {
    MCF_BS_EURO_CALL = 0,    /* p={S0,K,T,r,sigma}   sims/black_scholes.py:237-243 */
    MCF_BS_EURO_PUT = 1,
    MCF_BS_AMER_CALL = 2,    /* max_k intrinsic_k*exp(-r*dt*k)  sims/black_scholes.py:245-255 */
    MCF_BS_AMER_PUT = 3,
    MCF_PATH_TERMINAL = 4,   /* p={S0,r,sigma,T}     sims/black_scholes.py:399-401 */
    MCF_PORTFOLIO_GBM = 5,   /* p={V0,mu,sigma}      sims/portfolio.py:79-82 */
    MCF_PORTFOLIO_LOG1P = 6, /* p={V0,mu,sigma}      sims/portfolio.py:83-84 */
    MCF_NORMAL_SUM = 7,      /* sum of the n_steps standard normals (building block / tests) */
    MCF_NORMAL_LOC_SCALE = 8 /* p={loc,scale}: n_steps*loc + scale*sum(z); n_steps = 1 is rng.normal(loc, scale) bit for bit
                                (declared draw programs of user simulations, tests/conftest.py:9-14 of the reference) */
};

typedef struct mcf_gbm_spec {
    int32_t mode;
    int32_t reserved;
    double p[6];
} mcf_gbm_spec;

/* ---- library info -------------------------------------------------------------------- */
int mcf_abi_version(void);
const char *mcf_build_info(void); /* "sm_100a, nvcc 12.9, ..." */
int mcf_last_cuda_error(void);    /* cudaError_t behind the last MCF_ERR_CUDA */

/* ---- (1) Pi estimation: fixed draw count => exact jump-ahead, no planning pass ---------
 * replaces PiEstimationSimulation.single_simulation looped by _run_sequential/_work
 * (sims/pi.py:69-82; core.py:593-594,661-662).
 * d_hits[i]  : points with x*x+y*y<=1 of replication i (bit-exact integer)
 * d_out[i]   : 4.0*hits/n_points (what single_simulation returns); may be NULL */
/* h_streams: optional HOST copy of the records (may be NULL); with it block-aligned Philox streams get one launch per
 * stream with launch-constant round keys. */
int mcf_pi_hits(uint32_t gen_kind, const mcf_stream *d_streams, const mcf_stream *h_streams, int32_t n_streams,
                int64_t n_total, int64_t n_points, int32_t antithetic, int64_t sims_per_stream_hint, int64_t *d_hits,
                double *d_out, void *cuda_stream);

/* ---- (2) normal-consuming replications -------------------------------------------------
 * Step A, mcf_normal_plan: locate, for every replication, the position in its reference
 * stream where its first standard_normal() call starts.  The ziggurat consumes a data
 * dependent number of 64-bit draws, so this is a speculative chunked scan of each stream
 * (64 draws per thread) + entry fix-up + prefix sum + locate.  The plan depends only on the
 * streams and draws_per_sim, not on model parameters: Greeks re-use one plan for all bumps
 * (sims/black_scholes.py:286-362 re-seed with 42 before every revaluation).
 * Step B, mcf_gbm_terminal / mcf_gbm_paths / mcf_gbm_slices: one thread per replication replays the draws of
 * its stream range; the plan workspace (`d_plan_workspace`, the one mcf_normal_plan filled -- keep it alive)
 * holds one bit per draw saying "an attempt starting here emits a normal", so the replay is branch-light:
 * no acceptance tests are repeated, only emitted draws are converted.  Path state stays in registers. */
int64_t mcf_normal_plan_workspace_bytes(int32_t n_streams, int64_t max_sims_per_stream, int64_t normals_per_sim,
                                        double slack);
/* h_streams: optional HOST copy of the same records (may be NULL).  With it, block-aligned Philox streams are walked
 * by one launch per stream whose round keys are launch constants (no key-schedule instructions). */
int mcf_normal_plan(uint32_t gen_kind, const mcf_stream *d_streams, const mcf_stream *h_streams, int32_t n_streams,
                    int64_t n_total, int64_t max_sims_per_stream, int64_t normals_per_sim, double slack,
                    int32_t resolve_iters, void *d_workspace, int64_t workspace_bytes,
                    uint64_t *d_sim_start /* [n_total] */,
                    uint64_t *d_stream_end /* [n_streams] draws consumed per stream; may be NULL */, void *cuda_stream);
/* 0 if the last plan in this workspace was valid, else MCF_ERR_WINDOW_TOO_SMALL / MCF_ERR_NOT_CONVERGED.
 * Synchronises the stream (reads one word back). */
int mcf_workspace_status(const void *d_workspace, void *cuda_stream);

/* d_out[k*n_total + i] = value of replication i under specs[k] (common random numbers).
 * replaces BlackScholesSimulation / BlackScholesPathSimulation / PortfolioSimulation
 * .single_simulation under the replication loop.
 * h_streams: optional HOST copy of the stream records (may be NULL).  With it, affine bodies (European payoffs,
 * portfolio GBM, path terminal) on block-aligned Philox streams run one launch per stream and regenerate at most two
 * Philox blocks per replication (the shorter side of the replication boundary; the rest comes from the stored sums). */
int mcf_gbm_terminal(uint32_t gen_kind, const mcf_stream *d_streams, const mcf_stream *h_streams, int32_t n_streams,
                     int64_t n_total, int64_t n_steps, int64_t sims_per_stream_hint, const uint64_t *d_sim_start,
                     const void *d_plan_workspace, const mcf_gbm_spec *h_specs, int32_t n_specs, double *d_out,
                     void *cuda_stream);

/* Full GBM paths, (n_total, n_steps+1) row-major when time_major==0 (what simulate_paths returns,
 * sims/black_scholes.py:415-418) or (n_steps+1, n_total) when time_major==1 (LSM layout).
 * p={S0,r,sigma,T}. */
int mcf_gbm_paths(uint32_t gen_kind, const mcf_stream *d_streams, int32_t n_streams, int64_t n_total, int64_t n_steps,
                  int64_t sims_per_stream_hint, const uint64_t *d_sim_start, const void *d_plan_workspace,
                  const double *h_p4, int32_t time_major, double *d_paths, void *cuda_stream);

/* Stored path slices (BASELINE config 3; the reference stores either nothing, sims/portfolio.py:77-84, or every
 * step, sims/black_scholes.py:415-418): values at steps 0, slice_stride, 2*slice_stride, ... <= n_steps, i.e.
 * n_slices = n_steps/slice_stride + 1 per replication.  h_spec->mode is MCF_PORTFOLIO_GBM (p={V0,mu,sigma}, value
 * V0*exp(cumulative log-return)) or MCF_PATH_TERMINAL (p={S0,r,sigma,T}).  Output (n_slices, n_total) when
 * time_major (coalesced stores) else (n_total, n_slices). */
int mcf_gbm_slices(uint32_t gen_kind, const mcf_stream *d_streams, int32_t n_streams, int64_t n_total, int64_t n_steps,
                   int64_t sims_per_stream_hint, const uint64_t *d_sim_start, const void *d_plan_workspace,
                   const mcf_gbm_spec *h_spec, int64_t slice_stride, int32_t time_major, double *d_out,
                   void *cuda_stream);

/* Stored path slices from segment sums (BASELINE cfg 3).  mcf_normal_relocate re-runs only the locate step of a finished
 * plan for `new_normals_per_sim` (the same streams cut into shorter pseudo-replications; `d_streams_new` carries the scaled
 * first_sim / n_sims; max_sims_per_stream, normals_per_sim and slack are the ORIGINAL plan's).  With it every path is cut into
 * n_steps/stride segments, mcf_gbm_terminal(MCF_NORMAL_SUM) returns the segment sums at two Philox blocks per segment, and
 * mcf_gbm_slices_from_sums accumulates them into the slice values V0*exp(...) (MCF_PORTFOLIO_GBM) or exp(log S0 + ...)
 * (MCF_PATH_TERMINAL): d_out is (n_segments+1, n_total) when time_major, else (n_total, n_segments+1). */
int mcf_normal_relocate(const mcf_stream *d_streams_new, int32_t n_streams, int64_t max_sims_per_stream,
                        int64_t normals_per_sim, double slack, int64_t new_normals_per_sim, void *d_workspace,
                        uint64_t *d_sim_start_new, void *cuda_stream);
/* out_f32 != 0: d_out is float (accumulation and exp stay in fp64, only the store narrows) -- half the HBM bytes.
 * Time-major output with n_total % 2 == 0 (fp64) / % 4 == 0 (fp32) is written with one 128-bit store per lane. */
int mcf_gbm_slices_from_sums(const double *d_segment_sums, int64_t n_total, int64_t n_segments, int64_t segment_steps,
                             int64_t n_steps, const mcf_gbm_spec *h_spec, int32_t time_major, int32_t out_f32,
                             void *d_out, void *cuda_stream);

/* ---- (2b) declared integer draw programs ------------------------------------------------
 * A user simulation whose hook is `rng.integers(low, high, size=k)` followed by a reduction (the reference's dice
 * fixture core.py:22-25, the coin-flip streak of docs/source/getting-started.md:139-183) can declare that instead of
 * being looped in Python.  Replication i of a stream reads the 32-bit slots [i*k, (i+1)*k) (NumPy's buffered
 * next_uint32: low half, then high half of each 64-bit draw) through the 32-bit Lemire rule.  op 0: sum of the values;
 * op 1: longest run of values equal to `match`.  A rejected slot (probability < 2^-32*(high-low) per draw) would shift
 * every later slot: the kernel raises *d_rejected and the caller falls back to the reference's own loop.
 * Requires 2 <= high-low < 2^32. */
int mcf_draw_integers(uint32_t gen_kind, const mcf_stream *d_streams, int32_t n_streams, int64_t n_total,
                      int64_t sims_per_stream_hint, int64_t low, int64_t high, int32_t size, int32_t op, int64_t match,
                      double *d_out, int32_t *d_rejected, void *cuda_stream);

/* ---- (3) Longstaff-Schwartz backward induction ------------------------------------------
 * replaces _american_exercise_lsm (sims/black_scholes.py:109-193).  `d_paths_tm` is TIME-MAJOR:
 * element (step t, path i) at d_paths_tm[t*n_paths + i] (what mcf_gbm_paths writes with
 * time_major=1; mcf_transpose converts the (n_paths, n_steps+1) row-major matrix simulate_paths
 * returns).  Per step t = n_steps-1 .. 1: mcf_lsm_reduce leaves the 8 normal-equation sums of the
 * in-the-money paths in the first 8 doubles of the workspace (sum them across ranks when paths are
 * sharded), mcf_lsm_apply solves the 3x3 system (pseudo-inverse, like lstsq) and exercises where
 * intrinsic > fitted.  mcf_lsm runs the whole loop on one GPU. */
int64_t mcf_lsm_workspace_bytes(int64_t n_paths);
int mcf_lsm_begin(const double *d_paths_tm, int64_t n_paths, int64_t n_steps, double K, double r, double dt,
                  int32_t is_put, void *d_workspace, int64_t workspace_bytes, void *cuda_stream);
int mcf_lsm_reduce(const double *d_paths_tm, int64_t n_paths, int64_t t, double K, double r, double dt, int32_t is_put,
                   void *d_workspace, void *cuda_stream);
int mcf_lsm_apply(const double *d_paths_tm, int64_t n_paths, int64_t t, double K, double r, double dt, int32_t is_put,
                  void *d_workspace, void *cuda_stream);
/* d_price_sum[0] = sum_i cash_flow_i*exp(-r*dt*exercise_i) over this rank's paths; d_price[0] = that / n_paths.
 * Either may be NULL.  d_exercise_times (int32[n_paths], may be NULL) receives the exercise step of every path. */
int mcf_lsm_finish(int64_t n_paths, double r, double dt, void *d_workspace, double *d_price_sum, double *d_price,
                   int32_t *d_exercise_times, void *cuda_stream);
/* Measurement aid (bench.py's roofline denominator, measured on the device the benchmark runs on): issue rate, in warp
 * instructions per second over the whole GPU, of kind 0: IMAD.WIDE.U32, 1: IMAD (low), 2: LOP3 -- eight independent
 * chains per thread, 4 CTAs x 256 threads per SM (the loop of tools/probe_pipes.cu).  Blocks until measured. */
int mcf_probe_issue_rate(int32_t kind, double *warp_inst_per_s, void *cuda_stream);

/* Paths sharded over the GPUs of ONE box (world <= 8): the same fused step kernel as mcf_lsm, with the exchange of the 8
 * sums of a step done inside the kernel over peer memory -- the last CTA of every rank stores its sums into a slot of
 * every rank's exchange buffer (8-byte words, each tagged with the low 32 bits of the step's sequence number: no fence, no
 * separate flag), waits for the others' and adds them in rank order (no collective call per step).
 * mcf_peer_buffer_create allocates one exchange buffer (mcf_lsm_exchange_bytes()) and returns its 64-byte CUDA IPC
 * handle; the handles are exchanged by the caller (torch.distributed), opened with mcf_peer_buffer_open, and passed as
 * peer_buffers[world] (entry [rank] = the local buffer).  sequence_base: any number that grows by more than n_steps
 * from one sweep to the next (stale slots never match); the caller orders successive sweeps with a stream barrier.
 * mcf_lsm and mcf_lsm_sharded keep the cash flows L2-resident when they fit the device's persisting set-aside (see
 * INTEGRATION.md): such a call raises cudaLimitPersistingL2CacheSize for its duration, restores it, and returns when the
 * sweep has finished; MCF_LSM_L2_PERSIST=0 in the environment turns that off. */
int64_t mcf_lsm_exchange_bytes(void);
int mcf_peer_buffer_create(int64_t bytes, void **d_ptr, void *handle64);
int mcf_peer_buffer_open(const void *handle64, void **d_ptr);
int mcf_peer_buffer_close(void *d_ptr);
int mcf_peer_buffer_destroy(void *d_ptr);
int mcf_lsm_sharded(const double *d_paths_tm, int64_t n_paths_local, int64_t n_paths_global, int64_t n_steps, double K, double r,
                    double dt, int32_t is_put, void *d_workspace, int64_t workspace_bytes, void *const *peer_buffers,
                    int32_t rank, int32_t world, uint64_t sequence_base, double *d_price, int32_t *d_exercise_times,
                    void *cuda_stream);
int mcf_lsm(const double *d_paths_tm, int64_t n_paths, int64_t n_steps, double K, double r, double dt, int32_t is_put,
            void *d_workspace, int64_t workspace_bytes, double *d_price, int32_t *d_exercise_times, void *cuda_stream);
/* (n_rows, n_cols) row-major -> (n_cols, n_rows) row-major */
int mcf_transpose(const double *d_in, int64_t n_rows, int64_t n_cols, double *d_out, void *cuda_stream);

/* ---- (4) StatsEngine reductions --------------------------------------------------------
 * mcf_moments: ONE pass over x (stats_engine.py:734-742 isfinite mask folded into the load) giving
 * everything mean/std/skew/kurtosis/ci_mean/chebyshev/markov/bias/mse need (stats_engine.py:767-1472):
 *   out[0] n_finite  out[1] n_nan  out[2] n_posinf  out[3] n_neginf
 *   out[4] mean      out[5] M2=sum (x-mean)^2   out[6] M3   out[7] M4      (finite values only)
 *   out[8] sum (x-target)^2 (= M2 + n*(mean-target)^2)   out[9] sum x   out[10..11] reserved
 * Per-rank outputs merge on the host with the pairwise update formulas (see _stats_host.py). */
#define MCF_MOMENTS_OUT 12
int64_t mcf_moments_workspace_bytes(int64_t n);
int mcf_moments(const double *d_x, int64_t n, double target, double *d_out, void *d_workspace, int64_t workspace_bytes,
                void *cuda_stream);

/* mcf_select: k order statistics (0-based ranks in the sorted order np.percentile uses: NaN last) by
 * MSD radix selection, replaces the sort inside np.percentile (stats_engine.py:865,1197,1286;
 * core.py:757).  With omit_nonfinite != 0, NaN and +-inf are excluded (nan_policy="omit").
 * Multi-GPU: call begin, then per pass {hist on every rank; all-reduce(SUM, int64) the histogram
 * block [mcf_select_hist_offset, +mcf_select_hist_bytes) of the workspace; choose}, then finish. */
#define MCF_SELECT_MAX_RANKS 32
int64_t mcf_select_workspace_bytes(void);
int64_t mcf_select_hist_offset(void);
int64_t mcf_select_hist_bytes(void);
int32_t mcf_select_passes(void);
int mcf_select_begin(const int64_t *h_ranks, int32_t k, void *d_workspace, int64_t workspace_bytes, void *cuda_stream);
int mcf_select_hist(const double *d_x, int64_t n, int32_t omit_nonfinite, void *d_workspace, void *cuda_stream);
int mcf_select_choose(void *d_workspace, void *cuda_stream);
int mcf_select_finish(const void *d_workspace, double *d_out, void *cuda_stream);
/* Large samples (n >= 2^22), NaN-free semantics: ONE pass over x instead of two radix passes plus a compaction pass.  A
 * strided sub-sample of 2^20 values brackets every target rank (6 standard deviations of the sampling error); the pass
 * counts the keys below / inside every bracket EXACTLY and appends the inside keys (a few per cent of x) to the scratch;
 * the radix passes then run on those keys with rebased ranks.  h_ranks ascending, k <= 16.  *d_invalid != 0 afterwards:
 * not decided here (ties heavier than the scratch, > 8 disjoint brackets, a sampling miss) -- run mcf_select_compacting.
 * d_scratch: mcf_select_bracket_bytes(n).  Synchronises the stream once (uploads the bracket ranks). */
int64_t mcf_select_bracket_bytes(int64_t n);
int mcf_select_bracketed(const double *d_x, int64_t n, const int64_t *h_ranks, int32_t k, double *d_out, void *d_workspace,
                         int64_t workspace_bytes, void *d_scratch, int64_t scratch_bytes, int32_t *d_invalid,
                         void *cuda_stream);
int mcf_select(const double *d_x, int64_t n, int32_t omit_nonfinite, const int64_t *h_ranks, int32_t k, double *d_out,
               void *d_workspace, int64_t workspace_bytes, void *cuda_stream);
/* The same selection with a scratch buffer of mcf_select_compact_bytes(n) bytes (d_compact may be NULL = mcf_select): after
 * the second pass the candidates that still share a tracked 16-bit prefix (a few per cent of a smooth sample) are
 * compacted and the six remaining passes read only those; if they do not fit (heavily tied data) the passes keep
 * reading x.  Results are identical. */
int64_t mcf_select_compact_bytes(int64_t n);
int mcf_select_compacting(const double *d_x, int64_t n, int32_t omit_nonfinite, const int64_t *h_ranks, int32_t k,
                          double *d_out, void *d_workspace, int64_t workspace_bytes, void *d_compact, int64_t compact_bytes,
                          void *cuda_stream);

/* Bootstrap means, replaces _bootstrap_means (stats_engine.py:1019-1041):
 *   idx = rng.integers(0, n, size=(B, n)); means = x[idx].mean(axis=1)
 * The index matrix is the sequence of ACCEPTED 32-bit slots of the stream (Lemire rejection decides each
 * slot on its own value), so slots [slot_lo, slot_hi) can be processed independently:
 *   mcf_bootstrap_count      : accepted slots per chunk + their scan; *d_total = accepted in the range
 *   mcf_bootstrap_accumulate : regenerate the slots, gather x[idx], add into d_sums[ordinal / n] where
 *                              ordinal = ordinal_base + (accepted slots before it in the range);
 *                              ordinals >= B*n are ignored; the thread that produces ordinal B*n-1
 *                              stores (slot index + 1) in *d_end_slot (stream consumption of the call).
 * Multi-GPU: shard the slot range, all-gather the per-rank totals to get each rank's ordinal_base,
 * all-reduce d_sums.  d_sums must be zeroed by the caller; mcf_bootstrap_finalize divides by n.
 * Needs 2 <= n < 2^32 (NumPy's 32-bit path; n == 1 consumes no randomness). */
/* device-wide hint for L2's DRAM fetch size on a miss (32/64/128 bytes); the bootstrap's random gathers want 32 */
int mcf_set_l2_fetch_granularity(int32_t bytes);
int64_t mcf_bootstrap_workspace_bytes(int64_t n_slots, int32_t chunk_slots);
int mcf_bootstrap_count(uint32_t gen_kind, const mcf_stream *d_stream, int64_t n, uint64_t slot_lo, uint64_t slot_hi,
                        int32_t chunk_slots, void *d_workspace, int64_t workspace_bytes, uint64_t *d_total,
                        void *cuda_stream);
int mcf_bootstrap_accumulate(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n,
                             int64_t n_resamples, uint64_t slot_lo, uint64_t slot_hi, int32_t chunk_slots,
                             uint64_t ordinal_base, const void *d_workspace, int64_t workspace_bytes, double *d_sums,
                             uint64_t *d_end_slot, void *cuda_stream);
/* Single-pass form of the binned bootstrap.  The ordinal at slot s is s*(1-p) to within a few standard deviations of a
 * binomial count, so a CTA whose estimated ordinal range (+- 8 sigma) lies strictly inside ONE resample -- 96 % of them at
 * n = 1e8 -- bins into that resample in its only pass over its slots and counts what it accepted on the way; only the CTAs
 * near resample boundaries and near the end of the draw are counted first and binned with exact ordinals.  phase1: safe
 * CTAs (bin + gather, batch by batch), count of the others, prefix sums; *d_total = accepted slots of [slot_lo, slot_hi).
 * phase2 (ordinal_base = accepted slots of all lower ranks): the exact CTAs, then every safe CTA is checked against its
 * exact ordinal range: *d_flag != 0 (an 8-sigma miss) or *d_end_slot == 0 (window too small) => recompute with
 * mcf_bootstrap_count + mcf_bootstrap_accumulate_binned.  slot_total: slots of ALL ranks (sets the margin).  Workspaces as
 * for the two-pass form.  Both calls synchronise the stream. */
int mcf_bootstrap_fused_phase1(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n, int64_t n_resamples,
                               uint64_t slot_lo, uint64_t slot_hi, uint64_t slot_total, int32_t chunk_slots, void *d_workspace,
                               int64_t workspace_bytes, int32_t resamples_per_batch, int32_t region_shift,
                               void *d_bin_workspace, int64_t bin_workspace_bytes, double *d_sums, uint64_t *d_total,
                               void *cuda_stream);
int mcf_bootstrap_fused_phase2(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n, int64_t n_resamples,
                               uint64_t slot_lo, uint64_t slot_hi, uint64_t slot_total, int32_t chunk_slots,
                               uint64_t ordinal_base, void *d_workspace, int64_t workspace_bytes, int32_t resamples_per_batch,
                               int32_t region_shift, void *d_bin_workspace, int64_t bin_workspace_bytes, double *d_sums,
                               uint64_t *d_end_slot, int32_t *d_flag, void *cuda_stream);
int mcf_bootstrap_finalize(double *d_sums, int64_t n_resamples, int64_t n, void *cuda_stream);
/* Binned variant of mcf_bootstrap_accumulate for populations larger than L2 (a random 8-byte gather from HBM moves a
 * 128-byte line; from an L2-resident region it is 3-4x faster): the accepted indices of `resamples_per_batch`
 * resamples are first appended (4 bytes each) to per-(resample, 2^region_shift-element region of x) queues, then
 * summed region by region.  Same index stream, bit for bit; only the order of a resample's additions differs.
 * h_block_base: host copy of the per-CTA ordinal prefix of the count pass (mcf_bootstrap_blocks entries, filled by
 * mcf_bootstrap_block_base).  region_shift 0 = automatic (2^22-element = 32 MB regions, doubled until at most 32 regions cover x).  Needs 256*chunk_slots <= n,
 * chunk_slots % 8 == 0.  The workspace holds TWO queue sets: the gathers of one batch run on an internal
 * high-priority side stream while the bin pass of the next batch fills the other set on the caller's stream; the
 * side stream is joined into the caller's before the call returns.  Synchronises the stream (reads the
 * queue-overflow flag). */
int64_t mcf_bootstrap_blocks(int64_t n_slots, int32_t chunk_slots);
int mcf_bootstrap_block_base(const void *d_workspace, int64_t n_slots, int32_t chunk_slots, uint64_t *h_block_base,
                             void *cuda_stream);
int64_t mcf_bootstrap_binned_workspace_bytes(int64_t n, int32_t resamples_per_batch, int32_t region_shift);
int mcf_bootstrap_accumulate_binned(uint32_t gen_kind, const mcf_stream *d_stream, const double *d_x, int64_t n,
                                    int64_t n_resamples, uint64_t slot_lo, uint64_t slot_hi, int32_t chunk_slots,
                                    uint64_t ordinal_base, const void *d_workspace, int64_t workspace_bytes,
                                    const uint64_t *h_block_base, int32_t resamples_per_batch, int32_t region_shift,
                                    void *d_bin_workspace, int64_t bin_workspace_bytes, double *d_sums,
                                    uint64_t *d_end_slot, void *cuda_stream);
/* *d_count = #{i : d_x[i] < value}  (BCa bias correction z0, stats_engine.py:1075-1078) */
int mcf_count_less(const double *d_x, int64_t n, double value, int64_t *d_count, void *cuda_stream);
/* nan_policy="omit" (_clean, stats_engine.py:734-742): d_out[0..*d_count) = finite values of d_x, order kept. */
int64_t mcf_compact_workspace_bytes(int64_t n);
int mcf_compact_finite(const double *d_x, int64_t n, double *d_out, int64_t *d_count, void *d_workspace,
                       int64_t workspace_bytes, void *cuda_stream);
int mcf_stats_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MCF_B200_H */
