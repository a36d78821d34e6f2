"""MonteCarloSimulation / MonteCarloFramework / SimulationResult with the replication loop on the GPU.

Drop-in for the reference's ``mcframework/core.py`` (citations are to that file): same classes,
signatures, validation messages, metadata keys, seeding side effects and parallel block layout.
What changes is WHERE the replication loop runs.  A built-in simulation describes its
``single_simulation`` hook as a *replication program* (``_replication_program``); ``run`` then
describes the random streams the reference would have consumed -- the single PCG64 stream of
``self.rng`` for a sequential run (:583-597), or one Philox stream per ``SeedSequence``-spawned
block for a parallel run (:622-641, :657-663) -- and the CUDA kernels replay exactly those streams
(``_device`` -> C ABI).  Results stay on the device for the statistics and are copied to the host
once, as the ``results`` ndarray of the :class:`SimulationResult`.

A subclass whose hook is arbitrary Python (no replication program) is user plugin code: it is looped
on the host exactly as the reference loops it (sequentially or on thread/process pools), because a
Python callable cannot be turned into a kernel.  That loop is API compatibility for user hooks, not
a backend for the built-in simulations: those have no host implementation at all and raise when the
CUDA library or a device is missing.
"""
from __future__ import annotations

import logging
import multiprocessing as mp
import sys
import time
import weakref
from abc import ABC, abstractmethod
from concurrent.futures import ProcessPoolExecutor, ThreadPoolExecutor, as_completed
from dataclasses import dataclass, field
from typing import Any, Callable, Iterable, Mapping

import numpy as np

from ._native import DeviceUnavailableError, NativeError
from .stats_engine import DEFAULT_ENGINE, DeviceSample, StatsContext, StatsEngine, _ensure_ctx, ci_mean, mean, std
from .stats_engine import percentiles as pct
from .utils import autocrit

logger = logging.getLogger(__name__)
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s - %(message)s"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)


def _is_windows_platform() -> bool:
    return sys.platform.startswith("win") or sys.platform == "cli"


def _worker_run_chunk(sim: "MonteCarloSimulation", chunk_size: int, seed_seq: np.random.SeedSequence,
                      simulation_kwargs: dict[str, Any]) -> list[float]:
    """Process-pool worker for USER hooks (:71-104): one Philox stream per chunk."""
    local_rng = np.random.Generator(np.random.Philox(seed_seq))
    return [float(sim.single_simulation(_rng=local_rng, **simulation_kwargs)) for _ in range(chunk_size)]


def make_blocks(n: int, block_size: int = 10_000) -> list[tuple[int, int]]:
    """Half-open ``(start, stop)`` blocks covering ``range(n)`` (:107-134)."""
    return [(lo, min(lo + block_size, n)) for lo in range(0, n, block_size)] if n > 0 else []


# Page-locked host memory held by results the caller has not released yet (bytes), and the budget beyond which new
# results are copied into pageable memory instead (a caller that keeps hundreds of large results must not pin them all).
PINNED_RESULT_BUDGET = 16 << 30
_PINNED_LIVE = [0]


def _unpin_accounting(nbytes: int) -> None:
    _PINNED_LIVE[0] -= nbytes


class DeviceResults:
    """Result vector left in HBM (opt-in: ``sim.keep_results_on_device = True``).

    The reference hands back a host ndarray (``core.py:137-147``); between the two halves of the hot path --
    replications and statistics -- that is an 800 MB device-to-host copy nobody reads when only the summary is used.
    This object stands in for the ndarray: ``.device`` is the float64 CUDA tensor, and anything that needs host
    values (``np.asarray``, indexing, iteration, ``.numpy()``) materialises the host copy ONCE, on first use.
    """

    __array_priority__ = 100

    def __init__(self, dev):
        self._dev = dev
        self._host = None

    @property
    def device(self):
        return self._dev

    def numpy(self) -> np.ndarray:
        if self._host is None:
            self._host = MonteCarloSimulation._to_host(self._dev)
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    @property
    def materialised(self) -> bool:
        return self._host is not None

    size = property(lambda self: int(self._dev.numel()))
    shape = property(lambda self: (int(self._dev.numel()),))
    ndim = 1
    dtype = np.dtype(np.float64)

    def __len__(self) -> int:
        return int(self._dev.numel())

    def __getitem__(self, key):
        return self.numpy()[key]

    def __iter__(self):
        return iter(self.numpy())

    def __repr__(self) -> str:
        return f"DeviceResults(n={len(self)}, materialised={self.materialised})"


@dataclass
class SimulationResult:
    """Outcome of one run (:137-238).  ``results`` is a host float64 ndarray owned by this object."""

    results: np.ndarray
    n_simulations: int
    execution_time: float
    mean: float
    std: float
    percentiles: dict[int, float]
    stats: dict[str, Any] = field(default_factory=dict)
    metadata: dict[str, Any] = field(default_factory=dict)

    def result_to_string(self, confidence: float = 0.95, method: str = "auto") -> str:
        print("=" * 20 + " SIM RESULTS " + "=" * 20)
        name = self.metadata.get("simulation_name")
        out = [f"Results for simulation '{name}':" if name else "Results for simulation:"]
        n = int(self.n_simulations)
        crit, kind = autocrit(confidence, n, method)
        se = self.std / np.sqrt(max(1, n))
        lo, hi = self.mean - crit * se, self.mean + crit * se
        out.append(f"  Number of simulations: {self.n_simulations}")
        out.append(f"  Execution time: {self.execution_time:.2f} seconds")
        out.append(f"  Mean: {self.mean:.5f}   (SE: {se:.5f}, {int(confidence * 100)}% {kind}-CI: [{lo:.5f}, {hi:.5f}])")
        out.append(f"  Std Dev (sample): {self.std:.5f}")
        out.append("  Percentiles:")
        out.extend(f"    {p}th: {self.percentiles[p]:.5f}" for p in sorted(self.percentiles))
        ci = self.stats.get("ci_mean")
        if isinstance(ci, (tuple, list)) and len(ci) == 2 and all(isinstance(v, (int, float)) for v in ci):
            out.append(f"  (engine) CI: [{ci[0]:.5f}, {ci[1]:.5f}]")
        if self.stats:
            out.append("Additional Stats:")
            out.extend(f"  {k}: {v}" for k, v in self.stats.items())
        if self.metadata:
            out.append("Metadata:")
            out.extend(f"    {k}: {v}" for k, v in self.metadata.items())
        out.append("=" * 20 + " END " + "=" * 20)
        return "\n".join(out)


class MonteCarloSimulation(ABC):
    """Template-method base class: subclasses implement ``single_simulation`` (:241-855)."""

    _PCTS = (5, 25, 50, 75, 95)
    _PARALLEL_THRESHOLD = 20_000
    _CHUNKS_PER_WORKER = 8
    #: block-parallel runs with at least this many replications copy their result vector to the host in
    #: ``_PIPELINE_GROUPS`` pieces, each behind the kernels of the next group of streams (one GPU, host-resident result)
    _PIPELINE_MIN_SIMS = 1 << 24
    _PIPELINE_GROUPS = 4

    @staticmethod
    def _rng(rng: np.random.Generator | None, default: np.random.Generator | None = None) -> np.random.Generator:
        return rng if rng is not None else default

    def __init__(self, name: str = "Simulation"):
        self.name = name
        self.seed_seq: np.random.SeedSequence | None = None
        self.rng = np.random.default_rng()
        self.parallel_backend: str = "auto"

    # pickling (process backend for user hooks): generators are rebuilt from the seed sequence (:318-330)
    def __getstate__(self):
        state = self.__dict__.copy()
        state["rng"] = None
        state.pop("_device_sample", None)
        if not isinstance(state.get("shard_group"), (str, type(None))):
            state.pop("shard_group")  # a process group does not travel to worker processes
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.rng = np.random.default_rng(self.seed_seq) if self.seed_seq is not None else np.random.default_rng()

    @abstractmethod
    def single_simulation(self, *args, **kwargs) -> float:
        """One replication; receives ``_rng`` when run on a worker stream."""

    def set_seed(self, seed: int | None) -> None:
        self.seed_seq = np.random.SeedSequence(seed)
        self.rng = np.random.default_rng(self.seed_seq)

    # ------------------------------------------------------------------ device dispatch
    def _replication_program(self, **simulation_kwargs):
        """Built-in simulations return a ``_programs.ReplicationProgram`` describing their hook; ``None``
        means "arbitrary Python hook" (looped on the host as user plugin code)."""
        return None

    def draw_program(self, **simulation_kwargs):
        """Optional declaration (``mcframework_b200.programs``) of what ``single_simulation`` draws and computes;
        ``None`` (default) means the hook is arbitrary Python."""
        return None

    def _program_for(self, simulation_kwargs):
        declared = self.draw_program(**simulation_kwargs)
        if declared is not None:
            return declared
        prog = self._replication_program(**simulation_kwargs)
        if prog is None:
            return None
        owner = getattr(type(self), "_program_hook_owner", None)
        if owner is None or type(self).single_simulation is not owner.single_simulation:
            return None  # the subclass replaced the hook: its Python body is the definition now
        return prog

    #: Multi-GPU sharding is OPT-IN: set ``sim.shard_group`` to a ``torch.distributed`` process group (one process per
    #: GPU), or to the string ``"world"`` for the default group.  While it is set, ``run(parallel=True)`` and
    #: ``calculate_greeks(parallel=True)`` are COLLECTIVE calls: every rank of the group must make them with identical
    #: arguments (``n_simulations``, ``n_workers``, seed, simulation kwargs); each rank replays a contiguous range of the
    #: spawned blocks.  ``None`` (default): an initialised process group is ignored and every process runs the whole job.
    shard_group = None

    def _shard_group(self):
        """Process group over which a parallel run is sharded (one process per GPU), else None."""
        group = getattr(self, "shard_group", None)
        if group is None:
            return None
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("shard_group is set but torch.distributed is not initialised")
        if isinstance(group, str):
            if group != "world":
                raise ValueError("shard_group must be a process group, 'world' or None")
            group = dist.group.WORLD
        return group if dist.get_world_size(group) > 1 else None

    # ------------------------------------------------------------------ validation / stats
    def _validate_run_params(self, n_simulations: int, n_workers: int | None, confidence: float, ci_method: str) -> None:
        if n_simulations <= 0:
            raise ValueError("n_simulations must be positive")
        if n_workers is not None and n_workers <= 0:
            raise ValueError("n_workers must be positive")
        if not (0.0 < confidence < 1.0):
            raise ValueError("confidence must be in the interval (0, 1)")
        if ci_method not in ("auto", "z", "t", "bootstrap"):
            raise ValueError(f"ci_method must be one of 'auto', 'z', 't', 'bootstrap', got '{ci_method}'")

    def _compute_stats_with_engine(self, results, n_simulations: int, confidence: float, ci_method: str,
                                   stats_engine: StatsEngine | None, extra_context: Mapping[str, Any] | None
                                   ) -> tuple[dict[str, Any], dict[int, float]]:
        """Engine metrics merged over the baseline block; engine wins on clashes (:383-450)."""
        eng = stats_engine or DEFAULT_ENGINE
        if eng is None:
            return {}, {}
        from .stats_engine import CIMethod

        method = CIMethod(ci_method)
        base_kwargs = dict(n=n_simulations, percentiles=self._PCTS, confidence=confidence, ci_method=method)
        try:
            ctx = StatsContext(**base_kwargs, **(dict(extra_context) if extra_context else {}))
        except (TypeError, ValueError) as exc:
            logger.warning("Invalid context parameters: %s. Using defaults.", exc)
            ctx = StatsContext(**base_kwargs)
        sample = self._sample_for(results)
        try:
            outcome = eng.compute(sample, ctx)
            stats = outcome.metrics if hasattr(outcome, "metrics") else {}
        except (DeviceUnavailableError, NativeError):
            raise
        except Exception as exc:  # noqa: BLE001
            logger.error("Stats engine failed: %s", exc)
            stats = {}
        merged = dict(self._compute_stats_block(sample, ctx))
        if isinstance(stats, dict):
            merged.update(stats)
        engine_pcts = merged.pop("percentiles", None) or {}
        return merged, {int(k): float(v) for k, v in engine_pcts.items()}

    def _sample_for(self, results):
        """The device-resident sample of this run when it exists (no re-upload), else the host array."""
        cached = getattr(self, "_device_sample", None)
        if cached is not None and cached.host is results:
            return cached
        return results

    def _handle_percentiles(self, results, percentiles: Iterable[int] | None, compute_stats: bool,
                            percentile_map: dict[int, float]) -> tuple[dict[int, float], list[int], bool]:
        """Merge user-requested percentiles with the engine defaults (:452-484)."""
        given = percentiles is not None
        wanted = tuple(int(p) for p in (percentiles or ()))
        if not compute_stats:
            table = self._percentiles(self._sample_for(results), wanted) if (given and wanted) else {}
            return table, (list(wanted) if given else []), False
        if wanted:
            percentile_map.update(self._percentiles(self._sample_for(results), wanted))
        return percentile_map, list(wanted), True

    # ------------------------------------------------------------------ run
    def run(self, n_simulations: int, *, parallel: bool = False, n_workers: int | None = None,
            progress_callback: Callable[[int, int], None] | None = None, percentiles: Iterable[int] | None = None,
            compute_stats: bool = True, stats_engine: StatsEngine | None = None, confidence: float = 0.95,
            ci_method: str = "auto", extra_context: Mapping[str, Any] | None = None,
            **simulation_kwargs: Any) -> SimulationResult:
        """Run ``n_simulations`` replications and summarise them (:486-581)."""
        self._validate_run_params(n_simulations, n_workers, confidence, ci_method)
        self._device_sample = None
        try:
            started = time.time()
            if parallel:
                if n_workers is None:
                    n_workers = mp.cpu_count()
                logger.info("Computing %d simulations in parallel using %d workers...", n_simulations, n_workers)
                results = self._run_parallel(n_simulations, n_workers, progress_callback, **simulation_kwargs)
            else:
                logger.info("Computing %d simulations sequentially...", n_simulations)
                results = self._run_sequential(n_simulations, progress_callback, **simulation_kwargs)
            elapsed = time.time() - started

            stats: dict[str, Any] = {}
            pmap: dict[int, float] = {}
            if compute_stats:
                stats, pmap = self._compute_stats_with_engine(results, n_simulations, confidence, ci_method, stats_engine,
                                                              extra_context)
            pmap, requested, defaults_used = self._handle_percentiles(results, percentiles, compute_stats, pmap)
            return self._create_result(results, n_simulations, elapsed, pmap, stats, requested, defaults_used)
        finally:
            self._device_sample = None  # also when the statistics raise: the device copy must not outlive the call

    # ------------------------------------------------------------------ replication loop
    @staticmethod
    def _to_host(t) -> np.ndarray:
        """Device vector -> fresh host ndarray.  Large vectors go through pinned memory (PyTorch's caching host
        allocator recycles the buffer once the previous result is released): ~50 GB/s instead of ~2 GB/s pageable."""
        import torch

        nbytes = int(t.numel()) * int(t.element_size())
        # results the caller still holds stay page-locked: beyond the budget further ones are pageable (slower, unbounded)
        pin = t.numel() >= (1 << 20) and _PINNED_LIVE[0] + nbytes <= PINNED_RESULT_BUDGET
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin)
        if pin:
            _PINNED_LIVE[0] += nbytes
            weakref.finalize(host, _unpin_accounting, nbytes)
        host.copy_(t, non_blocking=False)
        return host.numpy()

    def _finish_device_run(self, values_dev, group=None, host=None) -> np.ndarray:
        """Device results -> (a) DeviceSample kept for the statistics, (b) the host ndarray of the result
        (``host``: already copied by the pipelined block run)."""
        if host is not None:  # already copied out (pipelined run); sharded runs keep the group for the statistics
            self._device_sample = DeviceSample(values_dev, host, group)
        elif group is not None:
            from . import _multi

            # "all": every rank returns the full vector (reference semantics on every rank);
            # "root": only rank 0 does, the others return their own shard (saves N-1 host copies);
            # "shm": as "root", without funnelling the vector through rank 0's GPU (opt-in)
            local = values_dev
            mode = getattr(self, "gather_mode", "all")
            if mode == "shm":  # every rank copies its own shard into one shared-memory buffer (see _multi)
                full = _multi.gather_results_shared(local, group)
                host = full if full is not None else self._to_host(local)
            elif mode == "root":
                full = _multi.gather_results_to_root(local, group)
                host = self._to_host(full if full is not None else local)
            else:
                host = self._to_host(_multi.gather_results(local, group))
            self._device_sample = DeviceSample(local, host, group)
        elif getattr(self, "keep_results_on_device", False):
            host = DeviceResults(values_dev)  # lazy host copy; the statistics read the device tensor
            self._device_sample = DeviceSample(values_dev, host, None)
        else:
            host = self._to_host(values_dev)
            self._device_sample = DeviceSample(values_dev, host, None)
        return host

    def _run_sequential(self, n_simulations: int, progress_callback: Callable[[int, int], None] | None,
                        **simulation_kwargs: Any) -> np.ndarray:
        """All replications draw from ``self.rng`` one after another (:583-597)."""
        step = max(1, n_simulations // 100)
        prog = self._program_for(simulation_kwargs)
        if prog is not None and isinstance(getattr(self.rng, "bit_generator", None), np.random.PCG64):
            from . import _programs

            try:
                values, draws = _programs.run_sequential(prog, self.rng, n_simulations)
            except _programs.HostLoopRequired:
                values = None  # the generator was not advanced: loop the hook below, exactly like the reference
            if values is not None:
                if draws:  # leave the stream where the reference would (64-bit draws never touch the buffered half)
                    from ._streams import advance_keeping_buffer

                    advance_keeping_buffer(self.rng, int(draws))
                results = self._finish_device_run(values)
                if progress_callback:
                    for done in range(step, n_simulations + 1, step):
                        progress_callback(done, n_simulations)
                    if n_simulations % step:
                        progress_callback(n_simulations, n_simulations)
                return results
        results = np.empty(n_simulations, dtype=float)
        for i in range(n_simulations):
            results[i] = float(self.single_simulation(**simulation_kwargs))
            if progress_callback and ((i + 1) % step == 0 or i + 1 == n_simulations):
                progress_callback(i + 1, n_simulations)
        return results

    def _resolve_parallel_backend(self, requested: str | None = None) -> str:
        backend = requested or getattr(self, "parallel_backend", "auto")
        if backend not in ("auto", "thread", "process"):
            logger.warning("Unknown parallel backend '%s'; defaulting to 'auto'", backend)
            backend = "auto"
        if backend != "auto":
            return backend
        if _is_windows_platform():
            logger.info("Parallel backend 'auto' resolved to 'process' on Windows platform.")
            return "process"
        return "thread"

    def _prepare_parallel_blocks(self, n_simulations: int, n_workers: int
                                 ) -> tuple[list[tuple[int, int]], list[np.random.SeedSequence]]:
        """Block partition + one spawned child sequence per block (:622-641)."""
        size = max(1, n_simulations // (n_workers * self._CHUNKS_PER_WORKER))
        blocks = make_blocks(n_simulations, size)
        if self.seed_seq is not None:
            children = self.seed_seq.spawn(len(blocks))
        else:
            children = [np.random.SeedSequence() for _ in blocks]
        return blocks, children

    def _run_with_threads(self, blocks, child_seqs, n_simulations: int, n_workers: int, progress_callback,
                          **simulation_kwargs: Any) -> np.ndarray:
        """Thread pool over blocks, for USER hooks (:643-674)."""
        results = np.empty(n_simulations, dtype=float)

        def work(job):
            (a, b), child = job
            local = np.random.Generator(np.random.Philox(child))
            vals = np.empty(b - a, dtype=float)
            for k in range(vals.size):
                vals[k] = float(self.single_simulation(_rng=local, **simulation_kwargs))
            return (a, b), vals

        done = 0
        with ThreadPoolExecutor(max_workers=min(n_workers, len(blocks))) as pool:
            pending = [pool.submit(work, job) for job in zip(blocks, child_seqs)]
            for fut in as_completed(pending):
                (a, b), vals = fut.result()
                results[a:b] = vals
                done += b - a
                if progress_callback:
                    progress_callback(done, n_simulations)
        return results

    def _run_with_processes(self, blocks, child_seqs, n_simulations: int, n_workers: int, progress_callback,
                            **simulation_kwargs: Any) -> np.ndarray:
        """Spawned process pool over blocks, for USER hooks (:676-712)."""
        results = np.empty(n_simulations, dtype=float)
        done = 0
        with ProcessPoolExecutor(max_workers=min(n_workers, len(blocks)), mp_context=mp.get_context("spawn")) as pool:
            pending = {}
            for (a, b), child in zip(blocks, child_seqs):
                pending[pool.submit(_worker_run_chunk, self, b - a, child, dict(simulation_kwargs))] = (a, b)
            try:
                for fut in as_completed(pending):
                    a, b = pending[fut]
                    results[a:b] = fut.result()
                    done += b - a
                    if progress_callback:
                        progress_callback(done, n_simulations)
            except KeyboardInterrupt:
                for fut in pending:
                    fut.cancel()
                raise
        return results

    def _run_parallel(self, n_simulations: int, n_workers: int | None, progress_callback, **simulation_kwargs: Any
                      ) -> np.ndarray:
        """Block-parallel run (:714-752); below the threshold or with one worker it is the sequential run."""
        if n_workers is None:
            n_workers = mp.cpu_count()
        if n_workers <= 1 or n_simulations < self._PARALLEL_THRESHOLD:
            return self._run_sequential(n_simulations, progress_callback, **simulation_kwargs)
        blocks, children = self._prepare_parallel_blocks(n_simulations, n_workers)
        backend = self._resolve_parallel_backend()
        prog = self._program_for(simulation_kwargs)
        if prog is not None:
            from . import _programs

            group = self._shard_group()
            host = None
            big = (prog.kind != "integers" and not getattr(self, "keep_results_on_device", False)
                   and n_simulations >= self._PIPELINE_MIN_SIMS and len(blocks) >= 2)
            try:
                if group is None and big:
                    # large host-resident result: copy it out group by group while the next group is computed
                    values, host = _programs.run_blocks_pipelined(prog, blocks, children, n_simulations,
                                                                  self._PIPELINE_GROUPS)
                elif group is not None and big and getattr(self, "gather_mode", "all") == "shm":
                    # sharded run, one host: every rank pipelines ITS shard straight into its slice of one shared,
                    # page-locked result buffer, over its own PCIe link -- no gather, one barrier
                    import torch.distributed as dist

                    from . import _multi

                    sizes = _programs.shard_sizes(blocks, dist.get_world_size(group))
                    whole, mine, seg = _multi.shared_result_buffer(sizes, group)
                    values, _ = _programs.run_blocks_pipelined(
                        prog, blocks, children, n_simulations, self._PIPELINE_GROUPS,
                        shard=(dist.get_rank(group), dist.get_world_size(group)), host_out=mine)
                    host = _multi.finish_shared_result(whole, seg, sizes, group)
                else:
                    values = _programs.run_blocks(prog, blocks, children, n_simulations, group)
            except _programs.HostLoopRequired:
                values = None  # spawned children are untouched: loop the hook over the same blocks below
        if prog is not None and values is not None:
            results = self._finish_device_run(values, group, host)
            if progress_callback:
                done = 0
                for a, b in blocks:
                    done += b - a
                    progress_callback(done, n_simulations)
            return results
        runner = self._run_with_threads if backend == "thread" else self._run_with_processes
        return runner(blocks, children, n_simulations, n_workers, progress_callback, **simulation_kwargs)

    # ------------------------------------------------------------------ result assembly
    @staticmethod
    def _percentiles(arr, ps: Iterable[int]) -> dict[int, float]:
        """``{p: percentile}`` of the results (:754-757), order statistics selected on the device."""
        ps = [int(p) for p in ps]
        if not ps:
            return {}
        from . import _stats_device as SD

        sample = DeviceSample.of(arr)
        s = sample.summary(None)
        vals = SD.percentiles(sample.dev, ps, n_valid=int(s.n + s.n_nonfinite), has_nan=s.n_nan > 0, group=sample.group)
        return {p: float(v) for p, v in zip(ps, vals)}

    @staticmethod
    def _compute_stats_block(results, ctx) -> dict[str, object]:
        """Baseline mean / std / z-t interval (:759-783); unguarded like the reference."""
        ctx = _ensure_ctx(ctx, results)
        sample = results if isinstance(results, DeviceSample) else np.asarray(results, dtype=float).ravel()
        if sample.size == 0:
            nan = float("nan")
            return {"mean": nan, "std": nan, "ci_mean": (nan, nan)}
        m, s, ci = mean(sample, ctx), std(sample, ctx), ci_mean(sample, ctx)
        return {
            "mean": float(m) if m is not None else float("nan"),
            "std": float(s) if s is not None else float("nan"),
            "ci_mean": (float(ci["low"]), float(ci["high"])),
            "confidence": float(ci["confidence"]),
            "method": ci["method"],
            "se": float(ci["se"]),
            "crit": float(ci["crit"]),
        }

    @staticmethod
    def _compute_percentiles_block(results, ctx) -> dict[float, float]:
        ctx = _ensure_ctx(ctx, results)
        wanted = list(getattr(ctx, "percentiles", None) or getattr(ctx, "requested_percentiles", None) or [])
        if not wanted:
            return {}
        vals = pct(results if isinstance(results, DeviceSample) else np.asarray(results, dtype=float).ravel(), ctx)
        if isinstance(vals, Mapping):
            return {float(q): float(vals[q]) for q in wanted}
        flat = np.asarray(vals, dtype=float).ravel()
        if flat.size != len(wanted):
            raise ValueError("pct() must return as many values as requested percentiles")
        return {float(q): float(v) for q, v in zip(wanted, flat)}

    def _create_result(self, results: np.ndarray, n_simulations: int, execution_time: float,
                       percentiles: dict[int, float], stats: dict[str, Any], requested_percentiles: list[int],
                       engine_defaults_used: bool) -> SimulationResult:
        """Assemble the result object; headline mean / sample std from the fused moments pass (:807-855)."""
        sample = self._sample_for(results)
        if isinstance(sample, DeviceSample):
            s = sample.summary(None)
            if s.n_nonfinite:
                mean_val = float(mean(sample, StatsContext(n=n_simulations)))
                std_val = float("nan") if results.size > 1 else 0.0
            else:
                mean_val = float(s.mean)
                std_val = float(np.sqrt(s.m2 / (s.n - 1.0))) if s.n > 1 else 0.0
        else:  # user hook looped on the host: summarise its (host) results on the device as well
            sample = DeviceSample.of(results)
            s = sample.summary(None)
            clean = s.n_nonfinite == 0
            mean_val = float(s.mean) if clean else float(mean(sample, StatsContext(n=n_simulations)))
            std_val = (float(np.sqrt(s.m2 / (s.n - 1.0))) if clean else float("nan")) if results.size > 1 else 0.0
        stats = dict(stats) if stats else {}
        for k, v in (stats.pop("percentiles", None) or {}).items():
            percentiles.setdefault(int(k), float(v))
        meta = {
            "simulation_name": self.name,
            "timestamp": time.time(),
            "n": n_simulations,
            "seed_entropy": self.seed_seq.entropy if self.seed_seq else None,
            "requested_percentiles": requested_percentiles,
            "engine_defaults_used": engine_defaults_used,
        }
        return SimulationResult(results=results, n_simulations=n_simulations, execution_time=execution_time,
                                mean=mean_val, std=std_val, percentiles=percentiles, stats=stats, metadata=meta)


class MonteCarloFramework:
    """Registry of named simulations and of their latest results (:858-989)."""

    def __init__(self):
        self.simulations: dict[str, MonteCarloSimulation] = {}
        self.results: dict[str, SimulationResult] = {}

    def register_simulation(self, simulation: MonteCarloSimulation, name: str | None = None):
        self.simulations[name or simulation.name] = simulation

    def run_simulation(self, name: str, n_simulations: int, **kwargs) -> SimulationResult:
        sim = self.simulations.get(name)
        if sim is None:
            raise ValueError(f"Simulation '{name}' not found")
        outcome = sim.run(n_simulations, **kwargs)
        self.results[name] = outcome
        return outcome

    def compare_results(self, names: list[str], metric: str = "mean") -> dict[str, float]:
        table: dict[str, float] = {}
        for name in names:
            if name not in self.results:
                raise ValueError(f"No results found for simulation '{name}'")
            r = self.results[name]
            if metric == "mean":
                table[name] = r.mean
            elif metric == "std":
                table[name] = r.std
            elif metric == "var":
                table[name] = r.std ** 2
            elif metric == "se":
                table[name] = r.std / np.sqrt(max(1, r.n_simulations))
            elif metric.lower().startswith("p") and metric[1:].isdigit():
                p = int(metric[1:])
                asked = r.metadata.get("requested_percentiles")
                if (asked and p not in {int(v) for v in asked}) or p not in r.percentiles:
                    raise ValueError(f"Percentile {p} not computed")
                table[name] = r.percentiles[p]
            else:
                raise ValueError(f"Unknown metric: {metric}")
        return table


__all__ = ["SimulationResult", "MonteCarloSimulation", "MonteCarloFramework", "make_blocks", "_worker_run_chunk"]
