"""Host-side description of the reference's random streams for the device kernels.

The reference seeds everything through NumPy (``core.py:363-364`` ``SeedSequence`` ->
``default_rng`` = PCG64; ``core.py:637,659`` ``seed_seq.spawn`` -> ``Philox`` per block).  We call
NumPy's *own* ``SeedSequence.spawn`` / ``generate_state`` / ``bit_generator.state`` here, so the
spawn tree (including the parent's ``n_children_spawned`` side effect) is identical by
construction, and hand the raw generator states to the GPU as ``mcf_stream`` records
(``include/mcf_b200.h``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

GEN_PCG64 = 0
GEN_PHILOX = 1

#: mirrors ``struct mcf_stream`` (80 bytes)
STREAM_DTYPE = np.dtype(
    [
        ("kind", np.uint32),
        ("buf_pos", np.uint32),
        ("has_uint32", np.uint32),
        ("uinteger", np.uint32),
        ("s", np.uint64, (6,)),
        ("first_sim", np.int64),
        ("n_sims", np.int64),
    ],
    align=True,
)
assert STREAM_DTYPE.itemsize == 80, STREAM_DTYPE.itemsize

_M64 = (1 << 64) - 1


def block_edges(n: int, block_size: int) -> list[tuple[int, int]]:
    """Half-open blocks covering ``[0, n)`` (same partition as ``make_blocks``, core.py:107-134)."""
    if block_size <= 0:
        raise ValueError("block_size must be positive")
    starts = range(0, n, block_size)
    return [(a, min(a + block_size, n)) for a in starts]


def pcg64_record(state: dict, first_sim: int, n_sims: int) -> np.ndarray:
    """One record from ``Generator.bit_generator.state`` of a PCG64."""
    if state.get("bit_generator") != "PCG64":
        raise TypeError(f"expected a PCG64 state, got {state.get('bit_generator')!r}")
    rec = np.zeros(1, dtype=STREAM_DTYPE)
    st, inc = int(state["state"]["state"]), int(state["state"]["inc"])
    rec["kind"] = GEN_PCG64
    rec["buf_pos"] = 0
    rec["has_uint32"] = int(state["has_uint32"])
    rec["uinteger"] = int(state["uinteger"])
    rec["s"][0, :4] = [st >> 64, st & _M64, inc >> 64, inc & _M64]
    rec["first_sim"] = first_sim
    rec["n_sims"] = n_sims
    return rec


# ---------------------------------------------------------------------------------------------------------------
# SeedSequence child keys, all children at once.  ``Philox(child)`` takes its key from ``child.generate_state(2,
# uint64)`` (core.py:659); with hundreds or thousands of spawned blocks (n_workers = cpu_count of a wide host) the
# per-child Python call is the slowest part of a small run.  The children of one ``spawn`` call share the parent's
# entropy and spawn-key prefix and differ in the LAST spawn-key word only, so NumPy's hash (bit_generator.pyx:
# SeedSequence.mix_entropy / generate_state) is evaluated on arrays over the children.  Bit-identical to NumPy
# (tests/test_host_cpu.py compares against ``generate_state`` itself); anything unusual falls back to NumPy.
# ---------------------------------------------------------------------------------------------------------------
_INIT_A, _MULT_A, _INIT_B, _MULT_B = 0x43B0D7E5, 0x931E8875, 0x8B51F9DD, 0x58F38DED
_MIX_L, _MIX_R, _XSHIFT, _M32 = 0xCA01F9DD, 0x4973F715, 16, 0xFFFFFFFF


def _uint32_words(value) -> list[int] | None:
    """NumPy's ``_coerce_to_uint32_array`` for ints and flat sequences of ints (little-endian words, at least one)."""
    if isinstance(value, (int, np.integer)):
        v = int(value)
        if v < 0:
            return None
        words = [v & _M32]
        v >>= 32
        while v:
            words.append(v & _M32)
            v >>= 32
        return words
    if isinstance(value, (tuple, list)):
        out: list[int] = []
        for item in value:
            w = _uint32_words(item) if isinstance(item, (int, np.integer)) else None
            if w is None:
                return None
            out += w
        return out
    return None


def child_philox_keys(children) -> np.ndarray | None:
    """``[c.generate_state(2, uint64) for c in children]`` as a (n, 2) uint64 array, for children of ONE spawn call
    (same entropy, same spawn-key prefix, consecutive last word); ``None`` when the fast form does not apply."""
    n = len(children)
    if n < 8:
        return None
    first = children[0]
    if first.pool_size != 4 or not first.spawn_key or children[-1].spawn_key[:-1] != first.spawn_key[:-1]:
        return None
    k0 = first.spawn_key[-1]
    if not isinstance(k0, (int, np.integer)) or children[-1].spawn_key[-1] != k0 + n - 1 or k0 + n - 1 > _M32:
        return None
    if children[-1].entropy is not first.entropy and children[-1].entropy != first.entropy:
        return None
    run = _uint32_words(first.entropy)
    prefix = _uint32_words(tuple(first.spawn_key[:-1]))
    if run is None or prefix is None:
        return None
    if len(run) < 4:
        run = run + [0] * (4 - len(run))  # spawned sequences pad the run entropy to the pool size
    words = run + prefix  # + the child index, one uint32 word that differs per child
    last = np.arange(k0, k0 + n, dtype=np.uint64).astype(np.uint32)
    hash_const = _INIT_A

    def hashmix(value):  # value: python int (shared by all children) or uint32 array
        nonlocal hash_const
        value = value ^ (hash_const if isinstance(value, int) else np.uint32(hash_const))
        hash_const = (hash_const * _MULT_A) & _M32
        if isinstance(value, int):
            value = (value * hash_const) & _M32
            return value ^ (value >> _XSHIFT)
        value = value * np.uint32(hash_const)
        return value ^ (value >> np.uint32(_XSHIFT))

    def mix(x, y):
        if isinstance(x, int) and isinstance(y, int):
            r = (_MIX_L * x - _MIX_R * y) & _M32
            return r ^ (r >> _XSHIFT)
        x = np.uint32(x) if isinstance(x, int) else x
        y = np.uint32(y) if isinstance(y, int) else y
        r = np.uint32(_MIX_L) * x - np.uint32(_MIX_R) * y
        return r ^ (r >> np.uint32(_XSHIFT))

    entropy = words + [last]
    with np.errstate(over="ignore"):
        mixer = [hashmix(entropy[i]) if i < len(entropy) else hashmix(0) for i in range(4)]
        for i_src in range(4):
            for i_dst in range(4):
                if i_src != i_dst:
                    mixer[i_dst] = mix(mixer[i_dst], hashmix(mixer[i_src]))
        for i_src in range(4, len(entropy)):
            for i_dst in range(4):
                mixer[i_dst] = mix(mixer[i_dst], hashmix(entropy[i_src]))
        # generate_state(2, uint64): four 32-bit words from the pool, cycled
        hc = _INIT_B
        out32 = np.empty((n, 4), dtype=np.uint32)
        for i_dst in range(4):
            v = mixer[i_dst % 4]
            v = (np.full(n, v, dtype=np.uint32) if isinstance(v, int) else v) ^ np.uint32(hc)
            hc = (hc * _MULT_B) & _M32
            v = v * np.uint32(hc)
            out32[:, i_dst] = v ^ (v >> np.uint32(_XSHIFT))
    return out32.view(np.uint64) if out32.dtype.byteorder in ("=", "<", "|") and np.little_endian else None


def philox_records(children, blocks) -> np.ndarray:
    """Records for ``Philox(child)`` of every spawned child (core.py:659): key from
    ``child.generate_state(2, uint64)``, counter 0, empty output buffer."""
    rec = np.zeros(len(children), dtype=STREAM_DTYPE)
    rec["kind"] = GEN_PHILOX
    rec["buf_pos"] = 4
    keys = child_philox_keys(children)
    if keys is None:
        keys = np.array([child.generate_state(2, np.uint64) for child in children], dtype=np.uint64).reshape(len(children), 2)
    rec["s"][:, :2] = keys
    edges = np.asarray(blocks, dtype=np.int64).reshape(len(children), 2)
    rec["first_sim"] = edges[:, 0]
    rec["n_sims"] = edges[:, 1] - edges[:, 0]
    return rec


@dataclass
class StreamLayout:
    """Which reference streams a run consumes and which replications each one feeds."""

    kind: int
    records: np.ndarray  # STREAM_DTYPE[n_streams]
    n_total: int
    hint: int  # uniform block size (0 when there is a single stream)

    @property
    def n_streams(self) -> int:
        return int(self.records.size)

    @property
    def max_sims(self) -> int:
        return int(self.records["n_sims"].max())

    def shard(self, rank: int, world: int) -> "StreamLayout":
        """Contiguous range of stream records for one GPU; results concatenate in reference order."""
        k = self.n_streams
        lo, hi = rank * k // world, (rank + 1) * k // world
        part = self.records[lo:hi].copy()
        if part.size:
            part["first_sim"] -= part["first_sim"][0]
        return StreamLayout(self.kind, part, int(part["n_sims"].sum()), self.hint)


def sequential_layout(rng: np.random.Generator, n: int) -> StreamLayout:
    """``_run_sequential`` (core.py:583-597): every replication draws from ``self.rng`` in turn."""
    return StreamLayout(GEN_PCG64, pcg64_record(rng.bit_generator.state, 0, n), n, 0)


def parallel_layout(seed_seq, n: int, n_workers: int, chunks_per_worker: int = 8) -> StreamLayout:
    """``_prepare_parallel_blocks`` (core.py:622-641): ``max(1, n // (8*W))``-sized blocks, one spawned
    child per block (a fresh OS-entropy SeedSequence per block when the simulation is unseeded)."""
    blocks = block_edges(n, max(1, n // (n_workers * chunks_per_worker)))
    if seed_seq is not None:
        children = seed_seq.spawn(len(blocks))
    else:
        children = [np.random.SeedSequence() for _ in blocks]
    return StreamLayout(GEN_PHILOX, philox_records(children, blocks), n, blocks[0][1] - blocks[0][0])


def generator_record(gen: np.random.Generator, first_sim: int = 0, n_sims: int = 1) -> tuple[int, np.ndarray]:
    """``mcf_stream`` record of a Generator's CURRENT state (PCG64 or Philox bit generators)."""
    st = gen.bit_generator.state
    name = st.get("bit_generator")
    if name == "PCG64":
        return GEN_PCG64, pcg64_record(st, first_sim, n_sims)
    if name == "Philox":
        rec = np.zeros(1, dtype=STREAM_DTYPE)
        rec["kind"] = GEN_PHILOX
        # the device cursor addresses block (counter + (buffer_pos + p) // 4), lane (buffer_pos + p) % 4
        rec["buf_pos"] = int(st["buffer_pos"])
        rec["has_uint32"] = int(st["has_uint32"])
        rec["uinteger"] = int(st["uinteger"])
        rec["s"][0, :2] = st["state"]["key"]
        rec["s"][0, 2:6] = st["state"]["counter"]
        rec["first_sim"] = first_sim
        rec["n_sims"] = n_sims
        return GEN_PHILOX, rec
    raise NotImplementedError(f"device streams support PCG64 and Philox bit generators, got {name!r}")


def _restore_buffered_half(bg, before) -> None:
    """``advance``/``random_raw`` drop NumPy's buffered 32-bit half draw (``has_uint32``/``uinteger``); the reference's
    ``random``/``standard_normal`` calls keep it, so it is written back after a jump over 64-bit draws."""
    if int(before.get("has_uint32", 0)):
        st = bg.state
        st["has_uint32"], st["uinteger"] = before["has_uint32"], before["uinteger"]
        bg.state = st


def advance_keeping_buffer(gen: np.random.Generator, draws: int) -> None:
    """PCG64: jump ``draws`` 64-bit outputs ahead; a buffered ``integers`` half survives, as in the reference."""
    bg = gen.bit_generator
    before = bg.state
    bg.advance(int(draws))
    _restore_buffered_half(bg, before)


def consume_draws(gen: np.random.Generator, draws: int) -> int:
    """Advance ``gen`` by ``draws`` 64-bit outputs without producing them; returns the LAST of those outputs
    (NumPy's buffered ``next_uint32`` needs its upper half).  PCG64: O(log n) jump.  Philox: counter arithmetic."""
    if draws <= 0:
        return 0
    bg = gen.bit_generator
    st = bg.state
    if st["bit_generator"] == "PCG64":
        bg.advance(draws - 1)
        last = int(bg.random_raw())
        _restore_buffered_half(bg, st)
        return last
    if st["bit_generator"] == "Philox":
        q = int(st["buffer_pos"]) + draws  # position after consumption, in units of 64-bit outputs
        ctr = sum(int(c) << (64 * i) for i, c in enumerate(st["state"]["counter"]))
        blk = (q - 1) // 4  # block (relative to the counter) holding the last consumed output
        if blk == 0:  # still inside the buffered block
            last = int(st["buffer"][q - 1])
            st["buffer_pos"] = q  # (the state dict written back still carries the buffered half)
            bg.state = st
            return last
        new_ctr = (ctr + blk - 1) % (1 << 256)  # pre-increment: the next generated block is new_ctr + 1
        st["state"]["counter"] = np.array([(new_ctr >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)
        st["buffer_pos"] = 4
        bg.state = st
        last = int(bg.random_raw((q - 1) % 4 + 1)[-1])
        _restore_buffered_half(bg, st)
        return last
    raise NotImplementedError(st["bit_generator"])
