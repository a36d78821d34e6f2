"""ctypes front-end of ``oracle/mcf_oracle.c`` (CPU oracle; test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmcf_oracle.so")

BS_EURO_CALL, BS_EURO_PUT, BS_AMER_CALL, BS_AMER_PUT, PATH_TERMINAL, PORTFOLIO_GBM, PORTFOLIO_LOG1P = range(7)


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (``make -C oracle``)."""
    src = os.path.join(_HERE, "mcf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mcfo_raw_consumed.restype = C.c_uint64
        _lib.mcfo_np_sum.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def seedseq_state(entropy: int, spawn_key=(), n_words64: int = 4) -> np.ndarray:
    """``SeedSequence(entropy, spawn_key=spawn_key).generate_state(n_words64, np.uint64)``."""
    words = []
    e = int(entropy)
    while True:
        words.append(e & 0xFFFFFFFF)
        e >>= 32
        if e == 0:
            break
    ent = np.array(words, dtype=np.uint32)
    sk = np.array(list(spawn_key), dtype=np.uint32)
    out = np.zeros(2 * n_words64, dtype=np.uint32)
    lib().mcfo_seedseq_state(_p(ent, C.c_uint32), len(ent), _p(sk, C.c_uint32), len(sk), len(out), _p(out, C.c_uint32))
    return out.view(np.uint64)


class Gen:
    """One oracle bit generator (PCG64 or Philox) with NumPy-identical stream semantics."""

    def __init__(self):
        self._buf = C.create_string_buffer(lib().mcfo_gen_sizeof())

    @classmethod
    def pcg64(cls, entropy: int, spawn_key=()):
        g = cls()
        w = seedseq_state(entropy, spawn_key, 4)
        lib().mcfo_init_pcg64(g._buf, _p(w, C.c_uint64))
        return g

    @classmethod
    def pcg64_from_state(cls, state: dict):
        g = cls()
        s, inc = int(state["state"]["state"]), int(state["state"]["inc"])
        m = (1 << 64) - 1
        lib().mcfo_init_pcg64_raw(g._buf, C.c_uint64(s >> 64), C.c_uint64(s & m), C.c_uint64(inc >> 64),
                                  C.c_uint64(inc & m), int(state["has_uint32"]), C.c_uint32(int(state["uinteger"])))
        return g

    @classmethod
    def philox(cls, entropy: int, spawn_key=()):
        g = cls()
        k = seedseq_state(entropy, spawn_key, 2)
        lib().mcfo_init_philox(g._buf, _p(k, C.c_uint64))
        return g

    @classmethod
    def philox_from_key(cls, key):
        g = cls()
        k = np.asarray(key, dtype=np.uint64)
        lib().mcfo_init_philox(g._buf, _p(k, C.c_uint64))
        return g

    def advance(self, delta: int):
        lib().mcfo_pcg64_advance(self._buf, C.c_uint64(delta >> 64), C.c_uint64(delta & ((1 << 64) - 1)))
        return self

    @property
    def consumed(self) -> int:
        return int(lib().mcfo_raw_consumed(self._buf))

    def raw64(self, n):
        out = np.empty(n, dtype=np.uint64)
        lib().mcfo_raw64(self._buf, C.c_int64(n), _p(out, C.c_uint64))
        return out

    def raw32(self, n):
        out = np.empty(n, dtype=np.uint32)
        lib().mcfo_raw32(self._buf, C.c_int64(n), _p(out, C.c_uint32))
        return out

    def random(self, n):
        out = np.empty(n)
        lib().mcfo_doubles(self._buf, C.c_int64(n), _p(out, C.c_double))
        return out

    def uniform(self, low, high, n):
        out = np.empty(n)
        lib().mcfo_uniform(self._buf, C.c_double(low), C.c_double(high), C.c_int64(n), _p(out, C.c_double))
        return out

    def standard_normal(self, n):
        out = np.empty(n)
        lib().mcfo_normals(self._buf, C.c_int64(n), _p(out, C.c_double))
        return out

    def integers(self, low, high, n):
        out = np.empty(n, dtype=np.int64)
        lib().mcfo_integers32(self._buf, C.c_int64(low), C.c_int64(high), C.c_int64(n), _p(out, C.c_int64))
        return out

    # ---- simulations: n_sims consecutive single_simulation() calls on this stream ----
    def pi(self, n_sims, n_points, antithetic=False):
        out = np.empty(n_sims)
        hits = np.empty(n_sims, dtype=np.int64)
        lib().mcfo_pi(self._buf, C.c_int64(n_sims), C.c_int64(n_points), int(bool(antithetic)), _p(out, C.c_double),
                      _p(hits, C.c_int64))
        return out, hits

    def gbm_terminal(self, n_sims, n_steps, mode, params):
        out = np.empty(n_sims)
        p = np.asarray(params, dtype=np.float64)
        lib().mcfo_gbm_terminal(self._buf, C.c_int64(n_sims), C.c_int64(n_steps), int(mode), _p(p, C.c_double),
                                _p(out, C.c_double))
        return out

    def gbm_paths(self, n_paths, n_steps, S0, r, sigma, T):
        out = np.empty((n_paths, n_steps + 1))
        lib().mcfo_gbm_paths(self._buf, C.c_int64(n_paths), C.c_int64(n_steps), C.c_double(S0), C.c_double(r),
                             C.c_double(sigma), C.c_double(T), _p(out, C.c_double))
        return out

    def bootstrap_means(self, x, B, keep_idx=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        means = np.empty(B)
        idx = np.empty(max(keep_idx, 1), dtype=np.int64)
        lib().mcfo_bootstrap_means(self._buf, _p(x, C.c_double), C.c_int64(x.size), C.c_int64(B), _p(means, C.c_double),
                                   _p(idx, C.c_int64), C.c_int64(keep_idx))
        return (means, idx[:keep_idx]) if keep_idx else means

    def dice_sums(self, n_sims, n_dice=2, low=1, high=7):
        out = np.empty(n_sims)
        lib().mcfo_dice_sums(self._buf, C.c_int64(n_sims), C.c_int64(n_dice), C.c_int64(low), C.c_int64(high),
                             _p(out, C.c_double))
        return out


def np_sum(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().mcfo_np_sum(_p(a, C.c_double), C.c_int64(a.size)))
