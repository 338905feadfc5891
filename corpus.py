"""Deterministic synthetic corpora shared by tests and bench.py.

Thin ctypes wrapper over tools/corpus_gen.c (built on demand with gcc).  Test/bench
infrastructure — not part of the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "tools", "corpus_gen.c")
_LIB = os.path.join(_HERE, "tools", "libcorpus.so")

SEED_C1 = 0xB2000001
SEED_C2 = 0xB2000002
SEED_C3 = 0xB2000003
SEED_C4 = 0xB2000004
SEED_C5 = 0xB2000005

_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-o", _LIB, _SRC, "-lm"])
    return _LIB


def _get():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        for name in ("corpus_random", "corpus_random_bytewise", "corpus_text", "corpus_source",
                     "corpus_binary", "corpus_mixed"):
            getattr(_lib, name).argtypes = [C.c_uint64, C.c_void_p, C.c_size_t]
            getattr(_lib, name).restype = None
    return _lib


def _gen(name, seed, n, out=None):
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n and out.flags["C_CONTIGUOUS"]
    getattr(_get(), name)(seed, out.ctypes.data_as(C.c_void_p), n)
    return out[:n]


def random_bytes(n, seed=SEED_C5, out=None):
    return _gen("corpus_random", seed, n, out)


def random_bytewise(n, seed=1, out=None):
    return _gen("corpus_random_bytewise", seed, n, out)


def text(n, seed=SEED_C1, out=None):
    return _gen("corpus_text", seed, n, out)


def source(n, seed=SEED_C2, out=None):
    return _gen("corpus_source", seed, n, out)


def binary(n, seed=SEED_C2, out=None):
    return _gen("corpus_binary", seed, n, out)


def mixed(n, seed=SEED_C2, out=None):
    return _gen("corpus_mixed", seed, n, out)


def periodic(n, unit):
    unit = np.frombuffer(bytes(unit), dtype=np.uint8)
    reps = -(-n // unit.size)
    return np.tile(unit, reps)[:n].copy()


def by_name(name, n, seed=None, out=None):
    table = {"text": (text, SEED_C1), "mixed": (mixed, SEED_C2), "random": (random_bytes, SEED_C5),
             "source": (source, SEED_C2), "binary": (binary, SEED_C2)}
    fn, default_seed = table[name]
    return fn(n, default_seed if seed is None else seed, out)
