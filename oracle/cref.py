"""ctypes wrapper of oracle/liboracle.so (oracle/quiver_oracle.c). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

COSINE, L2, DOT, SQL2, L1 = 0, 1, 2, 3, 4
ARITH_VECTORTYPES, ARITH_HNSW_F32 = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib_path() -> str:
    return os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, strict IEEE, no FMA contraction)."""
    srcs = sorted(os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c"))
    out = lib_path()
    newest = max(os.path.getmtime(s) for s in srcs + [os.path.join(_HERE, "synth.h")])
    if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", out]
                              + srcs + ["-lm", "-lpthread"])
    return out


def _load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path()):
            build()
        lib = C.CDLL(lib_path())
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        lib.qo_distance.argtypes = [i32, i32, vp, vp, i32]
        lib.qo_distance.restype = C.c_float
        lib.qo_distances.argtypes = [i32, i32, vp, i64, i32, vp, vp]
        lib.qo_exact_search.argtypes = [vp, i64, i32, i32, i32, vp, vp, i64, vp, vp]
        lib.qo_exact_search.restype = i64
        lib.qo_exact_search_batch.argtypes = [vp, i64, i32, i32, i32, vp, vp, i64, i64, i32, vp, vp, vp]
        lib.qo_synth_fill_mt.argtypes = [i32, C.c_uint64, i64, i64, i32, vp, i32]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def distance(metric: int, a, b, arith: int = ARITH_VECTORTYPES) -> np.float32:
    a, b = _f32(a), _f32(b)
    if a.shape != b.shape:
        raise ValueError("vectors must have the same length")  # distances.go:13-15 panics
    return np.float32(_load().qo_distance(metric, arith, _p(a), _p(b), a.size))


def distances(metric: int, corpus, q, arith: int = ARITH_VECTORTYPES) -> np.ndarray:
    corpus, q = _f32(corpus), _f32(q)
    out = np.empty(corpus.shape[0], dtype=np.float32)
    _load().qo_distances(metric, arith, _p(corpus), corpus.shape[0], corpus.shape[1], _p(q), _p(out))
    return out


def exact_search(corpus, q, k: int, metric: int, arith: int = ARITH_VECTORTYPES, live=None):
    """pkg/hybrid/exact.go:92-133. Returns (dist[k'], row[k']) ascending by (distance, row).
    Raises ValueError('k must be positive') like the reference's error (exact.go:104-106)."""
    corpus, q = _f32(corpus), _f32(q)
    n, d = corpus.shape if corpus.ndim == 2 else (0, q.size)
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    kk = max(int(k), 1)
    dist = np.empty(min(kk, max(n, 1)), dtype=np.float32)
    row = np.empty(min(kk, max(n, 1)), dtype=np.int64)
    got = _load().qo_exact_search(_p(corpus), n, d, metric, arith, _p(lv), _p(q), int(k), _p(dist), _p(row))
    if got == -1:
        raise ValueError("k must be positive")
    if got < 0:
        raise MemoryError()
    return dist[:got].copy(), row[:got].copy()


def exact_search_batch(corpus, queries, k: int, metric: int, arith: int = ARITH_VECTORTYPES, live=None,
                       threads: int = 1):
    """hybrid_index.go:703-795: one goroutine per query, each an ExactIndex.Search."""
    corpus, queries = _f32(corpus), _f32(queries)
    n, d = corpus.shape
    nq = queries.shape[0]
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    dist = np.empty((nq, k), dtype=np.float32)
    row = np.empty((nq, k), dtype=np.int64)
    cnt = np.zeros(nq, dtype=np.int64)
    rc = _load().qo_exact_search_batch(_p(corpus), n, d, metric, arith, _p(lv), _p(queries), nq, k, threads,
                                       _p(dist), _p(row), _p(cnt))
    if rc != 0:
        raise ValueError("k must be positive")
    return dist, row, cnt


def synth(kind: int, seed: int, row0: int, n: int, dim: int, threads: int = 8) -> np.ndarray:
    out = np.empty((n, dim), dtype=np.float32)
    _load().qo_synth_fill_mt(kind, seed, row0, n, dim, _p(out), threads)
    return out
