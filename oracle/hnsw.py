"""ctypes wrapper of oracle/hnsw_oracle.c (restatement of pkg/hnsw/hnsw.go). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import cref


class Graph:
    def __init__(self, corpus: np.ndarray, metric: int, arith: int = 0, M: int = 16, MaxM0: int = 32,
                 EfConstruction: int = 200, EfSearch: int = 100, MaxLevel: int = 16, seed: int = 1, flat: dict = None):
        """Builds the graph by inserting the rows of `corpus` one by one (hnsw.Insert), or — `flat` given —
        adopts a graph stored as the flat arrays export() returns."""
        lib = cref._load()
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        lib.qo_hnsw_import.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
        lib.qo_hnsw_import.restype = vp
        lib.qo_hnsw_search_batch_mt.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp]
        lib.qo_hnsw_build.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, i32, i32, C.c_uint64]
        lib.qo_hnsw_build.restype = vp
        lib.qo_hnsw_free.argtypes = [vp]
        lib.qo_hnsw_search.argtypes = [vp, vp, i32, vp, vp, C.POINTER(i64), C.POINTER(C.c_uint64)]
        lib.qo_hnsw_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
        lib.qo_hnsw_export.argtypes = [vp, vp, vp, vp, vp]
        self._lib = lib
        self.corpus = np.ascontiguousarray(corpus, dtype=np.float32)  # the graph borrows this buffer
        self.metric, self.arith, self.M, self.MaxM0, self.EfSearch = metric, arith, M, MaxM0, EfSearch
        n, d = self.corpus.shape
        if flat is not None:
            self._flat = {k: np.ascontiguousarray(flat[k]) for k in ("level", "adj0", "upper_off", "upper_adj")}
            f = self._flat
            self.h = lib.qo_hnsw_import(self.corpus.ctypes.data_as(vp), n, d, metric, arith, M, MaxM0, EfSearch,
                                        int(flat["entry"]), int(flat["current_level"]),
                                        f["level"].astype(np.int32).ctypes.data_as(vp),
                                        f["adj0"].astype(np.uint32).ctypes.data_as(vp),
                                        f["upper_off"].astype(np.int64).ctypes.data_as(vp),
                                        f["upper_adj"].astype(np.uint32).ctypes.data_as(vp))
        else:
            self.h = lib.qo_hnsw_build(self.corpus.ctypes.data_as(vp), n, d, metric, arith, M, MaxM0, EfConstruction,
                                       EfSearch, MaxLevel, seed)

    def __del__(self):
        try:
            self._lib.qo_hnsw_free(self.h)
        except Exception:
            pass

    def search(self, q, k: int):
        """HNSW.Search (hnsw.go:602-713) -> (dist, node ids, distance evaluations, visit-order hash)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        dist = np.empty(max(k, 1), dtype=np.float32)
        idx = np.empty(max(k, 1), dtype=np.uint32)
        ev, tr = C.c_int64(0), C.c_uint64(0)
        m = self._lib.qo_hnsw_search(self.h, q.ctypes.data_as(C.c_void_p), k, dist.ctypes.data_as(C.c_void_p),
                                     idx.ctypes.data_as(C.c_void_p), C.byref(ev), C.byref(tr))
        if m < 0:
            raise ValueError("k must be positive")
        return dist[:m].copy(), idx[:m].copy(), ev.value, tr.value

    def search_batch(self, queries, k: int, threads: int = 1):
        """hnsw.Search for every query on `threads` host threads -> (dist [nq,k], idx [nq,k], count, evals)."""
        qs = np.ascontiguousarray(queries, dtype=np.float32)
        nq = qs.shape[0]
        dist = np.full((nq, k), np.inf, dtype=np.float32)
        idx = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        cnt = np.zeros(nq, dtype=np.int32)
        ev = np.zeros(nq, dtype=np.int64)
        vp = C.c_void_p
        self._lib.qo_hnsw_search_batch_mt(self.h, qs.ctypes.data_as(vp), nq, k, threads, dist.ctypes.data_as(vp),
                                          idx.ctypes.data_as(vp), cnt.ctypes.data_as(vp), ev.ctypes.data_as(vp))
        return dist, idx, cnt, ev

    def export(self):
        """Flat arrays for the GPU-batched walk (see qh_hnsw_graph in include/quiver_host.h)."""
        n, entry, cur, ulen = C.c_int64(0), C.c_int(0), C.c_int(0), C.c_int64(0)
        self._lib.qo_hnsw_info(self.h, C.byref(n), C.byref(entry), C.byref(cur), C.byref(ulen))
        level = np.empty(n.value, dtype=np.int32)
        adj0 = np.empty((n.value, self.MaxM0), dtype=np.uint32)
        upper_off = np.empty(n.value + 1, dtype=np.int64)
        upper_adj = np.empty(max(ulen.value, 1), dtype=np.uint32)
        vp = C.c_void_p
        self._lib.qo_hnsw_export(self.h, level.ctypes.data_as(vp), adj0.ctypes.data_as(vp),
                                 upper_off.ctypes.data_as(vp), upper_adj.ctypes.data_as(vp))
        return {"n": n.value, "entry": entry.value, "current_level": cur.value, "level": level, "adj0": adj0,
                "upper_off": upper_off, "upper_adj": upper_adj, "M": self.M, "MaxM0": self.MaxM0,
                "EfSearch": self.EfSearch}
