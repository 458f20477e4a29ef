"""ctypes binding of libquiverhost.so (include/quiver_host.h) under the reference's names.

`HybridIndex`, `Collection`, `FluentSearch` and the facet filter constructors mirror
pkg/hybrid/hybrid_index.go, pkg/core/collection.go and pkg/facets/facets.go so that the parity
tests read like the reference's own tests. All logic (ID maps, validation, re-rank, predicate
compiler) is in the C++ library; this file only marshals arguments. No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import time
import json
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libquiverhost.so")
_lib = None

EXPORTED_SYMBOLS = [
    "qh_last_error", "qh_results_queries", "qh_results_count", "qh_results_id", "qh_results_distance",
    "qh_results_free", "qh_index_create", "qh_index_destroy", "qh_index_insert", "qh_index_insert_batch",
    "qh_index_delete", "qh_index_delete_batch", "qh_collection_delete_batch", "qh_collection_update",
    "qh_collection_update_batch",
    "qh_index_compact", "qh_collection_compact", "qh_index_size", "qh_index_search", "qh_index_batch_search", "qh_collection_create",
    "qh_collection_destroy", "qh_collection_add", "qh_collection_add_batch", "qh_collection_delete",
    "qh_collection_count", "qh_collection_set_facet_fields", "qh_collection_search",
    "qh_collection_search_with_facets", "qh_collection_filter_mask", "qh_collection_rows", "qh_collection_row_id",
    "qh_debug_sprint_v", "qh_debug_equal_fold", "qh_hnsw_search_batch", "qh_hnsw_upload", "qh_hnsw_dev_free",
    "qh_hnsw_search_device", "qh_hnsw_search_negative", "qh_results_score", "qh_results_vector",
    "qh_results_metadata", "qh_collection_search_request", "qh_collection_persistence_search",
]


class qh_filter(C.Structure):
    _fields_ = [("field", C.c_char_p), ("op", C.c_char_p), ("value_json", C.c_char_p)]


class qh_facet_filter(C.Structure):
    _fields_ = [("type", C.c_int), ("field", C.c_char_p), ("value_json", C.c_char_p), ("min_json", C.c_char_p),
                ("max_json", C.c_char_p), ("include_min", C.c_int), ("include_max", C.c_int),
                ("should_exist", C.c_int)]


class qh_search_options(C.Structure):
    _fields_ = [("include_vectors", C.c_int), ("include_metadata", C.c_int), ("exact_search", C.c_int),
                ("namespace_id", C.c_char_p)]


class qh_hnsw_graph(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("m", C.c_int), ("max_m0", C.c_int), ("entry_point", C.c_int),
                ("current_level", C.c_int), ("ef_search", C.c_int), ("level", C.c_void_p), ("adj0", C.c_void_p),
                ("upper_off", C.c_void_p), ("upper_adj", C.c_void_p)]


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("QH_LIB") or LIB_PATH  # QH_LIB: development aid (the ThreadSanitizer build)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `make -j8 all`")
    capi.load()  # libquivergpu.so first (resolved through $ORIGIN rpath as well)
    lib = C.CDLL(path)
    vp, i32, i64, cp = C.c_void_p, C.c_int, C.c_int64, C.c_char_p
    lib.qh_last_error.restype = cp
    lib.qh_results_queries.argtypes = [vp]
    lib.qh_results_count.argtypes = [vp, i32]
    lib.qh_results_id.argtypes = [vp, i32, i32]
    lib.qh_results_id.restype = cp
    lib.qh_results_distance.argtypes = [vp, i32, i32]
    lib.qh_results_distance.restype = C.c_float
    lib.qh_results_score.argtypes = [vp, i32, i32]
    lib.qh_results_score.restype = C.c_float
    lib.qh_results_vector.argtypes = [vp, i32, i32, C.POINTER(i32)]
    lib.qh_results_vector.restype = C.POINTER(C.c_float)
    lib.qh_results_metadata.argtypes = [vp, i32, i32]
    lib.qh_results_metadata.restype = cp
    lib.qh_collection_search_request.argtypes = [vp, vp, i32, i32, C.POINTER(qh_filter), i32,
                                                 C.POINTER(qh_search_options), C.POINTER(vp)]
    lib.qh_collection_persistence_search.argtypes = [vp, vp, i32, i32, C.POINTER(qh_facet_filter), i32, C.POINTER(vp)]
    lib.qh_hnsw_search_negative.argtypes = [vp, vp, vp, i32, vp, i32, C.c_float, i32, C.POINTER(vp)]
    lib.qh_results_free.argtypes = [vp]
    lib.qh_index_create.argtypes = [C.POINTER(vp), i32, cp, i32, i32]
    lib.qh_index_destroy.argtypes = [vp]
    lib.qh_index_insert.argtypes = [vp, cp, vp, i32]
    lib.qh_index_insert_batch.argtypes = [vp, C.POINTER(cp), vp, i64, i32]
    lib.qh_index_delete.argtypes = [vp, cp]
    lib.qh_index_delete_batch.argtypes = [vp, C.POINTER(cp), i64]
    lib.qh_collection_delete_batch.argtypes = [vp, C.POINTER(cp), i64]
    lib.qh_collection_update.argtypes = [vp, cp, vp, i32, cp]
    lib.qh_collection_update_batch.argtypes = [vp, C.POINTER(cp), vp, i64, i32, C.POINTER(cp)]
    lib.qh_index_compact.argtypes = [vp, C.POINTER(i64)]
    lib.qh_collection_compact.argtypes = [vp, C.POINTER(i64)]
    lib.qh_index_size.argtypes = [vp]
    lib.qh_index_size.restype = i64
    lib.qh_index_search.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    lib.qh_index_batch_search.argtypes = [vp, vp, i32, i32, i32, vp, i32, C.c_float, cp, C.POINTER(vp)]
    lib.qh_collection_create.argtypes = [C.POINTER(vp), cp, i32, cp, i32]
    lib.qh_collection_destroy.argtypes = [vp]
    lib.qh_collection_add.argtypes = [vp, cp, vp, i32, cp]
    lib.qh_collection_add_batch.argtypes = [vp, C.POINTER(cp), vp, i64, i32, C.POINTER(cp)]
    lib.qh_collection_delete.argtypes = [vp, cp]
    lib.qh_collection_count.argtypes = [vp]
    lib.qh_collection_count.restype = i64
    lib.qh_collection_set_facet_fields.argtypes = [vp, C.POINTER(cp), i32]
    lib.qh_collection_search.argtypes = [vp, vp, i32, i32, C.POINTER(qh_filter), i32, C.POINTER(vp)]
    lib.qh_collection_search_with_facets.argtypes = [vp, vp, i32, i32, C.POINTER(qh_facet_filter), i32, C.POINTER(vp)]
    lib.qh_collection_filter_mask.argtypes = [vp, i32, C.POINTER(qh_filter), C.POINTER(qh_facet_filter), i32, vp, i64]
    lib.qh_collection_rows.argtypes = [vp]
    lib.qh_collection_rows.restype = i64
    lib.qh_collection_row_id.argtypes = [vp, i64]
    lib.qh_collection_row_id.restype = cp
    lib.qh_hnsw_upload.argtypes = [vp, C.POINTER(qh_hnsw_graph), C.POINTER(vp)]
    lib.qh_hnsw_dev_free.argtypes = [vp]
    lib.qh_hnsw_search_device.argtypes = [vp, vp, vp, i32, i32, i32, C.POINTER(vp), vp, C.POINTER(i32)]
    lib.qh_hnsw_search_batch.argtypes = [vp, C.POINTER(qh_hnsw_graph), vp, i32, i32, i32, C.POINTER(vp), vp,
                                         C.POINTER(i64)]
    lib.qh_debug_sprint_v.argtypes = [cp, i32, C.c_char_p, i32]
    lib.qh_debug_equal_fold.argtypes = [cp, cp]
    _lib = lib
    return lib


class QuiverError(RuntimeError):
    """The Go `error` the reference would have returned (message text is the reference's)."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def _check(rc: int) -> None:
    if rc != 0:
        raise QuiverError(rc, (load().qh_last_error() or b"").decode("utf-8", "replace"))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _take(res: C.c_void_p) -> List[List[Tuple[str, np.float32]]]:
    """[[(ID, Distance), ...] per query] — types.BasicSearchResult (pkg/types/search.go:9-14)."""
    lib = load()
    out = []
    for q in range(lib.qh_results_queries(res)):
        out.append([(lib.qh_results_id(res, q, j).decode(), np.float32(lib.qh_results_distance(res, q, j)))
                    for j in range(lib.qh_results_count(res, q))])
    lib.qh_results_free(res)
    return out


def _take_items(res: C.c_void_p) -> List[dict]:
    """types.SearchResultItem per result of query 0 (pkg/types/search.go:31-42): ID, Distance, Score and — when
    the request's options asked for them — Vector and Metadata (absent keys = the reference's omitempty)."""
    lib = load()
    out = []
    for j in range(lib.qh_results_count(res, 0)):
        item = {"ID": lib.qh_results_id(res, 0, j).decode(), "Distance": np.float32(lib.qh_results_distance(res, 0, j)),
                "Score": np.float32(lib.qh_results_score(res, 0, j))}
        n = C.c_int(0)
        vp_ = lib.qh_results_vector(res, 0, j, C.byref(n))
        if n.value > 0:
            item["Vector"] = np.ctypeslib.as_array(vp_, shape=(n.value,)).copy()
        md = lib.qh_results_metadata(res, 0, j)
        if md is not None:
            item["Metadata"] = md.decode()
        out.append(item)
    lib.qh_results_free(res)
    return out


def sprint_v(value_json: str, typed: bool = True) -> str:
    buf = C.create_string_buffer(4096)
    n = load().qh_debug_sprint_v(value_json.encode(), 1 if typed else 0, buf, 4096)
    if n < 0:
        raise ValueError((load().qh_last_error() or b"").decode())
    return buf.value.decode()


def equal_fold(a: str, b: str) -> bool:
    return bool(load().qh_debug_equal_fold(a.encode(), b.encode()))


class HybridIndex:
    """hybrid.HybridIndex with ForceStrategy = exact (pkg/hybrid/hybrid_index.go)."""

    def __init__(self, dim: int, distance: str = "cosine", arith: int = 0, device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        _check(self._lib.qh_index_create(C.byref(h), dim, distance.encode(), arith, device))
        self.handle, self.dim = h, dim

    def close(self):
        if getattr(self, "handle", None):
            self._lib.qh_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Insert(self, id: str, vector) -> None:
        v = _f32(vector)
        _check(self._lib.qh_index_insert(self.handle, id.encode(), _ptr(v), v.size))

    def InsertBatch(self, vectors: dict) -> None:
        ids = list(vectors)
        if not ids:
            return
        mat = _f32(np.stack([_f32(vectors[i]) for i in ids]))
        arr = (C.c_char_p * len(ids))(*[i.encode() for i in ids])
        _check(self._lib.qh_index_insert_batch(self.handle, arr, _ptr(mat), len(ids), mat.shape[1]))

    def InsertBatchArrays(self, ids: Sequence[str], matrix) -> None:
        """InsertBatch for rows that already sit in one [n x dim] matrix (Arrow / Parquet ingest)."""
        mat = _f32(matrix)
        if len(ids) == 0:
            return
        arr = (C.c_char_p * len(ids))(*[i.encode() for i in ids])
        _check(self._lib.qh_index_insert_batch(self.handle, arr, _ptr(mat), len(ids), mat.shape[1]))

    def Delete(self, id: str) -> None:
        _check(self._lib.qh_index_delete(self.handle, id.encode()))

    def DeleteBatch(self, ids: Sequence[str]) -> None:
        arr = (C.c_char_p * max(1, len(ids)))(*[i.encode() for i in ids])
        _check(self._lib.qh_index_delete_batch(self.handle, arr, len(ids)))

    def Compact(self) -> int:
        """Drop the rows of deleted vectors from HBM; returns how many rows went (not in the reference: its map
        frees a vector on Delete, exact.go:61-70)."""
        removed = C.c_int64(0)
        _check(self._lib.qh_index_compact(self.handle, C.byref(removed)))
        return removed.value

    def Size(self) -> int:
        return int(self._lib.qh_index_size(self.handle))

    def Search(self, query, k: int):
        q = _f32(query)
        res = C.c_void_p()
        _check(self._lib.qh_index_search(self.handle, _ptr(q), q.size, k, C.byref(res)))
        return _take(res)[0]

    def SearchWithRequest(self, Query, K: int, ForceStrategy: str = "", NegativeExample=None,
                          NegativeWeight: float = 0.0):
        return self.BatchSearch([Query], K, ForceStrategy,
                                None if NegativeExample is None else [NegativeExample], NegativeWeight)[0]

    def BatchSearch(self, Queries: Sequence, K: int, ForceStrategy: str = "", NegativeExamples=None,
                    NegativeWeight: float = 0.0):
        if len(Queries) == 0:
            _check(self._lib.qh_index_batch_search(self.handle, None, 0, self.dim, K, None, 0, 0.0, b"", C.byref(C.c_void_p())))
        qs = _f32(np.stack([_f32(q) for q in Queries]))
        neg, neg_dim = None, 0
        if NegativeExamples is not None and len(NegativeExamples) > 0:
            if len(NegativeExamples) != len(Queries) and NegativeWeight > 0:
                raise QuiverError(1, "number of negative examples must match number of queries")
            neg = _f32(np.stack([_f32(n) for n in NegativeExamples]))
            neg_dim = neg.shape[1]
        res = C.c_void_p()
        _check(self._lib.qh_index_batch_search(self.handle, _ptr(qs), qs.shape[0], qs.shape[1], K, _ptr(neg), neg_dim,
                                               float(NegativeWeight), ForceStrategy.encode(), C.byref(res)))
        return _take(res)

    def FluentSearch(self, query):
        return FluentHybridSearch(self, query)

    def HNSWSearchBatch(self, graph: dict, queries, k: int):
        """hnsw.Search (pkg/hnsw/hnsw.go:602-713) for a batch of queries over the reference's graph
        (flat arrays, node id = insertion row), neighbour distances batched on the GPU.
        Returns (results per query, distance evaluations per query, lock-step rounds)."""
        qs = _f32(queries)
        level = np.ascontiguousarray(graph["level"], dtype=np.int32)
        adj0 = np.ascontiguousarray(graph["adj0"], dtype=np.uint32)
        uoff = np.ascontiguousarray(graph["upper_off"], dtype=np.int64)
        uadj = np.ascontiguousarray(graph["upper_adj"], dtype=np.uint32)
        g = qh_hnsw_graph(int(graph["n"]), int(graph["M"]), int(graph["MaxM0"]), int(graph["entry"]),
                          int(graph["current_level"]), int(graph["EfSearch"]), level.ctypes.data, adj0.ctypes.data,
                          uoff.ctypes.data, uadj.ctypes.data)
        evals = np.zeros(qs.shape[0], dtype=np.int64)
        steps = C.c_int64(0)
        res = C.c_void_p()
        t0 = time.perf_counter()
        _check(self._lib.qh_hnsw_search_batch(self.handle, C.byref(g), _ptr(qs), qs.shape[0], qs.shape[1], k,
                                              C.byref(res), _ptr(evals), C.byref(steps)))
        self.last_call_s = time.perf_counter() - t0
        return _take(res), evals, steps.value

    def HNSWUpload(self, graph: dict) -> "DeviceGraph":
        """The graph made resident on the device for HNSW searches that run entirely there."""
        return DeviceGraph(self, graph)


class DeviceGraph:
    """An HNSW graph resident on the index's device (qh_hnsw_upload). Keeps the host arrays alive: the
    lock-step fallback of qh_hnsw_search_device reads them."""

    def __init__(self, index: "HybridIndex", graph: dict):
        self._lib = index._lib
        self.index = index
        self._keep = (np.ascontiguousarray(graph["level"], dtype=np.int32),
                      np.ascontiguousarray(graph["adj0"], dtype=np.uint32),
                      np.ascontiguousarray(graph["upper_off"], dtype=np.int64),
                      np.ascontiguousarray(graph["upper_adj"], dtype=np.uint32))
        level, adj0, uoff, uadj = self._keep
        self._g = qh_hnsw_graph(int(graph["n"]), int(graph["M"]), int(graph["MaxM0"]), int(graph["entry"]),
                                int(graph["current_level"]), int(graph["EfSearch"]), level.ctypes.data,
                                adj0.ctypes.data, uoff.ctypes.data, uadj.ctypes.data)
        h = C.c_void_p()
        _check(self._lib.qh_hnsw_upload(index.handle, C.byref(self._g), C.byref(h)))
        self.handle = h

    def search(self, queries, k: int):
        """hnsw.Search for a batch, the whole walk on the device. Returns (results per query, distance
        evaluations per query, queries that fell back to the lock-step host walk)."""
        qs = _f32(queries)
        evals = np.zeros(qs.shape[0], dtype=np.int64)
        fb = C.c_int(0)
        res = C.c_void_p()
        t0 = time.perf_counter()
        _check(self._lib.qh_hnsw_search_device(self.index.handle, self.handle, _ptr(qs), qs.shape[0], qs.shape[1], k,
                                               C.byref(res), _ptr(evals), C.byref(fb)))
        self.last_call_s = time.perf_counter() - t0  # the C call alone (the list conversion below is Python's)
        return _take(res), evals, fb.value

    def search_with_negative_example(self, query, negative, negative_weight: float, k: int):
        """HNSWAdapter.SearchWithNegativeExample (pkg/hnsw/adapter.go:345-437) -> [(ID, adjusted Distance)]."""
        q = _f32(query)
        neg = None if negative is None or len(negative) == 0 else _f32(negative)
        res = C.c_void_p()
        _check(self._lib.qh_hnsw_search_negative(self.index.handle, self.handle, _ptr(q), q.size, _ptr(neg),
                                                 0 if neg is None else neg.size, float(negative_weight), k, C.byref(res)))
        return _take(res)[0]

    def close(self):
        if getattr(self, "handle", None):
            self._lib.qh_hnsw_dev_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FluentHybridSearch:
    """hybrid.FluentHybridSearch (hybrid_index.go:814-881): defaults k=10, negativeWeight=0.5."""

    def __init__(self, index: HybridIndex, query):
        self.index, self.query, self.k, self.strategy, self.neg, self.w = index, query, 10, "", None, 0.5

    def WithK(self, k):
        self.k = k
        return self

    def WithStrategy(self, s):
        self.strategy = s
        return self

    def WithNegativeExample(self, v):
        self.neg = v
        return self

    def WithNegativeWeight(self, w):
        self.w = w
        return self

    def Execute(self):
        return self.index.SearchWithRequest(self.query, self.k, self.strategy, self.neg, self.w if self.neg is not None else 0.0)


# ---- facets.Filter constructors (pkg/facets/facets.go:46,104,274,351) ---------------------------------
def _lit(v) -> bytes:
    """A Go literal as JSON text: Python int -> Go int, float -> float64 (always with '.' or 'e')."""
    if isinstance(v, bool) or v is None or isinstance(v, (str, list, dict)):
        return json.dumps(v).encode()
    if isinstance(v, int):
        return str(v).encode()
    if isinstance(v, float):
        s = repr(v)
        if "." not in s and "e" not in s and "E" not in s and "n" not in s:
            s += ".0"
        return s.encode()
    raise TypeError(type(v))


def _lit_list(vs) -> bytes:
    return b"[" + b",".join(_lit(v) for v in vs) + b"]"


def NewEqualityFilter(field, value):
    return qh_facet_filter(0, field.encode(), _lit(value), None, None, 1, 1, 1)


def NewRangeFilter(field, min, max, includeMin, includeMax):
    return qh_facet_filter(1, field.encode(), None, _lit(min), _lit(max), int(includeMin), int(includeMax), 1)


def NewSetFilter(field, values):
    return qh_facet_filter(2, field.encode(), _lit_list(values), None, None, 1, 1, 1)


def NewExistsFilter(field, shouldExist):
    return qh_facet_filter(3, field.encode(), None, None, None, 1, 1, int(shouldExist))


class Collection:
    """core.Collection (pkg/core/collection.go) over the GPU index."""

    def __init__(self, name: str, dim: int, distance: str = "cosine", device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        _check(self._lib.qh_collection_create(C.byref(h), name.encode(), dim, distance.encode(), device))
        self.handle, self.Dimension, self.Name = h, dim, name

    def close(self):
        if getattr(self, "handle", None):
            self._lib.qh_collection_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Add(self, id: str, vector, metadata=None) -> None:
        v = _f32(vector)
        md = None if metadata is None else (metadata if isinstance(metadata, (bytes, str)) else json.dumps(metadata))
        if isinstance(md, str):
            md = md.encode()
        _check(self._lib.qh_collection_add(self.handle, id.encode(), _ptr(v), v.size, md))

    def AddBatch(self, ids: Sequence[str], vectors, metadata: Optional[Sequence] = None) -> None:
        mat = _f32(vectors)
        n = len(ids)
        arr = (C.c_char_p * n)(*[i.encode() for i in ids])
        mds = None
        if metadata is not None:
            enc = []
            for m in metadata:
                if m is None:
                    enc.append(None)
                elif isinstance(m, bytes):
                    enc.append(m)
                elif isinstance(m, str):
                    enc.append(m.encode())
                else:
                    enc.append(json.dumps(m).encode())
            mds = (C.c_char_p * n)(*enc)
        _check(self._lib.qh_collection_add_batch(self.handle, arr, _ptr(mat), n, mat.shape[1], mds))

    def Delete(self, id: str) -> None:
        _check(self._lib.qh_collection_delete(self.handle, id.encode()))

    def DeleteBatch(self, ids: Sequence[str]) -> None:
        arr = (C.c_char_p * max(1, len(ids)))(*[i.encode() for i in ids])
        _check(self._lib.qh_collection_delete_batch(self.handle, arr, len(ids)))

    def Update(self, id: str, vector=None, metadata=None) -> None:
        v = None if vector is None else _f32(vector)
        md = None if metadata is None else (metadata if isinstance(metadata, (bytes, str)) else json.dumps(metadata))
        if isinstance(md, str):
            md = md.encode()
        _check(self._lib.qh_collection_update(self.handle, id.encode(), _ptr(v), 0 if v is None else v.size, md))

    def UpdateBatch(self, ids: Sequence[str], vectors, metadata: Optional[Sequence] = None) -> None:
        mat = _f32(vectors)
        n = len(ids)
        if n and mat.ndim == 1:
            mat = mat.reshape(n, -1)
        arr = (C.c_char_p * max(1, n))(*[i.encode() for i in ids])
        mds = None
        if metadata is not None:
            enc = [None if m is None else (m if isinstance(m, (bytes, str)) else json.dumps(m)) for m in metadata]
            mds = (C.c_char_p * max(1, n))(*[e.encode() if isinstance(e, str) else e for e in enc])
        _check(self._lib.qh_collection_update_batch(self.handle, arr, _ptr(mat), n, mat.shape[1] if n else 0, mds))

    def Compact(self) -> int:
        removed = C.c_int64(0)
        _check(self._lib.qh_collection_compact(self.handle, C.byref(removed)))
        return removed.value

    def Count(self) -> int:
        return int(self._lib.qh_collection_count(self.handle))

    def SetFacetFields(self, fields: Sequence[str]) -> None:
        arr = (C.c_char_p * len(fields))(*[f.encode() for f in fields])
        _check(self._lib.qh_collection_set_facet_fields(self.handle, arr, len(fields)))

    @staticmethod
    def _filters(filters):
        fs = [qh_filter(f.encode(), op.encode(), _lit(v) if not isinstance(v, bytes) else v) for f, op, v in filters]
        return (qh_filter * max(1, len(fs)))(*fs), len(fs)

    def Search(self, Vector, TopK: int, Filters: Sequence[Tuple[str, str, object]] = ()):
        """types.SearchRequest{Vector, TopK, Filters} -> [(ID, Distance)] (collection.go:637-807)."""
        q = _f32(Vector)
        arr, n = self._filters(Filters)
        res = C.c_void_p()
        _check(self._lib.qh_collection_search(self.handle, _ptr(q), q.size, TopK, arr, n, C.byref(res)))
        return _take(res)[0]

    def SearchRequest(self, Vector, TopK: int, Filters: Sequence[Tuple[str, str, object]] = (),
                      IncludeVectors: bool = False, IncludeMetadata: bool = False, ExactSearch: bool = False,
                      NamespaceID: str = ""):
        """Collection.Search(types.SearchRequest) -> types.SearchResponse.Results (collection.go:637-807)."""
        q = _f32(Vector)
        arr, n = self._filters(Filters)
        opt = qh_search_options(int(IncludeVectors), int(IncludeMetadata), int(ExactSearch), NamespaceID.encode())
        res = C.c_void_p()
        _check(self._lib.qh_collection_search_request(self.handle, _ptr(q), q.size, TopK, arr, n, C.byref(opt),
                                                      C.byref(res)))
        return _take_items(res)

    def PersistenceSearch(self, query, limit: int, filters: Sequence[qh_facet_filter] = ()):
        """persistence.Collection.Search / SearchWithFacets (pkg/persistence/collection.go:226-261, 327-378)."""
        q = None if query is None else _f32(query)
        arr = (qh_facet_filter * max(1, len(filters)))(*filters)
        res = C.c_void_p()
        _check(self._lib.qh_collection_persistence_search(self.handle, _ptr(q), 0 if q is None else q.size, limit, arr,
                                                          len(filters), C.byref(res)))
        return _take(res)[0]

    def SearchWithFacets(self, query, k: int, filters: Sequence[qh_facet_filter] = ()):
        q = _f32(query)
        arr = (qh_facet_filter * max(1, len(filters)))(*filters)
        res = C.c_void_p()
        _check(self._lib.qh_collection_search_with_facets(self.handle, _ptr(q), q.size, k, arr, len(filters), C.byref(res)))
        return _take(res)[0]

    def FluentSearch(self, vector):
        return FluentSearch(self, vector)

    def filter_mask(self, Filters=None, facet_filters=None) -> np.ndarray:
        """Row-pass bits of a predicate set (bit-exactness checks against the oracle)."""
        rows = int(self._lib.qh_collection_rows(self.handle))
        out = np.zeros(max(rows, 1), dtype=np.uint8)
        if facet_filters is not None:
            arr = (qh_facet_filter * max(1, len(facet_filters)))(*facet_filters)
            _check(self._lib.qh_collection_filter_mask(self.handle, 1, None, arr, len(facet_filters), _ptr(out), out.size))
        else:
            arr, n = self._filters(Filters or ())
            _check(self._lib.qh_collection_filter_mask(self.handle, 0, arr, None, n, _ptr(out), out.size))
        return out[:rows].astype(bool)


class FluentSearch:
    """core.FluentSearch (collection.go:874-1108): WithK, WithNamespace, IncludeVectors, IncludeMetadata,
    UseExactSearch, Filter, FilterNotEquals, FilterGreaterThan, FilterLessThan, FilterIn, Execute. Defaults:
    k = 10, IncludeMetadata = true (:887-895); k is clamped to Count() (:924-926). Once a builder call has
    failed the later ones are ignored and Execute returns that first error (the `valid` flag)."""

    def __init__(self, collection: Collection, vector):
        self.c, self.vector, self.k, self.filters, self.err = collection, vector, 10, [], None
        self.include_vectors, self.include_metadata, self.exact, self.namespace = False, True, False, ""
        if len(vector) != collection.Dimension:
            self.err = QuiverError(2, f"invalid vector dimension: expected {collection.Dimension}, got {len(vector)}")

    def WithNamespace(self, namespace: str):
        if self.err is None:
            self.namespace = namespace
        return self

    def IncludeVectors(self, include: bool):
        if self.err is None:
            self.include_vectors = bool(include)
        return self

    def IncludeMetadata(self, include: bool):
        if self.err is None:
            self.include_metadata = bool(include)
        return self

    def UseExactSearch(self):
        if self.err is None:
            self.exact = True
        return self

    def WithK(self, k):
        if k <= 0:
            self.err = self.err or QuiverError(3, "k must be greater than 0")
        self.k = k
        return self

    def Filter(self, field, value):
        self.filters.append((field, "=", value))
        return self

    def FilterNotEquals(self, field, value):
        self.filters.append((field, "!=", value))
        return self

    def FilterGreaterThan(self, field, value):
        self.filters.append((field, ">", value))
        return self

    def FilterLessThan(self, field, value):
        self.filters.append((field, "<", value))
        return self

    def FilterIn(self, field, values):
        if len(values) == 0:
            self.err = self.err or QuiverError(1, "values cannot be empty")
        self.filters.append((field, "in", list(values)))
        return self

    def Execute(self):
        if self.err is not None:
            raise self.err
        if self.k <= 0:
            raise QuiverError(3, "k must be greater than 0")
        k = min(self.k, max(self.c.Count(), 1))
        return self.c.Search(self.vector, k, self.filters)

    def ExecuteResponse(self):
        """Execute() with the full types.SearchResultItem list (Score, and Vector / Metadata per the options)."""
        if self.err is not None:
            raise self.err
        if self.k <= 0:
            raise QuiverError(3, "k must be greater than 0")
        k = min(self.k, max(self.c.Count(), 1))
        return self.c.SearchRequest(self.vector, k, self.filters, self.include_vectors, self.include_metadata,
                                    self.exact, self.namespace)
