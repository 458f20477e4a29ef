"""ctypes binding of libquivergpu.so (include/quiver_gpu.h).

This is test / benchmark plumbing: it only marshals numpy arrays and raw device pointers
into the C ABI. There is no Python implementation of any operation and no fallback — if the
CUDA library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libquivergpu.so")

COSINE, L2, DOT, SQL2, L1 = 0, 1, 2, 3, 4
ARITH_VECTORTYPES, ARITH_HNSW_F32 = 0, 1
FLAG_NO_BF16_COPY = 1  # qg_config.flags: no bf16 copy of the corpus (quiver_gpu.h)
METRIC_NAMES = {"cosine": COSINE, "euclidean": L2, "l2": L2, "dot_product": DOT, "dot": DOT,
                "squared_euclidean": SQL2, "manhattan": L1}

QG_OK, QG_ERR_INVALID, QG_ERR_DIM, QG_ERR_K, QG_ERR_CUDA, QG_ERR_OOM, QG_ERR_UNSUPPORTED, QG_ERR_RANGE = range(8)


class QuiverGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class qg_config(C.Structure):
    _fields_ = [("device", C.c_int), ("arith", C.c_int), ("reserve_rows", C.c_int64),
                ("select_margin", C.c_int), ("flags", C.c_int)]


class qg_clause(C.Structure):
    _fields_ = [("op", C.c_int32), ("field", C.c_int32), ("negate", C.c_int32), ("ia", C.c_int32),
                ("ib", C.c_int32), ("ic", C.c_int32), ("fa", C.c_double), ("fb", C.c_double)]


class qg_pred(C.Structure):
    _fields_ = [("first_clause", C.c_int32), ("n_clauses", C.c_int32), ("negate", C.c_int32),
                ("require_row", C.c_int32)]


class qg_scan_stats(C.Structure):
    _fields_ = [("rows_scanned", C.c_int64), ("bytes_algorithmic", C.c_int64), ("kernel_launches", C.c_int32),
                ("queries_per_pass", C.c_int32), ("passes", C.c_int32), ("escalations", C.c_int32),
                ("path", C.c_int32), ("reserved", C.c_int32)]


class qg_profile(C.Structure):
    _fields_ = [("scan_ms", C.c_double), ("finalize_ms", C.c_double), ("scan_launches", C.c_int64),
                ("finalize_launches", C.c_int64), ("prep_ms", C.c_double), ("prep_launches", C.c_int64)]


# every symbol include/quiver_gpu.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "qg_abi_version", "qg_last_error", "qg_device_count", "qg_device_info", "qg_index_create",
    "qg_index_destroy", "qg_index_upload", "qg_index_upload_device", "qg_index_upload_synthetic",
    "qg_index_tombstone", "qg_index_compact", "qg_index_size", "qg_index_rows", "qg_index_dim", "qg_index_metric",
    "qg_index_fetch", "qg_facets_set_column", "qg_facets_set_array_column", "qg_filter_compile", "qg_filter_eval", "qg_filter_destroy",
    "qg_search_batch", "qg_search_exhaustive", "qg_search_batch_device", "qg_search_shard_keys_device", "qg_merge_shard_keys_device",
    "qg_batch_distance", "qg_batch_distance_multi", "qg_last_scan_stats", "qg_index_set_profiling",
    "qg_index_read_profile", "qg_debug_tc_pass", "qg_queries_upload", "qg_queries_destroy",
    "qg_batch_distance_queries",
    "qg_hnsw_upload", "qg_hnsw_destroy", "qg_hnsw_search_batch", "qg_hnsw_build", "qg_hnsw_nodes",
    "qg_hnsw_upper_len", "qg_hnsw_export",
    "qg_comm_unique_id", "qg_comm_create_rank", "qg_comm_destroy", "qg_comm_world", "qg_comm_rank",
    "qg_comm_search_rows_device", "qg_comm_search_queries_device",
    "qg_group_create", "qg_group_destroy", "qg_group_devices", "qg_group_index", "qg_group_row_base",
    "qg_group_rows", "qg_group_layout", "qg_group_upload", "qg_group_upload_synthetic", "qg_group_search_batch",
]

_lib = None


def load() -> C.CDLL:
    """Load libquivergpu.so (built in-tree by `make lib` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    # QG_LIB: development aid (an instrumented build of the same library, see tools/tc_timing.py)
    path = os.environ.get("QG_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `make -j8 lib` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_float)
    lib.qg_abi_version.restype = i32
    lib.qg_last_error.restype = C.c_char_p
    lib.qg_device_count.argtypes = [C.POINTER(i32)]
    lib.qg_device_info.argtypes = [i32, C.c_char_p, C.c_size_t, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.qg_index_create.argtypes = [C.POINTER(vp), i32, i32, C.POINTER(qg_config)]
    lib.qg_index_destroy.argtypes = [vp]
    lib.qg_index_upload.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.qg_index_upload_device.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.qg_index_upload_synthetic.argtypes = [vp, i32, C.c_uint64, i64, i64, C.POINTER(i64)]
    lib.qg_index_tombstone.argtypes = [vp, vp, i64]
    lib.qg_index_compact.argtypes = [vp, vp, C.POINTER(i64)]
    lib.qg_index_size.argtypes = [vp]
    lib.qg_index_size.restype = i64
    lib.qg_index_rows.argtypes = [vp]
    lib.qg_index_rows.restype = i64
    lib.qg_index_dim.argtypes = [vp]
    lib.qg_index_metric.argtypes = [vp]
    lib.qg_index_fetch.argtypes = [vp, vp, i64, vp]
    lib.qg_facets_set_column.argtypes = [vp, i32, vp, vp, vp, vp, i64]
    lib.qg_facets_set_array_column.argtypes = [vp, i32, vp, vp, i64, i64]
    lib.qg_filter_compile.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32, C.POINTER(vp)]
    lib.qg_filter_eval.argtypes = [vp, vp, vp, C.POINTER(i64)]
    lib.qg_filter_destroy.argtypes = [vp]
    lib.qg_search_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.qg_search_exhaustive.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.qg_search_batch_device.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.qg_search_shard_keys_device.argtypes = [vp, vp, i32, i32, i32, vp, i64, vp, vp]
    lib.qg_merge_shard_keys_device.argtypes = [i32, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.qg_batch_distance.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.qg_batch_distance_multi.argtypes = [vp, vp, i32, i32, vp, i32, vp]
    lib.qg_last_scan_stats.argtypes = [vp, C.POINTER(qg_scan_stats)]
    lib.qg_index_set_profiling.argtypes = [vp, i32]
    lib.qg_index_read_profile.argtypes = [vp, C.POINTER(qg_profile)]
    lib.qg_debug_tc_pass.argtypes = [vp, vp, i32, i32, vp, vp, vp, C.POINTER(i32)]
    lib.qg_queries_upload.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    lib.qg_queries_destroy.argtypes = [vp]
    lib.qg_batch_distance_queries.argtypes = [vp, vp, vp, i32, vp]
    lib.qg_hnsw_upload.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(vp)]
    lib.qg_hnsw_destroy.argtypes = [vp]
    lib.qg_hnsw_build.argtypes = [vp, i32, i32, i32, i32, C.c_uint64, i32, C.POINTER(vp)]
    lib.qg_hnsw_nodes.argtypes = [vp]
    lib.qg_hnsw_nodes.restype = i64
    lib.qg_hnsw_upper_len.argtypes = [vp]
    lib.qg_hnsw_upper_len.restype = i64
    lib.qg_hnsw_export.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)]
    lib.qg_hnsw_search_batch.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.qg_comm_unique_id.argtypes = [vp]
    lib.qg_comm_create_rank.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    lib.qg_comm_destroy.argtypes = [vp]
    lib.qg_comm_world.argtypes = [vp]
    lib.qg_comm_rank.argtypes = [vp]
    lib.qg_comm_search_rows_device.argtypes = [vp, vp, vp, i32, i32, i32, vp, i64, vp, vp, vp, vp]
    lib.qg_comm_search_queries_device.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, vp, vp, vp, vp]
    lib.qg_group_create.argtypes = [vp, i32, i32, i32, C.POINTER(qg_config), C.POINTER(vp)]
    lib.qg_group_destroy.argtypes = [vp]
    lib.qg_group_devices.argtypes = [vp]
    lib.qg_group_index.argtypes = [vp, i32]
    lib.qg_group_index.restype = vp
    lib.qg_group_row_base.argtypes = [vp, i32]
    lib.qg_group_row_base.restype = i64
    lib.qg_group_rows.argtypes = [vp]
    lib.qg_group_rows.restype = i64
    lib.qg_group_layout.argtypes = [vp]
    lib.qg_group_upload.argtypes = [vp, vp, i64, i32]
    lib.qg_group_upload_synthetic.argtypes = [vp, i32, C.c_uint64, i64, i32]
    lib.qg_group_search_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        msg = load().qg_last_error()
        raise QuiverGpuError(rc, (msg or b"").decode("utf-8", "replace"))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().qg_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def device_info(device: int = 0):
    name = C.create_string_buffer(256)
    sm, ma, mi = C.c_int(0), C.c_int(0), C.c_int(0)
    _check(load().qg_device_info(device, name, 256, C.byref(sm), C.byref(ma), C.byref(mi)))
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (ma.value, mi.value)}


class Filter:
    """A compiled predicate program bound to one index (qg_filter)."""

    def __init__(self, index: "Index", preds: Sequence[qg_pred], clauses: Sequence[qg_clause],
                 iset: Sequence[int] = (), fset: Sequence[float] = ()):
        self._lib = load()
        self.index = index
        pa = (qg_pred * max(1, len(preds)))(*preds)
        ca = (qg_clause * max(1, len(clauses)))(*clauses)
        ia = np.asarray(list(iset), dtype=np.int32)
        fa = np.asarray(list(fset), dtype=np.float64)
        h = C.c_void_p()
        _check(self._lib.qg_filter_compile(index.handle, C.cast(pa, C.c_void_p), len(preds), C.cast(ca, C.c_void_p),
                                           len(clauses), _ptr(ia) if len(ia) else None, len(ia),
                                           _ptr(fa) if len(fa) else None, len(fa), C.byref(h)))
        self.handle = h

    def eval(self):
        """Returns (mask bits as a bool array over all uploaded rows, match count)."""
        n = self.index.rows
        words = np.zeros((n + 63) // 64, dtype=np.uint64)
        m = C.c_int64(0)
        _check(self._lib.qg_filter_eval(self.index.handle, self.handle, _ptr(words), C.byref(m)))
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        return bits, m.value

    def close(self):
        if self.handle:
            self._lib.qg_filter_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    """Thin handle wrapper: rows are int64 indices; see quiver_b200.hybrid for the string-ID mirror."""

    def __init__(self, dim: int, metric: int, device: int = 0, arith: int = ARITH_VECTORTYPES,
                 reserve_rows: int = 0, select_margin: int = 0, flags: int = 0):
        self._lib = load()
        cfg = qg_config(device, arith, reserve_rows, select_margin, flags)
        h = C.c_void_p()
        _check(self._lib.qg_index_create(C.byref(h), dim, metric, C.byref(cfg)))
        self.handle = h
        self.dim, self.metric, self.device, self.arith = dim, metric, device, arith

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None):
            self._lib.qg_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self) -> int:
        return int(self._lib.qg_index_size(self.handle))

    @property
    def rows(self) -> int:
        return int(self._lib.qg_index_rows(self.handle))

    def upload(self, rows: np.ndarray) -> int:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise QuiverGpuError(QG_ERR_DIM, f"vector dimension mismatch: expected {self.dim}, got {rows.shape[-1]}")
        first = C.c_int64(0)
        _check(self._lib.qg_index_upload(self.handle, _ptr(rows), rows.shape[0], C.byref(first)))
        return first.value

    def upload_device(self, dptr: int, n: int) -> int:
        first = C.c_int64(0)
        _check(self._lib.qg_index_upload_device(self.handle, C.c_void_p(dptr), n, C.byref(first)))
        return first.value

    def upload_synthetic(self, kind: int, seed: int, global_row0: int, n: int) -> int:
        first = C.c_int64(0)
        _check(self._lib.qg_index_upload_synthetic(self.handle, kind, seed, global_row0, n, C.byref(first)))
        return first.value

    def tombstone(self, rows) -> None:
        r = np.ascontiguousarray(rows, dtype=np.int64)
        _check(self._lib.qg_index_tombstone(self.handle, _ptr(r), r.size))

    def compact(self) -> np.ndarray:
        """Squeeze tombstoned rows out (qg_index_compact); returns old row -> new row (-1 = deleted)."""
        old_to_new = np.empty(self.rows, dtype=np.int64)
        n_new = C.c_int64(0)
        _check(self._lib.qg_index_compact(self.handle, _ptr(old_to_new), C.byref(n_new)))
        assert n_new.value == self.rows
        return old_to_new

    def fetch(self, rows) -> np.ndarray:
        r = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty((r.size, self.dim), dtype=np.float32)
        _check(self._lib.qg_index_fetch(self.handle, _ptr(r), r.size, _ptr(out)))
        return out

    # -- facets --------------------------------------------------------------------------------
    def set_column(self, field: int, kind: np.ndarray, num: np.ndarray, scode: np.ndarray, fcode: np.ndarray):
        kind = np.ascontiguousarray(kind, dtype=np.uint8)
        num = np.ascontiguousarray(num, dtype=np.float64)
        scode = np.ascontiguousarray(scode, dtype=np.int32)
        fcode = np.ascontiguousarray(fcode, dtype=np.int32)
        _check(self._lib.qg_facets_set_column(self.handle, field, _ptr(kind), _ptr(num), _ptr(scode), _ptr(fcode),
                                              kind.size))

    def set_array_column(self, field: int, offsets: np.ndarray, elem_codes: np.ndarray):
        """CSR element lists of the array-valued rows of a column already set (qg_facets_set_array_column)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        elem_codes = np.ascontiguousarray(elem_codes, dtype=np.int32)
        _check(self._lib.qg_facets_set_array_column(self.handle, field, _ptr(offsets),
                                                    _ptr(elem_codes) if elem_codes.size else None,
                                                    offsets.size - 1, elem_codes.size))

    # -- search --------------------------------------------------------------------------------
    def search(self, queries: np.ndarray, k: int, filter: Optional[Filter] = None,
               negatives: Optional[np.ndarray] = None, out=None):
        """Host-buffer search. Returns (dist [q,k], row [q,k], count [q], negdist or None).
        `out` = (dist, row, count[, negdist]) from an earlier call of the same shape reuses the result
        buffers (the library overwrites every element), as a caller holding result slices would."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        q, dim = queries.shape
        kk = max(k, 0)
        if out is not None and out[0].shape == (q, kk) and out[0].dtype == np.float32 and out[1].dtype == np.int64:
            dist, row, cnt = out[0], out[1], out[2]
        else:
            out = None
            dist = np.full((q, kk), np.inf, dtype=np.float32)
            row = np.full((q, kk), -1, dtype=np.int64)
            cnt = np.zeros(q, dtype=np.int32)
        neg = negd = None
        if negatives is not None:
            neg = np.ascontiguousarray(negatives, dtype=np.float32).reshape(q, -1)
            negd = out[3] if (out is not None and len(out) > 3 and out[3] is not None) else \
                np.full((q, kk), np.inf, dtype=np.float32)
        _check(self._lib.qg_search_batch(self.handle, _ptr(queries), q, dim, k,
                                         filter.handle if filter is not None else None, _ptr(neg), _ptr(dist),
                                         _ptr(negd), _ptr(row), _ptr(cnt)))
        return dist, row, cnt, negd

    def search_exhaustive(self, queries: np.ndarray, k: int, filter: Optional[Filter] = None):
        """The reference algorithm on the device (every row's exact distance + full sort): the GPU-side
        oracle of the full-size parity tests. Returns (dist [q,k], row [q,k], count [q])."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        q, dim = queries.shape
        dist = np.full((q, max(k, 0)), np.inf, dtype=np.float32)
        row = np.full((q, max(k, 0)), -1, dtype=np.int64)
        cnt = np.zeros(q, dtype=np.int32)
        _check(self._lib.qg_search_exhaustive(self.handle, _ptr(queries), q, dim, k,
                                              filter.handle if filter is not None else None, _ptr(dist), _ptr(row),
                                              _ptr(cnt)))
        return dist, row, cnt

    def search_device(self, d_queries: int, q: int, k: int, d_dist: int, d_row: int, d_count: int,
                      stream: int = 0, filter: Optional[Filter] = None, d_negatives: int = 0, d_negdist: int = 0):
        _check(self._lib.qg_search_batch_device(self.handle, C.c_void_p(d_queries), q, self.dim, k,
                                                filter.handle if filter is not None else None,
                                                C.c_void_p(d_negatives) if d_negatives else None,
                                                C.c_void_p(d_dist), C.c_void_p(d_negdist) if d_negdist else None,
                                                C.c_void_p(d_row), C.c_void_p(d_count),
                                                C.c_void_p(stream) if stream else None))

    def search_shard_keys_device(self, d_queries: int, q: int, k: int, row_base: int, d_keys: int, stream: int = 0,
                                 filter: Optional[Filter] = None):
        _check(self._lib.qg_search_shard_keys_device(self.handle, C.c_void_p(d_queries), q, self.dim, k,
                                                     filter.handle if filter is not None else None, row_base,
                                                     C.c_void_p(d_keys), C.c_void_p(stream) if stream else None))

    def batch_distance(self, queries: np.ndarray, rows: np.ndarray) -> np.ndarray:
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(queries.shape[0], -1)
        out = np.empty(rows.shape, dtype=np.float32)
        _check(self._lib.qg_batch_distance_multi(self.handle, _ptr(queries), queries.shape[0], queries.shape[1],
                                                 _ptr(rows), rows.shape[1], _ptr(out)))
        return out

    def debug_tc_pass(self, queries: np.ndarray, k: int):
        """One tensor-core pass without the re-rank: (tau [q], count [q], keys [q, cap] uint64)."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        q = queries.shape[0]
        tau = np.zeros(q, dtype=np.float32)
        cnt = np.zeros(q, dtype=np.int32)
        cand = np.zeros((q, 2048), dtype=np.uint64)
        cap = C.c_int(0)
        _check(self._lib.qg_debug_tc_pass(self.handle, _ptr(queries), q, k, _ptr(tau), _ptr(cnt), _ptr(cand),
                                          C.byref(cap)))
        return tau, cnt, cand[:, :cap.value]

    def set_profiling(self, on: bool) -> None:
        _check(self._lib.qg_index_set_profiling(self.handle, 1 if on else 0))

    def read_profile(self) -> dict:
        pr = qg_profile()
        _check(self._lib.qg_index_read_profile(self.handle, C.byref(pr)))
        return {f: getattr(pr, f) for f, _ in pr._fields_}

    def stats(self) -> dict:
        s = qg_scan_stats()
        _check(self._lib.qg_last_scan_stats(self.handle, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}


def merge_shard_keys_device(device: int, d_keys: int, world: int, q: int, k: int, d_dist: int, d_row: int,
                            d_count: int, stream: int = 0) -> None:
    _check(load().qg_merge_shard_keys_device(device, C.c_void_p(d_keys), world, q, k, C.c_void_p(d_dist),
                                             C.c_void_p(d_row), C.c_void_p(d_count),
                                             C.c_void_p(stream) if stream else None))


LAYOUT_ROWS, LAYOUT_QUERIES, LAYOUT_AUTO = 0, 1, 2
COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """128-byte id rank 0 creates and hands to the other ranks (qg_comm_unique_id)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(load().qg_comm_unique_id(buf))
    return buf.raw


class Comm:
    """One rank's NCCL communicator inside libquivergpu (qg_comm): the one-process-per-GPU shape."""

    def __init__(self, uid: bytes, world: int, rank: int, device: int):
        self._lib = load()
        h = C.c_void_p()
        _check(self._lib.qg_comm_create_rank(C.c_char_p(uid), world, rank, device, C.byref(h)))
        self.handle, self.world, self.rank, self.device = h, world, rank, device

    def search_rows_device(self, shard: "Index", d_queries: int, q: int, k: int, row_base: int, d_dist: int, d_row: int,
                           d_count: int, stream: int = 0, filter: Optional[Filter] = None) -> None:
        _check(self._lib.qg_comm_search_rows_device(self.handle, shard.handle, C.c_void_p(d_queries), q, shard.dim, k,
                                                    filter.handle if filter is not None else None, row_base,
                                                    C.c_void_p(d_dist), C.c_void_p(d_row), C.c_void_p(d_count),
                                                    C.c_void_p(stream) if stream else None))

    def search_queries_device(self, replica: "Index", d_queries: int, q: int, k: int, d_dist: int, d_row: int,
                              d_count: int, stream: int = 0, gather: bool = True,
                              filter: Optional[Filter] = None) -> None:
        _check(self._lib.qg_comm_search_queries_device(self.handle, replica.handle, C.c_void_p(d_queries), q,
                                                       replica.dim, k, filter.handle if filter is not None else None,
                                                       1 if gather else 0, C.c_void_p(d_dist), C.c_void_p(d_row),
                                                       C.c_void_p(d_count), C.c_void_p(stream) if stream else None))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.qg_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """All GPUs of one box driven by one host process (qg_group): the shape a Go host has."""

    def __init__(self, devices: Sequence[int], dim: int, metric: int, arith: int = ARITH_VECTORTYPES):
        self._lib = load()
        devs = (C.c_int * len(devices))(*devices)
        cfg = qg_config(0, arith, 0, 0, 0)
        h = C.c_void_p()
        _check(self._lib.qg_group_create(C.cast(devs, C.c_void_p), len(devices), dim, metric, C.byref(cfg), C.byref(h)))
        self.handle, self.dim, self.metric, self.n = h, dim, metric, len(devices)

    def upload(self, rows: np.ndarray, layout: int = LAYOUT_AUTO) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        _check(self._lib.qg_group_upload(self.handle, _ptr(rows), rows.shape[0], layout))

    def upload_synthetic(self, kind: int, seed: int, n: int, layout: int = LAYOUT_AUTO) -> None:
        _check(self._lib.qg_group_upload_synthetic(self.handle, kind, seed, n, layout))

    @property
    def layout(self) -> int:
        return int(self._lib.qg_group_layout(self.handle))

    @property
    def rows(self) -> int:
        return int(self._lib.qg_group_rows(self.handle))

    def row_base(self, i: int) -> int:
        return int(self._lib.qg_group_row_base(self.handle, i))

    def search(self, queries: np.ndarray, k: int):
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        q, dim = queries.shape
        kk = max(k, 0)
        dist = np.full((q, kk), np.inf, dtype=np.float32)
        row = np.full((q, kk), -1, dtype=np.int64)
        cnt = np.zeros(q, dtype=np.int32)
        _check(self._lib.qg_group_search_batch(self.handle, _ptr(queries), q, dim, k, _ptr(dist), _ptr(row), _ptr(cnt)))
        return dist, row, cnt

    def close(self):
        if getattr(self, "handle", None):
            self._lib.qg_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HnswGraph:
    """An HNSW graph on the index's device (qg_hnsw): built there (qg_hnsw_build) or uploaded (qg_hnsw_upload)."""

    def __init__(self, index: Index, handle):
        self._lib, self.index, self.handle = load(), index, handle

    @classmethod
    def build(cls, index: Index, M: int = 16, MaxM0: int = 32, EfConstruction: int = 200, MaxLevel: int = 16,
              seed: int = 1, max_batch: int = 0) -> "HnswGraph":
        h = C.c_void_p()
        _check(load().qg_hnsw_build(index.handle, M, MaxM0, EfConstruction, MaxLevel, seed, max_batch, C.byref(h)))
        g = cls(index, h)
        g.M, g.MaxM0 = M, MaxM0
        return g

    @classmethod
    def upload(cls, index: Index, graph: dict) -> "HnswGraph":
        level = np.ascontiguousarray(graph["level"], dtype=np.int32)
        adj0 = np.ascontiguousarray(graph["adj0"], dtype=np.uint32)
        uoff = np.ascontiguousarray(graph["upper_off"], dtype=np.int64)
        uadj = np.ascontiguousarray(graph["upper_adj"], dtype=np.uint32)
        h = C.c_void_p()
        _check(load().qg_hnsw_upload(index.handle, int(graph["n"]), int(graph["M"]), int(graph["MaxM0"]),
                                     int(graph["entry"]), int(graph["current_level"]), _ptr(level), _ptr(adj0),
                                     _ptr(uoff), _ptr(uadj), C.byref(h)))
        g = cls(index, h)
        g.M, g.MaxM0 = int(graph["M"]), int(graph["MaxM0"])
        return g

    def export(self, EfSearch: int = 128) -> dict:
        n = int(self._lib.qg_hnsw_nodes(self.handle))
        ulen = int(self._lib.qg_hnsw_upper_len(self.handle))
        level = np.empty(n, dtype=np.int32)
        adj0 = np.empty((n, self.MaxM0), dtype=np.uint32)
        uoff = np.empty(n + 1, dtype=np.int64)
        uadj = np.empty(max(ulen, 1), dtype=np.uint32)
        entry, cur = C.c_int(0), C.c_int(0)
        _check(self._lib.qg_hnsw_export(self.handle, _ptr(level), _ptr(adj0), _ptr(uoff), _ptr(uadj), C.byref(entry),
                                        C.byref(cur)))
        return {"n": n, "entry": entry.value, "current_level": cur.value, "level": level, "adj0": adj0,
                "upper_off": uoff, "upper_adj": uadj, "M": self.M, "MaxM0": self.MaxM0, "EfSearch": EfSearch}

    def search(self, queries: np.ndarray, k: int, ef_search: int = 128):
        """qg_hnsw_search_batch -> (idx [q,k] uint32, dist [q,k], count [q], evals [q])."""
        qs = np.ascontiguousarray(queries, dtype=np.float32)
        q = qs.shape[0]
        idx = np.full((q, k), 0xFFFFFFFF, dtype=np.uint32)
        dist = np.full((q, k), np.inf, dtype=np.float32)
        cnt = np.zeros(q, dtype=np.int32)
        ev = np.zeros(q, dtype=np.int64)
        _check(self._lib.qg_hnsw_search_batch(self.index.handle, self.handle, _ptr(qs), q, qs.shape[1], k, ef_search,
                                              _ptr(idx), _ptr(dist), _ptr(cnt), _ptr(ev)))
        return idx, dist, cnt, ev

    def close(self):
        if getattr(self, "handle", None):
            self._lib.qg_hnsw_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
