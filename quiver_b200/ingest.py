"""Arrow IPC / Parquet -> HBM ingest for the reference's two on-disk vector formats.

Plumbing only: pyarrow decodes the files, the rows go through the C ABI batch entry points
(`qh_index_insert_batch` / `qh_collection_add_batch` -> `qg_index_upload`, pinned staging, one H2D copy
per batch). No arithmetic happens here.

Formats (read from the reference's writers, not guessed):

* Arrow IPC file written by `ArrowHNSWIndex.Save` (index/arrow_hnsw.go:153-197): schema
  `id: utf8`, `vector: fixed_size_list<float32>[dim]`; any number of record batches; `Load`
  (:201-241) walks every batch and takes `dim` consecutive float32 values per row. The child values
  buffer of a fixed-size list is already the row-major [rows x dim] matrix the index wants, so a
  batch is handed over without a per-row copy.
* `vectors.parquet` written by `writeVectorsToParquet` (pkg/persistence/parquet.go:16-92): columns
  `id` (UTF8), `vector` (LIST of FLOAT), `metadata` (UTF8 JSON object of string -> string), SNAPPY.
  `readVectorsFromParquet` (:96-174) reads 1000 rows at a time, **skips rows with an empty id or an
  empty vector**, and replaces metadata that does not parse by an empty map; the loader then adds
  each row with its metadata (db.go:248-263).
"""
import json
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

BATCH_ROWS = 1000  # parquet.go:122 reads in batches of 1000 rows


def _pa():
    import pyarrow  # noqa: F401  (kept out of module import so the package loads without pyarrow)
    import pyarrow.ipc
    import pyarrow.parquet
    return pyarrow


# ---- Arrow IPC -----------------------------------------------------------------------------------
def iter_arrow_ipc(path: str, dim: Optional[int] = None) -> Iterator[Tuple[List[str], np.ndarray]]:
    """Yield (ids, [rows x dim] float32 matrix) per record batch of an Arrow IPC *file*."""
    pa = _pa()
    with pa.OSFile(path, "rb") as f:
        reader = pa.ipc.open_file(f)
        for i in range(reader.num_record_batches):
            rb = reader.get_batch(i)
            ids = rb.column(0)
            vec = rb.column(1)
            if not pa.types.is_fixed_size_list(vec.type) or not pa.types.is_float32(vec.type.value_type):
                raise ValueError(f"column 1 must be fixed_size_list<float32>[dim], got {vec.type}")
            d = vec.type.list_size
            if dim is not None and d != dim:
                # Load() would slice with the index's dim and silently misalign; fail loudly instead
                raise ValueError(f"vector dimension mismatch: expected {dim}, got {d}")
            n = rb.num_rows
            # the child array may carry an offset when the batch is a slice
            values = vec.flatten() if vec.offset == 0 else vec.flatten()
            mat = values.to_numpy(zero_copy_only=True).reshape(n, d)
            yield [s.as_py() for s in ids], mat


def load_arrow_ipc(path: str, index, dim: Optional[int] = None) -> int:
    """Insert every row of an Arrow IPC file into `index` (a hostapi.HybridIndex or anything with
    InsertBatch / AddBatch). Returns the number of rows inserted."""
    total = 0
    for ids, mat in iter_arrow_ipc(path, dim):
        _insert(index, ids, mat, None)
        total += len(ids)
    return total


def save_arrow_ipc(path: str, ids: Sequence[str], vectors: np.ndarray, batch_rows: int = 0) -> None:
    """Write the layout ArrowHNSWIndex.Save produces (one record batch unless batch_rows is set)."""
    pa = _pa()
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    n, d = vectors.shape
    schema = pa.schema([("id", pa.string()), ("vector", pa.list_(pa.float32(), d))])
    step = batch_rows if batch_rows > 0 else max(n, 1)
    with pa.OSFile(path, "wb") as f:
        with pa.ipc.new_file(f, schema) as w:
            for r0 in range(0, n, step):
                r1 = min(n, r0 + step)
                flat = pa.array(vectors[r0:r1].reshape(-1), type=pa.float32())
                col = pa.FixedSizeListArray.from_arrays(flat, d)
                w.write_batch(pa.record_batch([pa.array(list(ids[r0:r1]), type=pa.string()), col], schema=schema))


# ---- Parquet -------------------------------------------------------------------------------------
def iter_parquet(path: str, dim: Optional[int] = None,
                 batch_rows: int = BATCH_ROWS) -> Iterator[Tuple[List[str], np.ndarray, List[bytes]]]:
    """Yield (ids, matrix, metadata JSON bytes) per batch of a `vectors.parquet`, applying the
    reference reader's row rules (parquet.go:133-166)."""
    pa = _pa()
    pf = pa.parquet.ParquetFile(path)
    for rb in pf.iter_batches(batch_size=batch_rows, columns=["id", "vector", "metadata"]):
        ids_col = rb.column(0).to_pylist()
        vec_col = rb.column(1)
        md_col = rb.column(2).to_pylist()
        offsets = vec_col.offsets.to_numpy()
        values = vec_col.values.to_numpy(zero_copy_only=False).astype(np.float32, copy=False)
        keep_ids: List[str] = []
        keep_md: List[bytes] = []
        rows = []
        for i, id_ in enumerate(ids_col):
            lo, hi = int(offsets[i]), int(offsets[i + 1])
            if not id_ or hi == lo or not vec_col[i].is_valid:
                continue  # "Skip empty IDs" / "Skip empty vectors"
            if dim is not None and hi - lo != dim:
                raise ValueError(f"vector dimension mismatch: expected {dim}, got {hi - lo}")
            rows.append((lo, hi))
            keep_ids.append(id_)
            keep_md.append(_clean_metadata(md_col[i]))
        if not rows:
            continue
        d = rows[0][1] - rows[0][0]
        if any(hi - lo != d for lo, hi in rows):
            raise ValueError("rows of one batch have different vector lengths")
        if len(rows) == len(ids_col) and rows[0][0] == int(offsets[0]):
            mat = values[rows[0][0]:rows[-1][1]].reshape(len(rows), d)  # contiguous: no per-row copy
        else:
            mat = np.stack([values[lo:hi] for lo, hi in rows])
        yield keep_ids, np.ascontiguousarray(mat, dtype=np.float32), keep_md


def _clean_metadata(text) -> bytes:
    """parquet.go:141-150: metadata is a JSON object of string -> string; anything that does not
    unmarshal into that becomes the empty map."""
    if not text:
        return b"{}"
    try:
        obj = json.loads(text)
    except Exception:
        return b"{}"
    if obj is None:
        return b"{}"  # JSON null unmarshals into a nil map without error
    if not isinstance(obj, dict) or any(not isinstance(v, str) for v in obj.values()):
        return b"{}"  # json.Unmarshal into map[string]string fails on non-string values
    return json.dumps(obj, separators=(",", ":"), ensure_ascii=False).encode()


def load_parquet(path: str, collection, dim: Optional[int] = None, batch_rows: int = BATCH_ROWS) -> int:
    """Add every surviving row of a `vectors.parquet` to `collection` (hostapi.Collection) with its
    metadata. Returns the number of rows added."""
    total = 0
    for ids, mat, md in iter_parquet(path, dim, batch_rows):
        _insert(collection, ids, mat, md)
        total += len(ids)
    return total


def save_parquet(path: str, ids: Sequence[str], vectors: Sequence, metadata: Sequence[Optional[dict]],
                 parquet_go_layout: bool = False) -> None:
    """Write the layout writeVectorsToParquet produces (id, vector LIST<FLOAT>, metadata JSON; SNAPPY).

    parquet_go_layout: reproduce what xitongsys/parquet-go derives from the struct tags of
    ParquetVectorRecord (parquet.go:16-20) instead of pyarrow's defaults — every field REQUIRED (Go value
    types, not pointers), the LIST as the three-level `required group vector (LIST) { repeated group list
    { required float element } }`, dictionary encoding on `id` only (PLAIN_DICTIONARY there, PLAIN on
    `metadata`), data page v1, SNAPPY. (A file written by parquet-go itself cannot be produced in this image:
    there is no Go toolchain.)"""
    pa = _pa()
    md = pa.array([json.dumps(m if m is not None else {}) if not isinstance(m, str) else m for m in metadata],
                  type=pa.string())
    if not parquet_go_layout:
        vec = pa.array([None if v is None else [float(np.float32(x)) for x in v] for v in vectors],
                       type=pa.list_(pa.float32()))
        table = pa.table({"id": pa.array(list(ids), type=pa.string()), "vector": vec, "metadata": md})
        pa.parquet.write_table(table, path, compression="snappy")
        return
    ltype = pa.list_(pa.field("element", pa.float32(), nullable=False))
    schema = pa.schema([pa.field("id", pa.string(), nullable=False), pa.field("vector", ltype, nullable=False),
                        pa.field("metadata", pa.string(), nullable=False)])
    vec = pa.array([[] if v is None else [float(np.float32(x)) for x in v] for v in vectors], type=ltype)
    table = pa.Table.from_arrays([pa.array(list(ids), type=pa.string()), vec, md], schema=schema)
    pa.parquet.write_table(table, path, compression="snappy", use_dictionary=["id"], data_page_version="1.0",
                           version="1.0", use_compliant_nested_type=True, write_statistics=False)


def _insert(target, ids, mat, md) -> None:
    if hasattr(target, "AddBatch"):
        target.AddBatch(ids, mat, md)
    elif hasattr(target, "InsertBatchArrays"):
        target.InsertBatchArrays(ids, mat)
    else:
        target.InsertBatch({i: mat[j] for j, i in enumerate(ids)})
