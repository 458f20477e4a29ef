"""Arrow IPC / Parquet ingest: the reference's two on-disk vector formats, written here
with pyarrow in the schema its writers use (index/arrow_hnsw.go:153-197, persistence/parquet.go:16-92),
read back under its readers' row rules (arrow_hnsw.go:201-241, parquet.go:96-174)."""
import json
import os

import numpy as np
import pytest

pa = pytest.importorskip("pyarrow")

from quiver_b200 import ingest  # noqa: E402


def _rows(n, d, seed=0):
    rng = np.random.default_rng(seed)
    return [f"id{i}" for i in range(n)], rng.random((n, d), dtype=np.float32)


def test_arrow_ipc_round_trip_is_zero_copy(tmp_path):
    ids, vec = _rows(2500, 24)
    path = os.path.join(tmp_path, "index.arrow")
    ingest.save_arrow_ipc(path, ids, vec, batch_rows=1024)  # chunkSize 1024 (arrowindex/graph.go)
    got_ids, mats = [], []
    for b_ids, mat in ingest.iter_arrow_ipc(path, dim=24):
        assert mat.dtype == np.float32 and mat.flags["C_CONTIGUOUS"] and not mat.flags["OWNDATA"]
        got_ids += b_ids
        mats.append(mat.copy())
    assert got_ids == ids
    assert np.array_equal(np.concatenate(mats).view(np.uint32), vec.view(np.uint32))
    with pytest.raises(ValueError, match="dimension mismatch"):
        list(ingest.iter_arrow_ipc(path, dim=32))


def test_parquet_reader_rules(tmp_path):
    """parquet.go:133-166: empty id and empty vector rows are skipped, unparsable metadata becomes {}."""
    path = os.path.join(tmp_path, "vectors.parquet")
    ids = ["a", "", "c", "d", "e"]
    vecs = [[1, 2, 3], [4, 5, 6], [], [7, 8, 9], [10, 11, 12]]
    md = [{"category": "x"}, {"category": "y"}, {"category": "z"}, "not json", {"n": 5}]
    ingest.save_parquet(path, ids, vecs, md)
    schema = pa.parquet.read_schema(path)
    assert schema.names == ["id", "vector", "metadata"]
    batches = list(ingest.iter_parquet(path, dim=3))
    assert len(batches) == 1
    got_ids, mat, got_md = batches[0]
    assert got_ids == ["a", "d", "e"]
    assert mat.tolist() == [[1, 2, 3], [7, 8, 9], [10, 11, 12]]
    # "not json" does not parse; {"n": 5} is not a map[string]string: both become the empty map
    assert [json.loads(m) for m in got_md] == [{"category": "x"}, {}, {}]
    with pytest.raises(ValueError, match="dimension mismatch"):
        list(ingest.iter_parquet(path, dim=4))


def test_parquet_batches_of_1000(tmp_path):
    ids, vec = _rows(2300, 8, seed=3)
    path = os.path.join(tmp_path, "vectors.parquet")
    ingest.save_parquet(path, ids, vec, [{"k": str(i % 7)} for i in range(len(ids))])
    sizes, all_ids = [], []
    for b_ids, mat, md in ingest.iter_parquet(path):
        sizes.append(len(b_ids))
        all_ids += b_ids
        assert mat.shape == (len(b_ids), 8)
    assert all_ids == ids and max(sizes) <= ingest.BATCH_ROWS and sum(sizes) == 2300


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture():
    rows = json.load(open(os.path.join(GOLDEN, "storage_fixture_rows.json")))
    return rows["ids"], np.load(os.path.join(GOLDEN, "storage_fixture_vectors.npy")), rows["metadata"], rows


def test_parquet_go_layout_fixture():
    """tests/golden/vectors_parquetgo.parquet carries the physical layout xitongsys/parquet-go derives from
    ParquetVectorRecord (parquet.go:16-20): REQUIRED fields, three-level LIST named list / element, dictionary
    on `id` only, SNAPPY. The reader applies parquet.go:133-166 to it."""
    import pyarrow.parquet as pq
    path = os.path.join(GOLDEN, "vectors_parquetgo.parquet")
    pf = pq.ParquetFile(path)
    text = str(pf.schema)
    for line in ("required binary field_id=-1 id (String);", "required group field_id=-1 vector (List) {",
                 "repeated group field_id=-1 list {", "required float field_id=-1 element;",
                 "required binary field_id=-1 metadata (String);"):
        assert line in text, text
    rg = pf.metadata.row_group(0)
    enc = {rg.column(i).path_in_schema: (set(rg.column(i).encodings), rg.column(i).compression) for i in range(rg.num_columns)}
    assert "PLAIN_DICTIONARY" in enc["id"][0] and "PLAIN_DICTIONARY" not in enc["metadata"][0]
    assert all(c == "SNAPPY" for _, c in enc.values())
    ids, vec, md, rows = _fixture()
    got_ids, mats, got_md = [], [], []
    for b_ids, mat, b_md in ingest.iter_parquet(path, dim=vec.shape[1]):
        got_ids += b_ids
        mats.append(mat)
        got_md += b_md
    assert got_ids == ids  # the empty-id row and the empty-vector row are gone
    assert np.array_equal(np.concatenate(mats).view(np.uint32), vec.view(np.uint32))
    want_md = [dict(m) for m in md]
    want_md[rows["unparsable_metadata_row"]] = {}
    assert [json.loads(m) for m in got_md] == want_md


def test_arrow_hnsw_fixture():
    """tests/golden/index_arrow_hnsw.arrow: the file ArrowHNSWIndex.Save writes (arrow_hnsw.go:153-197)."""
    ids, vec, _, _ = _fixture()
    got_ids, mats = [], []
    for b_ids, mat in ingest.iter_arrow_ipc(os.path.join(GOLDEN, "index_arrow_hnsw.arrow"), dim=vec.shape[1]):
        assert not mat.flags["OWNDATA"]  # the child buffer IS the row-major matrix
        got_ids += b_ids
        mats.append(mat.copy())
    assert got_ids == ids and len(mats) == 3
    assert np.array_equal(np.concatenate(mats).view(np.uint32), vec.view(np.uint32))


@pytest.mark.gpu
def test_fixture_files_loaded_to_hbm_answer_like_the_oracle(oracle):
    """Upload path end to end: the two fixture files -> pinned staging -> HBM; the searches over the loaded
    indexes equal the CPU oracle over the fixture's vectors (not a second GPU index)."""
    from oracle import filters as F
    from oracle import rerank
    from quiver_b200 import hostapi
    ids, vec, md, rows = _fixture()
    d = vec.shape[1]
    rng = np.random.default_rng(2)
    idx = hostapi.HybridIndex(d, "euclidean")
    assert ingest.load_arrow_ipc(os.path.join(GOLDEN, "index_arrow_hnsw.arrow"), idx, dim=d) == len(ids)
    col = hostapi.Collection("loaded", d, "cosine")
    assert ingest.load_parquet(os.path.join(GOLDEN, "vectors_parquetgo.parquet"), col, dim=d) == len(ids)
    md_loaded = [dict(m) for m in md]
    md_loaded[rows["unparsable_metadata_row"]] = {}
    raw = [json.dumps(m) for m in md_loaded]
    for _ in range(5):
        q = rng.standard_normal(d).astype(np.float32)
        od, orow = oracle.exact_search(vec, q, 10, 1)
        got = idx.Search(q, 10)
        assert [g[0] for g in got] == [ids[r] for r in orow]
        assert [np.float32(g[1]).view(np.uint32) for g in got] == [x.view(np.uint32) for x in od]
        mask = np.array(F.metadata_mask(raw, [("category", "=", "cat2"), ("lang", "=", "en")]))
        want = rerank.filtered_search(vec, ids, q, 6, 0, mask)
        got = col.Search(q, 6, Filters=[("category", "=", "cat2"), ("lang", "=", "en")])
        assert [g[0] for g in got] == [w[0] for w in want]
        assert [np.float32(g[1]).view(np.uint32) for g in got] == [np.float32(w[1]).view(np.uint32) for w in want]
    idx.close()
    col.close()


@pytest.mark.gpu
def test_arrow_ipc_load_matches_direct_insert(tmp_path):
    from quiver_b200 import hostapi
    ids, vec = _rows(3000, 32, seed=5)
    path = os.path.join(tmp_path, "index.arrow")
    ingest.save_arrow_ipc(path, ids, vec, batch_rows=1024)
    a = hostapi.HybridIndex(32, "euclidean")
    assert ingest.load_arrow_ipc(path, a, dim=32) == 3000
    b = hostapi.HybridIndex(32, "euclidean")
    b.InsertBatchArrays(ids, vec)
    assert a.Size() == b.Size() == 3000
    q = np.random.default_rng(9).random((5, 32), dtype=np.float32)
    for i in range(5):
        ra, rb = a.Search(q[i], 10), b.Search(q[i], 10)
        assert [r[0] for r in ra] == [r[0] for r in rb]
        assert [np.float32(r[1]).view(np.uint32) for r in ra] == [np.float32(r[1]).view(np.uint32) for r in rb]


@pytest.mark.gpu
def test_parquet_load_then_filtered_search(tmp_path):
    """A persisted collection reloaded from vectors.parquet answers a metadata-filtered search like
    the collection it was written from (db.go:248-263 re-adds every row with its metadata)."""
    from quiver_b200 import hostapi
    ids, vec = _rows(2000, 16, seed=11)
    cats = ["cat%d" % (i % 5) for i in range(len(ids))]
    md = [{"category": c} for c in cats]
    path = os.path.join(tmp_path, "vectors.parquet")
    ingest.save_parquet(path, ids, vec, md)
    src = hostapi.Collection("src", 16, "cosine")
    src.AddBatch(ids, vec, md)
    dst = hostapi.Collection("dst", 16, "cosine")
    assert ingest.load_parquet(path, dst, dim=16) == 2000
    assert dst.Count() == 2000
    q = np.random.default_rng(13).random(16, dtype=np.float32)
    want = src.Search(q, 7, Filters=[("category", "=", "cat3")])
    got = dst.Search(q, 7, Filters=[("category", "=", "cat3")])
    assert [r[0] for r in got] == [r[0] for r in want] and len(got) == 7
    assert all(cats[int(r[0][2:])] == "cat3" for r in got)
