"""Tombstone compaction (qg_index_compact): after a burst of deletes the device index is squeezed so
that, like the reference's map after Delete (exact.go:61-70, hybrid_index.go:244-290), it holds live vectors only.
Bar: the compacted index answers exactly like a fresh index built from the surviving rows — distances
bit-identical to the oracle, row lists identical — in the flat and the tensor-core regime; filter masks over the
compacted facet columns bit-exact against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K_NUMBER, K_OTHER, K_MISSING, K_NOROW = 3, 5, 0, 6
OP_NUM_CMP, OP_ELEM_IN = 5, 15


def _parity(oracle, idx, corpus, queries, k, metric, which, label):
    dist, row, cnt, _ = idx.search(queries, k)
    for i in which:
        od, orow = oracle.exact_search(corpus, queries[i], k, metric, 0, None)
        n = len(od)
        assert cnt[i] == n, f"{label} q{i}: count {cnt[i]} != {n}"
        assert np.array_equal(row[i, :n], orow), f"{label} q{i}: rows {row[i, :n]} vs {orow} ({idx.stats()})"
        assert np.array_equal(dist[i, :n].view(np.uint32), od.view(np.uint32)), f"{label} q{i}: distances"
    return dist, row, cnt


@pytest.mark.parametrize("metric,d", [(1, 96), (0, 128), (0, 50), (2, 7), (4, 33), (0, 768), (1, 516)])
def test_compact_matches_fresh_index(capi, oracle, metric, d):
    rng = np.random.default_rng(100 + d)
    n = 70000
    corpus = rng.random((n, d), dtype=np.float32)
    queries = rng.random((40, d), dtype=np.float32)
    idx = capi.Index(d, metric)
    idx.upload(corpus)
    dead = rng.choice(n, 30000, replace=False)
    d0, r0, _, _ = idx.search(queries, 10)
    dead = np.unique(np.concatenate([dead, r0[:, :4].ravel(), [0, 31, 32, n - 1]]))  # word edges, current winners
    idx.tombstone(dead)
    before_d, before_r, before_c, _ = idx.search(queries, 10)
    live = np.ones(n, dtype=bool)
    live[dead] = False
    old_to_new = idx.compact()
    # the map: dead -> -1, live rows keep their order and become 0 .. size-1
    assert np.all(old_to_new[~live] == -1)
    assert np.array_equal(old_to_new[live], np.arange(int(live.sum())))
    assert idx.rows == idx.size == int(live.sum())
    kept = corpus[live]
    sample = np.array([0, 1, 31, 32, 33, 1023, 1024, len(kept) // 2, len(kept) - 1])
    assert np.array_equal(idx.fetch(sample).view(np.uint32), kept[sample].view(np.uint32))
    # batch (tensor-core regime where the metric has one) and single query (flat scan) against the oracle
    dist, row, cnt = _parity(oracle, idx, kept, queries, 10, metric, [0, 7, 39], f"compact/m{metric}/d{d}")
    _parity(oracle, idx, kept, queries[:1], 10, metric, [0], f"compact-flat/m{metric}/d{d}")
    # and against the same index before the squeeze, through the map
    assert np.array_equal(cnt, before_c)
    assert np.array_equal(row, old_to_new[before_r])
    assert np.array_equal(dist.view(np.uint32), before_d.view(np.uint32))
    # compacting a compact index is the identity
    again = idx.compact()
    assert np.array_equal(again, np.arange(len(kept)))
    # the index keeps growing after it shrank
    more = rng.random((5000, d), dtype=np.float32)
    first = idx.upload(more)
    assert first == len(kept)
    both = np.concatenate([kept, more])
    _parity(oracle, idx, both, queries, 10, metric, [0, 20], f"compact+upload/m{metric}/d{d}")
    idx.tombstone(np.arange(len(both)))
    assert idx.compact().tolist() == [-1] * len(both) and idx.rows == 0
    dist, row, cnt, _ = idx.search(queries[:2], 5)
    assert np.all(cnt == 0)
    first = idx.upload(more[:10])
    assert first == 0
    _parity(oracle, idx, more[:10], queries[:2], 5, metric, [0, 1], "compact-to-empty then upload")
    idx.close()


def test_compact_facet_columns_and_element_lists(capi):
    rng = np.random.default_rng(7)
    n, d = 20000, 16
    idx = capi.Index(d, 1)
    idx.upload(rng.random((n, d), dtype=np.float32))
    # column 0: numbers on the first 15 000 rows (the later rows have no metadata entry)
    n0 = 15000
    num = rng.integers(0, 100, n0).astype(np.float64)
    kind0 = np.full(n0, K_NUMBER, dtype=np.uint8)
    kind0[rng.random(n0) < 0.1] = K_MISSING
    idx.set_column(0, kind0, num, np.full(n0, -1, np.int32), np.full(n0, -1, np.int32))
    # column 1: every third row carries an array of 0..5 element codes
    is_arr = (np.arange(n) % 3) == 0
    lens = np.where(is_arr, rng.integers(0, 6, n), 0)
    off = np.zeros(n + 1, dtype=np.int32)
    off[1:] = np.cumsum(lens)
    codes = rng.integers(0, 12, int(off[-1])).astype(np.int32)
    kind1 = np.where(is_arr, K_OTHER | 0x80, K_NUMBER).astype(np.uint8)
    idx.set_column(1, kind1, np.zeros(n), np.full(n, -1, np.int32), np.full(n, -1, np.int32))
    idx.set_array_column(1, off, codes)

    f_num = capi.Filter(idx, [capi.qg_pred(0, 1, 0, 0)], [capi.qg_clause(OP_NUM_CMP, 0, 0, 2, 0, 0, 40.0, 0.0)])
    f_not = capi.Filter(idx, [capi.qg_pred(0, 1, 1, 1)], [capi.qg_clause(OP_NUM_CMP, 0, 0, 2, 0, 0, 40.0, 0.0)])
    f_elem = capi.Filter(idx, [capi.qg_pred(0, 1, 0, 0)], [capi.qg_clause(OP_ELEM_IN, 1, 0, 0, 0, 3, 0.0, 0.0)],
                         iset=[2, 5, 11])
    f_unknown = capi.Filter(idx, [capi.qg_pred(0, 1, 0, 0)], [capi.qg_clause(2, 9, 0, 1 << K_MISSING, 0, 0, 0.0, 0.0)])

    def want_num(k, v):
        return (k == K_NUMBER) & (v > 40.0)

    want0 = np.zeros(n, dtype=bool)
    want0[:n0] = want_num(kind0, num)
    has_row = np.zeros(n, dtype=bool)
    has_row[:n0] = True
    want_not = ~want0 & has_row  # require_row: rows without an entry never match
    want1 = np.array([bool(np.isin(codes[off[r]:off[r + 1]], [2, 5, 11]).any()) for r in range(n)])
    for f, w in ((f_num, want0), (f_not, want_not), (f_elem, want1), (f_unknown, np.ones(n, dtype=bool))):
        bits, m = f.eval()
        assert np.array_equal(bits, w) and m == int(w.sum())

    dead = rng.choice(n, 9000, replace=False)
    idx.tombstone(dead)
    live = np.ones(n, dtype=bool)
    live[dead] = False
    old_to_new = idx.compact()
    assert idx.rows == int(live.sum())
    # the filters compiled before the squeeze answer over the new numbering
    for f, w in ((f_num, want0), (f_not, want_not), (f_elem, want1), (f_unknown, np.ones(n, dtype=bool))):
        bits, m = f.eval()
        assert bits.shape[0] == int(live.sum())
        assert np.array_equal(bits, w[live]), "mask over the compacted columns"
        assert m == int(w[live].sum())
    # a filtered search returns rows of the new numbering that pass the predicate
    q = rng.random((3, d), dtype=np.float32)
    dist, row, cnt, _ = idx.search(q, 10, filter=f_elem)
    new_to_old = np.nonzero(live)[0]
    assert np.all(cnt == 10) and np.all(want1[new_to_old[row]])
    # rows uploaded after the squeeze have no entry in either column
    idx.upload(rng.random((100, d), dtype=np.float32))
    bits, m = f_num.eval()
    assert np.array_equal(bits[:-100], want0[live]) and not bits[-100:].any()
    bits, m = f_elem.eval()
    assert np.array_equal(bits[:-100], want1[live]) and not bits[-100:].any()
    for f in (f_num, f_not, f_elem, f_unknown):
        f.close()
    idx.close()


def test_compact_large_index_multi_round_scan(capi):
    """6M rows, 4.5M survivors: the prefix sum over the live mask runs through 46 scan blocks, the one over the
    element counts of the new rows through 1 099 blocks, i.e. two rounds of the single-block spine."""
    n, d = 6_000_000, 8
    idx = capi.Index(d, 1)
    idx.upload_synthetic(0, 42, 0, n)
    rng = np.random.default_rng(3)
    # one column: every third row is a one-element array
    is_arr = (np.arange(n) % 3) == 0
    off = np.zeros(n + 1, dtype=np.int32)
    off[1:] = np.cumsum(is_arr)
    codes = rng.integers(0, 50, int(off[-1])).astype(np.int32)
    kind = np.where(is_arr, K_OTHER | 0x80, K_MISSING).astype(np.uint8)
    idx.set_column(0, kind, np.zeros(n), np.full(n, -1, np.int32), np.full(n, -1, np.int32))
    idx.set_array_column(0, off, codes)
    f = capi.Filter(idx, [capi.qg_pred(0, 1, 0, 0)], [capi.qg_clause(OP_ELEM_IN, 0, 0, 0, 0, 4, 0.0, 0.0)],
                    iset=[1, 7, 19, 42])
    want = np.zeros(n, dtype=bool)
    want[is_arr] = np.isin(codes, [1, 7, 19, 42])
    bits, m = f.eval()
    assert np.array_equal(bits, want)
    dead = rng.choice(n, 1_500_000, replace=False)
    idx.tombstone(dead)
    live = np.ones(n, dtype=bool)
    live[dead] = False
    probe_old = np.nonzero(live)[0][::50021]
    before = idx.fetch(probe_old)
    old_to_new = idx.compact()
    assert idx.rows == idx.size == n - 1_500_000
    assert np.array_equal(old_to_new[live], np.arange(int(live.sum()))) and np.all(old_to_new[~live] == -1)
    assert np.array_equal(idx.fetch(old_to_new[probe_old]).view(np.uint32), before.view(np.uint32))
    bits, m = f.eval()
    assert np.array_equal(bits, want[live]) and m == int(want[live].sum())
    f.close()
    idx.close()


def test_collection_compact_keeps_ids_metadata_and_answers(oracle):
    """Collection level: after Delete x many + Compact, filtered and unfiltered searches return the same ids and
    bit-identical distances as before the squeeze; deleted ids can be added again."""
    import json
    from quiver_b200 import hostapi as H
    from oracle import filters as F
    from oracle import rerank
    H.load()
    rng = np.random.default_rng(21)
    n, d, k = 12000, 48, 10
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    ids = [f"v{i:05d}" for i in range(n)]
    rows = [{"category": f"cat{rng.integers(5)}", "labels": [f"l{x}" for x in rng.integers(0, 8, rng.integers(0, 4))]}
            for _ in range(n)]
    c = H.Collection("squeeze", d, "euclidean")
    c.AddBatch(ids, corpus, rows)
    c.SetFacetFields(["category", "labels"])
    q = rng.standard_normal(d).astype(np.float32)
    dead_rows = rng.choice(n, 7000, replace=False)
    for r in dead_rows:
        c.Delete(ids[r])
    live = np.ones(n, dtype=np.uint8)
    live[dead_rows] = 0
    flt = [("category", "=", "cat2")]
    facet = [H.NewSetFilter("labels", ["L3", "l5"]), H.NewEqualityFilter("category", "CAT2")]
    before = (c.Search(q, k), c.Search(q, k, flt), c.SearchWithFacets(q, k, facet))
    assert c.Compact() == 7000
    assert c.Count() == n - 7000 and c.Compact() == 0
    after = (c.Search(q, k), c.Search(q, k, flt), c.SearchWithFacets(q, k, facet))
    for b, a in zip(before, after):
        assert [x[0] for x in a] == [x[0] for x in b]
        assert [np.float32(x[1]).view(np.uint32) for x in a] == [np.float32(x[1]).view(np.uint32) for x in b]
    # and against the oracle over the survivors
    raw = [json.dumps(m) for m in rows]
    mask = np.array(F.metadata_mask(raw, [("category", "=", "cat2")]))
    want = rerank.filtered_search(corpus, ids, q, k, 1, mask, live=live)
    assert [x[0] for x in after[1]] == [w[0] for w in want]
    assert [np.float32(x[1]).view(np.uint32) for x in after[1]] == [np.float32(w[1]).view(np.uint32) for w in want]
    # a deleted id comes back under the same name with a new vector (exact.go:44-50 only rejects live ids)
    back = ids[int(dead_rows[0])]
    c.Add(back, q, {"category": "cat2", "labels": ["l5"]})
    got = c.Search(q, 1, flt)
    assert got[0][0] == back and got[0][1] == 0.0
    with pytest.raises(H.QuiverError, match="vector not found"):
        c.Delete(ids[int(dead_rows[1])])
    c.close()
