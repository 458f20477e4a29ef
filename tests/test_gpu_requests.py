"""Rows a14 / a17 / a19 / a21 of SURVEY 8: the request / response structs around the search path.
  * Collection.Search(types.SearchRequest) with SearchOptions (IncludeVectors / IncludeMetadata) and the
    FluentSearch builder calls that set them; SearchResultItem{ID, Distance, Score, Vector, Metadata}
    (pkg/core/collection.go:758-779, 946-985; pkg/types/search.go:31-52)
  * persistence.Collection.Search / SearchWithFacets (pkg/persistence/collection.go:226-261, 327-378)
  * HNSWAdapter.SearchWithNegativeExample with the weight clamp (pkg/hnsw/adapter.go:345-437)"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from quiver_b200 import hostapi
    hostapi.load()
    return hostapi


def test_search_request_options_and_result_items(H, oracle):
    rng = np.random.default_rng(21)
    n, d, k = 3000, 32, 7
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    ids = [f"v{i:05d}" for i in range(n)]
    docs = [None if i % 11 == 0 else json.dumps({"category": f"cat{i % 4}", "n": i}) for i in range(n)]
    c = H.Collection("items", d, "cosine")
    c.AddBatch(ids, corpus, docs)
    q = rng.standard_normal(d).astype(np.float32)
    od, orow = oracle.exact_search(corpus, q, k, 0)

    # FluentSearch defaults: k = 10, IncludeMetadata = true, no vectors (collection.go:887-895)
    items = c.FluentSearch(q).WithK(k).ExecuteResponse()
    assert [it["ID"] for it in items] == [ids[r] for r in orow]
    for it, dist, r in zip(items, od, orow):
        assert np.float32(it["Distance"]).view(np.uint32) == dist.view(np.uint32)
        assert np.float32(it["Score"]).view(np.uint32) == np.float32(np.float32(1.0) - dist).view(np.uint32)  # 1.0 - res.Distance
        assert "Vector" not in it
        if docs[r] is None:
            assert "Metadata" not in it          # no document stored: `omitempty`
        else:
            assert it["Metadata"] == docs[r]     # the stored json.RawMessage, byte for byte

    items = c.FluentSearch(q).WithK(k).IncludeVectors(True).IncludeMetadata(False).UseExactSearch() \
        .WithNamespace("tenant-a").ExecuteResponse()
    for it, r in zip(items, orow):
        assert "Metadata" not in it
        assert np.array_equal(it["Vector"].view(np.uint32), corpus[r].view(np.uint32))  # c.Vectors[res.ID]

    # with a filter, and after an Update that moves the row: the decoration follows the id
    items = c.SearchRequest(q, 5, [("category", "=", "cat2")], IncludeVectors=True, IncludeMetadata=True)
    assert all(json.loads(it["Metadata"])["category"] == "cat2" for it in items)
    first = items[0]["ID"]
    newv = rng.standard_normal(d).astype(np.float32)
    c.Update(first, newv, {"category": "cat2", "moved": True})
    it2 = [it for it in c.SearchRequest(newv, 1, [], IncludeVectors=True, IncludeMetadata=True)]
    assert it2[0]["ID"] == first and json.loads(it2[0]["Metadata"])["moved"] is True
    assert np.array_equal(it2[0]["Vector"], newv)

    # a failed builder call sticks (the `valid` flag): later calls are ignored, Execute returns the first error
    with pytest.raises(H.QuiverError, match="k must be greater than 0"):
        c.FluentSearch(q).WithK(0).IncludeVectors(True).ExecuteResponse()
    with pytest.raises(H.QuiverError, match="invalid vector dimension: expected 32, got 3"):
        c.FluentSearch([1, 2, 3]).WithK(5).ExecuteResponse()
    c.close()


def test_persistence_collection_search_semantics(H, oracle):
    """pkg/persistence/collection.go:226-261, 327-378 restated: filter first (rows without facet values never
    pass), distances of the survivors, ascending, limit <= 0 = everything, that type's error text."""
    from oracle import filters as F
    rng = np.random.default_rng(5)
    n, d = 4000, 24
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    ids = [f"p{i:05d}" for i in range(n)]
    rows = [None if i % 13 == 0 else {"category": f"cat{rng.integers(4)}", "price": float(rng.integers(0, 100))}
            for i in range(n)]
    c = H.Collection("persist", d, "euclidean")
    c.AddBatch(ids, corpus, [None if r is None else json.dumps(r) for r in rows])
    c.SetFacetFields(["category", "price"])
    q = rng.standard_normal(d).astype(np.float32)
    dist = oracle.distances(1, corpus, q)

    raw = ["" if r is None else json.dumps(r) for r in rows]

    def restated(limit, flt):
        # rows without facet values are skipped, the others must pass MatchesAllFilters (:346-355)
        mask = [True] * n if flt is None else F.facet_mask(raw, ["category", "price"], flt)
        keep = [i for i in range(n) if mask[i]]
        keep.sort(key=lambda i: (dist[i], i))
        return keep[:limit] if 0 < limit < len(keep) else keep

    for limit in (10, 0, -3):
        got = c.PersistenceSearch(q, limit)
        want = restated(limit, None)
        assert len(got) == len(want) == (10 if limit == 10 else n)
        assert [g[0] for g in got[:50]] == [ids[i] for i in want[:50]]
        assert np.array_equal(np.array([g[1] for g in got], dtype=np.float32).view(np.uint32), dist[want].view(np.uint32))
    flt_o = [F.EqualityFilter("category", "cat1"), F.RangeFilter("price", {"float": 10.0}, {"float": 60.0}, True, False)]
    assert 0 < sum(F.facet_mask(raw, ["category", "price"], flt_o)) < n // 4
    flt_h = [H.NewEqualityFilter("category", "cat1"), H.NewRangeFilter("price", 10.0, 60.0, True, False)]
    for limit in (5, 0):
        got = c.PersistenceSearch(q, limit, flt_h)
        want = restated(limit, flt_o)
        assert [g[0] for g in got] == [ids[i] for i in want]
    with pytest.raises(H.QuiverError, match="query vector dimension mismatch: got 3, expected 24"):
        c.PersistenceSearch([1, 2, 3], 5)
    with pytest.raises(H.QuiverError, match="query vector is nil"):
        c.PersistenceSearch(None, 5)
    c.close()


@pytest.mark.parametrize("distance,metric", [("euclidean", 1), ("cosine", 0)])
def test_hnsw_search_with_negative_example(H, oracle, distance, metric):
    """pkg/hnsw/adapter.go:345-437 restated on the oracle's walk: max(2k, 30) candidates, Distance - w * negDistance
    in float32 with w clamped to 1, stable (Distance, ID) order, first k; the returned Distance is the adjusted one."""
    from oracle import hnsw
    rng = np.random.default_rng(31 + metric)
    n, d, k = 5000, 32, 10
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    ids = [f"n{i:05d}" for i in range(n)]
    os.environ["QO_HNSW_STANDARD"] = "1"  # a graph whose walks fill max(2k, 30) results
    graph = hnsw.Graph(corpus, metric, M=16, MaxM0=32, EfConstruction=100, EfSearch=64, seed=3)
    os.environ["QO_HNSW_STANDARD"] = "0"
    idx = H.HybridIndex(d, distance)
    idx.InsertBatch({ids[i]: corpus[i] for i in range(n)})
    dg = idx.HNSWUpload(graph.export())
    for trial in range(6):
        q = rng.standard_normal(d).astype(np.float32)
        neg = rng.standard_normal(d).astype(np.float32)
        for w in (0.5, 1.0, 3.0):
            retrieve = max(2 * k, 30)
            od, oidx, _, _ = graph.search(q, retrieve)
            assert len(oidx) == retrieve
            wc = np.float32(min(w, 1.0))
            ext = []
            for dist, j in zip(od, oidx):
                nd = oracle.distance(metric, corpus[j], neg)
                ext.append((np.float32(dist - np.float32(wc * nd)), ids[j]))
            ext.sort(key=lambda e: (e[0], e[1]))
            got = dg.search_with_negative_example(q, neg, w, k)
            assert [g[0] for g in got] == [e[1] for e in ext[:k]], (trial, w)
            assert [np.float32(g[1]).view(np.uint32) for g in got] == [e[0].view(np.uint32) for e in ext[:k]]
        # no negative example / zero weight: the plain search truncated to k
        plain, _, _ = dg.search(q[None, :], k)
        assert dg.search_with_negative_example(q, None, 0.5, k) == plain[0]
        assert dg.search_with_negative_example(q, neg, 0.0, k) == plain[0]
        # a negative example of the wrong dimension makes every candidate's DistanceFunc fail: nothing is returned
        assert dg.search_with_negative_example(q, neg[:5], 0.5, k) == []
    with pytest.raises(H.QuiverError, match="k must be positive"):
        dg.search_with_negative_example(q, neg, 0.5, 0)
    dg.close()
    idx.close()
