"""The reference-facing host layer (libquiverhost.so: HybridIndex / Collection / FluentSearch names)
over the GPU C ABI, against the reference's own test cases (tests/golden) and the oracle."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def num(x):
    for key in ("int", "float", "float32"):
        if isinstance(x, dict) and key in x:
            return x[key]
    raise TypeError(x)


def vec(x):
    if isinstance(x, dict) and "f32vec" in x:
        return np.asarray(x["f32vec"], dtype=np.float32)
    if isinstance(x, dict) and "items" in x:
        return np.asarray([float(num(i)) for i in x["items"]], dtype=np.float32)
    return np.asarray(x, dtype=np.float32)


def py(v):
    """typed golden value -> the Python literal hostapi marshals as a Go literal"""
    if v is None or isinstance(v, (bool, str)):
        return v
    if "int" in v:
        return int(v["int"])
    if "float" in v:
        return float(v["float"])
    if "float32" in v:
        return float(v["float32"])
    if "strings" in v:
        return list(v["strings"])
    if "list" in v:
        return [py(x) for x in v["list"]]
    if "map" in v:
        return {k: py(x) for k, x in v["map"].items()}
    raise TypeError(v)


@pytest.fixture(scope="module")
def H():
    from quiver_b200 import hostapi
    hostapi.load()
    return hostapi


# ---- hybrid index --------------------------------------------------------------------------------
def test_exact_index_search_golden(H):
    """pkg/hybrid/exact_test.go:97-205 through the host index."""
    g = load("exact_search.json")["exact_search"]
    idx = H.HybridIndex(3, "cosine")
    for id_, v in g["vectors"].items():
        idx.Insert(id_, vec(v))
    for case in g["cases"]:
        res = idx.SearchWithRequest(vec(case["query"]), int(num(case["k"])), ForceStrategy="exact")
        ids = [r[0] for r in res]
        want = case["wantIDs"]["strings"]
        assert len(ids) == min(int(num(case["k"])), 3)
        assert all(res[i][1] <= res[i + 1][1] for i in range(len(res) - 1))
        assert ids == want if case["exactOrder"] else set(ids) == set(want), case["name"]
    idx.close()


def test_index_errors_have_the_reference_text(H):
    """exact.go:46-49,101-106; hybrid_index.go:88-96,247-250,392-402,581,678-680."""
    idx = H.HybridIndex(3, "cosine")
    assert idx.Search([1, 0, 0], 5) == []  # empty index: no results, no error
    idx.Insert("a", [1, 0, 0])
    with pytest.raises(H.QuiverError, match="vector with ID a already exists"):
        idx.Insert("a", [0, 1, 0])
    with pytest.raises(H.QuiverError, match="vector dimension mismatch: expected 3, got 2"):
        idx.Insert("b", [0, 1])
    with pytest.raises(H.QuiverError, match="query dimension mismatch: expected 3, got 4"):
        idx.Search([1, 0, 0, 0], 1)
    with pytest.raises(H.QuiverError, match="k must be positive"):
        idx.Search([1, 0, 0], 0)
    with pytest.raises(H.QuiverError, match="vector with ID zzz not found"):
        idx.Delete("zzz")
    with pytest.raises(H.QuiverError, match="invalid search strategy: bogus"):
        idx.SearchWithRequest([1, 0, 0], 1, ForceStrategy="bogus")
    with pytest.raises(H.QuiverError, match="negative example dimension mismatch: expected 3, got 2"):
        idx.SearchWithRequest([1, 0, 0], 1, NegativeExample=[1, 0], NegativeWeight=0.5)
    with pytest.raises(H.QuiverError, match="no queries provided"):
        idx.BatchSearch([], 3)
    v = np.array([0.0, 1.0, 0.0], dtype=np.float32)
    idx.Insert("c", v)
    v[:] = 9  # Insert copies (exact_test.go:56-60)
    assert idx.Search([0, 1, 0], 1)[0][0] == "c"
    idx.Delete("c")
    assert idx.Size() == 1 and [r[0] for r in idx.Search([0, 1, 0], 5)] == ["a"]
    idx.close()


def test_batch_search_golden(H):
    """hybrid_index_test.go:453-518: two queries, forced exact => vec1, vec2."""
    idx = H.HybridIndex(3, "cosine")
    idx.InsertBatch({"vec1": [1, 0, 0], "vec2": [0, 1, 0], "vec3": [0, 0, 1]})
    res = idx.BatchSearch([[0.9, 0.1, 0], [0.1, 0.9, 0]], 1, ForceStrategy="exact")
    assert [r[0][0] for r in res] == ["vec1", "vec2"]
    idx.close()


def test_negative_example_golden(H, oracle):
    """hybrid_index_test.go:541-657 and hybrid_index_rerank_test.go:9-47."""
    from oracle import rerank
    g = load("exact_search.json")["negative_example"]
    ids = list(g["vectors"])
    corpus = np.stack([vec(g["vectors"][i]) for i in ids])
    idx = H.HybridIndex(4, "cosine")
    for i, id_ in enumerate(ids):
        idx.Insert(id_, corpus[i])
    neg = corpus[ids.index(g["negative_id"])]
    q = np.asarray(g["query"], np.float32)
    res = idx.FluentSearch(q).WithK(g["k"]).WithStrategy("exact").WithNegativeExample(neg).WithNegativeWeight(g["weight"]).Execute()
    want = rerank.search_with_negative(corpus, ids, q, g["k"], 0, neg, g["weight"])
    assert len(res) == 3 and res[0][0] != g["negative_id"]
    assert [r[0] for r in res] == [w[0] for w in want]
    assert [np.float32(r[1]).view(np.uint32) for r in res] == [np.float32(w[1]).view(np.uint32) for w in want]
    idx.close()
    s = load("exact_search.json")["rerank_stability"]
    idx = H.HybridIndex(2, "euclidean")
    for id_ in sorted(s["vectors"]):
        idx.Insert(id_, s["vectors"][id_])
    res = idx.SearchWithRequest(s["query"], s["k"], "exact", s["negative"], s["weight"])
    assert [r[0] for r in res] == ["1", "2", "3"] and all(float(r[1]) == 0.0 for r in res)
    idx.close()


@pytest.mark.parametrize("distance,metric", [("cosine", 0), ("euclidean", 1), ("dot_product", 2)])
def test_negative_example_random_parity(H, oracle, distance, metric):
    from oracle import rerank
    rng = np.random.default_rng(7 + metric)
    n, d, k = 3000, 48, 10
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    ids = [f"id{j:05d}" for j in rng.permutation(n)]
    idx = H.HybridIndex(d, distance)
    idx.InsertBatch({ids[j]: corpus[j] for j in range(n)})
    # InsertBatch keeps dict order => row j holds ids[j]
    queries = rng.standard_normal((5, d)).astype(np.float32)
    negs = rng.standard_normal((5, d)).astype(np.float32)
    res = idx.BatchSearch(list(queries), k, "exact", list(negs), 0.5)
    for i in range(5):
        want = rerank.search_with_negative(corpus, ids, queries[i], k, metric, negs[i], 0.5)
        assert [r[0] for r in res[i]] == [w[0] for w in want], i
        got_bits = [np.float32(r[1]).view(np.uint32) for r in res[i]]
        assert got_bits == [np.float32(w[1]).view(np.uint32) for w in want], i
    # a zero weight ignores the negative example (hybrid_index.go:417)
    plain = idx.BatchSearch(list(queries), k, "exact")
    assert idx.BatchSearch(list(queries), k, "exact", list(negs), 0.0) == plain
    idx.close()


# ---- filters -------------------------------------------------------------------------------------
OPS = {"Equals": "=", "NotEquals": "!=", "GreaterThan": ">", "GreaterThanOrEqual": ">=", "LessThan": "<",
       "LessThanOrEqual": "<=", "In": "in", "NotIn": "not_in"}


def test_core_matches_filter_golden_through_the_device(H):
    """pkg/core/collection_test.go:727-816: every row of the truth table as a one-row collection."""
    for r in load("filters.json")["core_matches_filter"]:
        c = H.Collection("t", 2)
        c.Add("x", [1, 0], {k: py(v) for k, v in r["metadata"]["map"].items()})
        fl = r["filter"]["fields"]
        op = fl["Operator"]
        op = OPS[op["ident"]] if isinstance(op, dict) else op
        mask = c.filter_mask(Filters=[(fl["Field"], op, py(fl["Value"]))])
        assert bool(mask[0]) == r["want"], r["name"]
        c.close()


def _facet_filter(H, kind, r):
    if kind == "equality":
        return H.NewEqualityFilter(r["field"], py(r["value"]))
    if kind == "range":
        return H.NewRangeFilter(r["field"], py(r["min"]), py(r["max"]), r["includeMin"], r["includeMax"])
    if kind == "set":
        return H.NewSetFilter(r["field"], [py(v) for v in r["values"]["list"]])
    return H.NewExistsFilter(r["field"], r["shouldExist"])


@pytest.mark.parametrize("kind", ["equality", "range", "set", "exists"])
def test_facet_truth_tables_golden_through_the_device(H, kind):
    """pkg/facets/facets_test.go:10-196. A nil test value cannot be stored as a facet (ExtractFacets
    drops nil, facets.go:423), so those rows are checked through MatchesAllFilters' missing-field rule."""
    from oracle import filters as F
    seen = 0
    for r in load("filters.json")["facets"][kind]:
        tv = r["testVal"]
        c = H.Collection("t", 2)
        c.SetFacetFields([r["field"], "other"])
        md = {"other": 1}
        if tv is not None:
            md[r["field"]] = py(tv)
        c.Add("x", [1, 0], md)
        flt = _facet_filter(H, kind, r)
        # the oracle's MatchesAllFilters on the same facets is the expectation
        typed_md = {"map": {"other": {"float": 1.0}}}
        if tv is not None:
            from oracle.gotypes import from_json
            typed_md["map"][r["field"]] = from_json(py(tv))
        ofl = {"equality": lambda: F.EqualityFilter(r["field"], r["value"]),
               "range": lambda: F.RangeFilter(r["field"], r["min"], r["max"], r["includeMin"], r["includeMax"]),
               "set": lambda: F.SetFilter(r["field"], r["values"]["list"]),
               "exists": lambda: F.ExistsFilter(r["field"], r["shouldExist"])}[kind]()
        want = F.matches_all_filters(F.extract_facets(typed_md, [r["field"], "other"]), [ofl])
        if tv is not None and not isinstance(py(tv), list):
            assert want == r["expected"], r["name"]  # same answer as Filter.Match in the reference's table
        try:
            got = bool(c.filter_mask(facet_filters=[flt])[0])
        except H.QuiverError:
            # the one shape still rejected rather than answered wrongly: an Equality filter whose own
            # value is an array / map, against an array-valued facet (reflect.DeepEqual of two slices)
            assert kind == "equality" and isinstance(py(tv), list) and isinstance(py(r["value"]), (list, dict)), r["name"]
        else:
            assert got == want, r["name"]
            seen += 1
        c.close()
    assert seen >= 4


def _random_metadata(rng, n):
    cats = ["cat0", "cat1", "CAT2", "cat3", "Cat4"]
    rows = []
    for i in range(n):
        u = rng.random()
        if u < 0.05:
            rows.append(None)
            continue
        md = {}
        if rng.random() < 0.9:
            md["category"] = cats[rng.integers(len(cats))]
        if rng.random() < 0.9:
            md["tag"] = f"tag{rng.integers(20):02d}"
        r = rng.random()
        if r < 0.6:
            md["price"] = float(np.round(rng.random() * 200, 2))
        elif r < 0.7:
            md["price"] = int(rng.integers(0, 200))
        elif r < 0.8:
            md["price"] = str(int(rng.integers(0, 200)))
        elif r < 0.85:
            md["price"] = None
        if rng.random() < 0.5:
            md["active"] = bool(rng.random() < 0.5)
        if rng.random() < 0.3:
            md["big"] = float(rng.integers(1, 50)) * 1e5
        if rng.random() < 0.2:
            md["nested"] = {"level": int(rng.integers(0, 4)), "name": ""}
        a = rng.random()
        if a < 0.35:   # array-valued facet (facets.go:308-320): strings, numbers, mixed, sometimes empty
            pool = ["red", "Red", "blue", 1, 2.0, 2.5, True, None, "3"]
            md["labels"] = [pool[j] for j in rng.integers(0, len(pool), size=int(rng.integers(0, 4)))]
        elif a < 0.40:
            md["labels"] = {"k": int(rng.integers(0, 2))}   # a map is compared whole
        elif a < 0.50:
            md["labels"] = ["red", "blue", 2.0, None][int(rng.integers(0, 4))]  # scalar in the same column
        if u < 0.1:
            md = {}
        rows.append(md)
    return rows


def test_metadata_masks_bit_exact(H):
    """core.matchesFilter over random mixed-type metadata: every predicate's mask equals the oracle's."""
    from oracle import filters as F
    rng = np.random.default_rng(11)
    n = 3000
    rows = _random_metadata(rng, n)
    c = H.Collection("m", 4)
    c.AddBatch([f"r{i}" for i in range(n)], rng.random((n, 4), dtype=np.float32), rows)
    raw = [None if m is None else json.dumps(m) for m in rows]

    def typed(v):
        if isinstance(v, bool) or v is None or isinstance(v, str):
            return v
        if isinstance(v, int):
            return {"int": v}
        if isinstance(v, float):
            return {"float": v}
        return {"list": [typed(x) for x in v]}

    cases = [
        [("category", "=", "cat3")], [("category", "=", "CAT3")], [("category", "!=", "cat0")],
        [("category", "in", ["cat1", "cat3", "nope"])], [("category", "not_in", ["cat1", "cat3"])],
        [("category", "not_in", "cat1")], [("category", "in", "cat1")],
        [("price", ">", 100)], [("price", ">=", 100.0)], [("price", "<", 50)], [("price", "<=", 12.5)],
        [("price", "=", 100)], [("price", "=", "100")], [("price", "!=", 7)], [("price", "in", [1, 2, 3, "4", 5.5])],
        [("price", ">", "50")], [("price", "<", "zzz")], [("price", "=", None)],
        [("active", "=", True)], [("active", "=", "true")], [("active", "!=", False)], [("active", ">", False)],
        [("big", "=", 1e6)], [("big", "=", "1e+06")], [("big", "=", 1000000)], [("big", ">", "2e+06")],
        [("nested", "=", "map[level:1 name:]")], [("missing", "!=", 1)], [("missing", "not_in", [1])],
        [("category", "=", "cat3"), ("tag", "in", ["tag00", "tag01", "tag02", "tag03", "tag04"])],
        [("category", "=", "cat3"), ("price", ">", 20), ("active", "!=", True)],
        [("category", "bogus", "cat3")],
    ]
    for flt in cases:
        got = c.filter_mask(Filters=flt)
        want = np.array(F.metadata_mask(raw, [(f, op, typed(v)) for f, op, v in flt]))
        assert np.array_equal(got, want), (flt, np.nonzero(got != want)[0][:5], [rows[i] for i in np.nonzero(got != want)[0][:3]])
    c.close()


def test_facet_masks_bit_exact(H):
    from oracle import filters as F
    rng = np.random.default_rng(12)
    n = 3000
    rows = _random_metadata(rng, n)
    c = H.Collection("f", 4)
    fields = ["category", "price", "active", "nested.level", "nested.name", "absent", "labels"]
    c.AddBatch([f"r{i}" for i in range(n)], rng.random((n, 4), dtype=np.float32), rows)
    c.SetFacetFields(fields)
    raw = [None if m is None else json.dumps(m) for m in rows]
    I = lambda v: {"int": v}
    Fl = lambda v: {"float": v}
    cases = [
        ([H.NewEqualityFilter("category", "CAT3")], [F.EqualityFilter("category", "CAT3")]),
        ([H.NewEqualityFilter("category", "cat2")], [F.EqualityFilter("category", "cat2")]),
        ([H.NewEqualityFilter("price", 100)], [F.EqualityFilter("price", I(100))]),
        ([H.NewEqualityFilter("price", "100")], [F.EqualityFilter("price", "100")]),
        ([H.NewEqualityFilter("active", True)], [F.EqualityFilter("active", True)]),
        ([H.NewEqualityFilter("active", 1)], [F.EqualityFilter("active", I(1))]),
        ([H.NewEqualityFilter("price", None)], [F.EqualityFilter("price", None)]),
        ([H.NewRangeFilter("price", 10, 100.5, True, False)], [F.RangeFilter("price", I(10), Fl(100.5), True, False)]),
        ([H.NewRangeFilter("price", None, 50, True, True)], [F.RangeFilter("price", None, I(50), True, True)]),
        ([H.NewRangeFilter("price", 20.0, None, False, True)], [F.RangeFilter("price", Fl(20.0), None, False, True)]),
        ([H.NewRangeFilter("price", "10", 100, True, True)], [F.RangeFilter("price", "10", I(100), True, True)]),
        ([H.NewSetFilter("category", ["CAT0", "cat4", 3])], [F.SetFilter("category", ["CAT0", "cat4", I(3)])]),
        ([H.NewSetFilter("price", [1, 2.0, "3", True])], [F.SetFilter("price", [I(1), Fl(2.0), "3", True])]),
        ([H.NewSetFilter("active", [True])], [F.SetFilter("active", [True])]),
        ([H.NewSetFilter("category", [])], [F.SetFilter("category", [])]),
        ([H.NewExistsFilter("price", True)], [F.ExistsFilter("price", True)]),
        ([H.NewExistsFilter("price", False)], [F.ExistsFilter("price", False)]),
        ([H.NewExistsFilter("nested.name", True)], [F.ExistsFilter("nested.name", True)]),
        ([H.NewExistsFilter("absent", False)], [F.ExistsFilter("absent", False)]),
        ([H.NewEqualityFilter("nested.level", 2)], [F.EqualityFilter("nested.level", I(2))]),
        # array-valued facets: any element valuesEqual to any member (case-sensitive, numbers as float64)
        ([H.NewSetFilter("labels", ["red"])], [F.SetFilter("labels", ["red"])]),
        ([H.NewSetFilter("labels", ["RED", 2])], [F.SetFilter("labels", ["RED", I(2)])]),
        ([H.NewSetFilter("labels", [2.5, True, "3"])], [F.SetFilter("labels", [Fl(2.5), True, "3"])]),
        ([H.NewSetFilter("labels", [None])], [F.SetFilter("labels", [None])]),
        ([H.NewSetFilter("labels", [{"k": 1}])], [F.SetFilter("labels", [{"map": {"k": I(1)}}])]),
        ([H.NewSetFilter("labels", [{"k": 1.0}])], [F.SetFilter("labels", [{"map": {"k": Fl(1.0)}}])]),
        ([H.NewEqualityFilter("labels", "RED")], [F.EqualityFilter("labels", "RED")]),
        # a filter value that is itself an array / map: reflect.DeepEqual with the facet's value (facets.go:85) —
        # order- and type-sensitive (a Go int inside the literal never equals a decoded float64)
        ([H.NewEqualityFilter("labels", ["red"])], [F.EqualityFilter("labels", {"list": ["red"]})]),
        ([H.NewEqualityFilter("labels", ["Red"])], [F.EqualityFilter("labels", {"list": ["Red"]})]),
        ([H.NewEqualityFilter("labels", [])], [F.EqualityFilter("labels", {"list": []})]),
        ([H.NewEqualityFilter("labels", [1.0])], [F.EqualityFilter("labels", {"list": [Fl(1.0)]})]),
        ([H.NewEqualityFilter("labels", [1])], [F.EqualityFilter("labels", {"list": [I(1)]})]),
        ([H.NewEqualityFilter("labels", ["red", "blue"])], [F.EqualityFilter("labels", {"list": ["red", "blue"]})]),
        ([H.NewEqualityFilter("labels", ["blue", "red"])], [F.EqualityFilter("labels", {"list": ["blue", "red"]})]),
        ([H.NewEqualityFilter("labels", [None])], [F.EqualityFilter("labels", {"list": [None]})]),
        ([H.NewEqualityFilter("labels", [True, 2.5])], [F.EqualityFilter("labels", {"list": [True, Fl(2.5)]})]),
        ([H.NewEqualityFilter("labels", {"k": 1.0})], [F.EqualityFilter("labels", {"map": {"k": Fl(1.0)}})]),
        ([H.NewEqualityFilter("labels", {"k": 1})], [F.EqualityFilter("labels", {"map": {"k": I(1)}})]),
        ([H.NewEqualityFilter("labels", {"k": 0.0}), H.NewExistsFilter("category", True)],
         [F.EqualityFilter("labels", {"map": {"k": Fl(0.0)}}), F.ExistsFilter("category", True)]),
        ([H.NewEqualityFilter("category", ["cat3"])], [F.EqualityFilter("category", {"list": ["cat3"]})]),
        ([H.NewExistsFilter("labels", True)], [F.ExistsFilter("labels", True)]),
        ([H.NewRangeFilter("labels", 1, 3, True, True)], [F.RangeFilter("labels", I(1), I(3), True, True)]),
        ([H.NewSetFilter("labels", ["blue", 1]), H.NewEqualityFilter("category", "cat1")],
         [F.SetFilter("labels", ["blue", I(1)]), F.EqualityFilter("category", "cat1")]),
        ([H.NewEqualityFilter("category", "cat3"), H.NewRangeFilter("price", 0, 100, True, True),
          H.NewExistsFilter("active", True)],
         [F.EqualityFilter("category", "cat3"), F.RangeFilter("price", I(0), I(100), True, True),
          F.ExistsFilter("active", True)]),
    ]
    for hf, of in cases:
        got = c.filter_mask(facet_filters=hf)
        want = np.array(F.facet_mask(raw, fields, of))
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, ([(f.type, f.field, f.value_json) for f in hf], bad[:5], [rows[i] for i in bad[:3]])
    c.close()


def test_collection_search_with_filters(H, oracle):
    """Collection.Search / FluentSearch / SearchWithFacets: filter over the WHOLE ranking, not a top-k
    window (collection_test.go:549-594, collection_facets_test.go:522-556), ~10 % selectivity like
    BASELINE config 3, and deletes."""
    from oracle import filters as F
    from oracle import rerank
    rng = np.random.default_rng(13)
    n, d, k = 20000, 64, 10
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    ids = [f"v{i:06d}" for i in range(n)]
    rows = [{"category": f"cat{rng.integers(5)}", "tag": f"tag{rng.integers(20):02d}", "price": float(rng.integers(0, 1000))}
            for _ in range(n)]
    c = H.Collection("big", d, "cosine")
    c.AddBatch(ids, corpus, rows)
    c.SetFacetFields(["category", "tag", "price"])
    raw = [json.dumps(m) for m in rows]
    tags = [f"tag{t:02d}" for t in range(10)]
    q = rng.standard_normal(d).astype(np.float32)

    mask = np.array(F.metadata_mask(raw, [("category", "=", "cat3"), ("tag", "in", {"list": tags})]))
    assert 0.07 < mask.mean() < 0.13
    want = rerank.filtered_search(corpus, ids, q, k, 0, mask)
    got = c.FluentSearch(q).WithK(k).Filter("category", "cat3").FilterIn("tag", tags).Execute()
    assert [g[0] for g in got] == [w[0] for w in want]
    assert [np.float32(g[1]).view(np.uint32) for g in got] == [np.float32(w[1]).view(np.uint32) for w in want]

    got = c.SearchWithFacets(q, k, [H.NewEqualityFilter("category", "CAT3"), H.NewSetFilter("tag", tags)])
    assert [g[0] for g in got] == [w[0] for w in want]

    # a very selective filter: the match is far down the ranking
    rare = np.array(F.metadata_mask(raw, [("price", "=", {"int": 7}), ("category", "=", "cat1")]))
    want = rerank.filtered_search(corpus, ids, q, k, 0, rare)
    got = c.Search(q, k, [("price", "=", 7), ("category", "=", "cat1")])
    assert [g[0] for g in got] == [w[0] for w in want] and len(got) == min(k, int(rare.sum()))

    # deletes hide rows from filtered and unfiltered searches
    dead = [w[0] for w in rerank.filtered_search(corpus, ids, q, 3, 0, mask)]
    for id_ in dead:
        c.Delete(id_)
    live = np.ones(n, dtype=np.uint8)
    for id_ in dead:
        live[ids.index(id_)] = 0
    want = rerank.filtered_search(corpus, ids, q, k, 0, mask, live=live)
    got = c.FluentSearch(q).WithK(k).Filter("category", "cat3").FilterIn("tag", tags).Execute()
    assert [g[0] for g in got] == [w[0] for w in want]
    assert c.Count() == n - 3
    with pytest.raises(H.QuiverError, match="vector not found"):
        c.Delete(dead[0])
    with pytest.raises(H.QuiverError, match="top_k must be greater than 0"):
        c.Search(q, 0)
    with pytest.raises(H.QuiverError, match="invalid vector dimension: expected 64, got 3"):
        c.Search([1, 2, 3], 5)
    c.close()


def test_single_inserts_are_write_combined_and_visible_at_once(H, oracle):
    """Insert one vector at a time (Collection.Add -> Index.Insert, collection.go:178): the host layer buffers the
    rows and uploads them in one go before the next device read, so every Insert is visible to the very next
    Search, Delete or filter — interleaved here — and the final index equals one built with InsertBatch."""
    rng = np.random.default_rng(31)
    n, d, k = 6000, 24, 7
    corpus = rng.random((n, d), dtype=np.float32)
    ids = [f"s{i:05d}" for i in range(n)]
    idx = H.HybridIndex(d, "euclidean")
    live = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        idx.Insert(ids[i], corpus[i])
        live[i] = 1
        if i in (0, 1, 17, 4095, 4096, 4097, 5000):      # around the buffer capacity too
            res = idx.Search(corpus[i], 1)
            assert res[0][0] == ids[i] and res[0][1] == 0.0
        if i in (100, 4500):                              # delete a buffered row; a buffered id is a duplicate
            idx.Delete(ids[i - 1])
            live[i - 1] = 0
            assert idx.Size() == int(live.sum())
            with pytest.raises(H.QuiverError, match="already exists"):
                idx.Insert(ids[i], corpus[i])
    assert idx.Size() == int(live.sum())
    q = rng.random((5, d), dtype=np.float32)
    batch = H.HybridIndex(d, "euclidean")
    keep = np.nonzero(live)[0]
    batch.InsertBatchArrays([ids[i] for i in keep], corpus[keep])
    for qi in q:
        got = idx.Search(qi, k)
        od, orow = oracle.exact_search(corpus, qi, k, 1, 0, live)
        assert [g[0] for g in got] == [ids[r] for r in orow]
        assert [np.float32(g[1]).view(np.uint32) for g in got] == [x.view(np.uint32) for x in od]
        assert got == batch.Search(qi, k)
    idx.close()
    batch.close()


def test_concurrent_searches_while_inserting(H):
    """hybrid_stress_test.go:29-74 shape: goroutines searching a bare HybridIndex while another inserts. Searches
    share the lock, each Insert takes it exclusively and its buffered row is uploaded before the next search reads
    the device; every answer must be a valid, sorted list whose first hit is the query's own vector (inserted
    before the searchers start)."""
    import threading
    rng = np.random.default_rng(41)
    d, n0, n_more = 32, 3000, 3000
    base = rng.random((n0, d), dtype=np.float32)
    more = rng.random((n_more, d), dtype=np.float32)
    idx = H.HybridIndex(d, "euclidean")
    idx.InsertBatchArrays([f"b{i}" for i in range(n0)], base)
    errors = []
    stop = threading.Event()

    def searcher(seed):
        r = np.random.default_rng(seed)
        try:
            while not stop.is_set():
                i = int(r.integers(n0))
                res = idx.Search(base[i], 5)
                assert len(res) == 5 and res[0][0] == f"b{i}" and res[0][1] == 0.0
                assert all(res[j][1] <= res[j + 1][1] for j in range(4))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def inserter():
        try:
            for i in range(n_more):
                idx.Insert(f"m{i}", more[i])
                if i % 500 == 499:
                    idx.Delete(f"m{i - 1}")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=searcher, args=(s,)) for s in range(4)]
    for t in ts:
        t.start()
    w = threading.Thread(target=inserter)
    w.start()
    w.join()
    stop.set()
    for t in ts:
        t.join()
    assert not errors, errors[:3]
    assert idx.Size() == n0 + n_more - n_more // 500
    res = idx.Search(more[-1], 1)
    assert res[0][0] == f"m{n_more - 1}" and res[0][1] == 0.0
    idx.close()


def test_delete_batch_and_update(H, oracle):
    """HybridIndex.DeleteBatch (hybrid_index.go:293-375), Collection.DeleteBatch (collection.go:375-414) and
    Collection.Update (collection.go:417-466) with the reference's error text and all-or-nothing checks."""
    from oracle import filters as F
    from oracle import rerank
    rng = np.random.default_rng(51)
    n, d, k = 4000, 20, 6
    corpus = rng.random((n, d), dtype=np.float32)
    ids = [f"u{i:04d}" for i in range(n)]
    idx = H.HybridIndex(d, "cosine")
    idx.InsertBatchArrays(ids, corpus)
    with pytest.raises(H.QuiverError, match=r"some vectors not found: \[nope also-missing\]"):
        idx.DeleteBatch([ids[0], "nope", ids[1], "also-missing"])
    assert idx.Size() == n                               # nothing was deleted
    idx.DeleteBatch([])                                  # no-op
    gone = list(range(0, n, 3))
    idx.DeleteBatch([ids[i] for i in gone])
    live = np.ones(n, dtype=np.uint8)
    live[gone] = 0
    assert idx.Size() == int(live.sum())
    q = rng.random(d, dtype=np.float32)
    od, orow = oracle.exact_search(corpus, q, k, 0, 0, live)
    got = idx.Search(q, k)
    assert [g[0] for g in got] == [ids[r] for r in orow]
    assert [np.float32(g[1]).view(np.uint32) for g in got] == [x.view(np.uint32) for x in od]
    idx.close()

    rows = [{"category": f"cat{i % 4}", "price": float(i)} for i in range(n)]
    c = H.Collection("upd", d, "euclidean")
    c.AddBatch(ids, corpus, rows)
    with pytest.raises(H.QuiverError, match="vector not found: ghost"):
        c.DeleteBatch([ids[5], "ghost"])
    assert c.Count() == n
    c.DeleteBatch([ids[i] for i in range(10)])
    assert c.Count() == n - 10
    with pytest.raises(H.QuiverError, match="vector not found"):
        c.Update(ids[3], q)
    with pytest.raises(H.QuiverError, match="invalid vector dimension: expected 20, got 3"):
        c.Update(ids[100], [1, 2, 3])
    with pytest.raises(H.QuiverError, match="invalid metadata format"):
        c.Update(ids[100], None, "[1, 2]")
    # new vector, metadata kept: the id now sits at distance 0 from q and still passes its old filter
    c.Update(ids[100], q)
    got = c.Search(q, 1, [("category", "=", "cat0")])
    assert got[0][0] == ids[100] and got[0][1] == 0.0
    # new metadata only: the vector stays, the filter answer changes
    c.Update(ids[100], None, {"category": "moved", "price": -1.0})
    assert c.Search(q, 1, [("category", "=", "moved")])[0][0] == ids[100]
    assert c.Search(q, 1, [("category", "=", "cat0")])[0][0] != ids[100]
    # both at once, then the whole ranking against the oracle
    v2 = rng.random(d, dtype=np.float32)
    c.Update(ids[200], v2, {"category": "cat1", "price": 200.5})
    corpus2 = corpus.copy()
    corpus2[100] = q
    corpus2[200] = v2
    rows2 = list(rows)
    rows2[100] = {"category": "moved", "price": -1.0}
    rows2[200] = {"category": "cat1", "price": 200.5}
    live = np.ones(n, dtype=np.uint8)
    live[:10] = 0
    mask = np.array(F.metadata_mask([json.dumps(m) for m in rows2], [("category", "=", "cat1")]))
    want = rerank.filtered_search(corpus2, ids, v2, k, 1, mask, live=live)
    got = c.Search(v2, k, [("category", "=", "cat1")])
    assert sorted(g[0] for g in got) == sorted(w[0] for w in want) and got[0][0] == ids[200]
    assert [np.float32(g[1]).view(np.uint32) for g in got] == [np.float32(w[1]).view(np.uint32) for w in want]
    assert c.Count() == n - 10
    c.Compact()
    assert [g[0] for g in c.Search(v2, k, [("category", "=", "cat1")])] == [g[0] for g in got]
    c.close()


def test_update_batch(H, oracle):
    """Collection.UpdateBatch (collection.go:469-529): validation first, then every vector replaced."""
    rng = np.random.default_rng(61)
    n, d, k = 3000, 16, 5
    corpus = rng.random((n, d), dtype=np.float32)
    ids = [f"w{i:04d}" for i in range(n)]
    rows = [{"category": f"cat{i % 3}"} for i in range(n)]
    c = H.Collection("ub", d, "euclidean")
    c.AddBatch(ids, corpus, rows)
    with pytest.raises(H.QuiverError, match="no vectors provided for batch update"):
        c.UpdateBatch([], np.zeros((0, d), dtype=np.float32))
    with pytest.raises(H.QuiverError, match="vector ID cannot be empty"):
        c.UpdateBatch([ids[0], ""], corpus[:2])
    with pytest.raises(H.QuiverError, match="vector not found: missing"):
        c.UpdateBatch([ids[0], "missing"], corpus[:2])
    with pytest.raises(H.QuiverError, match=f"invalid vector dimension for vector {ids[0]}: expected 16, got 4"):
        c.UpdateBatch([ids[0]], np.zeros((1, 4), dtype=np.float32))
    with pytest.raises(H.QuiverError, match=f"invalid metadata format for vector {ids[1]}"):
        c.UpdateBatch([ids[0], ids[1]], corpus[:2], [None, "[]"])
    assert c.Count() == n
    upd = list(range(0, n, 7))
    new = rng.random((len(upd), d), dtype=np.float32)
    mds = [None if j % 2 else {"category": "fresh"} for j in range(len(upd))]
    c.UpdateBatch([ids[i] for i in upd], new, mds)
    assert c.Count() == n
    corpus2 = corpus.copy()
    corpus2[upd] = new
    q = rng.random(d, dtype=np.float32)
    od, orow = oracle.exact_search(corpus2, q, k, 1, 0, None)
    got = c.Search(q, k)
    assert sorted(g[0] for g in got) == sorted(ids[r] for r in orow)
    assert [np.float32(g[1]).view(np.uint32) for g in got] == [x.view(np.uint32) for x in od]
    fresh = {ids[upd[j]] for j in range(len(upd)) if j % 2 == 0}
    got = c.Search(q, len(fresh) + 5, [("category", "=", "fresh")])
    assert {g[0] for g in got} == fresh
    kept = c.Search(new[1], 1, [("category", "=", rows[upd[1]]["category"])])   # metadata kept for odd entries
    assert kept[0][0] == ids[upd[1]] and kept[0][1] == 0.0
    # the same id twice in one batch: applied in order, the last vector wins
    c.UpdateBatch([ids[1], ids[1]], np.stack([corpus[2], q]))
    assert c.Search(q, 1)[0][0] == ids[1] and c.Count() == n
    c.close()
