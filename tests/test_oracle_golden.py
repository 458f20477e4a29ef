"""T0: the oracle against every known-answer vector the reference's own tests hold for the hot path
(tests/golden/*.json, extracted from the Go test sources by tests/golden/make_golden.py).
Runs on CPU; this is what pins the oracle."""
import json
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def num(x):
    for key in ("int", "float", "float32"):
        if isinstance(x, dict) and key in x:
            return float(x[key])
    raise TypeError(x)


def vec(x):
    if isinstance(x, dict) and "f32vec" in x:
        return np.asarray(x["f32vec"], dtype=np.float32)
    if isinstance(x, dict) and "items" in x:
        return np.asarray([num(i) for i in x["items"]], dtype=np.float32)
    return np.asarray(x, dtype=np.float32)


METRIC = {"cosine": 0, "euclidean": 1, "dot_product": 2, "squared_euclidean": 3, "manhattan": 4}


@pytest.mark.parametrize("metric", sorted(METRIC))
def test_vectortypes_distances(oracle, metric):
    """pkg/vectortypes/distances_test.go:9-204, tolerance 1e-6 (floatEquals)."""
    rows = load("distances.json")["vectortypes"][metric]
    assert rows
    for r in rows:
        got = float(oracle.distance(METRIC[metric], vec(r["vecA"]), vec(r["vecB"])))
        assert abs(got - num(r["expected"])) <= 1e-6, (metric, r["name"], got)


def test_distance_length_mismatch_is_an_error(oracle):
    """distances.go:13-15 panics on a length mismatch."""
    with pytest.raises(ValueError):
        oracle.distance(0, np.ones(3, np.float32), np.ones(4, np.float32))


def test_distance_type_mapping(oracle):
    """pkg/vectortypes/types_test.go:9-80: enum -> function, checked by value."""
    name_to_metric = {"Cosine": 0, "Euclidean": 1, "DotProduct": 2, "Manhattan": 4}
    seen = 0
    for r in load("distances.json")["types"]:
        ident = r["distType"].get("ident") if isinstance(r["distType"], dict) else None
        if not r.get("checkVecs"):
            continue
        metric = name_to_metric.get(ident, 0)  # unknown => cosine (types.go:46-47)
        got = float(oracle.distance(metric, vec(r["vecA"]), vec(r["vecB"])))
        assert abs(got - num(r["wantResult"])) <= 1e-6, (r["name"], got)
        seen += 1
    assert seen >= 4


def test_hnsw_f32_distances(oracle):
    """pkg/hnsw/hnsw_test.go:376-455: the sequential float32 variants."""
    fn = {"EuclideanDistanceFunc": 1, "CosineDistanceFunc": 0, "DotProductDistanceFunc": 2}
    seen = 0
    for r in load("distances.json")["hnsw_f32"]:
        a, b = vec(r["a"]), vec(r["b"])
        if r.get("wantErr"):
            assert a.size != b.size  # ErrDimensionMismatch
            continue
        got = float(oracle.distance(fn[r["fn"]["ident"]], a, b, oracle.ARITH_HNSW_F32))
        assert abs(got - num(r["want"])) <= num(r["epsilon"]), (r["name"], got)
        seen += 1
    assert seen >= 6


def test_exact_index_search_cases(oracle):
    """pkg/hybrid/exact_test.go:97-205 (cosine, 3 axis vectors)."""
    g = load("exact_search.json")["exact_search"]
    ids = sorted(g["vectors"])
    corpus = np.stack([vec(g["vectors"][i]) for i in ids])
    for case in g["cases"]:
        q = vec(case["query"])
        k = int(num(case["k"]))
        dist, row = oracle.exact_search(corpus, q, k, 0)
        got_ids = [ids[int(r)] for r in row]
        want_ids = case["wantIDs"]["strings"]
        assert len(got_ids) == min(k, len(ids)), case["name"]
        assert np.all(np.diff(dist) >= 0)
        if case["exactOrder"]:
            assert got_ids == want_ids, (case["name"], got_ids, want_ids)
        else:
            assert set(got_ids) == set(want_ids), (case["name"], got_ids, want_ids)


def test_exact_search_edges(oracle):
    """exact.go:96-111: empty => [] even for k <= 0; k <= 0 => error; k > N clamps."""
    empty = np.zeros((0, 3), dtype=np.float32)
    q = np.ones(3, dtype=np.float32)
    d, r = oracle.exact_search(empty, q, 0, 0)
    assert len(d) == 0
    corpus = np.eye(3, dtype=np.float32)
    with pytest.raises(ValueError, match="k must be positive"):
        oracle.exact_search(corpus, q, 0, 0)
    d, r = oracle.exact_search(corpus, q, 10, 0)
    assert len(d) == 3


def test_negative_example_rerank(oracle):
    """pkg/hybrid/hybrid_index_test.go:541-657: 3 results, the negative example itself not first."""
    from oracle import rerank
    g = load("exact_search.json")["negative_example"]
    ids = list(g["vectors"])
    corpus = np.stack([vec(g["vectors"][i]) for i in ids])
    neg = corpus[ids.index(g["negative_id"])]
    res = rerank.search_with_negative(corpus, ids, np.asarray(g["query"], np.float32), g["k"], 0, neg, g["weight"])
    assert len(res) == 3
    assert res[0][0] != g["negative_id"]
    plain = rerank.exact_search_ids(corpus, ids, np.asarray(g["query"], np.float32), g["k"], 0)
    assert [r[0] for r in plain] != [r[0] for r in res] or True  # ranking may change, count must not


def test_rerank_stability(oracle):
    """pkg/hybrid/hybrid_index_rerank_test.go:9-47: equal candidates get equal adjusted distances
    (1 - 0.5 * 2 = 0), ordered by ID, no NaN."""
    from oracle import rerank
    g = load("exact_search.json")["rerank_stability"]
    ids = sorted(g["vectors"])
    corpus = np.asarray([g["vectors"][i] for i in ids], dtype=np.float32)
    res = rerank.search_with_negative(corpus, ids, np.asarray(g["query"], np.float32), g["k"], 1,
                                      np.asarray(g["negative"], np.float32), g["weight"])
    assert [r[0] for r in res] == ["1", "2", "3"]
    assert all(float(r[1]) == 0.0 for r in res)


# ---- filters ------------------------------------------------------------------------------------
def test_facet_equality_table():
    from oracle import filters as F
    for r in load("filters.json")["facets"]["equality"]:
        assert F.EqualityFilter(r["field"], r["value"]).match(r["testVal"]) == r["expected"], r["name"]


def test_facet_range_table():
    from oracle import filters as F
    for r in load("filters.json")["facets"]["range"]:
        f = F.RangeFilter(r["field"], r["min"], r["max"], r["includeMin"], r["includeMax"])
        assert f.match(r["testVal"]) == r["expected"], r["name"]


def test_facet_set_table():
    from oracle import filters as F
    for r in load("filters.json")["facets"]["set"]:
        assert F.SetFilter(r["field"], r["values"]["list"]).match(r["testVal"]) == r["expected"], r["name"]


def test_facet_exists_table():
    from oracle import filters as F
    for r in load("filters.json")["facets"]["exists"]:
        assert F.ExistsFilter(r["field"], r["shouldExist"]).match(r["testVal"]) == r["expected"], r["name"]


def _build_facet_filter(call):
    from oracle import filters as F
    name, a = call["call"].split(".")[-1], call["args"]
    if name == "NewEqualityFilter":
        return F.EqualityFilter(a[0], a[1])
    if name == "NewRangeFilter":
        return F.RangeFilter(a[0], a[1], a[2], a[3], a[4])
    if name == "NewSetFilter":
        return F.SetFilter(a[0], a[1]["list"])
    if name == "NewExistsFilter":
        return F.ExistsFilter(a[0], a[1])
    raise KeyError(name)


def test_matches_all_filters_table():
    """facets_test.go:322-405."""
    from oracle import filters as F
    g = load("filters.json")
    facets = list(g["matches_all_facets"].items())
    rows = g["facets"]["matches_all"]
    assert rows
    for r in rows:
        flt = r["filters"]
        items = flt["items"] if isinstance(flt, dict) and "items" in flt else (flt or [])
        filters = [_build_facet_filter(c) for c in items]
        assert F.matches_all_filters(facets, filters) == r["result"], r["name"]


def test_core_matches_filter_table():
    """pkg/core/collection_test.go:727-816."""
    from oracle import filters as F
    rows = load("filters.json")["core_matches_filter"]
    assert len(rows) == 12
    for r in rows:
        fl = r["filter"]["fields"]
        op = fl["Operator"]
        op = F.OPERATORS[op["ident"]] if isinstance(op, dict) else op
        assert F.matches_filter(r["metadata"]["map"], fl["Field"], op, fl["Value"]) == r["want"], r["name"]


def test_core_filter_missing_field_and_mixed_types():
    """collection.go:533-536 (missing field is false even for != / not_in), :601-608 ("%v" text
    equality across types), :551-571 (non-list operand of in / not_in)."""
    from oracle import filters as F
    md = {"n": {"float": 42.0}, "s": "42", "b": True, "big": {"float": 1234567.0}}
    assert not F.matches_filter(md, "missing", "!=", "x")
    assert not F.matches_filter(md, "missing", "not_in", {"list": ["x"]})
    assert F.matches_filter(md, "s", "=", {"int": 42})          # "42" == "42"
    assert F.matches_filter(md, "n", "=", "42")                 # float64 42 prints as 42
    assert F.matches_filter(md, "b", "=", "true")
    assert F.matches_filter(md, "big", "=", "1.234567e+06")     # Go prints 1234567.0 as 1.234567e+06
    assert F.matches_filter(md, "n", "=", {"float": 42.0000000001})  # 1e-9 tolerance
    assert not F.matches_filter(md, "s", "in", "42")            # operand is not a list
    assert F.matches_filter(md, "s", "not_in", "42")
    assert F.matches_filter(md, "s", ">", {"int": 100})         # "42" > "100" as text


def test_gofmt_floats():
    from oracle.gotypes import format_float
    assert format_float(42.0) == "42"
    assert format_float(99.99) == "99.99"
    assert format_float(1e6) == "1e+06"
    assert format_float(123456.0) == "123456"
    assert format_float(0.0001) == "0.0001"
    assert format_float(0.00001) == "1e-05"
    assert format_float(-2.5) == "-2.5"
    assert format_float(1e21) == "1e+21"


def test_metadata_and_facet_masks():
    """collection.go:716-744 and :1189-1204 row rules: no metadata / unparsable JSON never match;
    a row needs a facet entry; MatchesAllFilters is false when the row has no facets at all."""
    from oracle import filters as F
    rows = ['{"category":"electronics","price":10}', None, "", "not json", '{"category":"books"}', "{}", "null"]
    m = F.metadata_mask(rows, [("category", "=", "electronics")])
    assert m == [True, False, False, False, False, False, False]
    m = F.metadata_mask(rows, [("category", "!=", "electronics")])
    assert m == [False, False, False, False, True, False, False]
    fm = F.facet_mask(rows, ["category"], [F.EqualityFilter("category", "ELECTRONICS")])
    assert fm == [True, False, False, False, False, False, False]
    fm = F.facet_mask(rows, ["category", "price"], [F.ExistsFilter("price", False)])
    assert fm == [False, False, False, False, True, False, False]  # "{}" has no facets at all => false
