"""CPU restatement of the exact branch of HybridIndex.searchWithStrategy with a negative example
(pkg/hybrid/hybrid_index.go:515-570) and of Collection.Search's scan-until-k filter stage
(pkg/core/collection.go:679-752). TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import numpy as np

from . import cref


def exact_search_ids(corpus, ids, q, k, metric, arith=cref.ARITH_VECTORTYPES, live=None):
    """ExactIndex.Search (exact.go:92-133) with string IDs. Ties (which the reference leaves in
    map-iteration order) are broken by row index, like every other part of the oracle."""
    dist, row = cref.exact_search(corpus, q, k, metric, arith, live)
    return [(ids[int(r)], np.float32(d), int(r)) for d, r in zip(dist, row)]


def search_with_negative(corpus, ids, q, k, metric, negative, weight, arith=cref.ARITH_VECTORTYPES, live=None):
    """hybrid_index.go:515-570: fetch max(2k, 30) clamped to N, score = d - w * dist(vec, neg) in
    float32 (Go/amd64 never fuses the multiply-subtract), stable sort by (score, ID string), first k.
    The returned Distance IS the adjusted score (:552)."""
    n_live = corpus.shape[0] if live is None else int(np.count_nonzero(live))
    retrieve_k = max(2 * k, 30)
    retrieve_k = min(retrieve_k, n_live)
    if retrieve_k <= 0:
        return []
    base = exact_search_ids(corpus, ids, q, retrieve_k, metric, arith, live)
    w = np.float32(weight)
    rer = []
    for pos, (id_, d, row) in enumerate(base):
        nd = cref.distance(metric, corpus[row], negative, arith)  # distFunc(vector, negExample) :544
        prod = np.float32(w * np.float32(nd))
        score = np.float32(np.float32(d) - prod)
        rer.append((id_, score, row, pos))
    # sort.SliceStable with less(i,j) = score equal ? ID< : score<   (:555-560); NaN scores compare
    # false both ways and keep their position, which a stable merge sort reproduces for the
    # non-NaN inputs the tests use.
    import functools

    def cmp(a, b):
        if a[1] == b[1]:
            return -1 if a[0] < b[0] else (1 if a[0] > b[0] else 0)
        return -1 if a[1] < b[1] else 1

    rer.sort(key=functools.cmp_to_key(cmp))
    return [(id_, score, row) for id_, score, row, _ in rer[:k]]


def filtered_search(corpus, ids, q, k, metric, mask, arith=cref.ARITH_VECTORTYPES, live=None):
    """collection.go:679-752 / :1177-1204: rank ALL live rows (searchK = Index.Size()), walk the
    ranking keeping rows whose mask bit is set, stop at k."""
    lv = np.ones(corpus.shape[0], dtype=np.uint8) if live is None else np.asarray(live, dtype=np.uint8).copy()
    n_live = int(lv.sum())
    if n_live == 0:
        return []
    full = exact_search_ids(corpus, ids, q, n_live, metric, arith, lv)
    out = []
    for id_, d, row in full:
        if mask[row]:
            out.append((id_, d, row))
            if len(out) >= k:
                break
    return out
