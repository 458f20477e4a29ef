"""Batched searches behind a selective filter run over a dense VIEW of the passing rows (api.cu:
build_filter_view — one ordered gather, cached in the filter handle — then the tensor-core scan), instead of
many gather passes of the flat scan. Results must equal the oracle restricted to the passing rows, bit for
bit, including tie order, tombstones, row-sharded keys and the cache's invalidation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STRING = 1 << 2


def _index_with_category(capi, corpus, metric, cat):
    n = corpus.shape[0]
    idx = capi.Index(corpus.shape[1], metric)
    idx.upload(corpus)
    kind = np.full(n, 2, dtype=np.uint8) | np.uint8(0x80)
    idx.set_column(0, kind, np.zeros(n), cat.astype(np.int32), cat.astype(np.int32))
    return idx


def _eq(capi, idx, code):
    return capi.Filter(idx, [capi.qg_pred(0, 1, 0, 1)], [capi.qg_clause(7, 0, 0, int(code), STRING, 0, 0.0, 0.0)])


@pytest.mark.parametrize("metric,d", [(1, 128), (0, 96), (1, 768), (2, 64)])
def test_filtered_batches_over_the_view(capi, oracle, metric, d):
    rng = np.random.default_rng(d + metric)
    n, nq, k = 120_000, 70, 10
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    corpus[500:540] = corpus[100]  # exact duplicates: ties must resolve by row as over the whole index
    cat = rng.integers(0, 10, n)
    cat[100] = cat[500:540] = 3
    queries = np.concatenate([rng.standard_normal((nq - 1, d)).astype(np.float32), corpus[100:101]])
    idx = _index_with_category(capi, corpus, metric, cat)
    flt = _eq(capi, idx, 3)
    live = (cat == 3).astype(np.uint8)
    dist, row, cnt, _ = idx.search(queries, k, filter=flt)
    st = idx.stats()
    assert st["path"] == 3 and st["rows_scanned"] == int(live.sum()), st  # the dense view, not the whole index
    for i in (0, 1, nq // 2, nq - 1):
        od, orow = oracle.exact_search(corpus, queries[i], k, metric, 0, live)
        assert cnt[i] == len(od) and np.array_equal(row[i], orow), (i, row[i], orow)
        assert np.array_equal(dist[i].view(np.uint32), od.view(np.uint32))
    # cached view, second batch; then a tombstone invalidates it
    d2, r2, _, _ = idx.search(queries, k, filter=flt)
    assert np.array_equal(r2, row) and np.array_equal(d2.view(np.uint32), dist.view(np.uint32))
    dead = np.unique(row[:, :3].ravel())
    idx.tombstone(dead)
    live[dead] = 0
    d3, r3, c3, _ = idx.search(queries, k, filter=flt)
    for i in (0, nq - 1):
        od, orow = oracle.exact_search(corpus, queries[i], k, metric, 0, live)
        assert np.array_equal(r3[i, :len(orow)], orow) and np.array_equal(d3[i, :len(od)].view(np.uint32), od.view(np.uint32))
    # exhaustive GPU oracle agrees for every query of the batch
    xd, xr, xc = idx.search_exhaustive(queries, k, filter=flt)
    assert np.array_equal(r3, xr) and np.array_equal(d3.view(np.uint32), xd.view(np.uint32))
    # a single query rides the cached view; behind a filter that has none yet it takes the flat gather scan
    # (building a view for one query costs more than gathering the passing rows once)
    d1, r1, c1, _ = idx.search(queries[:1], k, filter=flt)
    assert idx.stats()["path"] == (3 if d <= 512 else 2)  # the bf16 copy (d <= 512) serves single queries too
    assert np.array_equal(r1[0], r3[0]) and np.array_equal(d1[0].view(np.uint32), d3[0].view(np.uint32))
    flt2 = _eq(capi, idx, 4)
    d4, r4, c4, _ = idx.search(queries[:1], k, filter=flt2)
    assert idx.stats()["path"] == 2
    live4 = (cat == 4).astype(np.uint8)
    live4[dead] = 0
    od, orow = oracle.exact_search(corpus, queries[0], k, metric, 0, live4)
    assert np.array_equal(r4[0, :len(orow)], orow) and np.array_equal(d4[0, :len(od)].view(np.uint32), od.view(np.uint32))
    flt.close()
    idx.close()


def test_view_with_negatives_and_shard_keys(capi, oracle):
    import torch
    rng = np.random.default_rng(4)
    n, d, nq, k = 90_000, 64, 48, 12
    corpus = rng.random((n, d), dtype=np.float32)
    cat = rng.integers(0, 8, n)
    queries = rng.random((nq, d), dtype=np.float32)
    negs = rng.random((nq, d), dtype=np.float32)
    idx = _index_with_category(capi, corpus, 1, cat)
    flt = _eq(capi, idx, 5)
    live = (cat == 5).astype(np.uint8)
    dist, row, cnt, negd = idx.search(queries, k, filter=flt, negatives=negs)
    assert idx.stats()["path"] == 3
    for i in (0, nq - 1):
        od, orow = oracle.exact_search(corpus, queries[i], k, 1, 0, live)
        assert np.array_equal(row[i], orow)
        want = np.array([oracle.distance(1, corpus[r], negs[i]) for r in orow], dtype=np.float32)
        assert np.array_equal(negd[i].view(np.uint32), want.view(np.uint32))
    # row-sharded keys: global row = row_base + index row, through the view
    dev = torch.device("cuda:0")
    dq = torch.from_numpy(queries).to(dev)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    base = 1_000_000
    idx.search_shard_keys_device(dq.data_ptr(), nq, k, base, keys.data_ptr(), filter=flt)
    torch.cuda.synchronize()
    kk = keys.cpu().numpy().view(np.uint64)
    assert np.array_equal((kk & np.uint64(0xFFFFFFFF)).astype(np.int64), row + base)
    flt.close()
    idx.close()
