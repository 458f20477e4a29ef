"""The full-size parity chain of SURVEY 7: (1) the exhaustive GPU path (qg_search_exhaustive: every row's
exact distance in the reference's arithmetic + full sort, exact.go:114-129 verbatim) equals the CPU oracle on
corpora the CPU finishes; (2) the fast regimes (flat fp32 scan, bf16 tensor-core scan) equal the exhaustive GPU
path at 1M rows on every synthetic kind, including the real-valued ones where the bf16 copy rounds every
element. tests/config_bench.py repeats step (2) at 10M x 768 and 100M x 96."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", [0, 1, 2, 3, 4])
def test_exhaustive_gpu_path_equals_the_cpu_oracle(capi, oracle, metric):
    n, d, k = 150_000, 96, 25
    kind = 2 if metric in (0, 2) else 0
    corpus = oracle.synth(kind, 42, 0, n, d)
    queries = oracle.synth(kind, 9999, 0, 6, d)
    idx = capi.Index(d, metric)
    idx.upload(corpus)
    dead = np.arange(0, n, 7)
    idx.tombstone(dead)
    live = np.ones(n, dtype=np.uint8)
    live[dead] = 0
    xd, xr, xc = idx.search_exhaustive(queries, k)
    od, orow, ocnt = oracle.exact_search_batch(corpus, queries, k, metric, live=live, threads=6)
    assert np.array_equal(xc, ocnt.astype(np.int32))
    assert np.array_equal(xr, orow)
    assert np.array_equal(xd.view(np.uint32), od.view(np.uint32))
    idx.close()


def test_exhaustive_gpu_path_edges(capi):
    idx = capi.Index(8, capi.L2)
    d, r, c = idx.search_exhaustive(np.zeros((2, 3), dtype=np.float32), 0)  # empty index: no error (exact.go:96-99)
    assert c.tolist() == [0, 0]
    idx.upload(np.arange(40, dtype=np.float32).reshape(5, 8))
    d, r, c = idx.search_exhaustive(np.zeros((1, 8), dtype=np.float32), 9)
    assert c[0] == 5 and r[0, :5].tolist() == [0, 1, 2, 3, 4] and r[0, 5] == -1 and np.isinf(d[0, 5])
    with pytest.raises(capi.QuiverGpuError) as e:
        idx.search_exhaustive(np.zeros((1, 7), dtype=np.float32), 3)
    assert e.value.code == capi.QG_ERR_DIM
    with pytest.raises(capi.QuiverGpuError) as e:
        idx.search_exhaustive(np.zeros((1, 8), dtype=np.float32), 0)
    assert e.value.code == capi.QG_ERR_K
    idx.close()


@pytest.mark.parametrize("kind,metric,dim", [(0, 1, 128), (1, 1, 128), (2, 1, 128), (2, 0, 128), (3, 1, 96), (0, 2, 128)])
def test_fast_regimes_equal_the_exhaustive_gpu_path_at_1m_rows(capi, oracle, kind, metric, dim):
    n, k = 1_000_000, 10
    idx = capi.Index(dim, metric, reserve_rows=n)
    idx.upload_synthetic(kind, 42, 0, n)
    queries = oracle.synth(kind, 9999, 0, 300, dim)
    # tensor-core regime: a batch of 300 (two passes), all queries compared with the GPU oracle for 40 of them
    dist, row, cnt, _ = idx.search(queries, k)
    st = idx.stats()
    assert st["path"] == 3, st
    ids = sorted(set(int(x) for x in np.linspace(0, 299, 40)))
    xd, xr, xc = idx.search_exhaustive(queries[ids], k)
    assert (cnt == k).all() and (xc == k).all()
    assert np.array_equal(row[ids], xr), st
    assert np.array_equal(dist[ids].view(np.uint32), xd.view(np.uint32))
    # single queries ride the bf16 copy too
    for j, i in enumerate(ids[:6]):
        d1, r1, c1, _ = idx.search(queries[i:i + 1], k)
        assert idx.stats()["path"] == 3
        assert np.array_equal(r1[0], xr[j]) and np.array_equal(d1[0].view(np.uint32), xd[j].view(np.uint32))
    idx.close()
    # flat fp32 scan (dense kernel): single queries and blocks of two on an index without the bf16 copy
    flat = capi.Index(dim, metric, reserve_rows=n, flags=capi.FLAG_NO_BF16_COPY)
    flat.upload_synthetic(kind, 42, 0, n)
    for j, i in enumerate(ids[:6]):
        d1, r1, c1, _ = flat.search(queries[i:i + 1], k)
        assert flat.stats()["path"] == 1
        assert np.array_equal(r1[0], xr[j]) and np.array_equal(d1[0].view(np.uint32), xd[j].view(np.uint32))
    pair = [ids[0], ids[1]]
    d2, r2, c2, _ = flat.search(queries[pair], k)
    assert flat.stats()["path"] == 1 and flat.stats()["queries_per_pass"] == 2
    assert np.array_equal(r2, xr[:2]) and np.array_equal(d2.view(np.uint32), xd[:2].view(np.uint32))
    flat.close()
