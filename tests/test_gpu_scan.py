"""Parity of the CUDA exact-search path (through the C ABI) against the CPU oracle.

Bar: distances bit-identical to the oracle (which restates the reference's float64 /
float32 arithmetic), row lists identical under the shared (distance, row) tie rule.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

METRICS = {"cosine": 0, "l2": 1, "dot": 2, "sql2": 3, "l1": 4}


def _check_against_oracle(oracle, idx, corpus, queries, k, metric, arith=0, live=None, label=""):
    dist, row, cnt, _ = idx.search(queries, k)
    for i in range(queries.shape[0]):
        od, orow = oracle.exact_search(corpus, queries[i], k, metric, arith, live)
        n = len(od)
        assert cnt[i] == n, f"{label} q{i}: count {cnt[i]} != oracle {n}"
        if not (np.array_equal(row[i, :n], orow) and np.array_equal(dist[i, :n].view(np.uint32), od.view(np.uint32))):
            bad = np.nonzero((row[i, :n] != orow) | (dist[i, :n] != od))[0][:5]
            raise AssertionError(
                f"{label} q{i}: mismatch at {bad}: gpu rows {row[i, bad]} d {dist[i, bad]!r} vs oracle rows "
                f"{orow[bad]} d {od[bad]!r}; stats {idx.stats()}")
        assert np.all(row[i, n:] == -1) and np.all(np.isinf(dist[i, n:]))
        assert np.all(np.diff(dist[i, :n]) >= 0)


@pytest.mark.parametrize("dim", [3, 4, 64, 96, 100, 128, 768])
@pytest.mark.parametrize("metric", ["cosine", "l2", "dot"])
def test_parity_dims_metrics(capi, oracle, dim, metric):
    rng = np.random.default_rng(dim * 7 + METRICS[metric])
    n = 6000 if dim <= 128 else 2500
    corpus = rng.random((n, dim), dtype=np.float32)
    queries = rng.random((5, dim), dtype=np.float32)
    idx = capi.Index(dim, METRICS[metric])
    idx.upload(corpus)
    _check_against_oracle(oracle, idx, corpus, queries, 10, METRICS[metric], label=f"d{dim}/{metric}")
    idx.close()


@pytest.mark.parametrize("metric", ["sql2", "l1"])
@pytest.mark.parametrize("dim", [4, 96, 128])
def test_parity_minor_metrics(capi, oracle, dim, metric):
    rng = np.random.default_rng(11 + dim)
    corpus = rng.standard_normal((3000, dim)).astype(np.float32)
    queries = rng.standard_normal((3, dim)).astype(np.float32)
    idx = capi.Index(dim, METRICS[metric])
    idx.upload(corpus)
    _check_against_oracle(oracle, idx, corpus, queries, 10, METRICS[metric], label=f"d{dim}/{metric}")
    idx.close()


@pytest.mark.parametrize("metric", ["cosine", "l2", "dot"])
def test_parity_hnsw_f32_arith(capi, oracle, metric):
    rng = np.random.default_rng(5)
    corpus = rng.standard_normal((4000, 128)).astype(np.float32)
    queries = rng.standard_normal((4, 128)).astype(np.float32)
    idx = capi.Index(128, METRICS[metric], arith=capi.ARITH_HNSW_F32)
    idx.upload(corpus)
    _check_against_oracle(oracle, idx, corpus, queries, 10, METRICS[metric], arith=1, label=f"f32/{metric}")
    idx.close()


@pytest.mark.parametrize("k", [1, 10, 30, 100, 200, 1000])
def test_parity_k(capi, oracle, k):
    rng = np.random.default_rng(k)
    corpus = rng.random((20000, 64), dtype=np.float32)
    queries = rng.random((3, 64), dtype=np.float32)
    idx = capi.Index(64, 1)
    idx.upload(corpus)
    _check_against_oracle(oracle, idx, corpus, queries, k, 1, label=f"k{k}")
    idx.close()


@pytest.mark.parametrize("nq", [1, 2, 3, 7, 8, 9, 32, 100])
def test_parity_query_batches(capi, oracle, nq):
    rng = np.random.default_rng(100 + nq)
    corpus = rng.random((10000, 128), dtype=np.float32)
    queries = rng.random((nq, 128), dtype=np.float32)
    for metric in (0, 1):
        idx = capi.Index(128, metric)
        idx.upload(corpus)
        _check_against_oracle(oracle, idx, corpus, queries, 10, metric, label=f"Q{nq}/m{metric}")
        idx.close()


def test_k_larger_than_n_and_tiny_indices(capi, oracle):
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 31, 33, 257):
        corpus = rng.random((n, 16), dtype=np.float32)
        queries = rng.random((2, 16), dtype=np.float32)
        idx = capi.Index(16, 0)
        idx.upload(corpus)
        _check_against_oracle(oracle, idx, corpus, queries, 10, 0, label=f"n{n}")
        _check_against_oracle(oracle, idx, corpus, queries, 2000, 0, label=f"n{n}/k2000")  # exhaustive path
        idx.close()


def test_large_k_exhaustive_path(capi, oracle):
    rng = np.random.default_rng(8)
    corpus = rng.random((5000, 32), dtype=np.float32)
    queries = rng.random((2, 32), dtype=np.float32)
    idx = capi.Index(32, 1)
    idx.upload(corpus)
    _check_against_oracle(oracle, idx, corpus, queries, 5000, 1, label="k=N")  # Collection.Search asks k = Size()
    idx.close()


def test_duplicates_and_ties(capi, oracle):
    """exact.go:124 leaves ties unordered; both sides use (distance, row). All-zero vectors is the
    reference's own benchmark corpus (hybrid/benchmark_test.go:10-26)."""
    corpus = np.zeros((1000, 64), dtype=np.float32)
    q = np.zeros((1, 64), dtype=np.float32)
    for metric in (0, 1, 2):
        idx = capi.Index(64, metric)
        idx.upload(corpus)
        _check_against_oracle(oracle, idx, corpus, q, 10, metric, label=f"zeros/m{metric}")
        idx.close()
    rng = np.random.default_rng(2)
    base = rng.random((50, 128), dtype=np.float32)
    corpus = np.repeat(base, 40, axis=0)  # every vector 40 times
    queries = base[:3] + 0.001
    idx = capi.Index(128, 1)
    idx.upload(corpus)
    _check_against_oracle(oracle, idx, corpus, queries.astype(np.float32), 10, 1, label="dups")
    _check_against_oracle(oracle, idx, corpus, queries.astype(np.float32), 100, 1, label="dups/k100")
    idx.close()


def test_near_duplicate_neighbours_cosine(capi, oracle):
    """fp32 cosine is off by percents for near-duplicates; the exact re-rank must
    still return the reference's float64 values."""
    rng = np.random.default_rng(9)
    q = rng.standard_normal((1, 768)).astype(np.float32)
    corpus = rng.standard_normal((3000, 768)).astype(np.float32)
    corpus[100:160] = q + 1e-2 * rng.standard_normal((60, 768)).astype(np.float32)
    for metric in (0, 2, 1):
        idx = capi.Index(768, metric)
        idx.upload(corpus)
        _check_against_oracle(oracle, idx, corpus, q, 10, metric, label=f"neardup/m{metric}")
        idx.close()


def test_tombstones(capi, oracle):
    rng = np.random.default_rng(12)
    corpus = rng.random((8000, 96), dtype=np.float32)
    queries = rng.random((4, 96), dtype=np.float32)
    idx = capi.Index(96, 1)
    idx.upload(corpus)
    dead = rng.choice(8000, 3000, replace=False)
    # delete the current nearest neighbours too
    d0, r0, _, _ = idx.search(queries, 10)
    dead = np.unique(np.concatenate([dead, r0[:, :5].ravel()]))
    idx.tombstone(dead)
    idx.tombstone(dead[:10])  # deleting twice is a no-op (exact.go:61-70)
    live = np.ones(8000, dtype=np.uint8)
    live[dead] = 0
    assert idx.size == int(live.sum()) and idx.rows == 8000
    _check_against_oracle(oracle, idx, corpus, queries, 10, 1, live=live, label="tombstones")
    idx.close()


def test_incremental_upload_matches_bulk(capi, oracle):
    rng = np.random.default_rng(13)
    corpus = rng.random((5000, 128), dtype=np.float32)
    queries = rng.random((2, 128), dtype=np.float32)
    idx = capi.Index(128, 0)
    off = 0
    for chunk in (1, 999, 1500, 2500):
        first = idx.upload(corpus[off:off + chunk])
        assert first == off
        off += chunk
    _check_against_oracle(oracle, idx, corpus, queries, 10, 0, label="incremental")
    np.testing.assert_array_equal(idx.fetch([0, 17, 4999]), corpus[[0, 17, 4999]])
    idx.close()


def test_edge_semantics(capi):
    """exact.go:96-111 order: empty -> [], nil even for k <= 0; then dimension; then k."""
    idx = capi.Index(8, 0)
    q = np.ones((1, 8), dtype=np.float32)
    dist, row, cnt, _ = idx.search(q, 0)
    assert cnt[0] == 0
    dist, row, cnt, _ = idx.search(q, 5)
    assert cnt[0] == 0 and np.all(row == -1)
    idx.upload(np.eye(8, dtype=np.float32))
    with pytest.raises(capi.QuiverGpuError) as e:
        idx.search(np.ones((1, 7), dtype=np.float32), 3)
    assert e.value.code == capi.QG_ERR_DIM and "query dimension mismatch: expected 8, got 7" in str(e.value)
    with pytest.raises(capi.QuiverGpuError) as e:
        idx.search(q, 0)
    assert e.value.code == capi.QG_ERR_K and "k must be positive" in str(e.value)
    with pytest.raises(capi.QuiverGpuError) as e:
        idx.search(q, -3)
    assert e.value.code == capi.QG_ERR_K
    dist, row, cnt, _ = idx.search(q, 100)
    assert cnt[0] == 8  # k clamps to N (exact.go:109-111)
    idx.tombstone(np.arange(8))
    dist, row, cnt, _ = idx.search(q, 0)  # empty again: no error
    assert cnt[0] == 0 and idx.size == 0
    idx.close()


@pytest.mark.parametrize("kind,dim", [(0, 128), (1, 128), (2, 768), (3, 96), (3, 100)])
def test_synthetic_generator_matches_oracle_bits(capi, oracle, kind, dim):
    n = 3000
    idx = capi.Index(dim, 1)
    idx.upload_synthetic(kind, 42, 1000, n)
    got = idx.fetch(np.arange(n))
    want = oracle.synth(kind, 42, 1000, n, dim)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    idx.close()


@pytest.mark.parametrize("metric", [0, 1, 2, 3, 4])
def test_batch_distance_exact(capi, oracle, metric):
    """hnsw.go:547 computeDistance, batched: bit-identical to the per-pair reference arithmetic."""
    rng = np.random.default_rng(21)
    corpus = rng.standard_normal((2000, 128)).astype(np.float32)
    queries = rng.standard_normal((6, 128)).astype(np.float32)
    idx = capi.Index(128, metric)
    idx.upload(corpus)
    rows = rng.integers(0, 2000, size=(6, 32)).astype(np.uint32)
    rows[0, 3] = 0xFFFFFFFF
    out = idx.batch_distance(queries, rows)
    for b in range(6):
        for j in range(32):
            if rows[b, j] == 0xFFFFFFFF:
                assert np.isinf(out[b, j])
            else:
                want = oracle.distance(metric, queries[b], corpus[rows[b, j]])
                assert out[b, j].view(np.uint32) == want.view(np.uint32), (metric, b, j, out[b, j], want)
    idx.close()


@pytest.mark.parametrize("dim,metric", [(32, "l2"), (64, "dot"), (96, "l2"), (128, "l2"), (128, "cosine"), (256, "dot"),
                                        (768, "cosine"), (1536, "l2")])
def test_dense_flat_scan_with_threshold_exchange(capi, oracle, dim, metric):
    """scan.cuh: scan_dense_kernel on an index WITHOUT the bf16 copy, large enough for the threshold exchange
    between the CTAs (several tiles per warp), with tombstones, exact duplicates, every query-block size and the
    k ranges that change the exchange rule (kp = 32 / 64 / 128: r = 1 / 2 / 4; kp >= 256: no exchange)."""
    m = METRICS[metric]
    rng = np.random.default_rng(dim + 31 * m)
    n = {32: 400_000, 64: 200_000, 96: 150_000, 128: 120_000, 256: 60_000, 768: 30_000, 1536: 12_000}[dim]
    corpus = rng.standard_normal((n, dim)).astype(np.float32)
    corpus[n // 2:n // 2 + 40] = corpus[17]  # ties: (distance, row) order must hold across CTAs
    queries = np.concatenate([rng.standard_normal((7, dim)).astype(np.float32), corpus[17:18]])
    idx = capi.Index(dim, m, flags=capi.FLAG_NO_BF16_COPY)
    idx.upload(corpus)
    dead = rng.choice(n, n // 50, replace=False)
    idx.tombstone(dead)
    live = np.ones(n, dtype=np.uint8)
    live[dead] = 0
    for k, nq in ((10, 1), (1, 1), (50, 1), (100, 1), (300, 1), (10, 2), (10, 3), (40, 7), (10, 5), (10, 4)):
        dist, row, cnt, _ = idx.search(queries[:nq], k)
        assert idx.stats()["path"] == 1, idx.stats()  # batches below 8 queries: flat scan (tf32 regime from 8 on)
        for i in {0, nq // 2, nq - 1}:
            od, orow = oracle.exact_search(corpus, queries[i], k, m, 0, live)
            assert cnt[i] == len(od), (k, nq, i, cnt[i])
            assert np.array_equal(row[i, :len(orow)], orow), (dim, metric, k, nq, i, row[i], orow)
            assert np.array_equal(dist[i, :len(od)].view(np.uint32), od.view(np.uint32))
    idx.close()


_ROUND1_KERNEL_SCRIPT = r"""
import numpy as np
from quiver_b200 import capi
rng = np.random.default_rng(11)
STRING = 1 << 2
for metric, d in ((1, 128), (0, 96), (2, 768)):
    n = 60_000 if d <= 128 else 15_000
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    queries = rng.standard_normal((5, d)).astype(np.float32)
    idx = capi.Index(d, metric, flags=capi.FLAG_NO_BF16_COPY)
    idx.upload(corpus)
    cat = rng.integers(0, 8, n).astype(np.int32)
    idx.set_column(0, np.full(n, 2, dtype=np.uint8) | np.uint8(0x80), np.zeros(n), cat, cat)
    flt = capi.Filter(idx, [capi.qg_pred(0, 1, 0, 1)], [capi.qg_clause(7, 0, 0, 3, STRING, 0, 0.0, 0.0)])
    for nq, k, f in ((1, 10, None), (3, 50, None), (1, 10, flt), (4, 20, flt)):
        dist, row, cnt, _ = idx.search(queries[:nq], k, filter=f)
        assert idx.stats()["path"] == (2 if f is not None else 1), idx.stats()
        xd, xr, xc = idx.search_exhaustive(queries[:nq], k, filter=f)
        assert np.array_equal(row, xr) and np.array_equal(dist.view(np.uint32), xd.view(np.uint32)), (metric, d, nq, k)
print("ok")
"""


def test_round1_fast_kernel_stays_bit_identical():
    """QG_SCAN_DENSE=0 (read once per process) routes flat scans back to scan_fast_kernel, the round-1 kernel that is
    kept as the opt-out: dense and row-list scans must still equal the exhaustive GPU path bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QG_SCAN_DENSE="0", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", _ROUND1_KERNEL_SCRIPT], env=env, cwd=root, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
