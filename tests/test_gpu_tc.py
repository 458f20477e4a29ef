"""Tensor-core regime (tc_scan.cu: TMA + tcgen05.mma kind::f16 / kind::tf32 + TMEM) against the CPU oracle.
The bf16 / tf32 scores only select candidates; the returned lists must still be bit-identical to the
oracle (exact re-rank + certificate), for every metric, ragged dimensions, masks and batch sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(oracle, idx, corpus, queries, k, metric, which, live=None, arith=0, label=""):
    dist, row, cnt, _ = idx.search(queries, k)
    st = idx.stats()
    for i in which:
        od, orow = oracle.exact_search(corpus, queries[i], k, metric, arith, live)
        n = len(od)
        assert cnt[i] == n, f"{label} q{i}: count {cnt[i]} != {n} ({st})"
        assert np.array_equal(row[i, :n], orow), f"{label} q{i}: rows {row[i, :n]} vs {orow} ({st})"
        assert np.array_equal(dist[i, :n].view(np.uint32), od.view(np.uint32)), f"{label} q{i}: distances"
    return st


@pytest.mark.parametrize("metric", [1, 0, 2, 3])
@pytest.mark.parametrize("nq", [8, 33, 256, 300])
def test_tc_parity_batches(capi, oracle, metric, nq):
    rng = np.random.default_rng(nq * 10 + metric)
    n, d = 60000, 128
    corpus = rng.random((n, d), dtype=np.float32)
    queries = rng.random((nq, d), dtype=np.float32)
    idx = capi.Index(d, metric)
    idx.upload(corpus)
    which = sorted(set([0, 1, nq // 2, nq - 1, min(nq - 1, 255), min(nq - 1, 256)]))
    st = _check(oracle, idx, corpus, queries, 10, metric, which, label=f"m{metric}/Q{nq}")
    assert st["path"] == 3, st
    assert st["escalations"] <= max(1, nq // 50), st
    idx.close()


@pytest.mark.parametrize("d", [32, 96, 100, 260, 768])
def test_tc_parity_dims(capi, oracle, d):
    rng = np.random.default_rng(d)
    n = 40000
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    queries = rng.standard_normal((24, d)).astype(np.float32)
    for metric in (1, 0):
        idx = capi.Index(d, metric)
        idx.upload(corpus)
        st = _check(oracle, idx, corpus, queries, 10, metric, [0, 7, 23], label=f"d{d}/m{metric}")
        assert st["path"] == 3, st
        idx.close()


def test_tc_k100_and_sift_like(capi, oracle):
    n, d = 100000, 128
    idx = capi.Index(d, 1)
    idx.upload_synthetic(1, 42, 0, n)
    corpus = oracle.synth(1, 42, 0, n, d)
    queries = oracle.synth(1, 9999, 0, 64, d)
    st = _check(oracle, idx, corpus, queries, 10, 1, [0, 31, 63], label="sift/k10")
    assert st["path"] == 3
    st = _check(oracle, idx, corpus, queries, 100, 1, [0, 63], label="sift/k100")
    assert st["path"] == 3
    idx.close()


def test_tc_with_tombstones(capi, oracle):
    rng = np.random.default_rng(5)
    n, d = 50000, 96
    corpus = rng.random((n, d), dtype=np.float32)
    queries = rng.random((40, d), dtype=np.float32)
    idx = capi.Index(d, 1)
    idx.upload(corpus)
    d0, r0, _, _ = idx.search(queries, 10)
    dead = np.unique(np.concatenate([rng.choice(n, n // 3, replace=False), r0[:, :5].ravel()]))
    idx.tombstone(dead)
    live = np.ones(n, dtype=np.uint8)
    live[dead] = 0
    st = _check(oracle, idx, corpus, queries, 10, 1, [0, 19, 39], live=live, label="tombstones")
    assert st["path"] == 3
    idx.close()


def test_tc_adversarial_order_falls_back_exactly(capi, oracle):
    """Neighbours clustered in one stretch of rows and near-duplicate rows: the sample threshold is
    a heuristic, so whatever it does the certified result must still equal the oracle's."""
    rng = np.random.default_rng(6)
    n, d = 40000, 64
    corpus = rng.random((n, d), dtype=np.float32)
    order = np.argsort(corpus[:, 0])
    corpus = np.ascontiguousarray(corpus[order])
    queries = rng.random((16, d), dtype=np.float32)
    corpus[1000:1400] = queries[0] + 1e-3 * rng.standard_normal((400, d)).astype(np.float32)
    for metric in (1, 0):
        idx = capi.Index(d, metric)
        idx.upload(corpus)
        _check(oracle, idx, corpus, queries, 10, metric, [0, 1, 15], label=f"adversarial/m{metric}")
        _check(oracle, idx, corpus, queries, 100, metric, [0, 15], label=f"adversarial/k100/m{metric}")
        idx.close()


def test_tc_device_api_matches_flat_path(capi):
    """The same batch through the tensor-core regime and, forced query by query, the flat scan."""
    import torch
    n, d, nq, k = 80000, 128, 64, 10
    idx = capi.Index(d, 1)
    idx.upload_synthetic(0, 7, 0, n)
    q = torch.rand((nq, d), device="cuda:0")
    dist = torch.empty((nq, k), dtype=torch.float32, device="cuda:0")
    row = torch.empty((nq, k), dtype=torch.int64, device="cuda:0")
    cnt = torch.empty((nq,), dtype=torch.int32, device="cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    idx.search_device(q.data_ptr(), nq, k, dist.data_ptr(), row.data_ptr(), cnt.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert idx.stats()["path"] == 3
    r1, d1 = [], []
    flat = capi.Index(d, 1, flags=capi.FLAG_NO_BF16_COPY)  # without the bf16 copy single queries take the flat scan
    flat.upload_synthetic(0, 7, 0, n)
    for i in range(4):  # one query at a time
        di, ri, _, _ = flat.search(q[i:i + 1].cpu().numpy(), k)
        assert flat.stats()["path"] == 1
        r1.append(ri[0])
        d1.append(di[0])
    flat.close()
    r1, d1 = np.stack(r1), np.stack(d1)
    ok = cnt[:4].cpu().numpy() >= 0
    assert ok.all()
    assert np.array_equal(row[:4].cpu().numpy(), r1)
    assert np.array_equal(dist[:4].cpu().numpy().view(np.uint32), d1.view(np.uint32))
    idx.close()


@pytest.mark.parametrize("nq,k", [(600, 10), (2100, 10), (2300, 100)])
def test_tc_multi_pass_groups_and_chunks(capi, oracle, nq, k):
    """Searches longer than one pass: the sample + threshold stage runs once for all passes, finalize
    once per group of passes, and the host-buffer call stages queries / results in chunks of 2 048 —
    with tombstones, so the masked (row-term) epilogue is the one exercised, and on raw L2 / cosine."""
    rng = np.random.default_rng(nq + k)
    n, d = 60000, 96  # 45 000 live rows after the tombstones: still above the tensor-core threshold
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    queries = rng.standard_normal((nq, d)).astype(np.float32)
    which = sorted(set([0, 255, 256, 511, nq // 2, 2047 % nq, 2048 % nq, nq - 1]))
    for metric in (1, 0):
        idx = capi.Index(d, metric)
        idx.upload(corpus)
        st = _check(oracle, idx, corpus, queries, k, metric, which, label=f"multi/m{metric}/Q{nq}/k{k}")
        assert st["path"] == 3 and st["passes"] == (nq + 255) // 256, st
        dead = rng.choice(n, n // 4, replace=False)
        idx.tombstone(dead)
        live = np.ones(n, dtype=np.uint8)
        live[dead] = 0
        st = _check(oracle, idx, corpus, queries, k, metric, which, live=live, label=f"multi+dead/m{metric}/Q{nq}")
        assert st["path"] == 3, st
        idx.close()


def test_concurrent_searches_share_one_index(capi):
    """The reference serves searches under a shared RLock (collection.go:647; hybrid_stress_test.go:29-74):
    many callers at once on one index. qg_search_batch is re-entrant (per-call workspace + streams); the
    answers of 8 threads hammering every regime must equal the ones computed one call at a time."""
    import threading
    rng = np.random.default_rng(77)
    n, d, k = 50000, 64, 10
    idx = capi.Index(d, 1)
    idx.upload(rng.random((n, d), dtype=np.float32))
    batches = [rng.random((q, d), dtype=np.float32) for q in (1, 4, 64, 300, 2, 700, 33, 1)]
    want = [idx.search(b, k) for b in batches]
    errors = []

    def worker(t):
        try:
            for it in range(6):
                j = (t + it) % len(batches)
                dist, row, cnt, _ = idx.search(batches[j], k)
                if not (np.array_equal(row, want[j][1]) and np.array_equal(cnt, want[j][2]) and
                        np.array_equal(dist.view(np.uint32), want[j][0].view(np.uint32))):
                    errors.append((t, it, j))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(120)
    assert not errors, errors[:5]
    idx.close()


_WARP_FINALIZE_SCRIPT = r"""
import numpy as np
from quiver_b200 import capi
rng = np.random.default_rng(5)
for metric, d in ((1, 128), (0, 64), (2, 96)):
    corpus = rng.standard_normal((150_000, d)).astype(np.float32)
    corpus[7000:7020] = corpus[11]
    queries = np.concatenate([rng.standard_normal((299, d)).astype(np.float32), corpus[11:12]])
    idx = capi.Index(d, metric)
    idx.upload(corpus)
    for k in (1, 10, 16):
        dist, row, cnt, _ = idx.search(queries, k)
        assert idx.stats()["path"] == 3, idx.stats()
        xd, xr, xc = idx.search_exhaustive(queries, k)
        assert np.array_equal(row, xr) and np.array_equal(dist.view(np.uint32), xd.view(np.uint32)), (metric, d, k)
print("ok")
"""


def test_warp_per_query_finalize_is_bit_identical():
    """finalize.cu: finalize_cand_warp_kernel is opt-in (QG_FINALIZE_WARP=1, read once per process): the same
    batches must equal the exhaustive GPU oracle bit for bit, duplicates and ties included."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QG_FINALIZE_WARP="1", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", _WARP_FINALIZE_SCRIPT], env=env, cwd=root, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
