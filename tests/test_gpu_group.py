"""Several GPUs behind the C ABI (qg_group_*: one host process, one worker thread + stream + NCCL communicator
per device; qg_comm_*: one process per GPU). Results must equal the CPU oracle / the single-GPU run bit for
bit in both layouts. The two-device cases skip on a one-GPU box; the one-device group runs everywhere."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same(oracle, corpus, queries, k, metric, dist, row, cnt, which):
    for i in which:
        od, orow = oracle.exact_search(corpus, queries[i], k, metric)
        assert cnt[i] == len(od)
        assert np.array_equal(row[i, :len(od)], orow), (i, row[i], orow)
        assert np.array_equal(dist[i, :len(od)].view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("layout", [0, 1])
def test_group_of_one_device(capi, oracle, layout):
    rng = np.random.default_rng(3)
    corpus = rng.random((30000, 64), dtype=np.float32)
    queries = rng.random((40, 64), dtype=np.float32)
    g = capi.Group([0], 64, capi.L2)
    g.upload(corpus, layout)
    assert g.layout == layout and g.rows == 30000
    dist, row, cnt = g.search(queries, 10)
    _same(oracle, corpus, queries, 10, capi.L2, dist, row, cnt, [0, 17, 39])
    dist, row, cnt = g.search(queries[:1], 5)
    _same(oracle, corpus, queries, 5, capi.L2, dist, row, cnt, [0])
    g.close()


def test_group_errors_follow_the_reference_order(capi):
    g = capi.Group([0], 8, capi.L2)
    d, r, c = g.search(np.zeros((2, 5), dtype=np.float32), 0)  # empty: no results, no error (exact.go:96-99)
    assert c.tolist() == [0, 0]
    g.upload(np.ones((10, 8), dtype=np.float32), 0)
    with pytest.raises(capi.QuiverGpuError) as e:
        g.search(np.zeros((1, 5), dtype=np.float32), 3)
    assert e.value.code == capi.QG_ERR_DIM and "query dimension mismatch: expected 8, got 5" in str(e.value)
    with pytest.raises(capi.QuiverGpuError) as e:
        g.search(np.zeros((1, 8), dtype=np.float32), 0)
    assert e.value.code == capi.QG_ERR_K and "k must be positive" in str(e.value)
    with pytest.raises(capi.QuiverGpuError):
        g.upload(np.ones((10, 8), dtype=np.float32), 0)  # a group is filled once
    g.close()


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("metric", [1, 0])
def test_group_of_two_devices(capi, oracle, layout, metric):
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(layout * 2 + metric)
    n, d = 70001, 96  # odd row count: uneven shards
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    queries = rng.standard_normal((301, d)).astype(np.float32)
    g = capi.Group([0, 1], d, metric)
    g.upload(corpus, layout)
    assert g.layout == layout
    if layout == 0:
        assert g.row_base(1) == (n + 1) // 2
    for nq, k in ((301, 10), (1, 10), (7, 100)):
        dist, row, cnt = g.search(queries[:nq], k)
        _same(oracle, corpus, queries, k, metric, dist, row, cnt, sorted({0, nq // 2, nq - 1}))
    g.close()


def test_comm_two_processes(capi):
    """qg_comm_* with one process per GPU (the bench / torchrun shape): both layouts against a single-GPU run."""
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29653",
                          os.path.join(ROOT, "tests", "comm_worker.py")], capture_output=True, text=True, env=env,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "comm worker ok" in out.stdout
