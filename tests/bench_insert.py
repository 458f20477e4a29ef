"""Single-vector Insert rate through the host layer (qh_index_insert), with and without the write-combining
buffer (QH_INSERT_BUFFER=0 uploads every Insert on its own: one pinned-staging H2D, three small kernels and a
stream synchronise per vector). The reference's Insert is a map store (exact.go:38-58).
usage: bench_insert.py [n] [dim]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, d):
    import numpy as np
    from quiver_b200 import hostapi as H
    H.load()
    rng = np.random.default_rng(1)
    corpus = rng.random((n, d), dtype=np.float32)
    ids = [f"v{i}" for i in range(n)]
    idx = H.HybridIndex(d, "euclidean")
    idx.Insert("warm", corpus[0])
    idx.Search(corpus[0], 1)
    t0 = time.perf_counter()
    for i in range(n):
        idx.Insert(ids[i], corpus[i])
    res = idx.Search(corpus[n - 1], 1)   # forces the last upload
    dt = time.perf_counter() - t0
    assert res[0][0] == ids[n - 1] and idx.Size() == n + 1
    idx.close()
    return n / dt


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    if os.environ.get("_BENCH_INSERT_CHILD"):
        print(run(n, d))
        sys.exit(0)
    out = {}
    for label, cap in (("buffered_4096", "4096"), ("unbuffered", "0")):
        env = dict(os.environ, _BENCH_INSERT_CHILD="1", QH_INSERT_BUFFER=cap)
        r = subprocess.run([sys.executable, __file__, str(n), str(d)], env=env, capture_output=True, text=True)
        out[label + "_inserts_per_s"] = float(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]
    print(json.dumps({"op": "qh_index_insert, one vector per call (Python ctypes caller)", "n": n, "dim": d, **out}))
