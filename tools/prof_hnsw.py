"""Builds an HNSW graph on the device and runs a few search batches (for ncu captures). argv: n d nq"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quiver_b200 import capi
n, d, nq = [int(x) for x in sys.argv[1:4]]
idx = capi.Index(d, capi.L2, reserve_rows=n)
idx.upload_synthetic(0, 42, 0, n)
g = capi.HnswGraph.build(idx, M=16, MaxM0=32, EfConstruction=200, seed=1)
q = np.random.default_rng(0).random((nq, d), dtype=np.float32)
for _ in range(3):
    i, dist, cnt, ev = g.search(q, 10, ef_search=128)
print("evals/query", ev.mean(), "filled", (cnt == 10).mean())
