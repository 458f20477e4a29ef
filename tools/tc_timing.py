"""Development aid: per-role cycle counters of the tensor-core scan (QG_TC_TIMING).
Needs the instrumented build: make -B lib TC_INSTRUMENT=1 (the counters are compiled out by default)."""
import os, sys
import numpy as np
os.environ["QG_TC_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quiver_b200 import capi
n, d, nq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
metric = int(sys.argv[4]) if len(sys.argv) > 4 else 1
idx = capi.Index(d, metric, reserve_rows=n)
idx.upload_synthetic(1, 42, 0, n)
q = np.floor(np.random.default_rng(0).random((nq, d), dtype=np.float32) * 218)
for _ in range(3):
    tau, cnt, cand = idx.debug_tc_pass(q, 10)
print("cand per query mean", cnt.mean(), "max", cnt.max())
