"""Builds the C5 graph (HNSW M=16, MaxM0=32, efConstruction=200 over 1M x 128 uniform vectors) with the oracle's
restatement of hnsw.Insert on the host and stores its flat arrays (tools/data/, not tracked): graph construction
on the GPU is SURVEY 8 row f-3, not part of the search path this measures.
usage: build_hnsw_graph.py rows standard(0|1) out.npz"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from oracle import hnsw  # noqa: E402

rows, standard, out = int(sys.argv[1]), sys.argv[2], sys.argv[3]
os.environ["QO_HNSW_STANDARD"] = standard
oracle.build()
corpus = oracle.synth(0, 42, 0, rows, 128, threads=8)
t = time.perf_counter()
g = hnsw.Graph(corpus, 1, M=16, MaxM0=32, EfConstruction=200, EfSearch=128, seed=1)
dt = time.perf_counter() - t
e = g.export()
np.savez(out, build_s=dt, **{k: np.asarray(v) for k, v in e.items()})
print(f"{out}: {rows} nodes built in {dt:.1f} s (one host thread)")
