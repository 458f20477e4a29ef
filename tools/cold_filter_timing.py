"""Where a filtered batch behind a FRESH filter handle spends its time (C3 shape scaled): mask evaluation, view build,
scan. usage: python tools/cold_filter_timing.py [rows]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quiver_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = 768
STRING = 1 << 2
idx = capi.Index(d, capi.COSINE, reserve_rows=n)
idx.upload_synthetic(2, 42, 0, n)
rng = np.random.default_rng(0)
cat = rng.integers(0, 10, n).astype(np.int32)
idx.set_column(0, np.full(n, 2, dtype=np.uint8) | np.uint8(0x80), np.zeros(n), cat, cat)
q = rng.standard_normal((32, d)).astype(np.float32)


def t(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, r


for rep in range(3):
    ms_c, f = t(lambda: capi.Filter(idx, [capi.qg_pred(0, 1, 0, 1)], [capi.qg_clause(7, 0, 0, 3, STRING, 0, 0.0, 0.0)]))
    ms_e, _ = t(lambda: f.eval())
    ms_1, _ = t(lambda: idx.search(q, 30, filter=f))
    ms_2, _ = t(lambda: idx.search(q, 30, filter=f))
    ms_x, _ = t(lambda: f.close())
    print(f"rep {rep}: compile {ms_c:.2f} ms | eval (mask) {ms_e:.2f} | first batch (view build + scan) {ms_1:.2f} | "
          f"second batch {ms_2:.2f} | close {ms_x:.2f}   path {idx.stats()['path']}", flush=True)
