"""Small run of every kernel family for compute-sanitizer (racecheck / memcheck): flat scan, tensor-core scan
(bf16 TS, tf32 SS), finalize, filters + view, tombstones + compaction, neighbour batches, the HNSW walk kernel.
usage: compute-sanitizer --tool racecheck python tools/sanitize_driver.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quiver_b200 import capi  # noqa: E402

rng = np.random.default_rng(0)
STRING = 1 << 2
# the dense flat scan with its threshold exchange (an index without the bf16 copy; several tiles per warp)
for metric, d, n in ((1, 128, 60000), (0, 96, 50000), (2, 768, 12000)):
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    idx = capi.Index(d, metric, flags=capi.FLAG_NO_BF16_COPY)
    idx.upload(corpus)
    q = rng.standard_normal((4, d)).astype(np.float32)
    idx.search(q[:1], 10)
    idx.search(q[:1], 100)
    idx.search(q[:2], 40)
    idx.search(q[:4], 10)
    idx.tombstone(np.arange(0, n, 5))
    idx.search(q[:1], 10)
    idx.close()
if os.environ.get("SAN_ONLY") == "dense":
    print("sanitize driver done (dense flat scan only)")
    sys.exit(0)
for metric, d, n in ((1, 128, 20000), (0, 96, 12000), (1, 768, 9000)):
    corpus = rng.standard_normal((n, d)).astype(np.float32)
    idx = capi.Index(d, metric)
    idx.upload(corpus)
    q = rng.standard_normal((40, d)).astype(np.float32)
    idx.search(q[:1], 10)            # flat scan + finalize
    idx.search(q, 10)                # tensor-core scan + finalize_cand
    cat = rng.integers(0, 4, n).astype(np.int32)
    kind = np.full(n, 2, dtype=np.uint8) | np.uint8(0x80)
    idx.set_column(0, kind, np.zeros(n), cat, cat)
    f = capi.Filter(idx, [capi.qg_pred(0, 1, 0, 1)], [capi.qg_clause(7, 0, 0, 1, STRING, 0, 0.0, 0.0)])
    f.eval()
    idx.search(q[:1], 10, filter=f)  # gather scan
    if n >= 4 * 8192:
        idx.search(q, 10, filter=f)  # dense view + tensor-core scan
    idx.tombstone(np.arange(0, n, 3))
    idx.search(q, 10)
    idx.compact()
    idx.search(q[:2], 10)
    idx.batch_distance(q[:4], rng.integers(0, idx.rows, (4, 16)).astype(np.uint32))
    idx.search_exhaustive(q[:2], 5)
    f.close()
    idx.close()
print("sanitize driver done")
