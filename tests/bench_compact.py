"""Measure tombstone compaction (qg_index_compact) on one B200: N x d synthetic rows, a random half deleted.
Prints one JSON line: wall time of the whole call (allocation of the new arrays, the scans, the row move, the
D2H of the old->new map, the frees), the algorithmic bytes of the row move, and a full-size parity property
(rows fetched after the squeeze are bit-identical to the generator's rows of the surviving old ids; a search
returns the same neighbours, renumbered).  Kernel durations come from the ncu launch list of this command
(profiles/r01_launches_compact.csv).   usage: bench_compact.py [rows] [dim] [dead_fraction]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quiver_b200 import capi  # noqa: E402
import oracle  # noqa: E402  (checker only)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    capi.load()
    oracle.build()
    idx = capi.Index(d, 1)
    idx.upload_synthetic(1, 42, 0, n)
    rng = np.random.default_rng(5)
    dead = rng.choice(n, int(n * frac), replace=False)
    q = oracle.synth(1, 9999, 0, 64, d)
    idx.tombstone(dead)
    d0, r0, c0, _ = idx.search(q, 10)
    t0 = time.perf_counter()
    old_to_new = idx.compact()
    wall_ms = (time.perf_counter() - t0) * 1e3
    n_live = idx.rows
    d1, r1, c1, _ = idx.search(q, 10)
    same = bool(np.array_equal(r1, old_to_new[r0]) and np.array_equal(d1.view(np.uint32), d0.view(np.uint32)))
    live_old = np.nonzero(old_to_new >= 0)[0]
    probe = live_old[:: max(1, len(live_old) // 64)][:64]
    got = idx.fetch(old_to_new[probe])
    want = np.stack([oracle.synth(1, 42, int(r), 1, d, threads=1)[0] for r in probe])
    rows_ok = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
    dp = (d + 3) & ~3
    extra = 3 if (d + 3 + 63) // 64 == (d + 63) // 64 else 0  # tc_extra_cols (tc_scan.cuh): L2 row-term columns
    dp16 = ((d + extra + 7) & ~7) if d <= 512 else 0
    row_bytes = dp * 4 + dp16 * 2 + 12
    algo = 2 * n_live * row_bytes + n * 4
    print(json.dumps({"op": "qg_index_compact", "rows_before": n, "rows_after": n_live, "dim": d,
                      "wall_ms": round(wall_ms, 3), "row_move_algorithmic_bytes": algo,
                      "row_move_floor_ms_at_measured_hbm": round(algo / 6551.4e9 * 1e3, 3),
                      "search_same_after": same, "rows_bit_identical": rows_ok,
                      "order_kept": bool(np.array_equal(old_to_new[live_old], np.arange(n_live)))}), flush=True)
    idx.close()
    if not (same and rows_ok):
        sys.exit(1)


if __name__ == "__main__":
    main()
