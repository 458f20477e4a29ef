"""Development aid: run one tensor-core pass and compare the raw tf32 scan scores with numpy.
usage: python tools/tc_debug.py [n] [d] [nq] [metric]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quiver_b200 import capi  # noqa: E402


def ordered_to_f32(u):
    u = u.astype(np.uint32)
    b = np.where(u & 0x80000000, u & 0x7FFFFFFF, ~u).astype(np.uint32)
    return b.view(np.float32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    metric = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    rng = np.random.default_rng(1)
    corpus = rng.random((n, d), dtype=np.float32)
    queries = rng.random((nq, d), dtype=np.float32)
    idx = capi.Index(d, metric)
    idx.upload(corpus)
    tau, cnt, cand = idx.debug_tc_pass(queries, 10)
    print("tau", tau[:8], "cnt", cnt[:16], "sum", int(cnt.sum()))
    c64, q64 = corpus.astype(np.float64), queries.astype(np.float64)
    worst = 0.0
    for qi in range(min(nq, 4)):
        m = min(int(cnt[qi]), cand.shape[1])
        keys = cand[qi, :m]
        rows = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
        sc = ordered_to_f32((keys >> np.uint64(32)).astype(np.uint32))
        if metric == 1:
            ref_all = (c64 * c64).sum(1) - 2.0 * (c64 @ q64[qi])
        elif metric == 2:
            ref_all = 1.0 - c64 @ q64[qi]
        else:
            ref_all = 1.0 - (c64 @ q64[qi]) / (np.linalg.norm(c64, axis=1) * np.linalg.norm(q64[qi]))
        ok_rows = rows < n
        err = np.abs(sc[ok_rows] - ref_all[rows[ok_rows]])
        expect = int((ref_all <= tau[qi]).sum())
        print(f"q{qi}: cand {m} (expected about {expect} rows <= tau {tau[qi]:.5f}), rows in range {ok_rows.mean():.3f}, "
              f"max |score - ref| {err.max() if m else 0:.3e}, dup rows {m - len(set(rows.tolist()))}")
        if m:
            worst = max(worst, float(err.max()))
            missing = set(np.nonzero(ref_all <= tau[qi] - 0.05 * abs(tau[qi]) - 1e-3)[0].tolist()) - set(rows.tolist())
            print(f"     clearly-below-tau rows missing from the list: {len(missing)}")
    dist, row, c, _ = idx.search(queries, 10)
    print("search stats", idx.stats(), "counts", c[:8])
    if metric == 1:
        bf = np.sqrt(np.maximum(((c64[None, :, :] - q64[:2, None, :]) ** 2).sum(-1), 0))
        for qi in range(2):
            want = np.argsort(bf[qi], kind="stable")[:10]
            print("q", qi, "rows equal:", np.array_equal(want, row[qi]), want[:5], row[qi][:5])
    print("worst score error", worst)


if __name__ == "__main__":
    main()
