"""Measure the BASELINE.json configurations other than the headline one (bench.py) on one B200.
Each line: config, shape, regime the library chose, ms, QPS, algorithmic GB/s, and a parity property
checked at full size (returned distances recomputed by the oracle's pairwise arithmetic on
regenerated rows, ascending order, no uncertified query).  usage: config_bench.py [c1,c3,c4] [scale]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quiver_b200 import capi  # noqa: E402
import oracle  # noqa: E402  (checker only)

DEV = torch.device("cuda:0")


def timed(idx, dq, q, k, flt=None, dneg=None, iters=10, warm=2):
    dist = torch.empty((q, k), dtype=torch.float32, device=DEV)
    row = torch.empty((q, k), dtype=torch.int64, device=DEV)
    cnt = torch.empty((q,), dtype=torch.int32, device=DEV)
    negd = torch.empty((q, k), dtype=torch.float32, device=DEV) if dneg is not None else None
    st = torch.cuda.current_stream().cuda_stream

    def run():
        idx.search_device(dq.data_ptr(), q, k, dist.data_ptr(), row.data_ptr(), cnt.data_ptr(), stream=st, filter=flt,
                          d_negatives=dneg.data_ptr() if dneg is not None else 0,
                          d_negdist=negd.data_ptr() if negd is not None else 0)
    for _ in range(warm):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, dist.cpu().numpy(), row.cpu().numpy(), cnt.cpu().numpy()


def check_rows(kind, seed, dim, metric, q_host, dist, row, cnt, k, normalize_rows=False):
    """Full-size parity property: the returned float32 distances are exactly the oracle's pairwise
    distances of the returned rows (rows regenerated from the counter-based generator)."""
    ok = True
    for i in range(min(len(cnt), 3)):
        if cnt[i] < 0:
            return "UNCERTIFIED"
        rows_i = row[i, :cnt[i]]
        vecs = np.stack([oracle.synth(kind, seed, int(r), 1, dim, threads=1)[0] for r in rows_i])
        want = np.array([oracle.distance(metric, q_host[i], v) for v in vecs], dtype=np.float32)
        ok &= np.array_equal(want.view(np.uint32), dist[i, :cnt[i]].view(np.uint32))
        ok &= bool(np.all(np.diff(dist[i, :cnt[i]]) >= 0))
    return "ok" if ok else "MISMATCH"


def check_membership(idx, q_host, dist, row, cnt, k, flt=None, n_check=32):
    """Full-size membership parity: the fast path's rows and float32 distances equal the exhaustive
    GPU path's (qg_search_exhaustive: every row's exact distance + full sort; validated against the CPU oracle
    in tests/test_gpu_parity_chain.py) for n_check queries."""
    ids = [i for i in range(min(len(cnt), n_check)) if cnt[i] >= 0]
    if not ids:
        return "UNCERTIFIED"
    xd, xr, xc = idx.search_exhaustive(q_host[ids], k, filter=flt)
    ok = np.array_equal(row[ids], xr) and np.array_equal(dist[ids].view(np.uint32), xd.view(np.uint32)) and \
        np.array_equal(cnt[ids], xc)
    return f"{len(ids)} queries == exhaustive GPU oracle" if ok else "MEMBERSHIP MISMATCH"


def emit(out, **rec):
    print(json.dumps(rec), flush=True)
    out.append(rec)


def c1(out):
    n, d, k = 10_000, 128, 10
    idx = capi.Index(d, capi.COSINE, reserve_rows=n)
    idx.upload_synthetic(0, 42, 0, n)
    qh = oracle.synth(0, 9999, 0, 100, d)
    dq = torch.from_numpy(qh).to(DEV)
    corpus = oracle.synth(0, 42, 0, n, d)
    for q in (1, 100):
        ms, dist, row, cnt = timed(idx, dq, q, k, iters=50)
        good = all(np.array_equal(row[i], oracle.exact_search(corpus, qh[i], k, 0)[1]) for i in range(min(q, 5)))
        emit(out, config="C1 hybrid-index exact 10k x 128 cosine", q=q, k=k, ms=round(ms, 4), qps=round(q / ms * 1e3),
             path=idx.stats()["path"], parity="ok" if good else "MISMATCH", note="L2-resident corpus: latency, not roofline")
    idx.close()


def c3(out, scale):
    n, d, k = int(10_000_000 * scale), 768, 10
    idx = capi.Index(d, capi.COSINE, reserve_rows=n)
    t0 = time.time()
    idx.upload_synthetic(2, 42, 0, n)
    rng = np.random.default_rng(43)
    cat = rng.integers(0, 5, n).astype(np.int32)
    tag = rng.integers(0, 20, n).astype(np.int32)
    kind = np.full(n, 2, dtype=np.uint8) | np.uint8(0x80)   # QG_KIND_STRING, non-empty
    zeros = np.zeros(n, dtype=np.float64)
    idx.set_column(0, kind, zeros, cat, cat)
    idx.set_column(1, kind, zeros, tag, tag)
    fill_s = time.time() - t0
    STRING = 1 << 2
    preds = [capi.qg_pred(0, 1, 0, 1), capi.qg_pred(1, 1, 0, 1)]
    clauses = [capi.qg_clause(7, 0, 0, 3, STRING, 0, 0.0, 0.0),       # category == "cat3"
               capi.qg_clause(10, 1, 0, 0, STRING, 10, 0.0, 0.0)]      # tag in {tag00..tag09}
    flt = capi.Filter(idx, preds, clauses, iset=list(range(10)))
    t0 = time.time()
    bits, matches = flt.eval()
    want_bits = (cat == 3) & (tag < 10)
    mask_ok = bool(np.array_equal(bits, want_bits))
    qh = oracle.synth(2, 9999, 0, 32, d)
    negh = oracle.synth(2, 7, 0, 32, d)
    dq, dneg = torch.from_numpy(qh).to(DEV), torch.from_numpy(negh).to(DEV)
    emit(out, config="C3 filter mask (category = cat3 AND tag IN 10 of 20)", rows=n, matches=int(matches),
         selectivity=round(matches / n, 4), mask_bit_exact=mask_ok, fill_s=round(fill_s, 1))
    for q, use_f, use_neg, kk in ((1, False, False, 10), (1, True, False, 10), (1, True, True, 30), (32, True, True, 30),
                                  (32, False, False, 10)):
        ms, dist, row, cnt = timed(idx, dq, q, kk, flt=flt if use_f else None, dneg=dneg if use_neg else None, iters=5)
        stt = idx.stats()
        par = check_rows(2, 42, d, 0, qh, dist, row, cnt, kk)
        if use_f and par == "ok":
            par = "ok" if bool(np.all(want_bits[row[:min(q, 3)].ravel()])) else "FILTER VIOLATED"
        member = check_membership(idx, qh, dist, row, cnt, kk, flt=flt if use_f else None)
        emit(out, config="C3 cosine 10M x 768" + (" + facet prefilter" if use_f else "") +
             (" + negative examples (window max(2k,30))" if use_neg else ""), rows=n, q=q, k=kk, ms=round(ms, 3),
             qps=round(q / ms * 1e3, 1), path=stt["path"], passes=stt["passes"],
             algorithmic_GB=round(stt["bytes_algorithmic"] / 1e9, 3),
             GBps=round(stt["bytes_algorithmic"] * stt["passes"] / (ms * 1e-3) / 1e9, 1), parity=par,
             membership=member)
    # the same filtered batch with a fresh filter handle per call: the view of the passing rows is built inside the call
    import time as _t
    cold = []
    for _ in range(3):
        f2 = capi.Filter(idx, preds, clauses, iset=list(range(10)))
        torch.cuda.synchronize()
        t0 = _t.perf_counter()
        timed(idx, dq, 32, 30, flt=f2, dneg=dneg, iters=1, warm=0)
        cold.append((_t.perf_counter() - t0) * 1e3)
        f2.close()
    emit(out, config="C3 filtered batch, cold filter handle (mask evaluation + ordered gather of the passing rows + scan in one call)",
         rows=n, q=32, k=30, ms=round(min(cold), 3))
    flt.close()
    idx.close()


def c4(out, scale):
    n, d, k = int(100_000_000 * scale), 96, 10
    idx = capi.Index(d, capi.L2, reserve_rows=n)
    t0 = time.time()
    idx.upload_synthetic(3, 42, 0, n)
    fill_s = time.time() - t0
    qh = oracle.synth(3, 9999, 0, 1024, d)
    dq = torch.from_numpy(qh).to(DEV)
    for q in (1, 32, 1024):
        ms, dist, row, cnt = timed(idx, dq, q, k, iters=5 if q < 1024 else 2)
        stt = idx.stats()
        emit(out, config="C4 L2 100M x 96 (Deep-shaped), all rows on ONE GPU", rows=n, q=q, k=k, ms=round(ms, 3),
             qps=round(q / ms * 1e3, 1), path=stt["path"], passes=stt["passes"],
             GBps=round(stt["bytes_algorithmic"] * stt["passes"] / (ms * 1e-3) / 1e9, 1),
             uncertified=int((cnt < 0).sum()), parity=check_rows(3, 42, d, 1, qh, dist, row, cnt, k),
             membership=check_membership(idx, qh, dist, row, cnt, k), fill_s=round(fill_s, 1))
    idx.close()


def main():
    which = (sys.argv[1] if len(sys.argv) > 1 else "c1,c3,c4").split(",")
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    out = []
    if "c1" in which:
        c1(out)
    if "c3" in which:
        c3(out, scale)
    if "c4" in which:
        c4(out, scale)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_" + "_".join(which) + ".json"), "w"), indent=1)


if __name__ == "__main__":
    main()
