"""C5 measurement (BASELINE configs[4]): HNSW (M = 16, efSearch = 128) on 1M x 128, hnsw.Search of a query batch
entirely on the device (qh_hnsw_search_device: one persistent kernel, a warp per query) against the reference's
walk on the host cores.

Not a pytest module (run by hand on a GPU box). The graph comes from tools/build_hnsw_graph.py (the oracle's
restatement of hnsw.Insert, built on the host and stored under tools/data/: construction is SURVEY 8 row f-3, not
part of the search path) or, for small sizes, is built here. Two graphs: the reference's own (whose connectNode
quirk fragments layer 0 — short walks; the under-fill exact pass answers most queries) and the textbook variant
(QO_HNSW_STANDARD=1, a measurement aid) that shows the regime the neighbour batches are meant for.
Checked: device results == host walk results (ids, float32 distances, evaluation counts) on a sample.

usage: python tests/bench_hnsw_c5.py [rows] [queries] [both|faithful|textbook|device]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run(standard: bool, n: int, nq: int, d: int = 128, k: int = 10, device_built: bool = False):
    import oracle
    from oracle import hnsw
    from quiver_b200 import capi, hostapi
    corpus = oracle.synth(0, 42, 0, n, d, threads=8)
    queries = oracle.synth(0, 9999, 0, nq, d, threads=1)
    stored = os.path.join(ROOT, "tools", "data", f"hnsw_{n // 1000000}m_{'textbook' if standard else 'faithful'}.npz")
    if device_built:
        # the graph is built on the device (qg_hnsw_build, batched inserts) and exported
        bidx = capi.Index(d, capi.L2, reserve_rows=n)
        bidx.upload(corpus)
        t = time.perf_counter()
        bg = capi.HnswGraph.build(bidx, M=16, MaxM0=32, EfConstruction=200, seed=1)
        t_build = time.perf_counter() - t
        g = bg.export(EfSearch=128)
        bg.close()
        bidx.close()
        graph = hnsw.Graph(corpus, 1, M=16, MaxM0=32, EfSearch=128, flat=g)
    elif n % 1000000 == 0 and os.path.exists(stored):
        z = np.load(stored)
        flat = {kk: z[kk] for kk in ("level", "adj0", "upper_off", "upper_adj")}
        flat.update(n=int(z["n"]), entry=int(z["entry"]), current_level=int(z["current_level"]), M=16, MaxM0=32, EfSearch=128)
        graph = hnsw.Graph(corpus, 1, M=16, MaxM0=32, EfSearch=128, flat=flat)
        t_build = float(z["build_s"])
        g = flat
    else:
        os.environ["QO_HNSW_STANDARD"] = "1" if standard else "0"
        t = time.perf_counter()
        graph = hnsw.Graph(corpus, 1, M=16, MaxM0=32, EfConstruction=200, EfSearch=128, seed=1)
        t_build = time.perf_counter() - t
        os.environ["QO_HNSW_STANDARD"] = "0"
        g = graph.export()
    ids = [f"n{i:07d}" for i in range(n)]
    idx = hostapi.HybridIndex(d, "euclidean")
    idx.InsertBatchArrays(ids, corpus)
    dg = idx.HNSWUpload(g)
    dg.search(queries[:256], k)  # warm-up
    res, evals, fallbacks = dg.search(queries, k)
    t_dev = dg.last_call_s  # qh_hnsw_search_device alone: host buffers in, results out (incl. the under-fill pass)
    # the round-1 path (heaps on the host, one distance launch per lock step) on a slice
    n_ls = min(nq, 2000)
    res_ls, _, steps = idx.HNSWSearchBatch(g, queries[:n_ls], k)
    t_ls = idx.last_call_s
    # the reference's walk on the host: one thread, and all cores (one query per thread at a time)
    cores = host_cores()
    n_host = min(nq, 2000)
    t = time.perf_counter()
    hd, hi, hc, hev = graph.search_batch(queries[:n_host], k, threads=cores)
    # hnsw.go:676-710: an under-filled walk is supplemented by ranking EVERY node (the same work as
    # ExactIndex.Search); the reference pays it inside Search, so it belongs to the host figure
    under = np.nonzero(hc < k)[0]
    n_under_timed = min(len(under), 4 * cores)
    t_host_all = time.perf_counter() - t
    if n_under_timed:
        t = time.perf_counter()
        oracle.exact_search_batch(corpus, queries[under[:n_under_timed]], k, 1, threads=cores)
        t_host_all += (time.perf_counter() - t) * len(under) / n_under_timed
    n_one = min(nq, 300)
    t = time.perf_counter()
    graph.search_batch(queries[:n_one], k, threads=1)
    t_host_one = time.perf_counter() - t
    same = True
    full = 0
    for i in range(n_host):
        same &= int(evals[i]) == int(hev[i])
        if hc[i] >= k:
            full += 1
            same &= [r[0] for r in res[i]] == [ids[j] for j in hi[i, :k]]
            same &= [np.float32(r[1]).view(np.uint32) for r in res[i]] == [x.view(np.uint32) for x in hd[i, :k]]
    same &= res_ls == res[:n_ls]
    # recall of the returned lists (after the under-fill pass) against the exact top-k
    n_rec = min(nq, 200)
    hit = 0
    ex = idx.BatchSearch(list(queries[:n_rec]), k)
    for i in range(n_rec):
        hit += len({r[0] for r in ex[i]} & {r[0] for r in res[i]})
    dg.close()
    idx.close()
    return {"graph": "built on the device (qg_hnsw_build, batched inserts, textbook descent)" if device_built else (
                "textbook entry-point descent (QO_HNSW_STANDARD=1)" if standard else "reference connectNode (faithful)"),
            "rows": n, "dim": d, "queries": nq, "k": k, "efSearch": 128, "M": 16,
            ("build_s_device" if device_built else "build_s_host_one_thread"): round(t_build, 1),
            "distance_evals_per_query": float(np.mean(evals)),
            "graph_walk_filled_k": full / n_host, "fallbacks_to_host_walk": int(fallbacks),
            "device_walk_qps": nq / t_dev, "device_walk_ms": t_dev * 1e3,
            "lockstep_walk_qps_round1": n_ls / t_ls, "lock_steps": int(steps),
            "host_walk_qps_all_cores": n_host / t_host_all, "host_cores": cores,
            "host_walk_qps_one_thread": n_one / t_host_one,
            "device_over_host_all_cores": (nq / t_dev) / (n_host / t_host_all),
            "same_results_and_eval_counts_as_host_walk": bool(same),
            "recall_at_10_vs_exact": hit / (n_rec * k)}


if __name__ == "__main__":
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    out = []
    if which in ("both", "faithful"):
        out.append(run(False, rows, nq))
    if which in ("both", "textbook"):
        out.append(run(True, rows, nq))
    if which in ("both", "device"):
        out.append(run(True, rows, nq, device_built=True))
    print(json.dumps({"metric": "HNSW search on the device vs the reference's walk on the host (C5)", "runs": out}))
