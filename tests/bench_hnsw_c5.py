"""C5 measurement: HNSW search with GPU-batched neighbour distances against the host walk.

Not a pytest module (run by hand on a GPU box): builds the graph with the oracle's restatement of
hnsw.Insert (graph construction is out of scope for the product), walks a query batch through
`qh_hnsw_search_batch` and through the oracle's single-threaded host walk, checks that both return the
same neighbours, and prints one JSON line. Two graphs: the reference's own (whose connectNode quirk
fragments layer 0 — short walks, recall near 0 on random data) and the textbook variant
(QO_HNSW_STANDARD=1, a measurement aid) that shows the regime the neighbour batches are meant for.

usage: python tests/bench_hnsw_c5.py [rows] [queries]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(standard: bool, n: int, nq: int, d: int = 128, k: int = 10):
    import oracle
    from oracle import hnsw
    from quiver_b200 import hostapi
    os.environ["QO_HNSW_STANDARD"] = "1" if standard else "0"
    corpus = oracle.synth(0, 42, 0, n, d, threads=8)
    queries = oracle.synth(0, 9999, 0, nq, d, threads=1)
    t = time.perf_counter()
    graph = hnsw.Graph(corpus, 1, M=16, MaxM0=32, EfConstruction=200, EfSearch=128, seed=1)
    t_build = time.perf_counter() - t
    ids = [f"n{i:07d}" for i in range(n)]
    idx = hostapi.HybridIndex(d, "euclidean")
    idx.InsertBatchArrays(ids, corpus)
    g = graph.export()
    idx.HNSWSearchBatch(g, queries[:64], k)  # warm-up
    t = time.perf_counter()
    res, evals, steps = idx.HNSWSearchBatch(g, queries, k)
    t_gpu = time.perf_counter() - t
    n_host = min(nq, 500)
    t = time.perf_counter()
    host = [graph.search(queries[i], k) for i in range(n_host)]
    t_host = (time.perf_counter() - t) / n_host
    same = all([r[0] for r in res[i]] == [ids[j] for j in host[i][1]] for i in range(n_host) if len(host[i][1]) >= k)
    hit = 0
    n_rec = min(nq, 100)
    for i in range(n_rec):
        od, orow = oracle.exact_search(corpus, queries[i], k, 1)
        hit += len({ids[j] for j in orow} & {r[0] for r in res[i]})
    idx.close()
    return {"graph": "textbook entry-point descent (QO_HNSW_STANDARD=1)" if standard else "reference connectNode (faithful)",
            "rows": n, "dim": d, "queries": nq, "k": k, "efSearch": 128, "build_s_host": round(t_build, 1),
            "distance_evals_per_query": float(np.mean(evals)), "lock_steps": int(steps),
            "gpu_batched_walk_qps": nq / t_gpu, "gpu_batched_walk_ms": t_gpu * 1e3,
            "host_walk_qps_one_thread": 1.0 / t_host, "same_results_as_host_walk": bool(same),
            "recall_at_10_vs_exact": hit / (n_rec * k)}


if __name__ == "__main__":
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    out = [run(False, rows, nq), run(True, rows, nq)]
    print(json.dumps({"metric": "HNSW search, GPU-batched neighbour distances vs host walk (C5)", "runs": out}))
