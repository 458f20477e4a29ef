#!/usr/bin/env python3
"""bench.py — exact k-NN QPS on the reference's headline configuration, through the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--queries Q] [--k 10]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): flat L2 exact search, 1M x 128 fp32 SIFT-shaped synthetic
corpus (oracle/synth.h kind 1, seed 42), k = 10, one step = one batch of Q queries (default 10 000,
the top of the config's "query batch 1-10k" range; the library serves it as passes of 256 queries
through the tensor-core regime; --queries 1 exercises the flat-scan regime).
  value   queries/s with the query batch already resident in HBM (device API, CUDA events)
  e2e     queries/s through qg_search_batch with HOST buffers (pinned staging, H2D + D2H inside)
  N > 1   one process per GPU; the data-path collective is NCCL inside libquivergpu (qg_comm_*). The
          headline corpus (0.5 GB) is replicated and the batch split (quiver_b200/sharded.choose_layout;
          --shard rows forces the row-sharded layout); the `row_sharded` sub-record measures north_star's
          layout — per-shard top-k + all-gather of the keys + merge — on C4-shaped shards of 12.5M x 96
          rows per GPU (100M x 96 at 8 GPUs) for batches of 1 / 32 / 1024 queries.
  real_valued (N = 1): the same 1M x 128 / 10 000-query step on real-valued corpora (U[0,1), N(0,1)),
          where every element is rounded in the bf16 copy.
The oracle (oracle/) is used only as the checker and for the cpu_baseline / --impl reference legs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "exact k-NN QPS (k=10, 1M x 128 L2)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--queries", type=int, default=10000, help="queries per step (batch; BASELINE config: 1..10k)")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="l2", choices=["l2", "cosine", "dot"])
    ap.add_argument("--kind", type=int, default=1, help="synthetic kind (oracle/synth.h)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true",
                    help="skip the real_valued / row_sharded sub-records and the ncu traffic step")
    ap.add_argument("--shard", default="auto", choices=["auto", "rows", "queries"],
                    help="multi-GPU layout (quiver_b200/sharded.py choose_layout): row-sharded corpus + all-gather "
                         "merge, or replicated corpus + query-split batch")
    return ap.parse_args()


METRIC_ID = {"cosine": 0, "l2": 1, "dot": 2}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_queries(oracle, args, nq):
    # queries come from the same generator, a different seed (query seed 9999)
    return oracle.synth(args.kind, 9999, 0, nq, args.dim, threads=min(8, host_cores()))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_qps(oracle, args, corpus, steps, warmup, budget_s):
    """The reference algorithm (oracle port of exact.go:92-133 under BatchSearch's goroutine-per-query
    model, hybrid_index.go:703-795) on all host cores. One step = one batch of `cores` queries."""
    cores = host_cores()
    mid = METRIC_ID[args.metric]
    q1 = make_queries(oracle, args, 1)
    t0 = time.perf_counter()
    oracle.exact_search_batch(corpus, q1, args.k, mid, threads=1)
    t_single = time.perf_counter() - t0
    # one step = one query per host core, all cores busy (a step takes about t_single of wall clock);
    # the run is bounded through the number of steps, never by leaving cores idle
    per_step = cores
    threads = cores
    affordable = max(1, int(budget_s / max(t_single, 1e-6)))
    if steps + warmup > affordable:
        warmup = min(warmup, max(1, affordable // 8))
        steps = max(1, affordable - warmup)
    qs = make_queries(oracle, args, per_step)
    for _ in range(warmup):
        oracle.exact_search_batch(corpus, qs, args.k, mid, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.exact_search_batch(corpus, qs, args.k, mid, threads=threads)
    dt = time.perf_counter() - t0
    qps = per_step * steps / dt
    return {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "steps": steps, "warmup": warmup,
            "sample": f"{steps} steps x {per_step} queries over the full {args.rows}x{args.dim} corpus, "
                      f"one query per thread on {threads} of {cores} host cores (full scan + full sort per query, "
                      f"exact.go:114-129); single query {t_single*1e3:.1f} ms"}, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    corpus = oracle.synth(args.kind, args.seed, 0, args.rows, args.dim, threads=min(16, host_cores()))
    # --steps / --warmup are honoured as given; only a run that would pass ~3 minutes of wall clock is cut
    base, ms = cpu_reference_qps(oracle, args, corpus, max(1, args.steps), max(1, args.warmup), budget_s=180.0)
    steps, warmup = base.pop("steps"), base.pop("warmup")
    line = {"metric": METRIC, "value": base["value"], "unit": "queries/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"flat {args.metric} exact search {args.rows}x{args.dim} fp32, k={args.k}, "
                                   "reference algorithm (C restatement of the Go path; no Go toolchain in the image)",
                       "rows": args.rows, "dim": args.dim, "k": args.k},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def measure_traffic(args, kernel_regex, skip):
    """dram__bytes_read + write of ONE launch of the dominant kernel, from a short ncu run of the same
    library on the same workload (tools/prof_once.py), outside every timed region. None when ncu cannot
    profile on this box."""
    import csv
    import io
    cmd = ["ncu", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "-k", f"regex:{kernel_regex}", "-s", str(skip), "-c", "1", sys.executable,
           os.path.join(ROOT, "tools", "prof_once.py"), str(args.rows), str(args.dim), str(METRIC_ID[args.metric]),
           str(min(args.queries, 256)), str(args.k), "3"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240).stdout
        rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 5]
        hdr = next(r for r in rows if "Metric Name" in r)
        i_name, i_unit, i_val = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        n = 0
        for r in rows:
            if r is hdr or len(r) <= i_val or not r[i_name].startswith("dram__bytes_"):
                continue
            tot += float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
            n += 1
        return int(tot) if n >= 2 else None
    except Exception:
        return None


def check_against_oracle(oracle, corpus, q_host, ids, k, mid, got_d, got_r, cnt_host):
    """Bit-identical rows and float32 distances for the listed queries (threaded oracle, one query per core)."""
    ids = [i for i in ids if cnt_host[i] >= 0]
    if not ids:
        return 0
    od, orow, ocnt = oracle.exact_search_batch(corpus, q_host[ids], k, mid, threads=host_cores())
    for j, i in enumerate(ids):
        n = int(ocnt[j])
        assert np.array_equal(got_r[i, :n], orow[j, :n]), (i, got_r[i], orow[j])
        assert np.array_equal(got_d[i, :n].view(np.uint32), od[j, :n].view(np.uint32)), (i, got_d[i], od[j])
    return len(ids)


def real_valued_record(capi, oracle, torch, args, dev, peaks):
    """SURVEY 8d's real-valued variants of the headline corpus (kind 0 = U[0,1), kind 2 = approx N(0,1)): every
    element is rounded when it becomes bf16, so the certificate's error bound is exercised at full size
    (kind 1's integers 0..217 are exact in bf16)."""
    out = []
    Q, k, d, mid = args.queries, args.k, args.dim, METRIC_ID[args.metric]
    st = torch.cuda.current_stream().cuda_stream
    for kind in (0, 2):
        idx = capi.Index(d, mid, device=dev.index or 0, reserve_rows=args.rows)
        idx.upload_synthetic(kind, args.seed, 0, args.rows)
        qh = oracle.synth(kind, 9999, 0, Q, d, threads=min(8, host_cores()))
        dq = torch.from_numpy(qh).to(dev)
        dd = torch.empty((Q, k), dtype=torch.float32, device=dev)
        dr = torch.empty((Q, k), dtype=torch.int64, device=dev)
        dc = torch.empty((Q,), dtype=torch.int32, device=dev)
        for _ in range(3):
            idx.search_device(dq.data_ptr(), Q, k, dd.data_ptr(), dr.data_ptr(), dc.data_ptr(), stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_it = 10
        e0.record()
        for _ in range(n_it):
            idx.search_device(dq.data_ptr(), Q, k, dd.data_ptr(), dr.data_ptr(), dc.data_ptr(), stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_it
        cnt = dc.cpu().numpy()
        unc = int((cnt < 0).sum())
        # the host entry point repeats uncertified queries (flat scan, then the exhaustive path)
        hd, hr, hc, _ = idx.search(qh, k)
        esc = idx.stats()["escalations"]
        assert (hc == min(k, args.rows)).all()
        corpus = oracle.synth(kind, args.seed, 0, args.rows, d, threads=min(16, host_cores()))
        ids = sorted(set(int(x) for x in np.linspace(0, Q - 1, 64)) | set(np.nonzero(cnt < 0)[0][:32].tolist()))
        n_chk = check_against_oracle(oracle, corpus, qh, ids, k, mid, hd, hr, hc)
        out.append({"kind": kind, "data": {0: "uniform [0,1)", 2: "approx normal (sum of 4 uniforms)"}[kind],
                    "rows": args.rows, "dim": d, "queries_per_step": Q, "k": k, "ms_per_step": ms,
                    "qps_device_api": (Q - unc) / (ms * 1e-3), "uncertified_device_api": unc,
                    "escalations_host_api": int(esc),
                    "parity": f"{n_chk} queries (spread + every uncertified one, through qg_search_batch) bit-identical "
                              f"to the oracle over all {args.rows} rows"})
        idx.close()
        del corpus
    return out


def row_sharded_record(capi, oracle, torch, dist, comm, world, rank, local_rank, dev, peaks):
    """north_star's layout on C4-shaped shards: 12.5M x 96 L2-normalised rows (synthetic kind 3) per GPU —
    100M x 96 at 8 GPUs — per-shard top-k, one NCCL all-gather of the keys inside libquivergpu
    (qg_comm_search_rows_device), merge. Weak scaling: the shard size is fixed, the corpus grows with N."""
    per, d, k, mid = 12_500_000, 96, 10, METRIC_ID["l2"]
    row0 = rank * per
    shard = capi.Index(d, mid, device=local_rank, reserve_rows=per)
    shard.upload_synthetic(3, 42, row0, per)
    st = torch.cuda.current_stream().cuda_stream
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    recs = []
    for Q in (1, 32, 1024):
        qh = oracle.synth(3, 9999, 0, Q, d, threads=min(8, host_cores()))
        dq = torch.from_numpy(qh).to(dev)
        dd = torch.empty((Q, k), dtype=torch.float32, device=dev)
        dr = torch.empty((Q, k), dtype=torch.int64, device=dev)
        dc = torch.empty((Q,), dtype=torch.int32, device=dev)
        keys = torch.empty((Q, k), dtype=torch.int64, device=dev)

        def step():
            comm.search_rows_device(shard, dq.data_ptr(), Q, k, row0, dd.data_ptr(), dr.data_ptr(), dc.data_ptr(), stream=st)
        for _ in range(3):
            step()
        dist.barrier()
        torch.cuda.synchronize()
        n_it = 20 if Q < 1024 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        for _ in range(n_it):
            step()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n_it], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        stats = shard.stats()
        # the shard scan alone (no exchange), for the share of the all-gather + merge
        e0.record()
        for _ in range(n_it):
            shard.search_shard_keys_device(dq.data_ptr(), Q, k, row0, keys.data_ptr(), stream=st)
        e1.record()
        torch.cuda.synchronize()
        t2 = torch.tensor([e0.elapsed_time(e1) / n_it], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        scan_ms = float(t2.item())
        # parity at full size: this rank's shard through the fast path == the exhaustive GPU oracle
        # (qg_search_exhaustive: every row's exact distance + full sort), then the merged list == the
        # merge of the per-rank oracle lists
        n_par = min(Q, 4)
        xd, xr, xc = shard.search_exhaustive(qh[:n_par], k)
        mine = torch.from_numpy(np.where(np.arange(k)[None, :] < xc[:, None],
                                         (capi_ordered(xd).astype(np.uint64) << np.uint64(32)) |
                                         (xr + row0).astype(np.uint64), np.uint64(0xFFFFFFFFFFFFFFFF)).view(np.int64)).to(dev)
        allk = torch.empty((world, n_par, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allk.view(world * n_par, k), mine)
        from quiver_b200 import sharded
        md, mr, mc = sharded.merge_keys(allk.cpu().numpy().view(np.uint64), k)
        gd, gr, gc = dd.cpu().numpy(), dr.cpu().numpy(), dc.cpu().numpy()
        unc = int((gc < 0).sum())
        assert np.array_equal(gr[:n_par], mr) and np.array_equal(gd[:n_par].view(np.uint32), md.view(np.uint32)), \
            (gr[:n_par], mr)
        bytes_rank = stats["bytes_algorithmic"] * stats["passes"]
        recs.append({"queries_per_step": Q, "ms_per_step": ms, "qps": Q / (ms * 1e-3),
                     "shard_scan_ms": scan_ms, "allgather_merge_us": max(0.0, (ms - scan_ms) * 1e3),
                     "path": {0: "exhaustive", 1: "flat scan", 2: "gather scan", 3: "tensor-core"}[stats["path"]],
                     "passes": stats["passes"],
                     "aggregate_GBps": world * bytes_rank / (ms * 1e-3) / 1e9,
                     "frac_of_aggregate_measured_hbm": bytes_rank / (ms * 1e-3) / 1e9 / hbm,
                     "uncertified": unc, "_rows": gr[0].copy(), "_dist": gd[0].copy(),
                     "parity": f"{n_par} queries: merged result == merge of the per-shard exhaustive GPU oracle lists "
                               f"(every row's exact distance + full sort on each rank), bit-identical"})
    shard.close()
    # The same single query over shards WITHOUT the bf16 copy (qg_config.flags = QG_FLAG_NO_BF16_COPY): the flat
    # fp32 stream of north_star's regime (a) — every GPU reads its 4.8 GB of fp32 rows once — as HBM evidence.
    # (With the copy, the default above, a single query streams half the bytes and is faster, see ms_per_step.)
    flat = capi.Index(d, mid, device=local_rank, reserve_rows=per, flags=capi.FLAG_NO_BF16_COPY)
    flat.upload_synthetic(3, 42, row0, per)
    qh = oracle.synth(3, 9999, 0, 1, d, threads=1)
    dq = torch.from_numpy(qh).to(dev)
    dd = torch.empty((1, k), dtype=torch.float32, device=dev)
    dr = torch.empty((1, k), dtype=torch.int64, device=dev)
    dc = torch.empty((1,), dtype=torch.int32, device=dev)
    for _ in range(3):
        comm.search_rows_device(flat, dq.data_ptr(), 1, k, row0, dd.data_ptr(), dr.data_ptr(), dc.data_ptr(), stream=st)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for _ in range(20):
        comm.search_rows_device(flat, dq.data_ptr(), 1, k, row0, dd.data_ptr(), dr.data_ptr(), dc.data_ptr(), stream=st)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fms = float(t.item())
    fst = flat.stats()
    fbytes = fst["bytes_algorithmic"] * fst["passes"]
    same = bool(np.array_equal(dr.cpu().numpy()[0], recs[0]["_rows"]) and
                np.array_equal(dd.cpu().numpy()[0].view(np.uint32), recs[0]["_dist"].view(np.uint32)))
    assert same, "flat fp32 stream and bf16 stream disagree"
    recs[0]["fp32_flat_stream"] = {"ms_per_step": fms, "path": {1: "flat scan", 3: "tensor-core"}.get(fst["path"], str(fst["path"])),
                                   "bytes_per_gpu": fbytes, "aggregate_GBps": world * fbytes / (fms * 1e-3) / 1e9,
                                   "frac_of_aggregate_measured_hbm": fbytes / (fms * 1e-3) / 1e9 / hbm,
                                   "uncertified": int((dc.cpu().numpy() < 0).sum()),
                                   "parity": "bit-identical to the default index's result for this query"}
    for r in recs:
        r.pop("_rows", None)
        r.pop("_dist", None)
    flat.close()
    return {"layout": "rows (north_star): per-shard top-k + NCCL all-gather of the keys + merge, all inside libquivergpu",
            "rows_per_gpu": per, "rows_total": per * world, "dim": d, "k": k, "metric": "l2",
            "data": "synthetic kind 3 (approx normal, L2-normalised: Deep-shaped)", "scaling": "weak",
            "hbm_peak_per_gpu_GBps": hbm, "batches": recs}


def capi_ordered(d):
    b = np.ascontiguousarray(d, dtype=np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def run_native(args):
    import torch
    import torch.distributed as dist
    from quiver_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and world > 1:
        args.gpus = world
    capi.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner (and anything NCCL_DEBUG asks for) to stdout by default; stdout
        # carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        # the data-path collective lives inside libquivergpu (qg_comm_*); torch.distributed only carries the
        # communicator id, the barriers and the max-over-ranks of the timings
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = capi.Comm(uid[0], world, rank, local_rank)

    import oracle  # checker + cpu_baseline only
    oracle.build()

    mid = METRIC_ID[args.metric]
    Q, k, d = args.queries, args.k, args.dim
    from quiver_b200 import sharded
    layout = "rows" if world == 1 else (args.shard if args.shard != "auto" else
                                        sharded.choose_layout(args.rows, d, Q, world))
    by_queries = layout == "queries"
    q0, q1 = 0, Q
    qper = Q
    if by_queries:
        # replicated corpus, this rank answers a contiguous slice of the batch
        row0, nloc = 0, args.rows
        qper = (Q + world - 1) // world
        q0, q1 = sharded.query_range(Q, world, rank)
    else:
        # row shard of this rank (contiguous block)
        per = (args.rows + world - 1) // world
        row0 = min(args.rows, rank * per)
        nloc = max(0, min(args.rows, row0 + per) - row0)
    idx = capi.Index(d, mid, device=local_rank, reserve_rows=max(nloc, 1))
    idx.upload_synthetic(args.kind, args.seed, row0, nloc)

    q_host = make_queries(oracle, args, Q)
    q_pin = torch.from_numpy(q_host).pin_memory()
    dq = q_pin.to(dev)
    st = torch.cuda.current_stream().cuda_stream
    d_dist = torch.empty((Q, k), dtype=torch.float32, device=dev)
    d_row = torch.empty((Q, k), dtype=torch.int64, device=dev)
    d_cnt = torch.empty((Q,), dtype=torch.int32, device=dev)

    def step_device(gather=True):
        if world == 1:
            idx.search_device(dq.data_ptr(), Q, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(), stream=st)
        elif by_queries:
            comm.search_queries_device(idx, dq.data_ptr(), Q, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(),
                                       stream=st, gather=gather)
        else:
            comm.search_rows_device(idx, dq.data_ptr(), Q, k, row0, d_dist.data_ptr(), d_row.data_ptr(),
                                    d_cnt.data_ptr(), stream=st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate before any timing is reported ------------------------------------------------
    step_device()
    torch.cuda.synchronize()
    checked = None
    host_corpus = None
    uncertified = 0  # queries the device API returned with count -1 (not counted as served)
    if not args.no_check and rank == 0:
        nchk = min(Q, 256)
        chk_ids = sorted(set(int(x) for x in np.linspace(0, Q - 1, nchk)))
        got_d, got_r = d_dist.cpu().numpy(), d_row.cpu().numpy()
        cnt_host = d_cnt.cpu().numpy()
        uncertified = int((cnt_host < 0).sum())
        # the device API marks a query whose selection could not be proven with count -1 (the host API
        # re-runs it through the flat scan); a sampled threshold makes that rare, not impossible
        assert uncertified <= max(1, Q // 2000), f"{uncertified} uncertified queries in the device-API run"
        # full oracle check when the corpus fits comfortably in host memory, else the GPU-side oracle
        if args.rows * d <= 300_000_000:
            corpus_chk = oracle.synth(args.kind, args.seed, 0, args.rows, d, threads=min(16, host_cores()))
            host_corpus = corpus_chk
            assert ((cnt_host == min(k, args.rows)) | (cnt_host < 0)).all()
            n_ok = check_against_oracle(oracle, corpus_chk, q_host, chk_ids, k, mid, got_d, got_r, cnt_host)
            checked = (f"{n_ok} queries spread over the batch bit-identical (rows and float32 distances) to the "
                       f"CPU oracle over all {args.rows} rows; {Q - uncertified} of {Q} queries certified in the "
                       f"device-API run")
        elif world == 1:
            # full-size membership: the exhaustive GPU oracle (qg_search_exhaustive: every row's exact distance
            # in the reference's arithmetic + full sort; itself checked against the CPU oracle in tests/)
            ids = [i for i in chk_ids[:32] if cnt_host[i] >= 0]
            xd, xr, xc = idx.search_exhaustive(q_host[ids], k)
            assert np.array_equal(got_r[ids], xr) and np.array_equal(got_d[ids].view(np.uint32), xd.view(np.uint32))
            checked = (f"{len(ids)} queries bit-identical to the exhaustive GPU oracle over all {args.rows} rows "
                       f"(every row's exact distance + full sort)")
        else:
            checked = "see row_sharded.batches[].parity (per-shard exhaustive GPU oracle)"

    # ---- value: device-resident, CUDA events on the launching stream ----------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    # the same K steps once more with an event pair around every kernel (prep / scan / finalize): the
    # per-kernel durations behind the roofline. Kept out of `value`: ~500 event records per step cost ~10 %.
    idx.read_profile()
    idx.set_profiling(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        step_device()
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1)
    clk = clocks.stop() if rank == 0 else None
    idx.set_profiling(False)
    prof = idx.read_profile()
    stats = idx.stats()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    qps = (Q - uncertified) / (ms_step * 1e-3)
    Q_gpu = (q1 - q0) if (world > 1 and by_queries) else Q  # queries one GPU's kernels serve per step

    # ---- e2e: host buffers through the C ABI call a Go caller would make ----------------------------------
    e2e_steps = max(10, min(args.steps, 100))
    if world == 1:
        e2e_out = [None]

        def step_e2e():
            e2e_out[0] = idx.search(q_host, k, out=e2e_out[0])  # result buffers reused from step to step
            return e2e_out[0]
        h2d, d2h = Q * d * 4, Q * k * 12 + Q * 4
        e2e_api = "qg_search_batch (host buffers, pinned staging)"
    else:
        # results land in pinned host buffers (a pageable .cpu() would add a staging copy per step); every
        # query's result reaches host memory exactly once: on the rank that answered it (replicas) or on
        # rank 0 (row shards, after the merge)
        nq_out = (q1 - q0) if by_queries else (Q if rank == 0 else 0)
        h_dist = torch.empty((max(nq_out, 1), k), dtype=torch.float32).pin_memory()
        h_row = torch.empty((max(nq_out, 1), k), dtype=torch.int64).pin_memory()
        h_cnt = torch.empty((max(nq_out, 1),), dtype=torch.int32).pin_memory()

        def step_e2e():
            if by_queries:
                if q1 > q0:
                    dq[q0:q1].copy_(q_pin[q0:q1], non_blocking=True)
                step_device(gather=False)  # no collective: the rank's block stays where it was computed
                if q1 > q0:
                    h_dist[:q1 - q0].copy_(d_dist[q0:q1], non_blocking=True)
                    h_row[:q1 - q0].copy_(d_row[q0:q1], non_blocking=True)
                    h_cnt[:q1 - q0].copy_(d_cnt[q0:q1], non_blocking=True)
            else:
                dq.copy_(q_pin, non_blocking=True)
                step_device()
                if rank == 0:
                    h_dist.copy_(d_dist, non_blocking=True)
                    h_row.copy_(d_row, non_blocking=True)
                    h_cnt.copy_(d_cnt, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d = Q * d * 4 if by_queries else world * Q * d * 4
        d2h = Q * k * 12 + Q * 4
        e2e_api = ("per rank: pinned H2D of its query block + qg_comm_search_queries_device(gather=0) + D2H of its "
                   "result block (no collective)" if by_queries else
                   "per rank: pinned H2D of the batch + qg_comm_search_rows_device (shard scan, NCCL all-gather, "
                   "merge); rank 0 copies the merged result to the host")
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_qps = Q * e2e_steps / float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    # ---- north_star's layout in the driver's record: C4-shaped row shards (every rank takes part) ----
    row_sharded = None
    if world > 1 and not args.no_subrecords:
        idx_keep = idx
        row_sharded = row_sharded_record(capi, oracle, torch, dist, comm, world, rank, local_rank, dev, peaks)
        idx = idx_keep

    if rank != 0:
        if world > 1:
            comm.close()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the scan), timed live by events inside the timed region ----
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    scan_launch_ms = prof["scan_ms"] / max(1, prof["scan_launches"])
    bytes_per_launch = stats["bytes_algorithmic"]  # rows*dim*4 (+ row-term / mask columns read); one corpus pass
    achieved = bytes_per_launch / (scan_launch_ms * 1e-3) / 1e9 if scan_launch_ms > 0 else 0.0
    tc = stats["path"] == 3
    bf16_stream = tc and stats.get("reserved", 0) == 1
    kernel_name = ("tc_ts_kernel (main scan, tcgen05.mma kind::%s, queries resident in TMEM)" % ("f16 on a bf16 copy of the corpus" if bf16_stream else "tf32")) if tc else "scan_dense_kernel"
    total_ms = ms_profiled
    hbm = {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
           "frac_of_nominal_8TBs": achieved / 8000.0}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": kernel_name, "launch_ms": scan_launch_ms,
                "bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
                "timed": f"CUDA events around every launch during a second pass of the same {args.steps} steps "
                         f"({ms_profiled / args.steps:.3f} ms per step with the events, {ms_step:.3f} without)",
                "scan_share_of_step": prof["scan_ms"] / total_ms if total_ms > 0 else None,
                "prep_share_of_step": prof["prep_ms"] / total_ms if total_ms > 0 else None,
                "finalize_share_of_step": prof["finalize_ms"] / total_ms if total_ms > 0 else None,
                "finalize_launch_ms": prof["finalize_ms"] / max(1, prof["finalize_launches"]),
                "prep_launch_ms": prof["prep_ms"] / max(1, prof["prep_launches"]) if prof["prep_launches"] else None}
    if tc:
        # a [rows x d] x [d x queries] contraction per pass: the binding roof is whichever floor is higher,
        # the corpus stream (HBM) or the MMA work (tensor pipe; tf32 runs at half the bf16 rate). The step is
        # tens of milliseconds at full clocks, not a seconds-long power-limited loop, so the tensor ceiling is
        # the BURST figure of MEASURED_PEAKS.json.
        qpp = min(Q_gpu, stats["queries_per_pass"])
        flops = 2.0 * qpp * nloc * d
        tfl = flops / (scan_launch_ms * 1e-3) / 1e12 if scan_launch_ms > 0 else 0.0
        bf16 = float(peaks.get("bf16_tflops", 1590.0))
        tpeak = bf16 if bf16_stream else bf16 / 2
        t_hbm, t_tensor = bytes_per_launch / (peak * 1e9), flops / (tpeak * 1e12)
        tensor = {"achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak, "queries_per_pass": qpp,
                  "flops_per_launch": flops,
                  "frac_of_sustained": tfl / float(peaks.get("bf16_tflops_sustained", 1400.0)) / (1 if bf16_stream else 0.5),
                  "peak_source": "measured burst (MEASURED_PEAKS.json bf16_tflops%s)" % ("" if bf16_stream else " / 2 for tf32")
                                 if "bf16_tflops" in peaks else "fallback 1590 TFLOP/s (B200_PROFILING.md)"}
        if t_tensor > t_hbm:
            roofline.update({"bound": "tensor", "achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak,
                             "peak_source": tensor["peak_source"], "flops_per_launch": flops,
                             "frac_of_sustained": tensor["frac_of_sustained"], "hbm": hbm})
        else:
            roofline["tensor"] = tensor
        roofline["floors_us"] = {"hbm": t_hbm * 1e6, "tensor": t_tensor * 1e6}
    if world == 1 and not args.no_subrecords:
        # one ncu launch of the same kernel on the same workload, outside the timed regions
        tr = measure_traffic(args, "tc_ts_kernel" if tc else "scan_dense_kernel", 3 if tc else 2)
        roofline["traffic"] = tr
        roofline["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu run of "
                                      "tools/prof_once.py inside this bench run" if tr else None)
    if roofline["traffic"] is None:
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            roofline["traffic"] = tr.get("tc_ts_kernel_bytes_per_launch" if tc else "scan_dense_kernel_bytes_per_launch")
            roofline["traffic_source"] = "profiles/traffic.json (an earlier ncu --set full capture; ncu could not run here)"
        except Exception:
            pass

    # ---- the small-batch regime: one query per call ----
    # (a) the fp32 flat scan (an index WITHOUT the bf16 copy, qg_config.flags = QG_FLAG_NO_BF16_COPY): the HBM
    #     roofline evidence north_star asks for; (b) the default index, whose single queries stream the bf16 copy
    #     through the tensor-core kernel (half the bytes, same exact results): the latency a caller sees.
    small = None
    if world == 1 and Q > 1:
        def one_query_record(index):
            index.read_profile()
            index.set_profiling(True)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(5):
                index.search_device(dq.data_ptr(), 1, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(), stream=st)
            torch.cuda.synchronize()
            index.read_profile()
            index.set_profiling(False)
            nsm = 100
            s0.record()
            for _ in range(nsm):
                index.search_device(dq.data_ptr(), 1, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(), stream=st)
            s1.record()
            torch.cuda.synchronize()
            per_call_ms = s0.elapsed_time(s1) / nsm  # without the per-kernel events
            index.set_profiling(True)
            for _ in range(nsm):
                index.search_device(dq.data_ptr(), 1, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(), stream=st)
            torch.cuda.synchronize()
            index.set_profiling(False)
            pr1 = index.read_profile()
            st1 = index.stats()
            l_ms = pr1["scan_ms"] / max(1, pr1["scan_launches"])
            gbps = st1["bytes_algorithmic"] / (l_ms * 1e-3) / 1e9 if l_ms > 0 else None
            kernel_label = ("tc_ts_kernel (bf16 copy of the corpus through the tensor-core scan, one query)" if st1["path"] == 3
                            else "scan_dense_kernel (fp32 rows, bulk-async tiles, one query)")
            return {"queries_per_step": 1, "qps": 1e3 / per_call_ms, "ms_per_query": per_call_ms, "path": st1["path"],
                    "kernel": kernel_label, "launch_ms": l_ms, "bytes_per_launch": st1["bytes_algorithmic"],
                    "achieved_GBps": gbps, "frac_of_measured_hbm": gbps / peak if gbps else None,
                    "uncertified": int((d_cnt[:1] < 0).sum().item())}

        flat_idx = capi.Index(d, mid, device=dev.index or 0, reserve_rows=args.rows, flags=capi.FLAG_NO_BF16_COPY)
        flat_idx.upload_synthetic(args.kind, args.seed, 0, args.rows)
        small = one_query_record(flat_idx)
        flat_idx.close()
        small["default_index"] = one_query_record(idx)

    real_valued = None
    if world == 1 and not args.no_subrecords and args.rows * d <= 300_000_000:
        real_valued = real_valued_record(capi, oracle, torch, args, dev, peaks)

    cpu_base = None
    if not args.no_cpu_baseline:
        corpus = host_corpus if host_corpus is not None else \
            oracle.synth(args.kind, args.seed, 0, args.rows, d, threads=min(16, host_cores()))
        cpu_base, _ = cpu_reference_qps(oracle, args, corpus, steps=2, warmup=1, budget_s=args.cpu_seconds)

    launches_per_step = stats["kernel_launches"] + (1 if (world > 1 and not by_queries) else 0)  # + merge kernel
    line = {
        "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 (returned distances: the reference's f32/f64 arithmetic; candidate selection: bf16 tensor-core scan for batches, f32 flat scan for single queries)", "data": "synthetic",
        "config": {"workload": f"flat {args.metric} exact search {args.rows}x{d} fp32 (SIFT-shaped synthetic, "
                               f"oracle/synth.h kind {args.kind} seed {args.seed}), query batch {Q}, k={k}",
                   "rows": args.rows, "dim": d, "k": k, "queries_per_step": Q,
                   "parallelism": "single GPU" if world == 1 else (
                       f"corpus replicated x{world} ({args.rows*d*6/1e9:.2f} GB per GPU with the bf16 copy), batch split "
                       f"in contiguous blocks of {qper} queries, one NCCL all-gather of the result blocks inside "
                       f"libquivergpu (qg_comm_search_queries_device); the row-sharded north_star layout is measured "
                       f"in row_sharded" if by_queries else
                       f"row-sharded x{world}, NCCL all-gather of per-shard top-k inside libquivergpu "
                       f"(qg_comm_search_rows_device)"),
                   "l2_policy": f"inputs larger than L2: every step streams the {args.rows*d*4/1e6:.0f} MB corpus "
                                f"({nloc*d*4/1e6:.0f} MB per GPU)",
                   "parity_check": checked},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": e2e_api},
        "gpu_launches": launches_per_step * args.steps,
        "kernels_per_step": {"passes": stats["passes"], "launches": stats["kernel_launches"],
                             "path": {0: "exhaustive", 1: "flat scan", 2: "gather scan", 3: "tensor-core"}[stats["path"]],
                             "queries_per_pass": stats["queries_per_pass"],
                             "merge": 1 if (world > 1 and not by_queries) else 0},
        "uncertified_queries_device_api": uncertified,
        "small_batch_regime": small,
        "real_valued": real_valued,
        "row_sharded": row_sharded,
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_base,
        "host_cores": host_cores(), "device": capi.device_info(local_rank)["name"],
    }
    emit(line)
    if world > 1:
        comm.close()
        dist.destroy_process_group()


_RESULT_OUT = None


def _claim_stdout():
    """stdout carries exactly one JSON line: keep a private handle on the real stdout for it and point
    fd 1 at stderr, so library chatter (NCCL's version banner, compiler notes ...) cannot join it."""
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
