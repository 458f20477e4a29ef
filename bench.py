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
  N > 1   the corpus is row-sharded across the ranks (contiguous blocks), every rank scans its
          shard for the replicated query batch, the per-shard top-k keys are exchanged with one
          NCCL all-gather and merged on every rank (strong scaling: the corpus is fixed).
The oracle (oracle/) is used only as the checker and for the cpu_baseline / --impl reference legs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

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


def run_native(args):
    import numpy as np
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner (and anything NCCL_DEBUG asks for) to stdout by default; stdout
        # carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    import oracle  # checker + cpu_baseline only
    oracle.build()

    mid = METRIC_ID[args.metric]
    Q, k, d = args.queries, args.k, args.dim
    from quiver_b200 import sharded
    layout = "rows" if world == 1 else (args.shard if args.shard != "auto" else
                                        sharded.choose_layout(args.rows, d, Q, world))
    by_queries = layout == "queries"
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
    if world > 1 and by_queries:
        blk = torch.zeros(sharded.ReplicatedIndex.block_bytes(qper, k), dtype=torch.uint8, device=dev)
        b_row, b_dist, b_cnt = sharded.unpack_block(blk, qper, k)
        d_allb = torch.empty(world * blk.numel(), dtype=torch.uint8, device=dev)  # rank-major result blocks
    elif world > 1:
        d_keys = torch.empty((Q, k), dtype=torch.int64, device=dev)  # packed u64 keys
        d_all = torch.empty((world * Q, k), dtype=torch.int64, device=dev)  # rank-major concatenation

    def step_device():
        if world == 1:
            idx.search_device(dq.data_ptr(), Q, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(), stream=st)
        elif by_queries:
            # the results stay in block layout (per rank: rows | distances | counts); every rank gets all
            if q1 > q0:
                idx.search_device(dq[q0:q1].data_ptr(), q1 - q0, k, b_dist.data_ptr(), b_row.data_ptr(),
                                  b_cnt.data_ptr(), stream=st)
            dist.all_gather_into_tensor(d_allb, blk)
        else:
            idx.search_shard_keys_device(dq.data_ptr(), Q, k, row0, d_keys.data_ptr(), stream=st)
            dist.all_gather_into_tensor(d_all, d_keys)
            capi.merge_shard_keys_device(local_rank, d_all.data_ptr(), world, Q, k, d_dist.data_ptr(),
                                         d_row.data_ptr(), d_cnt.data_ptr(), stream=st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate before any timing is reported ------------------------------------------------
    step_device()
    torch.cuda.synchronize()
    if world > 1 and by_queries:
        sharded.scatter_blocks(d_allb, world, qper, Q, k, d_dist, d_row, d_cnt)
        torch.cuda.synchronize()
    checked = None
    host_corpus = None
    uncertified = 0  # queries the device API returned with count -1 (not counted as served)
    if not args.no_check and rank == 0:
        nchk = min(Q, 8)
        chk_ids = sorted(set(int(x) for x in np.linspace(0, Q - 1, nchk)))
        # full oracle check when the corpus fits comfortably in host memory, else a 200k-row prefix
        sub_rows = args.rows if args.rows * d <= 300_000_000 else 200_000
        if sub_rows == args.rows:
            corpus_chk = oracle.synth(args.kind, args.seed, 0, sub_rows, d, threads=min(16, host_cores()))
            host_corpus = corpus_chk
            got_d, got_r = d_dist.cpu().numpy(), d_row.cpu().numpy()
            cnt_host = d_cnt.cpu().numpy()
            uncertified = int((cnt_host < 0).sum())
            # the device API marks a query whose selection could not be proven with count -1 (the host API
            # re-runs it through the flat scan); a sampled threshold makes that rare, not impossible
            assert uncertified <= max(1, Q // 2000), f"{uncertified} uncertified queries in the device-API run"
            assert ((cnt_host == min(k, args.rows)) | (cnt_host < 0)).all()
            chk_ids = [i for i in chk_ids if cnt_host[i] >= 0]
            for i in chk_ids:
                od, orow = oracle.exact_search(corpus_chk, q_host[i], k, mid)
                assert np.array_equal(got_r[i, :len(orow)], orow), (i, got_r[i], orow)
                assert np.array_equal(got_d[i, :len(od)].view(np.uint32), od.view(np.uint32))
            checked = (f"{len(chk_ids)} queries spread over the batch bit-identical to the oracle over all "
                       f"{sub_rows} rows; {Q - uncertified} of {Q} queries certified in the device-API run")
        else:
            # full-size property: distances of the returned rows equal the oracle's pairwise arithmetic,
            # ascending, and no row of a 200k-row prefix beats the k-th result
            corpus_chk = oracle.synth(args.kind, args.seed, 0, sub_rows, d, threads=min(16, host_cores()))
            got_d, got_r = d_dist.cpu().numpy(), d_row.cpu().numpy()
            cnt_host = d_cnt.cpu().numpy()
            uncertified = int((cnt_host < 0).sum())
            assert uncertified <= max(1, Q // 2000), f"{uncertified} uncertified queries in the device-API run"
            chk_ids = [i for i in chk_ids if cnt_host[i] >= 0]
            for i in chk_ids:
                rows_i = got_r[i]
                vecs = np.stack([oracle.synth(args.kind, args.seed, int(r), 1, d, threads=1)[0] for r in rows_i])
                want = np.array([oracle.distance(mid, q_host[i], v) for v in vecs], dtype=np.float32)
                assert np.array_equal(want.view(np.uint32), got_d[i].view(np.uint32)), (want, got_d[i])
                assert np.all(np.diff(got_d[i]) >= 0)
                od, orow = oracle.exact_search(corpus_chk, q_host[i], k, mid)
                inside = rows_i < sub_rows
                assert od[0] >= got_d[i][0] and set(orow[od < got_d[i][-1]]).issubset(set(rows_i[inside]))
            checked = (f"{nchk} queries: returned distances bit-identical to the oracle's pairwise arithmetic, "
                       f"ascending, and consistent with the oracle's exact top-{k} over a {sub_rows}-row prefix")

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
    else:
        # results land in pinned host buffers (a pageable .cpu() would add a staging copy per step)
        if by_queries:
            h_allb = torch.empty(d_allb.numel(), dtype=torch.uint8).pin_memory()
        else:
            h_dist = torch.empty((Q, k), dtype=torch.float32).pin_memory()
            h_row = torch.empty((Q, k), dtype=torch.int64).pin_memory()
            h_cnt = torch.empty((Q,), dtype=torch.int32).pin_memory()

        def step_e2e():
            if by_queries:
                dq[q0:q1].copy_(q_pin[q0:q1], non_blocking=True)
                step_device()
                h_allb.copy_(d_allb, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return h_allb
            dq.copy_(q_pin, non_blocking=True)
            step_device()
            h_dist.copy_(d_dist, non_blocking=True)
            h_row.copy_(d_row, non_blocking=True)
            h_cnt.copy_(d_cnt, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return h_dist, h_row, h_cnt
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the scan), timed live by events inside the timed region ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    scan_launch_ms = prof["scan_ms"] / max(1, prof["scan_launches"])
    bytes_per_launch = stats["bytes_algorithmic"]  # rows*dim*4 (+ row-term / mask columns read); one corpus pass
    achieved = bytes_per_launch / (scan_launch_ms * 1e-3) / 1e9 if scan_launch_ms > 0 else 0.0
    tc = stats["path"] == 3
    bf16_stream = tc and stats.get("reserved", 0) == 1
    kernel_name = ("tc_ts_kernel (main scan, tcgen05.mma kind::%s, queries resident in TMEM)" % ("f16 on a bf16 copy of the corpus" if bf16_stream else "tf32")) if tc else "scan_fast_kernel"
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
        # the corpus stream (HBM) or the MMA work (tensor pipe; tf32 runs at half the bf16 rate)
        qpp = min(Q_gpu, stats["queries_per_pass"])
        flops = 2.0 * qpp * nloc * d
        tfl = flops / (scan_launch_ms * 1e-3) / 1e12 if scan_launch_ms > 0 else 0.0
        bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
        tpeak = bf16 if bf16_stream else bf16 / 2
        t_hbm, t_tensor = bytes_per_launch / (peak * 1e9), flops / (tpeak * 1e12)
        tensor = {"achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak, "queries_per_pass": qpp,
                  "flops_per_launch": flops,
                  "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained%s)" % ("" if bf16_stream else " / 2 for tf32")
                                 if "bf16_tflops_sustained" in peaks else "fallback 1400 TFLOP/s"}
        if t_tensor > t_hbm:
            roofline.update({"bound": "tensor", "achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak,
                             "peak_source": tensor["peak_source"], "flops_per_launch": flops, "hbm": hbm})
        else:
            roofline["tensor"] = tensor
        roofline["floors_us"] = {"hbm": t_hbm * 1e6, "tensor": t_tensor * 1e6}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        roofline["traffic"] = tr.get("tc_ts_kernel_bytes_per_launch" if tc else "scan_fast_kernel_bytes_per_launch")
    except Exception:
        pass

    # ---- the small-batch regime (flat scan, one query per pass) for the HBM roofline it is judged on ----
    small = None
    if world == 1 and Q > 1:
        idx.read_profile()
        idx.set_profiling(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            idx.search_device(dq.data_ptr(), 1, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(), stream=st)
        torch.cuda.synchronize()
        idx.read_profile()
        s0.record()
        nsm = 100
        for _ in range(nsm):
            idx.search_device(dq.data_ptr(), 1, k, d_dist.data_ptr(), d_row.data_ptr(), d_cnt.data_ptr(), stream=st)
        s1.record()
        torch.cuda.synchronize()
        idx.set_profiling(False)
        pr1 = idx.read_profile()
        st1 = idx.stats()
        l_ms = pr1["scan_ms"] / max(1, pr1["scan_launches"])
        small = {"queries_per_step": 1, "qps": nsm / (s0.elapsed_time(s1) * 1e-3), "kernel": "scan_fast_kernel",
                 "launch_ms": l_ms, "achieved_GBps": st1["bytes_algorithmic"] / (l_ms * 1e-3) / 1e9 if l_ms > 0 else None,
                 "frac_of_measured_hbm": st1["bytes_algorithmic"] / (l_ms * 1e-3) / 1e9 / peak if l_ms > 0 else None}

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
                       f"in contiguous blocks of {qper} queries, one NCCL all-gather of the result blocks"
                       if by_queries else f"row-sharded x{world}, NCCL all-gather of per-shard top-k"),
                   "l2_policy": f"inputs larger than L2: every step streams the {args.rows*d*4/1e6:.0f} MB corpus "
                                f"({nloc*d*4/1e6:.0f} MB per GPU)",
                   "parity_check": checked},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": Q * d * 4,
                "d2h_bytes_per_step": Q * k * 12 + Q * 4, "steps": e2e_steps,
                "api": "qg_search_batch (host buffers, pinned staging)" if world == 1 else (
                       "pinned H2D of the rank's query block + search + all-gather + D2H of all results" if by_queries
                       else "pinned H2D + shard search + all-gather + merge + D2H")},
        "gpu_launches": launches_per_step * args.steps,
        "kernels_per_step": {"passes": stats["passes"], "launches": stats["kernel_launches"],
                             "path": {0: "exhaustive", 1: "flat scan", 2: "gather scan", 3: "tensor-core"}[stats["path"]],
                             "queries_per_pass": stats["queries_per_pass"],
                             "merge": 1 if (world > 1 and not by_queries) else 0},
        "uncertified_queries_device_api": uncertified,
        "small_batch_regime": small,
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_base,
        "host_cores": host_cores(), "device": capi.device_info(local_rank)["name"],
    }
    emit(line)
    if world > 1:
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
