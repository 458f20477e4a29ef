"""Quick device-resident timing sweep (development aid; bench.py is the judged harness)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quiver_b200 import capi  # noqa: E402


def time_search(idx, dq, q, k, iters=20, warm=3):
    dev = torch.device("cuda:0")
    dist = torch.empty((q, k), dtype=torch.float32, device=dev)
    row = torch.empty((q, k), dtype=torch.int64, device=dev)
    cnt = torch.empty((q,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(warm):
        idx.search_device(dq.data_ptr(), q, k, dist.data_ptr(), row.data_ptr(), cnt.data_ptr(), stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        idx.search_device(dq.data_ptr(), q, k, dist.data_ptr(), row.data_ptr(), cnt.data_ptr(), stream=st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, cnt.cpu().numpy()


def main():
    out = []
    configs = [(1_000_000, 128, 1, 1), (1_000_000, 96, 1, 3), (200_000, 768, 0, 2), (4_000_000, 128, 1, 1), (1_000_000, 128, 2, 1), (1_000_000, 128, 0, 0), (1_000_000, 128, 1, 0), (1_000_000, 128, 0, 1), (1_000_000, 128, 2, 0), (12_500_000, 96, 1, 3), (4_000_000, 256, 1, 2), (2_000_000, 768, 0, 2), (4_000_000, 192, 1, 2)]
    qs = (1, 2, 4, 8, 16, 64, 128, 256, 1024, 10000)
    if len(sys.argv) > 1:
        qs = tuple(int(x) for x in sys.argv[1].split(","))
    ks = (10, 100)
    if len(sys.argv) > 3:
        ks = tuple(int(x) for x in sys.argv[3].split(','))
    if len(sys.argv) > 2:
        configs = [configs[int(x)] for x in sys.argv[2].split(",")]
    for n, d, metric, kind in configs:
        idx = capi.Index(d, metric, reserve_rows=n)
        t0 = time.time()
        idx.upload_synthetic(kind, 42, 0, n)
        print(f"N={n} d={d} metric={metric} fill {time.time()-t0:.2f}s", flush=True)
        for q in qs:
            for k in ks:
                dq = torch.rand((q, d), device="cuda:0")
                if kind == 1:
                    dq = torch.floor(dq * 218)
                ms, cnt = time_search(idx, dq, q, k, iters=20 if q <= 256 else 5)
                st = idx.stats()
                gbs = st["bytes_algorithmic"] * st["passes"] / (ms * 1e-3) / 1e9
                tfl = 2.0 * q * n * d / (ms * 1e-3) / 1e12
                idx.read_profile()
                idx.set_profiling(True)
                time_search(idx, dq, q, k, iters=5, warm=0)
                idx.set_profiling(False)
                pr = idx.read_profile()
                prof = dict(scan_us=round(1e3 * pr["scan_ms"] / max(1, pr["scan_launches"]), 2),
                            fin_us=round(1e3 * pr["finalize_ms"] / max(1, pr["finalize_launches"]), 2),
                            prep_us=round(1e3 * pr["prep_ms"] / max(1, pr["prep_launches"]), 2))
                rec = dict(prof=prof, n=n, d=d, metric=metric, q=q, k=k, ms=round(ms, 4), qps=round(q / ms * 1e3, 1),
                           gbs=round(gbs, 1), tflops=round(tfl, 1), path=st["path"], passes=st["passes"], qb=st["queries_per_pass"],
                           launches=st["kernel_launches"], bad=int((cnt < 0).sum()))
                print(json.dumps(rec), flush=True)
                out.append(rec)
        idx.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/quickbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
