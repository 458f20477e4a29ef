"""Run a few device-resident searches (for ncu captures). argv: n d metric q k iters"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quiver_b200 import capi
n, d, metric, q, k, iters = [int(x) for x in sys.argv[1:7]]
idx = capi.Index(d, metric, reserve_rows=n, flags=int(os.environ.get("PROF_FLAGS", "0")))
idx.upload_synthetic(1 if metric == 1 else 2, 42, 0, n)
dq = torch.floor(torch.rand((q, d), device="cuda:0") * 218) if metric == 1 else torch.randn((q, d), device="cuda:0")
dist = torch.empty((q, k), dtype=torch.float32, device="cuda:0")
row = torch.empty((q, k), dtype=torch.int64, device="cuda:0")
cnt = torch.empty((q,), dtype=torch.int32, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
for _ in range(iters):
    idx.search_device(dq.data_ptr(), q, k, dist.data_ptr(), row.data_ptr(), cnt.data_ptr(), stream=st)
torch.cuda.synchronize()
print(idx.stats(), cnt[:4].tolist())
