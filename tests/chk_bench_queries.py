"""Development check (run by hand on a GPU box): how many of the bench batch's 10 000 queries come back
uncertified from the device API, and why (candidate count against the sampled threshold)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from quiver_b200 import capi
import oracle
oracle.build()
n, d, Q, k = 1_000_000, 128, 10000, 10
idx = capi.Index(d, 1, reserve_rows=n)
idx.upload_synthetic(1, 42, 0, n)
q = oracle.synth(1, 9999, 0, Q, d, threads=8)
dq = torch.from_numpy(q).cuda()
dist = torch.empty((Q, k), dtype=torch.float32, device="cuda"); row = torch.empty((Q, k), dtype=torch.int64, device="cuda"); cnt = torch.empty((Q,), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    idx.search_device(dq.data_ptr(), Q, k, dist.data_ptr(), row.data_ptr(), cnt.data_ptr(), stream=st)
    torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    bad = np.nonzero(c < 0)[0]
    print("rep", rep, "bad", bad.size, bad[:10], idx.stats())
if bad.size:
    tau, cn, cand = idx.debug_tc_pass(q[(bad[0]//256)*256:(bad[0]//256)*256+256], k)
    j = bad[0] % 256
    print("query", bad[0], "tau", tau[j], "cand count", cn[j], "mean cnt", cn.mean(), "max", cn.max())
