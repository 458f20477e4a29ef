"""Worker of tests/test_gpu_group.py::test_comm_two_processes: one process per GPU, NCCL inside libquivergpu
(qg_comm_*), torch.distributed (gloo) only hands the 128-byte id around. Both layouts must reproduce a
single-GPU search of the same synthetic corpus bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quiver_b200 import capi, sharded  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    capi.load()
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = capi.Comm(uid[0], world, rank, local)
    n, d, k, Q = 200_001, 128, 10, 517
    g = torch.Generator().manual_seed(7)
    queries = torch.floor(torch.rand((Q, d), generator=g) * 218).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    # reference: the whole corpus on this rank's GPU
    full = capi.Index(d, capi.L2, device=local)
    full.upload_synthetic(1, 42, 0, n)
    rd = torch.empty((Q, k), dtype=torch.float32, device=dev)
    rr = torch.empty((Q, k), dtype=torch.int64, device=dev)
    rc = torch.empty((Q,), dtype=torch.int32, device=dev)
    full.search_device(queries.data_ptr(), Q, k, rd.data_ptr(), rr.data_ptr(), rc.data_ptr(), stream=st)
    torch.cuda.synchronize()
    ok = rc >= 0
    od = torch.empty_like(rd); orow = torch.empty_like(rr); oc = torch.empty_like(rc)
    # rows layout
    row0, row1 = sharded.shard_range(n, world, rank)
    shard = capi.Index(d, capi.L2, device=local)
    shard.upload_synthetic(1, 42, row0, row1 - row0)
    comm.search_rows_device(shard, queries.data_ptr(), Q, k, row0, od.data_ptr(), orow.data_ptr(), oc.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert torch.equal(orow[ok], rr[ok]) and torch.equal(od[ok].view(torch.int32), rd[ok].view(torch.int32)), "rows layout"
    assert int((oc < 0).sum()) == 0
    # queries layout, gathered
    od.zero_(); orow.zero_(); oc.zero_()
    comm.search_queries_device(full, queries.data_ptr(), Q, k, od.data_ptr(), orow.data_ptr(), oc.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert torch.equal(orow[ok], rr[ok]) and torch.equal(od[ok].view(torch.int32), rd[ok].view(torch.int32)), "queries layout"
    # queries layout without the collective: only this rank's block is written
    od.fill_(-1); orow.fill_(-7)
    comm.search_queries_device(full, queries.data_ptr(), Q, k, od.data_ptr(), orow.data_ptr(), oc.data_ptr(), stream=st,
                               gather=False)
    torch.cuda.synchronize()
    q0, q1 = sharded.query_range(Q, world, rank)
    mine = torch.zeros(Q, dtype=torch.bool, device=dev); mine[q0:q1] = True
    assert torch.equal(orow[mine & ok], rr[mine & ok]) and bool((orow[~mine] == -7).all()), "queries layout, own block"
    dist.barrier()
    comm.close()
    if rank == 0:
        print("comm worker ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
