"""Row-sharded search. CPU: the exchange protocol (shard ranges, packed keys, all-gather,
merge) over gloo with world_size 2 and 3, the per-shard search played by the oracle. GPU: two
shards of one corpus on one device merged by qg_merge_shard_keys_device equal the unsharded index
bit for bit."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, k, nq, metric, out):
    import torch.distributed as dist
    import torch
    import oracle
    from quiver_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    corpus = oracle.synth(0, 42, 0, n, d, threads=1)  # every rank could regenerate any row
    queries = oracle.synth(0, 9999, 0, nq, d, threads=1)
    row0, row1 = sharded.shard_range(n, world, rank)
    shard = corpus[row0:row1]
    dist_l = np.full((nq, k), np.inf, dtype=np.float32)
    row_l = np.full((nq, k), -1, dtype=np.int64)
    cnt_l = np.zeros(nq, dtype=np.int32)
    for i in range(nq):
        if row1 > row0:
            od, orow = oracle.exact_search(shard, queries[i], k, metric)
            dist_l[i, :len(od)], row_l[i, :len(od)], cnt_l[i] = od, orow, len(od)
    keys = sharded.pack_keys(dist_l, row_l, cnt_l, row0)
    mine = torch.from_numpy(keys.view(np.int64).copy())
    allk = torch.empty((world * nq, k), dtype=torch.int64)  # rank-major concatenation
    dist.all_gather_into_tensor(allk, mine)
    gd, gr, gc = sharded.merge_keys(allk.numpy().view(np.uint64).reshape(world, nq, k), k)
    ok = True
    for i in range(nq):
        od, orow = oracle.exact_search(corpus, queries[i], k, metric)
        ok &= gc[i] == len(od) and np.array_equal(gr[i, :len(od)], orow) and \
            np.array_equal(gd[i, :len(od)].view(np.uint32), od.view(np.uint32))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5000), (3, 1001), (2, 1)])
def test_sharded_exchange_over_gloo(oracle, world, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 32, 10, 4, 1, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


def test_key_packing_orders_like_floats():
    from quiver_b200 import sharded
    d = np.array([-np.inf, -3.5, -0.0, 0.0, 1e-30, 1.0, 2.5, np.inf], dtype=np.float32)
    o = sharded.f32_to_ordered(d)
    assert np.all(np.diff(o.astype(np.int64)) >= 0)
    assert np.array_equal(sharded.ordered_to_f32(o).view(np.uint32), d.view(np.uint32))
    assert sharded.shard_range(10, 3, 0) == (0, 4) and sharded.shard_range(10, 3, 2) == (8, 10)
    assert sharded.shard_range(2, 4, 3) == (2, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("nq", [3, 40])
def test_two_shards_on_one_device_equal_the_unsharded_index(capi, oracle, nq):
    import torch
    from quiver_b200 import sharded
    n, d, k = 70000, 96, 10
    corpus = oracle.synth(3, 42, 0, n, d)
    queries = oracle.synth(3, 9999, 0, nq, d)
    full = capi.Index(d, 1)
    full.upload(corpus)
    fd, fr, fc, _ = full.search(queries, k)
    dq = torch.from_numpy(queries).cuda()
    world = 2
    allk = torch.empty((world, nq, k), dtype=torch.int64, device="cuda:0")
    shards = []
    for g in range(world):
        r0, r1 = sharded.shard_range(n, world, g)
        idx = capi.Index(d, 1)
        idx.upload(corpus[r0:r1])
        idx.search_shard_keys_device(dq.data_ptr(), nq, k, r0, allk[g].data_ptr())
        shards.append(idx)
    torch.cuda.synchronize()
    od = torch.empty((nq, k), dtype=torch.float32, device="cuda:0")
    orow = torch.empty((nq, k), dtype=torch.int64, device="cuda:0")
    oc = torch.empty((nq,), dtype=torch.int32, device="cuda:0")
    capi.merge_shard_keys_device(0, allk.data_ptr(), world, nq, k, od.data_ptr(), orow.data_ptr(), oc.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(orow.cpu().numpy(), fr)
    assert np.array_equal(od.cpu().numpy().view(np.uint32), fd.view(np.uint32))
    assert np.array_equal(oc.cpu().numpy(), fc)
    # the numpy merge used by the gloo test agrees with the device merge
    gd, gr, gc = sharded.merge_keys(allk.cpu().numpy().view(np.uint64), k)
    assert np.array_equal(gr, fr) and np.array_equal(gd.view(np.uint32), fd.view(np.uint32))
    for idx in shards + [full]:
        idx.close()


def _worker_queries(rank, world, port, n, d, k, nq, metric, out):
    """The "queries" layout: every rank holds the corpus, answers its slice, one all-gather of blocks."""
    import torch.distributed as dist
    import torch
    import oracle
    from quiver_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    corpus = oracle.synth(0, 42, 0, n, d, threads=1)
    queries = oracle.synth(0, 9999, 0, nq, d, threads=1)
    per = (nq + world - 1) // world
    q0, q1 = sharded.query_range(nq, world, rank)
    mine = torch.zeros(sharded.ReplicatedIndex.block_bytes(per, k), dtype=torch.uint8)
    rows, dists, cnts = sharded.unpack_block(mine, per, k)
    for i in range(q0, q1):
        od, orow = oracle.exact_search(corpus, queries[i], k, metric)
        dists[i - q0, :len(od)] = torch.from_numpy(od)
        rows[i - q0, :len(od)] = torch.from_numpy(orow.astype(np.int64))
        cnts[i - q0] = len(od)
    allb = torch.empty(world * mine.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(allb, mine)
    gd = torch.empty((nq, k), dtype=torch.float32)
    gr = torch.empty((nq, k), dtype=torch.int64)
    gc = torch.empty((nq,), dtype=torch.int32)
    sharded.scatter_blocks(allb, world, per, nq, k, gd, gr, gc)
    ok = True
    for i in range(nq):
        od, orow = oracle.exact_search(corpus, queries[i], k, metric)
        ok &= int(gc[i]) == len(od) and np.array_equal(gr[i, :len(od)].numpy(), orow) and \
            np.array_equal(gd[i, :len(od)].numpy().view(np.uint32), od.view(np.uint32))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nq", [(2, 7), (3, 10), (3, 2)])
def test_query_split_exchange_over_gloo(oracle, world, nq):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_queries, args=(r, world, port, 500, 16, 5, nq, 1, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


def test_layout_choice():
    from quiver_b200 import sharded
    assert sharded.choose_layout(1_000_000, 128, 10_000, 8) == "queries"     # 0.77 GB corpus, 1250 queries per GPU
    assert sharded.choose_layout(1_000_000, 128, 1, 8) == "rows"            # a single query is shortened by rows
    assert sharded.choose_layout(100_000_000, 96, 1024, 8) == "rows"        # 57.6 GB with the bf16 copy: shard it
    assert sharded.choose_layout(1_000_000, 128, 10_000, 1) == "rows"
    assert sharded.query_range(10, 3, 2) == (8, 10) and sharded.query_range(2, 3, 2) == (2, 2)


@pytest.mark.gpu
def test_shard_keys_repeat_uncertified_queries(capi, oracle):
    """The packed keys of a shard travel without counts, so a query whose selection could not be
    certified must already be exact when qg_search_shard_keys_device returns: adversarial rows (the
    neighbours of one query clustered in one stretch, 400 near-duplicates) through the keys API."""
    import torch
    from quiver_b200 import sharded
    rng = np.random.default_rng(6)
    n, d, k, nq = 40000, 64, 10, 16
    corpus = rng.random((n, d), dtype=np.float32)
    corpus = np.ascontiguousarray(corpus[np.argsort(corpus[:, 0])])
    queries = rng.random((nq, d), dtype=np.float32)
    corpus[1000:1400] = queries[0] + 1e-3 * rng.standard_normal((400, d)).astype(np.float32)
    idx = capi.Index(d, 1)
    idx.upload(corpus)
    dq = torch.from_numpy(queries).cuda()
    keys = torch.empty((1, nq, k), dtype=torch.int64, device="cuda:0")
    idx.search_shard_keys_device(dq.data_ptr(), nq, k, 0, keys.data_ptr())
    torch.cuda.synchronize()
    gd, gr, gc = sharded.merge_keys(keys.cpu().numpy().view(np.uint64), k)
    for i in range(nq):
        od, orow = oracle.exact_search(corpus, queries[i], k, 1)
        assert gc[i] == k and np.array_equal(gr[i], orow), (i, gr[i], orow, idx.stats())
        assert np.array_equal(gd[i].view(np.uint32), od.view(np.uint32))
    idx.close()
