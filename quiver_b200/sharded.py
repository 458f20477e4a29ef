"""Row-sharded exact search across GPUs (no reference counterpart — Quiver is single
process). One process per GPU: rank g owns the contiguous block [g*ceil(N/G), (g+1)*ceil(N/G)),
queries are replicated, every rank emits its shard's top-k as packed 64-bit keys
(order-preserving image of the EXACT float32 distance << 32 | global row), one all-gather
exchanges Q*k*8 bytes per rank, and every rank merges the G lists.

The key format and the merge are pure integer work, restated here in numpy so that the exchange
protocol is testable with the gloo backend on CPU; on the GPU the keys come from
qg_search_shard_keys_device and the merge is qg_merge_shard_keys_device.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

KEY_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_range(rows: int, world: int, rank: int) -> Tuple[int, int]:
    """[row0, row1) of `rank`: contiguous blocks of ceil(rows / world)."""
    per = (rows + world - 1) // world
    row0 = min(rows, rank * per)
    return row0, min(rows, row0 + per)


def query_range(queries: int, world: int, rank: int) -> Tuple[int, int]:
    """[q0, q1) of `rank` when a batch is split over replicas: contiguous blocks of ceil(queries / world)."""
    per = (queries + world - 1) // world
    q0 = min(queries, rank * per)
    return q0, min(queries, q0 + per)


def choose_layout(rows: int, dim: int, queries: int, world: int, hbm_bytes: float = 180e9) -> str:
    """How a batch is spread over `world` GPUs of one box.

    "rows"    — the corpus is row-sharded, every GPU scans its shard for all queries, the per-shard
                top-k lists are all-gathered and merged (the only layout for corpora that do
                not fit one GPU and the one that shortens a single query).
    "queries" — every GPU holds the whole corpus and answers a contiguous slice of the batch; one
                all-gather of the results. This is the reference's own parallel structure
                (`BatchSearch`: one goroutine per query over one shared index, hybrid_index.go:703-795)
                and the right one when the corpus (fp32 rows + bf16 copy = 6 bytes per element) is a
                small fraction of one GPU's HBM and each GPU still gets full tensor-core passes."""
    if world <= 1:
        return "rows"
    fits = rows * dim * 6.0 <= 0.25 * hbm_bytes
    return "queries" if fits and queries >= 256 * world else "rows"


def f32_to_ordered(d: np.ndarray) -> np.ndarray:
    """Same mapping as csrc/common.cuh: unsigned order == float order (negatives flipped)."""
    b = np.ascontiguousarray(d, dtype=np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def ordered_to_f32(k: np.ndarray) -> np.ndarray:
    k = k.astype(np.uint32)
    b = np.where(k & np.uint32(0x80000000), k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32)
    return b.view(np.float32)


def pack_keys(dist: np.ndarray, row: np.ndarray, count: np.ndarray, row_base: int) -> np.ndarray:
    """[Q, k] distances / local rows / counts -> [Q, k] uint64 keys (missing entries = all ones)."""
    q, k = dist.shape
    keys = (f32_to_ordered(dist).astype(np.uint64) << np.uint64(32)) | (row.astype(np.int64) + row_base).astype(np.uint64)
    valid = np.arange(k)[None, :] < np.asarray(count)[:, None]
    return np.where(valid, keys, KEY_NONE)


def merge_keys(gathered: np.ndarray, k: int):
    """[G, Q, k] keys -> (dist [Q, k], row [Q, k], count [Q]): the k smallest keys per query."""
    g, q, kk = gathered.shape
    allk = np.sort(gathered.transpose(1, 0, 2).reshape(q, g * kk), axis=1)[:, :k]
    valid = allk != KEY_NONE
    dist = np.where(valid, ordered_to_f32((allk >> np.uint64(32)).astype(np.uint32)), np.float32(np.inf)).astype(np.float32)
    row = np.where(valid, (allk & np.uint64(0xFFFFFFFF)).astype(np.int64), -1)
    return dist, row, valid.sum(axis=1).astype(np.int32)


class ShardedIndex:
    """One rank's shard plus the collective. Needs torch.distributed initialised (backend nccl)."""

    def __init__(self, dim: int, metric: int, rows_total: int, device: int = 0):
        import torch
        import torch.distributed as dist
        from . import capi
        self.torch, self.dist, self.capi = torch, dist, capi
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.row0, self.row1 = shard_range(rows_total, self.world, self.rank)
        self.device = device
        self.index = capi.Index(dim, metric, device=device, reserve_rows=max(1, self.row1 - self.row0))

    def search_device(self, d_queries, q: int, k: int, d_dist, d_row, d_count, stream: int = 0):
        """All tensors on this rank's device; results (global rows) are valid on every rank."""
        torch = self.torch
        if self.world == 1:
            self.index.search_device(d_queries.data_ptr(), q, k, d_dist.data_ptr(), d_row.data_ptr(),
                                     d_count.data_ptr(), stream=stream)
            return
        keys = torch.empty((q, k), dtype=torch.int64, device=d_queries.device)
        allk = torch.empty((self.world * q, k), dtype=torch.int64, device=d_queries.device)  # rank-major
        self.index.search_shard_keys_device(d_queries.data_ptr(), q, k, self.row0, keys.data_ptr(), stream=stream)
        self.dist.all_gather_into_tensor(allk, keys)
        self.capi.merge_shard_keys_device(self.device, allk.data_ptr(), self.world, q, k, d_dist.data_ptr(),
                                          d_row.data_ptr(), d_count.data_ptr(), stream=stream)


class ReplicatedIndex:
    """The "queries" layout of choose_layout(): one full copy of the corpus per rank, the batch split in
    contiguous blocks, one all-gather of the packed result blocks. Needs torch.distributed (nccl)."""

    def __init__(self, dim: int, metric: int, rows_total: int, device: int = 0):
        import torch
        import torch.distributed as dist
        from . import capi
        self.torch, self.dist, self.capi = torch, dist, capi
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = device
        self.index = capi.Index(dim, metric, device=device, reserve_rows=max(1, rows_total))

    @staticmethod
    def block_bytes(per: int, k: int) -> int:
        return per * (k * 8 + k * 4 + 4)  # rows (int64) first for alignment, then distances, then counts

    def search_device(self, d_queries, q: int, k: int, d_dist, d_row, d_count, stream: int = 0):
        """d_queries holds all q queries on every rank; every rank ends with all q results."""
        torch = self.torch
        if self.world == 1:
            self.index.search_device(d_queries.data_ptr(), q, k, d_dist.data_ptr(), d_row.data_ptr(),
                                     d_count.data_ptr(), stream=stream)
            return
        per = (q + self.world - 1) // self.world
        q0, q1 = query_range(q, self.world, self.rank)
        dev = d_queries.device
        mine = torch.zeros(self.block_bytes(per, k), dtype=torch.uint8, device=dev)
        rows, dists, cnts = unpack_block(mine, per, k)
        if q1 > q0:
            self.index.search_device(d_queries[q0:q1].data_ptr(), q1 - q0, k, dists.data_ptr(), rows.data_ptr(),
                                     cnts.data_ptr(), stream=stream)
        allb = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=dev)
        self.dist.all_gather_into_tensor(allb, mine)
        scatter_blocks(allb, self.world, per, q, k, d_dist, d_row, d_count)


def unpack_block(block, per: int, k: int):
    """Views (rows int64 [per,k], dists float32 [per,k], counts int32 [per]) into one result block."""
    import torch
    rows = block[:per * k * 8].view(torch.int64).view(per, k)
    dists = block[per * k * 8:per * k * 12].view(torch.float32).view(per, k)
    cnts = block[per * k * 12:per * k * 12 + per * 4].view(torch.int32)
    return rows, dists, cnts


def scatter_blocks(allb, world: int, per: int, q: int, k: int, d_dist, d_row, d_count) -> None:
    """Rank-major concatenation of result blocks -> [q, k] outputs (the last blocks may be partial)."""
    bb = per * (k * 12 + 4)
    for r in range(world):
        q0, q1 = query_range(q, world, r)
        if q1 <= q0:
            continue
        rows, dists, cnts = unpack_block(allb[r * bb:(r + 1) * bb], per, k)
        d_row[q0:q1].copy_(rows[:q1 - q0])
        d_dist[q0:q1].copy_(dists[:q1 - q0])
        d_count[q0:q1].copy_(cnts[:q1 - q0])
