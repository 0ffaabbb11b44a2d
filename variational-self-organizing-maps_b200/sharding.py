"""Host-side helpers for the two multi-GPU layouts.  No compute: partitioning and result merging only.

* rows   : batch scoring / evaluate / index shard by data rows, map replicated, no communication.
* nodes  : large-map training deals blocks of a few grid rows to the ranks round-robin (block b -> rank b % world); per
           sample the ranks exchange one packed 64-bit (distance, node) key and take the minimum (inside the persistent
           kernel, over NVLink peer memory).  The U-matrix needs one halo row of means per block border.
* row-shard combines: bmuHits are summed over the ranks; Som::evaluate's running mean is order dependent in f64, so the
           per-row distances are gathered in rank order and folded on one rank exactly like the reference does.
"""
from __future__ import annotations

import numpy as np


def row_shard(n_rows: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of `rank`; sizes differ by at most one row."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


DEFAULT_BLOCK_ROWS = 4


def shard_block_rows(height: int, world: int, block: int = DEFAULT_BLOCK_ROWS) -> int:
    """Rows per block — same rule as vsom_create_sharded: every rank holds at least one row."""
    if world == 1:
        return height
    return block if block * world <= height else height // world


def node_rows(height: int, rank: int, world: int, block: int = DEFAULT_BLOCK_ROWS) -> np.ndarray:
    """Global grid rows held by `rank`, in local order: block b of `block` rows belongs to rank b % world."""
    b = shard_block_rows(height, world, block)
    y = np.arange(height)
    return y[(y // b) % world == rank]


def node_ids(width: int, height: int, rank: int, world: int, block: int = DEFAULT_BLOCK_ROWS) -> np.ndarray:
    """Global node ids held by `rank`, in local order."""
    return (node_rows(height, rank, world, block)[:, None] * width + np.arange(width)[None, :]).reshape(-1)


def halo_rows(height: int, rank: int, world: int, block: int = DEFAULT_BLOCK_ROWS) -> np.ndarray:
    """Grid rows of OTHER ranks that border this rank's rows (what the sharded U-matrix fetches)."""
    mine = set(node_rows(height, rank, world, block).tolist())
    out = sorted({y + d for y in mine for d in (-1, 1) if 0 <= y + d < height} - mine)
    return np.array(out, dtype=np.int64)


def sum_hits(per_rank_hits):
    """bmuHits of row-sharded scoring / training replicas: plain sum over the ranks (what an NCCL sum-allreduce gives)."""
    return np.sum(np.stack([np.asarray(h, np.uint64) for h in per_rank_hits]), axis=0, dtype=np.uint64)


def evaluate_from_shards(per_rank_dists) -> float:
    """Som::evaluate (src/Som.cpp:490-523) over row shards: the f64 running mean  e += (d_i - e) / (i + 1)  depends on the
    row order, so the shards' per-row BMU distances are concatenated in rank order and folded exactly like the reference."""
    err = 0.0
    i = 0
    for d in per_rank_dists:
        for v in np.asarray(d, np.float32):
            err += 1.0 / (i + 1.0) * (float(v) + 0.0 - err)
            i += 1
    return err


def pack_key(dist: np.ndarray, node: np.ndarray) -> np.ndarray:
    """(f32 distance >= 0, node id) -> sortable uint64: IEEE bits of the distance in the high word, node in the low word.
    min() over keys is findBmu's rule: smallest distance, lowest index on ties; NaN never wins unless at node 0."""
    d = np.ascontiguousarray(dist, np.float32)
    bits = d.view(np.uint32).astype(np.uint64)
    node = np.asarray(node, np.uint64)
    nan = np.isnan(d)
    bits = np.where(nan, np.where(node == 0, np.uint64(0), np.uint64(0x7FFFFFFF)), bits)
    return (bits << np.uint64(32)) | node


def unpack_key(key: np.ndarray):
    key = np.asarray(key, np.uint64)
    dist = (key >> np.uint64(32)).astype(np.uint32).view(np.float32)
    return dist, (key & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def merge_owner_outputs(per_rank_dist):
    """Per-sample distances reported only by the rank that owns the BMU (NaN elsewhere) -> one array."""
    stack = np.stack([np.asarray(d, np.float32) for d in per_rank_dist])
    owners = (~np.isnan(stack)).sum(axis=0)
    if not np.all(owners <= 1):
        raise ValueError("a sample was claimed by more than one rank")
    return np.nanmax(np.where(np.isnan(stack), -np.inf, stack), axis=0).astype(np.float32)
