"""Host-side helpers for the two multi-GPU layouts.  No compute: partitioning and result merging only.

* rows   : batch scoring / evaluate / index shard by data rows, map replicated, no communication.
* nodes  : large-map training shards contiguous bands of grid rows across ranks; per sample the ranks exchange one packed
           64-bit (distance, node) key and take the minimum (inside the persistent kernel, over NVLink peer memory).
"""
from __future__ import annotations

import numpy as np


def row_shard(n_rows: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of `rank`; sizes differ by at most one row."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def node_band(width: int, height: int, rank: int, world: int):
    """(first_node, node_count) of the band of grid rows held by `rank` — same rule as vsom_create_sharded."""
    y0 = height * rank // world
    y1 = height * (rank + 1) // world
    return y0 * width, (y1 - y0) * width


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
