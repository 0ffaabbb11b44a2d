"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the node-band partition + packed-key min exchange that the
sharded online step performs inside the kernel, and the row partition of batch scoring, checked against the oracle."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, REPO


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module(PKG_NAME + ".sharding")
    from oracle import pyoracle as po

    W, H, D, n = 9, 7, 11, 40
    rng = np.random.default_rng(5)  # same stream on every rank: same map, same samples
    o = po.Oracle(W, H, D, po.STANDARD)
    o.random_initialize(3, 1.0)
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[7] = o.get_state()["mean"][0]  # exact hit on node 0
    n0, cnt = sh.node_band(W, H, rank, world)
    # ---- node sharding: each rank scans only its band, then one min over packed keys
    keys = np.empty(n, np.uint64)
    for r in range(n):
        d = o.all_dists(x[r]).astype(np.float32)[n0:n0 + cnt]
        k = sh.pack_key(d, np.arange(n0, n0 + cnt))
        keys[r] = k.min()
    t = torch.from_numpy(keys.view(np.int64).copy())  # keys < 2^63: signed min == unsigned min
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    gd, gnode = sh.unpack_key(t.numpy().view(np.uint64))
    ob, od = o.find_bmu(x)
    ok_nodes = np.array_equal(gnode, ob) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    # ---- row sharding: each rank scores its rows, results concatenated in rank order
    lo, hi = sh.row_shard(n, rank, world)
    mine = torch.from_numpy(o.find_bmu(x[lo:hi])[0].astype(np.int64))
    sizes = [sh.row_shard(n, r, world) for r in range(world)]
    parts = [torch.empty(b - a, dtype=torch.int64) for a, b in sizes]
    dist.all_gather(parts, mine)
    ok_rows = np.array_equal(torch.cat(parts).numpy().astype(np.uint32), ob)
    if rank == 0:
        q.put((ok_nodes, ok_rows, [sh.node_band(W, H, r, world) for r in range(world)]))
    dist.destroy_process_group()


def test_node_band_and_row_shard_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_nodes, ok_rows, bands = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_nodes, "min over packed band keys differs from the oracle's findBmu"
    assert ok_rows, "row-sharded scoring differs from the oracle"
    assert bands[0][0] == 0 and bands[0][0] + bands[0][1] == bands[1][0] and bands[1][0] + bands[1][1] == 63


def test_partition_helpers():
    sh = importlib.import_module(PKG_NAME + ".sharding")
    for n, w in [(10, 3), (7, 8), (100, 8), (0, 2)]:
        spans = [sh.row_shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for W, H, w in [(512, 512, 8), (9, 7, 2), (5, 3, 3)]:
        bands = [sh.node_band(W, H, r, w) for r in range(w)]
        assert bands[0][0] == 0 and sum(c for _, c in bands) == W * H and all(c % W == 0 and c > 0 for _, c in bands)
    d = np.array([1.5, 1.5, np.nan, 0.0], np.float32)
    k = sh.pack_key(d, np.array([5, 2, 1, 9]))
    assert k.argmin() == 3 and sorted(k.tolist())[1] == k[1]
    assert sh.pack_key(np.array([np.nan], np.float32), np.array([0]))[0] == 0
    out = sh.merge_owner_outputs([np.array([1.0, np.nan], np.float32), np.array([np.nan, 2.0], np.float32)])
    assert out.tolist() == [1.0, 2.0]
