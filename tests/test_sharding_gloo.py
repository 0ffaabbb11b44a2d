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
    ids = sh.node_ids(W, H, rank, world, block=2)
    # ---- node sharding: each rank scans only its blocks of grid rows, then one min over packed keys
    keys = np.empty(n, np.uint64)
    for r in range(n):
        d = o.all_dists(x[r]).astype(np.float32)[ids]
        k = sh.pack_key(d, ids)
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
    # ---- row-shard combines: hits = sum over ranks, evaluate = f64 running mean folded in rank order
    hits = torch.from_numpy(np.bincount(mine.numpy(), minlength=W * H).astype(np.int64))
    dist.all_reduce(hits)
    ok_rows = ok_rows and np.array_equal(hits.numpy(), np.bincount(ob, minlength=W * H))
    dl = [None] * world
    dist.all_gather_object(dl, o.find_bmu(x[lo:hi])[1])
    ok_rows = ok_rows and sh.evaluate_from_shards(dl) == o.evaluate(x)
    # ---- sharded U-matrix: own rows + halo rows of means are all a rank needs
    st = o.get_state()
    st["sigma"] = np.abs(rng.standard_normal(st["sigma"].shape)).astype(np.float32)
    o.set_state(**st)
    want = o.update_umatrix()
    rows = sh.node_rows(H, rank, world, block=2)
    have = np.concatenate([rows, sh.halo_rows(H, rank, world, block=2)])
    part = po.Oracle(W, H, D, po.STANDARD)
    mean = np.full_like(st["mean"], np.nan).reshape(H, W, D)
    mean[have] = st["mean"].reshape(H, W, D)[have]        # everything else is NaN: touching it would poison the result
    sig = np.full_like(st["sigma"], np.nan).reshape(H, W, D)
    sig[rows] = st["sigma"].reshape(H, W, D)[rows]
    part.set_state(mean=mean.reshape(-1, D), sigma=sig.reshape(-1, D))
    got = part.update_umatrix().reshape(H, W)[rows]
    ok_um = np.array_equal(got.view(np.uint64), want.reshape(H, W)[rows].view(np.uint64))
    if rank == 0:
        q.put((ok_nodes, ok_rows and ok_um, [sh.node_rows(H, r, world, block=2).tolist() for r in range(world)]))
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
    assert bands == [[0, 1, 4, 5], [2, 3, 6]]


def test_partition_helpers():
    sh = importlib.import_module(PKG_NAME + ".sharding")
    for n, w in [(10, 3), (7, 8), (100, 8), (0, 2)]:
        spans = [sh.row_shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for W, H, w in [(512, 512, 8), (9, 7, 2), (5, 3, 3), (4, 9, 2)]:
        rows = [sh.node_rows(H, r, w) for r in range(w)]
        assert sorted(np.concatenate(rows).tolist()) == list(range(H)) and all(len(r) > 0 for r in rows)
        ids = [sh.node_ids(W, H, r, w) for r in range(w)]
        assert sorted(np.concatenate(ids).tolist()) == list(range(W * H)) and all(np.all(np.diff(i) > 0) for i in ids)
        for r in range(w):
            assert not set(sh.halo_rows(H, r, w).tolist()) & set(rows[r].tolist())
    assert sh.node_rows(512, 3, 8).tolist()[:6] == [12, 13, 14, 15, 44, 45]
    assert sh.sum_hits([np.array([1, 2], np.uint64), np.array([3, 0], np.uint64)]).tolist() == [4, 2]
    d = np.array([1.5, 1.5, np.nan, 0.0], np.float32)
    k = sh.pack_key(d, np.array([5, 2, 1, 9]))
    assert k.argmin() == 3 and sorted(k.tolist())[1] == k[1]
    assert sh.pack_key(np.array([np.nan], np.float32), np.array([0]))[0] == 0
    out = sh.merge_owner_outputs([np.array([1.0, np.nan], np.float32), np.array([np.nan, 2.0], np.float32)])
    assert out.tolist() == [1.0, 2.0]
