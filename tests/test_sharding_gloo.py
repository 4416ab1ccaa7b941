"""N>1 host logic on CPU: two gloo ranks run the sharded-kNN plumbing of morb_slam_b200/sharding.py
(shard bounds, index_base, one all-gather, merge by (distance, index)) with the local scan / merge
kernels replaced by the oracle, and must reproduce the single brute-force scan. Also frame sharding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from morb_slam_b200 import sharding, synth
from oracle import oracle_py as op


def _oracle_merge(idx_parts, dist_parts):
    P, nq, _ = idx_parts.shape
    oi = np.full((nq, 2), -1, np.int32); od = np.full((nq, 2), -1, np.int32)
    for q in range(nq):
        c = [(int(dist_parts[p, q, k]), int(idx_parts[p, q, k])) for p in range(P) for k in range(2) if idx_parts[p, q, k] >= 0]
        c.sort()
        for k, (d, i) in enumerate(c[:2]):
            oi[q, k], od[q, k] = i, d
    return oi, od


def _worker(rank, world, port, ndb, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from morb_slam_b200 import capi
    # the CUDA entry points are unavailable on CPU: the checker stands in for the two kernels only
    capi.hamming_knn2 = lambda ex, q, db, index_base=0, flags=0, **kw: tuple(
        (lambda i, d: (np.where(i >= 0, i + index_base, -1).astype(np.int32), d))(*op.oracle_knn2(q, db)))
    capi.knn2_merge = lambda ex, ip, dp, flags=0, **kw: _oracle_merge(np.asarray(ip), np.asarray(dp))
    q = synth.random_descriptors(3, 64)
    db = synth.clustered_descriptors(4, q, ndb, max_flips=6)   # many exact ties across shards
    b, e = sharding.db_rows_of_rank(ndb, rank, world)
    idx, dd = sharding.sharded_knn2(None, torch.from_numpy(q), torch.from_numpy(db[b:e].copy()), b)
    if rank == 0:
        ret["idx"] = idx.numpy().copy(); ret["dist"] = dd.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ndb", [4001, 2])
def test_sharded_knn_two_gloo_ranks(ndb):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500) + ndb % 7
    mp.spawn(_worker, args=(world, port, ndb, ret), nprocs=world, join=True)
    q = synth.random_descriptors(3, 64)
    db = synth.clustered_descriptors(4, q, ndb, max_flips=6)
    io, do = op.oracle_knn2(q, db)
    assert np.array_equal(ret["idx"], io) and np.array_equal(ret["dist"], do)


def test_shard_bounds_and_frame_round_robin():
    for ndb in (0, 1, 7, 10_000_000):
        for world in (1, 2, 4, 8):
            rows = [sharding.db_rows_of_rank(ndb, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == ndb
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
    assert sharding.db_rows_of_rank(10_000_000, 3, 8) == (3_750_000, 5_000_000)
    fr = [sharding.frames_of_rank(10, r, 4) for r in range(4)]
    assert sorted(sum(fr, [])) == list(range(10)) and fr[1] == [1, 5, 9]
