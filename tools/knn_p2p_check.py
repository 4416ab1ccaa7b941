"""Multi-process check of the peer-memory kNN exchange (run under torchrun on >= 2 GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/knn_p2p_check.py
Every rank scans its shard; ShardedKnn (CUDA IPC buffers, NVLink peer stores, flag wait) must equal sharded_knn2 (NCCL all-gather +
merge kernel) and, on rank 0, one brute-force scan of the concatenated database. Prints the time of both routes."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morb_slam_b200 import capi, sharding, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = "cuda:%d" % local
    ex = capi.ORBextractor(1000, 1.2, 8, 20, 7, device=local)
    nq, rows = 1200, int(os.environ.get("KNN_ROWS", "1250000"))
    knn = sharding.ShardedKnn(ex, nq)
    q = torch.from_numpy(synth.random_descriptors(10, nq)).to(dev)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    db = torch.randint(0, 256, (rows, 32), dtype=torch.uint8, device=dev, generator=g)
    db[rank::7][:, :] = q[rank % nq]          # exact duplicates of one query in every shard: distance-0 ties across ranks
    torch.cuda.synchronize()
    base = rank * rows
    ok = True
    for rep in range(3):
        i1, d1 = knn.search(q, db, base)
        i0, d0 = sharding.sharded_knn2(ex, q, db, base)
        torch.cuda.synchronize()
        ok &= bool(torch.equal(i0, i1) and torch.equal(d0, d1))
    # every rank holds the same result
    chk = torch.tensor([int(i1.sum().item()), int(d1.sum().item())], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    ok &= all(torch.equal(c, allc[0]) for c in allc)
    # small brute force on rank 0: gather 20 k rows per rank
    sub = 20000
    parts = [torch.empty((sub, 32), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, db[:sub].contiguous())
    i2, d2 = knn.search(q, db[:sub].contiguous(), rank * sub)
    if rank == 0:
        full = torch.cat(parts).cpu().numpy()
        ib, dbst = capi.hamming_knn2(ex, q.cpu().numpy(), full)
        ok &= bool(np.array_equal(ib, i2.cpu().numpy()) and np.array_equal(dbst, d2.cpu().numpy()))

    def timed(fn, reps=10):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ex.sync(); torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps

    out = torch.empty((2, nq, 2), dtype=torch.int32, device=dev)
    timed(lambda: knn.search(q, db, base, out=out, flags=capi.ORB_ASYNC), reps=100)     # clocks up (a fresh box idles at low clocks)
    t_nccl = timed(lambda: sharding.sharded_knn2(ex, q, db, base))
    t_p2p = timed(lambda: knn.search(q, db, base, out=out, flags=capi.ORB_ASYNC))
    t = torch.tensor([t_p2p, t_nccl, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("knn_p2p_check world=%d rows/gpu=%d: %s  peer-memory exchange %.3f ms, NCCL all-gather route %.3f ms per search"
              % (world, rows, "EQUAL" if t[2].item() == 0 else "MISMATCH", t[0].item(), t[1].item()))
    knn.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
