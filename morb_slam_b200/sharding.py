"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

* Extraction + stereo matching: frames are independent units -> frame i goes to rank i mod G, no
  collective on the data path.
* Map-descriptor kNN (BASELINE.json configs[4]): database rows are sharded contiguously,
  rank g owns rows [g*D/G, (g+1)*D/G); every rank scans its shard for the local top-2 of each query,
  the per-rank (index, distance) lists are exchanged and merged by (distance, global index) - identical to a
  single brute-force scan because the shards are index-contiguous and each local scan keeps the lowest
  index among ties. Two exchanges: `ShardedKnn` fuses it into the kernels over peer memory (the merge kernel
  stores every rank's top-2 straight into all ranks' buffers over NVLink and a second kernel waits for the
  epoch flags - no collective call, no host synchronisation in between); `sharded_knn2` is the plain
  formulation with ONE all-gather (NCCL on GPUs, gloo in the CPU tests) and a merge kernel.
"""
import numpy as np


def frames_of_rank(n_frames, rank, world):
    """Indices of the frames rank `rank` processes (round robin; left/right of a pair stay together)."""
    return list(range(rank, n_frames, world))


def db_rows_of_rank(ndb, rank, world):
    """[begin, end) of the contiguous database shard of `rank`."""
    return (ndb * rank) // world, (ndb * (rank + 1)) // world


def sharded_knn2(ex, q, db_local, index_base, group=None):
    """Top-2 Hamming neighbours of q (torch uint8 [nq, 32]) in a row-sharded database.

    db_local: this rank's shard (torch uint8 [rows, 32]) on the same device as q; index_base: global
    index of its first row. Returns (idx, dist) torch int32 [nq, 2], identical on every rank.
    CUDA tensors go through the C ABI with device pointers (no host staging); the exchange is a single
    torch.distributed all_gather_into_tensor on `group` (NCCL for CUDA tensors).
    """
    import torch
    import torch.distributed as dist
    from . import capi
    nq = q.shape[0]
    dev = q.device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = torch.empty((2, nq, 2), dtype=torch.int32, device=dev)     # [idx | dist]
    if dev.type == "cuda":
        fl = capi.ORB_SRC_DEVICE | capi.ORB_DST_DEVICE
        capi.hamming_knn2(ex, q.data_ptr(), db_local.data_ptr(), index_base, fl, ndb=db_local.shape[0], nq=nq,
                          out=(local[0].data_ptr(), local[1].data_ptr()))
    else:
        i, d = capi.hamming_knn2(ex, q.numpy(), db_local.numpy(), index_base)
        local[0] = torch.from_numpy(i); local[1] = torch.from_numpy(d)
    if world == 1:
        return local[0], local[1]
    gathered = torch.empty((world * 2, nq, 2), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gathered, local, group=group)          # 16 * nq bytes per rank
    gathered = gathered.view(world, 2, nq, 2)
    idx_parts = gathered[:, 0].contiguous()
    dist_parts = gathered[:, 1].contiguous()
    out = torch.empty((2, nq, 2), dtype=torch.int32, device=dev)
    if dev.type == "cuda":
        capi.knn2_merge(ex, idx_parts.data_ptr(), dist_parts.data_ptr(), capi.ORB_SRC_DEVICE | capi.ORB_DST_DEVICE,
                        nparts=world, nq=nq, out=(out[0].data_ptr(), out[1].data_ptr()))
    else:
        i, d = capi.knn2_merge(ex, idx_parts.numpy(), dist_parts.numpy())
        out[0] = torch.from_numpy(np.asarray(i)); out[1] = torch.from_numpy(np.asarray(d))
    return out[0], out[1]


class ShardedKnn:
    """The sharded top-2 search with the peer-memory exchange (include/orb_b200.h: orb_hamming_knn2_sharded).

    One instance per rank (one process per GPU). Set-up exchanges the 64-byte CUDA IPC handles of the ranks' exchange buffers
    with one all_gather on `group`; after that search() only enqueues kernels on the extractor handle's stream."""

    def __init__(self, ex, max_nq, group=None):
        import torch
        import torch.distributed as dist
        from . import capi
        self.ex, self.max_nq = ex, max_nq
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.x = capi.KnnExchange(ex, self.rank, self.world, max_nq)
        if self.world > 1:
            dev = torch.device("cuda", ex.device)
            mine = torch.from_numpy(self.x.handle.copy()).to(dev)
            allh = torch.empty((self.world, capi.ORB_IPC_HANDLE_BYTES), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            self.x.connect(allh.cpu().numpy())
            dist.barrier(group)          # every rank has mapped every buffer before the first search writes into them

    def search(self, q, db_local, index_base, out=None, flags=0):
        """q: torch uint8 [nq, 32], db_local: torch uint8 [rows, 32], both on the handle's device. Returns (idx, dist) int32 [nq, 2],
        identical on every rank. With capi.ORB_ASYNC the call only enqueues (results valid after ex.sync())."""
        import torch
        nq = q.shape[0]
        if out is None:
            out = torch.empty((2, nq, 2), dtype=torch.int32, device=q.device)
        self.x.search(q.data_ptr(), nq, db_local.data_ptr(), db_local.shape[0], index_base, out[0].data_ptr(), out[1].data_ptr(), flags)
        return out[0], out[1]

    def check(self):
        """Completes the searches enqueued with ORB_ASYNC; raises if a peer stayed away."""
        self.x.check()

    def close(self):
        self.x.close()
