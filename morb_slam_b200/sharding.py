"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

* Extraction + stereo matching: frames are independent units -> frame i goes to rank i mod G, no
  collective on the data path.
* Map-descriptor kNN (BASELINE.json configs[4]): database rows are sharded contiguously,
  rank g owns rows [g*D/G, (g+1)*D/G); every rank scans its shard for the local top-2 of each query,
  the per-rank (index, distance) lists are exchanged with ONE all-gather (NCCL over NVLink on GPUs) and
  merged by (distance, global index) - identical to a single brute-force scan because the shards are
  index-contiguous and each local scan keeps the lowest index among ties.
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
