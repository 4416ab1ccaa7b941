"""The sharded top-2 search with the exchange fused over peer memory (include/orb_b200.h: orb_knn_exchange_*, orb_hamming_knn2_sharded)
against one brute-force scan of the whole database (oracle: cv::BFMatcher restatement, oracle_py.oracle_knn2). Bit-exact.
One GPU is enough: the ranks are separate handles (separate streams) in one process, connected with orb_knn_exchange_connect_local;
the multi-process / NVLink form (CUDA IPC) is exercised by tools/knn_p2p_check.py under torchrun and by bench.py --gpus N."""
import numpy as np
import pytest

from morb_slam_b200 import capi, synth
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu
DEV = capi.ORB_SRC_DEVICE | capi.ORB_DST_DEVICE


def _dev_copy(ex, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_single_rank_equals_plain_scan():
    import torch
    ex = capi.ORBextractor(1000, 1.2, 8, 20, 7)
    q = synth.random_descriptors(1, 700); db = synth.clustered_descriptors(2, q, 30000)
    x = capi.KnnExchange(ex, 0, 1, 1200)
    tq, tdb = _dev_copy(ex, q), _dev_copy(ex, db)
    out = torch.empty((2, 700, 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        x.search(tq.data_ptr(), 700, tdb.data_ptr(), len(db), 5, out[0].data_ptr(), out[1].data_ptr())
    io, do = op.oracle_knn2(q, db)
    assert np.array_equal(out[0].cpu().numpy(), io + 5) and np.array_equal(out[1].cpu().numpy(), do)
    x.close()


@pytest.mark.parametrize("world,rows", [(2, [20000, 20000]), (3, [15000, 0, 9001]), (4, [1, 300, 5000, 2])])
def test_local_ranks_equal_one_scan(world, rows):
    import torch
    exs = [capi.ORBextractor(1000, 1.2, 8, 20, 7) for _ in range(world)]
    xs = [capi.KnnExchange(exs[r], r, world, 1300) for r in range(world)]
    for x in xs:
        x.connect_local(xs)
    outs = [torch.empty((2, 1300, 2), dtype=torch.int32, device="cuda") for _ in range(world)]
    for epoch, nq in enumerate([1200, 37, 1300, 1200]):
        q = synth.random_descriptors(10 + epoch, nq)
        db = synth.clustered_descriptors(20 + epoch, q, sum(rows))      # near-duplicates: ties across the shard borders
        db[rows[0] - 1 if rows[0] else 0] = db[-1]                       # identical rows in different shards
        tq = _dev_copy(exs[0], q)
        bases = np.concatenate([[0], np.cumsum(rows)])
        shards = [_dev_copy(exs[0], db[bases[r]:bases[r + 1]]) if rows[r] else torch.empty((0, 32), dtype=torch.uint8, device="cuda") for r in range(world)]
        if epoch == 0:   # size the handles' scratch before ranks wait on each other (growing a buffer synchronises the device)
            for r in range(world):
                exs[r]._check(exs[r].L.orb_hamming_knn2(exs[r].h, _dev_copy(exs[r], synth.random_descriptors(0, 1300)).data_ptr(), 1300,
                                                        _dev_copy(exs[r], synth.random_descriptors(1, max(rows))).data_ptr(), max(rows), 0,
                                                        outs[r][0].data_ptr(), outs[r][1].data_ptr(), DEV))
        torch.cuda.synchronize()
        order = list(range(world)) if epoch % 2 == 0 else list(range(world))[::-1]
        for r in order:     # every rank enqueues; a rank's second kernel waits on the device for the others' flags
            xs[r].search(tq.data_ptr(), nq, shards[r].data_ptr() if rows[r] else 0, rows[r], int(bases[r]), outs[r][0].data_ptr(),
                         outs[r][1].data_ptr(), capi.ORB_ASYNC)
        for r in range(world):
            xs[r].check()
        io, do = op.oracle_knn2(q, db)
        for r in range(world):
            assert np.array_equal(outs[r][0, :nq].cpu().numpy(), io), (epoch, r)
            assert np.array_equal(outs[r][1, :nq].cpu().numpy(), do), (epoch, r)
    for x in xs:
        x.close()


def test_missing_peer_is_an_error_not_a_hang(monkeypatch):
    import torch
    monkeypatch.setenv("ORB_B200_KNN_TIMEOUT_S", "2")     # read by orb_knn_exchange_create (default 30 s)
    exs = [capi.ORBextractor(1000, 1.2, 8, 20, 7) for _ in range(2)]
    xs = [capi.KnnExchange(exs[r], r, 2, 64) for r in range(2)]
    with pytest.raises(capi.OrbError):       # not connected yet
        xs[0].search(1, 1, 1, 1, 0, 1, 1)
    for x in xs:
        x.connect_local(xs)
    q = _dev_copy(exs[0], synth.random_descriptors(3, 64)); db = _dev_copy(exs[0], synth.random_descriptors(4, 4096))
    out = torch.empty((2, 64, 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with pytest.raises(capi.OrbError) as e:  # rank 1 never calls: bounded wait, then ORB_ERR_STATE
        xs[0].search(q.data_ptr(), 64, db.data_ptr(), 4096, 0, out[0].data_ptr(), out[1].data_ptr())
    assert e.value.status == -6
    with pytest.raises(capi.OrbError):
        xs[0].search(q.data_ptr(), 65, db.data_ptr(), 4096, 0, out[0].data_ptr(), out[1].data_ptr())   # more queries than max_nq
    for x in xs:
        x.close()


def test_multi_process_exchange_over_nvlink():
    """the real thing: one process per GPU under torchrun, exchange buffers mapped with CUDA IPC, peer stores over NVLink.
    tools/knn_p2p_check.py asserts peer-memory route == NCCL all-gather route == brute force; skipped on a 1-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs on the box (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, KNN_ROWS="200000")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)), "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(root, "tools", "knn_p2p_check.py")], capture_output=True,
                       text=True, env=env, timeout=600)
    assert r.returncode == 0 and "EQUAL" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
