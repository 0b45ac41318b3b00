"""N>1 path on CPU: world_size-2 gloo processes exercise the row/frame sharding + the final gather
(the one collective of the path). The per-shard kernel is stood in by the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    import oracle as O
    from trackingbench_slam_b200.sharding import shard_range, gather_rows, cross_check
    from trackingbench_slam_b200.synth import synth_descriptors
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    Q, T = synth_descriptors(n, seed=7)
    lo, hi = shard_range(n, rank, world)
    idx, best, second = O.knn2(Q[lo:hi], T)
    packed = torch.from_numpy(np.stack([idx, best, second], 1))
    full = gather_rows(packed, n).numpy()
    # cross-check variant: the transposed problem sharded the same way + a second gather
    lo2, hi2 = shard_range(len(T), rank, world)
    idx2, best2, second2 = O.knn2(T[lo2:hi2], Q)
    full2 = gather_rows(torch.from_numpy(np.stack([idx2, best2, second2], 1)), len(T)).numpy()
    mutual = cross_check(torch.from_numpy(full[:, 0].copy()), torch.from_numpy(full2[:, 0].copy())).numpy()
    if rank == 0:
        i0, b0, s0 = O.knn2(Q, T)
        ok = bool(np.array_equal(full[:, 0], i0) and np.array_equal(full[:, 1], b0) and np.array_equal(full[:, 2], s0))
        bf = O.bf_cross(Q, T)      # cv::BFMatcher(HAMMING, crossCheck=true) restatement
        ok = ok and np.array_equal(np.flatnonzero(mutual), bf['queryIdx']) and np.array_equal(full[mutual, 0], bf['trainIdx'])
        ok = ok and np.array_equal(full[mutual, 1].astype(np.float32), bf['distance'])
        ret['ok'] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [1001, 64])
def test_row_sharded_knn_gather_world2(n):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
    assert ret.get('ok') is True


def test_shard_ranges_cover_everything():
    from trackingbench_slam_b200.sharding import shard_range, frame_shards
    for n in (0, 1, 7, 256, 200000):
        for w in (1, 2, 3, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    assert frame_shards(256, 8)[3] == (96, 128)
