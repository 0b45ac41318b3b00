"""Config 4 on more than one GPU (runs only where >= 2 CUDA devices are visible): two NCCL ranks shard the query rows, gather
with one all_gather and must reproduce the single-GPU result and the oracle, including the cross-check variant."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    import torch.distributed as dist
    import oracle as O
    import trackingbench_slam_b200 as tb
    from trackingbench_slam_b200.sharding import sharded_knn2_cross
    from trackingbench_slam_b200.synth import synth_descriptors
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    ctx = tb.Context(rank, st.cuda_stream)
    ctx.set_match_impl(1)
    Q, T = synth_descriptors(n, seed=5)
    q, t = torch.from_numpy(Q).cuda(), torch.from_numpy(T).cuda()
    idx, best, second, mutual = sharded_knn2_cross(ctx, q, t)
    torch.cuda.synchronize()
    if rank == 0:
        i0, b0, s0 = O.knn2(Q, T)
        bf = O.bf_cross(Q, T)
        mu = mutual.cpu().numpy()
        ret['ok'] = bool(np.array_equal(idx.cpu().numpy(), i0) and np.array_equal(best.cpu().numpy(), b0)
                         and np.array_equal(second.cpu().numpy(), s0) and np.array_equal(np.flatnonzero(mu), bf['queryIdx']))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('n', [4099, 4096])   # unequal shards (padded list gather) and equal shards (one all_gather_into_tensor)
def test_row_sharded_knn_two_gpus(n):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, ret), nprocs=2, join=True)
    assert ret.get('ok') is True
