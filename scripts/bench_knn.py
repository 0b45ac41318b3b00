#!/usr/bin/env python3
"""BASELINE config 4: brute-force Hamming kNN (best, second, argbest + ratio test) on n x n 256-bit
descriptors, query rows sharded over the ranks, one all_gather of (idx, best, second).
  python scripts/bench_knn.py --n 200000 [--impl tc|popc]          (torchrun for several GPUs)
Prints one JSON line with Gpairs/s (whole job) — supplementary to bench.py."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.sharding import sharded_knn2
from trackingbench_slam_b200.synth import synth_descriptors

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=200000)
ap.add_argument('--impl', default='tc')
ap.add_argument('--steps', type=int, default=3)
a = ap.parse_args()
rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = tb.Context(local, st.cuda_stream)
ctx.set_match_impl(1 if a.impl == 'tc' else 0)
Q, T = synth_descriptors(a.n, seed=42)
q, t = torch.from_numpy(Q).cuda(), torch.from_numpy(T).cuda()
for _ in range(2):
    idx, best, second = sharded_knn2(ctx, q, t)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
for _ in range(a.steps):
    idx, best, second = sharded_knn2(ctx, q, t)
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = (time.perf_counter() - t0) / a.steps
if rank == 0:
    b, s, i = best.cpu().numpy(), second.cpu().numpy(), idx.cpu().numpy()
    ratio_ok = int((b.astype(np.float32) < 0.7 * s.astype(np.float32)).sum())
    d = np.unpackbits(Q[:2000] ^ T[i[:2000]], axis=1).sum(1)
    print(json.dumps({'metric': 'Hamming Gpairs/s (kNN-2 %dx%d, rows sharded, all_gather)' % (a.n, a.n), 'value': a.n * a.n / dt / 1e9,
                      'unit': 'Gpairs/s', 'n_gpus': world, 'impl': a.impl, 'ms': dt * 1e3, 'ratio_test_pass': ratio_ok,
                      'spot_check_exact': bool(np.array_equal(d, b[:2000])), 'planted_found': int((b[:a.n // 2] < 60).sum())}))
