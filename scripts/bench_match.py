#!/usr/bin/env python3
"""Matcher kernels alone (for ncu and quick A/B): (a) the batched frame-pair searchByBF of bench.py's step — 255 pairs of ~2006
descriptors, both directions — through tb_frontend_search_by_bf on a frontend whose descriptors are filled with random bytes;
(b) kNN-2 n x n.  Prints ms and Tpairs/s per kernel (CUDA events on the ctx stream)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import synth_descriptors

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=200000)
ap.add_argument('--pairs', type=int, default=255)
ap.add_argument('--rows', type=int, default=2006)
ap.add_argument('--impl', default='tc')
ap.add_argument('--iters', type=int, default=10)
a = ap.parse_args()
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = tb.Context(0, st.cuda_stream)
ctx.set_match_impl(1 if a.impl == 'tc' else 0)
out = {}

def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.iters): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters

# (a) P independent pair problems through the single-pair device entry is launch bound; use one big "frame set" instead:
#     the kNN-2 entry with nq = nt = rows per problem is not batched, so build a frontend-free batch with knn2 on stacked sets
rng = np.random.default_rng(1)
B = a.pairs + 1
d = torch.from_numpy(rng.integers(0, 256, (B, a.rows, 32), dtype=np.uint8)).cuda()
lw, lh = [64], [64]
fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=a.rows)
cap = fe.kp_cap
kd, dd, cd = fe.results_device()
# fill the frontend's descriptor / count buffers directly (device plumbing with torch)
class _Arr:
    def __init__(self, ptr, shape, t): self.__cuda_array_interface__ = {'shape': shape, 'typestr': t, 'data': (ptr, False), 'version': 2}
D = torch.as_tensor(_Arr(dd, (B, cap, 32), '|u1'), device='cuda'); Cn = torch.as_tensor(_Arr(cd, (B,), '<i4'), device='cuda')
D[:, :a.rows] = d; Cn[:] = a.rows
m = torch.zeros((a.pairs, cap, 4), dtype=torch.int32, device='cuda'); c = torch.zeros(a.pairs, dtype=torch.int32, device='cuda')
qf, tf = np.arange(0, a.pairs, dtype=np.int32), np.arange(1, B, dtype=np.int32)
ms = timed(lambda: fe.search_by_bf(qf, tf, 10.0, 30.0, m.data_ptr(), cap, c.data_ptr()))
pe = 2.0 * a.pairs * a.rows * a.rows
out['batched'] = {'ms': ms, 'Tpairs_per_s': pe / (ms * 1e-3) / 1e12, 'pairs': a.pairs, 'rows': a.rows}
# (b) kNN-2
Q, T = synth_descriptors(a.n, seed=42)
q, t = torch.from_numpy(Q).cuda(), torch.from_numpy(T).cuda()
idx = torch.empty(a.n, dtype=torch.int32, device='cuda'); best = torch.empty_like(idx); second = torch.empty_like(idx)
ms = timed(lambda: ctx.knn2_device(q.data_ptr(), a.n, t.data_ptr(), a.n, idx.data_ptr(), best.data_ptr(), second.data_ptr()))
out['knn2'] = {'ms': ms, 'Tpairs_per_s': float(a.n) * a.n / (ms * 1e-3) / 1e12, 'n': a.n}
print(json.dumps(out))
