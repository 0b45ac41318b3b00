#!/usr/bin/env python3
"""Matcher::rejectWithF alone: P frame pairs of ~2000 tracks each (synthetic two-view geometry, 25 % gross outliers) through
tb_reject_with_f_device (one CTA per pair) against cv2.findFundamentalMat on one host thread.  Prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import trackingbench_slam_b200 as tb

ap = argparse.ArgumentParser()
ap.add_argument('--pairs', type=int, default=256)
ap.add_argument('--n', type=int, default=2000)
ap.add_argument('--iters', type=int, default=5)
a = ap.parse_args()
st_ = torch.cuda.Stream(); torch.cuda.set_stream(st_)
ctx = tb.Context(0, st_.cuda_stream)
P, n = a.pairs, a.n
cur = np.zeros((P, n, 2), np.float32); last = np.zeros((P, n, 2), np.float32)
K = np.array([[700, 0, 620], [0, 700, 188], [0, 0, 1.]])
for p in range(P):
    rs = np.random.default_rng(p)
    X = np.c_[rs.uniform(-5, 5, n), rs.uniform(-3, 3, n), rs.uniform(4, 30, n)]
    Rm, _ = cv2.Rodrigues(rs.uniform(-0.05, 0.05, 3)); tv = rs.uniform(-0.5, 0.5, 3)
    p1 = (K @ X.T).T; p1 = p1[:, :2] / p1[:, 2:]
    p2 = (K @ ((Rm @ X.T).T + tv).T).T; p2 = p2[:, :2] / p2[:, 2:]
    p1 += rs.normal(0, 0.3, p1.shape); p2 += rs.normal(0, 0.3, p2.shape)
    o = rs.choice(n, n // 4, replace=False); p2[o] += rs.uniform(-40, 40, (len(o), 2))
    cur[p] = p2; last[p] = p1
d_cur, d_last = torch.from_numpy(cur).cuda(), torch.from_numpy(last).cuda()
cnt = torch.full((P,), n, dtype=torch.int32, device='cuda')
it = torch.zeros(P, dtype=torch.int32, device='cuda')
def run():
    s = torch.ones((P, n), dtype=torch.uint8, device='cuda')
    ctx.reject_with_f_device(P, d_cur.data_ptr(), d_last.data_ptr(), cnt.data_ptr(), n, s.data_ptr(), None, it.data_ptr())
    return s
for _ in range(2): s = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st_)
for _ in range(a.iters): s = run()
e1.record(st_); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
S = s.cpu().numpy()
t0 = time.perf_counter(); bad = 0; m = min(P, 16)
cv2.setNumThreads(1)
for p in range(m):
    F, mask = cv2.findFundamentalMat(cur[p], last[p], cv2.FM_RANSAC, 1.0, 0.99)
    bad += int(not np.array_equal(mask.ravel(), S[p]))
cpu_ms = (time.perf_counter() - t0) * 1e3 / m
print(json.dumps({'pairs': P, 'tracks_per_pair': n, 'ms_per_batch': ms, 'pairs_per_s': P / ms * 1e3, 'mean_ransac_iterations': float(it.float().mean()),
                  'cv2_ms_per_pair_1_thread': cpu_ms, 'masks_differing_from_cv2': bad, 'pairs_checked': m}))
