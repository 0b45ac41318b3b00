#!/usr/bin/env python3
"""Optical-flow row, measured: Matcher::searchByOPFlow's pyramidal LK (21x21, 3 levels) over
consecutive frames of a device-resident 1080p batch (keypoints from the extractor), plus cv2's own
calcOpticalFlowPyrLK and the oracle on one pair, with the parity of that pair. Prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import synth_frame

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=10)
a = ap.parse_args()
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = tb.Context(0, st.cuda_stream)
W, H, nl, sc, tgt = 1920, 1080, 8, 1 / 1.2, 2000
sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
B = a.batch
fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=tgt)
cap = fe.kp_cap
base = [synth_frame(k, W, H) for k in range(4)]
frames = np.stack([np.roll(base[(k // 16) % 4], (k % 16, 2 * (k % 16)), axis=(0, 1)) for k in range(B)])   # a slow pan
fe.upload_frames(0, frames)
fe.build_pyramid(B)
fe.extract(B, sf, tgt, 20.0, 7.0)
kps, desc, counts = fe.get_results_batch(B)
pairs = [(k, k - 1) for k in range(1, B) if k % 16]
f1 = [p[0] for p in pairs]; f2 = [p[1] for p in pairs]
pts = torch.zeros((len(pairs), cap, 2), dtype=torch.float32, device='cuda')
stt = torch.zeros((len(pairs), cap), dtype=torch.uint8, device='cuda')
lk = tb.LKParams()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.steps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


ms = timed(lambda: fe.search_by_opflow(lk, B, f1, f2, 0, pts.data_ptr(), stt.data_ptr()))
ms_eq = timed(lambda: fe.search_by_opflow(lk, B, f1, f2, 1, pts.data_ptr(), stt.data_ptr()))
fe.search_by_opflow(lk, B, f1, f2, 0, pts.data_ptr(), stt.data_ptr()); ctx.sync()
n_pts = int(sum(counts[q] for q in f2))
# per-kernel split of one call (CUDA events around every launch)
ctx.profile_enable(True)
fe.search_by_opflow(lk, B, f1, f2, 1, pts.data_ptr(), stt.data_ptr()); ctx.sync()
split = {k: round(v[1], 4) for k, v in ctx.profile_report().items()}
ctx.profile_enable(False)
fe.search_by_opflow(lk, B, f1, f2, 0, pts.data_ptr(), stt.data_ptr()); ctx.sync()
import oracle as O
O.build()
q1, q2 = pairs[0]
n2 = counts[q2]
kp = np.stack([kps[q2, :n2]['x'], kps[q2, :n2]['y']], 1)
t0 = time.time(); en, es, _ = O.lk_flow(O.LKParams(), frames[q2], frames[q1], kp); t_or = time.time() - t0
cv2.setNumThreads(1)
t0 = time.time()
for _ in range(5):
    cn, cs, _ = cv2.calcOpticalFlowPyrLK(frames[q2], frames[q1], kp.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
t_cv = (time.time() - t0) / 5
got = pts[0, :n2].cpu().numpy(); gs = stt[0, :n2].cpu().numpy()
xi, yi = en[:, 0].astype(np.int32), en[:, 1].astype(np.int32)
es = es & ((xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)).astype(np.uint8)
cs = cs.ravel(); ok = (gs == 1) & (cs == 1)
d = np.abs(got - cn.reshape(-1, 2)).max(1)[ok]
print(json.dumps({'workload': '%d frame pairs of 1080p, %d tracked points (avg %d/pair), LK 21x21, 3 levels, 30 its' % (len(pairs), n_pts, n_pts // len(pairs)),
                  'opflow_ms': round(ms, 4), 'opflow_equalized_ms': round(ms_eq, 4), 'points_per_s': round(n_pts / ms * 1e3, 1),
                  'pairs_per_s': round(len(pairs) / ms * 1e3, 1), 'tracked_frac': float(stt.float().sum().item() / n_pts),
                  'per_kernel_ms_equalized_call': split, 'cv2_ms_per_pair_1thread': round(t_cv * 1e3, 2), 'oracle_ms_per_pair_1core': round(t_or * 1e3, 1),
                  'parity_pair0_vs_oracle': bool(got.tobytes() == en.tobytes() and (gs == es).all()),
                  'vs_cv2_pair0': {'status_equal': float((gs == cs).mean()), 'within_0.05px': float((d <= 0.05).mean()), 'median_px': float(np.median(d))}}))
fe.close(); ctx.close()
