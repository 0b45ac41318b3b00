#!/usr/bin/env python3
"""Stereo row, measured: the ORB row-band matcher (LocalBA.cpp:69-288) over the frame pairs
(2k, 2k+1) of a device-resident 1080p batch, plus the CPU oracle on one pair and the parity of that pair.
Prints one JSON line (supplementary to bench.py)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import synth_frame

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--steps', type=int, default=20)
a = ap.parse_args()
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = tb.Context(0, st.cuda_stream)
W, H, nl, sc, tgt = 1920, 1080, 8, 1 / 1.2, 2000
sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
inv_sf = (np.float32(1.0) / sf).astype(np.float32)
B = a.batch
fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=tgt)
cap = fe.kp_cap
base = [synth_frame(k, W, H) for k in range(8)]
# right view = left view shifted by a few pixels (a fronto-parallel scene), so that row bands hold true partners
frames = np.stack([base[(k // 2) % 8] if k % 2 == 0 else np.roll(base[(k // 2) % 8], -(3 + (k // 2) % 5), axis=1) for k in range(B)])
fe.upload_frames(0, frames)
fe.build_pyramid(B)
fe.extract(B, sf, tgt, 20.0, 7.0)
kps, desc, counts = fe.get_results_batch(B)
pairs = [(2 * k, 2 * k + 1) for k in range(B // 2)]
fl = [p[0] for p in pairs]; fr = [p[1] for p in pairs]
m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda'); c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
z = torch.zeros((len(pairs), cap), dtype=torch.float32, device='cuda')
p = tb.StereoParams(bf=386.1448, th_dist=75)


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


ms = timed(lambda: fe.stereo_match(p, fl, fr, inv_sf, m.data_ptr(), cap, c.data_ptr(), z.data_ptr()))
cn = c.cpu().numpy()
import oracle as O
O.build()
po = O.StereoParams(bf=386.1448, th_dist=75)
nL, nR = counts[0], counts[1]
t0 = time.time()
reps = 20
for _ in range(reps):
    em, ez = O.stereo_rowband(po, kps[0, :nL], desc[0, :nL], kps[1, :nR], desc[1, :nR], inv_sf, H)
cpu_ms = (time.time() - t0) / reps * 1e3
got = m[0, :cn[0]].cpu().numpy().view(np.uint8).tobytes()
print(json.dumps({'workload': '%d stereo pairs of 1080p frames, %d keypoints/frame avg' % (len(pairs), int(counts.mean())),
                  'stereo_rowband_ms': round(ms, 4), 'pairs_per_s': round(len(pairs) / ms * 1e3, 1),
                  'matches_per_pair_avg': float(cn.mean()), 'cpu_oracle_ms_per_pair_1core': round(cpu_ms, 3),
                  'parity_pair0': bool(got == em.tobytes() and z[0, :nL].cpu().numpy().tobytes() == ez.tobytes())}))
fe.close(); ctx.close()
