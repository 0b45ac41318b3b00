#!/usr/bin/env python3
"""Ingest row, measured: PinholeCamera::UndistortImage (cv::remap INTER_LINEAR) and Frame::Equalize (CLAHE 3.0, 8x8) on a
device-resident batch of 1080p frames, against the HBM roofline (algorithmic bytes = 1 B read + 1 B written per pixel),
plus the CPU oracle and cv2 on one frame. Prints one JSON line (supplementary to bench.py)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import synth_frame

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--steps', type=int, default=10)
a = ap.parse_args()
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = tb.Context(0, st.cuda_stream)
W, H, B = 1920, 1080, a.batch
cam = tb.Pinhole(1000.0, 1000.0, 960.0, 540.0, W, H, (tb.capi.C.c_float * 5)(-0.28, 0.07, 0.0002, 0.00002, 0.0))
m1, m2 = tb.init_undistort_rectify_map(cam)
und = tb.Undistorter(ctx, m1, m2)
base = np.stack([synth_frame(k, W, H) for k in range(8)])
src = torch.from_numpy(base).cuda().repeat((B + 7) // 8, 1, 1)[:B].contiguous()
dst = torch.zeros_like(src)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
except Exception:
    pass


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


ms_remap = timed(lambda: und.apply_device(src.data_ptr(), W, H, W, W * H, dst.data_ptr(), W, W * H, B))
remap_out = dst[0].cpu().numpy()
ms_clahe = timed(lambda: ctx.clahe_device(src.data_ptr(), W, H, W, W * H, dst.data_ptr(), W, W * H, B, 3.0, (8, 8)))
clahe_out = dst[0].cpu().numpy()
ctx.profile_enable(True)
ctx.clahe_device(src.data_ptr(), W, H, W, W * H, dst.data_ptr(), W, W * H, B, 3.0, (8, 8))
ctx.sync()
prof = ctx.profile_report()
ctx.profile_enable(False)
import oracle as O
t0 = time.time(); o_remap = O.remap_linear(base[0], m1, m2); cpu_remap = time.time() - t0
t0 = time.time(); o_clahe = O.clahe(base[0]); cpu_clahe = time.time() - t0
out = {'workload': '%d x 1920x1080 u8 frames, device-resident' % B,
       'remap_ms_per_batch': ms_remap, 'remap_frames_per_s': B / (ms_remap / 1e3), 'remap_GBps_algorithmic': 2 * W * H * B / ms_remap / 1e6,
       'clahe_ms_per_batch': ms_clahe, 'clahe_frames_per_s': B / (ms_clahe / 1e3), 'clahe_GBps_algorithmic': 3 * W * H * B / ms_clahe / 1e6,
       'clahe_kernels_ms': prof, 'cpu_oracle_remap_ms_per_frame': cpu_remap * 1e3, 'cpu_oracle_clahe_ms_per_frame': cpu_clahe * 1e3,
       'parity_frame0': bool(np.array_equal(remap_out, o_remap) and np.array_equal(clahe_out, o_clahe)),
       'hbm_peak_GBps': peaks.get('hbm_gbs')}
try:
    import cv2
    cv2.setNumThreads(1)
    t0 = time.time(); cv2.remap(base[0], m1, m2, cv2.INTER_LINEAR); out['cv2_remap_ms_per_frame_1thr'] = (time.time() - t0) * 1e3
    cl = cv2.createCLAHE(3.0, (8, 8)); t0 = time.time(); cl.apply(base[0]); out['cv2_clahe_ms_per_frame_1thr'] = (time.time() - t0) * 1e3
except Exception:
    pass
print(json.dumps(out))
