"""compute-sanitizer companion of sanitize_smoke.py for the large-batch code path: 33 small frames take direct launches, the
side-stream pyramid, per-level octree and describe launches and the tcgen05 matcher."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import synth_stream
W, H, B, nl, sc, tgt = 300, 170, 33, 4, 1 / 1.2, 300
ctx = tb.Context(0)
ctx.set_match_impl(1)
sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=tgt)
frames = synth_stream(B, W, H, seed0=3)
cap = fe.kp_cap
m = torch.zeros((B - 1, cap, 4), dtype=torch.int32, device='cuda'); c = torch.zeros(B - 1, dtype=torch.int32, device='cuda')
for th in ((20.0, 7.0), (40.0, 40.0)):
    fe.upload_frames(0, frames)
    fe.build_pyramid(B)
    fe.extract(B, sf, tgt, *th)
    fe.search_by_bf(np.arange(0, B - 1, dtype=np.int32), np.arange(1, B, dtype=np.int32), 10.0, 30.0, m.data_ptr(), cap, c.data_ptr())
    ctx.sync()
kps, desc, counts = fe.get_results_batch(B)
print('batch sanitizer run ok:', int(counts.sum()), 'keypoints')
fe.close(); ctx.close()
