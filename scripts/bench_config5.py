#!/usr/bin/env python3
"""BASELINE config 5: synthetic 4K 8-stream test_vo-style tracking — per frame: pyramid + ORB extract(2000) +
projection-guided match against the previous frame of the same stream (map points = previous keypoints at a
synthetic depth of 10 m, 1 px/frame pan).  Streams are sharded over the ranks (torchrun), no collective.
Device-resident; prints one JSON line (supplementary to bench.py)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import synth_frame

ap = argparse.ArgumentParser()
ap.add_argument('--streams', type=int, default=8)
ap.add_argument('--frames', type=int, default=8, help='frames per stream in one batch')
ap.add_argument('--steps', type=int, default=5)
a = ap.parse_args()
rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
W, H, nl, sc, tgt = 3840, 2160, 8, 1 / 1.2, 2000
my_streams = [s for s in range(a.streams) if s % world == rank]
B = len(my_streams) * a.frames
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = tb.Context(local, st.cuda_stream)
sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=tgt)
cap = fe.kp_cap
frames = np.stack([synth_frame(k, W, H, seed0=2000 + 1000 * s) for s in my_streams for k in range(a.frames)])
fe.upload_frames(0, frames)
pairs = [(i * a.frames + k, i * a.frames + k - 1) for i in range(len(my_streams)) for k in range(1, a.frames)]
f1 = np.array([p[0] for p in pairs], np.int32); f2 = np.array([p[1] for p in pairs], np.int32)
fx = fy = 2000.0; cx, cy = W / 2.0, H / 2.0
T = np.tile(np.eye(4, dtype=np.float32), (len(pairs), 1, 1)); T[:, 0, 3] = -10.0 / fx
cam = tb.Pinhole(fx, fy, cx, cy, W, H, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
p = tb.MatcherParams(30, 50, 30, 1, 30.0)
kps_d, desc_d, cnt_d = fe.results_device()
m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda'); c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
valid = torch.ones((B, cap), dtype=torch.uint8, device='cuda')
xyz = torch.zeros((B, cap, 3), dtype=torch.float32, device='cuda')

def step():
    fe.build_pyramid(B)
    fe.extract(B, sf, tgt, 20.0, 7.0)
    # map points of every frame = its keypoints back-projected at 10 m (plumbing on the device with torch)
    kt = _kp_tensor
    xyz[:, :, 0] = (kt[:, :, 0] - cx) / fx * 10.0
    xyz[:, :, 1] = (kt[:, :, 1] - cy) / fy * 10.0
    xyz[:, :, 2] = 10.0
    fe.search_by_projection(p, cam, f1, f2, T, None, valid.data_ptr(), xyz.data_ptr(), sf, m.data_ptr(), cap, c.data_ptr())

# a zero-copy tensor view of the keypoint buffer (B x cap x 7 float32) through the CUDA array interface
class _Arr:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {'shape': shape, 'typestr': '<f4', 'data': (ptr, False), 'version': 2}
_kp_tensor = torch.as_tensor(_Arr(kps_d, (B, cap, 7)), device='cuda')

for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(a.steps):
    step()
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    n_match = int(c.sum().item())
    print(json.dumps({'metric': 'frames/s 4K extract + projection-guided match (%d streams x %d frames, device-resident)' % (a.streams, a.frames),
                      'value': a.streams * a.frames / (ms / 1e3), 'unit': 'frames/s', 'n_gpus': world, 'ms_per_step': ms,
                      'projection_matches_rank0': n_match, 'pairs_rank0': len(pairs)}))
