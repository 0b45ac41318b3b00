"""Small end-to-end run for compute-sanitizer: extract (both passes, AddPoints), all matchers, FASTExtractor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import load_gray, synth_descriptors
left, right = load_gray('left.png')[:200, :420], load_gray('right.png')[:200, :420]
ctx = tb.Context(0)
for nl, sc, tgt, a, b in ((4, 0.8, 300, 40.0, 12.0), (3, 1 / 1.2, 200, 20.0, 7.0)):
    sf, lw, lh = tb.pyramid_geometry(left.shape[1], left.shape[0], nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=3, max_target=tgt)
    fe.upload_frames(0, np.stack([left, right, left]))
    fe.build_pyramid(3)
    fe.extract(3, sf, tgt, a, b)
    kps, desc, counts = fe.get_results_batch(3)
    fe.extract(3, sf, tgt, a, b, exit_kps=kps[0, :counts[0]:4])
    fe.extract(3, sf, tgt, a, b)
    kps, desc, counts = fe.get_results_batch(3)
    cap = fe.kp_cap
    for impl in (0, 1):
        ctx.set_match_impl(impl)
        m = torch.zeros((2, cap, 4), dtype=torch.int32, device='cuda'); c = torch.zeros(2, dtype=torch.int32, device='cuda')
        fe.search_by_bf([0, 1], [1, 2], 10.0, 60.0, m.data_ptr(), cap, c.data_ptr()); ctx.sync()
    k0, d0, k1, d1 = kps[0, :counts[0]], desc[0, :counts[0]], kps[1, :counts[1]], desc[1, :counts[1]]
    H, W = left.shape
    ctx.search_by_violence(tb.MatcherParams(50, 100, 30, 1, 0.9), k0, d0, k1, d1, W, H, 0, nl, 40.0)
    cam = tb.Pinhole(300.0, 300.0, W / 2, H / 2, W, H, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
    xyz = np.stack([(k0['x'] - W / 2) / 300 * 5, (k0['y'] - H / 2) / 300 * 5, np.full(len(k0), 5.0)], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32); T[0, 3] = -0.05
    ctx.search_by_projection(tb.MatcherParams(30, 80, 30, 1, 20.0), cam, T, k1, d1, None, sf, W, H, k0, np.ones(len(k0), np.uint8), xyz, d0)
    nrm = xyz / np.linalg.norm(xyz, axis=1, keepdims=True)
    ctx.search_by_projection_map(tb.MatcherParams(30, 80, 30, 1, 5.0), cam, T, k1, d1, None, sf, W, H, np.zeros(len(k0), np.uint8), xyz, nrm.astype(np.float32), d0, 0.9)
    ctx.features_in_area(k0, W, H, 100.0, 80.0, 30.0, 0, 2)
    fe.close()
Q, Tt = synth_descriptors(700, seed=3)
q, t = torch.from_numpy(Q).cuda(), torch.from_numpy(Tt).cuda()
for impl in (0, 1):
    ctx.set_match_impl(impl)
    i = torch.empty(700, dtype=torch.int32, device='cuda'); bb = torch.empty_like(i); s = torch.empty_like(i)
    ctx.knn2_device(q.data_ptr(), 700, t.data_ptr(), 700, i.data_ptr(), bb.data_ptr(), s.data_ptr())
    ctx.match_bf_cross_device(q.data_ptr(), 700, t.data_ptr(), 650)
img = np.ascontiguousarray(np.tile(load_gray('left.png')[:360, :640], (2, 1))[:480, :640])
ctx.fast_extract_host([img], np.ones(1, np.float32), 300, 20.0)
torch.cuda.synchronize()
print('sanitize smoke done')
