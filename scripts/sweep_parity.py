#!/usr/bin/env python3
"""Extended randomized parity sweep (not part of pytest): batches of 33-40 frames (per-level octree launches, direct
launch path) and single frames (graph replay path) of random sizes, scale factors, budgets and thresholds, every frame
compared bit for bit with the oracle; frames (0, 1) of each batch also go through the stereo row-band matcher and the LK
tracker (random parameters, some points pushed past the border). Usage: python scripts/sweep_parity.py [cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import trackingbench_slam_b200 as tb
import oracle as O
from trackingbench_slam_b200.synth import synth_frame

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
left = cv2.imread(os.path.join(ROOT, 'trackingbench_slam_b200/data/left.png'), 0)
ctx = tb.Context(0)
bad = 0
t0 = time.time()
for c in range(cases):
    w, h = int(rng.integers(200, 1300)), int(rng.integers(120, 700))
    B = int(rng.choice([1, 2, 17, 33, 40]))
    nl = int(rng.integers(2, 9))
    sc = float(rng.choice([0.8, 1 / 1.2, 0.9, 0.7, 0.6]))
    tgt = int(rng.integers(100, 2000))
    a, b = float(rng.choice([20, 40, 80])), float(rng.choice([5, 7, 20]))
    frames = np.stack([synth_frame(int(rng.integers(0, 1000)), w, h) for _ in range(B)])
    if rng.random() < 0.3:
        frames[0, :min(h, left.shape[0]), :min(w, left.shape[1])] = left[:h, :w]
    sf, lw, lh = tb.pyramid_geometry(w, h, nl, sc)
    if lw[-1] < 40 or lh[-1] < 40:
        continue
    fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=tgt)
    for rep in range(2):                       # second pass replays graphs / reuses workspaces
        fe.upload_frames(0, frames)
        fe.build_pyramid(B)
        fe.extract(B, sf, tgt, a, b)
    kps, desc, counts = fe.get_results_batch(B)
    check = range(B) if B <= 4 else sorted(set(int(x) for x in rng.integers(0, B, 4)) | {0, B - 1})
    for f in check:
        lv, sfo = O.pyramid(frames[f], nl, sc)
        for l in range(nl):
            if not np.array_equal(fe.download_level(f, l), lv[l]):
                print('PYRAMID MISMATCH', c, (w, h, B, nl, sc), 'frame', f, 'level', l); bad += 1
        ko, do, _ = O.orb_extract(lv, sfo, tgt, a, b)
        n = counts[f]
        if n != len(ko) or kps[f, :n].tobytes() != ko.tobytes() or desc[f, :n].tobytes() != do.tobytes():
            print('EXTRACT MISMATCH', c, (w, h, B, nl, sc, tgt, a, b), 'frame', f, n, len(ko)); bad += 1
    if B >= 2:   # stereo row-band matcher and LK tracker on frames (0, 1) of the batch, host entries vs oracle
        n0, n1 = counts[0], counts[1]
        inv_sf = (np.float32(1.0) / sf).astype(np.float32)
        th, tol = int(rng.choice([30, 75, 120])), int(rng.integers(0, 3))
        m, z = ctx.stereo_match(tb.StereoParams(bf=386.0, th_dist=th, octave_tol=tol), kps[0, :n0], desc[0, :n0], kps[1, :n1], desc[1, :n1], inv_sf, h)
        em, ez = O.stereo_rowband(O.StereoParams(bf=386.0, th_dist=th, octave_tol=tol), kps[0, :n0], desc[0, :n0], kps[1, :n1], desc[1, :n1], inv_sf, h)
        if m.tobytes() != em.tobytes() or z.tobytes() != ez.tobytes():
            print('STEREO MISMATCH', c, (w, h, th, tol), len(m), len(em)); bad += 1
        nxt = np.roll(frames[0], (int(rng.integers(-3, 4)), int(rng.integers(-6, 7))), axis=(0, 1)) if rng.random() < 0.7 else frames[1]
        lkp = dict(win=int(rng.choice([21, 21, 15, 9])), max_level=int(rng.integers(0, 5)), max_iter=int(rng.choice([30, 10, 3])))
        pts = np.stack([kps[0, :n0]['x'], kps[0, :n0]['y']], 1)
        if n0:
            pts[: max(1, n0 // 20)] += rng.uniform(-40, 40, (max(1, n0 // 20), 2)).astype(np.float32)   # some points near / past the border
        gp, gs, ge = ctx.lk_flow(tb.LKParams(**lkp), frames[0], nxt, pts)
        ep, es, ee = O.lk_flow(O.LKParams(**lkp), frames[0], nxt, pts)
        if gp.tobytes() != ep.tobytes() or not (gs == es).all() or ge.tobytes() != ee.tobytes():
            print('LK MISMATCH', c, (w, h), lkp, int((gs != es).sum()), float(np.abs(gp - ep).max())); bad += 1
    fe.close()
    print('case', c, (w, h, B, nl, round(sc, 3), tgt, a, b), 'ok' if not bad else 'BAD so far', flush=True)
print('done in %.1f s, mismatches: %d' % (time.time() - t0, bad))
sys.exit(1 if bad else 0)
