#!/usr/bin/env python3
"""Generate tests/golden/*.npz — run in the BUILD container only (needs /root/reference + cv2).

Sources of truth, strongest first:
  * extraction: oracle/_ref = the reference's own src/extractors/ORBextractor.cpp compiled unmodified
    (comparator pinned, oracle/ref_build.cpp) on pyramids made by the real cv2.resize;
  * primitives: the real OpenCV (cv2 4.13.0): resize / GaussianBlur level digests, BFMatcher lists;
  * windowed / projection matchers and the search grid: oracle/_ref/libtb_ref_matcher.so = the reference's own
    Frame.cpp / MapPoint.cpp / Map.cpp / CameraModel.cpp compiled unmodified + matcher.cpp's member functions by
    line range (oracle/ref_matcher_build.cpp).  The C restatement must reproduce every list (asserted here).
"""
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from tests.conftest import PARAM_SETS  # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')
imgs = {n: cv2.imread(os.path.join(ROOT, 'trackingbench_slam_b200', 'data', n + '.png'), 0) for n in ('left', 'right', 'disparity')}
assert O.ref_available(), 'oracle/_ref missing: run make -C oracle'
ref = O.RefExtractor()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cv2_pyramid(img, nl, sc):
    sf = O.scale_factors(nl, sc)
    sizes = O.level_sizes(img.shape[1], img.shape[0], sf)
    lv = [img]
    for i in range(1, nl):
        lv.append(cv2.resize(lv[-1], sizes[i]))
    return lv, sf


digests = {}
ext = {}
for pname, (nl, sc, tgt, a, b) in PARAM_SETS.items():
    for iname in ('left', 'right'):
        lv, sf = cv2_pyramid(imgs[iname], nl, sc)
        for l, im in enumerate(lv):
            digests['pyr/%s/%d/%.6f/%d' % (iname, nl, sc, l)] = sha(im)
            digests['blur/%s/%d/%.6f/%d' % (iname, nl, sc, l)] = sha(cv2.GaussianBlur(im, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101))
        k, d = ref(lv, sf, tgt, a, b)
        ext[(iname, pname)] = (k, d)
        np.savez_compressed(os.path.join(G, 'extract_%s_%s.npz' % (iname, pname)), kps=k, desc=d, sf=sf)
        print(iname, pname, len(k))
np.savez_compressed(os.path.join(G, 'cv2_digests.npz'), keys=np.array(sorted(digests)), vals=np.array([digests[k] for k in sorted(digests)]))

# matching goldens
bf = cv2.BFMatcher(cv2.NORM_HAMMING, True)
for pname in PARAM_SETS:
    kl, dl = ext[('left', pname)]
    kr, dr = ext[('right', pname)]
    ms = bf.match(dl, dr)
    cross = np.array([(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in ms], dtype=O.MATCH_DTYPE)
    mind = min(m.distance for m in ms)
    thr = min(np.float32(10.0) * np.float32(mind), np.float32(30.0))
    filt = cross[cross['distance'] < thr]  # searchByBF(F1,F2,0,nLevels,10,30): matcher.cpp:209-218
    out = dict(cross=cross, filtered=filt)
    # searchByViolence with setViolenceParam(30,100,30,true,5) / search_r 50 (test_matcher.cpp:70-72) — run by the reference's code
    H, W = imgs['left'].shape
    nl, sc = PARAM_SETS[pname][0], np.float32(PARAM_SETS[pname][1])
    fx, fy, cx, cy, base = 718.856, 718.856, 607.1928, 185.2157, 0.573
    rcam = O.RefCamera(W, H, fx, fy, cx, cy)
    FL, FR = O.RefFrame(imgs['left'], nl, sc, rcam), O.RefFrame(imgs['right'], nl, sc, rcam)
    FL.set_keys(kl, dl); FL.assign_grid()
    FR.set_keys(kr, dr); FR.assign_grid()
    assert O.RefMatcher().search_by_bf(FL, FR, 0, nl, 10.0, 30.0).tobytes() == filt.tobytes()
    p = O.MatcherParams(30, 100, 30, 1, 5.0)
    out['violence'] = O.RefMatcher(30, 100, 30, True, 5.0).search_by_violence(FL, FR, 0, nl, 50.0)
    assert out['violence'].tobytes() == O.search_by_violence(p, kl, dl, kr, dr, W, H, 0, nl, 50.0).tobytes()
    # config 2: stereo projection. Map points from the LEFT keypoints with depth = b*fx/disparity
    # (test_vo.cpp:414,496-498), right camera pose = translation (-0.573,0,0); searchByProjection(R, L)
    cam = O.Pinhole(fx, fy, cx, cy, W, H, (O.C.c_float * 5)(0, 0, 0, 0, 0))
    disp = imgs['disparity']
    u = np.clip(kl['x'].astype(np.int32), 0, W - 1)
    v = np.clip(kl['y'].astype(np.int32), 0, H - 1)
    dsp = disp[v, u].astype(np.float32)
    valid = (dsp > 0).astype(np.uint8)
    depth = np.where(dsp > 0, np.float32(base * fx) / np.maximum(dsp, 1), 0).astype(np.float32)
    xyz = np.stack([(kl['x'] - np.float32(cx)) / np.float32(fx) * depth, (kl['y'] - np.float32(cy)) / np.float32(fy) * depth, depth], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[0, 3] = -base
    FL.set_pose(np.eye(4, dtype=np.float32))
    FR.set_pose(T)
    # one MapPoint per left key (MapPoint::MapPoint: descriptor of the key, normal from the left camera centre); keys without
    # disparity carry a point flagged bad, so slot i of the Map == key i and trainIdx counts every slot (matcher.cpp:549-553)
    mps = O.RefMapPoints(FL, np.ones(len(kl), np.uint8), xyz, attach=True)
    for i in np.flatnonzero(valid == 0):
        mps.set_bad(i)
    sf = O.scale_factors(nl, sc)
    pp = O.MatcherParams(30, 50, 30, 1, 30.0)  # setProjectionParam(30,50,30,true,30)
    out['projection'] = O.RefMatcher(30, 50, 30, True, 30.0).search_by_projection(FR, FL)
    assert out['projection'].tobytes() == O.search_by_projection(pp, cam, T, kr, dr, None, sf, W, H, kl, valid, xyz, dl).tobytes()
    out['mp_valid'] = valid
    out['mp_xyz'] = xyz
    nrm = np.stack([mps.get(i)[1] for i in range(len(kl))]).astype(np.float32)
    nrm[valid == 0] = 0
    pm = O.MatcherParams(30, 100, 30, 1, 5.0)
    out['projection_map'] = O.RefMatcher(30, 100, 30, True, 5.0).search_by_projection_map(mps, FR, 0.8)
    assert out['projection_map'].tobytes() == O.search_by_projection_map(pm, cam, T, kr, dr, None, sf, W, H, 1 - valid, xyz, nrm, dl, 0.8).tobytes()
    out['mp_normal'] = nrm
    mps.close(); FL.close(); FR.close()
    np.savez_compressed(os.path.join(G, 'match_%s.npz' % pname), **out)
    print(pname, 'cross', len(cross), 'filtered', len(filt), 'violence', len(out['violence']),
          'projection', len(out['projection']), 'projection_map', len(out['projection_map']))
