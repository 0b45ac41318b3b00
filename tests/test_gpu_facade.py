"""The C++ facade (include/extractors, include/matchers, include/types keep the reference's class
names) driven by a headless restatement of test/test_matcher.cpp, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from .conftest import ROOT, PARAM_SETS

pytestmark = pytest.mark.gpu


def _fnv(b, h=1469598103934665603):
    for x in b:
        h ^= x
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _digest(m):
    h = 1469598103934665603
    for r in m:
        h = _fnv(np.array([r['queryIdx'], r['trainIdx'], r['imgIdx']], np.int32).tobytes(), h)
        h = _fnv(np.float32(r['distance']).tobytes(), h)
    return h


def test_headless_test_matcher(tmp_path, oracle, images):
    exe = str(tmp_path / 'test_matcher_headless')
    libdir = os.path.join(ROOT, 'trackingbench_slam_b200')
    subprocess.check_call(['g++', '-std=c++14', '-O2', '-I' + os.path.join(ROOT, 'include'), '-o', exe,
                           os.path.join(ROOT, 'examples', 'test_matcher_headless.cpp'), '-L' + libdir, '-ltb_b200',
                           '-Wl,-rpath,' + libdir])
    paths = []
    for n in ('left', 'right', 'disparity'):
        p = str(tmp_path / (n + '.pgm'))
        im = images[n]
        with open(p, 'wb') as f:
            f.write(b'P5\n%d %d\n255\n' % (im.shape[1], im.shape[0]))
            f.write(im.tobytes())
        paths.append(p)
    from trackingbench_slam_b200 import vocab as V
    voc = V.synth_vocabulary(10, 5, 4)   # levelsup is fixed to 4 by SetBow: buckets at level 1
    vpath = str(tmp_path / 'voc.txt')
    voc.save_text(vpath)
    paths.append(vpath)
    out = subprocess.check_output([exe] + paths).decode().split('\n')
    got = {l.split()[0]: l.split()[1:] for l in out if l.strip()}
    nl, sc, tgt, a, b = PARAM_SETS['A1000']
    ext = []
    for n in ('left', 'right'):
        lv, sf = oracle.pyramid(images[n], nl, sc)
        ext.append(oracle.orb_extract(lv, sf, tgt, a, b))
    (k1, d1, _), (k2, d2, _) = ext
    assert got['kps'] == [str(len(k1)), str(len(k2))]
    assert got['kps_digest'] == ['%x' % _fnv(k1.tobytes()), '%x' % _fnv(k2.tobytes())]
    assert got['desc_digest'] == ['%x' % _fnv(d1.tobytes()), '%x' % _fnv(d2.tobytes())]
    bf = oracle.search_by_bf(d1, d2, 10.0, 30.0)
    assert got['bf'] == [str(len(bf)), '%x' % _digest(bf)]
    H, W = images['left'].shape
    vio = oracle.search_by_violence(oracle.MatcherParams(30, 100, 30, 1, 5.0), k1, d1, k2, d2, W, H, 0, 5, 50.0)
    assert got['violence'] == [str(len(vio)), '%x' % _digest(vio)]
    fx, fy, cx, cy, base = np.float32(718.856), np.float32(718.856), np.float32(607.1928), np.float32(185.2157), np.float32(0.573)
    disp = images['disparity'][k1['y'].astype(np.int32), k1['x'].astype(np.int32)].astype(np.float32)
    valid = (disp > 0).astype(np.uint8)
    depth = np.where(disp > 0, (base * fx) / np.maximum(disp, 1), 0).astype(np.float32)
    xyz = np.stack([(k1['x'] - cx) / fx * depth, (k1['y'] - cy) / fy * depth, depth], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[0, 3] = -base
    cam = oracle.Pinhole(718.856, 718.856, 607.1928, 185.2157, W, H, (oracle.C.c_float * 5)(0, 0, 0, 0, 0))
    sf = oracle.scale_factors(nl, sc)
    proj = oracle.search_by_projection(oracle.MatcherParams(30, 50, 30, 1, 30.0), cam, T, k2, d2, None, sf, W, H, k1, valid, xyz, d1)
    assert got['projection'] == [str(len(proj)), '%x' % _digest(proj)]
    assert got['dist01'] == [str(oracle.hamming256(d1[0], d1[1]))]
    # AddPoints with a larger target than the last operator(): budget = that of operator()(…, 1000, …) (ORBextractor.cpp:840-904)
    lv, sf = oracle.pyramid(images['left'], nl, sc)
    ko2, _, _ = oracle.orb_extract(lv, sf, 1500, a, b, exit_kps=k1[::5], budget=oracle.orb_budget(nl, sf[1], 1000))
    assert got['addpoints'] == [str(len(ko2)), '%x' % _fnv(ko2.tobytes())]
    inside = k1[(k1['x'] < 101) & (k1['y'] < 57)]
    q = (inside['x'] >= 51).astype(int) + 2 * (inside['y'] >= 29).astype(int)
    assert got['dividenode'] == [str(len(inside))] + [str(int((q == i).sum())) for i in range(4)] + ['51', '29']
    # PinholeCamera::UndistortImage / UndistortPoint with a distorted camera
    dd = [-0.28, 0.07, 0.0002, 0.00002, 0.0]
    m1, m2 = oracle.init_undistort_rectify_map(718.856, 718.856, 607.1928, 185.2157, dd, W, H)
    rect = oracle.remap_linear(images['left'], m1, m2)
    un = oracle.undistort_points(718.856, 718.856, 607.1928, 185.2157, dd, np.stack([k1['x'], k1['y']], 1))
    assert got['undistort'] == ['%x' % _fnv(rect.tobytes()), '%x' % _fnv(un.tobytes())]
    # searchByOPFlow(F1 = right frame, F2 = left frame with its keypoints), raw and CLAHE'd current image
    for eq in (0, 1):
        img1 = oracle.clahe(images['right'], 3.0, (8, 8)) if eq else images['right']
        pts, st, _ = oracle.lk_flow(oracle.LKParams(), images['left'], img1, np.stack([k1['x'], k1['y']], 1))
        xi, yi = pts[:, 0].astype(np.int32), pts[:, 1].astype(np.int32)
        idx = np.nonzero((st == 1) & (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H))[0].astype(np.int32)
        h = _fnv(pts.tobytes())
        for i in idx:
            h = _fnv(np.int32(i).tobytes(), h)
        assert len(idx) > 300 and got['opflow%d' % eq] == [str(len(idx)), '%x' % h]
        if eq:   # reject = true: Matcher::rejectWithF on the tracks that are left (matcher.cpp:751-755)
            st2 = np.zeros(len(pts), np.uint8)
            st2[idx] = 1
            st3, _, it = oracle.reject_with_f(pts, np.stack([k1['x'], k1['y']], 1), st2)
            idx3 = np.nonzero(st3)[0].astype(np.int32)
            h = _fnv(pts.tobytes())
            for i in idx3:
                h = _fnv(np.int32(i).tobytes(), h)
            assert it > 0 and 8 <= len(idx3) < len(idx) and got['opflow_reject'] == [str(len(idx3)), '%x' % h]
    # bag of words: Frame::SetBow (levelsup 4) + searchByBow, with and without MapPointOnly
    w1, n1, wt1 = oracle.bow_transform(voc, d1, 4)
    w2, n2, wt2 = oracle.bow_transform(voc, d2, 4)
    ids, vals = V.bow_vector(w1, wt1, voc.weighting, voc.scoring)
    h = 1469598103934665603
    for i, v in zip(ids.tolist(), vals.tolist()):
        h = _fnv(np.uint32(i).tobytes(), h)
        h = _fnv(np.float64(v).tobytes(), h)
    assert got['bowvec'] == [str(len(ids)), '%x' % h]
    fv1, fv2 = V.feature_vector(n1, wt1), V.feature_vector(n2, wt2)
    pb = oracle.MatcherParams(50, 100, 30, 1, 0.9)
    bow = oracle.search_by_bow(pb, k1, d1, fv1, k2, d2, fv2)
    assert len(bow) > 10 and got['bow'] == [str(len(bow)), '%x' % _digest(bow)]
    bow_mp = oracle.search_by_bow(pb, k2, d2, fv2, k1, d1, fv1, valid, True)
    assert len(bow_mp) > 10 and got['bow_mp'] == [str(len(bow_mp)), '%x' % _digest(bow_mp)]
