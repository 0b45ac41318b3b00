"""Pin the oracle's OpenCV-primitive restatements against the real OpenCV (cv2 4.13): bit-exact.
(the reference's own tests hold no golden vectors for this path.)"""
import cv2
import numpy as np
import pytest


def test_resize_matches_cv2(oracle, images):
    rng = np.random.default_rng(0)
    for img in (images['left'], rng.integers(0, 256, (270, 480), dtype=np.uint8)):
        for nl, sc in ((8, 1 / 1.2), (5, 0.8), (5, 0.6), (4, 0.5)):
            lv, sf = oracle.pyramid(img, nl, sc)
            prev = img
            for i in range(1, nl):
                ref = cv2.resize(prev, (lv[i].shape[1], lv[i].shape[0]))
                assert np.array_equal(ref, lv[i]), (nl, sc, i)
                prev = ref


def test_resize_exact_half_uses_area(oracle):
    img = np.random.default_rng(1).integers(0, 256, (64, 96), dtype=np.uint8)
    assert np.array_equal(oracle.resize_linear(img, 48, 32), cv2.resize(img, (48, 32)))


def test_level_sizes_float32_rule(oracle):
    sf = oracle.scale_factors(8, 1 / 1.2)
    assert oracle.level_sizes(1920, 1080, sf)[1] == (1600, 900)  # 1920*0.83333331f rounds to 1600.0f
    assert [s[0] for s in oracle.level_sizes(1241, 376, sf)] == [1241, 1034, 861, 718, 598, 498, 415, 346]


def test_gauss_matches_cv2(oracle, images):
    rng = np.random.default_rng(2)
    for img in (images['left'], rng.integers(0, 256, (61, 97), dtype=np.uint8), rng.integers(0, 256, (9, 7), dtype=np.uint8)):
        ref = cv2.GaussianBlur(img, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, oracle.gauss7x7(img))


@pytest.mark.parametrize('T', [7, 20, 30, 80])
def test_fast_matches_cv2(oracle, images, T):
    rng = np.random.default_rng(3)
    for img in (images['left'][:200, :400], images['left'][100:140, 200:237], rng.integers(0, 256, (64, 80), dtype=np.uint8)):
        det = cv2.FastFeatureDetector_create(T, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        kps = det.detect(img)
        mine = oracle.fast9(img, T)
        assert len(kps) == len(mine)
        for a, b in zip(kps, mine):
            assert a.pt == (b['x'], b['y']) and a.response == b['response']


def test_fast_atan2_matches_cv2(oracle):
    rng = np.random.default_rng(4)
    ys = rng.normal(0, 1000, 5000).astype(np.float32)
    xs = rng.normal(0, 1000, 5000).astype(np.float32)
    for y, x in list(zip(ys, xs)) + [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (5, 5), (-5, 5)]:
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(oracle.fast_atan2(y, x))


def test_bf_cross_matches_cv2(oracle):
    rng = np.random.default_rng(5)
    d1 = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    d2[:100] = d1[50:150]
    d2[7] = d2[8]
    d1[3] = d1[4]  # duplicates => ties, lowest index must win
    ms = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(d1, d2)
    mine = oracle.bf_cross(d1, d2)
    assert len(ms) == len(mine)
    for a, b in zip(ms, mine):
        assert (a.queryIdx, a.trainIdx, a.distance) == (b['queryIdx'], b['trainIdx'], b['distance'])


def test_hamming_swar(oracle):
    rng = np.random.default_rng(6)
    a = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert oracle.hamming256(x, y) == int(np.unpackbits(x ^ y).sum())


# ---- frame ingest: undistortion maps, remap, CLAHE — pinned to the real OpenCV
CAMERAS = [  # (w, h, fx, fy, cx, cy, d[5])
    (1241, 376, 718.856, 718.856, 607.1928, 185.2157, [-0.28, 0.07, 0.0002, 0.00002, 0.0]),
    (752, 480, 458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0]),
    (640, 480, 517.3, 516.5, 318.6, 255.3, [0.2624, -0.9531, -0.0054, 0.0026, 1.1633]),
]


def _cv_maps(w, h, fx, fy, cx, cy, d):
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)          # CameraModel.cpp:15-16: float matrices
    D = np.array(d, np.float32).reshape(1, 5)
    return cv2.initUndistortRectifyMap(K, D, None, K, (w, h), cv2.CV_32F)


@pytest.mark.parametrize('cam', CAMERAS)
def test_undistort_maps_and_remap_match_cv2(oracle, cam):
    w, h, fx, fy, cx, cy, d = cam
    m1, m2 = _cv_maps(*cam)
    a1, a2 = oracle.init_undistort_rectify_map(fx, fy, cx, cy, d, w, h)
    assert a1.tobytes() == m1.tobytes() and a2.tobytes() == m2.tobytes()
    img = np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(oracle.remap_linear(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))


def test_remap_borders_and_nonfinite_match_cv2(oracle):
    rng = np.random.default_rng(0)
    m1 = rng.uniform(-50, 700, (300, 400)).astype(np.float32)
    m2 = rng.uniform(-50, 530, (300, 400)).astype(np.float32)
    m1[0, 0] = np.nan; m2[1, 1] = np.inf; m1[2, 2] = 1e9; m1[3, 3] = -1e9; m2[4, 4] = -np.inf
    m1[5, :8] = [-1.0, -0.5, -0.015625, 0.0, 638.5, 639.0, 639.5, 640.0]
    m2[5, :8] = 100.0
    m2[6, :8] = [-1.0, -0.5, -0.015625, 0.0, 478.5, 479.0, 479.5, 480.0]
    m1[6, :8] = 100.0
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    assert np.array_equal(oracle.remap_linear(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))


@pytest.mark.parametrize('shape', [(376, 1241), (480, 640), (1080, 1920), (377, 1243), (301, 535), (64, 72), (9, 17)])
def test_clahe_matches_cv2(oracle, images, shape):
    h, w = shape
    rng = np.random.default_rng(h * w)
    cases = [rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(100, 130, (h, w)).astype(np.uint8)]
    if h <= 376 and w <= 1241:
        cases.append(np.ascontiguousarray(images['left'][:h, :w]))
    for img in cases:
        ref = cv2.createCLAHE(3.0, (8, 8)).apply(img)                          # Frame::Equalize, Frame.cpp:453-458
        assert np.array_equal(oracle.clahe(img, 3.0, (8, 8)), ref)
    ref = cv2.createCLAHE(40.0, (4, 6)).apply(cases[0])
    assert np.array_equal(oracle.clahe(cases[0], 40.0, (4, 6)), ref)


@pytest.mark.parametrize('shape', [(376, 1241), (61, 83), (100, 101), (23, 45), (1, 7), (7, 1)])
def test_pyrdown_matches_cv2(oracle, shape):
    img = np.random.default_rng(shape[0]).integers(0, 256, shape, dtype=np.uint8)
    assert (oracle.pyrdown(img) == cv2.pyrDown(img)).all()


def lk_inputs(images):
    """(prev, next, pts) cases for the LK tests (shared with tests/test_gpu_flow.py): the KITTI stereo pair with the golden
    ORB keypoints, a shifted/noisy synthetic pair with points up to the image border, and a small image that stops the
    pyramid early."""
    import os
    from .conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'extract_left_B1000.npz'))
    cases = [('kitti', images['left'], images['right'], np.stack([g['kps']['x'], g['kps']['y']], 1).astype(np.float32))]
    rng = np.random.default_rng(5)
    base = cv2.GaussianBlur(rng.integers(0, 256, (200, 260), dtype=np.uint8), (0, 0), 2.0)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    M = np.float32([[1.01, 0.01, 2.3], [-0.01, 0.99, -1.7]])
    nxt = cv2.warpAffine(base, M, (260, 200), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    pts = np.stack([rng.uniform(-3, 263, 400), rng.uniform(-3, 203, 400)], 1).astype(np.float32)
    pts[:8] = [[0, 0], [259, 199], [0.5, 100], [130, 0.25], [259.5, 50], [100, 199.75], [-30, 20], [20, 400]]
    cases.append(('affine', base, nxt, pts))
    small = base[:50, :70].copy()
    cases.append(('small', small, np.roll(small, 1, axis=1), np.stack([rng.uniform(0, 70, 60), rng.uniform(0, 50, 60)], 1).astype(np.float32)))
    flat = np.full((120, 160), 77, np.uint8)
    cases.append(('flat', flat, flat, np.stack([rng.uniform(0, 160, 30), rng.uniform(0, 120, 30)], 1).astype(np.float32)))
    return cases


@pytest.mark.parametrize('max_level', [3, 0])
def test_lk_flow_matches_cv2_within_tolerance(oracle, images, max_level):
    """cv2's SIMD build sums the 2x2 system in float lanes, the restatement in exact integers: most tracks are bit-identical,
    the rest differ by rounding, and a stop decision on the 0.01-pixel threshold may flip (then the difference is of that
    order). Tolerance: status equal on >= 99 % of the points; of the tracks both call valid, >= 98 % within 0.05 px and
    the median within 1e-3 px."""
    for name, prev, nxt, pts in lk_inputs(images):
        n2, s2, e2 = oracle.lk_flow(oracle.LKParams(max_level=max_level), prev, nxt, pts)
        n1, s1, e1 = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=max_level)
        n1, s1, e1 = n1.reshape(-1, 2), s1.ravel(), e1.ravel()
        assert (s1 == s2).mean() >= 0.99, (name, (s1 == s2).mean())
        ok = (s1 == 1) & (s2 == 1)
        if name == 'flat':
            assert ok.sum() == 0      # no texture: minEig below the threshold everywhere
            continue
        assert ok.sum() > len(pts) // 3, name
        d = np.abs(n1 - n2).max(1)[ok]
        assert (d <= 0.05).mean() >= 0.98 and np.median(d) <= 1e-3, (name, (d <= 0.05).mean(), np.median(d))
        close = ok.copy()
        close[ok] = d <= 1e-3
        assert np.abs(e1 - e2)[close].max() <= 0.05, name


POINT_CAMERAS = CAMERAS + [
    (1241, 376, 718.856, 700.5, 607.1928, 185.2157, [-3.0, 2.0, 0.1, 0.1, -1.0]),     # icdist < 0 branch for far points
    (1241, 376, 718.856, 718.856, 607.1928, 185.2157, [1e-8, 0.2, 0.0, 0.0, 0.0]),    # mbDistortion false: identity
]


@pytest.mark.parametrize('cam', POINT_CAMERAS)
def test_undistort_points_match_cv2(oracle, cam):
    """PinholeCamera::UndistortPoint = cv::undistortPoints(src, dst, K, D, Mat(), K) (CameraModel.cpp:104-121): bit-exact."""
    w, h, fx, fy, cx, cy, d = cam
    rng = np.random.default_rng(w + h)
    pts = np.stack([rng.uniform(-200, w + 200, 3000), rng.uniform(-100, h + 100, 3000)], 1).astype(np.float32)
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    D = np.array(d, np.float32)
    got = oracle.undistort_points(K[0, 0], K[1, 1], K[0, 2], K[1, 2], D, pts)
    if abs(D[0]) > 1e-7:
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K.astype(np.float64), D.astype(np.float64), None, K.astype(np.float64)).reshape(-1, 2)
    else:
        ref = pts
    assert got.tobytes() == ref.tobytes()


def fm_cases(images, seeds=range(12), max_n=2200):
    """(name, cur_pts, last_pts, status) inputs of Matcher::rejectWithF: synthetic two-view tracks (pinhole camera, small
    rotation + translation, 0.3 px noise, up to 45 % gross outliers, 10 % of the tracks already lost) and the real LK tracks
    of the KITTI pair."""
    cases = []
    for seed in seeds:
        rs = np.random.default_rng(seed)
        n = int(rs.integers(15, max_n))
        X = np.c_[rs.uniform(-5, 5, n), rs.uniform(-3, 3, n), rs.uniform(4, 30, n)]
        K = np.array([[700, 0, 620], [0, 700, 188], [0, 0, 1.]])
        Rm, _ = cv2.Rodrigues(rs.uniform(-0.05, 0.05, 3))
        tv = rs.uniform(-0.5, 0.5, 3)
        p1 = (K @ X.T).T
        p1 = p1[:, :2] / p1[:, 2:]
        p2 = (K @ ((Rm @ X.T).T + tv).T).T
        p2 = p2[:, :2] / p2[:, 2:]
        p1 += rs.normal(0, 0.3, p1.shape)
        p2 += rs.normal(0, 0.3, p2.shape)
        o = rs.choice(n, int(n * rs.uniform(0, 0.45)), replace=False)
        p2[o] += rs.uniform(-40, 40, (len(o), 2))
        st = (rs.random(n) >= 0.1).astype(np.uint8)
        cases.append(('synthetic%d' % seed, p2.astype(np.float32), p1.astype(np.float32), st))
    name, prev, nxt, pts = lk_inputs(images)[0]
    n1, s1, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    cases.append(('kitti_lk', n1.reshape(-1, 2).astype(np.float32), pts, s1.ravel().astype(np.uint8)))
    few = cases[0]
    cases.append(('too_few', few[1][:40], few[2][:40], (np.arange(40) < 14).astype(np.uint8)))   # 14 valid tracks: OpenCV would run LMedS; unchanged
    cases.append(('n8', few[1][:8], few[2][:8], np.ones(8, np.uint8)))                             # matcher.cpp:870: not called
    return cases


def cv2_reject_with_f(cur, last, status):
    """Matcher::rejectWithF with the real OpenCV (matcher.cpp:853-881)."""
    st = status.copy()
    ids = np.flatnonzero(st)
    if len(cur) > 8 and len(ids) >= 15:
        F, mask = cv2.findFundamentalMat(cur[ids], last[ids], cv2.FM_RANSAC, 1.0, 0.99)
        if mask is not None:
            st[ids[mask.ravel() == 0]] = 0
    return st


def test_reject_with_f_matches_cv2(oracle, images):
    """The restated RANSAC (cv::RNG, getSubset, 7-point solver, solveCubic, error, iteration update) keeps exactly the tracks
    cv2.findFundamentalMat keeps — tolerance 0 on the masks; F agrees to 1e-6 relative (the null-space basis differs)."""
    for name, cur, last, st in fm_cases(images, seeds=range(40)):
        exp = cv2_reject_with_f(cur, last, st)
        got, F, it = oracle.reject_with_f(cur, last, st)
        assert np.array_equal(got, exp), (name, int(exp.sum()), int(got.sum()))
        if name.startswith('synthetic') or name == 'kitti_lk':
            assert 0 < it <= 1000 and got.sum() < st.sum() and got.sum() >= 8
            Fc, _ = cv2.findFundamentalMat(cur[st > 0], last[st > 0], cv2.FM_RANSAC, 1.0, 0.99)
            assert np.abs(F - Fc).max() <= 1e-6 * np.abs(Fc).max(), name
        else:
            assert it == 0 and np.array_equal(got, st)


def test_reject_with_f_golden(oracle):
    import os
    from .conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'reject_f.npz'))
    for i in range(int(g['n_cases'])):
        got, F, it = oracle.reject_with_f(g['cur%d' % i], g['last%d' % i], g['status_in%d' % i])
        assert np.array_equal(got, g['status_out%d' % i]) and it == int(g['iters%d' % i]), i
