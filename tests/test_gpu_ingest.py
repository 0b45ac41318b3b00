"""GPU parity for the ingest row (PinholeCamera::UndistortImage remap, Frame::Equalize CLAHE) through the C ABI:
bit-exact against the real OpenCV (cv2, where importable) and the oracle."""
import numpy as np
import pytest
import torch

from .conftest import PARAM_SETS

pytestmark = pytest.mark.gpu

CAMERAS = [  # (w, h, fx, fy, cx, cy, d[5])
    (1241, 376, 718.856, 718.856, 607.1928, 185.2157, [-0.28, 0.07, 0.0002, 0.00002, 0.0]),
    (752, 480, 458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0]),
    (640, 480, 517.3, 516.5, 318.6, 255.3, [0.2624, -0.9531, -0.0054, 0.0026, 1.1633]),
]


def _cam(tb, c):
    w, h, fx, fy, cx, cy, d = c
    return tb.Pinhole(fx, fy, cx, cy, w, h, (tb.capi.C.c_float * 5)(*d))


@pytest.mark.parametrize('c', CAMERAS)
def test_undistort_image_matches_cv2_and_oracle(tb, ctx, oracle, c):
    import cv2
    w, h, fx, fy, cx, cy, d = c
    m1, m2 = tb.init_undistort_rectify_map(_cam(tb, c))
    o1, o2 = oracle.init_undistort_rectify_map(fx, fy, cx, cy, d, w, h)
    assert m1.tobytes() == o1.tobytes() and m2.tobytes() == o2.tobytes()
    und = tb.Undistorter(ctx, m1, m2)
    rng = np.random.default_rng(w + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    got = und.apply_host(img)
    assert np.array_equal(got, oracle.remap_linear(img, m1, m2))
    assert np.array_equal(got, cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    und.close()


def test_remap_arbitrary_maps_borders_nonfinite(tb, ctx, oracle):
    rng = np.random.default_rng(0)
    m1 = rng.uniform(-50, 700, (301, 403)).astype(np.float32)
    m2 = rng.uniform(-50, 530, (301, 403)).astype(np.float32)
    m1[0, 0] = np.nan; m2[1, 1] = np.inf; m1[2, 2] = 1e9; m1[3, 3] = -1e9; m2[4, 4] = -np.inf; m1[7, 7] = 3e38
    m1[5, :8] = [-1.0, -0.5, -0.015625, 0.0, 638.5, 639.0, 639.5, 640.0]; m2[5, :8] = 100.0
    m2[6, :8] = [-1.0, -0.5, -0.015625, 0.0, 478.5, 479.0, 479.5, 480.0]; m1[6, :8] = 100.0
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)     # source and destination sizes differ
    und = tb.Undistorter(ctx, m1, m2)
    assert np.array_equal(und.apply_host(img), oracle.remap_linear(img, m1, m2))
    und.close()


def test_remap_batch_device_strided(tb, ctx, oracle):
    c = CAMERAS[1]
    w, h = c[0], c[1]
    m1, m2 = tb.init_undistort_rectify_map(_cam(tb, c))
    und = tb.Undistorter(ctx, m1, m2)
    rng = np.random.default_rng(3)
    B, sp, dp = 7, 768, 832                                    # 7 frames: exercises the partial group of 4
    raw = np.zeros((B, h, sp), np.uint8)
    raw[:, :, :w] = rng.integers(0, 256, (B, h, w), dtype=np.uint8)
    t_src = torch.from_numpy(raw).cuda()
    t_dst = torch.full((B, h, dp), 77, dtype=torch.uint8, device='cuda')
    und.apply_device(t_src.data_ptr(), w, h, sp, sp * h, t_dst.data_ptr(), dp, dp * h, B)
    ctx.sync()
    out = t_dst.cpu().numpy()
    for f in range(B):
        assert np.array_equal(out[f, :, :w], oracle.remap_linear(raw[f, :, :w], m1, m2)), f
    assert np.all(out[:, :, w:] == 77)
    und.close()


@pytest.mark.parametrize('shape', [(376, 1241), (480, 640), (1080, 1920), (377, 1243), (301, 535), (64, 72), (9, 17)])
def test_clahe_matches_cv2_and_oracle(tb, ctx, oracle, images, shape):
    import cv2
    h, w = shape
    rng = np.random.default_rng(h * w)
    cases = [rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(100, 130, (h, w)).astype(np.uint8)]
    if h <= 376 and w <= 1241:
        cases.append(np.ascontiguousarray(images['left'][:h, :w]))
    for img in cases:
        got = ctx.clahe_host(img, 3.0, (8, 8))
        assert np.array_equal(got, oracle.clahe(img, 3.0, (8, 8)))
        assert np.array_equal(got, cv2.createCLAHE(3.0, (8, 8)).apply(img))
    got = ctx.clahe_host(cases[0], 40.0, (4, 6))
    assert np.array_equal(got, oracle.clahe(cases[0], 40.0, (4, 6)))
    got = ctx.clahe_host(cases[0], 0.0, (8, 8))                 # clip limit 0 = plain adaptive equalisation
    assert np.array_equal(got, oracle.clahe(cases[0], 0.0, (8, 8)))


def test_clahe_batch_device_strided(tb, ctx, oracle, images):
    from trackingbench_slam_b200.synth import synth_frame
    W, H, B, pitch = 1241, 376, 5, 1280
    fr = np.zeros((B, H, pitch), np.uint8)
    fr[0, :, :W] = images['left']; fr[1, :, :W] = images['right']
    for k in range(2, B):
        fr[k, :, :W] = synth_frame(k, W, H)
    t_src = torch.from_numpy(fr).cuda()
    t_dst = torch.zeros_like(t_src)
    ctx.clahe_device(t_src.data_ptr(), W, H, pitch, pitch * H, t_dst.data_ptr(), pitch, pitch * H, B, 3.0, (8, 8))
    ctx.sync()
    out = t_dst.cpu().numpy()
    for k in range(B):
        assert np.array_equal(out[k, :, :W], oracle.clahe(fr[k, :, :W], 3.0, (8, 8))), k


def test_frontend_upload_undistort_then_extract(tb, ctx, oracle, images):
    """test_vo.cpp:178 flow: UndistortImage -> Frame (pyramid) -> extract, with the remap writing level 0 in place of an upload."""
    c = CAMERAS[0]
    W, H = c[0], c[1]
    nl, sc, tgt, a, b = PARAM_SETS['A1000']
    m1, m2 = tb.init_undistort_rectify_map(_cam(tb, c))
    und = tb.Undistorter(ctx, m1, m2)
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=2, max_target=tgt)
    try:
        raw = np.stack([images['left'], images['right']])
        fe.upload_frames_undistort(und, 0, raw)
        fe.build_pyramid(2)
        fe.extract(2, sf, tgt, a, b)
        for f in range(2):
            rect = oracle.remap_linear(raw[f], m1, m2)
            assert np.array_equal(fe.download_level(f, 0), rect)
            lv, osf = oracle.pyramid(rect, nl, sc)
            k, d, _ = oracle.orb_extract(lv, osf, tgt, a, b)
            gk, gd = fe.get_results(f)
            assert gk.tobytes() == k.tobytes() and gd.tobytes() == d.tobytes()
    finally:
        fe.close()
        und.close()


def test_undistort_points_host_device_frontend(tb, ctx, oracle, images):
    """PinholeCamera::UndistortPoint / Frame::UnDistortPoints: double arithmetic in OpenCV's order, bit-exact."""
    from .test_oracle_cv2 import POINT_CAMERAS
    for c in POINT_CAMERAS:
        w, h, fx, fy, cx, cy, d = c
        rng = np.random.default_rng(w * 3 + h)
        pts = np.stack([rng.uniform(-200, w + 200, 5000), rng.uniform(-100, h + 100, 5000)], 1).astype(np.float32)
        exp = oracle.undistort_points(np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy), d, pts)
        got = ctx.undistort_points(_cam(tb, c), pts)
        assert got.tobytes() == exp.tobytes(), c
    assert ctx.undistort_points(_cam(tb, POINT_CAMERAS[0]), np.zeros((0, 2), np.float32)).shape == (0, 2)
    # Frame::UnDistortPoints over the keypoints of an extraction (tb_keypoint stride), device-resident
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    c = POINT_CAMERAS[0]
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=2, max_target=tgt)
    try:
        fe.upload_frames(0, np.stack([images['left'], images['right']]))
        fe.build_pyramid(2)
        fe.extract(2, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(2)
        out = torch.full((2, fe.kp_cap, 2), -7.0, dtype=torch.float32, device='cuda')
        fe.undistort_points(_cam(tb, c), 2, out.data_ptr())
        ctx.sync()
        o = out.cpu().numpy()
        for f in range(2):
            n = counts[f]
            exp = oracle.undistort_points(np.float32(c[2]), np.float32(c[3]), np.float32(c[4]), np.float32(c[5]), c[6], np.stack([kps[f, :n]['x'], kps[f, :n]['y']], 1))
            assert o[f, :n].tobytes() == exp.tobytes() and (o[f, n:] == 0).all()
    finally:
        fe.close()
