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
