"""Oracle restatement vs the committed golden vectors (made by scripts/make_golden.py from the
reference's own ORBextractor.cpp compiled unmodified = oracle/_ref, and from cv2) and, where
oracle/_ref is present (build container), vs that build directly."""
import hashlib
import os

import numpy as np
import pytest

from .conftest import GOLDEN, PARAM_SETS, kp_bytes_equal


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize('pname', list(PARAM_SETS))
@pytest.mark.parametrize('iname', ['left', 'right'])
def test_extract_matches_golden(oracle, images, iname, pname):
    nl, sc, tgt, a, b = PARAM_SETS[pname]
    g = np.load(os.path.join(GOLDEN, 'extract_%s_%s.npz' % (iname, pname)))
    lv, sf = oracle.pyramid(images[iname], nl, sc)
    assert np.array_equal(sf, g['sf'])
    k, d, _ = oracle.orb_extract(lv, sf, tgt, a, b)
    assert kp_bytes_equal(k, g['kps'])
    assert np.array_equal(d, g['desc'])


def test_pyramid_and_blur_digests(oracle, images):
    g = np.load(os.path.join(GOLDEN, 'cv2_digests.npz'))
    dig = dict(zip(g['keys'].tolist(), g['vals'].tolist()))
    for iname in ('left', 'right'):
        for nl, sc in ((5, 0.8), (8, 1 / 1.2)):
            lv, _ = oracle.pyramid(images[iname], nl, sc)
            for l, im in enumerate(lv):
                assert _sha(im) == dig['pyr/%s/%d/%.6f/%d' % (iname, nl, sc, l)]
                assert _sha(oracle.gauss7x7(im)) == dig['blur/%s/%d/%.6f/%d' % (iname, nl, sc, l)]


@pytest.mark.parametrize('pname', list(PARAM_SETS))
def test_matchers_match_golden(oracle, images, pname):
    nl, sc = PARAM_SETS[pname][:2]
    gl = np.load(os.path.join(GOLDEN, 'extract_left_%s.npz' % pname))
    gr = np.load(os.path.join(GOLDEN, 'extract_right_%s.npz' % pname))
    gm = np.load(os.path.join(GOLDEN, 'match_%s.npz' % pname))
    H, W = images['left'].shape
    assert oracle.bf_cross(gl['desc'], gr['desc']).tobytes() == gm['cross'].tobytes()
    assert oracle.search_by_bf(gl['desc'], gr['desc'], 10.0, 30.0).tobytes() == gm['filtered'].tobytes()
    p = oracle.MatcherParams(30, 100, 30, 1, 5.0)
    assert oracle.search_by_violence(p, gl['kps'], gl['desc'], gr['kps'], gr['desc'], W, H, 0, nl, 50.0).tobytes() == gm['violence'].tobytes()
    cam = oracle.Pinhole(718.856, 718.856, 607.1928, 185.2157, W, H, (oracle.C.c_float * 5)(0, 0, 0, 0, 0))
    T = np.eye(4, dtype=np.float32)
    T[0, 3] = -0.573
    pp = oracle.MatcherParams(30, 50, 30, 1, 30.0)
    got = oracle.search_by_projection(pp, cam, T, gr['kps'], gr['desc'], None, gl['sf'], W, H, gl['kps'], gm['mp_valid'], gm['mp_xyz'], gl['desc'])
    assert got.tobytes() == gm['projection'].tobytes()
    pm = oracle.MatcherParams(30, 100, 30, 1, 5.0)
    got = oracle.search_by_projection_map(pm, cam, T, gr['kps'], gr['desc'], None, gl['sf'], W, H, 1 - gm['mp_valid'], gm['mp_xyz'], gm['mp_normal'], gl['desc'], 0.8)
    assert got.tobytes() == gm['projection_map'].tobytes()


def test_counts_match_survey():
    # SURVEY.md Appendix C (verbatim reference build, pinned comparator)
    n = {p: len(np.load(os.path.join(GOLDEN, 'extract_left_%s.npz' % p))['kps']) for p in PARAM_SETS}
    assert n == {'A1000': 1003, 'A2000': 2003, 'B1000': 1010, 'B2000': 2007}
    c = {p: (len(np.load(os.path.join(GOLDEN, 'match_%s.npz' % p))['cross']), len(np.load(os.path.join(GOLDEN, 'match_%s.npz' % p))['filtered'])) for p in PARAM_SETS}
    assert c == {'A1000': (443, 121), 'A2000': (934, 266), 'B1000': (425, 124), 'B2000': (842, 250)}


def test_restatement_equals_reference_build(oracle, images):
    if not oracle.ref_available():
        pytest.skip('oracle/_ref not built here (reference sources absent)')
    ref = oracle.RefExtractor()
    rng = np.random.default_rng(7)
    for img in (images['left'][:, :700], rng.integers(0, 256, (240, 320), dtype=np.uint8)):
        for nl, sc, tgt, a, b in ((5, 0.8, 500, 80, 30), (8, 1 / 1.2, 800, 20, 7)):
            lv, sf = oracle.pyramid(img, nl, sc)
            k1, d1, _ = oracle.orb_extract(lv, sf, tgt, a, b)
            k2, d2 = ref(lv, sf, tgt, a, b)
            assert kp_bytes_equal(k1, k2) and np.array_equal(d1, d2)
            # AddPoints with every 5th keypoint as "existing" (test_frame.cpp:36 flow)
            ex = k1[::5]
            budget = oracle.orb_budget(nl, sf[1], tgt)
            k3, d3, _ = oracle.orb_extract(lv, sf, tgt, a, b, exit_kps=ex, budget=budget)
            k4, d4 = ref.add_points(lv, sf, tgt, a, b, ex)
            assert kp_bytes_equal(k3, k4) and np.array_equal(d3, d4)


def test_edge_cases(oracle):
    flat = np.full((120, 160), 77, np.uint8)
    lv, sf = oracle.pyramid(flat, 3, 0.8)
    k, d, npl = oracle.orb_extract(lv, sf, 100, 20, 7)
    assert len(k) == 0 and list(npl) == [0, 0, 0]
    tiny = np.random.default_rng(8).integers(0, 256, (40, 50), dtype=np.uint8)  # no 30-px cell fits
    k, d, _ = oracle.orb_extract([tiny], np.ones(2, np.float32), 100, 20, 7)
    assert len(k) == 0
    assert len(oracle.bf_cross(np.zeros((0, 32), np.uint8), np.zeros((5, 32), np.uint8))) == 0
    assert oracle.three_maxima([0, 5, 0, 4, 1]) == (1, 3, 4)
    assert oracle.three_maxima([0, 50, 0, 4, 1]) == (1, -1, -1)
