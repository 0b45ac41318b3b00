"""Pins the restated matcher / search-grid / Frame / FASTExtractor oracle (oracle/tb_oracle.c) to REFERENCE-RUN output:
oracle/_ref/libtb_ref_matcher.so = the reference's own src/types/{Frame,MapPoint,Map}.cpp, src/camera/CameraModel.cpp and
src/extractors/FASTextractor.cpp compiled unmodified plus the hot-path member functions of src/matchers/matcher.cpp cut out
by line range at build time (oracle/ref_matcher_build.cpp, oracle/Makefile).  Build container only (needs /root/reference);
the goldens these runs produce are committed (tests/golden/match_*.npz, scripts/make_golden.py) and checked everywhere by
tests/test_oracle_golden.py.  Stated pins: Eigen's 3-term reduction order, Map iteration order = slot order, FAST-10 detector
(uzh-rpg/fast absent) — see oracle/ref_matcher_build.cpp."""
import os

import numpy as np
import pytest

import oracle as O
from .conftest import GOLDEN, PARAM_SETS

pytestmark = pytest.mark.skipif(not (O.ref_matcher_available() and os.path.exists('/root/reference')),
                                reason='reference matcher build (oracle/_ref/libtb_ref_matcher.so) not present')

FX, FY, CX, CY, BASE = 718.856, 718.856, 607.1928, 185.2157, 0.573   # test_vo.cpp:414


def _ext(iname, pname):
    e = np.load(os.path.join(GOLDEN, 'extract_%s_%s.npz' % (iname, pname)))
    return e['kps'], e['desc']


def _rand_keys(rng, n, w, h, n_levels, clustered=False):
    k = np.zeros(n, O.KP_DTYPE)
    if clustered:   # many keys per grid cell and keys exactly on cell borders
        k['x'] = np.round(rng.uniform(0, w, n) / 7) * 7
        k['y'] = np.round(rng.uniform(0, h, n) / 5) * 5
    else:
        k['x'] = rng.uniform(-3, w + 3, n)
        k['y'] = rng.uniform(-3, h + 3, n)
    k['angle'] = rng.uniform(0, 360, n).astype(np.float32)
    k['angle'][rng.random(n) < 0.05] = 0
    k['octave'] = rng.integers(0, n_levels, n)
    k['size'] = 31
    k['response'] = rng.integers(1, 200, n)
    return k


def _rand_desc(rng, n, base=None, flip=0.06):
    if base is None:
        return rng.integers(0, 256, (n, 32), dtype=np.uint8)
    src = base[rng.integers(0, len(base), n)]
    return src ^ np.packbits(rng.random((n, 256)) < flip, axis=1)


def _frame(img_shape, levels, scale, cam, kps, desc, rng=None):
    h, w = img_shape
    img = (rng.integers(0, 256, (h, w), dtype=np.uint8) if rng is not None else np.zeros((h, w), np.uint8))
    f = O.RefFrame(img, levels, scale, cam)
    f.set_keys(kps, desc)
    f.assign_grid()
    return f


# ------------------------------------------------------------------ a1: scale tables + ComputePyramid (Frame.cpp:13-33, 414-427)
@pytest.mark.parametrize('shape,levels,scale', [((376, 1241), 5, 0.8), ((376, 1241), 8, 1 / 1.2), ((270, 480), 8, 1 / 1.2),
                                                ((135, 240), 4, 0.5), ((97, 131), 6, 0.77), ((64, 64), 3, 0.5)])
def test_frame_constructor_equals_oracle(images, shape, levels, scale):
    rng = np.random.default_rng(shape[0] + levels)
    img = images['left'][:shape[0], :shape[1]] if shape[0] <= 376 and shape[1] <= 1241 else rng.integers(0, 256, shape, dtype=np.uint8)
    cam = O.RefCamera(shape[1], shape[0], FX, FY, CX, CY)
    f = O.RefFrame(img, levels, np.float32(scale), cam)
    g = f.geometry()
    lv, sf = O.pyramid(img, levels, np.float32(scale))
    assert g['sf'].tobytes() == sf.tobytes()
    assert [(int(a), int(b)) for a, b in zip(g['w'], g['h'])] == O.level_sizes(shape[1], shape[0], sf)
    for l in range(levels):
        assert np.array_equal(f.level(l), lv[l]), 'pyramid level %d' % l
    f.close()


# ------------------------------------------------------------------ a12 / a17
def test_descriptor_distance_and_three_maxima():
    rng = np.random.default_rng(0)
    M = O.RefMatcher()
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b = _rand_desc(rng, 500, a, 0.2)
    a[0], b[0] = 0, 255
    b[1] = a[1]
    for i in range(500):
        assert M.descriptor_distance(a[i], b[i]) == O.hamming256(a[i], b[i])
    for t in range(400):
        L = int(rng.integers(1, 40))
        hs = rng.integers(0, [2, 5, 40, 1000][t % 4], L)
        if t % 7 == 0:
            hs[:] = hs[0]
        assert M.three_maxima(hs) == O.three_maxima(hs), hs


# ------------------------------------------------------------------ a16: search grid (Frame.cpp:187-265), incl. the swapped 36/rows, 120/cols
@pytest.mark.parametrize('shape', [(376, 1241), (1080, 1920), (480, 640), (100, 100), (2160, 3840)])
def test_grid_equals_reference(shape):
    h, w = shape
    rng = np.random.default_rng(h)
    cam = O.RefCamera(w, h, FX, FY, CX, CY)
    for clustered in (False, True):
        k = _rand_keys(rng, 1500, w, h, 8, clustered)
        f = O.RefFrame(np.zeros((h, w), np.uint8), 1, 1.0, cam)
        f.set_keys(k, np.zeros((len(k), 32), np.uint8))
        f.assign_grid()
        for i in range(0, len(k), 7):
            ok, px, py = f.pos_in_grid(k[i])
            # Frame.cpp:30-31,259-260: posX = round(x * 36/rows), posY = round(y * 120/cols) (the swapped constants are the reference's)
            ex = int(np.round(np.float32(np.float32(k['x'][i]) * (np.float32(36) / np.float32(h)))))
            ey = int(np.round(np.float32(np.float32(k['y'][i]) * (np.float32(120) / np.float32(w)))))
            assert ok == (0 <= px < 120 and 0 <= py < 36)
            if abs(np.float32(k['x'][i]) * (np.float32(36) / np.float32(h)) % 1 - 0.5) > 1e-3:
                assert px == ex
            if abs(np.float32(k['y'][i]) * (np.float32(120) / np.float32(w)) % 1 - 0.5) > 1e-3:
                assert py == ey
        for q in range(300):
            x, y = rng.uniform(-20, w + 20), rng.uniform(-20, h + 20)
            r = float(rng.choice([0.5, 3, 10, 50, 200, 5000]))
            lo, hi = [(-1, -1), (0, 7), (2, 3), (3, -1), (0, 0), (-1, 2)][q % 6]
            ref = f.features_in_area(x, y, r, lo, hi)
            got = O.features_in_area(k, w, h, x, y, r, lo, hi)
            assert np.array_equal(ref, got), (x, y, r, lo, hi)
        f.close()


# ------------------------------------------------------------------ a11: searchByBF (matcher.cpp:168-228), all-levels branch
@pytest.mark.parametrize('pname', sorted(PARAM_SETS))
def test_search_by_bf_equals_reference(images, pname):
    nl, sc = PARAM_SETS[pname][:2]
    h, w = images['left'].shape
    cam = O.RefCamera(w, h, FX, FY, CX, CY)
    kl, dl = _ext('left', pname)
    kr, dr = _ext('right', pname)
    fl, fr = _frame((h, w), nl, sc, cam, kl, dl), _frame((h, w), nl, sc, cam, kr, dr)
    M = O.RefMatcher()
    for ratio, th in [(10.0, 30.0), (2.0, 64.0), (1.5, 100.0), (0.5, 30.0)]:
        ref = M.search_by_bf(fl, fr, 0, nl, ratio, th)
        got = O.search_by_bf(dl, dr, ratio, th)
        assert ref.tobytes() == got.tobytes()
    fl.close(); fr.close()


# ------------------------------------------------------------------ a13: searchByViolence (matcher.cpp:299-395)
@pytest.mark.parametrize('pname', sorted(PARAM_SETS))
def test_violence_equals_reference_on_kitti(images, pname):
    nl, sc = PARAM_SETS[pname][:2]
    h, w = images['left'].shape
    cam = O.RefCamera(w, h, FX, FY, CX, CY)
    kl, dl = _ext('left', pname)
    kr, dr = _ext('right', pname)
    fl, fr = _frame((h, w), nl, sc, cam, kl, dl), _frame((h, w), nl, sc, cam, kr, dr)
    for (low, high, histo, check, ratio), (lo, hi, r) in [((30, 100, 30, True, 5.0), (0, nl, 50.0)), ((50, 100, 30, True, 0.8), (0, 1, 10.0)),
                                                          ((80, 100, 30, False, 0.9), (-1, -1, 30.0)), ((64, 100, 20, True, 0.7), (1, 3, 80.0)),
                                                          ((256, 100, 45, True, 1.0), (0, nl, 25.0))]:
        M = O.RefMatcher(low, high, histo, check, ratio)
        ref = M.search_by_violence(fl, fr, lo, hi, r)
        got = O.search_by_violence(O.MatcherParams(low, high, histo, int(check), ratio), kl, dl, kr, dr, w, h, lo, hi, r)
        assert len(ref) > 0 or low < 40
        assert ref.tobytes() == got.tobytes(), (low, high, histo, check, ratio, lo, hi, r)
    fl.close(); fr.close()


def test_violence_equals_reference_random():
    rng = np.random.default_rng(11)
    for case in range(12):
        h, w = [(376, 1241), (1080, 1920), (480, 640)][case % 3]
        cam = O.RefCamera(w, h, FX, FY, CX, CY)
        n1, n2 = int(rng.integers(0, 900)), int(rng.integers(1, 900))
        k2 = _rand_keys(rng, n2, w, h, 8, clustered=case % 2 == 1)
        d2 = _rand_desc(rng, n2)
        k1 = _rand_keys(rng, n1, w, h, 8)
        if n1:
            src = rng.integers(0, n2, n1)
            k1['x'] = k2['x'][src] + rng.normal(0, 4, n1)
            k1['y'] = k2['y'][src] + rng.normal(0, 4, n1)
            k1['angle'] = np.mod(k2['angle'][src] + rng.normal(3, 6, n1), 360).astype(np.float32)
            d1 = d2[src] ^ np.packbits(rng.random((n1, 256)) < 0.07, axis=1)
        else:
            d1 = np.zeros((0, 32), np.uint8)
        f1, f2 = _frame((h, w), 1, 1.0, cam, k1, d1), _frame((h, w), 1, 1.0, cam, k2, d2)
        low, high = int(rng.integers(20, 120)), 100
        histo = int(rng.integers(19, 60))
        check = bool(case % 4 != 3)
        ratio = float(rng.choice([0.6, 0.8, 1.0, 5.0]))
        lo, hi = [(0, 8), (-1, -1), (0, 1), (2, 5)][case % 4]
        r = float(rng.choice([5.0, 12.0, 40.0]))
        ref = O.RefMatcher(low, high, histo, check, ratio).search_by_violence(f1, f2, lo, hi, r)
        got = O.search_by_violence(O.MatcherParams(low, high, histo, int(check), ratio), k1, d1, k2, d2, w, h, lo, hi, r)
        assert ref.tobytes() == got.tobytes(), case
        f1.close(); f2.close()


# ------------------------------------------------------------------ a14 / a15: searchByProjection x2 + IsInFrustum
def _stereo_case(images, pname):
    """BASELINE config 2: map points from the LEFT keypoints with depth = b*fx/disparity (test_vo.cpp:496-498)."""
    h, w = images['left'].shape
    kl, dl = _ext('left', pname)
    kr, dr = _ext('right', pname)
    disp = images['disparity']
    u = np.clip(kl['x'].astype(np.int32), 0, w - 1)
    v = np.clip(kl['y'].astype(np.int32), 0, h - 1)
    dsp = disp[v, u].astype(np.float32)
    valid = (dsp > 0).astype(np.uint8)
    depth = np.where(dsp > 0, np.float32(BASE * FX) / np.maximum(dsp, 1), 0).astype(np.float32)
    xyz = np.stack([(kl['x'] - np.float32(CX)) / np.float32(FX) * depth, (kl['y'] - np.float32(CY)) / np.float32(FY) * depth, depth], 1).astype(np.float32)
    return h, w, kl, dl, kr, dr, valid, xyz


def _pose(rng, tx=-BASE, rot=0.0):
    T = np.eye(4, dtype=np.float32)
    if rot:
        a = rng.normal(0, rot, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        T[:3, :3] = (Rx @ Ry @ Rz).astype(np.float32)
    T[:3, 3] = np.array([tx, rng.normal(0, 0.05) if rot else 0, rng.normal(0, 0.05) if rot else 0], np.float32)
    return T


@pytest.mark.parametrize('pname', sorted(PARAM_SETS))
def test_projection_equals_reference_on_kitti(images, pname):
    nl, sc = PARAM_SETS[pname][:2]
    h, w, kl, dl, kr, dr, valid, xyz = _stereo_case(images, pname)
    rng = np.random.default_rng(len(kl))
    for case, d in enumerate([(0, 0, 0, 0, 0), (-0.3, 0.1, 0.001, 0.0005, 0.02)]):
        cam = O.RefCamera(w, h, FX, FY, CX, CY, d)
        pin = O.Pinhole(FX, FY, CX, CY, w, h, (O.C.c_float * 5)(*d))
        sf = O.scale_factors(nl, np.float32(sc))
        for it in range(3):
            fl, fr = _frame((h, w), nl, np.float32(sc), cam, kl, dl), _frame((h, w), nl, np.float32(sc), cam, kr, dr)
            fl.set_pose(np.eye(4, dtype=np.float32))
            T = _pose(rng, rot=0.0 if it == 0 else 0.01)
            fr.set_pose(T)
            mps = O.RefMapPoints(fl, valid, xyz, attach=True)
            bad = np.zeros(len(kl), np.uint8)
            matched = np.zeros(len(kr), np.uint8)
            if it == 2:   # bad flags on F2's map points, F1 keys that already carry an observed map point
                for i in np.flatnonzero(valid)[::9]:
                    mps.set_bad(i); bad[i] = 1
                # every 5th F1 key already carries a map point; it blocks the key only if that point has observations
                src = np.flatnonzero(valid)
                obs = {}
                for j in range(0, len(kr), 5):
                    i = int(src[(j // 5) % len(src)])
                    fr.attach_map_point(j, mps, i)
                    obs[i] = 1 if (j // 5) % 2 == 0 else 0
                for i, o in obs.items():
                    mps.set_obs(i, o)
                for j in range(0, len(kr), 5):
                    matched[j] = obs[int(src[(j // 5) % len(src)])]
            for (low, high, histo, check, ratio) in [(30, 50, 30, True, 30.0), (30, 100, 30, False, 15.0), (30, 64, 24, True, 8.0)]:
                M = O.RefMatcher(low, high, histo, check, ratio)
                ref = M.search_by_projection(fr, fl)
                mpv = (valid.astype(bool) & ~bad.astype(bool)).astype(np.uint8)
                got = O.search_by_projection(O.MatcherParams(low, high, histo, int(check), ratio), pin, T, kr, dr,
                                             matched if it == 2 else None, sf, w, h, kl, mpv, xyz, dl)
                assert ref.tobytes() == got.tobytes(), (pname, case, it, low, high, histo, check, ratio)
                if it == 0 and case == 0 and high == 50:
                    assert len(ref) > 50
            mps.close(); fl.close(); fr.close()


@pytest.mark.parametrize('pname', ['A1000', 'B2000'])
def test_projection_map_and_frustum_equal_reference(images, pname):
    nl, sc = PARAM_SETS[pname][:2]
    h, w, kl, dl, kr, dr, valid, xyz = _stereo_case(images, pname)
    rng = np.random.default_rng(3 + len(kl))
    sf = O.scale_factors(nl, np.float32(sc))
    for case, d in enumerate([(0, 0, 0, 0, 0), (-0.3, 0.1, 0.001, 0.0005, 0.02)]):
        cam = O.RefCamera(w, h, FX, FY, CX, CY, d)
        pin = O.Pinhole(FX, FY, CX, CY, w, h, (O.C.c_float * 5)(*d))
        for it in range(3):
            fl, fr = _frame((h, w), nl, np.float32(sc), cam, kl, dl), _frame((h, w), nl, np.float32(sc), cam, kr, dr)
            Tl = _pose(rng, tx=0.0, rot=0.0 if it == 0 else 0.02)
            fl.set_pose(Tl)          # the map points take their normal from this frame's camera centre (MapPoint.cpp:24-26)
            T = _pose(rng, rot=0.0 if it == 0 else 0.03)
            fr.set_pose(T)
            mps = O.RefMapPoints(fl, valid, xyz, attach=False)
            idx = np.flatnonzero(valid)
            bad = np.zeros(len(idx), np.uint8)
            if it == 2:
                for j, i in enumerate(idx):
                    if j % 6 == 0:
                        mps.set_bad(i); bad[j] = 1
            got_xyz = np.zeros((len(idx), 3), np.float32)
            nrm = np.zeros((len(idx), 3), np.float32)
            mpd = np.zeros((len(idx), 32), np.uint8)
            for j, i in enumerate(idx):
                got_xyz[j], nrm[j], mpd[j] = mps.get(i)
            assert np.array_equal(got_xyz, xyz[idx]) and np.array_equal(mpd, dl[idx])
            # IsInFrustum of every point, through the map matcher with TH_HIGH = 256 and a huge window: a point is listed iff
            # it passed the frustum and had a candidate; plus the direct call for the projected coordinates
            for j in range(0, len(idx), 17):
                ok, lvl, x, y, vc = fr.is_in_frustum(mps, idx[j])
                if ok:
                    assert lvl == 0 and 0 <= int(x) < w and 0 <= int(y) < h and vc >= 0.5
            for (low, high, histo, check, ratio), radio in [((30, 100, 30, True, 5.0), 0.8), ((30, 60, 30, True, 1.0), 0.6), ((30, 256, 30, False, 3.0), 2.0)]:
                M = O.RefMatcher(low, high, histo, check, ratio)
                ref = M.search_by_projection_map(mps, fr, radio)
                got = O.search_by_projection_map(O.MatcherParams(low, high, histo, int(check), ratio), pin, T, kr, dr, None, sf, w, h,
                                                 bad, got_xyz, nrm, mpd, radio)
                assert ref.tobytes() == got.tobytes(), (pname, case, it, low, high, radio)
                if it == 0 and case == 0 and high == 100:
                    assert len(ref) > 20
            mps.close(); fl.close(); fr.close()


# ------------------------------------------------------------------ f2: searchByBow (matcher.cpp:619-721)
@pytest.mark.parametrize('pname', ['A1000', 'B2000'])
def test_search_by_bow_equals_reference(images, pname, tmp_path):
    from trackingbench_slam_b200 import vocab as V
    nl, sc = PARAM_SETS[pname][:2]
    h, w = images['left'].shape
    cam = O.RefCamera(w, h, FX, FY, CX, CY)
    kl, dl = _ext('left', pname)
    kr, dr = _ext('right', pname)
    L = O.ref_matcher_lib()
    for (k, depth, seed, levelsup) in [(10, 4, 7, 2), (6, 5, 3, 4)]:
        voc = V.synth_vocabulary(k, depth, seed)
        path = str(tmp_path / ('voc_%d_%d.txt' % (k, depth)))
        voc.save_text(path)
        hv = O.C.c_void_p(L.tbrefm_vocab_load_text(path.encode()))
        assert hv
        fl, fr = _frame((h, w), nl, np.float32(sc), cam, kl, dl), _frame((h, w), nl, np.float32(sc), cam, kr, dr)
        # Frame::SetBow hard-codes levelsup = 4 (Frame.cpp:269); other depths go through set_feature_vector
        _, n1, w1 = O.bow_transform(voc, dl, levelsup)
        _, n2, w2 = O.bow_transform(voc, dr, levelsup)
        node1, node2 = V.feature_vector(n1, w1), V.feature_vector(n2, w2)
        if levelsup == 4:
            fl.set_bow(hv); fr.set_bow(hv)
            assert np.array_equal(fl.feature_vector(), node1) and np.array_equal(fr.feature_vector(), node2)
        else:
            fl.set_feature_vector(node1); fr.set_feature_vector(node2)
        valid = np.ones(len(kr), np.uint8)
        valid[::3] = 0
        mps = O.RefMapPoints(fr, valid, np.ones((len(kr), 3), np.float32), attach=True)
        for (low, high, histo, check, ratio), mp_only in [((50, 100, 30, True, 0.8), False), ((80, 100, 30, False, 0.95), False),
                                                         ((64, 100, 25, True, 0.7), True), ((30, 100, 30, True, 5.0), True)]:
            ref = O.RefMatcher(low, high, histo, check, ratio).search_by_bow(fl, fr, mp_only)
            got = O.search_by_bow(O.MatcherParams(low, high, histo, int(check), ratio), kl, dl, node1, kr, dr, node2, valid, mp_only)
            assert ref.tobytes() == got.tobytes(), (pname, k, depth, low, mp_only)
        mps.close(); fl.close(); fr.close()
        L.tbrefm_vocab_free(hv)


# ------------------------------------------------------------------ a10: FASTExtractor (FASTextractor.cpp, whole file)
def test_fast_extractor_equals_reference(images):
    rng = np.random.default_rng(2)
    left = images['left']
    for img, nl, sc in [(left, 3, 0.5), (left[:200, :300], 1, 1.0), (images['right'], 4, 0.8)]:
        lv, sf = O.pyramid(img, nl, np.float32(sc))
        inv = (np.float32(1) / sf).astype(np.float32)
        for nf, th in [(100, 20.0), (300, 5.0), (1000, 40.0), (37, 0.0)]:
            cell, ncols, nrows = O.fast_grid(img.shape[1], img.shape[0], nf)

            def in_grid(kps):
                # FASTextractor.cpp:58 reads grid_occupancy_[k] unchecked: corners of the strip below the last full cell row
                # (k >= ncols*nrows) read past the vector<bool> — heap garbage, not reproducible run to run.  The oracle and
                # the CUDA path define those cells as free; the comparison with the reference build covers the cells inside.
                kk = (kps['y'] / np.float32(cell)).astype(np.int32) * ncols + (kps['x'] / np.float32(cell)).astype(np.int32)
                return kps[kk < ncols * nrows]

            ref = O.ref_fast_extract(lv, inv, nf, th)
            got = O.fast_extract(lv, inv, nf, th)
            assert (ref is None) == (got is None), (nf, th)
            if ref is not None:
                assert in_grid(ref).tobytes() == in_grid(got).tobytes(), (nf, th)
                assert len(in_grid(ref)) > 10
                # AddPoints: cells of the exit points are skipped
                ex = in_grid(ref)[::3]
                occ = np.zeros(ncols * nrows, np.uint8)
                for e in ex:
                    occ[int(np.float32(e['y']) / np.float32(cell)) * ncols + int(np.float32(e['x']) / np.float32(cell))] = 1
                ref2 = O.ref_fast_extract(lv, inv, nf, th, ex)
                got2 = O.fast_extract(lv, inv, nf, th, occ)
                assert (ref2 is None) == (got2 is None)
                if ref2 is not None:
                    assert in_grid(ref2).tobytes() == in_grid(got2).tobytes()
                    assert len(in_grid(ref2)) < len(in_grid(ref))
    for _ in range(300):
        u, v = int(rng.integers(0, left.shape[1])), int(rng.integers(0, left.shape[0]))
        a, b = O.ref_shi_tomasi(left, u, v), O.shi_tomasi(left, u, v)
        assert np.float32(a).tobytes() == np.float32(b).tobytes()


# ------------------------------------------------------------------ f3 / f4 glue: searchByOPFlow's IsInFrame filter, camera and CLAHE call sites
def test_opflow_and_ingest_glue_equal_reference(images):
    left, right = images['left'], images['right']
    h, w = left.shape
    kl, dl = _ext('left', 'A1000')
    for equalized in (False, True):
        cam = O.RefCamera(w, h, FX, FY, CX, CY)
        f1, f2 = _frame((h, w), 1, 1.0, cam, kl[:0], dl[:0]), _frame((h, w), 1, 1.0, cam, kl, dl)
        f1.close(); f2.close()
        f1 = O.RefFrame(right, 1, 1.0, cam)
        f2 = O.RefFrame(left, 1, 1.0, cam)
        f2.set_keys(kl, dl)
        ref_m, ref_pts = O.RefMatcher().search_by_opflow(f1, f2, equalized)
        img1 = O.clahe(right) if equalized else right
        if equalized:
            assert np.array_equal(f1.equalize(), img1)
        pts, st, _ = O.lk_flow(O.LKParams(), left, img1, np.stack([kl['x'], kl['y']], 1))
        xi, yi = pts[:, 0].astype(np.int32), pts[:, 1].astype(np.int32)
        keep = (st != 0) & (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        assert np.array_equal(ref_pts, pts)
        assert np.array_equal(ref_m['queryIdx'], np.flatnonzero(keep)) and np.array_equal(ref_m['trainIdx'], np.flatnonzero(keep))
        assert keep.sum() > 300
        f1.close(); f2.close()
    # PinholeCamera: UndistortPoint / UndistortImage / World2Cam call sites (CameraModel.cpp:63-121)
    d = (-0.28, 0.07, 0.0002, 0.00002, 0.0)
    cam = O.RefCamera(w, h, FX, FY, CX, CY, d)
    m1, m2 = O.init_undistort_rectify_map(FX, FY, CX, CY, d, w, h)
    assert np.array_equal(cam.undistort_image(left), O.remap_linear(left, m1, m2))
    pts = np.stack([kl['x'], kl['y']], 1)[:200]
    exp = O.undistort_points(FX, FY, CX, CY, d, pts)
    for i in range(len(pts)):
        assert np.array_equal(cam.undistort_point(pts[i]), exp[i])
    f = O.RefFrame(left, 1, 1.0, cam)
    f.set_keys(kl[:200], dl[:200], undistort=True)
    assert np.array_equal(f.undistort_points(), exp)
    f.close()


# ------------------------------------------------------------------ f3: ORB stereo row-band matcher (LocalBA.cpp:69-288, dead code in the
# reference: its text is compiled by line range as a function body — the strongest pin a block that the reference never builds can have)
@pytest.mark.parametrize('pname', sorted(PARAM_SETS))
def test_stereo_rowband_equals_reference_text(images, pname):
    nl, sc = PARAM_SETS[pname][:2]
    h, w = images['left'].shape
    cam = O.RefCamera(w, h, FX, FY, CX, CY)
    kl, dl = _ext('left', pname)
    kr, dr = _ext('right', pname)
    fl, fr = O.RefFrame(images['left'], nl, np.float32(sc), cam), O.RefFrame(images['right'], nl, np.float32(sc), cam)
    inv_sf = fl.geometry()['inv_sf']

    def inside(k):
        # LocalBA.cpp:89-90 writes vRowIndices[yi] unchecked: a key whose band [floor(y-r), ceil(y+r)] leaves the image is an
        # out-of-bounds write in the reference (it crashes on the 8-level sets, where the `pt *= sf` keys sit near row 0).
        # The oracle and the CUDA path drop those rows; the comparison with the reference's text uses the keys it can take.
        r = np.float32(2) * inv_sf[k['octave']]
        return (np.floor(k['y'] - r) >= 0) & (np.ceil(k['y'] + r) <= h - 1)

    ml, mr = inside(kl), inside(kr)
    kl, dl, kr, dr = kl[ml], dl[ml], kr[mr], dr[mr]
    fl.set_keys(kl, dl); fr.set_keys(kr, dr)
    for (cur, ste, kc, dc, ks, ds) in [(fl, fr, kl, dl, kr, dr), (fr, fl, kr, dr, kl, dl)]:
        ref_m, ref_d = O.ref_stereo_rowband(cur, ste, BASE * FX, FX)
        got_m, got_d = O.stereo_rowband(O.StereoParams(bf=BASE * FX), kc, dc, ks, ds, inv_sf, h)
        assert len(ref_m) > 30
        assert ref_m.tobytes() == got_m.tobytes()
        assert ref_d.tobytes() == got_d.tobytes()
    fl.close(); fr.close()
