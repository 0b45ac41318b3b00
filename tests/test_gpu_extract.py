"""GPU parity tests (through the C ABI): pyramid, FAST candidates, blur, octree, orientation,
descriptors vs the oracle and the committed golden vectors.  Bit-exact for pixels, coordinates,
octaves, responses and descriptors; angles are compared as float32 bit patterns too (the CUDA
fastAtan2 follows OpenCV's float polynomial without FMA) — see DESIGN.md for the sin/cos note."""
import os

import numpy as np
import pytest

from .conftest import GOLDEN, PARAM_SETS, ROOT, kp_bytes_equal

pytestmark = pytest.mark.gpu


def _frontend(tb, ctx, img_shape, nl, sc, batch, target):
    sf, lw, lh = tb.pyramid_geometry(img_shape[1], img_shape[0], nl, sc)
    return tb.Frontend(ctx, lw, lh, max_batch=batch, max_target=target), sf


def _report_desc_flips(dg, do):
    rows = int((dg != do).any(1).sum())
    return rows, rows / max(len(do), 1)


@pytest.mark.parametrize('pname', list(PARAM_SETS))
@pytest.mark.parametrize('iname', ['left', 'right'])
def test_extract_matches_golden_and_oracle(tb, ctx, oracle, images, iname, pname):
    nl, sc, tgt, a, b = PARAM_SETS[pname]
    img = images[iname]
    g = np.load(os.path.join(GOLDEN, 'extract_%s_%s.npz' % (iname, pname)))
    fe, sf = _frontend(tb, ctx, img.shape, nl, sc, 1, tgt)
    try:
        fe.upload_frames(0, img[None])
        fe.build_pyramid(1)
        lv, sfo = oracle.pyramid(img, nl, sc)
        assert np.array_equal(sf, sfo)
        for l in range(nl):
            assert np.array_equal(fe.download_level(0, l), lv[l]), 'pyramid level %d' % l
        fe.extract(1, sf, tgt, a, b)
        for l in range(nl):
            co, cg = oracle.cell_candidates(lv[l], a, b), fe.get_candidates(0, l)
            assert len(co) == len(cg) and np.array_equal(co['x'], cg['x']) and np.array_equal(co['y'], cg['y']) \
                and np.array_equal(co['response'], cg['response']), 'FAST candidates level %d' % l
            assert np.array_equal(fe.download_level(0, l, blurred=True), oracle.gauss7x7(lv[l])), 'blur level %d' % l
        k, d = fe.get_results(0)
        ko, do, npl = oracle.orb_extract(lv, sf, tgt, a, b)
        assert np.array_equal(fe.get_level_counts(0), npl)
        assert kp_bytes_equal(k, ko)
        assert kp_bytes_equal(k, g['kps'])                      # = the reference's own code (oracle/_ref)
        rows, frac = _report_desc_flips(d, g['desc'])
        assert rows == 0, 'descriptor rows differing: %d (%.4f%%)' % (rows, 100 * frac)
    finally:
        fe.close()


def test_orb_extract_host_entry_point(tb, ctx, oracle, images):
    """tb_orb_extract_host = what the C++ facade ORBExtractor::operator() calls (pyramid given by the caller)."""
    nl, sc, tgt, a, b = PARAM_SETS['A1000']
    lv, sf = oracle.pyramid(images['left'], nl, sc)
    fe = tb.Frontend(ctx, [l.shape[1] for l in lv], [l.shape[0] for l in lv], 1, tgt)
    try:
        k, d = fe.orb_extract_host(lv, sf, tgt, a, b)
        ko, do, _ = oracle.orb_extract(lv, sf, tgt, a, b)
        assert kp_bytes_equal(k, ko) and np.array_equal(d, do)
        # AddPoints (test_frame.cpp:36 flow): every 5th keypoint is "existing"
        ex = ko[::5]
        k2, d2 = fe.orb_extract_host(lv, sf, tgt, a, b, exit_kps=ex)
        ko2, do2, _ = oracle.orb_extract(lv, sf, tgt, a, b, exit_kps=ex, budget=oracle.orb_budget(nl, sf[1], tgt))
        assert len(ko2) < len(ko)
        assert kp_bytes_equal(k2, ko2) and np.array_equal(d2, do2)
    finally:
        fe.close()


def test_batch_slots_and_ragged_batch(tb, ctx, oracle, images):
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    L, R = images['left'], images['right']
    fe, sf = _frontend(tb, ctx, L.shape, nl, sc, 5, tgt)
    try:
        fe.upload_frames(0, np.stack([L, R, L]))          # 3 of 5 slots used
        fe.build_pyramid(3)
        fe.extract(3, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(3)
        ref = {}
        for name, img in (('L', L), ('R', R)):
            lv, _ = oracle.pyramid(img, nl, sc)
            ref[name] = oracle.orb_extract(lv, sf, tgt, a, b)
        for f, name in enumerate('LRL'):
            ko, do, _ = ref[name]
            assert counts[f] == len(ko)
            assert kp_bytes_equal(kps[f, :counts[f]], ko) and np.array_equal(desc[f, :counts[f]], do)
        # idempotence: running again gives the same bytes
        fe.extract(3, sf, tgt, a, b)
        kps2, desc2, counts2 = fe.get_results_batch(3)
        assert np.array_equal(counts, counts2) and kps2.tobytes() == kps.tobytes() and np.array_equal(desc, desc2)
    finally:
        fe.close()


def test_synthetic_1080p_frame(tb, ctx, oracle):
    from trackingbench_slam_b200.synth import synth_frame
    nl, sc, tgt, a, b = PARAM_SETS['B2000']
    img = synth_frame(3, 1920, 1080)
    fe, sf = _frontend(tb, ctx, img.shape, nl, sc, 1, tgt)
    try:
        fe.upload_frames(0, img[None])
        fe.build_pyramid(1)
        fe.extract(1, sf, tgt, a, b)
        k, d = fe.get_results(0)
        lv, _ = oracle.pyramid(img, nl, sc)
        for l in range(nl):
            assert np.array_equal(fe.download_level(0, l), lv[l])
        ko, do, _ = oracle.orb_extract(lv, sf, tgt, a, b)
        assert kp_bytes_equal(k, ko)
        rows, frac = _report_desc_flips(d, do)
        assert frac < 0.001, 'descriptor flips %d' % rows   # north_star bound; observed 0
        assert rows == 0
    finally:
        fe.close()


@pytest.mark.parametrize('shape', [(40, 50), (70, 70), (97, 161), (376, 1241)])
def test_edge_shapes_and_flat_images(tb, ctx, oracle, shape):
    rng = np.random.default_rng(shape[0])
    for img in (rng.integers(0, 256, shape, dtype=np.uint8), np.full(shape, 128, np.uint8)):
        nl, sc, tgt, a, b = 3, 0.8, 300, 20.0, 7.0
        fe, sf = _frontend(tb, ctx, img.shape, nl, sc, 1, tgt)
        try:
            fe.upload_frames(0, img[None])
            fe.build_pyramid(1)
            fe.extract(1, sf, tgt, a, b)
            k, d = fe.get_results(0)
            lv, _ = oracle.pyramid(img, nl, sc)
            ko, do, _ = oracle.orb_extract(lv, sf, tgt, a, b)
            assert kp_bytes_equal(k, ko) and np.array_equal(d, do)
        finally:
            fe.close()


@pytest.mark.parametrize('width', [282, 283, 538, 539, 290, 547])
def test_fast_tile_geometry_edges(tb, ctx, oracle, width):
    """Widths that put the FAST tile at its limits: cells of 32 columns (a tile of eight is the full 256 interior columns of
    the compile-time tile pitch), tile origins that are / are not 16-byte aligned (the tile is staged in 16- or 8-byte units),
    ragged last tiles. Three frames so that the batch path runs too; textured and noisy content."""
    from trackingbench_slam_b200.synth import synth_frame
    h = 150
    rng = np.random.default_rng(width)
    frames = np.stack([synth_frame(width, width, h), rng.integers(0, 256, (h, width), dtype=np.uint8), synth_frame(7, width, h)])
    nl, sc, tgt, a, b = 2, 0.9, 400, 20.0, 7.0
    fe, sf = _frontend(tb, ctx, (h, width), nl, sc, 3, tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(3)
        fe.extract(3, sf, tgt, a, b)
        for f in range(3):
            k, d = fe.get_results(f)
            lv, _ = oracle.pyramid(frames[f], nl, sc)
            ko, do, _ = oracle.orb_extract(lv, sf, tgt, a, b)
            assert kp_bytes_equal(k, ko) and np.array_equal(d, do), (width, f)
    finally:
        fe.close()


def test_pyramid_exact_half_scale(tb, ctx, oracle):
    """Frame(...,4,0.5) (test_projection.cpp:723): OpenCV switches to INTER_AREA on exact 2x."""
    img = np.random.default_rng(11).integers(0, 256, (480, 640), dtype=np.uint8)
    fe, sf = _frontend(tb, ctx, img.shape, 4, 0.5, 1, 100)
    try:
        fe.upload_frames(0, img[None])
        fe.build_pyramid(1)
        lv, _ = oracle.pyramid(img, 4, 0.5)
        for l in range(4):
            assert np.array_equal(fe.download_level(0, l), lv[l])
    finally:
        fe.close()


def test_invalid_arguments_fail_loudly(tb, ctx):
    with pytest.raises(tb.TbError):
        tb.Frontend(ctx, [5000], [100], 1, 100)      # above the 12-bit coordinate packing
    fe = tb.Frontend(ctx, [200, 160], [100, 80], 1, 100)
    try:
        with pytest.raises(tb.TbError):
            fe.extract(1, np.array([1, 0.8], np.float32), 500, 20, 7)   # target > max_target
        with pytest.raises(tb.TbError):
            fe.extract(1, np.array([1, 0.8], np.float32), 50, 20, 7, exit_kps=np.zeros(1, tb.KP_DTYPE))  # AddPoints before operator()
    finally:
        fe.close()


def test_config5_4k_extract_and_projection_match(tb, ctx, oracle):
    """BASELINE config 5 at full size (one stream step): 3840x2160, extract(2000), projection-guided match
    F_k -> F_{k-1} with synthetic depth 10 m and a 1 px/frame pan."""
    from trackingbench_slam_b200.synth import synth_frame
    nl, sc, tgt, a, b = PARAM_SETS['B2000']
    W, H = 3840, 2160
    frames = np.stack([synth_frame(k, W, H, seed0=2000) for k in range(2)])
    fe, sf = _frontend(tb, ctx, (H, W), nl, sc, 2, tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(2)
        fe.extract(2, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(2)
        ref = []
        for f in range(2):
            lv, _ = oracle.pyramid(frames[f], nl, sc)
            ko, do, _ = oracle.orb_extract(lv, sf, tgt, a, b)
            assert kp_bytes_equal(kps[f, :counts[f]], ko)
            assert np.array_equal(desc[f, :counts[f]], do)
            ref.append((ko, do))
        # map points from frame 0's keypoints at 10 m; frame 1's camera moved so that the image pans by ~1 px
        fx = fy = 2000.0
        cx, cy = W / 2.0, H / 2.0
        k0, d0 = ref[0]
        depth = np.full(len(k0), 10.0, np.float32)
        xyz = np.stack([(k0['x'] - np.float32(cx)) / np.float32(fx) * depth, (k0['y'] - np.float32(cy)) / np.float32(fy) * depth, depth], 1).astype(np.float32)
        T = np.eye(4, dtype=np.float32)
        T[0, 3] = -10.0 / fx
        valid = np.ones(len(k0), np.uint8)
        p = tb.MatcherParams(30, 50, 30, 1, 30.0)
        po = oracle.MatcherParams(30, 50, 30, 1, 30.0)
        cam = tb.Pinhole(fx, fy, cx, cy, W, H, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
        camo = oracle.Pinhole(fx, fy, cx, cy, W, H, (oracle.C.c_float * 5)(0, 0, 0, 0, 0))
        k1, d1 = ref[1]
        got = ctx.search_by_projection(p, cam, T, k1, d1, None, sf, W, H, k0, valid, xyz, d0)
        exp = oracle.search_by_projection(po, camo, T, k1, d1, None, sf, W, H, k0, valid, xyz, d0)
        assert got.tobytes() == exp.tobytes()
        assert len(exp) > 50
    finally:
        fe.close()


def test_fast_extractor_vs_oracle(tb, ctx, oracle, images):
    """FASTExtractor (SVO-style grid FAST-10 + Shi-Tomasi). The detector library is absent from the reference
    tree, so parity is pinned to the restatement only; Shi-Tomasi scores are compared bit-exact
    (north_star asks for 1e-4 relative)."""
    rng = np.random.default_rng(21)
    base = images['left'][:360, :640]
    for img, nf in ((np.ascontiguousarray(np.tile(base, (2, 1))[:480, :640]), 300), (rng.integers(0, 256, (480, 640), dtype=np.uint8), 300)):
        lv, sf = oracle.pyramid(img, 3, 0.8)
        inv = (np.float32(1.0) / sf).astype(np.float32)   # FASTextractor.cpp multiplies by the factors it is given
        # keep every level's corners inside the 20x15 grid so that gridPoints.at(k) never throws
        assert oracle.fast_grid(640, 480, nf) == (32, 20, 15)
        inv_used = np.minimum(inv, np.float32(1.0))
        exp = oracle.fast_extract(lv, inv_used, nf, 20.0)
        got = ctx.fast_extract_host(lv, inv_used, nf, 20.0)
        assert exp is not None and len(exp) > 20
        assert kp_bytes_equal(got, exp)
        # AddPoints: cells of every 3rd keypoint are occupied
        ex = exp[::3]
        cell, ncols, nrows = oracle.fast_grid(640, 480, nf)
        occ = np.zeros(ncols * nrows, np.uint8)
        occ[(ex['y'] / np.float32(cell)).astype(np.int32) * ncols + (ex['x'] / np.float32(cell)).astype(np.int32)] = 1
        exp2 = oracle.fast_extract(lv, inv_used, nf, 20.0, occupied=occ)
        got2 = ctx.fast_extract_host(lv, inv_used, nf, 20.0, exit_kps=ex)
        assert kp_bytes_equal(got2, exp2) and len(exp2) < len(exp)
    # the reference throws std::out_of_range on ordinary inputs (bottom-row corners map to k >= nFeatures)
    lv, sf = oracle.pyramid(images['left'], 5, 0.8)
    inv = (np.float32(1.0) / sf).astype(np.float32)
    assert oracle.fast_extract(lv, inv, 500, 20.0) is None
    with pytest.raises(tb.TbError) as ei:
        ctx.fast_extract_host(lv, inv, 500, 20.0)
    assert ei.value.code == tb.capi.TB_E_INVALID


@pytest.mark.parametrize('seed', range(4))
def test_random_shapes_and_thresholds_sweep(tb, ctx, oracle, images, seed):
    """Property sweep: odd sizes (cell widths 30..40, tiles whose origin is not 8-byte aligned, clipped last cells,
    1-3 initial octree nodes), random thresholds / budgets / level counts; textured and noisy content."""
    rng = np.random.default_rng(100 + seed)
    L = images['left']
    for _ in range(6):
        w = int(rng.choice([int(rng.integers(120, 700)), int(rng.integers(353, 363)), int(rng.integers(290, 300))]))
        h = int(rng.integers(100, 380))
        x0, y0 = int(rng.integers(0, L.shape[1] - w)), int(rng.integers(0, L.shape[0] - h + 1)) if h < L.shape[0] else 0
        img = np.ascontiguousarray(L[y0:y0 + h, x0:x0 + w])
        if rng.random() < 0.4:
            img = np.clip(img.astype(np.int16) + rng.integers(-12, 13, img.shape), 0, 255).astype(np.uint8)
        nl = int(rng.integers(1, 5))
        sc = float(rng.choice([0.8, 1 / 1.2, 0.6, 0.9]))
        tgt = int(rng.integers(50, 900))
        a = float(rng.choice([20, 40, 80, 130]))
        b = float(rng.choice([5, 7, 30]))
        if b > a:
            a, b = b, a
        sf, lw, lh = tb.pyramid_geometry(w, h, max(nl, 2), sc)
        sf, lw, lh = sf[:max(nl, 2)], lw[:max(nl, 2)], lh[:max(nl, 2)]
        fe = tb.Frontend(ctx, lw, lh, max_batch=1, max_target=tgt)
        try:
            fe.upload_frames(0, img[None])
            fe.build_pyramid(1)
            fe.extract(1, sf, tgt, a, b)
            k, d = fe.get_results(0)
            lv, sfo = oracle.pyramid(img, max(nl, 2), sc)
            ko, do, _ = oracle.orb_extract(lv, sfo, tgt, a, b)
            assert kp_bytes_equal(k, ko), (w, h, nl, sc, tgt, a, b, len(k), len(ko))
            assert np.array_equal(d, do), (w, h, nl, sc, tgt, a, b)
        finally:
            fe.close()


def test_fast_corner_list_overflow_path_is_exact():
    """fast9_cells keeps a bounded corner list per tile and falls back to scanning its score map when a tile has more
    corners. TB_FAST_LISTCAP=8 forces that path for every tile (read once per process, hence the subprocess)."""
    import subprocess
    import sys
    code = r'''
import os, sys, numpy as np, cv2
sys.path.insert(0, %r)
import trackingbench_slam_b200 as tb
from tests.conftest import PARAM_SETS, GOLDEN, DATA
ctx = tb.Context(0)
for pname in ('A1000', 'B2000'):
    nl, sc, tgt, a, b = PARAM_SETS[pname]
    for iname in ('left', 'right'):
        img = cv2.imread(os.path.join(DATA, iname + '.png'), 0)
        g = np.load(os.path.join(GOLDEN, 'extract_%%s_%%s.npz' %% (iname, pname)))
        sf, lw, lh = tb.pyramid_geometry(img.shape[1], img.shape[0], nl, sc)
        fe = tb.Frontend(ctx, lw, lh, max_batch=1, max_target=tgt)
        fe.upload_level(0, 0, img); fe.build_pyramid(1); fe.extract(1, sf, tgt, a, b)
        k, d = fe.get_results(0)
        assert k.tobytes() == g['kps'].tobytes() and d.tobytes() == g['desc'].tobytes(), (pname, iname)
        fe.close()
print('overflow path ok')
''' % ROOT
    env = dict(os.environ, TB_FAST_LISTCAP='8')
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and 'overflow path ok' in out.stdout, out.stdout + out.stderr
