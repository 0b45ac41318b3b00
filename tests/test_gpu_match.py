"""GPU parity tests for the Hamming matchers (through the C ABI) vs the oracle and golden vectors.
Integer work: everything must be bit-exact, ties broken by lowest index."""
import os

import numpy as np
import pytest
import torch

from .conftest import GOLDEN, PARAM_SETS

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _rand_desc(rng, n1, n2, planted=True):
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    if planted and n1 > 20 and n2 > 20:
        m = min(n1, n2) // 2
        d2[:m] = d1[n1 - m:]
        d2[:m, 0] ^= rng.integers(0, 4, m, dtype=np.uint8)
        d2[5] = d2[6]
        d1[3] = d1[4]
    return d1, d2


@pytest.mark.parametrize('n1,n2', [(1, 1), (5, 300), (300, 5), (128, 128), (129, 127), (1000, 1003), (2009, 2007)])
def test_bf_cross_and_filter_vs_oracle(tb, ctx, oracle, n1, n2):
    rng = np.random.default_rng(n1 * 7 + n2)
    d1, d2 = _rand_desc(rng, n1, n2)
    t1, t2 = dev(d1), dev(d2)
    got = ctx.match_bf_cross_device(t1.data_ptr(), n1, t2.data_ptr(), n2)
    assert got.tobytes() == oracle.bf_cross(d1, d2).tobytes()
    got = ctx.search_by_bf_device(t1.data_ptr(), n1, t2.data_ptr(), n2, 10.0, 30.0)
    assert got.tobytes() == oracle.search_by_bf(d1, d2, 10.0, 30.0).tobytes()


def test_bf_empty_inputs(tb, ctx):
    t = dev(np.zeros((4, 32), np.uint8))
    assert len(ctx.match_bf_cross_device(t.data_ptr(), 0, t.data_ptr(), 4)) == 0
    assert len(ctx.match_bf_cross_device(t.data_ptr(), 4, t.data_ptr(), 0)) == 0


@pytest.mark.parametrize('pname', list(PARAM_SETS))
def test_bf_golden(tb, ctx, pname):
    gl = np.load(os.path.join(GOLDEN, 'extract_left_%s.npz' % pname))
    gr = np.load(os.path.join(GOLDEN, 'extract_right_%s.npz' % pname))
    gm = np.load(os.path.join(GOLDEN, 'match_%s.npz' % pname))
    t1, t2 = dev(gl['desc']), dev(gr['desc'])
    assert ctx.match_bf_cross_device(t1.data_ptr(), len(gl['desc']), t2.data_ptr(), len(gr['desc'])).tobytes() == gm['cross'].tobytes()   # = cv2.BFMatcher
    assert ctx.search_by_bf_device(t1.data_ptr(), len(gl['desc']), t2.data_ptr(), len(gr['desc']), 10.0, 30.0).tobytes() == gm['filtered'].tobytes()


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('nq,nt', [(1, 1), (7, 1), (100, 257), (513, 1024), (3000, 2500)])
def test_knn2_vs_oracle(tb, ctx, oracle, nq, nt, impl):
    ctx.set_match_impl(impl)
    rng = np.random.default_rng(nq + nt)
    q, t = _rand_desc(rng, nq, nt)
    tq, tt = dev(q), dev(t)
    idx = torch.empty(nq, dtype=torch.int32, device='cuda')
    best = torch.empty_like(idx)
    second = torch.empty_like(idx)
    ctx.knn2_device(tq.data_ptr(), nq, tt.data_ptr(), nt, idx.data_ptr(), best.data_ptr(), second.data_ptr())
    torch.cuda.synchronize()
    ctx.set_match_impl(0)
    io, bo, so = oracle.knn2(q, t)
    so = np.where(so == np.iinfo(np.int32).max, 0x7fffffff, so)
    assert np.array_equal(idx.cpu().numpy(), io) and np.array_equal(best.cpu().numpy(), bo) and np.array_equal(second.cpu().numpy(), so)


def test_knn2_sliced_train_set(tb, ctx, oracle):
    """A query count that leaves the SMs half empty (25 000 rows = 196 CTAs, a rank's share of config 4 on 8 GPUs) makes the
    tensor-core kNN-2 slice the train set and merge the partial results: same (argbest, best, second) as the oracle, with
    duplicated train rows across the slice boundaries so that the lowest-index tie-break is exercised."""
    nq, nt = 25000, 12288 + 37
    rng = np.random.default_rng(3)
    q, t = _rand_desc(rng, nq, nt)
    t[4096:4096 + 50] = t[100:150]          # exact duplicates in a later slice: the earlier row must win
    t[8200:8230] = t[4100:4130]
    q[:300] = t[rng.integers(0, nt, 300)]   # queries with distance 0 somewhere
    ctx.set_match_impl(1)
    tq, tt = dev(q), dev(t)
    idx = torch.empty(nq, dtype=torch.int32, device='cuda')
    best = torch.empty_like(idx)
    second = torch.empty_like(idx)
    ctx.knn2_device(tq.data_ptr(), nq, tt.data_ptr(), nt, idx.data_ptr(), best.data_ptr(), second.data_ptr())
    torch.cuda.synchronize()
    ctx.set_match_impl(0)
    io, bo, so = oracle.knn2(q, t)
    so = np.where(so == np.iinfo(np.int32).max, 0x7fffffff, so)
    assert np.array_equal(idx.cpu().numpy(), io) and np.array_equal(best.cpu().numpy(), bo) and np.array_equal(second.cpu().numpy(), so)


def test_knn2_large_properties(tb, ctx):
    """Config-4 shaped data at 20k x 20k: planted pairs are found, result is invariant to a train permutation."""
    from trackingbench_slam_b200.synth import synth_descriptors
    n = 20000
    Q, T = synth_descriptors(n, seed=42)
    tq, tt = dev(Q), dev(T)
    idx = torch.empty(n, dtype=torch.int32, device='cuda')
    best = torch.empty_like(idx)
    second = torch.empty_like(idx)
    ctx.knn2_device(tq.data_ptr(), n, tt.data_ptr(), n, idx.data_ptr(), best.data_ptr(), second.data_ptr())
    torch.cuda.synchronize()
    i1, b1, s1 = idx.cpu().numpy(), best.cpu().numpy(), second.cpu().numpy()
    d = np.unpackbits(Q ^ T[i1], axis=1).sum(1)
    assert np.array_equal(d, b1) and (b1 <= s1).all()
    assert (b1[:n // 2] < 60).all()                 # planted (Binomial(256,0.08)) matches
    assert (b1[n // 2:] > 60).mean() > 0.99          # unplanted rows sit near the random floor
    perm = np.random.default_rng(0).permutation(n)
    tt2 = dev(T[perm])
    ctx.knn2_device(tq.data_ptr(), n, tt2.data_ptr(), n, idx.data_ptr(), best.data_ptr(), second.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(best.cpu().numpy(), b1) and np.array_equal(second.cpu().numpy(), s1)


def test_hamming_pairs(tb, ctx, oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    out = torch.empty(1000, dtype=torch.int32, device='cuda')
    ta, tb_ = dev(a), dev(b)   # keep the tensors alive: data_ptr() of a temporary is recycled at once
    ctx.hamming_pairs_device(ta.data_ptr(), tb_.data_ptr(), 1000, out.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), np.unpackbits(a ^ b, axis=1).sum(1))
    assert oracle.hamming256(a[0], b[0]) == int(out[0])


def test_frontend_batched_search_by_bf(tb, ctx, oracle, images):
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    L, R = images['left'], images['right']
    sf, lw, lh = tb.pyramid_geometry(L.shape[1], L.shape[0], nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=3, max_target=tgt)
    try:
        fe.upload_frames(0, np.stack([L, R, L]))
        fe.build_pyramid(3)
        fe.extract(3, sf, tgt, a, b)
        cap = fe.kp_cap
        m = torch.zeros((3, cap, 4), dtype=torch.int32, device='cuda')
        c = torch.zeros(3, dtype=torch.int32, device='cuda')
        fe.search_by_bf([0, 1, 0], [1, 2, 2], 10.0, 30.0, m.data_ptr(), cap, c.data_ptr())
        ctx.sync()
        kps, desc, counts = fe.get_results_batch(3)
        cn = c.cpu().numpy()
        mm = m.cpu().numpy().view(np.uint8).reshape(3, cap, 16)
        for p, (qf, tf) in enumerate([(0, 1), (1, 2), (0, 2)]):
            exp = oracle.search_by_bf(desc[qf, :counts[qf]], desc[tf, :counts[tf]], 10.0, 30.0)
            assert cn[p] == len(exp)
            assert mm[p, :cn[p]].tobytes() == exp.tobytes()
        # identical frames: min distance 0, so the reference's strict `<` filter keeps nothing
        assert cn[2] == 0
    finally:
        fe.close()


def _stereo(images, gl, gm):
    H, W = images['left'].shape
    T = np.eye(4, dtype=np.float32)
    T[0, 3] = -0.573
    return H, W, T


@pytest.mark.parametrize('pname', list(PARAM_SETS))
def test_window_and_projection_matchers_golden(tb, ctx, oracle, images, pname):
    nl = PARAM_SETS[pname][0]
    gl = np.load(os.path.join(GOLDEN, 'extract_left_%s.npz' % pname))
    gr = np.load(os.path.join(GOLDEN, 'extract_right_%s.npz' % pname))
    gm = np.load(os.path.join(GOLDEN, 'match_%s.npz' % pname))
    H, W, T = _stereo(images, gl, gm)
    p = tb.MatcherParams(30, 100, 30, 1, 5.0)
    got = ctx.search_by_violence(p, gl['kps'], gl['desc'], gr['kps'], gr['desc'], W, H, 0, nl, 50.0)
    assert got.tobytes() == gm['violence'].tobytes()
    cam = tb.Pinhole(718.856, 718.856, 607.1928, 185.2157, W, H, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
    pp = tb.MatcherParams(30, 50, 30, 1, 30.0)
    got = ctx.search_by_projection(pp, cam, T, gr['kps'], gr['desc'], None, gl['sf'], W, H, gl['kps'], gm['mp_valid'], gm['mp_xyz'], gl['desc'])
    assert got.tobytes() == gm['projection'].tobytes()
    pm = tb.MatcherParams(30, 100, 30, 1, 5.0)
    got = ctx.search_by_projection_map(pm, cam, T, gr['kps'], gr['desc'], None, gl['sf'], W, H, 1 - gm['mp_valid'], gm['mp_xyz'], gm['mp_normal'], gl['desc'], 0.8)
    assert got.tobytes() == gm['projection_map'].tobytes()


def test_window_matchers_variants_vs_oracle(tb, ctx, oracle, images):
    gl = np.load(os.path.join(GOLDEN, 'extract_left_B2000.npz'))
    gr = np.load(os.path.join(GOLDEN, 'extract_right_B2000.npz'))
    gm = np.load(os.path.join(GOLDEN, 'match_B2000.npz'))
    H, W, T = _stereo(images, gl, gm)
    rng = np.random.default_rng(3)
    # no orientation check / different radii, level ranges, thresholds; some F1 features already matched
    for (lo, hi, hist, chk, ratio, r, l0, l1) in ((50, 100, 30, 0, 0.9, 20.0, 0, 8), (40, 80, 30, 1, 0.7, 80.0, 2, 5), (256, 256, 30, 1, 100.0, 5.0, -1, -1)):
        p = tb.MatcherParams(lo, hi, hist, chk, ratio)
        po = oracle.MatcherParams(lo, hi, hist, chk, ratio)
        got = ctx.search_by_violence(p, gl['kps'], gl['desc'], gr['kps'], gr['desc'], W, H, l0, l1, r)
        exp = oracle.search_by_violence(po, gl['kps'], gl['desc'], gr['kps'], gr['desc'], W, H, l0, l1, r)
        assert got.tobytes() == exp.tobytes()
    matched = (rng.random(len(gr['kps'])) < 0.3).astype(np.uint8)
    cam = tb.Pinhole(718.856, 718.856, 607.1928, 185.2157, W, H, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
    camo = oracle.Pinhole(718.856, 718.856, 607.1928, 185.2157, W, H, (oracle.C.c_float * 5)(0, 0, 0, 0, 0))
    R = np.array([[0.9998, -0.01, 0.017, -0.55], [0.0101, 0.99995, -0.002, 0.01], [-0.017, 0.0022, 0.99985, 0.05], [0, 0, 0, 1]], np.float32)
    for chk, nratio in ((1, 15.0), (0, 40.0)):
        p = tb.MatcherParams(30, 70, 30, chk, nratio)
        po = oracle.MatcherParams(30, 70, 30, chk, nratio)
        got = ctx.search_by_projection(p, cam, R, gr['kps'], gr['desc'], matched, gl['sf'], W, H, gl['kps'], gm['mp_valid'], gm['mp_xyz'], gl['desc'])
        exp = oracle.search_by_projection(po, camo, R, gr['kps'], gr['desc'], matched, gl['sf'], W, H, gl['kps'], gm['mp_valid'], gm['mp_xyz'], gl['desc'])
        assert got.tobytes() == exp.tobytes() and len(exp) > 10
        got = ctx.search_by_projection_map(p, cam, R, gr['kps'], gr['desc'], matched, gl['sf'], W, H, 1 - gm['mp_valid'], gm['mp_xyz'], gm['mp_normal'], gl['desc'], 0.9)
        exp = oracle.search_by_projection_map(po, camo, R, gr['kps'], gr['desc'], matched, gl['sf'], W, H, 1 - gm['mp_valid'], gm['mp_xyz'], gm['mp_normal'], gl['desc'], 0.9)
        assert got.tobytes() == exp.tobytes()
    # radial distortion branch of World2Cam
    d = (tb.capi.C.c_float * 5)(-0.28, 0.07, 0.0002, 0.00002, 0.0)
    camd = tb.Pinhole(718.856, 718.856, 607.1928, 185.2157, W, H, d)
    camdo = oracle.Pinhole(718.856, 718.856, 607.1928, 185.2157, W, H, (oracle.C.c_float * 5)(-0.28, 0.07, 0.0002, 0.00002, 0.0))
    p = tb.MatcherParams(30, 70, 30, 1, 25.0)
    po = oracle.MatcherParams(30, 70, 30, 1, 25.0)
    got = ctx.search_by_projection(p, camd, T, gr['kps'], gr['desc'], None, gl['sf'], W, H, gl['kps'], gm['mp_valid'], gm['mp_xyz'], gl['desc'])
    exp = oracle.search_by_projection(po, camdo, T, gr['kps'], gr['desc'], None, gl['sf'], W, H, gl['kps'], gm['mp_valid'], gm['mp_xyz'], gl['desc'])
    assert got.tobytes() == exp.tobytes()


def test_features_in_area_vs_oracle(tb, ctx, oracle):
    gl = np.load(os.path.join(GOLDEN, 'extract_left_B2000.npz'))
    rng = np.random.default_rng(5)
    for (cols, rows) in ((1241, 376), (1920, 1080)):   # 16:9 hits the swapped-scale quirk
        k = gl['kps'].copy()
        k['x'] *= cols / 1241.0
        k['y'] *= rows / 376.0
        for _ in range(25):
            x, y = rng.uniform(-50, cols + 50), rng.uniform(-50, rows + 50)
            r = rng.choice([0.0, 3.0, 30.0, 200.0])
            lo, hi = rng.choice([-1, 0, 2]), rng.choice([-1, 1, 7])
            got = ctx.features_in_area(k, cols, rows, x, y, r, int(lo), int(hi))
            exp = oracle.features_in_area(k, cols, rows, x, y, r, int(lo), int(hi))
            assert np.array_equal(got, exp)


@pytest.mark.parametrize('impl', [0, 1])
def test_batched_bf_popc_and_tcgen05_agree_with_oracle(tb, ctx, oracle, images, impl):
    """K7 (LOP3+POPC) and K8 (tcgen05 int8, hamming = (256 - dot)/2) must give identical bytes."""
    from trackingbench_slam_b200.synth import synth_frame
    nl, sc, tgt, a, b = 5, 0.8, 700, 40.0, 15.0
    frames = np.stack([synth_frame(k, 640, 360) for k in range(4)] + [np.full((360, 640), 90, np.uint8)])   # last frame: no keypoints
    sf, lw, lh = tb.pyramid_geometry(640, 360, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=5, max_target=tgt)
    try:
        ctx.set_match_impl(impl)
        fe.upload_frames(0, frames)
        fe.build_pyramid(5)
        fe.extract(5, sf, tgt, a, b)
        cap = fe.kp_cap
        pairs = [(0, 1), (1, 2), (3, 2), (0, 4), (4, 1), (2, 2)]
        m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda')
        c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
        fe.search_by_bf([p[0] for p in pairs], [p[1] for p in pairs], 3.0, 60.0, m.data_ptr(), cap, c.data_ptr())
        ctx.sync()
        kps, desc, counts = fe.get_results_batch(5)
        assert counts[4] == 0 and counts[0] > 300
        cn = c.cpu().numpy()
        mm = m.cpu().numpy().view(np.uint8).reshape(len(pairs), cap, 16)
        for p, (qf, tf) in enumerate(pairs):
            exp = oracle.search_by_bf(desc[qf, :counts[qf]], desc[tf, :counts[tf]], 3.0, 60.0)
            assert cn[p] == len(exp), (impl, p)
            assert mm[p, :cn[p]].tobytes() == exp.tobytes(), (impl, p)
    finally:
        ctx.set_match_impl(0)
        fe.close()


def test_frontend_batched_projection_device_resident(tb, ctx, oracle, images):
    """tb_frontend_search_by_projection: map points, descriptors and results stay on the device (config 5 path)."""
    from trackingbench_slam_b200.synth import synth_frame
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    frames = np.stack([images['left'], images['right'], synth_frame(1, W, H), synth_frame(2, W, H)])
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=4, max_target=tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(4)
        fe.extract(4, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(4)
        cap = fe.kp_cap
        fx, fy, cx, cy = 718.856, 718.856, 607.1928, 185.2157
        rng = np.random.default_rng(9)
        valid = np.zeros((4, cap), np.uint8)
        xyz = np.zeros((4, cap, 3), np.float32)
        matched = (rng.random((4, cap)) < 0.2).astype(np.uint8)
        for f in range(4):
            n = counts[f]
            depth = rng.uniform(4.0, 40.0, n).astype(np.float32)
            valid[f, :n] = rng.random(n) < 0.9
            xyz[f, :n, 0] = (kps[f, :n]['x'] - np.float32(cx)) / np.float32(fx) * depth
            xyz[f, :n, 1] = (kps[f, :n]['y'] - np.float32(cy)) / np.float32(fy) * depth
            xyz[f, :n, 2] = depth
        pairs = [(1, 0), (0, 1), (3, 2), (2, 3), (0, 0)]
        T = np.tile(np.eye(4, dtype=np.float32), (len(pairs), 1, 1))
        T[0, 0, 3] = -0.573
        T[1, 0, 3] = 0.573
        T[2, 0, 3], T[2, 1, 3] = -0.02, 0.01
        T[3, :3, :3] = np.array([[0.9999, -0.01, 0.0], [0.01, 0.9999, 0.0], [0.0, 0.0, 1.0]], np.float32)
        cam = tb.Pinhole(fx, fy, cx, cy, W, H, (tb.capi.C.c_float * 5)(0, 0, 0, 0, 0))
        camo = oracle.Pinhole(fx, fy, cx, cy, W, H, (oracle.C.c_float * 5)(0, 0, 0, 0, 0))
        for chk, use_matched in ((1, False), (0, True)):
            p = tb.MatcherParams(30, 60, 30, chk, 25.0)
            po = oracle.MatcherParams(30, 60, 30, chk, 25.0)
            d_valid, d_xyz, d_matched = dev(valid), dev(xyz), dev(matched)
            m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda')
            c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
            fe.search_by_projection(p, cam, [q[0] for q in pairs], [q[1] for q in pairs], T,
                                    d_matched.data_ptr() if use_matched else None, d_valid.data_ptr(), d_xyz.data_ptr(), sf,
                                    m.data_ptr(), cap, c.data_ptr())
            ctx.sync()
            cn = c.cpu().numpy()
            mm = m.cpu().numpy().view(np.uint8).reshape(len(pairs), cap, 16)
            total = 0
            for i, (f1, f2) in enumerate(pairs):
                n1, n2 = counts[f1], counts[f2]
                exp = oracle.search_by_projection(po, camo, T[i], kps[f1, :n1], desc[f1, :n1], matched[f1, :n1] if use_matched else None, sf, W, H,
                                                  kps[f2, :n2], valid[f2, :n2], xyz[f2, :n2], desc[f2, :n2])
                assert cn[i] == len(exp), (chk, i, cn[i], len(exp))
                assert mm[i, :cn[i]].tobytes() == exp.tobytes(), (chk, i)
                total += len(exp)
            assert total > 200
    finally:
        fe.close()


def test_frontend_batched_violence_device_resident(tb, ctx, oracle, images):
    """tb_frontend_search_by_violence: keypoints, descriptors, grids and matches stay on the device."""
    from trackingbench_slam_b200.synth import synth_frame
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    frames = np.stack([images['left'], images['right'], synth_frame(1, W, H), synth_frame(2, W, H)])
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=4, max_target=tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(4)
        fe.extract(4, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(4)
        cap = fe.kp_cap
        pairs = [(0, 1), (1, 0), (2, 3), (3, 2), (0, 0), (0, 3)]
        for chk, lo, hi, r, ratio in ((1, 0, nl, 50.0, 5.0), (0, 0, 1, 10.0, 0.8), (1, 1, 3, 30.0, 0.9)):
            p = tb.MatcherParams(30, 100, 30, chk, ratio)
            po = oracle.MatcherParams(30, 100, 30, chk, ratio)
            m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda')
            c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
            fe.search_by_violence(p, [q[0] for q in pairs], [q[1] for q in pairs], lo, hi, r, m.data_ptr(), cap, c.data_ptr())
            ctx.sync()
            cn = c.cpu().numpy()
            mm = m.cpu().numpy().view(np.uint8).reshape(len(pairs), cap, 16)
            total = 0
            for i, (f1, f2) in enumerate(pairs):
                n1, n2 = counts[f1], counts[f2]
                exp = oracle.search_by_violence(po, kps[f1, :n1], desc[f1, :n1], kps[f2, :n2], desc[f2, :n2], W, H, lo, hi, r)
                assert cn[i] == len(exp), (chk, i, cn[i], len(exp))
                assert mm[i, :cn[i]].tobytes() == exp.tobytes(), (chk, i)
                total += len(exp)
            assert total > 50
    finally:
        fe.close()
