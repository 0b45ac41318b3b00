"""GPU parity tests of the pyramidal Lucas-Kanade tracker (Matcher::searchByOPFlow's cv::calcOpticalFlowPyrLK call,
matcher.cpp:745) through the C ABI against the oracle.  The 2x2 sums are exact integers on both sides
and every float step is one IEEE operation in the same order, so points, status and err must be bit-identical (tolerance
0); the oracle itself is pinned to cv2 within the tolerance stated in tests/test_oracle_cv2.py."""
import numpy as np
import pytest
import torch

from .conftest import PARAM_SETS
from .test_oracle_cv2 import lk_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('params', [dict(), dict(max_level=0), dict(max_level=5, win=15), dict(win=9, max_iter=5, eps=0.03), dict(min_eig=1e-2)])
def test_lk_host_entry_vs_oracle(tb, ctx, oracle, images, params):
    for name, prev, nxt, pts in lk_inputs(images):
        en, es, ee = oracle.lk_flow(oracle.LKParams(**params), prev, nxt, pts)
        n, s, e = ctx.lk_flow(tb.LKParams(**params), prev, nxt, pts)
        assert (s == es).all(), (name, params)
        assert n.tobytes() == en.tobytes(), (name, params, np.abs(n - en).max())
        assert e.tobytes() == ee.tobytes(), (name, params)
        if name != 'flat':
            assert s.sum() > len(pts) // 4


def test_lk_close_to_cv2(tb, ctx, images):
    import cv2
    name, prev, nxt, pts = lk_inputs(images)[0]
    n, s, e = ctx.lk_flow(tb.LKParams(), prev, nxt, pts)
    n1, s1, e1 = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
    n1, s1 = n1.reshape(-1, 2), s1.ravel()
    ok = (s == 1) & (s1 == 1)
    d = np.abs(n - n1).max(1)[ok]
    assert (s == s1).mean() >= 0.99 and (d <= 0.05).mean() >= 0.98 and np.median(d) <= 1e-3


def test_lk_device_batched_and_frontend_opflow(tb, ctx, oracle, images):
    """tb_lk_flow_device on stacks + tb_frontend_search_by_opflow (CLAHE'd current frame, IsInFrame filter)."""
    from trackingbench_slam_b200.synth import synth_frame
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    s2 = synth_frame(2, W, H)
    frames = np.stack([images['left'], images['right'], s2, np.roll(s2, 2, axis=1)])
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=4, max_target=tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(4)
        fe.extract(4, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(4)
        cap = fe.kp_cap
        pairs = [(1, 0), (0, 1), (3, 2), (2, 2)]          # (F1 = current, F2 = previous with keys)
        for eq in (0, 1):
            pts = torch.zeros((len(pairs), cap, 2), dtype=torch.float32, device='cuda')
            st = torch.zeros((len(pairs), cap), dtype=torch.uint8, device='cuda')
            er = torch.zeros((len(pairs), cap), dtype=torch.float32, device='cuda')
            fe.search_by_opflow(tb.LKParams(), 4, [q[0] for q in pairs], [q[1] for q in pairs], eq, pts.data_ptr(), st.data_ptr(), er.data_ptr())
            ctx.sync()
            P, S, E = pts.cpu().numpy(), st.cpu().numpy(), er.cpu().numpy()
            tracked = 0
            for i, (f1, f2) in enumerate(pairs):
                n2 = counts[f2]
                img1 = oracle.clahe(frames[f1], 3.0, (8, 8)) if eq else frames[f1]
                kp = np.stack([kps[f2, :n2]['x'], kps[f2, :n2]['y']], 1)
                en, es, ee = oracle.lk_flow(oracle.LKParams(), frames[f2], img1, kp)
                xi, yi = en[:, 0].astype(np.int32), en[:, 1].astype(np.int32)          # matcher.cpp:747-749
                es = es & ((xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)).astype(np.uint8)
                assert (S[i, :n2] == es).all() and (S[i, n2:] == 0).all(), (eq, i)
                assert P[i, :n2].tobytes() == en.tobytes() and E[i, :n2].tobytes() == ee.tobytes(), (eq, i)
                tracked += int(es.sum())
            assert tracked > 500
        # the generic device entry on the same stacks: float2 points per pair
        base, pitch, fstride = fe.level_layout(0)
        n0 = int(counts[0])
        pin = torch.zeros((2, cap, 2), dtype=torch.float32, device='cuda')
        pin[0, :n0] = torch.from_numpy(np.stack([kps[0, :n0]['x'], kps[0, :n0]['y']], 1)).cuda()
        pin[1, :50] = pin[0, :50]
        cnt = torch.tensor([n0, 50], dtype=torch.int32, device='cuda')
        pts = torch.zeros((2, cap, 2), dtype=torch.float32, device='cuda')
        st = torch.zeros((2, cap), dtype=torch.uint8, device='cuda')
        ctx.lk_flow_device(tb.LKParams(), base, pitch, fstride, 4, base, pitch, fstride, 4, W, H, [0, 0], [1, 1], pin.data_ptr(), cnt.data_ptr(), cap,
                           pts.data_ptr(), st.data_ptr())
        ctx.sync()
        en, es, _ = oracle.lk_flow(oracle.LKParams(), frames[0], frames[1], pin[0, :n0].cpu().numpy())
        assert pts[0, :n0].cpu().numpy().tobytes() == en.tobytes() and (st[0, :n0].cpu().numpy() == es).all()
        assert pts[1, :50].cpu().numpy().tobytes() == en[:50].tobytes() and (st[1, 50:].cpu().numpy() == 0).all()
    finally:
        fe.close()


def test_lk_edge_cases(tb, ctx, oracle, images):
    prev, nxt = images['left'], images['right']
    n, s, e = ctx.lk_flow(tb.LKParams(), prev, nxt, np.zeros((0, 2), np.float32))          # no points
    assert n.shape == (0, 2) and len(s) == 0
    with pytest.raises(tb.TbError) as ei:                                                    # window larger than the kernel supports
        ctx.lk_flow(tb.LKParams(win=25), prev, nxt, np.float32([[100, 100]]))
    assert ei.value.code == -4
    pts = np.float32([[1e6, 5], [-1e6, -1e6], [np.nan, 3], [620.5, 188.25]])                # far outside / NaN points
    en, es, ee = oracle.lk_flow(oracle.LKParams(), prev, nxt, pts)
    n, s, e = ctx.lk_flow(tb.LKParams(), prev, nxt, pts)
    assert (s == es).all() and s[3] == 1 and not s[:2].any()
    assert n[3].tobytes() == en[3].tobytes() and n[:2].tobytes() == en[:2].tobytes()


def test_frontend_stereo_depth_by_opflow(tb, ctx, oracle, images):
    """LocalBA::AddMapPointsByStereo's live body (LocalBA.cpp:46-68) without the RANSAC rejection: LK tracks of the left
    keypoints into the CLAHE'd right image, Depth = bf / |x_tracked - x_left|."""
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    frames = np.stack([images['left'], images['right']])
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=2, max_target=tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(2)
        fe.extract(2, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(2)
        cap = fe.kp_cap
        bf = np.float32(0.573 * 718.856)
        for eq in (1, 0):
            z = torch.zeros((1, cap), dtype=torch.float32, device='cuda')
            pts = torch.zeros((1, cap, 2), dtype=torch.float32, device='cuda')
            st = torch.zeros((1, cap), dtype=torch.uint8, device='cuda')
            fe.stereo_depth_by_opflow(tb.LKParams(), bf, 2, [0], [1], eq, z.data_ptr(), pts.data_ptr(), st.data_ptr())
            ctx.sync()
            n = counts[0]
            kp = np.stack([kps[0, :n]['x'], kps[0, :n]['y']], 1)
            img1 = oracle.clahe(frames[1], 3.0, (8, 8)) if eq else frames[1]
            en, es, _ = oracle.lk_flow(oracle.LKParams(), frames[0], img1, kp)
            xi, yi = en[:, 0].astype(np.int32), en[:, 1].astype(np.int32)
            es = es & ((xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)).astype(np.uint8)
            with np.errstate(divide='ignore'):
                ez = np.where(es == 1, bf / np.abs(en[:, 0] - kp[:, 0]), np.float32(-1.0)).astype(np.float32)
            got = z[0].cpu().numpy()
            assert (st[0, :n].cpu().numpy() == es).all() and got[:n].tobytes() == ez.tobytes() and (got[n:] == -1).all()
            good = (es == 1) & (ez > 0) & (ez < 200)
            assert good.sum() > 300
    finally:
        fe.close()


def test_reject_with_f_host_entry(tb, ctx, oracle, images):
    """Matcher::rejectWithF (matcher.cpp:853-881) through tb_reject_with_f_host: the kept tracks are exactly those the oracle
    and the real cv2.findFundamentalMat(FM_RANSAC, 1.0, 0.99) keep (tolerance 0 on the masks), the same number of RANSAC
    iterations is replayed, F agrees with the oracle to 1e-9 relative (CUDA's and glibc's acos / cos / pow differ in the last
    bits; same algorithm otherwise)."""
    from .test_oracle_cv2 import fm_cases, cv2_reject_with_f
    for name, cur, last, st in fm_cases(images, seeds=range(16)):
        es, eF, eit = oracle.reject_with_f(cur, last, st)
        s, F, it = ctx.reject_with_f(cur, last, st)
        assert np.array_equal(s, es), (name, int(s.sum()), int(es.sum()))
        assert np.array_equal(s, cv2_reject_with_f(cur, last, st)), name
        assert it == eit, (name, it, eit)
        if eit:
            assert np.abs(F - eF).max() <= 1e-9 * np.abs(eF).max(), name
    s, F, it = ctx.reject_with_f(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.uint8))   # no tracks
    assert len(s) == 0 and it == 0


def test_reject_with_f_golden(tb, ctx):
    import os
    from .conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'reject_f.npz'))
    for i in range(int(g['n_cases'])):
        s, F, it = ctx.reject_with_f(g['cur%d' % i], g['last%d' % i], g['status_in%d' % i])
        assert np.array_equal(s, g['status_out%d' % i]) and it == int(g['iters%d' % i]), i


def test_reject_with_f_device_batched_and_frontend(tb, ctx, oracle, images):
    """tb_reject_with_f_device on stacked pairs (one CTA per pair) and tb_frontend_reject_with_f behind
    tb_frontend_search_by_opflow = Matcher::searchByOPFlow(F1, F2, pts, equalized, reject = true)."""
    from .test_oracle_cv2 import fm_cases
    cases = fm_cases(images, seeds=range(7))
    cap = max(len(c[1]) for c in cases) + 5
    P = len(cases)
    cur = torch.zeros((P, cap, 2), dtype=torch.float32, device='cuda')
    last = torch.zeros((P, cap, 2), dtype=torch.float32, device='cuda')
    st = torch.ones((P, cap), dtype=torch.uint8, device='cuda')      # bytes past the count must stay untouched
    cnt = torch.tensor([len(c[1]) for c in cases], dtype=torch.int32, device='cuda')
    for i, (name, c, l, s) in enumerate(cases):
        cur[i, :len(c)] = torch.from_numpy(c).cuda(); last[i, :len(c)] = torch.from_numpy(l).cuda(); st[i, :len(c)] = torch.from_numpy(s).cuda()
    F = torch.zeros((P, 9), dtype=torch.float64, device='cuda')
    it = torch.zeros(P, dtype=torch.int32, device='cuda')
    ctx.reject_with_f_device(P, cur.data_ptr(), last.data_ptr(), cnt.data_ptr(), cap, st.data_ptr(), F.data_ptr(), it.data_ptr())
    ctx.sync()
    S, IT = st.cpu().numpy(), it.cpu().numpy()
    for i, (name, c, l, s) in enumerate(cases):
        es, eF, eit = oracle.reject_with_f(c, l, s)
        assert np.array_equal(S[i, :len(c)], es) and (S[i, len(c):] == 1).all() and IT[i] == eit, name
    # behind the frontend's flow entry
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    frames = np.stack([images['left'], images['right']])
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=2, max_target=tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(2)
        fe.extract(2, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(2)
        kc = fe.kp_cap
        pairs = [(1, 0), (0, 1)]
        pts = torch.zeros((2, kc, 2), dtype=torch.float32, device='cuda')
        s2 = torch.zeros((2, kc), dtype=torch.uint8, device='cuda')
        fe.search_by_opflow(tb.LKParams(), 2, [q[0] for q in pairs], [q[1] for q in pairs], 1, pts.data_ptr(), s2.data_ptr())
        ctx.sync()
        before = s2.cpu().numpy().copy()
        fe.reject_with_f([q[1] for q in pairs], pts.data_ptr(), s2.data_ptr())
        ctx.sync()
        after, Pn = s2.cpu().numpy(), pts.cpu().numpy()
        for i, (f1, f2) in enumerate(pairs):
            n2 = counts[f2]
            kp = np.stack([kps[f2, :n2]['x'], kps[f2, :n2]['y']], 1)
            es, _, eit = oracle.reject_with_f(Pn[i, :n2], kp, before[i, :n2])
            assert np.array_equal(after[i, :n2], es) and eit > 0 and es.sum() < before[i, :n2].sum(), i
    finally:
        fe.close()
