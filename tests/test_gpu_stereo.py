"""GPU parity tests of the ORB stereo row-band matcher (LocalBA.cpp:69-288) through the C ABI
against the oracle: match lists and depths must be bit-identical (integer distances, one IEEE division)."""
import os

import numpy as np
import pytest
import torch

from .conftest import GOLDEN, PARAM_SETS
from .test_oracle_stereo import stereo_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', stereo_cases(), ids=lambda c: c[0])
@pytest.mark.parametrize('th,tol,band', [(30, 1, 2.0), (75, 0, 1.0), (257, 7, 3.0)])
def test_stereo_host_entry_vs_oracle(tb, ctx, oracle, case, th, tol, band):
    _, kL, dL, kR, dR, inv_sf, rows = case
    po = oracle.StereoParams(bf=386.1448, band=band, th_dist=th, octave_tol=tol)
    p = tb.StereoParams(bf=386.1448, band=band, th_dist=th, octave_tol=tol)
    em, ez = oracle.stereo_rowband(po, kL, dL, kR, dR, inv_sf, rows)
    m, z = ctx.stereo_match(p, kL, dL, kR, dR, inv_sf, rows)
    assert m.tobytes() == em.tobytes()
    assert z.tobytes() == ez.tobytes()


def test_stereo_capacity_error(tb, ctx):
    _, kL, dL, kR, dR, inv_sf, rows = stereo_cases()[1]
    import ctypes as C
    from trackingbench_slam_b200.capi import lib, MATCH_DTYPE, _vp
    p = tb.StereoParams(bf=1.0, th_dist=257, octave_tol=7)
    out = np.zeros(4, MATCH_DTYPE)
    n = C.c_int()
    rc = lib().tb_stereo_match(ctx.h, C.byref(p), _vp(kL), _vp(dL), len(kL), _vp(kR), _vp(dR), len(kR), _vp(inv_sf), len(inv_sf), rows,
                               _vp(out), 4, C.byref(n), None)
    assert rc == -3 and n.value > 4 and (np.diff(out['queryIdx']) > 0).all()


def test_frontend_stereo_device_resident(tb, ctx, oracle, images):
    """tb_frontend_stereo_match on extracted left/right frames: keypoints, descriptors, matches and depths stay in HBM."""
    from trackingbench_slam_b200.synth import synth_frame
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    frames = np.stack([images['left'], images['right'], synth_frame(1, W, H), synth_frame(2, W, H)])
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    inv_sf = (np.float32(1.0) / sf).astype(np.float32)
    fe = tb.Frontend(ctx, lw, lh, max_batch=4, max_target=tgt)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(4)
        fe.extract(4, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(4)
        cap = fe.kp_cap
        pairs = [(0, 1), (1, 0), (2, 3), (0, 0), (3, 1)]
        for th, tol in ((30, 1), (75, 1), (100, 0)):
            p = tb.StereoParams(bf=386.1448, th_dist=th, octave_tol=tol)
            po = oracle.StereoParams(bf=386.1448, th_dist=th, octave_tol=tol)
            m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda')
            c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
            z = torch.zeros((len(pairs), cap), dtype=torch.float32, device='cuda')
            fe.stereo_match(p, [q[0] for q in pairs], [q[1] for q in pairs], inv_sf, m.data_ptr(), cap, c.data_ptr(), z.data_ptr())
            ctx.sync()
            cn, zz = c.cpu().numpy(), z.cpu().numpy()
            mm = m.cpu().numpy().view(np.uint8).reshape(len(pairs), cap, 16)
            total = 0
            for i, (fl, fr) in enumerate(pairs):
                nL, nR = counts[fl], counts[fr]
                em, ez = oracle.stereo_rowband(po, kps[fl, :nL], desc[fl, :nL], kps[fr, :nR], desc[fr, :nR], inv_sf, H)
                assert cn[i] == len(em), (th, i, cn[i], len(em))
                assert mm[i, :cn[i]].tobytes() == em.tobytes(), (th, i)
                assert zz[i, :nL].tobytes() == ez.tobytes() and (zz[i, nL:] == -1).all()
                total += len(em)
            assert total > (20 if th > 30 else 0)
    finally:
        fe.close()


def test_stereo_golden_extract(tb, ctx, oracle):
    nl, sc = PARAM_SETS['B1000'][:2]
    gl = np.load(os.path.join(GOLDEN, 'extract_left_B1000.npz'))
    gr = np.load(os.path.join(GOLDEN, 'extract_right_B1000.npz'))
    inv_sf = (1.0 / np.cumprod(np.r_[1.0, np.full(nl - 1, sc)]).astype(np.float32)).astype(np.float32)
    em, ez = oracle.stereo_rowband(oracle.StereoParams(bf=386.1448, th_dist=75), gl['kps'], gl['desc'], gr['kps'], gr['desc'], inv_sf, 376)
    m, z = ctx.stereo_match(tb.StereoParams(bf=386.1448, th_dist=75), gl['kps'], gl['desc'], gr['kps'], gr['desc'], inv_sf, 376)
    assert len(em) > 20 and m.tobytes() == em.tobytes() and z.tobytes() == ez.tobytes()
