"""CPU tests of the ORB stereo row-band matcher oracle: the C restatement of the block-commented
body of LocalBA::AddMapPointsByStereo (LocalBA.cpp:69-288) against an independent, literal Python restatement of the same
lines (row table as lists in push order, first strict minimum). Parity is unpinned against the reference itself: that
block is dead code there."""
import math
import os

import numpy as np
import pytest

from .conftest import GOLDEN, PARAM_SETS


def stereo_cases():
    """(name, kL, dL, kR, dR, inv_sf, n_rows) — shared with tests/test_gpu_stereo.py."""
    from oracle import KP_DTYPE
    out = []
    inv_sf = (1.0 / np.float32(1.2) ** np.arange(8, dtype=np.float32)).astype(np.float32)
    for seed, nL, nR, rows, frac in ((1, 300, 280, 120, True), (2, 1500, 1600, 376, True), (3, 40, 900, 64, False), (4, 0, 50, 100, True),
                                     (5, 50, 0, 100, True), (6, 2000, 2000, 1080, True), (7, 513, 511, 33, True)):
        rng = np.random.default_rng(seed)
        def keys(n):
            k = np.zeros(n, KP_DTYPE)
            k['x'] = rng.uniform(0, 1300, n).astype(np.float32)
            y = rng.uniform(0, rows - 1e-3, n)
            k['y'] = (y if frac else np.floor(y)).astype(np.float32)
            k['octave'] = rng.integers(0, 8, n)
            k['size'] = 31
            return k
        kL, kR = keys(nL), keys(nR)
        dL = rng.integers(0, 256, (nL, 32), dtype=np.uint8)
        dR = rng.integers(0, 256, (nR, 32), dtype=np.uint8)
        m = min(nL, nR)
        if m > 10:
            # planted stereo partners: same row band, near-identical descriptor, a few exact copies (distance 0 is
            # rejected by `dist > 0`), duplicated right descriptors (first index must win) and x ties (depth -1)
            sel = rng.permutation(m)[: m * 3 // 4]
            kR['y'][sel] = kL['y'][sel] + rng.uniform(-2.5, 2.5, len(sel)).astype(np.float32)
            kR['octave'][sel] = np.clip(kL['octave'][sel] + rng.integers(-2, 3, len(sel)), 0, 7)
            dR[sel] = dL[sel]
            flips = rng.integers(0, 40, len(sel))
            for j, i in enumerate(sel):
                bits = rng.permutation(256)[: flips[j]]
                for b in bits:
                    dR[i, b >> 3] ^= np.uint8(1 << (b & 7))
            dup = sel[:8]
            kR['y'][dup[1::2]] = kR['y'][dup[0::2]]
            kR['octave'][dup[1::2]] = kR['octave'][dup[0::2]]
            dR[dup[1::2]] = dR[dup[0::2]]
            kR['x'][sel[8:12]] = kL['x'][sel[8:12]]
            kR['y'][sel[12:14]] = np.float32(-0.5)        # row band reaching above the image
            kR['y'][sel[14:16]] = np.float32(rows - 0.25)  # and below it
            kR['x'][sel[16:18]] = np.float32(200000.0)    # outside [min_u, max_u]
        out.append(('s%d' % seed, kL, dL, kR, dR, inv_sf, rows))
    return out


def literal_stereo(p, kL, dL, kR, dR, inv_sf, n_rows):
    """LocalBA.cpp:73-172 line by line (float32 arithmetic where the reference has float)."""
    f32 = np.float32
    rows = [[] for _ in range(n_rows)]
    for iR in range(len(kR)):
        kpY = f32(kR['y'][iR])
        r = f32(p.band) * f32(inv_sf[int(kR['octave'][iR])])
        max_r, min_r = int(math.ceil(f32(kpY + r))), int(math.floor(f32(kpY - r)))
        for yi in range(min_r, max_r + 1):
            if 0 <= yi < n_rows:
                rows[yi].append(iR)
    pop = np.array([bin(i).count('1') for i in range(256)], np.int32)
    matches, depth = [], np.full(len(kL), -1.0, np.float32)
    for iL in range(len(kL)):
        vL = float(kL['y'][iL])
        if not (-1.0 < vL < n_rows):
            continue
        best, bi = 2 ** 31 - 1, 0
        lv = int(kL['octave'][iL])
        for iR in rows[int(vL)]:
            oc = int(kR['octave'][iR])
            if oc < lv - p.octave_tol or oc > lv + p.octave_tol:
                continue
            uR = f32(kR['x'][iR])
            if uR >= f32(p.min_u) and uR <= f32(p.max_u):
                d = int(pop[dL[iL] ^ dR[iR]].sum())
                if d < best and d > 0:
                    best, bi = d, iR
        if best < p.th_dist:
            matches.append((iL, bi, -1, float(best)))
            dx = abs(f32(kL['x'][iL]) - f32(kR['x'][bi]))
            depth[iL] = f32(p.bf) / dx if dx > 0 else f32(-1.0)
    return matches, depth


@pytest.mark.parametrize('case', stereo_cases(), ids=lambda c: c[0])
@pytest.mark.parametrize('th,tol,band', [(30, 1, 2.0), (75, 0, 1.0), (257, 7, 3.0)])
def test_stereo_oracle_vs_literal(case, th, tol, band):
    import oracle
    _, kL, dL, kR, dR, inv_sf, rows = case
    p = oracle.StereoParams(bf=386.1448, band=band, th_dist=th, octave_tol=tol)
    m, z = oracle.stereo_rowband(p, kL, dL, kR, dR, inv_sf, rows)
    em, ez = literal_stereo(p, kL, dL, kR, dR, inv_sf, rows)
    assert [tuple(x) for x in m.tolist()] == em
    assert z.tobytes() == ez.tobytes()
    if len(kL) > 100 and len(kR) > 100:
        assert len(em) > 10 and (ez > 0).sum() > 10
        if th == 257:   # the planted x ties are matched -> depth -1 on a matched key (:168-171)
            assert ((ez == -1) & np.isin(np.arange(len(kL)), [x[0] for x in em])).sum() >= 1


def test_stereo_oracle_on_golden_extract():
    import oracle
    nl, sc = PARAM_SETS['B1000'][:2]
    gl = np.load(os.path.join(GOLDEN, 'extract_left_B1000.npz'))
    gr = np.load(os.path.join(GOLDEN, 'extract_right_B1000.npz'))
    inv_sf = (1.0 / np.cumprod(np.r_[1.0, np.full(nl - 1, sc)]).astype(np.float32)).astype(np.float32)
    p = oracle.StereoParams(bf=386.1448, th_dist=75)
    m, z = oracle.stereo_rowband(p, gl['kps'], gl['desc'], gr['kps'], gr['desc'], inv_sf, 376)
    em, ez = literal_stereo(p, gl['kps'], gl['desc'], gr['kps'], gr['desc'], inv_sf, 376)
    assert [tuple(x) for x in m.tolist()] == em and z.tobytes() == ez.tobytes()
    assert (np.diff(m['queryIdx']) > 0).all()
