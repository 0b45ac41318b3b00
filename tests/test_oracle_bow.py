"""Bag-of-words row: the oracle's vocabulary transform against the reference's own vendored
DBoW2 (oracle/_ref/libtb_ref_bow.so, build container) and against the committed fixtures (everywhere); searchByBow
against an independent dict-based restatement of matcher.cpp:619-721."""
import os

import numpy as np
import pytest

import oracle as O
from trackingbench_slam_b200 import vocab as V

G = os.path.join(os.path.dirname(__file__), 'golden')
VOCABS = {'k10L4s7': (10, 4, 7, 2), 'k6L5s3': (6, 5, 3, 4)}


def _load(iname, pname):
    e = np.load(os.path.join(G, 'extract_%s_%s.npz' % (iname, pname)))
    return e['kps'], e['desc']


@pytest.mark.parametrize('vname', sorted(VOCABS))
@pytest.mark.parametrize('pname', ['A1000', 'B2000'])
def test_transform_matches_golden(vname, pname):
    k, L, seed, levelsup = VOCABS[vname]
    voc = V.synth_vocabulary(k, L, seed)
    g = np.load(os.path.join(G, 'bow_%s_%s.npz' % (vname, pname)))
    for iname in ('left', 'right'):
        _, d = _load(iname, pname)
        wid, nid, w = O.bow_transform(voc, d, levelsup)
        assert np.array_equal(wid, g[iname + '_word'])
        assert np.array_equal(V.feature_vector(nid, w), g[iname + '_node'])
        ids, vals = V.bow_vector(wid, w, voc.weighting, voc.scoring)
        assert np.array_equal(ids, g[iname + '_bow_ids'])
        assert np.array_equal(vals, g[iname + '_bow_vals'])   # bit-equal doubles: same accumulation order


@pytest.mark.skipif(not (O.ref_bow_available() and os.path.exists('/root/reference')), reason='reference DBoW2 build not present')
def test_transform_matches_reference_dbow2(tmp_path):
    rng = np.random.default_rng(5)
    for (k, L, seed, levelsup) in [(10, 3, 1, 4), (10, 3, 1, 1), (4, 6, 2, 4), (7, 4, 9, 2), (3, 2, 0, 0)]:
        voc = V.synth_vocabulary(k, L, seed, stop_frac=0.1)
        path = str(tmp_path / ('voc_%d_%d_%d.txt' % (k, L, seed)))
        voc.save_text(path)
        back = V.Vocabulary.load_text(path)
        assert np.array_equal(back.parent, voc.parent) and np.array_equal(back.desc, voc.desc)
        assert np.array_equal(back.weight, voc.weight) and np.array_equal(back.is_leaf, voc.is_leaf)
        ref = O.RefVocabulary(path)
        assert ref.size() == int(voc.is_leaf.sum())
        # features: noisy copies of random leaves + pure noise
        leaves = np.flatnonzero(voc.is_leaf)
        d = voc.desc[rng.choice(leaves, 700)].copy()
        noise = (rng.random((700, 256)) < 0.1)
        d ^= np.packbits(noise, axis=1)
        d = np.concatenate([d, rng.integers(0, 256, (300, 32), dtype=np.uint8)])
        nid_r, wid_r, ids_r, vals_r = ref.set_bow(d, levelsup)
        wid, nid, w = O.bow_transform(voc, d, levelsup)
        assert np.array_equal(wid, wid_r)
        assert np.array_equal(V.feature_vector(nid, w), nid_r)
        ids, vals = V.bow_vector(wid, w, voc.weighting, voc.scoring)
        assert np.array_equal(ids, ids_r) and np.array_equal(vals, vals_r)
        ref.close()


def _bow_py(p, k1, d1, n1, k2, d2, n2, mv, mpo):
    """matcher.cpp:619-721 with std::map semantics spelled out in Python."""
    fv1, fv2 = {}, {}
    for i, n in enumerate(n1.tolist()):
        if n >= 0:
            fv1.setdefault(n, []).append(i)
    for i, n in enumerate(n2.tolist()):
        if n >= 0:
            fv2.setdefault(n, []).append(i)
    b1, b2 = np.unpackbits(d1, axis=1).astype(np.int32), np.unpackbits(d2, axis=1).astype(np.int32)
    raw, bins = [], []
    for node in sorted(fv1):
        if node not in fv2:
            continue
        for i1 in fv1[node]:
            best1, best2, bidx = 256, 256, -1
            for i2 in fv2[node]:
                if mpo and not mv[i2]:
                    continue
                dist = int(np.abs(b1[i1] - b2[i2]).sum())
                if dist < best1:
                    best2, best1, bidx = best1, dist, i2
                elif dist < best2:
                    best2 = dist
            if best1 < p.TH_LOW and np.float32(best1) < np.float32(p.nRatio) * np.float32(best2):
                raw.append((i1, bidx, -1, float(best1)))   # cv::DMatch(q, t, dist) leaves imgIdx = -1
                if p.checkOrientation:
                    rot = np.float32(k1['angle'][i1]) - np.float32(k2['angle'][bidx])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360.0))
                    b = int(np.round(np.float32(rot * (np.float32(1.0) / np.float32(p.HISTO_LENGTH)))))   # C round(): half away
                    bins.append(0 if b == p.HISTO_LENGTH else b)
    raw = np.array(raw, dtype=O.MATCH_DTYPE) if raw else np.zeros(0, O.MATCH_DTYPE)
    if not p.checkOrientation:
        return raw
    bins = np.array(bins, np.int64)
    hist = [np.flatnonzero(bins == b) for b in range(p.HISTO_LENGTH)]
    i1, i2, i3 = O.three_maxima([len(h) for h in hist])
    keep = [raw[hist[b]] for b in range(p.HISTO_LENGTH) if b in (i1, i2, i3)]
    return np.concatenate(keep) if keep else raw[:0]


@pytest.mark.parametrize('check', [0, 1])
@pytest.mark.parametrize('mpo', [False, True])
def test_search_by_bow_vs_python(check, mpo):
    voc = V.synth_vocabulary(10, 4, 7)
    k1, d1 = _load('left', 'A1000')
    k2, d2 = _load('right', 'A1000')
    k1, d1, k2, d2 = k1[:400], d1[:400], k2[:450], d2[:450]
    w1, n1, wt1 = O.bow_transform(voc, d1, 2)
    w2, n2, wt2 = O.bow_transform(voc, d2, 2)
    n1, n2 = V.feature_vector(n1, wt1), V.feature_vector(n2, wt2)
    mv = (np.arange(len(k2)) % 4 != 1).astype(np.uint8)
    p = O.MatcherParams(60, 100, 30, check, 0.9)
    got = O.search_by_bow(p, k1, d1, n1, k2, d2, n2, mv, mpo)
    want = _bow_py(p, k1, d1, n1, k2, d2, n2, mv, mpo)
    assert len(got) > 20
    assert np.array_equal(got, want)


@pytest.mark.parametrize('vname', sorted(VOCABS))
def test_search_by_bow_golden(vname):
    for pname in ('A1000', 'B2000'):
        g = np.load(os.path.join(G, 'bow_%s_%s.npz' % (vname, pname)))
        kl, dl = _load('left', pname)
        kr, dr = _load('right', pname)
        p = O.MatcherParams(50, 100, 30, 1, 0.9)
        assert np.array_equal(O.search_by_bow(p, kl, dl, g['left_node'], kr, dr, g['right_node']), g['bow'])
        mv = (np.arange(len(kr)) % 3 != 0).astype(np.uint8)
        assert np.array_equal(O.search_by_bow(p, kl, dl, g['left_node'], kr, dr, g['right_node'], mv, True), g['bow_mp_only'])
        p2 = O.MatcherParams(80, 100, 30, 0, 0.95)
        assert np.array_equal(O.search_by_bow(p2, kl, dl, g['left_node'], kr, dr, g['right_node']), g['bow_nocheck'])


def test_vocabulary_text_roundtrip_and_csr(tmp_path):
    """Runs everywhere (no reference build needed): ORBvoc.txt layout round trip and the flattened tree's invariants."""
    voc = V.synth_vocabulary(4, 3, 5, stop_frac=0.2)
    path = str(tmp_path / 'voc.txt')
    voc.save_text(path)
    assert not open(path).read().endswith('\n')       # a trailing newline would add a bogus root child in the reference loader
    back = V.Vocabulary.load_text(path)
    for name in ('parent', 'is_leaf', 'desc', 'weight'):
        assert np.array_equal(getattr(back, name), getattr(voc, name)), name
    assert (back.k, back.L, back.scoring, back.weighting) == (voc.k, voc.L, voc.scoring, voc.weighting)
    start, children, word = voc.csr()
    n = voc.n_nodes
    assert start[0] == 0 and start[-1] == n - 1 and len(children) == n - 1
    for i in range(n):
        kids = children[start[i]:start[i + 1]]
        assert np.all(voc.parent[kids] == i) and np.all(np.diff(kids) > 0)      # children in ascending id
    leaves = np.flatnonzero(voc.is_leaf)
    assert np.array_equal(word[leaves], np.arange(len(leaves)))                  # word ids in file order
    assert np.all(word[(start[1:] > start[:-1])] == -1)                          # inner nodes carry no word
    fast = V.synth_vocabulary_fast(3, 3, 1)
    assert fast.n_nodes == 1 + 3 + 9 + 27 and int(fast.is_leaf.sum()) == 27
    wid, nid, w = O.bow_transform(fast, fast.desc[-5:], 1)                       # leaves map to themselves
    assert np.array_equal(wid, np.arange(22, 27))
