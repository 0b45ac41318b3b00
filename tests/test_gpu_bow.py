"""GPU parity for the bag-of-words row (Frame::SetBow's vocabulary transform + Matcher::searchByBow) through the
C ABI: bit-exact against the fixtures made with the reference's own DBoW2 and against the oracle."""
import os

import numpy as np
import pytest
import torch

from trackingbench_slam_b200 import vocab as V

from .conftest import GOLDEN

pytestmark = pytest.mark.gpu
VOCABS = {'k10L4s7': (10, 4, 7, 2), 'k6L5s3': (6, 5, 3, 4)}


def _load(iname, pname):
    e = np.load(os.path.join(GOLDEN, 'extract_%s_%s.npz' % (iname, pname)))
    return e['kps'], e['desc']


@pytest.mark.parametrize('vname', sorted(VOCABS))
def test_bow_transform_and_search_golden(tb, ctx, vname):
    k, L, seed, levelsup = VOCABS[vname]
    voc = V.synth_vocabulary(k, L, seed)
    dv = ctx.vocab_create(voc)
    for pname in ('A1000', 'B2000'):
        g = np.load(os.path.join(GOLDEN, 'bow_%s_%s.npz' % (vname, pname)))
        nodes = {}
        for iname in ('left', 'right'):
            _, d = _load(iname, pname)
            wid, nid, w = ctx.bow_transform(dv, d, levelsup)
            assert np.array_equal(wid, g[iname + '_word'])                 # = reference DBoW2 transform(feature)
            nodes[iname] = V.feature_vector(nid, w)
            assert np.array_equal(nodes[iname], g[iname + '_node'])       # = reference FeatureVector
            ids, vals = V.bow_vector(wid, w, voc.weighting, voc.scoring)
            assert np.array_equal(ids, g[iname + '_bow_ids']) and np.array_equal(vals, g[iname + '_bow_vals'])
        kl, dl = _load('left', pname)
        kr, dr = _load('right', pname)
        p = tb.MatcherParams(50, 100, 30, 1, 0.9)
        assert np.array_equal(ctx.search_by_bow(p, kl, dl, nodes['left'], kr, dr, nodes['right']), g['bow'])
        mv = (np.arange(len(kr)) % 3 != 0).astype(np.uint8)
        assert np.array_equal(ctx.search_by_bow(p, kl, dl, nodes['left'], kr, dr, nodes['right'], mv, True), g['bow_mp_only'])
        p2 = tb.MatcherParams(80, 100, 30, 0, 0.95)
        assert np.array_equal(ctx.search_by_bow(p2, kl, dl, nodes['left'], kr, dr, nodes['right']), g['bow_nocheck'])
    dv.close()


@pytest.mark.parametrize('k,L,seed,levelsup,n', [(10, 3, 1, 4, 1), (10, 3, 1, 1, 777), (4, 6, 2, 4, 5000), (7, 4, 9, 2, 2049), (3, 2, 0, 0, 64)])
def test_bow_transform_vs_oracle(tb, ctx, oracle, k, L, seed, levelsup, n):
    voc = V.synth_vocabulary(k, L, seed, stop_frac=0.1)
    dv = ctx.vocab_create(voc)
    rng = np.random.default_rng(seed + n)
    leaves = np.flatnonzero(voc.is_leaf)
    d = voc.desc[rng.choice(leaves, n)].copy()
    d ^= np.packbits(rng.random((n, 256)) < 0.12, axis=1)
    d[::3] = rng.integers(0, 256, (len(d[::3]), 32), dtype=np.uint8)
    wid_o, nid_o, w_o = oracle.bow_transform(voc, d, levelsup)
    wid, nid, w = ctx.bow_transform(dv, d, levelsup)
    assert np.array_equal(wid, wid_o) and np.array_equal(nid, nid_o) and np.array_equal(w, w_o)
    # device-resident entry
    t = torch.from_numpy(d).cuda()
    tw = torch.zeros(n, dtype=torch.int32, device='cuda'); tn = torch.zeros_like(tw)
    twt = torch.zeros(n, dtype=torch.float64, device='cuda')
    ctx.bow_transform_device(dv, t.data_ptr(), n, levelsup, tw.data_ptr(), tn.data_ptr(), twt.data_ptr())
    ctx.sync()
    assert np.array_equal(tw.cpu().numpy(), wid_o) and np.array_equal(tn.cpu().numpy(), nid_o) and np.array_equal(twt.cpu().numpy(), w_o)
    dv.close()


def test_search_by_bow_random_vs_oracle(tb, ctx, oracle):
    rng = np.random.default_rng(11)
    for n1, n2, nb in [(1, 1, 1), (300, 5, 3), (700, 900, 40), (2100, 1900, 100)]:
        d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        m = min(n1, n2) // 2
        d2[:m] = d1[n1 - m:]
        d2[:m] ^= np.packbits(rng.random((m, 256)) < 0.05, axis=1)
        if n2 > 8:
            d2[5] = d2[6]                       # duplicate candidates: lowest index must win, second = same distance
        node1 = rng.integers(-1, nb, n1).astype(np.int32) * 17
        node2 = rng.integers(-1, nb, n2).astype(np.int32) * 17
        node1[node1 < 0] = -1; node2[node2 < 0] = -1
        if m:
            node2[:m] = node1[n1 - m:]
        k1 = np.zeros(n1, tb.KP_DTYPE); k2 = np.zeros(n2, tb.KP_DTYPE)
        k1['angle'] = rng.uniform(0, 360, n1).astype(np.float32)
        k2['angle'] = rng.uniform(0, 360, n2).astype(np.float32)
        if m:
            k2['angle'][:m] = np.mod(k1['angle'][n1 - m:] + rng.normal(0, 8, m).astype(np.float32), 360).astype(np.float32)
        mv = (rng.random(n2) < 0.7).astype(np.uint8)
        for check in (0, 1):
            for mpo in (False, True):
                p = tb.MatcherParams(70, 100, 30, check, 0.9)
                po = oracle.MatcherParams(70, 100, 30, check, 0.9)
                got = ctx.search_by_bow(p, k1, d1, node1, k2, d2, node2, mv, mpo)
                want = oracle.search_by_bow(po, k1, d1, node1, k2, d2, node2, mv, mpo)
                assert np.array_equal(got, want), (n1, n2, check, mpo)


def test_vocab_create_rejects_bad_trees(tb, ctx):
    voc = V.synth_vocabulary(3, 2, 0)
    bad = voc.parent.copy(); bad[2] = 5            # parent after child
    import ctypes as C
    h = C.c_void_p()
    rc = tb.lib().tb_vocab_create(ctx.h, voc.n_nodes, voc.L, bad.ctypes.data_as(C.c_void_p), voc.is_leaf.ctypes.data_as(C.c_void_p),
                                  voc.desc.ctypes.data_as(C.c_void_p), voc.weight.ctypes.data_as(C.c_void_p), C.byref(h))
    assert rc == tb.TB_E_INVALID


def test_frontend_batched_bow_device_resident(tb, ctx, oracle, images):
    """tb_frontend_set_bow + tb_frontend_search_by_bow: descriptors, FeatureVectors and matches never leave the device."""
    from trackingbench_slam_b200.synth import synth_frame
    from .conftest import PARAM_SETS
    nl, sc, tgt, a, b = PARAM_SETS['B1000']
    W, H = 1241, 376
    frames = np.stack([images['left'], images['right'], synth_frame(1, W, H), synth_frame(2, W, H)])
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=4, max_target=tgt)
    voc = V.synth_vocabulary(10, 4, 7)
    dv = ctx.vocab_create(voc)
    try:
        fe.upload_frames(0, frames)
        fe.build_pyramid(4)
        fe.extract(4, sf, tgt, a, b)
        kps, desc, counts = fe.get_results_batch(4)
        cap = fe.kp_cap
        node = torch.full((4, cap), 12345, dtype=torch.int32, device='cuda')
        word = torch.zeros((4, cap), dtype=torch.int32, device='cuda')
        wt = torch.zeros((4, cap), dtype=torch.float64, device='cuda')
        fe.set_bow(dv, 4, 2, node.data_ptr(), word.data_ptr(), wt.data_ptr())
        ctx.sync()
        node_h, word_h, wt_h = node.cpu().numpy(), word.cpu().numpy(), wt.cpu().numpy()
        fvs = []
        for f in range(4):
            n = counts[f]
            wid, nid, w = oracle.bow_transform(voc, desc[f, :n], 2)
            fv = V.feature_vector(nid, w)
            fvs.append(fv)
            assert np.array_equal(node_h[f, :n], fv) and np.all(node_h[f, n:] == -1)
            assert np.array_equal(word_h[f, :n], wid) and np.array_equal(wt_h[f, :n], w)
        rng = np.random.default_rng(3)
        valid = (rng.random((4, cap)) < 0.8).astype(np.uint8)
        d_valid = torch.from_numpy(valid).cuda()
        pairs = [(0, 1), (1, 0), (2, 3), (3, 2), (0, 0), (0, 2)]
        for chk, mpo in ((1, False), (0, True), (1, True)):
            p = tb.MatcherParams(60, 100, 30, chk, 0.9)
            po = oracle.MatcherParams(60, 100, 30, chk, 0.9)
            m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda')
            c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
            fe.search_by_bow(p, [q[0] for q in pairs], [q[1] for q in pairs], node.data_ptr(), d_valid.data_ptr() if mpo else None, mpo,
                             m.data_ptr(), cap, c.data_ptr())
            ctx.sync()
            cn = c.cpu().numpy()
            mm = m.cpu().numpy().view(np.uint8).reshape(len(pairs), cap, 16)
            total = 0
            for i, (f1, f2) in enumerate(pairs):
                n1, n2 = counts[f1], counts[f2]
                exp = oracle.search_by_bow(po, kps[f1, :n1], desc[f1, :n1], fvs[f1], kps[f2, :n2], desc[f2, :n2], fvs[f2], valid[f2, :n2], mpo)
                assert cn[i] == len(exp), (chk, mpo, i, cn[i], len(exp))
                assert mm[i, :cn[i]].tobytes() == exp.tobytes(), (chk, mpo, i)
                total += len(exp)
            assert total > 100
    finally:
        dv.close()
        fe.close()


def test_bow_and_ingest_edge_cases(tb, ctx, oracle):
    """Empty inputs and invalid arguments of the next-row entry points behave like the rest of the ABI."""
    import ctypes as C
    voc = V.synth_vocabulary(3, 2, 0)
    dv = ctx.vocab_create(voc)
    wid, nid, w = ctx.bow_transform(dv, np.zeros((0, 32), np.uint8), 4)
    assert len(wid) == 0 and len(nid) == 0 and len(w) == 0
    k0, d0 = np.zeros(0, tb.KP_DTYPE), np.zeros((0, 32), np.uint8)
    k3, d3 = np.zeros(3, tb.KP_DTYPE), np.zeros((3, 32), np.uint8)
    p = tb.MatcherParams(50, 100, 30, 1, 0.9)
    assert len(ctx.search_by_bow(p, k0, d0, np.zeros(0, np.int32), k3, d3, np.zeros(3, np.int32))) == 0
    assert len(ctx.search_by_bow(p, k3, d3, np.zeros(3, np.int32), k0, d0, np.zeros(0, np.int32))) == 0
    # all features of F1 unbucketed (-1): nothing to match
    assert len(ctx.search_by_bow(p, k3, d3, np.full(3, -1, np.int32), k3, d3, np.zeros(3, np.int32))) == 0
    # identical descriptors in one bucket: best == second, ratio test rejects (bestDist1 < nRatio * bestDist2 with 0 < 0 false)
    assert len(ctx.search_by_bow(p, k3, d3, np.zeros(3, np.int32), k3, d3, np.zeros(3, np.int32))) == 0
    assert len(oracle.search_by_bow(oracle.MatcherParams(50, 100, 30, 1, 0.9), k3, d3, np.zeros(3, np.int32), k3, d3, np.zeros(3, np.int32))) == 0
    dv.close()
    # ingest: bad arguments
    m = np.zeros((8, 8), np.float32)
    und = tb.Undistorter(ctx, m, m)
    rc = tb.lib().tb_undistort_image_device(ctx.h, und.h, C.c_void_p(0), 8, 8, C.c_size_t(8), C.c_size_t(64), C.c_void_p(0), C.c_size_t(8), C.c_size_t(64), 1)
    assert rc == tb.TB_E_INVALID
    und.close()
    img = np.zeros((4, 4), np.uint8)
    with pytest.raises(tb.TbError):
        ctx.clahe_host(img, 3.0, (8, 8))          # padding to the tile grid would exceed the image: unsupported, not silently wrong
