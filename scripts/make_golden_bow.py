#!/usr/bin/env python3
"""Generate tests/golden/bow_*.npz — run in the BUILD container only (needs /root/reference).

Source of truth for the vocabulary transform: oracle/_ref/libtb_ref_bow.so = the reference's vendored
third_part/DBoW2 compiled unmodified (oracle/ref_bow_build.cpp); the vocabulary is a seeded synthetic tree
(trackingbench_slam_b200.vocab.synth_vocabulary) written in the ORBvoc.txt text format and loaded by the
reference's own loadFromTextFile. The searchByBow lists come from oracle/_ref/libtb_ref_matcher.so = matcher.cpp:619-721
cut out by line range and compiled against the reference's own Frame/MapPoint sources (oracle/ref_matcher_build.cpp);
the C restatement must reproduce every list (asserted here).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from trackingbench_slam_b200 import vocab as V  # noqa: E402
from tests.conftest import PARAM_SETS  # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')
assert O.ref_bow_available(), 'oracle/_ref/libtb_ref_bow.so missing: run make -C oracle'

VOCABS = {'k10L4s7': (10, 4, 7, 2), 'k6L5s3': (6, 5, 3, 4)}   # name -> (k, L, seed, levelsup)
for vname, (k, L, seed, levelsup) in VOCABS.items():
    voc = V.synth_vocabulary(k, L, seed)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'voc.txt')
        voc.save_text(path)
        ref = O.RefVocabulary(path)
    for pname in ('A1000', 'B2000'):
        out = {}
        feats = {}
        for iname in ('left', 'right'):
            e = np.load(os.path.join(G, 'extract_%s_%s.npz' % (iname, pname)))
            nid, wid, ids, vals = ref.set_bow(e['desc'], levelsup)
            out[iname + '_node'] = nid; out[iname + '_word'] = wid
            out[iname + '_bow_ids'] = ids; out[iname + '_bow_vals'] = vals
            feats[iname] = (e['kps'], e['desc'], nid)
        p = O.MatcherParams(50, 100, 30, 1, 0.9)    # TH_LOW 50, nRatio 0.9 (ORB-SLAM style values; the reference's set*Param)
        kl, dl, nl = feats['left']
        kr, dr, nr = feats['right']
        import cv2
        left, right = (cv2.imread(os.path.join(ROOT, 'trackingbench_slam_b200', 'data', n + '.png'), 0) for n in ('left', 'right'))
        nlv, sc = PARAM_SETS[pname][0], np.float32(PARAM_SETS[pname][1])
        cam = O.RefCamera(left.shape[1], left.shape[0], 718.856, 718.856, 607.1928, 185.2157)
        FL, FR = O.RefFrame(left, nlv, sc, cam), O.RefFrame(right, nlv, sc, cam)
        FL.set_keys(kl, dl); FR.set_keys(kr, dr)
        FL.set_feature_vector(nl); FR.set_feature_vector(nr)
        mv = (np.arange(len(kr)) % 3 != 0).astype(np.uint8)
        mps = O.RefMapPoints(FR, mv, np.ones((len(kr), 3), np.float32), attach=True)   # F2->GetMapPoint(idx2) != nullptr
        out['bow'] = O.RefMatcher(50, 100, 30, True, 0.9).search_by_bow(FL, FR, False)
        assert out['bow'].tobytes() == O.search_by_bow(p, kl, dl, nl, kr, dr, nr).tobytes()
        out['bow_mp_only'] = O.RefMatcher(50, 100, 30, True, 0.9).search_by_bow(FL, FR, True)
        assert out['bow_mp_only'].tobytes() == O.search_by_bow(p, kl, dl, nl, kr, dr, nr, mv, True).tobytes()
        p2 = O.MatcherParams(80, 100, 30, 0, 0.95)
        out['bow_nocheck'] = O.RefMatcher(80, 100, 30, False, 0.95).search_by_bow(FL, FR, False)
        assert out['bow_nocheck'].tobytes() == O.search_by_bow(p2, kl, dl, nl, kr, dr, nr).tobytes()
        mps.close(); FL.close(); FR.close()
        np.savez_compressed(os.path.join(G, 'bow_%s_%s.npz' % (vname, pname)), **out)
        print(vname, pname, 'buckets', len(np.unique(nl)), 'bow', len(out['bow']), 'mp_only', len(out['bow_mp_only']),
              'nocheck', len(out['bow_nocheck']), 'stopped', int((nl < 0).sum()))
    ref.close()
