#!/usr/bin/env python3
"""Generate tests/golden/bow_*.npz — run in the BUILD container only (needs /root/reference).

Source of truth for the vocabulary transform: oracle/_ref/libtb_ref_bow.so = the reference's vendored
third_part/DBoW2 compiled unmodified (oracle/ref_bow_build.cpp); the vocabulary is a seeded synthetic tree
(trackingbench_slam_b200.vocab.synth_vocabulary) written in the ORBvoc.txt text format and loaded by the
reference's own loadFromTextFile. searchByBow lives only in matcher.cpp (needs Eigen/OpenCV): its lists come from
the C restatement (oracle/tb_oracle.c) and are stored so later edits cannot drift silently.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from trackingbench_slam_b200 import vocab as V  # noqa: E402

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
        out['bow'] = O.search_by_bow(p, kl, dl, nl, kr, dr, nr)
        mv = (np.arange(len(kr)) % 3 != 0).astype(np.uint8)
        out['bow_mp_only'] = O.search_by_bow(p, kl, dl, nl, kr, dr, nr, mv, True)
        p2 = O.MatcherParams(80, 100, 30, 0, 0.95)
        out['bow_nocheck'] = O.search_by_bow(p2, kl, dl, nl, kr, dr, nr)
        np.savez_compressed(os.path.join(G, 'bow_%s_%s.npz' % (vname, pname)), **out)
        print(vname, pname, 'buckets', len(np.unique(nl)), 'bow', len(out['bow']), 'mp_only', len(out['bow_mp_only']),
              'nocheck', len(out['bow_nocheck']), 'stopped', int((nl < 0).sum()))
    ref.close()
