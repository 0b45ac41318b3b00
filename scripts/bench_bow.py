#!/usr/bin/env python3
"""Bag-of-words row, measured: Frame::SetBow (vocabulary transform, ORBvoc-sized synthetic tree k=10 L=6, levelsup 4)
and Matcher::searchByBow over consecutive frames of a 1080p batch, device-resident, plus the CPU oracle on a sample.
Prints one JSON line (supplementary to bench.py)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200 import vocab as V
from trackingbench_slam_b200.synth import synth_frame

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--k', type=int, default=10)
ap.add_argument('--L', type=int, default=6)
a = ap.parse_args()
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ctx = tb.Context(0, st.cuda_stream)
W, H, nl, sc, tgt = 1920, 1080, 8, 1 / 1.2, 2000
sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
B = a.batch
fe = tb.Frontend(ctx, lw, lh, max_batch=B, max_target=tgt)
cap = fe.kp_cap
fe.upload_frames(0, np.stack([synth_frame(k, W, H) for k in range(B)]))
fe.build_pyramid(B)
fe.extract(B, sf, tgt, 20.0, 7.0)
kps, desc, counts = fe.get_results_batch(B)
t0 = time.time()
voc = V.synth_vocabulary_fast(a.k, a.L, 1)
dv = ctx.vocab_create(voc)
t_voc = time.time() - t0
node = torch.zeros((B, cap), dtype=torch.int32, device='cuda')
word = torch.zeros((B, cap), dtype=torch.int32, device='cuda')
wt = torch.zeros((B, cap), dtype=torch.float64, device='cuda')
pairs = [(k, k - 1) for k in range(1, B)]
f1 = [p[0] for p in pairs]; f2 = [p[1] for p in pairs]
m = torch.zeros((len(pairs), cap, 4), dtype=torch.int32, device='cuda'); c = torch.zeros(len(pairs), dtype=torch.int32, device='cuda')
p = tb.MatcherParams(50, 100, 30, 1, 0.9)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.steps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


ms_set = timed(lambda: fe.set_bow(dv, B, 4, node.data_ptr(), word.data_ptr(), wt.data_ptr()))
ms_search = timed(lambda: fe.search_by_bow(p, f1, f2, node.data_ptr(), None, False, m.data_ptr(), cap, c.data_ptr()))
n_feat = int(counts.sum())
# CPU oracle on one frame pair (same vocabulary), and parity of that pair
import oracle as O
n0, n1 = counts[0], counts[1]
t0 = time.time(); w0, nd0, wt0 = O.bow_transform(voc, desc[0, :n0], 4); w1, nd1, wt1 = O.bow_transform(voc, desc[1, :n1], 4); cpu_set = (time.time() - t0) / 2
fv0, fv1 = V.feature_vector(nd0, wt0), V.feature_vector(nd1, wt1)
po = O.MatcherParams(50, 100, 30, 1, 0.9)
t0 = time.time(); exp = O.search_by_bow(po, kps[1, :n1], desc[1, :n1], fv1, kps[0, :n0], desc[0, :n0], fv0); cpu_search = time.time() - t0
node_h = node.cpu().numpy()
ok = bool(np.array_equal(node_h[0, :n0], fv0) and np.array_equal(node_h[1, :n1], fv1))
cn = c.cpu().numpy(); mm = m.cpu().numpy().view(np.uint8).reshape(len(pairs), cap, 16)
ok = ok and cn[0] == len(exp) and mm[0, :cn[0]].tobytes() == exp.tobytes()
print(json.dumps({'workload': '%d x 1080p frames, %d features, vocabulary k=%d L=%d (%d nodes, %.0f MB), levelsup 4' %
                  (B, n_feat, a.k, a.L, voc.n_nodes, voc.n_nodes * 32 / 1e6),
                  'set_bow_ms_per_batch': ms_set, 'set_bow_Mfeatures_per_s': n_feat / ms_set / 1e3,
                  'search_by_bow_ms_per_batch': ms_search, 'search_by_bow_pairs_per_s': len(pairs) / (ms_search / 1e3),
                  'bow_matches_total': int(cn.sum()), 'cpu_oracle_set_bow_ms_per_frame': cpu_set * 1e3,
                  'cpu_oracle_search_by_bow_ms_per_pair': cpu_search * 1e3, 'parity_pair0': ok,
                  'vocab_build_upload_s': t_voc}))
