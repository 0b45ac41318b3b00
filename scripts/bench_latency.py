#!/usr/bin/env python3
"""BASELINE config 1 (the reference's own CPU-runnable case), measured as single-frame LATENCY through the drop-in host
entries: data/left.png / right.png -> Frame pyramid (5 levels x 0.8) -> ORBExtractor()(…, 1000, 80, 30) -> searchByBF,
host buffers in, host buffers out, one frame at a time (what test/test_matcher.cpp does).  Also a single 1080p frame with
the config-3 parameters.  The reference's own extractor (oracle/_ref) is timed beside it on one host core.  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import trackingbench_slam_b200 as tb
from trackingbench_slam_b200.synth import synth_frame
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = tb.Context(0)
out = {}


def med(fn, n=30):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


for name, (imgs, nl, sc, tgt, a, b) in {
        'kitti_1241x376_5lvl_0.8_1000_80_30': ([cv2.imread(os.path.join(ROOT, 'trackingbench_slam_b200/data', n + '.png'), 0) for n in ('left', 'right')], 5, 0.8, 1000, 80.0, 30.0),
        'synthetic_1920x1080_8lvl_1.2_2000_20_7': ([synth_frame(0, 1920, 1080), synth_frame(1, 1920, 1080)], 8, 1 / 1.2, 2000, 20.0, 7.0)}.items():
    H, W = imgs[0].shape
    sf, lw, lh = tb.pyramid_geometry(W, H, nl, sc)
    fe = tb.Frontend(ctx, lw, lh, max_batch=1, max_target=tgt)
    res = {}

    def frame(img):
        fe.upload_level(0, 0, img)                       # Frame ctor: ComputePyramid on the GPU
        fe.build_pyramid(1)
        fe.extract(1, sf, tgt, a, b)
        return fe.get_results(0)                         # keypoints + descriptors back on the host

    res['gpu_pyramid_extract_ms'] = med(lambda: frame(imgs[0]))
    (k1, d1), (k2, d2) = frame(imgs[0]), frame(imgs[1])
    t1, t2 = None, None
    import torch
    t1, t2 = torch.from_numpy(d1).cuda(), torch.from_numpy(d2).cuda()
    res['gpu_search_by_bf_ms'] = med(lambda: ctx.search_by_bf_device(t1.data_ptr(), len(d1), t2.data_ptr(), len(d2), 10.0, 30.0))
    lv, osf = O.pyramid(imgs[0], nl, sc)
    ref = O.RefExtractor() if O.ref_available() else None
    if ref is not None:
        res['cpu_reference_extract_ms_1core'] = med(lambda: ref(lv, osf, tgt, a, b), 5)
        kr, dr = ref(lv, osf, tgt, a, b)
        res['parity'] = bool(kr.tobytes() == k1.tobytes() and dr.tobytes() == d1.tobytes())
    cv2.setNumThreads(1)
    res['cpu_cv2_pyramid_ms_1core'] = med(lambda: O.pyramid(imgs[0], nl, sc) if False else [cv2.resize(imgs[0], (int(lw[i]), int(lh[i]))) for i in range(1, nl)], 5)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, True)
    res['cpu_cv2_bfmatcher_ms_1core'] = med(lambda: bf.match(d1, d2), 5)
    res['keypoints'] = [len(k1), len(k2)]
    out[name] = res
    fe.close()
print(json.dumps(out))
