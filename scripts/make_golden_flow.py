#!/usr/bin/env python3
"""Generate tests/golden/reject_f.npz: inputs and outputs of Matcher::rejectWithF (matcher.cpp:853-881) with the real
cv2.findFundamentalMat(cur, last, FM_RANSAC, 1.0, 0.99) as the source of truth — synthetic two-view tracks with outliers and
the Lucas-Kanade tracks of data/left.png -> data/right.png.  The C restatement must reproduce every mask (asserted here)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from tests.test_oracle_cv2 import fm_cases, cv2_reject_with_f  # noqa: E402

imgs = {n: cv2.imread(os.path.join(ROOT, 'trackingbench_slam_b200', 'data', n + '.png'), 0) for n in ('left', 'right')}
out = {}
for i, (name, cur, last, st) in enumerate(fm_cases(imgs, seeds=range(6), max_n=700)):
    exp = cv2_reject_with_f(cur, last, st)
    got, F, it = O.reject_with_f(cur, last, st)
    assert np.array_equal(got, exp), name
    out['cur%d' % i] = cur; out['last%d' % i] = last; out['status_in%d' % i] = st; out['status_out%d' % i] = exp
    out['iters%d' % i] = np.int32(it)
    print(name, len(cur), int(st.sum()), '->', int(exp.sum()), 'iterations', it)
out['n_cases'] = np.int32(i + 1)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'reject_f.npz'), **out)
