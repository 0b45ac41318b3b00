"""The C-ABI library loads and exports every symbol include/tb_capi.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from .conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'tb_capi.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol(tb):
    L = tb.capi.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert sorted(tb.capi.EXPORTS) == names


def test_no_cpu_fallback_without_cuda(tb):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    rc = tb.capi.lib().tb_ctx_create(0, None, ctypes.byref(h))
    assert rc == tb.capi.TB_E_CUDA and not h.value
    with pytest.raises(tb.TbError):
        tb.Context(0)


def test_host_geometry_matches_oracle(tb, oracle):
    for (w, h, nl, sc) in ((1241, 376, 8, 1 / 1.2), (1920, 1080, 8, 1 / 1.2), (3840, 2160, 8, 1 / 1.2), (1241, 376, 5, 0.8), (752, 480, 5, 0.6)):
        sf, lw, lh = tb.pyramid_geometry(w, h, nl, sc)
        sfo = oracle.scale_factors(nl, sc)
        assert np.array_equal(sf, sfo)
        assert list(zip(lw.tolist(), lh.tolist())) == oracle.level_sizes(w, h, sfo)


def test_three_maxima_matches_oracle(tb, oracle):
    rng = np.random.default_rng(0)
    for _ in range(200):
        hs = rng.integers(0, 12, 30).astype(np.int32)
        a, b, c = ctypes.c_int(-1), ctypes.c_int(-1), ctypes.c_int(-1)
        tb.capi.lib().tb_three_maxima(hs.ctypes.data_as(ctypes.c_void_p), 30, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        assert (a.value, b.value, c.value) == oracle.three_maxima(hs)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'trackingbench_slam_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'tb_oracle.h' not in txt and 'libtb_oracle' not in txt, f
    for dp, _, fs in os.walk(os.path.join(ROOT, 'include')):
        for f in fs:
            assert 'tb_oracle' not in open(os.path.join(dp, f)).read(), f


def test_init_undistort_rectify_map_host_matches_cv2_and_oracle():
    """tb_init_undistort_rectify_map is host arithmetic (one-time camera setup): checked here without a GPU."""
    import cv2
    import numpy as np
    import oracle as O
    import trackingbench_slam_b200 as tb
    for (w, h, fx, fy, cx, cy, d) in [(1241, 376, 718.856, 718.856, 607.1928, 185.2157, [-0.28, 0.07, 0.0002, 0.00002, 0.0]),
                                      (640, 480, 517.3, 516.5, 318.6, 255.3, [0.2624, -0.9531, -0.0054, 0.0026, 1.1633])]:
        cam = tb.Pinhole(fx, fy, cx, cy, w, h, (tb.capi.C.c_float * 5)(*d))
        m1, m2 = tb.init_undistort_rectify_map(cam)
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        c1, c2 = cv2.initUndistortRectifyMap(K, np.array(d, np.float32).reshape(1, 5), None, K, (w, h), cv2.CV_32F)
        assert m1.tobytes() == c1.tobytes() and m2.tobytes() == c2.tobytes()
        o1, o2 = O.init_undistort_rectify_map(fx, fy, cx, cy, d, w, h)
        assert m1.tobytes() == o1.tobytes() and m2.tobytes() == o2.tobytes()
