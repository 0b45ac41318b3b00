"""ctypes bindings for the CPU oracle (oracle/libtb_oracle.so, oracle/_ref/libtb_ref.so).

TEST INFRASTRUCTURE ONLY — see oracle/tb_oracle.h.  May be imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.  The product package
(trackingbench_slam_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KP_DTYPE = np.dtype([('x', '<f4'), ('y', '<f4'), ('size', '<f4'), ('angle', '<f4'),
                     ('response', '<f4'), ('octave', '<i4'), ('class_id', '<i4')])
MATCH_DTYPE = np.dtype([('queryIdx', '<i4'), ('trainIdx', '<i4'), ('imgIdx', '<i4'), ('distance', '<f4')])


class MatcherParams(C.Structure):
    _fields_ = [('TH_LOW', C.c_int), ('TH_HIGH', C.c_int), ('HISTO_LENGTH', C.c_int),
                ('checkOrientation', C.c_int), ('nRatio', C.c_float)]


class StereoParams(C.Structure):
    _fields_ = [('bf', C.c_float), ('band', C.c_float), ('min_u', C.c_float), ('max_u', C.c_float),
                ('th_dist', C.c_int), ('octave_tol', C.c_int)]

    def __init__(self, bf=1.0, band=2.0, min_u=0.0, max_u=100000.0, th_dist=30, octave_tol=1):
        super().__init__(bf, band, min_u, max_u, th_dist, octave_tol)


class Pinhole(C.Structure):
    _fields_ = [('fx', C.c_float), ('fy', C.c_float), ('cx', C.c_float), ('cy', C.c_float),
                ('width', C.c_int), ('height', C.c_int), ('d', C.c_float * 5)]


def build(force=False):
    """Compile the C restatement (always) and oracle/_ref (only where /root/reference exists)."""
    so = os.path.join(_HERE, 'libtb_oracle.so')
    src_newer = (not os.path.exists(so)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(so)
        for f in ('tb_oracle.c', 'tb_oracle.h', 'orb_pattern.inc'))
    if force or src_newer or (os.path.exists('/root/reference/src/extractors/ORBextractor.cpp')
                              and not all(os.path.exists(os.path.join(_HERE, '_ref', n))
                                          for n in ('libtb_ref.so', 'libtb_ref_bow.so', 'libtb_ref_matcher.so'))):
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))


_lib = None
_ref = {}


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, 'libtb_oracle.so'))
        _lib.tbo_fast_atan2.restype = C.c_float
        _lib.tbo_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.tbo_ic_angle.restype = C.c_float
        _lib.tbo_shi_tomasi.restype = C.c_float
        _lib.tbo_grid_build.restype = C.c_void_p
    return _lib


def ref_available(unpinned=False):
    name = 'libtb_ref_unpinned.so' if unpinned else 'libtb_ref.so'
    return os.path.exists(os.path.join(_HERE, '_ref', name))


def ref_lib(unpinned=False):
    name = 'libtb_ref_unpinned.so' if unpinned else 'libtb_ref.so'
    if name not in _ref:
        build()
        L = C.CDLL(os.path.join(_HERE, '_ref', name))
        L.tbref_extractor_new.restype = C.c_void_p
        L.tbref_extractor_delete.argtypes = [C.c_void_p]
        _ref[name] = L
    return _ref[name]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


# ---------------------------------------------------------------- pyramid
def scale_factors(n_levels, scale):
    sf = np.zeros(n_levels, np.float32)
    lib().tbo_scale_factors(n_levels, C.c_float(scale), _p(sf))
    return sf


def level_sizes(w0, h0, sf):
    out = []
    for s in sf:
        w, h = C.c_int(), C.c_int()
        lib().tbo_level_size(w0, h0, C.c_float(float(s)), C.byref(w), C.byref(h))
        out.append((w.value, h.value))
    return out


def resize_linear(src, dw, dh):
    src = _u8(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().tbo_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def pyramid(img, n_levels, scale):
    """Frame::ComputePyramid: level i = resize(level i-1 -> int(cols*sf[i]) x int(rows*sf[i]))."""
    img = _u8(img)
    sf = scale_factors(n_levels, scale)
    sizes = level_sizes(img.shape[1], img.shape[0], sf)
    levels = [img]
    for i in range(1, n_levels):
        levels.append(resize_linear(levels[-1], sizes[i][0], sizes[i][1]))
    return levels, sf


# ---------------------------------------------------------------- primitives
def fast9(img, T, nonmax=True):
    img = _u8(img)
    cap = img.size // 2 + 16
    out = np.zeros(cap, KP_DTYPE)
    n = lib().tbo_fast9(_p(img), img.strides[0], img.shape[1], img.shape[0], int(T), int(nonmax), _p(out), cap)
    return out[:n]


def gauss7x7(img):
    img = _u8(img)
    dst = np.empty_like(img)
    lib().tbo_gauss7x7(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(dst), dst.strides[0])
    return dst


def fast_atan2(y, x):
    return lib().tbo_fast_atan2(float(y), float(x))


def ic_angle(img, cx, cy):
    img = _u8(img)
    return lib().tbo_ic_angle(_p(img), img.strides[0], int(cx), int(cy))


def rbrief(blur, cx, cy, angle_deg):
    blur = _u8(blur)
    d = np.zeros(32, np.uint8)
    lib().tbo_rbrief(_p(blur), blur.strides[0], int(cx), int(cy), C.c_float(angle_deg), _p(d))
    return d


def shi_tomasi(img, u, v):
    img = _u8(img)
    return lib().tbo_shi_tomasi(_p(img), img.shape[1], img.shape[0], img.strides[0], int(u), int(v))


# ---------------------------------------------------------------- extractor
def orb_budget(n_levels, factor, target):
    b = np.zeros(n_levels, np.int32)
    lib().tbo_orb_budget(n_levels, C.c_float(float(factor)), int(target), _p(b))
    return b


def cell_candidates(img, initTh, minTh):
    img = _u8(img)
    cap = img.size // 4 + 16
    out = np.zeros(cap, KP_DTYPE)
    n = lib().tbo_cell_candidates(_p(img), img.shape[1], img.shape[0], img.strides[0],
                                  C.c_float(initTh), C.c_float(minTh), _p(out), cap)
    return out[:n]


def distribute_octree(cand, exit_kps, minX, maxX, minY, maxY, N):
    cand = np.ascontiguousarray(cand, KP_DTYPE)
    ex = np.ascontiguousarray(exit_kps if exit_kps is not None else np.zeros(0, KP_DTYPE), KP_DTYPE)
    out = np.zeros(len(cand) + 8, KP_DTYPE)
    n = lib().tbo_distribute_octree(_p(cand), len(cand), _p(ex), len(ex), minX, maxX, minY, maxY, int(N),
                                    _p(out), len(out))
    return out[:n]


def _level_args(levels):
    levels = [_u8(l) for l in levels]
    n = len(levels)
    ptrs = (C.c_void_p * n)(*[l.ctypes.data for l in levels])
    w = (C.c_int * n)(*[l.shape[1] for l in levels])
    h = (C.c_int * n)(*[l.shape[0] for l in levels])
    s = (C.c_int * n)(*[l.strides[0] for l in levels])
    return levels, ptrs, w, h, s


def orb_extract(levels, sf, target, initTh, minTh, exit_kps=None, budget=None, cap=None):
    """ORBExtractor::operator() (budget None) or AddPoints (budget = previous per-level budget)."""
    levels, ptrs, w, h, s = _level_args(levels)
    n = len(levels)
    sf = np.ascontiguousarray(sf, np.float32)
    cap = cap or (4 * target + 256)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    npl = np.zeros(n, np.int32)
    ex = np.ascontiguousarray(exit_kps if exit_kps is not None else np.zeros(0, KP_DTYPE), KP_DTYPE)
    bud = None if budget is None else np.ascontiguousarray(budget, np.int32)
    cnt = lib().tbo_orb_extract(ptrs, w, h, s, n, _p(sf), int(target), C.c_float(initTh), C.c_float(minTh),
                                _p(ex), len(ex), None if bud is None else _p(bud), _p(kps), _p(desc), cap, _p(npl))
    assert cnt <= cap, (cnt, cap)
    return kps[:cnt].copy(), desc[:cnt].copy(), npl


class RefExtractor:
    """The reference's own ORBExtractor (oracle/_ref, compiled from /root/reference unmodified)."""

    def __init__(self, unpinned=False):
        self.L = ref_lib(unpinned)
        self.h = C.c_void_p(self.L.tbref_extractor_new())

    def __del__(self):
        try:
            self.L.tbref_extractor_delete(self.h)
        except Exception:
            pass

    def __call__(self, levels, sf, target, initTh, minTh, cap=None):
        levels, ptrs, w, h, s = _level_args(levels)
        sf = np.ascontiguousarray(sf, np.float32)
        cap = cap or (4 * target + 256)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.tbref_extract(self.h, ptrs, w, h, s, len(levels), _p(sf), int(target), C.c_float(initTh),
                                 C.c_float(minTh), _p(kps), _p(desc), cap)
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def add_points(self, levels, sf, target, initTh, minTh, exit_kps, cap=None):
        levels, ptrs, w, h, s = _level_args(levels)
        sf = np.ascontiguousarray(sf, np.float32)
        ex = np.ascontiguousarray(exit_kps, KP_DTYPE)
        cap = cap or (4 * target + 256)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.tbref_add_points(self.h, ptrs, w, h, s, len(levels), _p(sf), int(target), C.c_float(initTh),
                                    C.c_float(minTh), _p(ex), len(ex), _p(kps), _p(desc), cap)
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()


def fast10(img, barrier=20):
    img = _u8(img)
    cap = img.size // 2 + 16
    xy = np.zeros((cap, 2), np.int32)
    sc = np.zeros(cap, np.int32)
    n = lib().tbo_fast10(_p(img), img.shape[1], img.shape[0], img.strides[0], int(barrier), _p(xy), _p(sc), cap)
    return xy[:n].copy(), sc[:n].copy()


def fast_extract(levels, inv_sf, n_features, threshold, occupied=None):
    """FASTExtractor::operator() (occupied None) / AddPoints (occupied = grid_occupancy_ flags). Returns None
    where the reference would throw std::out_of_range."""
    levels, ptrs, w, h, s = _level_args(levels)
    inv_sf = np.ascontiguousarray(inv_sf, np.float32)
    out = np.zeros(n_features + 1, KP_DTYPE)
    occ = None if occupied is None else _u8(occupied)
    n = lib().tbo_fast_extract(ptrs, w, h, s, len(levels), _p(inv_sf), int(n_features), C.c_float(threshold),
                               None if occ is None else _p(occ), _p(out), len(out))
    return None if n < 0 else out[:n].copy()


def fast_grid(w0, h0, n_features):
    cell = int(np.sqrt(np.float32(np.float32(w0) * np.float32(h0) / np.float32(n_features))))
    return cell, int(np.float32(w0) / np.float32(cell)), int(np.float32(h0) / np.float32(cell))


# ---------------------------------------------------------------- matchers
def hamming256(a, b):
    a, b = _u8(a), _u8(b)
    return lib().tbo_hamming256(_p(a), _p(b))


def knn2(q, t):
    q, t = _u8(q), _u8(t)
    idx = np.zeros(len(q), np.int32)
    best = np.zeros(len(q), np.int32)
    second = np.zeros(len(q), np.int32)
    lib().tbo_knn2(_p(q), len(q), _p(t), len(t), _p(idx), _p(best), _p(second))
    return idx, best, second


def bf_cross(d1, d2):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros(max(len(d1), 1), MATCH_DTYPE)
    n = lib().tbo_bf_cross(_p(d1), len(d1), _p(d2), len(d2), _p(out))
    return out[:n].copy()


def search_by_bf(d1, d2, ratio, minTh):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros(max(len(d1), 1), MATCH_DTYPE)
    n = lib().tbo_search_by_bf(_p(d1), len(d1), _p(d2), len(d2), C.c_float(ratio), C.c_float(minTh), _p(out))
    return out[:n].copy()


def features_in_area(kps, cols, rows, x, y, r, minLevel=-1, maxLevel=-1):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    g = C.c_void_p(lib().tbo_grid_build(_p(kps), len(kps), cols, rows))
    out = np.zeros(len(kps) + 1, np.int32)
    n = lib().tbo_features_in_area(g, C.c_float(x), C.c_float(y), C.c_float(r), minLevel, maxLevel, _p(out), len(out))
    lib().tbo_grid_free(g)
    return out[:n].copy()


def search_by_violence(params, k1, d1, k2, d2, cols2, rows2, min_level=0, max_level=1, search_r=10.0):
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros(len(k1) + 1, MATCH_DTYPE)
    n = lib().tbo_search_by_violence(C.byref(params), _p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2),
                                     cols2, rows2, min_level, max_level, C.c_float(search_r), _p(out))
    return out[:n].copy()


class LKParams(C.Structure):
    _fields_ = [('win', C.c_int), ('max_level', C.c_int), ('max_iter', C.c_int), ('eps', C.c_float), ('min_eig', C.c_float)]

    def __init__(self, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=1e-4):
        super().__init__(win, max_level, max_iter, eps, min_eig)


def pyrdown(img):
    img = _u8(img)
    h, w = img.shape
    out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().tbo_pyrdown_u8(_p(img), w, h, img.strides[0], _p(out), out.strides[0])
    return out


def lk_flow(params, prev, nxt, pts):
    """cv::calcOpticalFlowPyrLK(prev, next, pts, ..., Size(21,21), 3) as Matcher::searchByOPFlow calls it
    (matcher.cpp:745) -> (next_pts [n,2] f32, status [n] u8, err [n] f32)."""
    prev, nxt = _u8(prev), _u8(nxt)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    h, w = prev.shape
    out = np.zeros_like(pts)
    st = np.zeros(len(pts), np.uint8)
    err = np.zeros(len(pts), np.float32)
    lib().tbo_lk_flow(C.byref(params), _p(prev), _p(nxt), w, h, prev.strides[0], _p(pts), len(pts), _p(out), _p(st), _p(err))
    return out, st, err


def reject_with_f(cur_pts, last_pts, status, threshold=1.0, confidence=0.99, max_iters=1000):
    """Matcher::rejectWithF (matcher.cpp:853-881): returns (status after the RANSAC rejection, F (3x3), iterations run)."""
    cur = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2); last = np.ascontiguousarray(last_pts, np.float32).reshape(-1, 2)
    st = np.ascontiguousarray(status, np.uint8).copy()
    F = np.zeros(9, np.float64)
    f = lib().tbo_reject_with_f
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p]
    it = f(_p(cur), _p(last), len(cur), _p(st), threshold, confidence, max_iters, _p(F))
    return st, F.reshape(3, 3), it


def stereo_rowband(params, kL, dL, kR, dR, inv_sf, n_rows):
    """LocalBA.cpp:69-288 (block-commented ORB stereo matcher) -> (matches, depth[nL])."""
    kL, kR = np.ascontiguousarray(kL, KP_DTYPE), np.ascontiguousarray(kR, KP_DTYPE)
    dL, dR = _u8(dL), _u8(dR)
    inv_sf = np.ascontiguousarray(inv_sf, np.float32)
    out = np.zeros(len(kL) + 1, MATCH_DTYPE)
    depth = np.zeros(len(kL), np.float32)
    n = lib().tbo_stereo_rowband(C.byref(params), _p(kL), _p(dL), len(kL), _p(kR), _p(dR), len(kR),
                                 _p(inv_sf), len(inv_sf), int(n_rows), _p(out), _p(depth))
    return out[:n].copy(), depth


def search_by_projection(params, cam, Tcw1, k1, d1, f1_matched, sf1, cols1, rows1, k2, mp_valid, mp_xyz, mp_desc):
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    d1, mp_desc = _u8(d1), _u8(mp_desc)
    Tcw1 = np.ascontiguousarray(Tcw1, np.float32)
    sf1 = np.ascontiguousarray(sf1, np.float32)
    mp_valid = _u8(mp_valid)
    mp_xyz = np.ascontiguousarray(mp_xyz, np.float32)
    fm = None if f1_matched is None else _u8(f1_matched)
    out = np.zeros(len(k2) + 1, MATCH_DTYPE)
    n = lib().tbo_search_by_projection(C.byref(params), C.byref(cam), _p(Tcw1), _p(k1), _p(d1), len(k1),
                                       None if fm is None else _p(fm), _p(sf1), len(sf1), cols1, rows1,
                                       _p(k2), len(k2), _p(mp_valid), _p(mp_xyz), _p(mp_desc), _p(out))
    return out[:n].copy()


def search_by_projection_map(params, cam, Tcw1, k1, d1, f1_matched, sf1, cols1, rows1,
                             mp_bad, mp_xyz, mp_normal, mp_desc, radio):
    k1 = np.ascontiguousarray(k1, KP_DTYPE)
    d1, mp_desc = _u8(d1), _u8(mp_desc)
    Tcw1 = np.ascontiguousarray(Tcw1, np.float32)
    sf1 = np.ascontiguousarray(sf1, np.float32)
    mp_bad = _u8(mp_bad)
    mp_xyz = np.ascontiguousarray(mp_xyz, np.float32)
    mp_normal = np.ascontiguousarray(mp_normal, np.float32)
    fm = None if f1_matched is None else _u8(f1_matched)
    out = np.zeros(len(mp_bad) + 1, MATCH_DTYPE)
    n = lib().tbo_search_by_projection_map(C.byref(params), C.byref(cam), _p(Tcw1), _p(k1), _p(d1), len(k1),
                                           None if fm is None else _p(fm), _p(sf1), cols1, rows1,
                                           len(mp_bad), _p(mp_bad), _p(mp_xyz), _p(mp_normal), _p(mp_desc),
                                           C.c_float(radio), _p(out))
    return out[:n].copy()


def three_maxima(hist_sizes):
    hs = np.ascontiguousarray(hist_sizes, np.int32)
    a, b, c = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    lib().tbo_three_maxima(_p(hs), len(hs), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


# ---- bag of words -------------------------------------------------------------------------------
class _Vocab(C.Structure):
    _fields_ = [('n_nodes', C.c_int), ('L', C.c_int), ('parent', C.c_void_p), ('child_start', C.c_void_p),
                ('children', C.c_void_p), ('word_id', C.c_void_p), ('desc', C.c_void_p), ('weight', C.c_void_p)]


def bow_transform(voc, desc, levelsup=4):
    """voc: trackingbench_slam_b200.vocab.Vocabulary (flat arrays only). Returns (word_id, node_id, weight)."""
    start, children, word = voc.csr()
    keep = (start, children, word, voc.parent, voc.desc, voc.weight)
    v = _Vocab(voc.n_nodes, voc.L, _p(voc.parent).value, _p(start).value, _p(children).value, _p(word).value,
               _p(voc.desc).value, _p(voc.weight).value)
    desc = _u8(desc)
    n = len(desc)
    wid, nid, w = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
    lib().tbo_bow_transform(C.byref(v), _p(desc), n, int(levelsup), _p(wid), _p(nid), _p(w))
    del keep
    return wid, nid, w


def search_by_bow(params, k1, d1, node1, k2, d2, node2, mp_valid2=None, map_point_only=False):
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    d1, d2 = _u8(d1), _u8(d2)
    node1, node2 = np.ascontiguousarray(node1, np.int32), np.ascontiguousarray(node2, np.int32)
    mv = None if mp_valid2 is None else _u8(mp_valid2)
    out = np.zeros(len(k1) + 1, MATCH_DTYPE)
    n = lib().tbo_search_by_bow(C.byref(params), _p(k1), _p(d1), _p(node1), len(k1), _p(k2), _p(d2), _p(node2), len(k2),
                                None if mv is None else _p(mv), int(bool(map_point_only)), _p(out))
    return out[:n].copy()


def ref_bow_available():
    return os.path.exists(os.path.join(_HERE, '_ref', 'libtb_ref_bow.so'))


class RefVocabulary:
    """The reference's vendored DBoW2 (oracle/_ref/libtb_ref_bow.so): loadFromTextFile + transform as Frame::SetBow."""

    def __init__(self, path):
        if 'bow' not in _ref:
            L = C.CDLL(os.path.join(_HERE, '_ref', 'libtb_ref_bow.so'))
            L.tbref_vocab_load_text.restype = C.c_void_p
            L.tbref_vocab_free.argtypes = [C.c_void_p]
            L.tbref_vocab_size.argtypes = [C.c_void_p]
            _ref['bow'] = L
        self.L = _ref['bow']
        self.h = self.L.tbref_vocab_load_text(path.encode())
        if not self.h:
            raise RuntimeError('reference loadFromTextFile failed for %s' % path)

    def size(self):
        return self.L.tbref_vocab_size(self.h)

    def set_bow(self, desc, levelsup=4):
        """Returns (node_id per feature (-1 = not in FeatureVector), word_id per feature, bow ids, bow values)."""
        desc = _u8(desc)
        n = len(desc)
        nid, wid = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ids, vals = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
        m = self.L.tbref_set_bow(C.c_void_p(self.h), _p(desc), n, int(levelsup), _p(nid), _p(wid), _p(ids), _p(vals), n + 1)
        return nid, wid, ids[:m].copy(), vals[:m].copy()

    def close(self):
        if self.h:
            self.L.tbref_vocab_free(self.h)
            self.h = None


# ---- frame ingest -------------------------------------------------------------------------------
def init_undistort_rectify_map(fx, fy, cx, cy, d, w, h):
    d = np.ascontiguousarray(d, np.float32)
    m1, m2 = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    lib().tbo_init_undistort_rectify_map(C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(d), w, h, _p(m1), _p(m2))
    return m1, m2


def undistort_points(fx, fy, cx, cy, d, pts):
    """PinholeCamera::UndistortPoint (CameraModel.cpp:104-121) on n points -> [n, 2] float32."""
    d = np.ascontiguousarray(d, np.float32)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(pts)
    lib().tbo_undistort_points(C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(d), _p(pts), len(pts), _p(out))
    return out


def remap_linear(src, map1, map2):
    src = _u8(src)
    map1, map2 = np.ascontiguousarray(map1, np.float32), np.ascontiguousarray(map2, np.float32)
    dst = np.zeros(map1.shape, np.uint8)
    lib().tbo_remap_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(map1), _p(map2),
                              map1.shape[1], map1.shape[0], _p(dst), dst.strides[0])
    return dst


def clahe(src, clip_limit=3.0, tiles=(8, 8)):
    src = _u8(src)
    dst = np.zeros_like(src)
    lib().tbo_clahe_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], C.c_double(clip_limit), int(tiles[0]), int(tiles[1]),
                       _p(dst), dst.strides[0])
    return dst


# ---- the reference's own Frame / MapPoint / Map / CameraModel / Matcher / FASTExtractor code -------------------
# oracle/_ref/libtb_ref_matcher.so (oracle/ref_matcher_build.cpp): Frame.cpp, MapPoint.cpp, Map.cpp, CameraModel.cpp and
# FASTextractor.cpp compiled unmodified, matcher.cpp's hot-path member functions by line range.
def ref_matcher_available():
    return os.path.exists(os.path.join(_HERE, '_ref', 'libtb_ref_matcher.so'))


def ref_matcher_lib():
    if 'matcher' not in _ref:
        build()
        L = C.CDLL(os.path.join(_HERE, '_ref', 'libtb_ref_matcher.so'))
        for n in ('tbrefm_camera_new', 'tbrefm_frame_new', 'tbrefm_map_points_new', 'tbrefm_matcher_new', 'tbrefm_vocab_load_text'):
            getattr(L, n).restype = C.c_void_p
        L.tbrefm_shi_tomasi.restype = C.c_float
        _ref['matcher'] = L
    return _ref['matcher']


class RefCamera:
    """TRACKING_BENCH::PinholeCamera (src/camera/CameraModel.cpp, compiled unmodified)."""

    def __init__(self, w, h, fx, fy, cx, cy, d=(0, 0, 0, 0, 0)):
        self.L = ref_matcher_lib()
        dd = np.ascontiguousarray(d, np.float32)
        self.w, self.h = int(w), int(h)
        self.h_ = C.c_void_p(self.L.tbrefm_camera_new(int(w), int(h), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(dd)))

    def world2cam(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32)
        uv = np.zeros(2, np.float32)
        self.L.tbrefm_camera_world2cam(self.h_, _p(xyz), _p(uv))
        return uv

    def is_in_frame(self, x, y):
        return bool(self.L.tbrefm_camera_is_in_frame(self.h_, int(x), int(y)))

    def undistort_point(self, uv):
        uv = np.ascontiguousarray(uv, np.float32)
        out = np.zeros(2, np.float32)
        self.L.tbrefm_camera_undistort_point(self.h_, _p(uv), _p(out))
        return out

    def undistort_image(self, raw):
        raw = _u8(raw)
        out = np.zeros((self.h, self.w), np.uint8)
        self.L.tbrefm_camera_undistort_image(self.h_, _p(raw), raw.shape[1], raw.shape[0], raw.strides[0], _p(out))
        return out


class RefFrame:
    """TRACKING_BENCH::Frame (src/types/Frame.cpp, compiled unmodified): constructor = scale tables + ComputePyramid."""

    def __init__(self, img, levels, scale, cam):
        self.L = ref_matcher_lib()
        img = _u8(img)
        self.levels, self.cam, self.n = int(levels), cam, 0
        self.h_ = C.c_void_p(self.L.tbrefm_frame_new(_p(img), img.shape[1], img.shape[0], img.strides[0], int(levels), C.c_float(scale), cam.h_))

    def close(self):
        if self.h_:
            self.L.tbrefm_frame_free(self.h_)
            self.h_ = None

    def geometry(self):
        n = self.levels
        f = [np.zeros(n, np.float32) for _ in range(4)]
        w, h = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.tbrefm_frame_geometry(self.h_, _p(f[0]), _p(f[1]), _p(f[2]), _p(f[3]), _p(w), _p(h))
        return dict(sf=f[0], inv_sf=f[1], sigma2=f[2], inv_sigma2=f[3], w=w, h=h)

    def level(self, l):
        g = self.geometry()
        out = np.zeros((int(g['h'][l]), int(g['w'][l])), np.uint8)
        self.L.tbrefm_frame_level(self.h_, int(l), _p(out))
        return out

    def equalize(self):
        g = self.geometry()
        out = np.zeros((int(g['h'][0]), int(g['w'][0])), np.uint8)
        self.L.tbrefm_frame_equalize(self.h_, _p(out))
        return out

    def set_keys(self, kps, desc, undistort=False):
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        desc = _u8(desc)
        self.n = len(kps)
        self.L.tbrefm_frame_set_keys(self.h_, _p(kps), _p(desc), len(kps), int(undistort))

    def assign_grid(self):
        self.L.tbrefm_frame_assign_grid(self.h_)

    def set_pose(self, Tcw):
        T = np.ascontiguousarray(Tcw, np.float32)
        self.L.tbrefm_frame_set_pose(self.h_, _p(T))

    def camera_center(self):
        o = np.zeros(3, np.float32)
        self.L.tbrefm_frame_camera_center(self.h_, _p(o))
        return o

    def pos_in_grid(self, kp):
        kp = np.ascontiguousarray(kp, KP_DTYPE).reshape(1)
        x, y = C.c_int(), C.c_int()
        ok = self.L.tbrefm_frame_pos_in_grid(self.h_, _p(kp), C.byref(x), C.byref(y))
        return bool(ok), x.value, y.value

    def features_in_area(self, x, y, r, minLevel=-1, maxLevel=-1):
        out = np.zeros(self.n + 1, np.int32)
        n = self.L.tbrefm_frame_features_in_area(self.h_, C.c_float(x), C.c_float(y), C.c_float(r), int(minLevel), int(maxLevel), _p(out), len(out))
        return out[:n].copy()

    def undistort_points(self):
        out = np.zeros((self.n, 2), np.float32)
        self.L.tbrefm_frame_undistort_points(self.h_, _p(out))
        return out

    def set_feature_vector(self, node):
        node = np.ascontiguousarray(node, np.int32)
        self.L.tbrefm_frame_set_feature_vector(self.h_, _p(node), len(node))

    def set_bow(self, voc_handle):
        self.L.tbrefm_frame_set_bow(self.h_, voc_handle)

    def feature_vector(self):
        node = np.zeros(self.n, np.int32)
        self.L.tbrefm_frame_get_feature_vector(self.h_, _p(node), self.n)
        return node

    def attach_map_point(self, idx, mps, i):
        self.L.tbrefm_frame_attach_map_point(self.h_, int(idx), mps.h_, int(i))

    def is_in_frustum(self, mps, i, view_cos_limit=0.5):
        lvl, x, y, vc = C.c_int(), C.c_float(), C.c_float(), C.c_float()
        ok = self.L.tbrefm_frame_is_in_frustum(self.h_, mps.h_, int(i), C.c_float(view_cos_limit), C.byref(lvl), C.byref(x), C.byref(y), C.byref(vc))
        return bool(ok), lvl.value, x.value, y.value, vc.value


class RefMapPoints:
    """MapPoint::MapPoint (src/types/MapPoint.cpp:13-44) for every valid slot, created from key i of `frame` and held in a
    Map; the points lie in one arena so that the Map's std::set (address order) iterates in slot order."""

    def __init__(self, frame, valid, xyz, attach=True):
        self.L = ref_matcher_lib()
        self.valid = _u8(valid)
        xyz = np.ascontiguousarray(xyz, np.float32)
        self.h_ = C.c_void_p(self.L.tbrefm_map_points_new(frame.h_, _p(self.valid), _p(xyz), len(self.valid), int(attach)))

    def close(self):
        if self.h_:
            self.L.tbrefm_map_points_free(self.h_)
            self.h_ = None

    def set_bad(self, i):
        self.L.tbrefm_map_point_set_bad(self.h_, int(i))

    def set_obs(self, i, n):
        self.L.tbrefm_map_point_set_obs(self.h_, int(i), int(n))

    def get(self, i):
        xyz, nrm, d = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(32, np.uint8)
        self.L.tbrefm_map_point_get(self.h_, int(i), _p(xyz), _p(nrm), _p(d))
        return xyz, nrm, d


class RefMatcher:
    """TRACKING_BENCH::Matcher: the member functions of src/matchers/matcher.cpp cut out by line range at build time."""

    def __init__(self, low=50, high=100, histo=30, check=True, ratio=0.0):
        self.L = ref_matcher_lib()
        self.h_ = C.c_void_p(self.L.tbrefm_matcher_new())
        self.set(low, high, histo, check, ratio)

    def __del__(self):
        try:
            self.L.tbrefm_matcher_free(self.h_)
        except Exception:
            pass

    def set(self, low, high, histo, check, ratio):
        self.L.tbrefm_matcher_set(self.h_, int(low), int(high), int(histo), int(bool(check)), C.c_float(ratio))

    @staticmethod
    def _out(n):
        return np.zeros(n + 1, MATCH_DTYPE)

    def search_by_bf(self, f1, f2, minLevel, maxLevel, ratio, minTh):
        out = self._out(f1.n)
        n = self.L.tbrefm_search_by_bf(self.h_, f1.h_, f2.h_, int(minLevel), int(maxLevel), C.c_float(ratio), C.c_float(minTh), _p(out), len(out))
        return out[:n].copy()

    def search_by_violence(self, f1, f2, min_level=0, max_level=1, search_r=10.0, mp_only=False):
        out = self._out(f1.n)
        n = self.L.tbrefm_search_by_violence(self.h_, f1.h_, f2.h_, int(min_level), int(max_level), C.c_float(search_r), int(mp_only), _p(out), len(out))
        return out[:n].copy()

    def search_by_projection(self, f1, f2):
        out = self._out(f2.n)
        n = self.L.tbrefm_search_by_projection(self.h_, f1.h_, f2.h_, _p(out), len(out))
        return out[:n].copy()

    def search_by_projection_map(self, mps, f1, radio):
        out = self._out(len(mps.valid))
        n = self.L.tbrefm_search_by_projection_map(self.h_, mps.h_, f1.h_, C.c_float(radio), _p(out), len(out))
        return out[:n].copy()

    def search_by_bow(self, f1, f2, mp_only=False):
        out = self._out(f1.n)
        n = self.L.tbrefm_search_by_bow(self.h_, f1.h_, f2.h_, int(mp_only), _p(out), len(out))
        return out[:n].copy()

    def search_by_opflow(self, f1, f2, equalized=False):
        out = self._out(f2.n)
        cur = np.zeros((f2.n, 2), np.float32)
        n = self.L.tbrefm_search_by_opflow(self.h_, f1.h_, f2.h_, int(equalized), _p(cur), _p(out), len(out))
        return out[:n].copy(), cur

    def descriptor_distance(self, a, b):
        a, b = _u8(a), _u8(b)
        return self.L.tbrefm_descriptor_distance(_p(a), _p(b))

    def three_maxima(self, hist_sizes):
        hs = np.ascontiguousarray(hist_sizes, np.int32)
        a, b, c = C.c_int(-1), C.c_int(-1), C.c_int(-1)
        self.L.tbrefm_three_maxima(_p(hs), len(hs), C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


def ref_fast_extract(levels, inv_sf, n_features, threshold, exit_kps=None):
    """The reference's FASTExtractor (src/extractors/FASTextractor.cpp compiled unmodified; uzh-rpg/fast stand-in):
    operator() (exit_kps None) / AddPoints. None where it throws std::out_of_range."""
    L = ref_matcher_lib()
    levels, ptrs, w, h, s = _level_args(levels)
    inv_sf = np.ascontiguousarray(inv_sf, np.float32)
    out = np.zeros(n_features + 1, KP_DTYPE)
    ex = None if exit_kps is None else np.ascontiguousarray(exit_kps, KP_DTYPE)
    n = L.tbrefm_fast_extract(ptrs, w, h, s, len(levels), _p(inv_sf), int(n_features), C.c_float(threshold),
                              None if ex is None else _p(ex), -1 if ex is None else len(ex), _p(out), len(out))
    return None if n < 0 else out[:n].copy()


def ref_shi_tomasi(img, u, v):
    img = _u8(img)
    return ref_matcher_lib().tbrefm_shi_tomasi(_p(img), img.shape[1], img.shape[0], img.strides[0], int(u), int(v))


def ref_stereo_rowband(current, stereo, bf, fx):
    """The block-commented ORB stereo matcher of LocalBA::AddMapPointsByStereo (src/mapping/LocalBA.cpp:69-288) compiled from
    its own text (oracle/ref_matcher_build.cpp): current/stereo = RefFrame with keys -> (matches, depth[N])."""
    L = ref_matcher_lib()
    depth = np.zeros(current.n, np.float32)
    out = np.zeros(current.n + 1, MATCH_DTYPE)
    n = L.tbrefm_stereo_rowband(current.h_, stereo.h_, C.c_float(bf), C.c_float(fx), _p(depth), _p(out), len(out))
    return out[:n].copy(), depth
