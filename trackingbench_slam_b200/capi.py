"""ctypes binding of libtb_b200.so — exactly the C ABI declared in include/tb_capi.h.

This module is plumbing: it loads the CUDA library (failing loudly if it is missing — there is no
CPU fallback) and converts numpy arrays / raw device pointers to the POD arguments.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TB_LIB_PATH') or os.path.join(_HERE, 'libtb_b200.so')   # TB_LIB_PATH: developer knob (A/B builds)

KP_DTYPE = np.dtype([('x', '<f4'), ('y', '<f4'), ('size', '<f4'), ('angle', '<f4'),
                     ('response', '<f4'), ('octave', '<i4'), ('class_id', '<i4')])
MATCH_DTYPE = np.dtype([('queryIdx', '<i4'), ('trainIdx', '<i4'), ('imgIdx', '<i4'), ('distance', '<f4')])

TB_OK, TB_E_INVALID, TB_E_CUDA, TB_E_CAPACITY, TB_E_UNSUPPORTED, TB_E_NOMEM = 0, -1, -2, -3, -4, -5
TB_MAX_LEVELS = 16

# every symbol include/tb_capi.h declares (tests check that the library exports all of them)
EXPORTS = [
    'tb_ctx_create', 'tb_ctx_destroy', 'tb_sync', 'tb_last_error', 'tb_version', 'tb_launch_count',
    'tb_profile_enable', 'tb_profile_report', 'tb_set_match_impl',
    'tb_pyramid_geometry',
    'tb_frontend_create', 'tb_frontend_destroy', 'tb_frontend_kp_capacity', 'tb_frontend_level_layout',
    'tb_frontend_get_budget', 'tb_frontend_set_budget',
    'tb_frontend_upload_level', 'tb_frontend_download_level', 'tb_frontend_upload_frames',
    'tb_frontend_build_pyramid', 'tb_frontend_extract', 'tb_frontend_results_device',
    'tb_frontend_get_results', 'tb_frontend_get_results_batch', 'tb_frontend_get_results_batch_async',
    'tb_frontend_status', 'tb_frontend_get_candidates',
    'tb_frontend_get_level_counts', 'tb_orb_extract_host', 'tb_fast_extract_host',
    'tb_hamming_pairs_device', 'tb_knn2_device', 'tb_match_bf_cross_device', 'tb_search_by_bf_device',
    'tb_match_bf_cross_host', 'tb_search_by_bf_host', 'tb_hamming_pairs_host',
    'tb_frontend_search_by_bf', 'tb_frontend_search_by_projection', 'tb_features_in_area', 'tb_search_by_violence', 'tb_search_by_projection',
    'tb_search_by_projection_map', 'tb_three_maxima',
    'tb_vocab_create', 'tb_vocab_destroy', 'tb_bow_transform', 'tb_bow_transform_device', 'tb_search_by_bow',
    'tb_frontend_set_bow', 'tb_frontend_search_by_bow', 'tb_frontend_search_by_violence',
    'tb_init_undistort_rectify_map', 'tb_undistort_create', 'tb_undistort_destroy', 'tb_undistort_image_device',
    'tb_undistort_image_host', 'tb_frontend_undistort_device', 'tb_frontend_upload_frames_undistort',
    'tb_clahe_device', 'tb_clahe_host',
    'tb_stereo_match', 'tb_frontend_stereo_match',
    'tb_lk_flow_host', 'tb_lk_flow_device', 'tb_frontend_search_by_opflow',
    'tb_undistort_points', 'tb_undistort_points_device', 'tb_frontend_undistort_points',
    'tb_frontend_stereo_depth_by_opflow',
    'tb_reject_with_f_host', 'tb_reject_with_f_device', 'tb_frontend_reject_with_f',
]


class MatcherParams(C.Structure):
    _fields_ = [('TH_LOW', C.c_int32), ('TH_HIGH', C.c_int32), ('HISTO_LENGTH', C.c_int32),
                ('checkOrientation', C.c_int32), ('nRatio', C.c_float)]


class StereoParams(C.Structure):
    """tb_stereo_params; defaults = the constants of LocalBA.cpp:69-288 (bf is the caller's)."""
    _fields_ = [('bf', C.c_float), ('band', C.c_float), ('min_u', C.c_float), ('max_u', C.c_float),
                ('th_dist', C.c_int32), ('octave_tol', C.c_int32)]

    def __init__(self, bf=1.0, band=2.0, min_u=0.0, max_u=100000.0, th_dist=30, octave_tol=1):
        super().__init__(bf, band, min_u, max_u, th_dist, octave_tol)


class LKParams(C.Structure):
    """tb_lk_params; defaults = the arguments of matcher.cpp:745 and OpenCV's defaults."""
    _fields_ = [('win', C.c_int32), ('max_level', C.c_int32), ('max_iter', C.c_int32), ('eps', C.c_float), ('min_eig', C.c_float)]

    def __init__(self, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=1e-4):
        super().__init__(win, max_level, max_iter, eps, min_eig)


class RansacParams(C.Structure):
    """tb_ransac_params; defaults = the arguments of matcher.cpp:872 (cv::findFundamentalMat(.., FM_RANSAC, 1.0, 0.99, ..))."""
    _fields_ = [('threshold', C.c_float), ('confidence', C.c_float), ('max_iters', C.c_int32)]

    def __init__(self, threshold=1.0, confidence=0.99, max_iters=1000):
        super().__init__(threshold, confidence, max_iters)


class Pinhole(C.Structure):
    _fields_ = [('fx', C.c_float), ('fy', C.c_float), ('cx', C.c_float), ('cy', C.c_float),
                ('width', C.c_int32), ('height', C.c_int32), ('d', C.c_float * 5)]


class TbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('tb error %d: %s' % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libtb_b200.so. Raises if the CUDA extension has not been built — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('%s is missing: run `python -m trackingbench_slam_b200.build` '
                              '(there is no CPU fallback)' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.tb_last_error.restype = C.c_char_p
        L.tb_version.restype = C.c_char_p
        L.tb_launch_count.restype = C.c_int64
        for name in EXPORTS:
            getattr(L, name)
        _lib = L
    return _lib


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def pyramid_geometry(width, height, n_levels, scale):
    sf = np.zeros(n_levels, np.float32)
    lw = np.zeros(n_levels, np.int32)
    lh = np.zeros(n_levels, np.int32)
    rc = lib().tb_pyramid_geometry(width, height, n_levels, C.c_float(scale), _vp(sf), _vp(lw), _vp(lh))
    if rc:
        raise TbError(rc, 'tb_pyramid_geometry')
    return sf, lw, lh


class Context:
    def __init__(self, device=0, stream=None):
        self.h = C.c_void_p()
        rc = lib().tb_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc:
            raise TbError(rc, 'tb_ctx_create failed (no usable CUDA device?)')

    def check(self, rc, allow=()):
        if rc and rc not in allow:
            raise TbError(rc, lib().tb_last_error(self.h).decode())
        return rc

    def sync(self):
        self.check(lib().tb_sync(self.h))

    def launch_count(self):
        return lib().tb_launch_count(self.h)

    def set_match_impl(self, impl):
        """0 = LOP3+POPC (K7), 1 = tcgen05 int8 (K8)."""
        self.check(lib().tb_set_match_impl(self.h, int(impl)))

    def profile_enable(self, on=True):
        self.check(lib().tb_profile_enable(self.h, int(on)))

    def profile_report(self):
        """{kernel name: (launches, total_ms)} since the last report."""
        buf = C.create_string_buffer(1 << 16)
        self.check(lib().tb_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    def close(self):
        if self.h:
            lib().tb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    # ---- matching on device pointers (ints) or numpy arrays uploaded by the caller elsewhere
    def hamming_pairs_device(self, a_dptr, b_dptr, n, out_dptr):
        self.check(lib().tb_hamming_pairs_device(self.h, _vp(a_dptr), _vp(b_dptr), int(n), _vp(out_dptr)))

    def knn2_device(self, q_dptr, nq, t_dptr, nt, idx_dptr, best_dptr, second_dptr):
        self.check(lib().tb_knn2_device(self.h, _vp(q_dptr), int(nq), _vp(t_dptr), int(nt),
                                        _vp(idx_dptr), _vp(best_dptr), _vp(second_dptr)))

    def match_bf_cross_device(self, d1_dptr, n1, d2_dptr, n2):
        out = np.zeros(max(int(n1), 1), MATCH_DTYPE)
        n = C.c_int()
        self.check(lib().tb_match_bf_cross_device(self.h, _vp(d1_dptr), int(n1), _vp(d2_dptr), int(n2),
                                                  _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def search_by_bf_device(self, d1_dptr, n1, d2_dptr, n2, ratio, minTh):
        out = np.zeros(max(int(n1), 1), MATCH_DTYPE)
        n = C.c_int()
        self.check(lib().tb_search_by_bf_device(self.h, _vp(d1_dptr), int(n1), _vp(d2_dptr), int(n2),
                                                C.c_float(ratio), C.c_float(minTh), _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def fast_extract_host(self, levels, inv_sf, n_features, threshold, exit_kps=None):
        """FASTExtractor::operator() / AddPoints. Raises TbError(TB_E_INVALID) where the reference throws."""
        levels = [np.ascontiguousarray(l, np.uint8) for l in levels]
        n = len(levels)
        ptrs = (C.c_void_p * n)(*[l.ctypes.data for l in levels])
        w = np.array([l.shape[1] for l in levels], np.int32)
        h = np.array([l.shape[0] for l in levels], np.int32)
        strides = (C.c_size_t * n)(*[l.strides[0] for l in levels])
        inv_sf = np.ascontiguousarray(inv_sf, np.float32)
        ex = None if exit_kps is None else np.ascontiguousarray(exit_kps, KP_DTYPE)
        out = np.zeros(n_features + 1, KP_DTYPE)
        cnt = C.c_int()
        self.check(lib().tb_fast_extract_host(self.h, ptrs, _vp(w), _vp(h), strides, n, _vp(inv_sf), int(n_features), C.c_float(threshold),
                                              _vp(ex), 0 if ex is None else len(ex), _vp(out), len(out), C.byref(cnt)))
        return out[:cnt.value].copy()

    # ---- windowed matchers (host in / host out)
    def features_in_area(self, kps, cols, rows, x, y, r, minLevel=-1, maxLevel=-1):
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        out = np.zeros(len(kps) + 1, np.int32)
        n = C.c_int()
        self.check(lib().tb_features_in_area(self.h, _vp(kps), len(kps), int(cols), int(rows), C.c_float(x), C.c_float(y),
                                             C.c_float(r), int(minLevel), int(maxLevel), _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def search_by_violence(self, params, k1, d1, k2, d2, cols2, rows2, min_level=0, max_level=1, search_r=10.0):
        k1 = np.ascontiguousarray(k1, KP_DTYPE)
        k2 = np.ascontiguousarray(k2, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8)
        d2 = np.ascontiguousarray(d2, np.uint8)
        out = np.zeros(len(k1) + 1, MATCH_DTYPE)
        n = C.c_int()
        self.check(lib().tb_search_by_violence(self.h, C.byref(params), _vp(k1), _vp(d1), len(k1), _vp(k2), _vp(d2), len(k2),
                                               int(cols2), int(rows2), int(min_level), int(max_level), C.c_float(search_r),
                                               _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def stereo_match(self, params, kL, dL, kR, dR, inv_sf, n_rows):
        """ORB stereo row-band matcher (LocalBA.cpp:69-288) -> (matches, depth[nL])."""
        kL = np.ascontiguousarray(kL, KP_DTYPE)
        kR = np.ascontiguousarray(kR, KP_DTYPE)
        dL = np.ascontiguousarray(dL, np.uint8)
        dR = np.ascontiguousarray(dR, np.uint8)
        inv_sf = np.ascontiguousarray(inv_sf, np.float32)
        out = np.zeros(len(kL) + 1, MATCH_DTYPE)
        depth = np.zeros(len(kL), np.float32)
        n = C.c_int()
        self.check(lib().tb_stereo_match(self.h, C.byref(params), _vp(kL), _vp(dL), len(kL), _vp(kR), _vp(dR), len(kR),
                                         _vp(inv_sf), len(inv_sf), int(n_rows), _vp(out), len(out), C.byref(n), _vp(depth)))
        return out[:n.value].copy(), depth

    def lk_flow(self, params, prev, nxt, pts):
        """cv::calcOpticalFlowPyrLK(prev, next, pts, ...) as matcher.cpp:745 -> (next_pts [n,2], status [n], err [n])."""
        prev = np.ascontiguousarray(prev, np.uint8)
        nxt = np.ascontiguousarray(nxt, np.uint8)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        h, w = prev.shape
        assert nxt.shape == prev.shape
        out = np.zeros_like(pts)
        st = np.zeros(len(pts), np.uint8)
        err = np.zeros(len(pts), np.float32)
        self.check(lib().tb_lk_flow_host(self.h, C.byref(params), _vp(prev), _vp(nxt), w, h, C.c_size_t(prev.strides[0]), _vp(pts), len(pts),
                                         _vp(out), _vp(st), _vp(err)))
        return out, st, err

    def reject_with_f(self, cur_pts, last_pts, status, params=None):
        """Matcher::rejectWithF (matcher.cpp:853-881), host in/out -> (status after the rejection, F [3,3], iterations)."""
        params = params or RansacParams()
        cur = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2)
        last = np.ascontiguousarray(last_pts, np.float32).reshape(-1, 2)
        st = np.ascontiguousarray(status, np.uint8).copy()
        F = np.zeros(9, np.float64)
        it = C.c_int()
        self.check(lib().tb_reject_with_f_host(self.h, C.byref(params), _vp(cur), _vp(last), len(cur), _vp(st), _vp(F), C.byref(it)))
        return st, F.reshape(3, 3), it.value

    def reject_with_f_device(self, n_pairs, cur_dptr, last_dptr, counts_dptr, pts_cap, status_dptr, F_dptr=None, iters_dptr=None, params=None):
        params = params or RansacParams()
        self.check(lib().tb_reject_with_f_device(self.h, C.byref(params), int(n_pairs), _vp(cur_dptr), _vp(last_dptr), _vp(counts_dptr),
                                                 int(pts_cap), _vp(status_dptr), _vp(F_dptr), _vp(iters_dptr)))

    def lk_flow_device(self, params, prev_dptr, prev_pitch, prev_frame_stride, n_prev, next_dptr, next_pitch, next_frame_stride, n_next, w, h,
                       prev_idx, next_idx, pts_dptr, counts_dptr, pts_cap, next_pts_dptr, status_dptr, err_dptr=None):
        a = np.ascontiguousarray(prev_idx, np.int32)
        b = np.ascontiguousarray(next_idx, np.int32)
        self.check(lib().tb_lk_flow_device(self.h, C.byref(params), _vp(prev_dptr), C.c_size_t(prev_pitch), C.c_size_t(prev_frame_stride), int(n_prev),
                                           _vp(next_dptr), C.c_size_t(next_pitch), C.c_size_t(next_frame_stride), int(n_next), int(w), int(h),
                                           len(a), _vp(a), _vp(b), _vp(pts_dptr), _vp(counts_dptr), int(pts_cap), _vp(next_pts_dptr),
                                           _vp(status_dptr), _vp(err_dptr)))

    def undistort_points(self, cam, pts):
        """PinholeCamera::UndistortPoint on n points (host in/out) -> [n, 2] float32."""
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        out = np.zeros_like(pts)
        self.check(lib().tb_undistort_points(self.h, C.byref(cam), _vp(pts), len(pts), _vp(out)))
        return out

    def undistort_points_device(self, cam, pts_dptr, stride, n, out_dptr):
        self.check(lib().tb_undistort_points_device(self.h, C.byref(cam), _vp(pts_dptr), C.c_size_t(stride), C.c_longlong(n), _vp(out_dptr)))

    def vocab_create(self, voc):
        """voc: vocab.Vocabulary -> DeviceVocabulary (tree resident in HBM)."""
        return DeviceVocabulary(self, voc)

    def bow_transform(self, dvoc, desc, levelsup=4):
        """Frame::SetBow's per-feature transform: (word_id, node_id, weight)."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        wid, nid, w = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        self.check(lib().tb_bow_transform(self.h, dvoc.h, _vp(desc), n, int(levelsup), _vp(wid), _vp(nid), _vp(w)))
        return wid, nid, w

    def bow_transform_device(self, dvoc, desc_dptr, n, levelsup, word_dptr, node_dptr, weight_dptr):
        self.check(lib().tb_bow_transform_device(self.h, dvoc.h, C.c_void_p(desc_dptr), C.c_longlong(n), int(levelsup),
                                                 C.c_void_p(word_dptr), C.c_void_p(node_dptr), C.c_void_p(weight_dptr)))

    def search_by_bow(self, params, k1, d1, node1, k2, d2, node2, mp_valid2=None, map_point_only=False):
        k1 = np.ascontiguousarray(k1, KP_DTYPE)
        k2 = np.ascontiguousarray(k2, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8)
        d2 = np.ascontiguousarray(d2, np.uint8)
        node1 = np.ascontiguousarray(node1, np.int32)
        node2 = np.ascontiguousarray(node2, np.int32)
        mv = None if mp_valid2 is None else np.ascontiguousarray(mp_valid2, np.uint8)
        out = np.zeros(len(k1) + 1, MATCH_DTYPE)
        n = C.c_int()
        self.check(lib().tb_search_by_bow(self.h, C.byref(params), _vp(k1), _vp(d1), _vp(node1), len(k1), _vp(k2), _vp(d2),
                                          _vp(node2), len(k2), _vp(mv), int(bool(map_point_only)), _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def clahe_host(self, img, clip_limit=3.0, tiles=(8, 8)):
        """Frame::Equalize on one host image."""
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros_like(img)
        self.check(lib().tb_clahe_host(self.h, _vp(img), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]), _vp(out),
                                       C.c_size_t(out.strides[0]), C.c_double(clip_limit), int(tiles[0]), int(tiles[1])))
        return out

    def clahe_device(self, src_dptr, w, h, src_pitch, src_frame_stride, dst_dptr, dst_pitch, dst_frame_stride, n_frames,
                     clip_limit=3.0, tiles=(8, 8)):
        self.check(lib().tb_clahe_device(self.h, C.c_void_p(src_dptr), int(w), int(h), C.c_size_t(src_pitch), C.c_size_t(src_frame_stride),
                                         C.c_void_p(dst_dptr), C.c_size_t(dst_pitch), C.c_size_t(dst_frame_stride), int(n_frames),
                                         C.c_double(clip_limit), int(tiles[0]), int(tiles[1])))

    def search_by_projection(self, params, cam, Tcw1, k1, d1, f1_matched, sf1, cols1, rows1, k2, mp_valid, mp_xyz, mp_desc):
        k1 = np.ascontiguousarray(k1, KP_DTYPE)
        k2 = np.ascontiguousarray(k2, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        Tcw1 = np.ascontiguousarray(Tcw1, np.float32)
        sf1 = np.ascontiguousarray(sf1, np.float32)
        mp_valid = np.ascontiguousarray(mp_valid, np.uint8)
        mp_xyz = np.ascontiguousarray(mp_xyz, np.float32)
        fm = None if f1_matched is None else np.ascontiguousarray(f1_matched, np.uint8)
        out = np.zeros(len(k2) + 1, MATCH_DTYPE)
        n = C.c_int()
        self.check(lib().tb_search_by_projection(self.h, C.byref(params), C.byref(cam), _vp(Tcw1), _vp(k1), _vp(d1), len(k1),
                                                 _vp(fm), _vp(sf1), len(sf1), int(cols1), int(rows1), _vp(k2), len(k2),
                                                 _vp(mp_valid), _vp(mp_xyz), _vp(mp_desc), _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def search_by_projection_map(self, params, cam, Tcw1, k1, d1, f1_matched, sf1, cols1, rows1,
                                 mp_bad, mp_xyz, mp_normal, mp_desc, radio):
        k1 = np.ascontiguousarray(k1, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        Tcw1 = np.ascontiguousarray(Tcw1, np.float32)
        sf1 = np.ascontiguousarray(sf1, np.float32)
        mp_bad = np.ascontiguousarray(mp_bad, np.uint8)
        mp_xyz = np.ascontiguousarray(mp_xyz, np.float32)
        mp_normal = np.ascontiguousarray(mp_normal, np.float32)
        fm = None if f1_matched is None else np.ascontiguousarray(f1_matched, np.uint8)
        out = np.zeros(len(mp_bad) + 1, MATCH_DTYPE)
        n = C.c_int()
        self.check(lib().tb_search_by_projection_map(self.h, C.byref(params), C.byref(cam), _vp(Tcw1), _vp(k1), _vp(d1), len(k1),
                                                     _vp(fm), _vp(sf1), int(cols1), int(rows1), len(mp_bad), _vp(mp_bad),
                                                     _vp(mp_xyz), _vp(mp_normal), _vp(mp_desc), C.c_float(radio),
                                                     _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()


def init_undistort_rectify_map(cam):
    """cv::initUndistortRectifyMap as PinholeCamera's constructor calls it: (map1 = x, map2 = y), float32 [h, w]."""
    m1 = np.zeros((cam.height, cam.width), np.float32)
    m2 = np.zeros((cam.height, cam.width), np.float32)
    rc = lib().tb_init_undistort_rectify_map(C.byref(cam), _vp(m1), _vp(m2))
    if rc:
        raise TbError(rc, 'tb_init_undistort_rectify_map')
    return m1, m2


class Undistorter:
    """tb_undistort: PinholeCamera's undistortion maps in fixed point on the device."""

    def __init__(self, ctx, map1, map2):
        map1 = np.ascontiguousarray(map1, np.float32)
        map2 = np.ascontiguousarray(map2, np.float32)
        assert map1.shape == map2.shape and map1.ndim == 2
        self.ctx, self.h_, self.w_ = ctx, map1.shape[0], map1.shape[1]
        self.h = C.c_void_p()
        ctx.check(lib().tb_undistort_create(ctx.h, self.w_, self.h_, _vp(map1), _vp(map2), C.byref(self.h)))

    def apply_host(self, img):
        """PinholeCamera::UndistortImage on one host image."""
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros((self.h_, self.w_), np.uint8)
        self.ctx.check(lib().tb_undistort_image_host(self.ctx.h, self.h, _vp(img), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]),
                                                     _vp(out), C.c_size_t(out.strides[0])))
        return out

    def apply_device(self, src_dptr, src_w, src_h, src_pitch, src_frame_stride, dst_dptr, dst_pitch, dst_frame_stride, n_frames):
        self.ctx.check(lib().tb_undistort_image_device(self.ctx.h, self.h, C.c_void_p(src_dptr), int(src_w), int(src_h), C.c_size_t(src_pitch),
                                                       C.c_size_t(src_frame_stride), C.c_void_p(dst_dptr), C.c_size_t(dst_pitch),
                                                       C.c_size_t(dst_frame_stride), int(n_frames)))

    def close(self):
        if self.h:
            lib().tb_undistort_destroy(self.h)
            self.h = C.c_void_p()


class DeviceVocabulary:
    """tb_vocab: the flattened ORBVocabulary on the device."""

    def __init__(self, ctx, voc):
        self.ctx = ctx
        self.h = C.c_void_p()
        ctx.check(lib().tb_vocab_create(ctx.h, voc.n_nodes, voc.L, _vp(voc.parent), _vp(voc.is_leaf), _vp(voc.desc),
                                        _vp(voc.weight), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().tb_vocab_destroy(self.h)
            self.h = C.c_void_p()


class Frontend:
    """tb_frontend: device-resident batch of equally sized frames."""

    def __init__(self, ctx, level_w, level_h, max_batch=1, max_target=2000):
        self.ctx = ctx
        self.level_w = np.ascontiguousarray(level_w, np.int32)
        self.level_h = np.ascontiguousarray(level_h, np.int32)
        self.n_levels = len(self.level_w)
        self.max_batch = max_batch
        self.h = C.c_void_p()
        ctx.check(lib().tb_frontend_create(ctx.h, self.n_levels, _vp(self.level_w), _vp(self.level_h),
                                           int(max_batch), int(max_target), C.byref(self.h)))
        self.kp_cap = lib().tb_frontend_kp_capacity(self.h)

    def close(self):
        if self.h:
            lib().tb_frontend_destroy(self.h)
            self.h = C.c_void_p()

    def level_layout(self, level, blurred=False):
        base, pitch, fs = C.c_void_p(), C.c_size_t(), C.c_size_t()
        self.ctx.check(lib().tb_frontend_level_layout(self.h, level, int(blurred), C.byref(base), C.byref(pitch), C.byref(fs)))
        return base.value, pitch.value, fs.value

    def upload_level(self, frame, level, img):
        img = np.ascontiguousarray(img, np.uint8)
        assert img.shape == (self.level_h[level], self.level_w[level]), (img.shape, level)
        self.ctx.check(lib().tb_frontend_upload_level(self.h, frame, level, _vp(img), img.strides[0]))
        self.ctx.sync()

    def upload_frames(self, first, frames):
        """frames: (n, H, W) u8 contiguous numpy array or a host pointer tuple (ptr, n)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.shape[1:] == (self.level_h[0], self.level_w[0])
        self.ctx.check(lib().tb_frontend_upload_frames(self.h, first, frames.shape[0], _vp(frames), frames.strides[1]))

    def upload_frames_ptr(self, first, n, host_ptr, host_stride):
        self.ctx.check(lib().tb_frontend_upload_frames(self.h, first, n, C.c_void_p(host_ptr), host_stride))

    def download_level(self, frame, level, blurred=False):
        out = np.empty((self.level_h[level], self.level_w[level]), np.uint8)
        self.ctx.check(lib().tb_frontend_download_level(self.h, frame, level, int(blurred), _vp(out), out.strides[0]))
        return out

    def build_pyramid(self, n_frames):
        self.ctx.check(lib().tb_frontend_build_pyramid(self.h, n_frames))

    def extract(self, n_frames, sf, target, initTh, minTh, exit_kps=None):
        sf = np.ascontiguousarray(sf, np.float32)
        ex = None if exit_kps is None else np.ascontiguousarray(exit_kps, KP_DTYPE)
        self.ctx.check(lib().tb_frontend_extract(self.h, n_frames, _vp(sf), int(target), C.c_float(initTh), C.c_float(minTh),
                                                 _vp(ex), 0 if ex is None else len(ex)))

    def get_budget(self):
        """ORBExtractor::mnFeaturesPerLevel of the last extract (empty before the first one)."""
        b = np.zeros(16, np.int32)
        n = C.c_int()
        self.ctx.check(lib().tb_frontend_get_budget(self.h, _vp(b), C.byref(n)))
        return b[:n.value].copy()

    def set_budget(self, per_level):
        b = np.ascontiguousarray(per_level, np.int32)
        self.ctx.check(lib().tb_frontend_set_budget(self.h, _vp(b), len(b)))

    def results_device(self):
        k, d, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().tb_frontend_results_device(self.h, C.byref(k), C.byref(d), C.byref(c)))
        return k.value, d.value, c.value

    def get_results(self, frame):
        kps = np.zeros(self.kp_cap, KP_DTYPE)
        desc = np.zeros((self.kp_cap, 32), np.uint8)
        n = C.c_int()
        self.ctx.check(lib().tb_frontend_get_results(self.h, frame, _vp(kps), _vp(desc), self.kp_cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def get_results_batch(self, n_frames, kps=None, desc=None, counts=None):
        kps = np.zeros((n_frames, self.kp_cap), KP_DTYPE) if kps is None else kps
        desc = np.zeros((n_frames, self.kp_cap, 32), np.uint8) if desc is None else desc
        counts = np.zeros(n_frames, np.int32) if counts is None else counts
        self.ctx.check(lib().tb_frontend_get_results_batch(self.h, n_frames, _vp(kps), _vp(desc), self.kp_cap, _vp(counts)))
        return kps, desc, counts

    def get_results_batch_async(self, n_frames, kps, desc, counts):
        """Enqueue the D2H copies only (pinned numpy views); call status() before reading."""
        self.ctx.check(lib().tb_frontend_get_results_batch_async(self.h, n_frames, _vp(kps), _vp(desc), self.kp_cap, _vp(counts)))

    def status(self):
        self.ctx.check(lib().tb_frontend_status(self.h))

    def get_candidates(self, frame, level):
        cap = int(self.level_w[level]) * int(self.level_h[level]) // 4 + 64
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int()
        self.ctx.check(lib().tb_frontend_get_candidates(self.h, frame, level, _vp(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def get_level_counts(self, frame):
        out = np.zeros(self.n_levels, np.int32)
        self.ctx.check(lib().tb_frontend_get_level_counts(self.h, frame, _vp(out)))
        return out

    def orb_extract_host(self, levels, sf, target, initTh, minTh, exit_kps=None):
        levels = [np.ascontiguousarray(l, np.uint8) for l in levels]
        n = len(levels)
        ptrs = (C.c_void_p * n)(*[l.ctypes.data for l in levels])
        strides = (C.c_size_t * n)(*[l.strides[0] for l in levels])
        sf = np.ascontiguousarray(sf, np.float32)
        ex = None if exit_kps is None else np.ascontiguousarray(exit_kps, KP_DTYPE)
        kps = np.zeros(self.kp_cap, KP_DTYPE)
        desc = np.zeros((self.kp_cap, 32), np.uint8)
        cnt = C.c_int()
        self.ctx.check(lib().tb_orb_extract_host(self.h, ptrs, strides, _vp(sf), int(target), C.c_float(initTh), C.c_float(minTh),
                                                 _vp(ex), 0 if ex is None else len(ex), _vp(kps), _vp(desc), self.kp_cap, C.byref(cnt)))
        return kps[:cnt.value].copy(), desc[:cnt.value].copy()

    def search_by_bf(self, q_frames, t_frames, ratio, minTh, matches_dptr, capacity, counts_dptr):
        q = np.ascontiguousarray(q_frames, np.int32)
        t = np.ascontiguousarray(t_frames, np.int32)
        self.ctx.check(lib().tb_frontend_search_by_bf(self.h, len(q), _vp(q), _vp(t), C.c_float(ratio), C.c_float(minTh),
                                                      _vp(matches_dptr), int(capacity), _vp(counts_dptr)))


def _frontend_search_by_projection(self, params, cam, f1_frames, f2_frames, Tcw1, f1_matched_dptr, mp_valid_dptr, mp_xyz_dptr,
                                   sf1, matches_dptr, capacity, counts_dptr):
    """Batched device-resident searchByProjection(F1,F2); Tcw1: [n_pairs,4,4] float32 host array."""
    f1 = np.ascontiguousarray(f1_frames, np.int32)
    f2 = np.ascontiguousarray(f2_frames, np.int32)
    T = np.ascontiguousarray(Tcw1, np.float32).reshape(len(f1), 16)
    sf1 = np.ascontiguousarray(sf1, np.float32)
    self.ctx.check(lib().tb_frontend_search_by_projection(self.h, C.byref(params), C.byref(cam), len(f1), _vp(f1), _vp(f2), _vp(T),
                                                          _vp(f1_matched_dptr), _vp(mp_valid_dptr), _vp(mp_xyz_dptr), _vp(sf1),
                                                          _vp(matches_dptr), int(capacity), _vp(counts_dptr)))


Frontend.search_by_projection = _frontend_search_by_projection


def _frontend_set_bow(self, dvoc, n_frames, levelsup, node_dptr, word_dptr=None, weight_dptr=None):
    """Frame::SetBow for every frame of the last extract; outputs stay on the device ([frame][kp_capacity])."""
    self.ctx.check(lib().tb_frontend_set_bow(self.h, dvoc.h, int(n_frames), int(levelsup), _vp(node_dptr), _vp(word_dptr), _vp(weight_dptr)))


def _frontend_search_by_bow(self, params, f1_frames, f2_frames, node_dptr, mp_valid_dptr, map_point_only, matches_dptr, capacity, counts_dptr):
    f1 = np.ascontiguousarray(f1_frames, np.int32)
    f2 = np.ascontiguousarray(f2_frames, np.int32)
    self.ctx.check(lib().tb_frontend_search_by_bow(self.h, C.byref(params), len(f1), _vp(f1), _vp(f2), _vp(node_dptr), _vp(mp_valid_dptr),
                                                   int(bool(map_point_only)), _vp(matches_dptr), int(capacity), _vp(counts_dptr)))


Frontend.set_bow = _frontend_set_bow
Frontend.search_by_bow = _frontend_search_by_bow


def _frontend_upload_frames_undistort(self, und, first, frames):
    """Raw (distorted) host frames [n, h, w] -> remap -> level 0 of frames first.. (test_vo.cpp:178 before the Frame ctor)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    self.ctx.check(lib().tb_frontend_upload_frames_undistort(self.h, und.h, int(first), n, _vp(frames), w, h, C.c_size_t(frames.strides[1])))


def _frontend_undistort_device(self, und, first, n, raw_dptr, raw_w, raw_h, raw_pitch, raw_frame_stride):
    self.ctx.check(lib().tb_frontend_undistort_device(self.h, und.h, int(first), int(n), C.c_void_p(raw_dptr), int(raw_w), int(raw_h),
                                                      C.c_size_t(raw_pitch), C.c_size_t(raw_frame_stride)))


Frontend.upload_frames_undistort = _frontend_upload_frames_undistort
Frontend.undistort_device = _frontend_undistort_device


def _frontend_search_by_violence(self, params, f1_frames, f2_frames, min_level, max_level, search_r, matches_dptr, capacity, counts_dptr):
    f1 = np.ascontiguousarray(f1_frames, np.int32)
    f2 = np.ascontiguousarray(f2_frames, np.int32)
    self.ctx.check(lib().tb_frontend_search_by_violence(self.h, C.byref(params), len(f1), _vp(f1), _vp(f2), int(min_level), int(max_level),
                                                        C.c_float(search_r), _vp(matches_dptr), int(capacity), _vp(counts_dptr)))


Frontend.search_by_violence = _frontend_search_by_violence


def _frontend_stereo_match(self, params, left_frames, right_frames, inv_sf, matches_dptr, capacity, counts_dptr, depth_dptr=None):
    fl = np.ascontiguousarray(left_frames, np.int32)
    fr = np.ascontiguousarray(right_frames, np.int32)
    inv_sf = np.ascontiguousarray(inv_sf, np.float32)
    self.ctx.check(lib().tb_frontend_stereo_match(self.h, C.byref(params), len(fl), _vp(fl), _vp(fr), _vp(inv_sf), _vp(matches_dptr),
                                                  int(capacity), _vp(counts_dptr), _vp(depth_dptr)))


Frontend.stereo_match = _frontend_stereo_match


def _frontend_search_by_opflow(self, params, n_frames, f1_frames, f2_frames, equalized, next_pts_dptr, status_dptr, err_dptr=None):
    f1 = np.ascontiguousarray(f1_frames, np.int32)
    f2 = np.ascontiguousarray(f2_frames, np.int32)
    self.ctx.check(lib().tb_frontend_search_by_opflow(self.h, C.byref(params), int(n_frames), len(f1), _vp(f1), _vp(f2), int(bool(equalized)),
                                                      _vp(next_pts_dptr), _vp(status_dptr), _vp(err_dptr)))


Frontend.search_by_opflow = _frontend_search_by_opflow


def _frontend_reject_with_f(self, f2_frames, next_pts_dptr, status_dptr, F_dptr=None, iters_dptr=None, params=None):
    """rejectWithF behind search_by_opflow / stereo_depth_by_opflow: status is updated in place on the device."""
    params = params or RansacParams()
    f2 = np.ascontiguousarray(f2_frames, np.int32)
    self.ctx.check(lib().tb_frontend_reject_with_f(self.h, C.byref(params), len(f2), _vp(f2), _vp(next_pts_dptr), _vp(status_dptr),
                                                   _vp(F_dptr), _vp(iters_dptr)))


Frontend.reject_with_f = _frontend_reject_with_f


def _frontend_undistort_points(self, cam, n_frames, out_dptr):
    self.ctx.check(lib().tb_frontend_undistort_points(self.h, C.byref(cam), int(n_frames), _vp(out_dptr)))


Frontend.undistort_points = _frontend_undistort_points


def _frontend_stereo_depth_by_opflow(self, params, bf, n_frames, current_frames, stereo_frames, equalized, depth_dptr, next_pts_dptr, status_dptr):
    fc = np.ascontiguousarray(current_frames, np.int32)
    fs = np.ascontiguousarray(stereo_frames, np.int32)
    self.ctx.check(lib().tb_frontend_stereo_depth_by_opflow(self.h, C.byref(params), C.c_float(bf), int(n_frames), len(fc), _vp(fc), _vp(fs),
                                                            int(bool(equalized)), _vp(depth_dptr), _vp(next_pts_dptr), _vp(status_dptr)))


Frontend.stereo_depth_by_opflow = _frontend_stereo_depth_by_opflow
