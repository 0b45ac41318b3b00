/* tb_oracle.h — CPU ORACLE for the trackingBench-SLAM feature-tracking front end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it, and
 * only as the checker (or the timed CPU baseline), never as the thing shipped or measured as
 * "ours".  The product (trackingbench_slam_b200/csrc, include/) never links or calls this.
 *
 * It is a plain-C restatement of the reference algorithm (file:line cited per function, all
 * paths relative to the reference repo root) plus restatements of the OpenCV primitives the
 * reference calls (cv::resize, cv::FAST, cv::GaussianBlur, cv::fastAtan2, cv::BFMatcher) — those
 * live in an un-vendored dependency (OpenCV; the reference's prebuilt libDBoW2.so was linked
 * against 3.3, the only OpenCV available here is cv2 4.13.0).
 *
 * PARITY PINNING: the reference's own tests hold no golden vectors (they are imshow demos), so
 * the oracle is pinned by (a) tests/test_oracle_cv2.py: every OpenCV primitive bit-exact against cv2
 * 4.13.0 on images and random data (calcOpticalFlowPyrLK within a stated tolerance); (b) reference-run
 * output: oracle/_ref = the reference's OWN sources compiled here from /root/reference —
 * ORBextractor.cpp (libtb_ref.so, tests/test_oracle_golden.py), the vendored DBoW2 (libtb_ref_bow.so,
 * tests/test_oracle_bow.py), and Frame.cpp / MapPoint.cpp / Map.cpp / CameraModel.cpp /
 * FASTextractor.cpp unmodified plus matcher.cpp's hot-path member functions and LocalBA.cpp's
 * block-commented stereo matcher by line range (libtb_ref_matcher.so, tests/test_oracle_ref_matcher.py);
 * (c) committed golden vectors in tests/golden written from (b) by scripts/make_golden*.py.
 * Deliberate pins, each stated in DESIGN.md section 7: the octree sort comparator, Eigen's 3-term
 * reduction order, the Map's iteration order, and the FAST-10 detector of FASTExtractor (uzh-rpg/fast is
 * absent from the reference tree: that detector alone is PARITY-UNPINNED).
 */
#ifndef TB_ORACLE_H
#define TB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint-shaped POD (7 x 4 bytes). */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} tbo_kp;

/* cv::DMatch-shaped POD. */
typedef struct {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} tbo_match;

/* ---- Frame: scale tables + pyramid (src/types/Frame.cpp:18-29, 414-427) ---- */
void tbo_scale_factors(int n_levels, float scale, float *sf);
void tbo_level_size(int w0, int h0, float sf_level, int *w, int *h);
void tbo_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                          uint8_t *dst, int dw, int dh, int dstride);

/* ---- OpenCV primitives used by ORBextractor.cpp ---- */
/* cv::FAST(img, kps, T, nonmax) TYPE_9_16 on a w x h ROI; returns count (writes at most cap). */
int tbo_fast9(const uint8_t *img, int stride, int w, int h, int T, int nonmax, tbo_kp *out, int cap);
/* cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on 8U (OpenCV >= 3.4.2 fixed-point path). */
void tbo_gauss7x7(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
float tbo_fast_atan2(float y, float x);

/* ---- ORBExtractor (src/extractors/ORBextractor.cpp) ---- */
void tbo_orb_budget(int n_levels, float factor, int target, int *per_level); /* :919-930 */
/* operator() (:906-978) when budget_in == NULL; AddPoints (:840-904) when budget_in != NULL
 * (the reference reuses mnFeaturesPerLevel from the previous operator() call).
 * levels[l] is a w[l] x h[l] u8 image with row stride stride[l]. Returns n keypoints (may exceed
 * cap; only cap are written). desc is n x 32. n_per_level (optional) receives per-level counts. */
int tbo_orb_extract(const uint8_t *const *levels, const int *w, const int *h, const int *stride,
                    int n_levels, const float *sf, int target, float initTh, float minTh,
                    const tbo_kp *exit_kps, int n_exit, const int *budget_in,
                    tbo_kp *kps, uint8_t *desc, int cap, int *n_per_level);
/* Intermediate stages, exposed for stage-by-stage parity tests. */
int tbo_cell_candidates(const uint8_t *img, int w, int h, int stride, float initTh, float minTh,
                        tbo_kp *out, int cap); /* :743-804, coords relative to (16,16) */
int tbo_distribute_octree(const tbo_kp *cand, int n_cand, const tbo_kp *exit_kps, int n_exit,
                          int minX, int maxX, int minY, int maxY, int N,
                          tbo_kp *out, int cap); /* :494-733 */
float tbo_ic_angle(const uint8_t *img, int stride, int cx, int cy); /* :17-44 */
void tbo_rbrief(const uint8_t *blur, int stride, int cx, int cy, float angle_deg, uint8_t *desc32); /* :48-87 */

/* ---- Matcher (src/matchers/matcher.cpp) ---- */
int tbo_hamming256(const uint8_t *a, const uint8_t *b); /* :793-808 */
/* BFMatcher(NORM_HAMMING, crossCheck=true).match (ctor :20, call :207): returns count. */
int tbo_bf_cross(const uint8_t *d1, int n1, const uint8_t *d2, int n2, tbo_match *out);
/* best / second-best / argbest per query row (lowest index on ties). */
void tbo_knn2(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *idx, int32_t *best, int32_t *second);
/* searchByBF all-levels branch (:168-228). */
int tbo_search_by_bf(const uint8_t *d1, int n1, const uint8_t *d2, int n2, float ratio, float minTh, tbo_match *out);

/* Frame search grid (src/types/Frame.cpp:30-31,187-265). Opaque handle. */
typedef struct tbo_grid tbo_grid;
tbo_grid *tbo_grid_build(const tbo_kp *kps, int n, int img_cols, int img_rows);
void tbo_grid_free(tbo_grid *g);
int tbo_features_in_area(const tbo_grid *g, float x, float y, float r, int minLevel, int maxLevel,
                         int32_t *out, int cap);

typedef struct {
    int TH_LOW, TH_HIGH, HISTO_LENGTH, checkOrientation;
    float nRatio;
} tbo_matcher_params; /* matcher.h:23-27 */

/* searchByViolence (:299-395): F1 keys/desc vs grid of F2. */
int tbo_search_by_violence(const tbo_matcher_params *p, const tbo_kp *k1, const uint8_t *d1, int n1,
                           const tbo_kp *k2, const uint8_t *d2, int n2, int cols2, int rows2,
                           int min_level, int max_level, float search_r, tbo_match *out);

/* ---- frame ingest (rank-4 "next" row of SURVEY.md 8f): undistortion remap and CLAHE ----
 * Arithmetic lives in OpenCV (un-vendored dependency of the reference); restated from its published algorithm and
 * pinned bit-exact against cv2 4.13 in tests/test_oracle_cv2.py. */
/* cv::initUndistortRectifyMap(K, D, Mat(), K, size, CV_32F, map1, map2) as PinholeCamera's constructor calls it
 * (src/camera/CameraModel.cpp:15-18): K, D are float matrices promoted to double; map1 = x, map2 = y. */
void tbo_init_undistort_rectify_map(float fx, float fy, float cx, float cy, const float d[5], int w, int h,
                                    float *map1, float *map2);
/* PinholeCamera::UndistortPoint (CameraModel.cpp:104-121) = cv::undistortPoints(src, dst, K, D, Mat(), K): 5 fixed-point
 * iterations in double (radtan, k1 k2 p1 p2 k3), re-projection with K; identity when |d[0]| <= 1e-7 (:9). Pinned
 * bit-exact against cv2 4.13 in tests/test_oracle_cv2.py. xy_in/xy_out: n x 2 floats. */
void tbo_undistort_points(float fx, float fy, float cx, float cy, const float d[5], const float *xy_in, int n, float *xy_out);
/* cv::remap(src, dst, map1, map2, INTER_LINEAR) with float maps, BORDER_CONSTANT 0 (PinholeCamera::UndistortImage,
 * CameraModel.cpp:96-102): 5 fractional bits, weights (32-fy)(32-fx)*32 of 2^15, (sum + 2^14) >> 15. */
void tbo_remap_linear_u8(const uint8_t *src, int sw, int sh, int sstride, const float *map1, const float *map2,
                         int dw, int dh, uint8_t *dst, int dstride);
/* cv::createCLAHE(clip, Size(tiles_x, tiles_y))->apply (Frame::Equalize, src/types/Frame.cpp:453-458 with 3.0, 8x8). */
void tbo_clahe_u8(const uint8_t *src, int w, int h, int stride, double clip_limit, int tiles_x, int tiles_y,
                  uint8_t *dst, int dstride);

/* ---- bag of words (rank-2 "next" row of SURVEY.md 8f) ----
 * Flattened DBoW2 vocabulary in node-id order as loadFromTextFile builds it (third_part/DBoW2/DBoW2/
 * TemplatedVocabulary.h:1338-1420): node 0 = root, node i+1 = line i of the file, children of a node in
 * ascending id, word ids handed out to leaves in file order. */
typedef struct {
    int n_nodes, L;              /* L = m_L (depth levels) */
    const int32_t *parent;       /* [n_nodes], parent[0] = -1 */
    const int32_t *child_start;  /* [n_nodes+1] CSR into children */
    const int32_t *children;     /* node ids */
    const int32_t *word_id;      /* [n_nodes], -1 for inner nodes */
    const uint8_t *desc;         /* [n_nodes*32] (row 0 unused) */
    const double *weight;        /* [n_nodes] */
} tbo_vocab;
/* TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup) (:1218-1259) for n features.
 * node_id = the node visited at level L-levelsup (0 if that level <= 0).  If a leaf is reached above that level
 * the reference leaves *nid unset (UB); here node_id = that leaf. */
void tbo_bow_transform(const tbo_vocab *v, const uint8_t *desc, int n, int levelsup,
                       int32_t *word_id, int32_t *node_id, double *weight);
/* Matcher::searchByBow (matcher.cpp:619-721).  node1/node2 = FeatureVector node per feature (-1 = feature not
 * in the FeatureVector, i.e. stopped word); buckets iterate in ascending node id, members in feature order. */
int tbo_search_by_bow(const tbo_matcher_params *p, const tbo_kp *k1, const uint8_t *d1, const int32_t *node1, int n1,
                      const tbo_kp *k2, const uint8_t *d2, const int32_t *node2, int n2,
                      const uint8_t *mp_valid2, int map_point_only, tbo_match *out);

/* ---- pyramidal Lucas-Kanade flow (second half of SURVEY.md 8f rank 3) ----
 * Matcher::searchByOPFlow (matcher.cpp:724-768) calls cv::calcOpticalFlowPyrLK(img2, img1, keys, pts, status, err,
 * Size(21,21), 3) (:745).  The arithmetic lives in OpenCV (un-vendored dependency): restated from its published
 * algorithm (video/lkpyramid.cpp: pyrDown pyramid with REFLECT_101 border, Scharr derivatives with zero border,
 * 14-bit fixed-point bilinear patches, iterative 2x2 solve) and pinned against cv2 4.13 in tests/test_oracle_cv2.py:
 * pyrDown bit-exact; points within a stated tolerance (cv2's SIMD build accumulates the 2x2 system in float lanes, this
 * restatement in exact integers, so borderline iteration stops may differ). */
void tbo_pyrdown_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride); /* dst (w+1)/2 x (h+1)/2 */
typedef struct {
    int32_t win;       /* 21 (square window) */
    int32_t max_level; /* 3 */
    int32_t max_iter;  /* 30 (TermCriteria default) */
    float eps;         /* 0.01 */
    float min_eig;     /* 1e-4 */
} tbo_lk_params;
/* prev_xy/next_xy: n x 2 floats; status n bytes; err n floats (may be NULL). Returns the number of pyramid levels - 1. */
int tbo_lk_flow(const tbo_lk_params *p, const uint8_t *prev, const uint8_t *next, int w, int h, int stride,
                const float *prev_xy, int n, float *next_xy, uint8_t *status, float *err);

/* ---- ORB stereo row-band matcher (rank-3 "next" row of SURVEY.md 8f) ----
 * Restates the block-commented body of LocalBA::AddMapPointsByStereo (src/mapping/LocalBA.cpp:69-288) as its text
 * stands: the lines inside the block that carry their own // (SAD sub-pixel refinement :175-238, median filter
 * :241-253) are not part of it.  Parity unpinned: the block is dead code in the reference (never compiled), so there is
 * no reference output to pin against; tests pin this C restatement against an independent numpy restatement.
 * current (left) frame = kL/dL, stereo (right) frame = kR/dR; inv_sf = Frame::GetInverseScaleFactors();
 * n_rows = rows of pyramid level 0.  Row table entries outside [0, n_rows) are dropped (the reference would write out
 * of bounds), left keys with (size_t)y >= n_rows are skipped.  depth[nL] gets bf/|xL-xR| or -1; returns the number of
 * matches (queryIdx = iL, trainIdx = best iR, distance = Hamming), in iL order. */
typedef struct {
    float bf;          /* baseline * fx */
    float band;        /* 2.f   (:87  r = 2.f * inverse scale factor of the right key's octave) */
    float min_u, max_u;/* 0, 100000 (:117-118) */
    int32_t th_dist;   /* 30    (:148) */
    int32_t octave_tol;/* 1     (:131) */
} tbo_stereo_params;
int tbo_stereo_rowband(const tbo_stereo_params *p, const tbo_kp *kL, const uint8_t *dL, int nL,
                       const tbo_kp *kR, const uint8_t *dR, int nR, const float *inv_sf, int n_levels, int n_rows,
                       tbo_match *out, float *depth);

typedef struct {
    float fx, fy, cx, cy;
    int width, height;
    float d[5]; /* radtan distortion, used iff |d[0]| > 1e-7 */
} tbo_pinhole; /* CameraModel.cpp:63-94; CameraModel.h:33-39 */

/* searchByProjection(F1,F2) (:405-531). Map points of F2 flattened: mp_valid[i2] (non-null &&
 * !isBad), mp_xyz[3*i2], mp_desc[32*i2]. f1_matched[i1] != 0 means F1's map point i1 has
 * Observations()>0 (skipped). Tcw1 = row-major 4x4 pose of F1. sf1 = F1 scale factors. */
int tbo_search_by_projection(const tbo_matcher_params *p, const tbo_pinhole *cam, const float *Tcw1,
                             const tbo_kp *k1, const uint8_t *d1, int n1, const uint8_t *f1_matched,
                             const float *sf1, int n_levels1, int cols1, int rows1,
                             const tbo_kp *k2, int n2, const uint8_t *mp_valid, const float *mp_xyz,
                             const uint8_t *mp_desc, tbo_match *out);

/* searchByProjection(map,F1,r) (:539-617) with Frame::IsInFrustum (Frame.cpp:370-412). Map points
 * given in iteration order (the reference iterates a std::set<shared_ptr> = pointer order; the
 * oracle defines the order as the array order). normals: 3 floats each; min/max distance per
 * MapPoint.cpp:207-217 are constants 1 / 1000. */
int tbo_search_by_projection_map(const tbo_matcher_params *p, const tbo_pinhole *cam, const float *Tcw1,
                                 const tbo_kp *k1, const uint8_t *d1, int n1, const uint8_t *f1_matched,
                                 const float *sf1, int cols1, int rows1,
                                 int n_mp, const uint8_t *mp_bad, const float *mp_xyz, const float *mp_normal,
                                 const uint8_t *mp_desc, float radio, tbo_match *out);

void tbo_three_maxima(const int *hist_sizes, int L, int *ind1, int *ind2, int *ind3); /* :810-851 */

/* FASTExtractor::shiTomasiScore (src/extractors/FASTextractor.cpp:87-127). */
float tbo_shi_tomasi(const uint8_t *img, int w, int h, int stride, int u, int v);

/* FASTExtractor::operator() / AddPoints (src/extractors/FASTextractor.cpp:8-80, :129-150).
 * The FAST-10 detector / score / 3x3 non-max suppression come from uzh-rpg/fast (third_part/fast_lib,
 * NOT in the reference tree, no version pin): restated from the published algorithm (Rosten's FAST-10:
 * 10 contiguous ring pixels brighter than p+b or darker than p-b; score = largest b for which the pixel
 * is still a corner = m10-1; nonmax keeps a corner unless an 8-neighbour corner has a score >= its own).
 * PARITY UNPINNED for this function (no reference test calls it, test_frame.cpp:48-85 is commented out).
 * occupied: grid_occupancy_ after AddPoints marked the exit points (NULL = operator() with reset).
 * Returns the number of keypoints, or -1 where the reference would throw std::out_of_range
 * (gridPoints.at(k) with k >= nFeatures). */
int tbo_fast10(const uint8_t *img, int w, int h, int stride, int barrier, int32_t *xy, int32_t *scores, int cap);
int tbo_fast_extract(const uint8_t *const *levels, const int *w, const int *h, const int *stride, int n_levels,
                     const float *inv_sf, int nFeatures, float threshold, const uint8_t *occupied,
                     tbo_kp *out, int cap);

/* Matcher::rejectWithF (src/matchers/matcher.cpp:853-881): the tracks with status != 0 go through
 * cv::findFundamentalMat(cur, last, cv::FM_RANSAC, 1.0, 0.99, mask) and the outliers' status is cleared.  OpenCV is an
 * un-vendored dependency: its RANSAC is restated here (calib3d fundam.cpp / ptsetreg.cpp, 4.x: RNG((uint64)-1), getSubset
 * with the collinearity check, run7Point, solveCubic, the symmetric epipolar error in double rounded to float,
 * RANSACUpdateNumIters) and pinned to cv2 4.13 by tests/test_oracle_cv2.py (inlier masks identical).  The two null vectors of
 * the 7x9 system come from a Jacobi eigen-decomposition of A^T A instead of OpenCV's SVD: the models are the members of that
 * pencil with det = 0, which do not depend on the basis.  Fewer than 15 valid tracks (OpenCV switches to LMedS below that)
 * or n <= 8 (the reference does not call findFundamentalMat) leave status unchanged and return 0.
 * Returns the number of RANSAC iterations run; F9 (may be NULL) receives the best model. */
int tbo_reject_with_f(const float *cur_xy, const float *last_xy, int n, uint8_t *status, double threshold, double confidence,
                      int max_iters, double *F9);

#ifdef __cplusplus
}
#endif
#endif
