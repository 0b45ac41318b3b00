/* tb_capi.h — C ABI of libtb_b200.so, the B200 (sm_100a) implementation of the
 * trackingBench-SLAM feature-tracking front end (pyramid -> FAST/ORB extract -> Hamming match).
 *
 * This is the drop-in boundary: plain pointers, sizes and POD structs; no C++/torch types; no
 * exceptions cross it.  Everything above it (the C++ facade classes in include/types,
 * include/extractors, include/matchers that keep the reference's names, and the Python ctypes
 * mirror in trackingbench_slam_b200/) only converts cv::Mat / cv::KeyPoint / cv::DMatch to these
 * PODs.  Each entry point cites the reference interface (file:line, relative to the reference
 * repo) whose body it replaces.
 *
 * Conventions
 *  - every function returns tb_status (0 = ok, negative = error); tb_last_error(ctx) gives text.
 *  - one tb_ctx = one CUDA device + one stream; NOT thread-safe; use one host thread per ctx.
 *    Calls are stream-ordered; functions with host outputs synchronise before returning,
 *    *_device functions do not (call tb_sync or consume on the same stream).
 *  - outputs go to caller-allocated buffers with a capacity; the true count is always reported;
 *    count > capacity => TB_E_CAPACITY (the first `capacity` items are valid).
 *  - "dptr" arguments are raw CUDA device pointers valid on the ctx's device (e.g. a torch
 *    tensor's data_ptr()); ownership stays with the caller.
 *  - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef TB_CAPI_H
#define TB_CAPI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TB_OK = 0,
    TB_E_INVALID = -1,     /* bad argument */
    TB_E_CUDA = -2,        /* CUDA runtime error (text in tb_last_error) */
    TB_E_CAPACITY = -3,    /* an output or internal buffer was too small; counts are still valid */
    TB_E_UNSUPPORTED = -4, /* shape outside what the kernels support (documented per call) */
    TB_E_NOMEM = -5
} tb_status;

#define TB_MAX_LEVELS 16
#define TB_DESC_BYTES 32

/* cv::KeyPoint layout (7 x 4 bytes): pt.x pt.y size angle response octave class_id */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} tb_keypoint;

/* cv::DMatch layout */
typedef struct {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} tb_match;

/* Matcher public parameters — include/matchers/matcher.h:23-27 */
typedef struct {
    int32_t TH_LOW, TH_HIGH, HISTO_LENGTH, checkOrientation;
    float nRatio;
} tb_matcher_params;

/* PinholeCamera — include/camera/CameraModel.h:43-87 (World2Cam src/camera/CameraModel.cpp:63-94) */
typedef struct {
    float fx, fy, cx, cy;
    int32_t width, height;
    float d[5]; /* radtan; used iff |d[0]| > 1e-7 (CameraModel.cpp:9) */
} tb_pinhole;

typedef struct tb_ctx tb_ctx;
typedef struct tb_frontend tb_frontend;

/* ------------------------------------------------------------------ context */
/* stream: a cudaStream_t to run on (e.g. torch.cuda.current_stream().cuda_stream), or NULL to
 * create an owned non-blocking stream. */
int tb_ctx_create(int device, void *stream, tb_ctx **out);
void tb_ctx_destroy(tb_ctx *ctx);
int tb_sync(tb_ctx *ctx);
const char *tb_last_error(const tb_ctx *ctx);
const char *tb_version(void);
/* number of kernel launches issued through this ctx since creation (bench.py's gpu_launches) */
int64_t tb_launch_count(const tb_ctx *ctx);
/* Per-kernel timing for the roofline report: when enabled every kernel launch is bracketed by CUDA
 * events on the ctx stream; tb_profile_report synchronises and writes "name launches total_ms" lines. */
int tb_profile_enable(tb_ctx *ctx, int on);
int tb_profile_report(tb_ctx *ctx, char *buf, size_t cap);

/* ------------------------------------------------------------------ pyramid geometry (host, pure)
 * Frame::Frame scale tables (src/types/Frame.cpp:18-29) and the level sizes of
 * Frame::ComputePyramid (src/types/Frame.cpp:414-427): sf[i] = sf[i-1]*scale (float32 chain),
 * size_i = (int(cols*sf[i]), int(rows*sf[i])) with the product taken in float32. */
int tb_pyramid_geometry(int width, int height, int n_levels, float scale,
                        float *sf /*[n_levels]*/, int32_t *level_w, int32_t *level_h);

/* ------------------------------------------------------------------ front end object
 * Owns the device buffers for a batch of up to max_batch equally-sized frames: the image pyramid,
 * the blurred pyramid, FAST candidates, octree workspace and the keypoint/descriptor outputs.
 * Level sizes are given explicitly so that pyramids built elsewhere (any vector<cv::Mat> handed
 * to ORBExtractor::operator()) can be processed too. max_target bounds `target` of later extract
 * calls (outputs hold up to tb_frontend_kp_capacity() keypoints per frame). */
int tb_frontend_create(tb_ctx *ctx, int n_levels, const int32_t *level_w, const int32_t *level_h,
                       int max_batch, int max_target, tb_frontend **out);
void tb_frontend_destroy(tb_frontend *fe);
int tb_frontend_kp_capacity(const tb_frontend *fe);

/* Device layout of level `level`: frame f starts at base + f*frame_stride, rows are `pitch` apart.
 * (Batches of >= 32 frames build levels >= 1 on a side stream of the ctx; every tb_* call orders itself behind that. Once a
 * pointer to a level >= 1 has been handed out here, the frontend builds its pyramid on the ctx's own stream instead, so that
 * the caller's kernels on that stream see complete levels.) */
int tb_frontend_level_layout(const tb_frontend *fe, int level, int blurred, void **base_dptr,
                             size_t *pitch, size_t *frame_stride);
/* H2D / D2H of one level of one frame (host row stride in bytes). */
int tb_frontend_upload_level(tb_frontend *fe, int frame, int level, const uint8_t *host, size_t host_stride);
int tb_frontend_download_level(tb_frontend *fe, int frame, int level, int blurred, uint8_t *host, size_t host_stride);
/* H2D of level 0 of n_frames frames packed back to back (frame stride = host_stride*height). */
int tb_frontend_upload_frames(tb_frontend *fe, int first_frame, int n_frames, const uint8_t *host, size_t host_stride);

/* Frame::ComputePyramid (src/types/Frame.cpp:414-427): levels 1.. from level 0 by chained
 * cv::resize INTER_LINEAR (OpenCV fixed-point semantics incl. the exact-2x INTER_AREA switch). */
int tb_frontend_build_pyramid(tb_frontend *fe, int n_frames);

/* ORBExtractor::operator() (src/extractors/ORBextractor.cpp:906-978) on frames [0,n_frames) whose
 * pyramid levels are resident in the frontend (uploaded or built). sf = the scale factors the
 * reference caller passes (Frame::GetScaleFactors). Results stay on the device.
 * ORBExtractor::AddPoints (src/extractors/ORBextractor.cpp:840-904) when exit_kps != NULL
 * (host array, shared by all frames of the call) — it reuses the per-level budget of the previous
 * tb_frontend_extract call on this frontend, as the reference reuses mnFeaturesPerLevel. */
int tb_frontend_extract(tb_frontend *fe, int n_frames, const float *sf, int target, float initTh, float minTh,
                        const tb_keypoint *exit_kps, int n_exit);

/* ORBExtractor::mnFeaturesPerLevel (include/extractors/ORBextractor.h:88; written by operator(), read by AddPoints): read it
 * out of a frontend (n_levels = 0 while no extract has run) and install it into another one of the same level count. */
int tb_frontend_get_budget(const tb_frontend *fe, int32_t *per_level /*[TB_MAX_LEVELS]*/, int *n_levels);
int tb_frontend_set_budget(tb_frontend *fe, const int32_t *per_level, int n_levels);

/* Results of the last extract. Device views: keypoints of frame f at kps + f*kp_capacity,
 * descriptors at desc + f*kp_capacity*32, counts[f]. */
int tb_frontend_results_device(const tb_frontend *fe, void **kps_dptr, void **desc_dptr, void **counts_dptr);
/* D2H of one frame's results; *n = true count. Synchronises. */
int tb_frontend_get_results(tb_frontend *fe, int frame, tb_keypoint *kps, uint8_t *desc, int capacity, int *n);
/* D2H of all counts + packed results of frames [0,n_frames): kps/desc hold n_frames*capacity slots. */
int tb_frontend_get_results_batch(tb_frontend *fe, int n_frames, tb_keypoint *kps, uint8_t *desc, int capacity, int32_t *counts);
/* Pipelined variant: only enqueues the D2H copies on the ctx stream (use pinned host buffers) and
 * returns at once; call tb_frontend_status (or tb_sync) before reading them. Lets one host thread keep
 * two ctx/frontend pairs busy so that H2D of batch k+1 overlaps the kernels of batch k. */
int tb_frontend_get_results_batch_async(tb_frontend *fe, int n_frames, tb_keypoint *kps, uint8_t *desc, int capacity, int32_t *counts);
/* Synchronises the ctx stream and reports TB_E_CAPACITY if the last extract overflowed an internal buffer. */
int tb_frontend_status(tb_frontend *fe);
/* Stage outputs for stage-by-stage parity tests: FAST candidates of (frame, level) as keypoints
 * relative to the (16,16) border in the reference's vToDistributeKeys order (ORBextractor.cpp:743-804). */
int tb_frontend_get_candidates(tb_frontend *fe, int frame, int level, tb_keypoint *out, int capacity, int *n);
int tb_frontend_get_level_counts(tb_frontend *fe, int frame, int32_t *n_per_level /*[n_levels]*/);

/* One-call host convenience used by the C++ facade ORBExtractor::operator()/AddPoints: pyramid
 * levels in host memory (as vector<cv::Mat>), results to host. */
int tb_orb_extract_host(tb_frontend *fe, const uint8_t *const *levels, const size_t *host_strides,
                        const float *sf, int target, float initTh, float minTh,
                        const tb_keypoint *exit_kps, int n_exit,
                        tb_keypoint *kps, uint8_t *desc, int capacity, int *n);

/* FASTExtractor::operator() (src/extractors/FASTextractor.cpp:8-80) when exit_kps == NULL, AddPoints
 * (:129-150) otherwise: FAST-10 (barrier 20) + 3x3 non-max + Shi-Tomasi score (:87-127), best corner per
 * grid cell over all levels, emitted in cell order if response > threshold. inv_sf = the per-level factors
 * the caller passes (coordinates are multiplied by them). The FAST-10 detector comes from a library that is
 * absent from the reference tree (uzh-rpg/fast): parity of that part is unpinned.
 * Returns TB_E_INVALID where the reference throws std::out_of_range (gridPoints.at(k), k >= nFeatures). */
int tb_fast_extract_host(tb_ctx *ctx, const uint8_t *const *levels, const int32_t *w, const int32_t *h, const size_t *host_strides,
                         int n_levels, const float *inv_sf, int n_features, float threshold,
                         const tb_keypoint *exit_kps, int n_exit, tb_keypoint *out, int capacity, int *n);

/* ------------------------------------------------------------------ matching (device-resident inputs)
 * Descriptors are rows of 32 bytes. */

/* Matcher::DescriptorDistance (src/matchers/matcher.cpp:793-808), elementwise over n pairs. */
int tb_hamming_pairs_device(tb_ctx *ctx, const void *a_dptr, const void *b_dptr, int n, void *dist_i32_dptr);

/* best / second-best / argbest (lowest index on ties) of every query row against all train rows:
 * the kernel shared by BFMatcher::match and the best/second loops (matcher.cpp:333-350,582-605). */
int tb_knn2_device(tb_ctx *ctx, const void *q_dptr, int nq, const void *t_dptr, int nt,
                   void *idx_i32_dptr, void *best_i32_dptr, void *second_i32_dptr);

/* cv::BFMatcher(NORM_HAMMING, crossCheck=true).match (matcher.cpp:20,207): ascending queryIdx. */
int tb_match_bf_cross_device(tb_ctx *ctx, const void *d1_dptr, int n1, const void *d2_dptr, int n2,
                             tb_match *out_host, int capacity, int *n);

/* Matcher::searchByBF all-levels branch (matcher.cpp:168-228): cross-check then
 * dist < fmin(ratio*min_dist, minTh). */
int tb_search_by_bf_device(tb_ctx *ctx, const void *d1_dptr, int n1, const void *d2_dptr, int n2,
                           float ratio, float minTh, tb_match *out_host, int capacity, int *n);

/* Host-buffer variants used by the C++ facade (Matcher::searchByBF / DescriptorDistance): stage the
 * descriptors to the device, run the same kernels, copy the result back. */
int tb_match_bf_cross_host(tb_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2,
                           tb_match *out, int capacity, int *n);
int tb_search_by_bf_host(tb_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2,
                         float ratio, float minTh, tb_match *out, int capacity, int *n);
int tb_hamming_pairs_host(tb_ctx *ctx, const uint8_t *a, const uint8_t *b, int n, int32_t *dist);

/* Implementation of the batched brute-force nearest-neighbour search: 0 = LOP3+POPC kernel,
 * 1 = exact +-1 int8 tcgen05 MMA formulation (hamming = (256 - dot)/2). Results are identical. */
int tb_set_match_impl(tb_ctx *ctx, int impl);

/* Batched searchByBF over frame pairs of one frontend's last extract: pair p matches frame
 * q_frames[p] (query/F1) against t_frames[p] (train/F2). Outputs on the device:
 * matches_dptr[p*capacity + i], counts_dptr[p]. No host synchronisation. */
int tb_frontend_search_by_bf(tb_frontend *fe, int n_pairs, const int32_t *q_frames, const int32_t *t_frames,
                             float ratio, float minTh, void *matches_dptr, int capacity, void *counts_i32_dptr);

/* Device-resident, batched Matcher::searchByProjection(F1,F2) (matcher.cpp:405-531) over frames of the frontend's
 * last extract (test_vo-style tracking, BASELINE config 5): pair p projects the map points attached to frame
 * f2_frames[p]'s keypoints into frame f1_frames[p] whose pose is Tcw1[16*p..] (host, row-major). Map points are
 * device arrays indexed [frame][tb_frontend_kp_capacity()]: mp_valid (u8), mp_xyz (3 x f32); a map point carries
 * the descriptor of the keypoint it was created from. f1_matched (u8, same indexing) may be NULL.
 * Outputs on the device: matches_dptr[p*capacity + i] (cv::DMatch), counts_dptr[p]. No host synchronisation
 * besides staging the small pair table. */
int tb_frontend_search_by_projection(tb_frontend *fe, const tb_matcher_params *p, const tb_pinhole *cam, int n_pairs,
                                     const int32_t *f1_frames, const int32_t *f2_frames, const float *Tcw1,
                                     const void *f1_matched_dptr, const void *mp_valid_dptr, const void *mp_xyz_dptr,
                                     const float *sf1, void *matches_dptr, int capacity, void *counts_i32_dptr);

/* Frame::AssignFeaturesToGrid + GetFeaturesInArea (src/types/Frame.cpp:187-265) as a query:
 * indices in the reference's traversal order. kps on the host. */
int tb_features_in_area(tb_ctx *ctx, const tb_keypoint *kps, int n, int img_cols, int img_rows,
                        float x, float y, float r, int minLevel, int maxLevel, int32_t *out, int capacity, int *n_out);

/* Matcher::searchByViolence (matcher.cpp:299-395), MapPointOnly=false. Host in/out. */
int tb_search_by_violence(tb_ctx *ctx, const tb_matcher_params *p,
                          const tb_keypoint *k1, const uint8_t *d1, int n1,
                          const tb_keypoint *k2, const uint8_t *d2, int n2, int cols2, int rows2,
                          int min_level, int max_level, float search_r,
                          tb_match *out, int capacity, int *n);

/* ---- frame ingest (the step in front of the pyramid in test_vo.cpp:178 and Frame::Equalize) ----
 * cv::initUndistortRectifyMap(K, D, Mat(), K, Size(width,height), CV_32F, map1, map2) exactly as PinholeCamera's
 * constructor calls it (src/camera/CameraModel.cpp:15-18; K and D are float, promoted to double): map1 = x, map2 = y,
 * both width*height floats. Host computation (one-time camera setup). */
int tb_init_undistort_rectify_map(const tb_pinhole *cam, float *map1, float *map2);
/* The maps of PinholeCamera (mUndistMap1/2, include/camera/CameraModel.h:86) converted once to fixed point on the device. */
typedef struct tb_undistort tb_undistort;
int tb_undistort_create(tb_ctx *ctx, int dst_w, int dst_h, const float *map1, const float *map2, tb_undistort **out);
void tb_undistort_destroy(tb_undistort *u);
/* PinholeCamera::UndistortImage = cv::remap(raw, rectified, map1, map2, INTER_LINEAR), BORDER_CONSTANT 0
 * (src/camera/CameraModel.cpp:96-102), for n_frames images laid out at *_frame_stride bytes. Never in place. */
int tb_undistort_image_device(tb_ctx *ctx, const tb_undistort *u, const void *src_dptr, int src_w, int src_h, size_t src_pitch,
                              size_t src_frame_stride, void *dst_dptr, size_t dst_pitch, size_t dst_frame_stride, int n_frames);
int tb_undistort_image_host(tb_ctx *ctx, const tb_undistort *u, const uint8_t *src, int src_w, int src_h, size_t src_stride,
                            uint8_t *dst, size_t dst_stride);
/* Raw (distorted) frames -> level 0 of frames [first, first+n) through the remap, from device or host memory. */
int tb_frontend_undistort_device(tb_frontend *fe, const tb_undistort *u, int first, int n, const void *raw_dptr, int raw_w, int raw_h,
                                 size_t raw_pitch, size_t raw_frame_stride);
int tb_frontend_upload_frames_undistort(tb_frontend *fe, const tb_undistort *u, int first, int n, const uint8_t *host_frames,
                                        int raw_w, int raw_h, size_t host_stride);
/* PinholeCamera::UndistortPoint (src/camera/CameraModel.cpp:104-121) = cv::undistortPoints(src, dst, K, D, Mat(), K): five
 * iterations in double, bit-exact vs OpenCV; identity when |d[0]| <= 1e-7. xy arrays are n x 2 floats (host); the device
 * form reads point i at pts_dptr + i*stride (x, y = first two floats: float2 or tb_keypoint) and writes float2.
 * tb_frontend_undistort_points = Frame::UnDistortPoints (src/types/Frame.cpp:165-171) over the keypoints of the last
 * extract: out_f32x2_dptr[frame][tb_frontend_kp_capacity()], zeros beyond a frame's count. */
int tb_undistort_points(tb_ctx *ctx, const tb_pinhole *cam, const float *xy_in, int n, float *xy_out);
int tb_undistort_points_device(tb_ctx *ctx, const tb_pinhole *cam, const void *pts_dptr, size_t stride, long long n, void *out_f32x2_dptr);
int tb_frontend_undistort_points(tb_frontend *fe, const tb_pinhole *cam, int n_frames, void *out_f32x2_dptr);
/* Frame::Equalize = cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y))->apply (src/types/Frame.cpp:453-458 uses 3.0, 8x8). */
int tb_clahe_device(tb_ctx *ctx, const void *src_dptr, int w, int h, size_t src_pitch, size_t src_frame_stride, void *dst_dptr,
                    size_t dst_pitch, size_t dst_frame_stride, int n_frames, double clip_limit, int tiles_x, int tiles_y);
int tb_clahe_host(tb_ctx *ctx, const uint8_t *src, int w, int h, size_t src_stride, uint8_t *dst, size_t dst_stride,
                  double clip_limit, int tiles_x, int tiles_y);

/* ---- bag of words ----
 * ORBVocabulary (include/types/Frame.h:22 = DBoW2::TemplatedVocabulary<FORB>) flattened in node-id order as
 * loadFromTextFile builds it (third_part/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1420): node 0 = root, node i+1 =
 * line i of the vocabulary file; parent[i] < i; is_leaf[i] = the file's leaf flag (word ids follow it in file
 * order); desc[32*i]; weight[i]. depth_levels = m_L. The tree is copied to the device once. */
typedef struct tb_vocab tb_vocab;
int tb_vocab_create(tb_ctx *ctx, int n_nodes, int depth_levels, const int32_t *parent, const uint8_t *is_leaf,
                    const uint8_t *desc, const double *weight, tb_vocab **out);
void tb_vocab_destroy(tb_vocab *v);
/* Frame::SetBow (src/types/Frame.cpp:267-270) -> TemplatedVocabulary::transform(feature, word, weight, nid, levelsup)
 * (:1218-1259) for n descriptors: word id, FeatureVector node id (level L-levelsup) and word weight per feature. A
 * feature enters the BowVector / FeatureVector only if weight > 0 (:1157, :1185). Any output may be NULL. */
int tb_bow_transform(tb_ctx *ctx, const tb_vocab *v, const uint8_t *desc, int n, int levelsup,
                     int32_t *word_id, int32_t *node_id, double *weight);
int tb_bow_transform_device(tb_ctx *ctx, const tb_vocab *v, const uint8_t *d_desc, long long n, int levelsup,
                            int32_t *d_word_id, int32_t *d_node_id, double *d_weight);
/* Matcher::searchByBow (matcher.cpp:619-721). node1/node2 = FeatureVector node of each feature, -1 = feature not in
 * the FeatureVector; mp_valid2[i2] = (F2->GetMapPoint(i2) != nullptr), only read when map_point_only. Host in/out. */
int tb_search_by_bow(tb_ctx *ctx, const tb_matcher_params *p,
                     const tb_keypoint *k1, const uint8_t *d1, const int32_t *node1, int n1,
                     const tb_keypoint *k2, const uint8_t *d2, const int32_t *node2, int n2,
                     const uint8_t *mp_valid2, int map_point_only, tb_match *out, int capacity, int *n);

/* Device-resident, batched Matcher::searchByViolence (matcher.cpp:299-395, MapPointOnly=false) over frames of the
 * frontend's last extract: pair p matches the keypoints of frame f1_frames[p] inside windows of half-size search_r on the
 * search grid of frame f2_frames[p]. Outputs on the device: matches_dptr[p*capacity + i], counts_dptr[p]. */
int tb_frontend_search_by_violence(tb_frontend *fe, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames,
                                   const int32_t *f2_frames, int min_level, int max_level, float search_r,
                                   void *matches_dptr, int capacity, void *counts_i32_dptr);

/* Device-resident, batched Frame::SetBow + Matcher::searchByBow over the frames of the frontend's last extract.
 * tb_frontend_set_bow writes the FeatureVector of every frame as node_dptr[frame][tb_frontend_kp_capacity()] (i32; -1 =
 * stopped word or slot beyond the frame's keypoint count) and optionally the per-feature word id (i32) / weight (f64).
 * tb_frontend_search_by_bow matches pair p = (f1_frames[p], f2_frames[p]); mp_valid_dptr (u8, same indexing, may be
 * NULL) is only read when map_point_only. Outputs on the device: matches_dptr[p*capacity + i], counts_dptr[p]. */
int tb_frontend_set_bow(tb_frontend *fe, const tb_vocab *voc, int n_frames, int levelsup,
                        void *node_i32_dptr, void *word_i32_dptr, void *weight_f64_dptr);
int tb_frontend_search_by_bow(tb_frontend *fe, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames, const int32_t *f2_frames,
                              const void *node_i32_dptr, const void *mp_valid_dptr, int map_point_only,
                              void *matches_dptr, int capacity, void *counts_i32_dptr);

/* ORB stereo row-band matcher: the block-commented body of LocalBA::AddMapPointsByStereo (src/mapping/LocalBA.cpp:69-288)
 * as its text stands (coarse ORB match per image row band + depth = bf/|xL-xR|; the SAD refinement and median filter carry
 * their own // inside that block and are not part of it).  Right key iR is a candidate of left key iL when
 * floor(yR-r) <= (size_t)yL <= ceil(yR+r), r = band * inv_sf[octaveR] (:84-92), |octaveR-octaveL| <= octave_tol (:131) and
 * min_u <= xR <= max_u (:136); the first candidate in index order with the smallest non-zero distance wins (:140); it is a
 * match when that distance < th_dist (:148).  Matches come out in iL order with imgIdx = -1 (cv::DMatch default). */
typedef struct {
    float bf;           /* baseline * fx */
    float band;         /* reference: 2.f */
    float min_u, max_u; /* reference: 0, 100000 */
    int32_t th_dist;    /* reference: 30 */
    int32_t octave_tol; /* reference: 1 */
} tb_stereo_params;
/* Host in/out. inv_sf = Frame::GetInverseScaleFactors() (n_levels floats), n_rows = rows of pyramid level 0;
 * depth (may be NULL) receives nL floats, -1 where there is no match. */
int tb_stereo_match(tb_ctx *ctx, const tb_stereo_params *p, const tb_keypoint *kL, const uint8_t *dL, int nL,
                    const tb_keypoint *kR, const uint8_t *dR, int nR, const float *inv_sf, int n_levels, int n_rows,
                    tb_match *out, int capacity, int *n, float *depth);
/* Device-resident, batched over frames of the frontend's last extract: pair p = (left_frames[p], right_frames[p]).
 * Outputs on the device: matches_dptr[p*capacity + i], counts_dptr[p], depth_f32_dptr[p*tb_frontend_kp_capacity() + iL]
 * (may be NULL). */
int tb_frontend_stereo_match(tb_frontend *fe, const tb_stereo_params *p, int n_pairs, const int32_t *left_frames,
                             const int32_t *right_frames, const float *inv_sf, void *matches_dptr, int capacity,
                             void *counts_i32_dptr, void *depth_f32_dptr);

/* Pyramidal Lucas-Kanade tracking = the cv::calcOpticalFlowPyrLK(img2, img1, F2->GetCVKeys(), cur_points, status, err,
 * cv::Size(21,21), 3) call of Matcher::searchByOPFlow (src/matchers/matcher.cpp:745): OpenCV's fixed-point patches and
 * stop rules; the 2x2 sums are exact integers (OpenCV's SIMD build sums them in float lanes), so results agree with
 * OpenCV within a tolerance (tests/test_oracle_cv2.py) and with the oracle bit for bit. status = 1 where the track is
 * valid; err = OpenCV's mean absolute patch difference / 32. rejectWithF (RANSAC, matcher.cpp:751-755) is tb_reject_with_f_* below. */
typedef struct {
    int32_t win;       /* reference: 21 (3..21) */
    int32_t max_level; /* reference: 3; at most 7 levels above level 0 are built (else TB_E_UNSUPPORTED) */
    int32_t max_iter;  /* OpenCV default TermCriteria: 30 */
    float eps;         /* OpenCV default: 0.01 */
    float min_eig;     /* OpenCV default minEigThreshold: 1e-4 */
} tb_lk_params;
/* Host in/out, one image pair: prev_xy/next_xy are n x 2 floats, status n bytes, err n floats (may be NULL). */
int tb_lk_flow_host(tb_ctx *ctx, const tb_lk_params *p, const uint8_t *prev, const uint8_t *next, int w, int h, size_t stride,
                    const float *prev_xy, int n, float *next_xy, uint8_t *status, float *err);
/* Device, batched: two stacks of w x h images (frame f at base + f*frame_stride; the stacks may be the same memory);
 * pair q tracks the points pts[q][0..counts[q]) (float2, pts_cap per pair, counts on the device) from image
 * prev_idx[q] of the first stack into image next_idx[q] of the second. Outputs [pair][pts_cap]. */
int tb_lk_flow_device(tb_ctx *ctx, const tb_lk_params *p, const void *prev_dptr, size_t prev_pitch, size_t prev_frame_stride, int n_prev,
                      const void *next_dptr, size_t next_pitch, size_t next_frame_stride, int n_next, int w, int h,
                      int n_pairs, const int32_t *prev_idx, const int32_t *next_idx, const void *pts_f32x2_dptr,
                      const void *counts_i32_dptr, int pts_cap, void *next_pts_f32x2_dptr, void *status_u8_dptr, void *err_f32_dptr);
/* Matcher::searchByOPFlow(F1, F2, cur_points, equalized, reject=false) (matcher.cpp:724-768; follow it with
 * tb_frontend_reject_with_f for reject = true) over frames [0, n_frames) of
 * the frontend (level 0 resident, keypoints from the last extract): pair q tracks the keypoints of frame f2_frames[q] into
 * frame f1_frames[q] (CLAHE'd first when equalized), then clears status where IsInFrame(Vector2i(pt)) fails (:747-749).
 * The reference's match list is {i,i} for every status[i] != 0. Outputs on the device, [pair][tb_frontend_kp_capacity()]. */
int tb_frontend_search_by_opflow(tb_frontend *fe, const tb_lk_params *p, int n_frames, int n_pairs, const int32_t *f1_frames,
                                 const int32_t *f2_frames, int equalized, void *next_pts_f32x2_dptr, void *status_u8_dptr,
                                 void *err_f32_dptr);

/* LocalBA::AddMapPointsByStereo's live body (src/mapping/LocalBA.cpp:46-68) without rejectWithF (run
 * tb_frontend_reject_with_f on the returned tracks and drop the depths whose status it cleared): pair q tracks the keypoints of
 * frame current_frames[q] into frame stereo_frames[q] (searchByOPFlow(stereo, current, pts, equalized, ...), :54) and writes
 * depth_f32_dptr[q][i] = bf / |pts[i].x - kp[i].x| for every valid track, -1 elsewhere (:61-65). The tracks and their status
 * are returned as well ([pair][tb_frontend_kp_capacity()]) so that the caller can run the reference's RANSAC rejection. */
int tb_frontend_stereo_depth_by_opflow(tb_frontend *fe, const tb_lk_params *p, float bf, int n_frames, int n_pairs,
                                       const int32_t *current_frames, const int32_t *stereo_frames, int equalized,
                                       void *depth_f32_dptr, void *next_pts_f32x2_dptr, void *status_u8_dptr);

/* Matcher::rejectWithF (src/matchers/matcher.cpp:853-881), the step searchByOPFlow(reject = true) runs after the IsInFrame
 * filter (:751-755; every reference call site passes reject = true): the tracks with status != 0 go through
 * cv::findFundamentalMat(cur, last, cv::FM_RANSAC, 1.0, 0.99, mask) and the outliers' status is cleared.  OpenCV's RANSAC is
 * reproduced iteration for iteration (cv::RNG((uint64)-1), getSubset with its collinearity test, the 7-point solver,
 * solveCubic, the symmetric epipolar error rounded to float, RANSACUpdateNumIters): inlier masks identical to cv2 4.13 on
 * every tested case.  Fewer than 15 tracks with status != 0 (OpenCV switches to LMedS below that) or n <= 8 (the reference
 * does not call findFundamentalMat) leave status unchanged. */
typedef struct {
    float threshold;    /* reference: 1.0 (pixels) */
    float confidence;   /* reference: 0.99 */
    int32_t max_iters;  /* OpenCV default: 1000 */
} tb_ransac_params;
/* Host in/out, one pair: cur_xy / last_xy are n x 2 floats, status n bytes (updated in place); F9 (row-major 3x3, may be
 * NULL) receives the best model, iterations (may be NULL) the number of RANSAC iterations OpenCV would have run. */
int tb_reject_with_f_host(tb_ctx *ctx, const tb_ransac_params *p, const float *cur_xy, const float *last_xy, int n, uint8_t *status,
                          double *F9, int *iterations);
/* Device, batched: [pair][pts_cap] float2 points, [pair][pts_cap] status bytes (in/out), counts[pair] on the device;
 * F_f64x9_dptr [pair][9] and iters_i32_dptr [pair] may be NULL.  One CTA per pair, no host wait. */
int tb_reject_with_f_device(tb_ctx *ctx, const tb_ransac_params *p, int n_pairs, const void *cur_pts_f32x2_dptr,
                            const void *last_pts_f32x2_dptr, const void *counts_i32_dptr, int pts_cap, void *status_u8_dptr,
                            void *F_f64x9_dptr, void *iters_i32_dptr);
/* The same behind tb_frontend_search_by_opflow / tb_frontend_stereo_depth_by_opflow: pair q's "last" points are the keypoints
 * of frame f2_frames[q] (the frame the tracks started from), next_pts / status are that call's outputs
 * ([pair][tb_frontend_kp_capacity()]). */
int tb_frontend_reject_with_f(tb_frontend *fe, const tb_ransac_params *p, int n_pairs, const int32_t *f2_frames,
                              const void *next_pts_f32x2_dptr, void *status_u8_dptr, void *F_f64x9_dptr, void *iters_i32_dptr);

/* Matcher::searchByProjection(F1,F2) (matcher.cpp:405-531). F2's map points flattened:
 * mp_valid[i2] = (GetMapPoint(i2) != null && !isBad()), mp_xyz[3*i2], mp_desc[32*i2];
 * f1_matched[i1] != 0 iff F1->GetMapPoint(i1)->Observations() > 0 (may be NULL). Tcw1 row-major 4x4. */
int tb_search_by_projection(tb_ctx *ctx, const tb_matcher_params *p, const tb_pinhole *cam, const float *Tcw1,
                            const tb_keypoint *k1, const uint8_t *d1, int n1, const uint8_t *f1_matched,
                            const float *sf1, int n_levels1, int cols1, int rows1,
                            const tb_keypoint *k2, int n2, const uint8_t *mp_valid, const float *mp_xyz,
                            const uint8_t *mp_desc, tb_match *out, int capacity, int *n);

/* Matcher::searchByProjection(map,F1,radio) (matcher.cpp:539-617) + Frame::IsInFrustum
 * (src/types/Frame.cpp:370-412). Map points in the caller's iteration order. */
int tb_search_by_projection_map(tb_ctx *ctx, const tb_matcher_params *p, const tb_pinhole *cam, const float *Tcw1,
                                const tb_keypoint *k1, const uint8_t *d1, int n1, const uint8_t *f1_matched,
                                const float *sf1, int cols1, int rows1,
                                int n_mp, const uint8_t *mp_bad, const float *mp_xyz, const float *mp_normal,
                                const uint8_t *mp_desc, float radio, tb_match *out, int capacity, int *n);

/* Matcher::ComputeThreeMaxima (matcher.cpp:810-851) — host helper kept for API completeness. */
void tb_three_maxima(const int32_t *hist_sizes, int L, int *ind1, int *ind2, int *ind3);

#ifdef __cplusplus
}
#endif
#endif /* TB_CAPI_H */
