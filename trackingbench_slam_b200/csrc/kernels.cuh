// kernels.cuh — argument structs and host launchers shared between the kernel files and frontend.cu.
#pragma once
#include "common.cuh"

struct ResizeTab { int32_t ofs; int16_t c0, c1; };
void tb_build_resize_tab(int sn, int dn, bool is_x, std::vector<ResizeTab> &tab, std::vector<int32_t> &ofs1);
bool tb_resize_is_area2(int sw, int sh, int dw, int dh);
bool tb_resize_span_ok(const std::vector<ResizeTab> &tx, const std::vector<int32_t> &x1);
bool tb_resize_span8_ok(const std::vector<ResizeTab> &tx, const std::vector<int32_t> &x1);
void tb_resize_box(const std::vector<ResizeTab> &tx, const std::vector<int32_t> &x1, const std::vector<ResizeTab> &ty,
                   const std::vector<int32_t> &y1, int *bp, int *bh);
int tb_launch_resize(tb_ctx *ctx, const uint8_t *src, int spitch, size_t sframe, uint8_t *dst, int dpitch,
                     size_t dframe, int dw, int dh, const int4 *xtab, const int4 *ytab, bool area2, int n_frames, int bp, int bh,
                     const CUtensorMap *tmap, bool span8);
bool tb_resize_tensor_map(CUtensorMap *tm, const uint8_t *src, int w, int h, int n_frames, int pitch, size_t frame_stride, int box_rows);
// tile pitch of fast9_cells in shared memory (compile-time: ring and neighbour reads use immediate offsets): 8-byte aligned
// origin (<= 7) + up to 256 interior columns + 6 halo, rounded to 16. The E-side word of the last prefilter group reads the
// first bytes of the next row: those lanes belong to columns outside the interior and are masked.
#ifndef TB_FAST_TPITCH
#define TB_FAST_TPITCH 272
#endif
#ifndef TB_FAST_NT
#define TB_FAST_NT 128       // threads per fast9_cells CTA: four warps on a 256-pixel-wide tile (twice the pixels per warp of
                             // eight warps amortise the per-warp fixed work: set-up, queue flush, NMS rounds)
#endif
#ifndef TB_FAST_CTAS
#define TB_FAST_CTAS 8   // resident CTAs per SM the shared-memory layout and the register cap aim for
#endif
struct FastArgs {
    const uint8_t *img; int pitch; size_t frame_stride;
    int w, h;
    int nCols, nRows, wCell, hCell, maxBX, maxBY;
    int G, tpitch, scw;
    int list_cap;      // entries of the per-tile corner list (more corners than that: the NMS scans the score map)
    uint32_t magic_groups, magic_w8, magic_wcell;   // 2^32/n rounded up for the full-width tile (0 when n == 1)
    int initTh, minTh;
    uint32_t one;      // 1 (a multiplier ptxas cannot fold: keeps some adds on the FMA pipe)
    uint32_t *cand; int cand_cap;
    int *n_cand; int n_cand_stride;
    int *overflow;
};
size_t tb_fast_smem_bytes(int hCell, int scw, int tpitch, int list_cap);
int tb_launch_fast(tb_ctx *ctx, const FastArgs &a, int nColsRun, int nRowsRun, int n_frames, size_t smem);
struct OctArgs {
    const uint32_t *cand; int cand_cap;
    const int *n_cand; int n_stride;
    uint16_t *nid;
    int N;
    int W, H;
    int nIni; float hX;
    int wCell, hCell;
    const float2 *exit_xy; int n_exit;
    int16_t *enid;
    uint32_t *kept; int kept_cap;
    int *n_kept;
    int *overflow;
    int node_cap;
};
size_t tb_octree_smem_bytes(int node_cap);
int tb_launch_octree(tb_ctx *ctx, const OctArgs *levels, int n_levels, int n_frames, cudaStream_t stream);
int tb_launch_gauss(tb_ctx *ctx, cudaStream_t stream, const uint8_t *src, uint8_t *dst, int w, int h, int pitch, size_t frame_stride, int n_frames);
struct DescLevel {
    const uint8_t *raw, *blur;
    int pitch; size_t frame_stride;
    int w, h;
    const uint32_t *kept; int kept_cap;
    float sf; float size;
};
struct DescArgs {
    DescLevel lv[TB_MAX_LEVELS];
    int n_levels;
    int level_base;    // first level of this launch (blockIdx.y counts from it); the launch holding the last level writes n_kp
    const int *n_kept;
    tb_keypoint *kps; uint8_t *desc; int kp_cap;
    int *n_kp;
    int *overflow;
};
int tb_launch_describe(tb_ctx *ctx, const DescArgs &a, int max_kept_per_level, int n_frames, int level_base, int n_levels_launch, cudaStream_t stream);
bool tb_describe_fused();   // blur on demand inside the describe kernel (default) instead of dense Gaussian levels
// one brute-force matching problem: nq query rows against nt train rows of 32-byte descriptors (device pointers)
struct PairDesc {
    const uint8_t *q, *t;
    int nq, nt;
};
#define TB_TC_MIN_ROWS 256   // single-pair searchByBF takes the tensor-core kernel from this many rows on (one full 256-row CTA)
int tb_match_knn_tc(tb_ctx *ctx, int n_pairs, const PairDesc *d_pairs, int kp_cap, int *best_idx, int *best_dist);
int tb_knn2_tc(tb_ctx *ctx, const PairDesc *d_pairs, int nq, int n_splits, int *idx, int *best, int *second);
int tb_knn2_tc_splits(int nq, int nt, int num_sms);
int tb_knn2_tc_merge(tb_ctx *ctx, const int *p_idx, const int *p_best, const int *p_second, int nq, int n_splits, int rows_per_split,
                     int *idx, int *best, int *second);
int tb_match_bf_batched(tb_ctx *ctx, int n_pairs, const int32_t *q_frames_host, const int32_t *t_frames_host, int n_frames_total,
                        const uint8_t *desc, size_t frame_stride, const int *counts, int kp_cap,
                        float ratio, float minTh, tb_match *matches, int capacity, int *out_counts);

struct tb_vocab;
int tb_window_set_bow_batched(tb_ctx *ctx, const tb_vocab *v, int n_frames, int levelsup, const uint8_t *desc, int kp_cap, const int *counts,
                              int32_t *node_id, int32_t *word_id, double *weight);
int tb_window_bow_batched(tb_ctx *ctx, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames, const int32_t *f2_frames, int n_frames,
                          const tb_keypoint *kps, const uint8_t *desc, int kp_cap, const int *counts, const int32_t *node,
                          const uint8_t *mp_valid, int map_point_only, tb_match *matches, int capacity, int *out_counts);
int tb_undistort_points_launch(tb_ctx *ctx, const tb_pinhole *cam, const uint8_t *pts, size_t stride, size_t slot_stride, const int *counts,
                               int cap, int n_slots, float2 *out);
int tb_flow_batched(tb_ctx *ctx, const tb_lk_params *p, const uint8_t *prev, size_t prev_pitch, size_t prev_frame_stride, int n_prev,
                    const uint8_t *next, size_t next_pitch, size_t next_frame_stride, int n_next, int w, int h,
                    int n_pairs, const int32_t *prev_idx, const int32_t *next_idx,
                    const uint8_t *pts, size_t pts_stride, size_t pts_slot_stride, const int32_t *pts_slot, const int *counts, int pts_cap,
                    int check_in_frame, float2 *out_pts, uint8_t *out_status, float *out_err);
int tb_flow_eq_reserve(tb_ctx *ctx, size_t bytes, uint8_t **out);
int tb_flow_stereo_depth(tb_ctx *ctx, int n_pairs, size_t pts_stride, int pts_cap, float bf, const float2 *tracked, const uint8_t *status, float *depth);
int tb_window_stereo_batched(tb_ctx *ctx, const tb_stereo_params *p, int n_pairs, const int32_t *left_frames, const int32_t *right_frames,
                             const tb_keypoint *kps, const uint8_t *desc, int kp_cap, const int *counts, const float *inv_sf, int n_levels,
                             int n_rows, tb_match *matches, int capacity, int *out_counts, float *depth);
int tb_window_violence_batched(tb_ctx *ctx, const tb_matcher_params *p, int n_pairs, const int32_t *f1_frames, const int32_t *f2_frames,
                               int n_frames, const tb_keypoint *kps, const uint8_t *desc, int kp_cap, const int *counts, int cols, int rows,
                               int min_level, int max_level, float search_r, tb_match *matches, int capacity, int *out_counts);
int tb_window_projection_batched(tb_ctx *ctx, const tb_matcher_params *p, const tb_pinhole *cam, int n_pairs, const int32_t *f1_frames,
                                 const int32_t *f2_frames, const float *Tcw1, int n_frames, const tb_keypoint *kps, const uint8_t *desc,
                                 int kp_cap, const int *counts, const uint8_t *f1_matched, const uint8_t *mp_valid, const float *mp_xyz,
                                 const float *sf1, int n_levels, int cols, int rows, tb_match *matches, int capacity, int *out_counts);
int tb_undistort_width(const tb_undistort *u);
int tb_undistort_height(const tb_undistort *u);
// per-device kernel attributes (dynamic shared memory opt-in); called by tb_ctx_create for its device
void tb_fast_init_device();
void tb_fund_init_device();
// rejectWithF (fundam.cu): one frame pair = tracked points, the points they came from (float2 at last + i * last_stride),
// their number (device) and the status bytes that are updated in place
struct RfPairHost { const float *cur; const uint8_t *last; int last_stride; const int32_t *count; uint8_t *status; };
int tb_reject_with_f_batched(tb_ctx *ctx, const tb_ransac_params *p, int n_pairs, const std::vector<RfPairHost> &hp, int pts_cap,
                             double *F_out, int32_t *iters_out);
void tb_octree_init_device();
void tb_pyramid_init_device();
void tb_match_tc_init_device();
void tb_describe_init_device();
void tb_window_init_device();
void tb_flow_init_device();
