// ref_matcher_build.cpp — builds oracle/_ref/libtb_ref_matcher.so: the reference's OWN Frame / MapPoint / Map /
// CameraModel / FASTExtractor sources compiled UNMODIFIED from where they lie under /root/reference, plus the
// hot-path member functions of src/matchers/matcher.cpp taken by LINE RANGE at build time (oracle/Makefile writes
// them to oracle/_ref/matcher_extract.inc; the rest of that file needs Sophus/boost/full Eigen and is out of scope:
// searchByNN, searchByDirect, SparseImageAlign, FeaturesAlign, Align1D/2D).  Compiled against oracle/shim (OpenCV
// names forwarding to the cv2-pinned C restatements, a fixed-size Eigen stand-in, a fast/fast.h stand-in) with
// -fno-access-control so the C wrappers below can reach private members.  ORACLE / TEST INFRASTRUCTURE ONLY.
// No reference source is copied into this repo.
//
// Deliberate pins (each is stated again in DESIGN.md):
//  * Eigen's fixed-size reduction order x0 + (x1 + x2) (oracle/shim/Eigen/Core) — Eigen itself is absent here.
//  * searchByProjection(map, F1, r) walks std::set<shared_ptr<MapPoint>> = ascending ADDRESS order (Map.cpp:71-76);
//    tbrefm_map_points_new places the MapPoints in one arena so that address order == the caller's array order.
//  * uzh-rpg/fast (FAST-10 detector of FASTExtractor) is absent: stand-in forwards to tbo_fast10, detector unpinned.
#include <algorithm>
#include <climits>
#include <cstring>
#include <memory>
#include <new>
#include <sstream>
#include <stdexcept>
#include <vector>
#include "cvshim_more.h"
#include "DBoW2/FORB.cpp"
#include "DBoW2/BowVector.cpp"
#include "DBoW2/FeatureVector.cpp"
#include "DBoW2/ScoringObject.cpp"
#include "DUtils/Random.cpp"
#include "DUtils/Timestamp.cpp"
#include "types/Frame.h"
#include "types/MapPoint.h"
#include "types/Map.h"
#include "camera/CameraModel.h"
#include "matchers/matcher.h"
#include "extractors/FASTextractor.h"

#include "src/camera/CameraModel.cpp"
#include "src/types/Frame.cpp"
#include "src/types/MapPoint.cpp"
#include "src/types/Map.cpp"
#include "src/extractors/FASTextractor.cpp"
#include "_ref/matcher_extract.inc"

// LocalBA::AddMapPointsByStereo's block-commented ORB row-band matcher (LocalBA.cpp:69-288) compiled from its own text as
// the body of a function with the live function's parameters; its final cv::imshow("Stereo", show) is the hook that
// hands the local `matches` vector out.
namespace cv { inline void tb_capture_matches(const std::vector<DMatch>& m); }
namespace TRACKING_BENCH {
static std::vector<cv::DMatch> tb_stereo_matches;
#define imshow(name, img) tb_capture_matches(matches)
static std::vector<float> tb_stereo_rowband_body(const std::shared_ptr<Frame>& current_frame, const std::shared_ptr<Frame>& stereo_frame,
                                                 const float bf, const float fx)
{
#include "_ref/stereo_extract.inc"
}
#undef imshow
}  // namespace TRACKING_BENCH
namespace cv { inline void tb_capture_matches(const std::vector<DMatch>& m) { TRACKING_BENCH::tb_stereo_matches = m; } }

using namespace TRACKING_BENCH;

typedef std::shared_ptr<Frame> FramePtr;
typedef std::shared_ptr<CameraModel> CamPtr;
typedef std::shared_ptr<Map> MapPtr;
typedef std::shared_ptr<MapPoint> MPPtr;

struct MapPointArena {
    char* mem;
    int n;
    std::vector<MPPtr> pts;  // null where the slot holds no map point
    MapPtr map;
};

static int put_matches(const std::vector<cv::DMatch>& m, tbo_match* out, int cap)
{
    const int n = (int)m.size();
    for (int i = 0; i < n && i < cap; i++) {
        tbo_match t = {m[i].queryIdx, m[i].trainIdx, m[i].imgIdx, m[i].distance};
        out[i] = t;
    }
    return n;
}

extern "C" {

// ---- PinholeCamera (CameraModel.cpp:6-19)
void* tbrefm_camera_new(int w, int h, float fx, float fy, float cx, float cy, const float* d)
{
    return new CamPtr(new PinholeCamera(w, h, fx, fy, cx, cy, d[0], d[1], d[2], d[3], d[4]));
}
void tbrefm_camera_free(void* c) { delete (CamPtr*)c; }
void tbrefm_camera_world2cam(void* c, const float* xyz, float* uv)
{
    Eigen::Vector2f r = (*(CamPtr*)c)->World2Cam(Eigen::Vector3f(xyz[0], xyz[1], xyz[2]));
    uv[0] = r[0]; uv[1] = r[1];
}
int tbrefm_camera_is_in_frame(void* c, int x, int y) { return (*(CamPtr*)c)->IsInFrame(Eigen::Vector2i(x, y)) ? 1 : 0; }
void tbrefm_camera_undistort_point(void* c, const float* uv, float* out)
{
    Eigen::Vector2f r = (*(CamPtr*)c)->UndistortPoint(Eigen::Vector2f(uv[0], uv[1]));
    out[0] = r[0]; out[1] = r[1];
}
// PinholeCamera::UndistortImage (CameraModel.cpp:96-102) into a w x h buffer
void tbrefm_camera_undistort_image(void* c, const uint8_t* raw, int w, int h, int stride, uint8_t* out)
{
    cv::Mat in(h, w, CV_8UC1, (void*)raw, (size_t)stride), o;
    (*(CamPtr*)c)->UndistortImage(in, o);
    for (int r = 0; r < o.rows; r++) std::memcpy(out + (size_t)r * o.cols, o.ptr(r), (size_t)o.cols);
}

// ---- Frame::Frame (Frame.cpp:13-33): scale tables, grid scale, ComputePyramid (:414-427)
void* tbrefm_frame_new(const uint8_t* img, int w, int h, int stride, int levels, float scale, void* cam)
{
    cv::Mat m = cv::Mat(h, w, CV_8UC1, (void*)img, (size_t)stride).clone();
    return new FramePtr(new Frame(m, 0.0, levels, scale, *(CamPtr*)cam));
}
void tbrefm_frame_free(void* f)
{
    FramePtr* fp = (FramePtr*)f;
    for (auto& k : (*fp)->GetKeys()) k->frame.reset();  // Feature::frame points back at the Frame (cycle)
    delete fp;
}
void tbrefm_frame_geometry(void* f, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2, int* w, int* h)
{
    Frame& F = **(FramePtr*)f;
    for (int i = 0; i < F.GetLevels(); i++) {
        sf[i] = F.GetScaleFactors()[i]; inv_sf[i] = F.GetInverseScaleFactors()[i];
        sigma2[i] = F.GetScaleSigmaSquares()[i]; inv_sigma2[i] = F.GetInverseScaleSigmaSquares()[i];
        w[i] = F.GetImagePyramid()[i].cols; h[i] = F.GetImagePyramid()[i].rows;
    }
}
void tbrefm_frame_level(void* f, int level, uint8_t* out)
{
    const cv::Mat& m = (*(FramePtr*)f)->GetImagePyramid()[level];
    for (int r = 0; r < m.rows; r++) std::memcpy(out + (size_t)r * m.cols, m.ptr(r), (size_t)m.cols);
}
// Frame::Equalize (Frame.cpp:453-458)
void tbrefm_frame_equalize(void* f, uint8_t* out)
{
    cv::Mat m = (*(FramePtr*)f)->Equalize();
    for (int r = 0; r < m.rows; r++) std::memcpy(out + (size_t)r * m.cols, m.ptr(r), (size_t)m.cols);
}
// Frame::SetKeys (Frame.cpp:94-116) + AssignFeaturesToGrid (:187-200)
void tbrefm_frame_set_keys(void* f, const tbo_kp* kps, const uint8_t* desc, int n, int undistort)
{
    FramePtr& F = *(FramePtr*)f;
    std::vector<cv::KeyPoint> pts;
    for (int i = 0; i < n; i++) pts.push_back(cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id));
    cv::Mat d;
    if (desc && n > 0) d = cv::Mat(n, 32, CV_8UC1, (void*)desc, 32).clone();
    F->SetKeys(pts, F, d, undistort != 0);
}
void tbrefm_frame_assign_grid(void* f) { (*(FramePtr*)f)->AssignFeaturesToGrid(); }
void tbrefm_frame_set_pose(void* f, const float* Tcw_rowmajor)
{
    Eigen::Matrix4f T;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = Tcw_rowmajor[4 * r + c];
    (*(FramePtr*)f)->SetPose(T);
}
void tbrefm_frame_camera_center(void* f, float* ow)
{
    Eigen::Vector3f o = (*(FramePtr*)f)->GetCameraCenter();
    ow[0] = o[0]; ow[1] = o[1]; ow[2] = o[2];
}
int tbrefm_frame_pos_in_grid(void* f, const tbo_kp* kp, int* px, int* py)
{
    cv::KeyPoint k(kp->x, kp->y, kp->size, kp->angle, kp->response, kp->octave, kp->class_id);
    return (*(FramePtr*)f)->PosInGrid(k, *px, *py) ? 1 : 0;
}
int tbrefm_frame_features_in_area(void* f, float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap)
{
    std::vector<size_t> v = (*(FramePtr*)f)->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}
// Frame::UnDistortPoints (Frame.cpp:165-171): px_un of every key
void tbrefm_frame_undistort_points(void* f, float* xy_out)
{
    Frame& F = **(FramePtr*)f;
    F.UnDistortPoints();
    for (size_t i = 0; i < F.GetKeys().size(); i++) { xy_out[2 * i] = F.GetKeys()[i]->px_un[0]; xy_out[2 * i + 1] = F.GetKeys()[i]->px_un[1]; }
}
// FeatureVector as the matcher reads it: either through the reference's own SetBow (vocabulary handle from
// tbrefm_vocab_load_text) or set directly from per-feature node ids (-1 = not in the FeatureVector)
void* tbrefm_vocab_load_text(const char* path)
{
    std::shared_ptr<ORBVocabulary>* v = new std::shared_ptr<ORBVocabulary>(new ORBVocabulary());
    if (!(*v)->loadFromTextFile(path) || (*v)->empty()) { delete v; return nullptr; }
    return v;
}
void tbrefm_vocab_free(void* v) { delete (std::shared_ptr<ORBVocabulary>*)v; }
void tbrefm_frame_set_bow(void* f, void* voc) { (*(FramePtr*)f)->SetBow(*(std::shared_ptr<ORBVocabulary>*)voc); }
void tbrefm_frame_set_feature_vector(void* f, const int32_t* node, int n)
{
    DBoW2::FeatureVector& fv = (*(FramePtr*)f)->GetFeatureVector();
    fv.clear();
    for (int i = 0; i < n; i++) if (node[i] >= 0) fv.addFeature((DBoW2::NodeId)node[i], (unsigned int)i);
}
void tbrefm_frame_get_feature_vector(void* f, int32_t* node, int n)
{
    for (int i = 0; i < n; i++) node[i] = -1;
    DBoW2::FeatureVector& fv = (*(FramePtr*)f)->GetFeatureVector();
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t k = 0; k < it->second.size(); k++) node[it->second[k]] = (int)it->first;
}

// ---- MapPoints: MapPoint::MapPoint (MapPoint.cpp:13-44) for every slot with valid[i] != 0, created from key i of
// `frame` (descriptor row i, normal = normalized(Pos - camera centre of `frame`)), attached to the frame
// (Frame::AddMapPoint) when attach != 0 and inserted into the arena's Map (Map::AddMapPoint).
void* tbrefm_map_points_new(void* frame, const uint8_t* valid, const float* xyz, int n, int attach)
{
    FramePtr& F = *(FramePtr*)frame;
    MapPointArena* A = new MapPointArena();
    A->mem = (char*)::operator new(sizeof(MapPoint) * (size_t)(n > 0 ? n : 1));
    A->n = n;
    A->map = MapPtr(new Map());
    A->pts.resize((size_t)n);
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;
        std::shared_ptr<Feature> ft = F->GetKey((size_t)i);
        MapPoint* raw = new (A->mem + sizeof(MapPoint) * (size_t)i)
            MapPoint(Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), A->map, F, ft, cv::Mat());
        A->pts[i] = MPPtr(raw, [](MapPoint* p) { p->~MapPoint(); });
        A->map->AddMapPoint(A->pts[i]);
        if (attach) F->AddMapPoint(A->pts[i], (size_t)i);
    }
    return A;
}
void tbrefm_map_points_free(void* a)
{
    MapPointArena* A = (MapPointArena*)a;
    A->map->clear();
    for (auto& p : A->pts) if (p) { p->mpMap.reset(); p->mpRefFeature.reset(); p->mFeatures.clear(); }
    A->map.reset();
    A->pts.clear();
    ::operator delete(A->mem);
    delete A;
}
void tbrefm_map_point_set_bad(void* a, int i) { MapPointArena* A = (MapPointArena*)a; if (A->pts[i]) A->pts[i]->mbBad = true; }
void tbrefm_map_point_set_obs(void* a, int i, int n_obs) { MapPointArena* A = (MapPointArena*)a; if (A->pts[i]) A->pts[i]->nObs = n_obs; }
void tbrefm_map_point_get(void* a, int i, float* xyz, float* normal, uint8_t* desc)
{
    MapPointArena* A = (MapPointArena*)a;
    Eigen::Vector3f p = A->pts[i]->GetWorldPos(), nn = A->pts[i]->GetNormal();
    for (int k = 0; k < 3; k++) { xyz[k] = p[k]; normal[k] = nn[k]; }
    cv::Mat d = A->pts[i]->GetDescriptor();
    std::memcpy(desc, d.data, 32);
}
// attach map point i of the arena to key idx of another frame (Frame::AddMapPoint), as tracking does after a match
void tbrefm_frame_attach_map_point(void* f, int idx, void* a, int i)
{
    MapPointArena* A = (MapPointArena*)a;
    (*(FramePtr*)f)->AddMapPoint(A->pts[i], (size_t)idx);
}
// Frame::IsInFrustum (Frame.cpp:370-412) for map point i; returns 0/1; out = level, x, y, viewCos
int tbrefm_frame_is_in_frustum(void* f, void* a, int i, float view_cos_limit, int* level, float* x, float* y, float* view_cos)
{
    MapPointArena* A = (MapPointArena*)a;
    float v = view_cos_limit;
    const bool ok = (*(FramePtr*)f)->IsInFrustum(A->pts[i], *level, *x, *y, v);
    *view_cos = v;
    return ok ? 1 : 0;
}

// ---- Matcher (matcher.cpp line ranges)
void* tbrefm_matcher_new() { return new Matcher(); }
void tbrefm_matcher_free(void* m) { delete (Matcher*)m; }
void tbrefm_matcher_set(void* m, int low, int high, int histo, int check, float ratio)
{
    ((Matcher*)m)->setProjectionParam(low, high, histo, check != 0, ratio);  // the three set*Param bundles are identical (matcher.h:48-90)
}
int tbrefm_search_by_bf(void* m, void* f1, void* f2, int minLevel, int maxLevel, float ratio, float minTh, tbo_match* out, int cap)
{
    return put_matches(((Matcher*)m)->searchByBF(*(FramePtr*)f1, *(FramePtr*)f2, minLevel, maxLevel, ratio, minTh, false), out, cap);
}
int tbrefm_search_by_violence(void* m, void* f1, void* f2, int min_level, int max_level, float r, int mp_only, tbo_match* out, int cap)
{
    return put_matches(((Matcher*)m)->searchByViolence(*(FramePtr*)f1, *(FramePtr*)f2, min_level, max_level, r, mp_only != 0), out, cap);
}
int tbrefm_search_by_projection(void* m, void* f1, void* f2, tbo_match* out, int cap)
{
    return put_matches(((Matcher*)m)->searchByProjection(*(FramePtr*)f1, *(FramePtr*)f2), out, cap);
}
int tbrefm_search_by_projection_map(void* m, void* arena, void* f1, float radio, tbo_match* out, int cap)
{
    return put_matches(((Matcher*)m)->searchByProjection(((MapPointArena*)arena)->map, *(FramePtr*)f1, radio), out, cap);
}
int tbrefm_search_by_bow(void* m, void* f1, void* f2, int mp_only, tbo_match* out, int cap)
{
    return put_matches(((Matcher*)m)->searchByBow(*(FramePtr*)f1, *(FramePtr*)f2, mp_only != 0), out, cap);
}
// searchByOPFlow(F1, F2, cur_points, equalized, reject = false) (matcher.cpp:724-768)
int tbrefm_search_by_opflow(void* m, void* f1, void* f2, int equalized, float* cur_xy, tbo_match* out, int cap)
{
    std::vector<cv::Point2f> cur;
    std::vector<cv::DMatch> r = ((Matcher*)m)->searchByOPFlow(*(FramePtr*)f1, *(FramePtr*)f2, cur, equalized != 0, false, false);
    for (size_t i = 0; i < cur.size(); i++) { cur_xy[2 * i] = cur[i].x; cur_xy[2 * i + 1] = cur[i].y; }
    return put_matches(r, out, cap);
}
int tbrefm_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    return Matcher::DescriptorDistance(cv::Mat(1, 32, CV_8UC1, (void*)a, 32), cv::Mat(1, 32, CV_8UC1, (void*)b, 32));
}
void tbrefm_three_maxima(const int32_t* hist_sizes, int L, int* i1, int* i2, int* i3)
{
    std::vector<std::vector<int> > h((size_t)L);
    for (int i = 0; i < L; i++) h[i].resize((size_t)hist_sizes[i]);
    Matcher::ComputeThreeMaxima(h.data(), L, *i1, *i2, *i3);
}

// ---- ORB stereo row-band matcher (LocalBA.cpp:69-288, dead code in the reference): depth[N] and the match list
int tbrefm_stereo_rowband(void* current, void* stereo, float bf, float fx, float* depth, tbo_match* out, int cap)
{
    std::vector<float> d = tb_stereo_rowband_body(*(FramePtr*)current, *(FramePtr*)stereo, bf, fx);
    for (size_t i = 0; i < d.size(); i++) depth[i] = d[i];
    return put_matches(tb_stereo_matches, out, cap);
}

// ---- FASTExtractor (FASTextractor.cpp, whole file): operator() when n_exit < 0, AddPoints otherwise.
// Returns the keypoint count, or -1 where the reference throws std::out_of_range (gridPoints.at(k)).
int tbrefm_fast_extract(const uint8_t* const* levels, const int* w, const int* h, const int* stride, int n_levels, const float* inv_sf,
                        int n_features, float threshold, const tbo_kp* exit_kps, int n_exit, tbo_kp* out, int cap)
{
    std::vector<cv::Mat> images;
    for (int l = 0; l < n_levels; l++) images.push_back(cv::Mat(h[l], w[l], CV_8UC1, (void*)levels[l], (size_t)stride[l]));
    std::vector<float> isf(inv_sf, inv_sf + n_levels);
    std::vector<cv::KeyPoint> kps;
    cv::Mat d;
    cv::_OutputArray oa(d);
    FASTExtractor ex;
    try {
        if (n_exit < 0) {
            ex(images, isf, n_features, threshold, kps, oa);
        } else {
            std::vector<cv::KeyPoint> ek;
            for (int i = 0; i < n_exit; i++) ek.push_back(cv::KeyPoint(exit_kps[i].x, exit_kps[i].y, exit_kps[i].size, exit_kps[i].angle, exit_kps[i].response, exit_kps[i].octave, exit_kps[i].class_id));
            ex.AddPoints(images, isf, n_features, threshold, ek, kps, oa);
        }
    } catch (const std::out_of_range&) {
        return -1;
    }
    const int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; i++) {
        tbo_kp k = {kps[i].pt.x, kps[i].pt.y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id};
        out[i] = k;
    }
    return n;
}
float tbrefm_shi_tomasi(const uint8_t* img, int w, int h, int stride, int u, int v)
{
    return FASTExtractor::shiTomasiScore(cv::Mat(h, w, CV_8UC1, (void*)img, (size_t)stride), u, v);
}

}  // extern "C"
