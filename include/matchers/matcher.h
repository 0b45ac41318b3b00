// matcher.h — drop-in facade of the reference's include/matchers/matcher.h:17-183 for the
// descriptor matchers on the hot path (searchByBF, searchByViolence, searchByProjection x2, searchByBow,
// DescriptorDistance, ComputeThreeMaxima).  Public members, setters, method names and argument
// meaning are the reference's; bodies (src/matchers/matcher.cpp:168-617,793-851) run on the GPU.
// searchByOPFlow (matcher.cpp:724) runs its pyramidal LK and its rejectWithF (cv::findFundamentalMat RANSAC, :853-881) on the GPU.
// searchByNN (approximate FLANN-LSH) and the direct-alignment methods are out of scope.
#ifndef TRACKING_BENCH_MATCHER_H
#define TRACKING_BENCH_MATCHER_H
#include <cassert>
#include <memory>
#include <vector>
#include "../tb_cv_compat.h"
#include "../tb_runtime.h"
#include "../types/Frame.h"

namespace TRACKING_BENCH {
class Matcher {
public:
    Matcher() = default;
    ~Matcher() = default;
    int TH_LOW = 50;
    int TH_HIGH = 100;
    int HISTO_LENGTH = 30;
    bool checkOrientation = true;
    float nRatio{};

    void setViolenceParam(int low, int high, int histo_length, bool check, float ratio) { TH_LOW = low; TH_HIGH = high; HISTO_LENGTH = histo_length; checkOrientation = check; nRatio = ratio; }
    void setProjectionParam(int low, int high, int histo_length, bool check, float ratio) { setViolenceParam(low, high, histo_length, check, ratio); }
    void setBowParam(int low, int high, int histo_length, bool check, float ratio) { setViolenceParam(low, high, histo_length, check, ratio); }

    // matcher.cpp:168 — only the all-levels call is defined behaviour in the reference.
    std::vector<cv::DMatch> searchByBF(const std::shared_ptr<Frame> &F1, const std::shared_ptr<Frame> &F2, int MinLevel, int MaxLevel,
                                       float ratio, float minTh, bool MapPointOnly = false)
    {
        assert(MinLevel <= MaxLevel);
        if (!(MinLevel == 0 && MaxLevel == F1->GetMaxLevel() && !MapPointOnly))
            throw std::invalid_argument("searchByBF: the partial-level branch writes into an empty cv::Mat in the reference (matcher.cpp:183-203); not supported");
        tb_ctx *ctx = tb_default_ctx();
        cv::Mat d1 = contiguous(F1->GetDescriptors()), d2 = contiguous(F2->GetDescriptors());
        // host descriptors -> device through the windowed-matcher staging path would copy twice; use the
        // one-shot host entry (violence with an infinite window is NOT the same thing), so stage manually:
        std::vector<tb_match> out((size_t)(d1.rows > 0 ? d1.rows : 1));
        int n = 0;
        tb_check(tb_search_by_bf_host(ctx, d1.data, d1.rows, d2.data, d2.rows, ratio, minTh, out.data(), (int)out.size(), &n), ctx, "tb_search_by_bf_host");
        return to_dmatch(out, n);
    }
    // matcher.cpp:299
    std::vector<cv::DMatch> searchByViolence(const std::shared_ptr<Frame> &F1, const std::shared_ptr<Frame> &F2, int min_level = 0,
                                             int max_level = 1, float search_r = 10, bool MapPointOnly = false)
    {
        if (MapPointOnly) throw std::invalid_argument("searchByViolence(MapPointOnly=true) is not on the accelerated path");
        tb_ctx *ctx = tb_default_ctx();
        cv::Mat d1 = contiguous(F1->GetDescriptors()), d2 = contiguous(F2->GetDescriptors());
        tb_matcher_params p = params();
        std::vector<tb_match> out(F1->GetKeys().size() + 1);
        int n = 0;
        tb_check(tb_search_by_violence(ctx, &p, kp(F1), d1.data, (int)F1->GetKeys().size(), kp(F2), d2.data, (int)F2->GetKeys().size(),
                                       F2->cols(), F2->rows(), min_level, max_level, search_r, out.data(), (int)out.size(), &n), ctx, "tb_search_by_violence");
        return to_dmatch(out, n);
    }
    // matcher.cpp:405 — F2's map points are projected into F1.
    std::vector<cv::DMatch> searchByProjection(const std::shared_ptr<Frame> &F1, const std::shared_ptr<Frame> &F2)
    {
        tb_ctx *ctx = tb_default_ctx();
        const int n1 = (int)F1->GetKeys().size(), n2 = (int)F2->GetKeys().size();
        std::vector<uint8_t> valid(n2 + 1, 0), mdesc((size_t)32 * (n2 + 1), 0), matched(n1 + 1, 0);
        std::vector<float> xyz((size_t)3 * (n2 + 1), 0.f);
        for (int i = 0; i < n2; i++) {
            auto mp = F2->GetMapPoint(i);
            if (mp && !mp->bad) { valid[i] = 1; std::memcpy(&xyz[3 * i], mp->pos, 12); std::memcpy(&mdesc[(size_t)32 * i], mp->desc, 32); }
        }
        for (int i = 0; i < n1; i++) { auto mp = F1->GetMapPoint(i); matched[i] = (mp && mp->observations > 0) ? 1 : 0; }
        cv::Mat d1 = contiguous(F1->GetDescriptors());
        tb_matcher_params p = params();
        tb_pinhole cam = F1->GetCameraModel()->pod();
        std::vector<float> sf = F1->GetScaleFactors();
        std::vector<tb_match> out(n2 + 1);
        int n = 0;
        tb_check(tb_search_by_projection(ctx, &p, &cam, F1->GetPose(), kp(F1), d1.data, n1, matched.data(), sf.data(), (int)sf.size(),
                                         F1->cols(), F1->rows(), kp(F2), n2, valid.data(), xyz.data(), mdesc.data(), out.data(), (int)out.size(), &n),
                 ctx, "tb_search_by_projection");
        return to_dmatch(out, n);
    }
    // matcher.cpp:539 — map points in the caller's iteration order (the reference iterates a std::set of
    // shared_ptr, i.e. allocator order).
    std::vector<cv::DMatch> searchByProjection(const std::vector<std::shared_ptr<MapPointLite>> &map, const std::shared_ptr<Frame> &F1, float r)
    {
        tb_ctx *ctx = tb_default_ctx();
        const int n1 = (int)F1->GetKeys().size(), nm = (int)map.size();
        std::vector<uint8_t> bad(nm + 1, 1), mdesc((size_t)32 * (nm + 1), 0), matched(n1 + 1, 0);
        std::vector<float> xyz((size_t)3 * (nm + 1), 0.f), nrm((size_t)3 * (nm + 1), 0.f);
        for (int i = 0; i < nm; i++) {
            bad[i] = map[i]->bad ? 1 : 0;
            std::memcpy(&xyz[3 * i], map[i]->pos, 12); std::memcpy(&nrm[3 * i], map[i]->normal, 12); std::memcpy(&mdesc[(size_t)32 * i], map[i]->desc, 32);
        }
        for (int i = 0; i < n1; i++) { auto mp = F1->GetMapPoint(i); matched[i] = (mp && mp->observations > 0) ? 1 : 0; }
        cv::Mat d1 = contiguous(F1->GetDescriptors());
        tb_matcher_params p = params();
        tb_pinhole cam = F1->GetCameraModel()->pod();
        std::vector<float> sf = F1->GetScaleFactors();
        std::vector<tb_match> out(nm + 1);
        int n = 0;
        tb_check(tb_search_by_projection_map(ctx, &p, &cam, F1->GetPose(), kp(F1), d1.data, n1, matched.data(), sf.data(), F1->cols(), F1->rows(),
                                             nm, bad.data(), xyz.data(), nrm.data(), mdesc.data(), r, out.data(), (int)out.size(), &n),
                 ctx, "tb_search_by_projection_map");
        return to_dmatch(out, n);
    }
    // matcher.cpp:619 — features that fell into the same vocabulary node (Frame::SetBow first).
    std::vector<cv::DMatch> searchByBow(const std::shared_ptr<Frame> &F1, const std::shared_ptr<Frame> &F2, bool MapPointOnly = false)
    {
        tb_ctx *ctx = tb_default_ctx();
        const int n1 = (int)F1->GetKeys().size(), n2 = (int)F2->GetKeys().size();
        std::vector<int32_t> node1(n1 + 1, -1), node2(n2 + 1, -1);
        for (const auto &e : F1->GetFeatureVector()) for (unsigned i : e.second) node1[i] = (int32_t)e.first;
        for (const auto &e : F2->GetFeatureVector()) for (unsigned i : e.second) node2[i] = (int32_t)e.first;
        std::vector<uint8_t> has_mp(n2 + 1, 0);
        if (MapPointOnly) for (int i = 0; i < n2; i++) has_mp[i] = F2->GetMapPoint(i) ? 1 : 0;
        cv::Mat d1 = contiguous(F1->GetDescriptors()), d2 = contiguous(F2->GetDescriptors());
        tb_matcher_params p = params();
        std::vector<tb_match> out(n1 + 1);
        int n = 0;
        tb_check(tb_search_by_bow(ctx, &p, kp(F1), d1.data, node1.data(), n1, kp(F2), d2.data, node2.data(), n2, has_mp.data(),
                                  MapPointOnly ? 1 : 0, out.data(), (int)out.size(), &n), ctx, "tb_search_by_bow");
        return to_dmatch(out, n);
    }
    // matcher.cpp:724 — LK tracks of F2's keypoints into F1 (cv::calcOpticalFlowPyrLK, Size(21,21), 3 levels), status
    // cleared outside the frame (:747-749), then — reject = true, as every caller in the reference passes — rejectWithF
    // (:751-755); matches are {i, i} for the tracks that are left.
    std::vector<cv::DMatch> searchByOPFlow(const std::shared_ptr<Frame> &F1, const std::shared_ptr<Frame> &F2, std::vector<cv::Point2f> &cur_points,
                                           bool equalized, bool reject, bool MapPointOnly = false)
    {
        (void)MapPointOnly;   // unused by the reference body as well
        tb_ctx *ctx = tb_default_ctx();
        cv::Mat img1 = contiguous(equalized ? F1->Equalize() : F1->GetImage()), img2 = contiguous(F2->GetImage());
        const std::vector<cv::Point2f> prev = F2->GetCVKeys();
        const int n = (int)prev.size();
        cur_points.assign(n, cv::Point2f());
        std::vector<uint8_t> status(n + 1, 0);
        tb_lk_params lk = {21, 3, 30, 0.01f, 1e-4f};
        static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f must be two packed floats");
        if (n) tb_check(tb_lk_flow_host(ctx, &lk, img2.data, img1.data, img2.cols, img2.rows, img2.step, (const float *)prev.data(), n,
                                        (float *)cur_points.data(), status.data(), nullptr), ctx, "tb_lk_flow_host");
        const int W = F1->GetCameraModel()->Width(), H = F1->GetCameraModel()->Height();
        for (int i = 0; i < n; i++) {
            const int xi = (int)cur_points[i].x, yi = (int)cur_points[i].y;
            if (status[i] && !(xi >= 0 && xi < W && yi >= 0 && yi < H)) status[i] = 0;
        }
        if (reject) rejectWithF(cur_points, prev, status);
        std::vector<cv::DMatch> matches;
        for (int i = 0; i < n; i++)
            if (status[i]) { cv::DMatch m; m.queryIdx = i; m.trainIdx = i; matches.push_back(m); }
        return matches;
    }
    // matcher.cpp:853 — cv::findFundamentalMat(cur, last, FM_RANSAC, 1.0, 0.99) over the tracks with status != 0; outliers lose
    // their status.  (The reference reads an empty mask when OpenCV returns none — fewer than 15 valid tracks run LMedS or
    // nothing there; here status stays as it is.)
    void rejectWithF(std::vector<cv::Point2f> &cur_pts, const std::vector<cv::Point2f> &last_pts, std::vector<uint8_t> &status)
    {
        tb_ctx *ctx = tb_default_ctx();
        const int n = (int)std::min(cur_pts.size(), last_pts.size());
        if (n == 0) return;
        if (status.size() < (size_t)n) status.resize(n, 0);
        const tb_ransac_params rp = {1.0f, 0.99f, 1000};
        tb_check(tb_reject_with_f_host(ctx, &rp, (const float *)cur_pts.data(), (const float *)last_pts.data(), n, status.data(), nullptr, nullptr),
                 ctx, "tb_reject_with_f_host");
    }
    // matcher.cpp:793
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
    {
        tb_ctx *ctx = tb_default_ctx();
        int d = -1;
        tb_check(tb_hamming_pairs_host(ctx, a.ptr(0), b.ptr(0), 1, &d), ctx, "tb_hamming_pairs_host");
        return d;
    }
    // matcher.cpp:810
    static void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3)
    {
        std::vector<int32_t> hs(L);
        for (int i = 0; i < L; i++) hs[i] = (int32_t)histo[i].size();
        tb_three_maxima(hs.data(), L, &ind1, &ind2, &ind3);
    }

private:
    tb_matcher_params params() const { tb_matcher_params p = {TH_LOW, TH_HIGH, HISTO_LENGTH, checkOrientation ? 1 : 0, nRatio}; return p; }
    static const tb_keypoint *kp(const std::shared_ptr<Frame> &F)
    {
        static_assert(sizeof(cv::KeyPoint) == sizeof(tb_keypoint), "cv::KeyPoint layout");
        static const tb_keypoint none = {0, 0, 0, 0, 0, 0, 0};
        return F->GetKeys().empty() ? &none : reinterpret_cast<const tb_keypoint *>(F->GetKeys().data());
    }
    static cv::Mat contiguous(const cv::Mat &m) { return (m.empty() || (size_t)m.step == (size_t)m.cols) ? m : m.clone(); }
    static std::vector<cv::DMatch> to_dmatch(const std::vector<tb_match> &m, int n)
    {
        std::vector<cv::DMatch> out;
        out.reserve(n);
        for (int i = 0; i < n; i++) out.push_back(cv::DMatch(m[i].queryIdx, m[i].trainIdx, m[i].imgIdx, m[i].distance));   // imgIdx: 0 from BFMatcher, -1 from DMatch(q,t,dist)
        return out;
    }
};
}  // namespace TRACKING_BENCH
#endif  // TRACKING_BENCH_MATCHER_H
