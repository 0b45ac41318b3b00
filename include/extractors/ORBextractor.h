// ORBextractor.h — drop-in facade of the reference's include/extractors/ORBextractor.h:23-92.
// Same namespace, class, method names and argument meaning; the bodies
// (src/extractors/ORBextractor.cpp:840-978) are replaced by one call into the CUDA library.
#ifndef TRACKING_BENCH_ORBEXTRACTOR_H
#define TRACKING_BENCH_ORBEXTRACTOR_H
#include <cmath>
#include <list>
#include <vector>
#include "../tb_cv_compat.h"
#include "../tb_runtime.h"

namespace TRACKING_BENCH {
// ORBextractor.h:9-21 — the quadtree node of DistributeOctTree.  The distribution itself runs on the device
// (csrc/octree.cu); the type is kept so that code written against the reference header still compiles.  DivideNode
// (ORBextractor.cpp:416-491): children split at ceil(w/2), ceil(h/2); a key goes left/up when its coordinate is strictly
// below the split; a child holding exactly one key is final (bNoMore).
class ExtractorNode {
public:
    ExtractorNode() : bNoMore(false) {}
    void DivideNode(ExtractorNode &n1, ExtractorNode &n2, ExtractorNode &n3, ExtractorNode &n4)
    {
        ExtractorNode *child[4] = {&n1, &n2, &n3, &n4};
        const int mx = UL.x + (int)std::ceil(static_cast<float>(UR.x - UL.x) / 2), my = UL.y + (int)std::ceil(static_cast<float>(BR.y - UL.y) / 2);
        const int xs[3] = {UL.x, mx, UR.x}, ys[3] = {UL.y, my, BL.y};
        for (int q = 0; q < 4; q++) {   // q = 2*row + column: n1 upper-left, n2 upper-right, n3 lower-left, n4 lower-right
            const int cx = q & 1, cy = q >> 1;
            child[q]->UL = cv::Point2i(xs[cx], ys[cy]);     child[q]->UR = cv::Point2i(xs[cx + 1], ys[cy]);
            child[q]->BL = cv::Point2i(xs[cx], ys[cy + 1]); child[q]->BR = cv::Point2i(xs[cx + 1], ys[cy + 1]);
            child[q]->vKeys.reserve(vKeys.size());
            child[q]->vExitKeys.reserve(vExitKeys.size());
        }
        n2.BR.y = my; n4.BR = BR;   // right column ends at this node's own right edge (UR.x / BR), as in the reference
        auto quadrant = [&](const cv::KeyPoint &kp) { return (kp.pt.x < (float)mx ? 0 : 1) + (kp.pt.y < (float)my ? 0 : 2); };
        for (const cv::KeyPoint &kp : vKeys) child[quadrant(kp)]->vKeys.push_back(kp);
        for (const cv::KeyPoint &kp : vExitKeys) child[quadrant(kp)]->vExitKeys.push_back(kp);
        for (int q = 0; q < 4; q++) if (child[q]->vKeys.size() == 1) child[q]->bNoMore = true;
    }
    std::vector<cv::KeyPoint> vKeys;
    std::vector<cv::KeyPoint> vExitKeys;
    cv::Point2i UL, UR, BL, BR;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore;
};

class ORBExtractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    ORBExtractor() = default;
    ~ORBExtractor() { if (fe_) tb_frontend_destroy(fe_); }
    ORBExtractor(const ORBExtractor &) = delete;
    ORBExtractor &operator=(const ORBExtractor &) = delete;

    // ORBextractor.cpp:906 — silent return on empty input, like the reference (:914-915).
    void operator()(std::vector<cv::Mat> &images, std::vector<float> mvScaleFactor, int targetNum, float initTh, float minTH,
                    std::vector<cv::KeyPoint> &keypoints, cv::Mat &descriptors)
    {
        if (images.empty() || images.at(0).empty()) return;
        run(images, mvScaleFactor, targetNum, initTh, minTH, nullptr, 0, keypoints, descriptors);
    }
    // ORBextractor.cpp:840 — reuses the per-level budget of the last operator() call.
    void AddPoints(std::vector<cv::Mat> &images, std::vector<float> &mvScaleFactor, int targetNum, float initTh, float minTH,
                   const std::vector<cv::KeyPoint> &exitPoints, std::vector<cv::KeyPoint> &newPoints, cv::Mat &descriptors)
    {
        if (images.empty() || images.at(0).empty()) return;
        static_assert(sizeof(cv::KeyPoint) == sizeof(tb_keypoint), "cv::KeyPoint layout");
        static const tb_keypoint none = {0, 0, 0, 0, 0, 0, 0};
        const tb_keypoint *ex = exitPoints.empty() ? &none : reinterpret_cast<const tb_keypoint *>(exitPoints.data());
        run(images, mvScaleFactor, targetNum, initTh, minTH, ex, (int)exitPoints.size(), newPoints, descriptors);
    }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

protected:
    void run(std::vector<cv::Mat> &images, const std::vector<float> &sf, int target, float initTh, float minTh,
             const tb_keypoint *ex, int n_exit, std::vector<cv::KeyPoint> &kps, cv::Mat &desc)
    {
        tb_ctx *ctx = tb_default_ctx();
        const int L = (int)images.size();
        std::vector<int32_t> w(L), h(L);
        for (int l = 0; l < L; l++) { w[l] = images[l].cols; h[l] = images[l].rows; }
        if (!fe_ || w != w_ || h != h_ || target > max_target_) {   // (re)size the device workspace
            // mnFeaturesPerLevel (ORBextractor.h:88) outlives the workspace: AddPoints after a re-size still finds the budget
            int32_t budget[TB_MAX_LEVELS] = {0};
            int nb = 0;
            if (fe_) { tb_frontend_get_budget(fe_, budget, &nb); tb_frontend_destroy(fe_); }
            fe_ = nullptr;
            max_target_ = target > max_target_ ? target : max_target_;
            for (int l = 0; l < nb; l++) if (budget[l] > max_target_) max_target_ = budget[l];
            tb_check(tb_frontend_create(ctx, L, w.data(), h.data(), 1, max_target_, &fe_), ctx, "tb_frontend_create");
            if (nb == L) tb_check(tb_frontend_set_budget(fe_, budget, nb), ctx, "tb_frontend_set_budget");
            w_ = w; h_ = h;
        }
        std::vector<const uint8_t *> ptr(L);
        std::vector<size_t> stride(L);
        for (int l = 0; l < L; l++) { ptr[l] = images[l].data; stride[l] = (size_t)images[l].step; }
        const int cap = tb_frontend_kp_capacity(fe_);
        std::vector<tb_keypoint> out(cap);
        std::vector<uint8_t> d((size_t)cap * 32);
        int n = 0;
        tb_check(tb_orb_extract_host(fe_, ptr.data(), stride.data(), sf.data(), target, initTh, minTh, ex, n_exit,
                                     out.data(), d.data(), cap, &n), ctx, "tb_orb_extract_host");
        kps.clear();
        kps.reserve(n);
        for (int i = 0; i < n; i++) kps.push_back(cv::KeyPoint(out[i].x, out[i].y, out[i].size, out[i].angle, out[i].response, out[i].octave, out[i].class_id));
        if (n == 0) { desc.release(); return; }
        desc.create(n, 32, CV_8U);
        for (int i = 0; i < n; i++) std::memcpy(desc.ptr(i), &d[(size_t)32 * i], 32);
    }
    tb_frontend *fe_ = nullptr;
    std::vector<int32_t> w_, h_;
    int max_target_ = 0;
    std::vector<float> mvLevelSigma2, mvInvLevelSigma2;
};
}  // namespace TRACKING_BENCH
#endif  // TRACKING_BENCH_ORBEXTRACTOR_H
