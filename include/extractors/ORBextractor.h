// ORBextractor.h — drop-in facade of the reference's include/extractors/ORBextractor.h:23-92.
// Same namespace, class, method names and argument meaning; the bodies
// (src/extractors/ORBextractor.cpp:840-978) are replaced by one call into the CUDA library.
#ifndef TRACKING_BENCH_ORBEXTRACTOR_H
#define TRACKING_BENCH_ORBEXTRACTOR_H
#include <vector>
#include "../tb_cv_compat.h"
#include "../tb_runtime.h"

namespace TRACKING_BENCH {
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
            if (fe_) tb_frontend_destroy(fe_);
            fe_ = nullptr;
            max_target_ = target > max_target_ ? target : max_target_;
            tb_check(tb_frontend_create(ctx, L, w.data(), h.data(), 1, max_target_, &fe_), ctx, "tb_frontend_create");
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
