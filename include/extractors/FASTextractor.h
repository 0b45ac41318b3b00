// FASTextractor.h — drop-in facade of the reference's include/extractors/FASTextractor.h:11-41.
// Same namespace, class and method names; bodies (src/extractors/FASTextractor.cpp) run on the GPU.
// `mvScaleFactor` are the factors the caller passes (the reference multiplies coordinates by them, its
// .cpp calls the parameter invScaleFactor).  Where the reference throws std::out_of_range
// (gridPoints.at(k), grid_occupancy_.at(k)) this facade throws std::out_of_range too.
#ifndef TRACKING_BENCH_FASTEXTRACTOR_H
#define TRACKING_BENCH_FASTEXTRACTOR_H
#include <stdexcept>
#include <vector>
#include "../tb_cv_compat.h"
#include "../tb_runtime.h"

namespace TRACKING_BENCH {
class FASTExtractor {
public:
    FASTExtractor() = default;
    ~FASTExtractor() = default;
    // detection (FASTextractor.cpp:8-80). With reset == false the occupancy set by AddPoints is honoured.
    void operator()(std::vector<cv::Mat> &images, std::vector<float> &mvScaleFactor, int targetNum, float threshold,
                    std::vector<cv::KeyPoint> &keypoints, cv::Mat & /*descriptors: none are produced*/, bool reset = true)
    {
        if (images.empty() || images.at(0).empty()) return;
        run(images, mvScaleFactor, targetNum, threshold, reset ? nullptr : &exit_, keypoints);
        resetGrid();
    }
    // FASTextractor.cpp:129-150
    void AddPoints(std::vector<cv::Mat> &images, std::vector<float> &mvScaleFactor, int targetNum, float threshold,
                   const std::vector<cv::KeyPoint> &exitPoints, std::vector<cv::KeyPoint> &newPoints, cv::Mat &descriptors)
    {
        exit_ = exitPoints;
        operator()(images, mvScaleFactor, targetNum, threshold, newPoints, descriptors, false);
    }
    void resetGrid() { exit_.clear(); }

private:
    void run(std::vector<cv::Mat> &images, std::vector<float> &sf, int target, float threshold,
             const std::vector<cv::KeyPoint> *exitPoints, std::vector<cv::KeyPoint> &kps)
    {
        tb_ctx *ctx = tb_default_ctx();
        const int L = (int)images.size();
        std::vector<const uint8_t *> ptr(L);
        std::vector<int32_t> w(L), h(L);
        std::vector<size_t> stride(L);
        for (int l = 0; l < L; l++) { ptr[l] = images[l].data; w[l] = images[l].cols; h[l] = images[l].rows; stride[l] = (size_t)images[l].step; }
        static_assert(sizeof(cv::KeyPoint) == sizeof(tb_keypoint), "cv::KeyPoint layout");
        static const tb_keypoint none = {0, 0, 0, 0, 0, 0, 0};
        const tb_keypoint *ex = nullptr;
        int n_exit = 0;
        if (exitPoints) { ex = exitPoints->empty() ? &none : reinterpret_cast<const tb_keypoint *>(exitPoints->data()); n_exit = (int)exitPoints->size(); }
        std::vector<tb_keypoint> out((size_t)target + 1);
        int n = 0;
        int rc = tb_fast_extract_host(ctx, ptr.data(), w.data(), h.data(), stride.data(), L, sf.data(), target, threshold, ex, n_exit,
                                      out.data(), (int)out.size(), &n);
        if (rc == TB_E_INVALID) throw std::out_of_range(tb_last_error(ctx));
        tb_check(rc, ctx, "tb_fast_extract_host");
        kps.clear();
        for (int i = 0; i < n; i++) kps.push_back(cv::KeyPoint(out[i].x, out[i].y, out[i].size, out[i].angle, out[i].response, out[i].octave, out[i].class_id));
    }
    std::vector<cv::KeyPoint> exit_;
};
}  // namespace TRACKING_BENCH
#endif  // TRACKING_BENCH_FASTEXTRACTOR_H
