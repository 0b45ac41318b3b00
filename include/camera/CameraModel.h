// CameraModel.h — facade of the reference's include/camera/CameraModel.h:9-87 (namespace, class and
// method names kept) without Eigen: only what the projection matchers consume.  The projection math
// itself (World2Cam, IsInFrame — src/camera/CameraModel.cpp:63-94, CameraModel.h:33-39) runs inside
// the CUDA matchers; this class carries the intrinsics and, for a distorted camera, the undistortion entry points
// (UndistortImage / UndistortPoint, src/camera/CameraModel.cpp:96-121; cv::Point2f stands in for Eigen::Vector2f).
#ifndef TRACKING_BENCH_CAMERAMODEL_H
#define TRACKING_BENCH_CAMERAMODEL_H
#include <cmath>
#include <memory>
#include <vector>
#include "../tb_capi.h"
#include "../tb_cv_compat.h"
#include "../tb_runtime.h"
namespace TRACKING_BENCH {
class CameraModel {
protected:
    int mnWidth{}, mnHeight{};
public:
    CameraModel() = default;
    CameraModel(int width, int height) : mnWidth(width), mnHeight(height) {}
    virtual ~CameraModel() = default;
    inline int Width() const { return mnWidth; }
    inline int Height() const { return mnHeight; }
    virtual tb_pinhole pod() const = 0;
    virtual void UndistortImage(const cv::Mat &raw, cv::Mat &rectified) = 0;
};
class PinholeCamera : public CameraModel {
public:
    PinholeCamera(int width, int height, float fx, float fy, float cx, float cy,
                  float d0 = 0.0, float d1 = 0.0, float d2 = 0.0, float d3 = 0.0, float d4 = 0.0)
        : CameraModel(width, height), mFx(fx), mFy(fy), mCx(cx), mCy(cy) { md[0] = d0; md[1] = d1; md[2] = d2; md[3] = d3; md[4] = d4; }
    inline float fx() const { return mFx; }
    inline float fy() const { return mFy; }
    inline float cx() const { return mCx; }
    inline float cy() const { return mCy; }
    tb_pinhole pod() const override
    {
        tb_pinhole p = {mFx, mFy, mCx, mCy, mnWidth, mnHeight, {md[0], md[1], md[2], md[3], md[4]}};
        return p;
    }
    // CameraModel.cpp:96-102: cv::remap(raw, rectified, map1, map2, INTER_LINEAR) when distorted, a copy otherwise
    void UndistortImage(const cv::Mat &raw, cv::Mat &rectified) override
    {
        if (!(std::fabs(md[0]) > 1e-7f)) { rectified = raw.clone(); return; }
        tb_ctx *ctx = tb_default_ctx();
        if (!mUndist) {   // CameraModel.cpp:15-18: maps once per camera
            tb_pinhole cam = pod();
            std::vector<float> m1((size_t)mnWidth * mnHeight), m2((size_t)mnWidth * mnHeight);
            tb_check(tb_init_undistort_rectify_map(&cam, m1.data(), m2.data()), ctx, "tb_init_undistort_rectify_map");
            tb_undistort *u = nullptr;
            tb_check(tb_undistort_create(ctx, mnWidth, mnHeight, m1.data(), m2.data(), &u), ctx, "tb_undistort_create");
            mUndist = std::shared_ptr<tb_undistort>(u, tb_undistort_destroy);
        }
        cv::Mat dst(mnHeight, mnWidth, CV_8U);
        tb_check(tb_undistort_image_host(ctx, mUndist.get(), raw.data, raw.cols, raw.rows, raw.step, dst.data, dst.step), ctx, "tb_undistort_image_host");
        rectified = dst;
    }
    // CameraModel.cpp:104-121: cv::undistortPoints(src, dst, K, D, Mat(), K) on one point / on a list
    cv::Point2f UndistortPoint(const cv::Point2f &uv_pt) const
    {
        std::vector<cv::Point2f> v(1, uv_pt);
        return UndistortPoints(v)[0];
    }
    std::vector<cv::Point2f> UndistortPoints(const std::vector<cv::Point2f> &pts) const
    {
        std::vector<cv::Point2f> out(pts.size());
        if (pts.empty()) return out;
        tb_ctx *ctx = tb_default_ctx();
        tb_pinhole cam = pod();
        tb_check(tb_undistort_points(ctx, &cam, (const float *)pts.data(), (int)pts.size(), (float *)out.data()), ctx, "tb_undistort_points");
        return out;
    }
private:
    const float mFx, mFy, mCx, mCy;
    float md[5]{};
    std::shared_ptr<tb_undistort> mUndist;
};
}  // namespace TRACKING_BENCH
#endif
