// CameraModel.h — facade of the reference's include/camera/CameraModel.h:9-87 (namespace, class and
// method names kept) without Eigen: only what the projection matchers consume.  The projection math
// itself (World2Cam, IsInFrame — src/camera/CameraModel.cpp:63-94, CameraModel.h:33-39) runs inside
// the CUDA matchers; this class only carries the intrinsics.
#ifndef TRACKING_BENCH_CAMERAMODEL_H
#define TRACKING_BENCH_CAMERAMODEL_H
#include "../tb_capi.h"
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
private:
    const float mFx, mFy, mCx, mCy;
    float md[5]{};
};
}  // namespace TRACKING_BENCH
#endif
