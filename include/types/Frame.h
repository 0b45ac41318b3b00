// Frame.h — facade of the reference's include/types/Frame.h:53-202 restricted to what the feature
// tracking path touches: construction (scale tables + pyramid, src/types/Frame.cpp:13-33, :414-427),
// keypoints / descriptors (SetKeys :69), the pose, the flattened map-point side data the projection
// matchers read, and the search-grid query (AssignFeaturesToGrid / GetFeaturesInArea :187-265, run
// on the GPU).  Eigen / DBoW2 / Map bookkeeping of the reference Frame are out of scope.
#ifndef TRACKING_BENCH_FRAME_H
#define TRACKING_BENCH_FRAME_H
#include <memory>
#include <vector>
#include "../camera/CameraModel.h"
#include "../tb_cv_compat.h"
#include "../tb_runtime.h"

namespace TRACKING_BENCH {
#define FRAME_GRID_ROWS 36
#define FRAME_GRID_COLS 120
const int EDGE_THRESHOLD = 19;

// Flattened MapPoint (include/types/MapPoint.h): what Matcher::searchByProjection reads.
struct MapPointLite {
    float pos[3] = {0, 0, 0};      // GetWorldPos()
    float normal[3] = {0, 0, 1};   // GetNormal()
    uint8_t desc[32] = {0};        // GetDescriptor()
    bool bad = false;              // isBad()
    int observations = 0;          // Observations()
};

class Frame {
public:
    Frame(const cv::Mat &imGray, const double &timeStamp, int level, float scale, std::shared_ptr<CameraModel> camera)
        : mTimeStamp(timeStamp), mpCamera(std::move(camera)), nLevels(level), scaleFactor(scale)
    {
        tb_ctx *ctx = tb_default_ctx();
        mvScaleFactor.resize(nLevels);
        std::vector<int32_t> w(nLevels), h(nLevels);
        tb_check(tb_pyramid_geometry(imGray.cols, imGray.rows, nLevels, scale, mvScaleFactor.data(), w.data(), h.data()), ctx, "tb_pyramid_geometry");
        mvInvScaleFactor.assign(nLevels, 1.f);
        for (int i = 1; i < nLevels; i++) mvInvScaleFactor[i] = mvInvScaleFactor[i - 1] / scale;
        cols_ = imGray.cols; rows_ = imGray.rows;
        ComputePyramid(imGray, w, h);
        for (int i = 0; i < 16; i++) mTcw[i] = (i % 5 == 0) ? 1.f : 0.f;
    }
    // pose (row-major 4x4 Tcw)
    void SetPose(const float *Tcw16) { for (int i = 0; i < 16; i++) mTcw[i] = Tcw16[i]; }
    const float *GetPose() const { return mTcw; }
    // features
    void SetKeys(std::vector<cv::KeyPoint> &pts, cv::Mat descriptors = cv::Mat())
    {
        mvKeys = pts;
        mDescriptors = descriptors;
        mvpMapPoints.assign(pts.size(), nullptr);
    }
    std::vector<cv::KeyPoint> &GetKeys() { return mvKeys; }
    cv::Mat GetDescriptors() const { return mDescriptors; }
    void AssignFeaturesToGrid() {}   // the grid is (re)built on the device inside each windowed query
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, int minLevel = -1, int maxLevel = -1) const
    {
        tb_ctx *ctx = tb_default_ctx();
        std::vector<int32_t> out(mvKeys.size() + 1);
        int n = 0;
        tb_check(tb_features_in_area(ctx, reinterpret_cast<const tb_keypoint *>(mvKeys.data()), (int)mvKeys.size(), cols_, rows_,
                                     x, y, r, minLevel, maxLevel, out.data(), (int)out.size(), &n), ctx, "tb_features_in_area");
        return std::vector<size_t>(out.begin(), out.begin() + n);
    }
    // map points
    void AddMapPoint(std::shared_ptr<MapPointLite> &pMP, const size_t &idx) { mvpMapPoints.at(idx) = pMP; }
    std::shared_ptr<MapPointLite> GetMapPoint(const size_t &idx) { return mvpMapPoints.at(idx); }
    // pyramid / scale
    int GetMaxLevel() const { return nLevels; }
    int GetLevels() const { return nLevels; }
    float GetScaleFactor() const { return scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<cv::Mat> &GetImagePyramid() { return mvImagePyramid; }
    cv::Mat GetImage() { return mvImagePyramid[0]; }
    std::shared_ptr<CameraModel> GetCameraModel() { return mpCamera; }
    int cols() const { return cols_; }
    int rows() const { return rows_; }

protected:
    // Frame::ComputePyramid (Frame.cpp:414-427) on the GPU: level 0 shares the caller's buffer.
    void ComputePyramid(const cv::Mat &image, const std::vector<int32_t> &w, const std::vector<int32_t> &h)
    {
        tb_ctx *ctx = tb_default_ctx();
        tb_frontend *fe = nullptr;
        tb_check(tb_frontend_create(ctx, nLevels, w.data(), h.data(), 1, 16, &fe), ctx, "tb_frontend_create");
        tb_check(tb_frontend_upload_level(fe, 0, 0, image.data, (size_t)image.step), ctx, "upload");
        tb_check(tb_frontend_build_pyramid(fe, 1), ctx, "tb_frontend_build_pyramid");
        mvImagePyramid.clear();
        mvImagePyramid.push_back(image);
        for (int l = 1; l < nLevels; l++) {
            cv::Mat m(h[l], w[l], CV_8U);
            tb_check(tb_frontend_download_level(fe, 0, l, 0, m.data, (size_t)m.step), ctx, "download");
            mvImagePyramid.push_back(m);
        }
        tb_frontend_destroy(fe);
    }
    double mTimeStamp{};
    float mTcw[16];
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    std::vector<std::shared_ptr<MapPointLite>> mvpMapPoints;
    std::shared_ptr<CameraModel> mpCamera;
    int nLevels;
    float scaleFactor;
    int cols_ = 0, rows_ = 0;
    std::vector<float> mvScaleFactor, mvInvScaleFactor;
    std::vector<cv::Mat> mvImagePyramid;
};
}  // namespace TRACKING_BENCH
#endif  // TRACKING_BENCH_FRAME_H
