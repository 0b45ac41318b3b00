// cvshim_more.h — the further OpenCV names the reference's Frame / MapPoint / CameraModel / Matcher sources use, on
// top of cvshim.h, so that src/types/{Frame,MapPoint,Map}.cpp and src/camera/CameraModel.cpp compile UNMODIFIED and
// the hot-path member functions of src/matchers/matcher.cpp compile from line ranges taken at build time
// (oracle/Makefile, oracle/ref_matcher_build.cpp).  ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/tb_oracle.h).
// Every image primitive forwards to a C restatement in tb_oracle.c that tests/test_oracle_cv2.py pins against the
// real cv2 4.13: resize, remap, initUndistortRectifyMap, undistortPoints, CLAHE, BFMatcher(HAMMING, crossCheck),
// calcOpticalFlowPyrLK (the last within the tolerance stated there).  Calls the path never makes abort().
#ifndef TB_CVSHIM_MORE_H
#define TB_CVSHIM_MORE_H
#include "cvshim.h"
#include "cvshim_filestorage.h"
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#define CV_INTER_LINEAR 1

namespace cv {

enum { NORM_HAMMING = 6, INTER_LINEAR = 1, FM_RANSAC = 8 };

struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };

// OpenCV: DMatch() = (-1,-1,-1,FLT_MAX); DMatch(q,t,dist) leaves imgIdx = -1
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

template <typename T> struct tb_type_of;
template <> struct tb_type_of<float> { enum { value = CV_32F }; };
template <> struct tb_type_of<double> { enum { value = CV_64F }; };
template <> struct tb_type_of<uchar> { enum { value = CV_8U }; };

template <typename T> struct Mat_ : Mat {
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, (int)tb_type_of<T>::value) {}
};
// (Mat_<float>(3,3) << a, b, c, ...) — CameraModel.cpp:13-14
template <typename T> struct MatCommaInitializer_ {
    Mat_<T> m;
    int n;
    template <typename V> MatCommaInitializer_& operator,(V v) { ((T*)m.data)[n++] = (T)v; return *this; }
    operator Mat() const { return m; }
    operator Mat_<T>() const { return m; }
};
template <typename T, typename V> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v)
{
    MatCommaInitializer_<T> ci;
    ci.m = m; ci.n = 0;
    ci, v;
    return ci;
}

template <typename T> using Ptr = std::shared_ptr<T>;

// cv::resize(src, dst, dsize) with the default INTER_LINEAR — Frame.cpp:423
inline void resize(const Mat& src, Mat& dst, Size dsize)
{
    assert(src.type() == CV_8UC1);
    Mat out(dsize.height, dsize.width, CV_8UC1);
    tbo_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step.v, out.data, out.cols, out.rows, (int)out.step.v);
    dst = out;
}

// cv::initUndistortRectifyMap(K, D, Mat(), K, size, CV_32F, map1, map2) — CameraModel.cpp:15-16
inline void initUndistortRectifyMap(const Mat& K, const Mat& D, const Mat& R, const Mat& newK, Size size, int m1type, Mat& map1, Mat& map2)
{
    if (!R.empty() || newK.data != K.data || m1type != CV_32F || K.type() != CV_32F || D.type() != CV_32F) { std::fprintf(stderr, "cvshim: unsupported initUndistortRectifyMap call\n"); std::abort(); }
    float d[5];
    for (int i = 0; i < 5; i++) d[i] = ((const float*)D.data)[i];
    map1.create(size.height, size.width, CV_32F);
    map2.create(size.height, size.width, CV_32F);
    tbo_init_undistort_rectify_map(K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d, size.width, size.height,
                                   (float*)map1.data, (float*)map2.data);
}

// cv::undistortPoints(src, dst, K, D, R, P): only the P = K form of PinholeCamera::UndistortPoint (CameraModel.cpp:112) is restated
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D, const Mat& R = Mat(), const Mat& P = Mat())
{
    if (!R.empty() || P.data != K.data) { std::fprintf(stderr, "cvshim: undistortPoints without P = K is not on the path\n"); std::abort(); }
    float d[5];
    for (int i = 0; i < 5; i++) d[i] = ((const float*)D.data)[i];
    tbo_undistort_points(K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d, (const float*)src.data,
                         src.rows * src.cols, (float*)dst.data);
}

// cv::remap(raw, rectified, map1, map2, INTER_LINEAR), BORDER_CONSTANT 0 — CameraModel.cpp:99
inline void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interp)
{
    assert(interp == INTER_LINEAR && src.type() == CV_8UC1 && map1.type() == CV_32F);
    (void)interp;
    Mat in = src;  // shares the buffer: stays alive if dst is the same matrix
    Mat out(map1.rows, map1.cols, CV_8UC1);
    tbo_remap_linear_u8(in.data, in.cols, in.rows, (int)in.step.v, (const float*)map1.data, (const float*)map2.data, map1.cols, map1.rows,
                        out.data, (int)out.step.v);
    dst = out;
}

// cv::createCLAHE(clip, tiles)->apply — Frame.cpp:455-456
struct CLAHE {
    double clip;
    Size tiles;
    void apply(const Mat& src, Mat& dst)
    {
        Mat out(src.rows, src.cols, CV_8UC1);
        tbo_clahe_u8(src.data, src.cols, src.rows, (int)src.step.v, clip, tiles.width, tiles.height, out.data, (int)out.step.v);
        dst = out;
    }
};
inline Ptr<CLAHE> createCLAHE(double clip, Size tiles)
{
    Ptr<CLAHE> p(new CLAHE());
    p->clip = clip; p->tiles = tiles;
    return p;
}

// cv::BFMatcher(NORM_HAMMING, crossCheck = true).match — matcher.cpp:20, :207
struct BFMatcher {
    BFMatcher(int normType, bool crossCheck)
    {
        if (normType != NORM_HAMMING || !crossCheck) { std::fprintf(stderr, "cvshim: only BFMatcher(NORM_HAMMING, true)\n"); std::abort(); }
    }
    void match(const Mat& d1, const Mat& d2, std::vector<DMatch>& matches) const
    {
        matches.clear();
        if (d1.rows == 0 || d2.rows == 0) return;
        assert(d1.cols == 32 && d2.cols == 32 && d1.step.v == 32 && d2.step.v == 32);
        std::vector<tbo_match> tmp((size_t)d1.rows);
        const int n = tbo_bf_cross(d1.data, d1.rows, d2.data, d2.rows, tmp.data());
        for (int i = 0; i < n; i++) matches.push_back(DMatch(tmp[i].queryIdx, tmp[i].trainIdx, tmp[i].imgIdx, tmp[i].distance));
    }
};

namespace flann {
struct IndexParams { virtual ~IndexParams() {} };
struct LshIndexParams : IndexParams { LshIndexParams(int, int, int) {} };
}  // namespace flann
// member of Matcher (matcher.h:152); searchByNN (approximate LSH) is out of scope and never called
struct FlannBasedMatcher {
    FlannBasedMatcher(flann::IndexParams* p = nullptr) { delete p; }
};

// cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(21,21), 3) — matcher.cpp:745
inline void calcOpticalFlowPyrLK(const Mat& prev, const Mat& next, const std::vector<Point2f>& prevPts, std::vector<Point2f>& nextPts,
                                 std::vector<uchar>& status, std::vector<float>& err, Size win, int maxLevel)
{
    assert(win.width == win.height && prev.rows == next.rows && prev.cols == next.cols && prev.step.v == next.step.v);
    tbo_lk_params p = {win.width, maxLevel, 30, 0.01f, 1e-4f};
    const int n = (int)prevPts.size();
    nextPts.resize((size_t)n); status.resize((size_t)n); err.resize((size_t)n);
    if (n == 0) return;
    tbo_lk_flow(&p, prev.data, next.data, prev.cols, prev.rows, (int)prev.step.v, (const float*)prevPts.data(), n, (float*)nextPts.data(),
                status.data(), err.data());
}

// rejectWithF's RANSAC (matcher.cpp:873) is host geometry outside the path
inline Mat findFundamentalMat(const std::vector<Point2f>&, const std::vector<Point2f>&, int, double, double, std::vector<uchar>&)
{
    std::fprintf(stderr, "cvshim: findFundamentalMat is not on the path\n");
    std::abort();
}

// drawing / display calls of the block-commented stereo matcher (LocalBA.cpp:104-105,153-158,279-286): no-ops
#define CV_GRAY2BGR 8
struct Scalar { Scalar(double = 0, double = 0, double = 0, double = 0) {} };
inline void cvtColor(const Mat&, Mat&, int) {}
template <typename P> inline void circle(Mat&, P, int, Scalar, int = 1) {}
template <typename P, typename Q> inline void line(Mat&, P, Q, Scalar, int = 1) {}

}  // namespace cv
#endif
