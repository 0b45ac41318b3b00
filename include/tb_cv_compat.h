// tb_cv_compat.h — the few OpenCV types the reference's extractor / matcher / Frame interfaces use
// (cv::Mat for 8-bit 2-D images and N x 32 descriptor matrices, cv::KeyPoint, cv::DMatch, cv::Point2f,
// cv::Size).  When a real OpenCV is installed (<opencv2/core.hpp> found) it is used instead and this
// header defines nothing.  Layouts match OpenCV's so the PODs can be handed to the C ABI unchanged.
#ifndef TB_CV_COMPAT_H
#define TB_CV_COMPAT_H
#if defined(__has_include) && __has_include(<opencv2/core.hpp>) && !defined(TB_FORCE_CV_COMPAT)
#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>
#else
#define TB_USING_CV_COMPAT 1
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
typedef unsigned char uchar;
#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#endif
namespace cv {
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct KeyPoint {   // 7 x 4 bytes, same order as OpenCV
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
};
// 2-D, 8-bit, single channel, reference counted; enough for images and descriptor matrices.
class Mat {
public:
    int rows, cols;
    uchar *data;
    size_t step;
    Mat() : rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int /*type*/) : rows(0), cols(0), data(nullptr), step(0) { create(r, c, CV_8U); }
    Mat(int r, int c, int /*type*/, void *ext, size_t s = 0) : rows(r), cols(c), data((uchar *)ext), step(s ? s : (size_t)c) {}
    void create(int r, int c, int /*type*/)
    {
        if (data && r == rows && c == cols && buf_) return;
        buf_.reset(new uchar[(size_t)r * c + 64], std::default_delete<uchar[]>());
        data = buf_.get(); rows = r; cols = c; step = (size_t)c;
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T &at(int r, int c) { return *(T *)(data + (size_t)r * step + c * sizeof(T)); }
    Mat row(int r) const { Mat m(*this); m.data = data + (size_t)r * step; m.rows = 1; return m; }
    Mat clone() const
    {
        Mat m; m.create(rows, cols, CV_8U);
        for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols);
        return m;
    }
private:
    std::shared_ptr<uchar> buf_;
};
typedef Mat &OutputArray;
}  // namespace cv
#endif
#endif  // TB_CV_COMPAT_H
