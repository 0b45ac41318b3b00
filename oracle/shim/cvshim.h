// cvshim.h — minimal stand-in for the handful of OpenCV 3.x types/functions that the reference's
// src/extractors/ORBextractor.{h,cpp} use, so that file can be compiled UNMODIFIED in a container
// without OpenCV.  ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/tb_oracle.h).  The three image
// primitives (FAST, GaussianBlur, fastAtan2) forward to the C restatements in tb_oracle.c, each of
// which is checked bit-exact against cv2 4.13 in tests/test_oracle_cv2.py.
#ifndef TB_CVSHIM_H
#define TB_CVSHIM_H
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <emmintrin.h>
#include "../tb_oracle.h"

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

static inline int cvRound(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }
static inline int cvRound(float v) { return _mm_cvtss_si32(_mm_set_ss(v)); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (v < i); }
static inline int cvCeil(double v) { int i = (int)v; return i + (v > i); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    template <typename A, typename B> Point_(A a, B b) : x((T)a), y((T)b) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

enum { BORDER_REFLECT_101 = 4 };

struct MatStep {
    size_t v;
    MatStep(size_t s = 0) : v(s) {}
    operator size_t() const { return v; }
};

struct MatZeros { int rows, cols; };

class Mat {
public:
    std::shared_ptr<uchar> buf;
    uchar* data;
    int rows, cols;
    MatStep step;
    Mat() : data(nullptr), rows(0), cols(0), step(0) {}
    Mat(int r, int c, int /*type*/) : data(nullptr), rows(0), cols(0), step(0) { create(r, c, 0); }
    // wrap external memory (no ownership)
    Mat(int r, int c, int /*type*/, void* p, size_t s) : data((uchar*)p), rows(r), cols(c), step(s) {}
    void create(int r, int c, int /*type*/) {
        if (data && r == rows && c == cols) return;  // cv::Mat::create keeps the buffer when size/type match
        buf.reset(new uchar[(size_t)r * c + 64], std::default_delete<uchar[]>());
        data = buf.get(); rows = r; cols = c; step = (size_t)c;
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step.v; }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step.v + c); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step.v + c); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.v); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.v); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step.v; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
    Mat clone() const {
        Mat m; m.create(rows, cols, 0);
        for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols);
        return m;
    }
    static MatZeros zeros(int r, int c, int) { return MatZeros{r, c}; }
    Mat& operator=(const MatZeros& z) {  // MatExpr assignment: create() (no realloc on match) then fill
        create(z.rows, z.cols, 0);
        for (int r = 0; r < rows; r++) std::memset(ptr(r), 0, (size_t)cols);
        return *this;
    }
};

class _OutputArray {
public:
    Mat* m;
    _OutputArray(Mat& mm) : m(&mm) {}
    void release() const { m->release(); }
    void create(int r, int c, int t) const { m->create(r, c, t); }
    Mat getMat() const { return *m; }
};
typedef const _OutputArray& OutputArray;

inline float fastAtan2(float y, float x) { return tbo_fast_atan2(y, x); }

inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmax) {
    kps.clear();
    int cap = (img.rows * img.cols) / 2 + 16;
    std::vector<tbo_kp> tmp((size_t)cap);
    int n = tbo_fast9(img.data, (int)img.step.v, img.cols, img.rows, threshold, nonmax ? 1 : 0, tmp.data(), cap);
    for (int i = 0; i < n; i++) kps.push_back(KeyPoint(tmp[i].x, tmp[i].y, tmp[i].size, tmp[i].angle, tmp[i].response));
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size ks, double sx, double sy, int border) {
    assert(ks.width == 7 && ks.height == 7 && sx == 2 && sy == 2 && border == BORDER_REFLECT_101);
    (void)ks; (void)sx; (void)sy; (void)border;
    Mat in = (src.data == dst.data) ? src.clone() : src;
    dst.create(src.rows, src.cols, 0);
    tbo_gauss7x7(in.data, in.cols, in.rows, (int)in.step.v, dst.data, (int)dst.step.v);
}

}  // namespace cv
#endif
