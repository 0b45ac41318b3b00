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
#define CV_16SC2 11
#define CV_32F 5
#define CV_32FC1 5
#define CV_32FC2 13
#define CV_64F 6

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
    size_t buf[2];
    size_t* p;  // OpenCV exposes step.p[0] (FASTextractor.cpp:104)
    MatStep(size_t s = 0) : v(s), p(buf) { buf[0] = s; buf[1] = 0; }
    MatStep(const MatStep& o) : v(o.v), p(buf) { buf[0] = o.buf[0]; buf[1] = o.buf[1]; }
    MatStep& operator=(const MatStep& o) { v = o.v; buf[0] = o.buf[0]; buf[1] = o.buf[1]; return *this; }
    MatStep& operator=(size_t s) { v = s; buf[0] = s; return *this; }
    operator size_t() const { return v; }
};

struct MatZeros { int rows, cols, type; };

// bytes per element of a CV_MAKETYPE(depth, channels) type code
static inline size_t tb_elem_size(int type)
{
    static const size_t depth_size[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return depth_size[type & 7] * (size_t)((type >> 3) + 1);
}

class Mat {
public:
    std::shared_ptr<uchar> buf;
    uchar* data;
    int rows, cols;
    MatStep step;
    int tp;
    Mat() : data(nullptr), rows(0), cols(0), step(0), tp(0) {}
    Mat(int r, int c, int type) : data(nullptr), rows(0), cols(0), step(0), tp(0) { create(r, c, type); }
    Mat(Size sz, int type) : data(nullptr), rows(0), cols(0), step(0), tp(0) { create(sz.height, sz.width, type); }
    // wrap external memory (no ownership); s == 0 is OpenCV's AUTO_STEP
    Mat(int r, int c, int type, void* p, size_t s = 0) : data((uchar*)p), rows(r), cols(c), step(s ? s : (size_t)c * tb_elem_size(type)), tp(type) {}
    size_t elemSize() const { return tb_elem_size(tp); }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == tp) return;  // cv::Mat::create keeps the buffer when size/type match
        const size_t esz = tb_elem_size(type);
        buf.reset(new uchar[(size_t)r * c * esz + 64], std::default_delete<uchar[]>());
        data = buf.get(); rows = r; cols = c; tp = type; step = (size_t)c * esz;
    }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return tp; }
    size_t step1() const { return step.v; }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step.v + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step.v + (size_t)c * sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.v); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.v); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step.v; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat clone() const {
        Mat m; m.create(rows, cols, tp);
        for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& m) const {
        m.create(rows, cols, tp);
        for (int r = 0; r < rows; r++) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
    }
    static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }
    Mat& operator=(const MatZeros& z) {  // MatExpr assignment: create() (no realloc on match) then fill
        create(z.rows, z.cols, z.type);
        for (int r = 0; r < rows; r++) std::memset(ptr(r), 0, (size_t)cols * elemSize());
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
