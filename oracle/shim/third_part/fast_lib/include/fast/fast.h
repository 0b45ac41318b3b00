// fast/fast.h — stand-in for uzh-rpg/fast (third_part/fast_lib is ABSENT from the reference tree and has no version
// pin), so that the reference's src/extractors/FASTextractor.cpp compiles UNMODIFIED: its grid / Shi-Tomasi /
// selection logic is then the reference's own code.  ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/tb_oracle.h).
// The detector itself stays PARITY-UNPINNED: the three calls forward to tbo_fast10 (tb_oracle.c), the restatement of
// Rosten's published FAST-10 (10 contiguous ring pixels brighter than p+b or darker than p-b, score = largest b that
// keeps the corner, 3x3 non-max with >=), which already returns the surviving corners in raster order.
#ifndef TB_FAST_SHIM_H
#define TB_FAST_SHIM_H
#include <vector>
#include "../../../../../tb_oracle.h"

namespace fast {

typedef unsigned char fast_byte;
struct fast_xy {
    short x, y;
    fast_xy(short x_, short y_) : x(x_), y(y_) {}
};

// scores of the corners returned by the last detect call (the reference calls detect, score, nonmax back to back)
inline std::vector<int>& tb_last_scores() { static thread_local std::vector<int> s; return s; }

inline void fast_corner_detect_10_sse2(const fast_byte* img, int w, int h, int stride, short barrier, std::vector<fast_xy>& corners)
{
    const int cap = w * h / 2 + 16;
    std::vector<int32_t> xy(2 * (size_t)cap), sc((size_t)cap);
    const int n = tbo_fast10(img, w, h, stride, barrier, xy.data(), sc.data(), cap);
    corners.clear();
    tb_last_scores().assign(sc.begin(), sc.begin() + n);
    for (int i = 0; i < n; i++) corners.push_back(fast_xy((short)xy[2 * i], (short)xy[2 * i + 1]));
}
inline void fast_corner_detect_10(const fast_byte* img, int w, int h, int stride, short barrier, std::vector<fast_xy>& corners)
{
    fast_corner_detect_10_sse2(img, w, h, stride, barrier, corners);
}
inline void fast_corner_score_10(const fast_byte*, const int, const std::vector<fast_xy>& corners, const int, std::vector<int>& scores)
{
    scores = tb_last_scores();
    scores.resize(corners.size());
}
inline void fast_nonmax_3x3(const std::vector<fast_xy>& corners, const std::vector<int>&, std::vector<int>& nonmax_corners)
{
    nonmax_corners.clear();
    for (size_t i = 0; i < corners.size(); i++) nonmax_corners.push_back((int)i);  // tbo_fast10 already suppressed
}

}  // namespace fast
#endif
