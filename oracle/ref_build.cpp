// ref_build.cpp — builds oracle/_ref/libtb_ref.so: the reference's OWN extractor source
// (/root/reference/src/extractors/ORBextractor.cpp) compiled UNMODIFIED from where it lies, against
// oracle/shim.  ORACLE / TEST INFRASTRUCTURE ONLY.  No reference source is copied into this repo.
//
// One deliberate pin: ORBextractor.cpp:643 sorts vector<pair<int, ExtractorNode*>> with the default
// operator<, i.e. ties on node size are broken by HEAP ADDRESS, which makes the reference's output
// differ run to run.  `sort` is redirected to tb_sort, which orders by
// (size, UL.x, UL.y) — same rule as tb_oracle.c:sz_cmp.  Build with -DTB_REF_UNPINNED to get the
// raw, address-ordered behaviour (used by a test that documents the drift).
#include <algorithm>
#include <list>
#include <utility>
#include <vector>
#include "extractors/ORBextractor.h"

namespace TRACKING_BENCH {
template <typename It> static void tb_sort(It first, It last)
{
#ifdef TB_REF_UNPINNED
    std::sort(first, last);
#else
    std::sort(first, last, [](const std::pair<int, ExtractorNode*>& a, const std::pair<int, ExtractorNode*>& b) {
        if (a.first != b.first) return a.first < b.first;
        if (a.second->UL.x != b.second->UL.x) return a.second->UL.x < b.second->UL.x;
        return a.second->UL.y < b.second->UL.y;
    });
#endif
}
}  // namespace TRACKING_BENCH

#define sort tb_sort
#include "src/extractors/ORBextractor.cpp"
#undef sort

using namespace TRACKING_BENCH;

static int run(bool add_points, ORBExtractor* ex, const uint8_t* const* levels, const int* w, const int* h,
               const int* stride, int n_levels, const float* sf, int target, float initTh, float minTh,
               const tbo_kp* exit_kps, int n_exit, tbo_kp* kps, uint8_t* desc, int cap)
{
    std::vector<cv::Mat> images;
    for (int l = 0; l < n_levels; l++) images.push_back(cv::Mat(h[l], w[l], CV_8UC1, (void*)levels[l], (size_t)stride[l]));
    std::vector<float> vsf(sf, sf + n_levels);
    std::vector<cv::KeyPoint> out;
    cv::Mat d;
    if (!add_points) {
        (*ex)(images, vsf, target, initTh, minTh, out, d);
    } else {
        std::vector<cv::KeyPoint> ek;
        for (int i = 0; i < n_exit; i++)
            ek.push_back(cv::KeyPoint(exit_kps[i].x, exit_kps[i].y, exit_kps[i].size, exit_kps[i].angle,
                                      exit_kps[i].response, exit_kps[i].octave, exit_kps[i].class_id));
        cv::_OutputArray oa(d);
        ex->AddPoints(images, vsf, target, initTh, minTh, ek, out, oa);
    }
    int n = (int)out.size();
    for (int i = 0; i < n && i < cap; i++) {
        tbo_kp k = {out[i].pt.x, out[i].pt.y, out[i].size, out[i].angle, out[i].response, out[i].octave, out[i].class_id};
        kps[i] = k;
        std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    }
    return n;
}

extern "C" {
void* tbref_extractor_new() { return new ORBExtractor(); }
void tbref_extractor_delete(void* p) { delete (ORBExtractor*)p; }
// ORBExtractor::operator() — ORBextractor.cpp:906
int tbref_extract(void* ex, const uint8_t* const* levels, const int* w, const int* h, const int* stride, int n_levels,
                  const float* sf, int target, float initTh, float minTh, tbo_kp* kps, uint8_t* desc, int cap)
{
    return run(false, (ORBExtractor*)ex, levels, w, h, stride, n_levels, sf, target, initTh, minTh, nullptr, 0, kps, desc, cap);
}
// ORBExtractor::AddPoints — ORBextractor.cpp:840 (reuses the budget of the previous operator() call)
int tbref_add_points(void* ex, const uint8_t* const* levels, const int* w, const int* h, const int* stride, int n_levels,
                     const float* sf, int target, float initTh, float minTh, const tbo_kp* exit_kps, int n_exit,
                     tbo_kp* kps, uint8_t* desc, int cap)
{
    return run(true, (ORBExtractor*)ex, levels, w, h, stride, n_levels, sf, target, initTh, minTh, exit_kps, n_exit, kps, desc, cap);
}
}
