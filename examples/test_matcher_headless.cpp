// test_matcher_headless.cpp — headless restatement of the reference's test/test_matcher.cpp call
// sequence (Frame(5,0.8) -> ORBExtractor()(...,1000,80,30) -> SetKeys -> searchByBF(F1,F2,0,5,10,30)
// -> searchByViolence -> searchByProjection) on data/left.png / right.png, written against the SAME
// class interfaces (include/types, include/extractors, include/matchers).  Instead of imshow it prints
// counts and FNV-1a digests that tests/test_gpu_facade.py compares with the oracle.
//   usage: test_matcher_headless left.pgm right.pgm disparity.pgm   (binary P5 files)
#include <cstdio>
#include <fstream>
#include <string>
#include "camera/CameraModel.h"
#include "extractors/ORBextractor.h"
#include "matchers/matcher.h"
#include "types/Frame.h"
using namespace TRACKING_BENCH;

static cv::Mat read_pgm(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    std::string magic; int w, h, maxv;
    f >> magic >> w >> h >> maxv;
    f.get();
    cv::Mat m(h, w, CV_8U);
    f.read((char *)m.data, (size_t)w * h);
    if (!f || magic != "P5") { std::fprintf(stderr, "bad pgm %s\n", path); std::exit(2); }
    return m;
}
static unsigned long long fnv(const void *p, size_t n, unsigned long long h = 1469598103934665603ull)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static unsigned long long digest(const std::vector<cv::DMatch> &m)
{
    unsigned long long h = 1469598103934665603ull;
    for (auto &x : m) { int v[3] = {x.queryIdx, x.trainIdx, x.imgIdx}; h = fnv(v, sizeof v, h); h = fnv(&x.distance, 4, h); }
    return h;
}

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: %s left.pgm right.pgm disparity.pgm\n", argv[0]); return 2; }
    cv::Mat img1 = read_pgm(argv[1]), img2 = read_pgm(argv[2]), disp = read_pgm(argv[3]);
    auto camera_ptr = std::make_shared<PinholeCamera>(img1.cols, img1.rows, 718.856f, 718.856f, 607.1928f, 185.2157f);
    auto frame1_ptr = std::make_shared<Frame>(img1, 0, 5, 0.8f, camera_ptr);
    auto frame2_ptr = std::make_shared<Frame>(img2, 0, 5, 0.8f, camera_ptr);
    auto extractor_ptr = std::make_shared<ORBExtractor>();
    std::vector<cv::KeyPoint> keypoints1, keypoints2;
    cv::Mat descriptors1, descriptors2;
    extractor_ptr->operator()(frame1_ptr->GetImagePyramid(), frame1_ptr->GetScaleFactors(), 1000, 80, 30, keypoints1, descriptors1);
    frame1_ptr->SetKeys(keypoints1, descriptors1);
    extractor_ptr->operator()(frame2_ptr->GetImagePyramid(), frame2_ptr->GetScaleFactors(), 1000, 80, 30, keypoints2, descriptors2);
    frame2_ptr->SetKeys(keypoints2, descriptors2);
    std::printf("kps %zu %zu\n", keypoints1.size(), keypoints2.size());
    std::printf("kps_digest %llx %llx\n", fnv(keypoints1.data(), keypoints1.size() * sizeof(cv::KeyPoint)), fnv(keypoints2.data(), keypoints2.size() * sizeof(cv::KeyPoint)));
    std::printf("desc_digest %llx %llx\n", fnv(descriptors1.data, (size_t)descriptors1.rows * 32), fnv(descriptors2.data, (size_t)descriptors2.rows * 32));

    auto matcher_ptr = std::make_shared<Matcher>();
    auto matches = matcher_ptr->searchByBF(frame1_ptr, frame2_ptr, 0, 5, 10, 30);
    std::printf("bf %zu %llx\n", matches.size(), digest(matches));

    frame2_ptr->AssignFeaturesToGrid();
    matcher_ptr->setViolenceParam(30, 100, 30, true, 5);
    matches = matcher_ptr->searchByViolence(frame1_ptr, frame2_ptr, 0, 5, 50);
    std::printf("violence %zu %llx\n", matches.size(), digest(matches));

    // stereo map points from the left keypoints (test_vo.cpp:414,496-498), right camera 0.573 m to the right
    const float fx = 718.856f, fy = 718.856f, cx = 607.1928f, cy = 185.2157f, base = 0.573f;
    for (size_t i = 0; i < keypoints1.size(); i++) {
        int u = (int)keypoints1[i].pt.x, v = (int)keypoints1[i].pt.y;
        float d = (float)disp.ptr(v)[u];
        if (d <= 0) continue;
        auto mp = std::make_shared<MapPointLite>();
        float depth = (base * fx) / d;
        mp->pos[0] = (keypoints1[i].pt.x - cx) / fx * depth; mp->pos[1] = (keypoints1[i].pt.y - cy) / fy * depth; mp->pos[2] = depth;
        std::memcpy(mp->desc, descriptors1.ptr((int)i), 32);
        frame1_ptr->AddMapPoint(mp, i);
    }
    float T[16] = {1, 0, 0, -base, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    frame2_ptr->SetPose(T);
    matcher_ptr->setProjectionParam(30, 50, 30, true, 30);
    matches = matcher_ptr->searchByProjection(frame2_ptr, frame1_ptr);
    std::printf("projection %zu %llx\n", matches.size(), digest(matches));
    std::printf("dist01 %d\n", Matcher::DescriptorDistance(descriptors1.row(0), descriptors1.row(1)));

    // AddPoints after the extractor's workspace was re-sized by a larger target: the per-level budget of the last operator()
    // (mnFeaturesPerLevel) must survive, as it does in the reference (test_frame.cpp:36 flow)
    {
        std::vector<cv::KeyPoint> exitPoints, newPoints;
        for (size_t i = 0; i < keypoints1.size(); i += 5) exitPoints.push_back(keypoints1[i]);
        cv::Mat nd;
        extractor_ptr->operator()(frame1_ptr->GetImagePyramid(), frame1_ptr->GetScaleFactors(), 1000, 80, 30, keypoints1, descriptors1);
        std::vector<float> sfv = frame1_ptr->GetScaleFactors();   // AddPoints takes the factors by non-const reference (ORBextractor.h:58)
        extractor_ptr->AddPoints(frame1_ptr->GetImagePyramid(), sfv, 1500, 80, 30, exitPoints, newPoints, nd);
        std::printf("addpoints %zu %llx\n", newPoints.size(), fnv(newPoints.data(), newPoints.size() * sizeof(cv::KeyPoint)));
        ExtractorNode root, c1, c2, c3, c4;
        root.UL = cv::Point2i(0, 0); root.UR = cv::Point2i(101, 0); root.BL = cv::Point2i(0, 57); root.BR = cv::Point2i(101, 57);
        for (auto &k : keypoints1) if (k.pt.x < 101 && k.pt.y < 57) root.vKeys.push_back(k);
        root.DivideNode(c1, c2, c3, c4);
        std::printf("dividenode %zu %zu %zu %zu %zu %d %d\n", root.vKeys.size(), c1.vKeys.size(), c2.vKeys.size(), c3.vKeys.size(), c4.vKeys.size(), c1.UR.x, c1.BL.y);
    }

    // a distorted camera: PinholeCamera::UndistortImage / UndistortPoint (CameraModel.cpp:96-121, test_vo.cpp:178)
    {
        PinholeCamera dcam(img1.cols, img1.rows, 718.856f, 718.856f, 607.1928f, 185.2157f, -0.28f, 0.07f, 0.0002f, 0.00002f, 0.f);
        cv::Mat rect;
        dcam.UndistortImage(img1, rect);
        const std::vector<cv::Point2f> un = dcam.UndistortPoints(frame1_ptr->GetCVKeys());
        std::printf("undistort %llx %llx\n", fnv(rect.data, (size_t)rect.rows * rect.cols), fnv(un.data(), un.size() * sizeof(cv::Point2f)));
    }

    // optical flow; LocalBA.cpp:54 and test_vo.cpp:716 call searchByOPFlow(.., equalized = true, reject = true)
    for (int eq = 0; eq < 2; eq++) {
        std::vector<cv::Point2f> cur_points;
        matches = matcher_ptr->searchByOPFlow(frame2_ptr, frame1_ptr, cur_points, eq != 0, false);
        unsigned long long h = fnv(cur_points.data(), cur_points.size() * sizeof(cv::Point2f));
        for (const auto &m : matches) h = fnv(&m.queryIdx, sizeof m.queryIdx, h);
        std::printf("opflow%d %zu %llx\n", eq, matches.size(), h);
    }
    {
        std::vector<cv::Point2f> cur_points;
        matches = matcher_ptr->searchByOPFlow(frame2_ptr, frame1_ptr, cur_points, true, true);
        unsigned long long h = fnv(cur_points.data(), cur_points.size() * sizeof(cv::Point2f));
        for (const auto &m : matches) h = fnv(&m.queryIdx, sizeof m.queryIdx, h);
        std::printf("opflow_reject %zu %llx\n", matches.size(), h);
    }

    // bag of words (test_vo.cpp:661-705 loads ORBvoc.txt and calls SetBow per keyframe); optional 4th argument
    if (argc > 4) {
        auto vocabulary = std::make_shared<ORBVocabulary>();
        if (!vocabulary->loadFromTextFile(argv[4])) { std::fprintf(stderr, "Failed to open vocabulary\n"); return 2; }
        frame1_ptr->SetBow(vocabulary);
        frame2_ptr->SetBow(vocabulary);
        unsigned long long h = 1469598103934665603ull;
        for (const auto &e : frame1_ptr->GetBowVector()) {
            h = fnv(&e.first, sizeof e.first, h);
            h = fnv(&e.second, sizeof e.second, h);
        }
        std::printf("bowvec %zu %llx\n", frame1_ptr->GetBowVector().size(), h);
        matcher_ptr->setBowParam(50, 100, 30, true, 0.9f);
        matches = matcher_ptr->searchByBow(frame1_ptr, frame2_ptr);
        std::printf("bow %zu %llx\n", matches.size(), digest(matches));
        matches = matcher_ptr->searchByBow(frame2_ptr, frame1_ptr, true);   // only features of F1 that carry a map point
        std::printf("bow_mp %zu %llx\n", matches.size(), digest(matches));
    }
    return 0;
}
