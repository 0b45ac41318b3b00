// Frame.h — facade of the reference's include/types/Frame.h:53-202 restricted to what the feature
// tracking path touches: construction (scale tables + pyramid, src/types/Frame.cpp:13-33, :414-427),
// keypoints / descriptors (SetKeys :69), the pose, the flattened map-point side data the projection
// matchers read, and the search-grid query (AssignFeaturesToGrid / GetFeaturesInArea :187-265, run
// on the GPU), and the bag-of-words vectors (SetBow :267-270: the vocabulary walk runs on the GPU).
// Eigen / Map bookkeeping of the reference Frame are out of scope.
#ifndef TRACKING_BENCH_FRAME_H
#define TRACKING_BENCH_FRAME_H
#include <cmath>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
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

// ORBVocabulary (reference include/types/Frame.h:22 = DBoW2::TemplatedVocabulary<FORB>): the text-file loader
// (third_part/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1420) keeps the tree as flat arrays and hands it to the device;
// transform() is Frame::SetBow's call (TemplatedVocabulary.h:1127-1194) with the tree walk on the GPU and the
// BowVector accumulation / normalisation (BowVector.cpp:62-84) on the host in the reference's order.
typedef std::map<unsigned int, double> BowVector;                       // DBoW2::BowVector
typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector;   // DBoW2::FeatureVector
class ORBVocabulary {
public:
    ORBVocabulary() = default;
    ~ORBVocabulary() { if (voc_) tb_vocab_destroy(voc_); }
    ORBVocabulary(const ORBVocabulary &) = delete;
    ORBVocabulary &operator=(const ORBVocabulary &) = delete;
    bool empty() const { return voc_ == nullptr; }
    unsigned int size() const { return n_words_; }
    bool loadFromTextFile(const std::string &filename)
    {
        std::ifstream f(filename.c_str());
        if (!f.is_open()) return false;
        std::string s;
        std::getline(f, s);
        std::stringstream ss(s);
        int n1 = -1, n2 = -1;
        k_ = L_ = -1;
        ss >> k_ >> L_ >> n1 >> n2;
        if (k_ < 0 || k_ > 20 || L_ < 1 || L_ > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;   // :1359
        scoring_ = n1; weighting_ = n2;
        std::vector<int32_t> parent(1, -1);
        std::vector<uint8_t> leaf(1, 0), desc(32, 0);
        std::vector<double> weight(1, 0.0);
        while (std::getline(f, s)) {
            std::stringstream sn(s);
            int pid, isLeaf;
            if (!(sn >> pid >> isLeaf)) continue;   // blank line: see trackingbench_slam_b200/vocab.py:save_text
            parent.push_back(pid); leaf.push_back(isLeaf > 0 ? 1 : 0);
            for (int i = 0; i < 32; i++) { int b = 0; sn >> b; desc.push_back((uint8_t)b); }
            double w = 0; sn >> w; weight.push_back(w);
        }
        n_words_ = 0;
        for (size_t i = 0; i < leaf.size(); i++) n_words_ += leaf[i];
        tb_ctx *ctx = tb_default_ctx();
        if (voc_) { tb_vocab_destroy(voc_); voc_ = nullptr; }
        return tb_vocab_create(ctx, (int)parent.size(), L_, parent.data(), leaf.data(), desc.data(), weight.data(), &voc_) == TB_OK;
    }
    // transform(features, v, fv, levelsup): descriptors as one contiguous n x 32 cv::Mat (Frame keeps them that way)
    void transform(const cv::Mat &descriptors, BowVector &v, FeatureVector &fv, int levelsup) const
    {
        v.clear(); fv.clear();
        if (empty() || descriptors.rows == 0) return;
        tb_ctx *ctx = tb_default_ctx();
        const int n = descriptors.rows;
        cv::Mat d = ((size_t)descriptors.step == (size_t)descriptors.cols) ? descriptors : descriptors.clone();
        std::vector<int32_t> word(n), node(n);
        std::vector<double> w(n);
        tb_check(tb_bow_transform(ctx, voc_, d.data, n, levelsup, word.data(), node.data(), w.data()), ctx, "tb_bow_transform");
        const bool tf = (weighting_ == 0 || weighting_ == 1);          // TF_IDF, TF accumulate; IDF, BINARY keep the first
        for (int i = 0; i < n; i++) {
            if (!(w[i] > 0)) continue;                                  // stopped word (:1157, :1185)
            if (tf) v[(unsigned)word[i]] += w[i];
            else if (v.find((unsigned)word[i]) == v.end()) v[(unsigned)word[i]] = w[i];
            fv[(unsigned)node[i]].push_back((unsigned)i);
        }
        const bool must = (scoring_ == 0 || scoring_ == 1);            // L1_NORM, L2_NORM
        if (tf && !v.empty() && !must) { const double nd = (double)v.size(); for (auto &e : v) e.second /= nd; }
        if (must) {
            double norm = 0.0;
            if (scoring_ == 0) for (auto &e : v) norm += std::fabs(e.second);
            else { for (auto &e : v) norm += e.second * e.second; norm = std::sqrt(norm); }
            if (norm > 0.0) for (auto &e : v) e.second /= norm;
        }
    }
    tb_vocab *handle() const { return voc_; }

private:
    tb_vocab *voc_ = nullptr;
    int k_ = 0, L_ = 0, scoring_ = 0, weighting_ = 0;
    unsigned int n_words_ = 0;
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
        mvLevelSigma2.assign(nLevels, 1.f); mvInvLevelSigma2.assign(nLevels, 1.f);           // Frame.cpp:20-28
        for (int i = 1; i < nLevels; i++) { mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.f / mvLevelSigma2[i]; }
        mfGridElementHeightInv = static_cast<float>(FRAME_GRID_COLS) / (float)imGray.cols;    // Frame.cpp:30-31 (names as in the reference)
        mfGridElementWidthInv = static_cast<float>(FRAME_GRID_ROWS) / (float)imGray.rows;
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
    cv::Mat GetDescriptor(int id) const { return mDescriptors.row(id); }
    std::vector<cv::Mat> GetVectorDescriptors() const { std::vector<cv::Mat> v; v.reserve(mDescriptors.rows); for (int j = 0; j < mDescriptors.rows; j++) v.push_back(mDescriptors.row(j)); return v; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
    // Frame.cpp:257-265
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY) const
    {
        posX = (int)std::round((kp.pt.x) * mfGridElementWidthInv);
        posY = (int)std::round((kp.pt.y) * mfGridElementHeightInv);
        return !(posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS);
    }
    bool IsInImage(const float &, const float &) const { return false; }   // Frame.cpp:364-368: the reference's stub
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
    // bag of words (Frame.cpp:267-270)
    void SetBow(const std::shared_ptr<ORBVocabulary> &voc) { voc->transform(mDescriptors, mBowVec, mFeatVec, 4); }
    const BowVector &GetBowVector() const { return mBowVec; }
    const FeatureVector &GetFeatureVector() const { return mFeatVec; }
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
    // Frame::Equalize (Frame.cpp:453-458): cv::createCLAHE(3.0, Size(8,8))->apply(level 0)
    cv::Mat Equalize()
    {
        tb_ctx *ctx = tb_default_ctx();
        const cv::Mat &src = mvImagePyramid[0];
        cv::Mat dst(src.rows, src.cols, CV_8U);
        tb_check(tb_clahe_host(ctx, src.data, src.cols, src.rows, src.step, dst.data, dst.step, 3.0, 8, 8), ctx, "tb_clahe_host");
        return dst;
    }
    // Frame::UnDistortPoints (Frame.cpp:165-171): px_un of every key through PinholeCamera::UndistortPoint
    void UnDistortPoints()
    {
        auto cam = std::dynamic_pointer_cast<PinholeCamera>(mpCamera);
        mvKeysUn = cam ? cam->UndistortPoints(GetCVKeys()) : GetCVKeys();
    }
    const std::vector<cv::Point2f> &GetUndistortedKeys() const { return mvKeysUn; }
    // Frame::UnDistortImage (Frame.cpp:173-176) on level 0
    void UnDistortImage() { cv::Mat out; mpCamera->UndistortImage(mvImagePyramid[0], out); mvImagePyramid[0] = out; }
    // Frame::GetCVKeys: the keypoint positions as cv::Point2f (the prevPts of searchByOPFlow)
    std::vector<cv::Point2f> GetCVKeys() { std::vector<cv::Point2f> v; v.reserve(mvKeys.size()); for (const auto &k : mvKeys) v.push_back(k.pt); return v; }
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
    std::vector<cv::Point2f> mvKeysUn;
    cv::Mat mDescriptors;
    std::vector<std::shared_ptr<MapPointLite>> mvpMapPoints;
    BowVector mBowVec;
    FeatureVector mFeatVec;
    std::shared_ptr<CameraModel> mpCamera;
    int nLevels;
    float scaleFactor;
    int cols_ = 0, rows_ = 0;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    float mfGridElementWidthInv = 0.f, mfGridElementHeightInv = 0.f;
    std::vector<cv::Mat> mvImagePyramid;
};
}  // namespace TRACKING_BENCH
#endif  // TRACKING_BENCH_FRAME_H
