/*
 * HFNetB200Backends.h -- reference-side bindings of libhfnet_b200.so for the three other classes on the hot path:
 *
 *   HFNetB200Matcher            the descriptor stages of Matcher (include/Matcher.h:43-89): SearchByBoW,
 *                               SearchForTriangulation, the windowed SearchByProjection family
 *   HFNetB200KeyFrameDatabase   KeyFrameDatabase::add / erase / clear / clearMap and the scan behind DetectNBestCandidates /
 *                               DetectRelocalizationCandidates (include/KeyFrameDatabase.h:59-69)
 *   HFNetB200Frame              Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:760-825)
 *   HFNetB200Optimizer          the numeric cores of Optimizer::LocalBundleAdjustment and Optimizer::PoseOptimization
 *                               (include/Optimizer.h:59,61)
 *
 * Header-only, C++11, OpenCV core types only (cv::Mat descriptors as the reference stores them: N x 256 CV_32F,
 * continuous; 4096 x 1 CV_32F global descriptors).  The reference's methods walk KeyFrame / MapPoint / Frame objects;
 * these wrappers take exactly the arrays those walks read, so the replacement inside each reference method is the few
 * lines shown in INTEGRATION.md.  All of them borrow an hfb_ctx (one per calling thread: Tracking, LocalMapping and
 * LoopClosing each own one, like they each construct their own Matcher on the stack).
 * Build without OpenCV: define HFNET_B200_SHIM_STANDALONE and provide cv::Mat stand-ins (tests/native/).
 */
#ifndef HFNETB200BACKENDS_H
#define HFNETB200BACKENDS_H

#include <cstdint>
#include <utility>
#include <vector>

#ifndef HFNET_B200_SHIM_STANDALONE
#include <opencv2/core.hpp>
#endif
#include "hfnet_b200.h"

namespace ORB_SLAM3
{

class HFNetB200Matcher
{
public:
    explicit HFNetB200Matcher(hfb_ctx *ctx) : mCtx(ctx) {}

    /* Matcher::SearchByBoW (src/Matcher.cc:220-263, :561-621): cv::BFMatcher(NORM_L2, crossCheck).match(desc1, desc2)
     * + distance < TH_LOW.  vnMatches12[i] = row of desc2 or -1; returns the number of matches.  The caller keeps its
     * map-point bookkeeping (:236-262) around this call. */
    int SearchByBoW(const cv::Mat &desc1, const cv::Mat &desc2, std::vector<int> &vnMatches12, std::vector<float> *pvDist = nullptr,
                    float thLow = 0.6f)
    {
        return Mutual(0, desc1, desc2, thLow, vnMatches12, pvDist);
    }

    /* The descriptor stage of Matcher::SearchForTriangulation (src/Matcher.cc:845-889): D1 * D2^T, per-row arg-max above
     * 1 - 0.5 * TH_HIGH^2, column cross-check.  The epipolar test (:894-909) stays with the caller. */
    int SearchForTriangulation(const cv::Mat &desc1, const cv::Mat &desc2, std::vector<std::pair<size_t, size_t> > &vMatchedPairs,
                               float thHigh = 0.75f)
    {
        std::vector<int> m;
        const int n = Mutual(1, desc1, desc2, -0.5f * thHigh * thHigh + 1.f, m, nullptr);
        vMatchedPairs.clear();
        for (size_t i = 0; i < m.size(); ++i)
            if (m[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m[i]));
        return n;
    }

    /* The window search inside every SearchByProjection / Fuse variant (src/Matcher.cc:78-117 and the like +
     * Frame::GetFeaturesInArea, src/Frame.cc:659-725): for query i (descriptor row i of queryDesc, projection uv[i],
     * radius[i], octave range [minLevel[i], maxLevel[i]], maxLevel < 0 = unbounded) the 4 nearest frame features inside
     * its window, ascending distance.  Outputs are nq x 4 (index or -1, distance, octave). */
    bool ProjectionCandidates(const cv::Mat &queryDesc, const std::vector<float> &uv, const std::vector<float> &radius,
                              const std::vector<int> &minLevel, const std::vector<int> &maxLevel, const cv::Mat &frameDesc,
                              const std::vector<float> &frameXY, const std::vector<int> &frameOctave,
                              const std::vector<unsigned char> *pvSkip, std::vector<int> &candIdx, std::vector<float> &candDist,
                              std::vector<int> &candLevel)
    {
        const int nq = queryDesc.rows, nf = frameDesc.rows;
        candIdx.assign((size_t)nq * HFB_PROJ_TOPK, -1);
        candDist.assign((size_t)nq * HFB_PROJ_TOPK, 3.402823466e38f);
        candLevel.assign((size_t)nq * HFB_PROJ_TOPK, -1);
        if (nq == 0 || nf == 0) return true;
        return hfb_match_projection(mCtx, queryDesc.ptr<float>(), nq, uv.data(), radius.data(), minLevel.data(), maxLevel.data(),
                                    frameDesc.ptr<float>(), nf, frameXY.data(), frameOctave.data(),
                                    pvSkip ? pvSkip->data() : nullptr, candIdx.data(), candDist.data(), candLevel.data()) == HFB_OK;
    }

private:
    int Mutual(int mode, const cv::Mat &d1, const cv::Mat &d2, float thr, std::vector<int> &m12, std::vector<float> *pv)
    {
        m12.assign((size_t)d1.rows, -1);
        std::vector<float> val((size_t)d1.rows, 0.f);
        int32_t n = 0;
        if (d1.rows > 0 && d2.rows > 0)
        {
            const int rc = mode == 0 ? hfb_match_mutual_l2(mCtx, d1.ptr<float>(), d1.rows, d2.ptr<float>(), d2.rows, thr, m12.data(), val.data(), &n)
                                     : hfb_match_mutual_cos(mCtx, d1.ptr<float>(), d1.rows, d2.ptr<float>(), d2.rows, thr, m12.data(), val.data(), &n);
            if (rc != HFB_OK) return -1;
        }
        if (pv) pv->swap(val);
        return n;
    }
    hfb_ctx *mCtx;
};

class HFNetB200KeyFrameDatabase
{
public:
    HFNetB200KeyFrameDatabase(hfb_ctx *ctx, int capacity = 65536) : mDb(nullptr) { hfb_kfdb_create(ctx, HFB_GLOBAL_DIM, capacity, &mDb); }
    ~HFNetB200KeyFrameDatabase() { if (mDb) hfb_kfdb_destroy(mDb); }
    HFNetB200KeyFrameDatabase(const HFNetB200KeyFrameDatabase &) = delete;
    HFNetB200KeyFrameDatabase &operator=(const HFNetB200KeyFrameDatabase &) = delete;
    bool IsValid() const { return mDb != nullptr; }

    /* KeyFrameDatabase::add(pKF): id = pKF->mnId, mapId = pKF->GetMap()->GetId(), desc = pKF->mGlobalDescriptors. */
    bool add(long unsigned int id, long unsigned int mapId, const cv::Mat &globalDescriptor)
    {
        const int64_t i = (int64_t)id, m = (int64_t)mapId;
        return hfb_kfdb_add_tagged(mDb, &i, &m, globalDescriptor.ptr<float>(), 1) == HFB_OK;
    }
    void erase(long unsigned int id) { hfb_kfdb_erase(mDb, (int64_t)id); }
    void clear() { hfb_kfdb_clear(mDb); }
    void clearMap(long unsigned int mapId) { hfb_kfdb_clear_map(mDb, (int64_t)mapId); }
    int size() const { return hfb_kfdb_size(mDb); }

    /* The scan of DetectNBestCandidates (src/KeyFrameDatabase.cc:85-104; rel 0.8, floor 0) and of
     * DetectRelocalizationCandidates (:177-192; floor 0.5): every keyframe scored 1 - |q - d|, candidates above
     * max(floor, rel * best), ids ascending.  The covisibility accumulation (:111-137) stays with the caller, which asks
     * ScoresOf for the neighbours' mPlaceRecognitionScore. */
    bool Query(const cv::Mat &globalDescriptor, float rel, float floor, std::vector<long unsigned int> &vIds, std::vector<float> &vScores,
               float &bestScore)
    {
        const int cap = hfb_kfdb_size(mDb) > 0 ? hfb_kfdb_size(mDb) : 1;
        std::vector<int64_t> ids((size_t)cap);
        vScores.assign((size_t)cap, 0.f);
        int32_t n = 0;
        if (hfb_kfdb_query(mDb, globalDescriptor.ptr<float>(), rel, floor, ids.data(), vScores.data(), cap, &n, &bestScore) != HFB_OK) return false;
        vIds.resize((size_t)n);
        vScores.resize((size_t)n);
        for (int i = 0; i < n; ++i) vIds[i] = (long unsigned int)ids[i];
        return true;
    }
    bool ScoresOf(const std::vector<long unsigned int> &vIds, std::vector<float> &vScores)
    {
        std::vector<int64_t> ids(vIds.begin(), vIds.end());
        vScores.assign(vIds.size(), -1.f);
        return hfb_kfdb_scores_of(mDb, ids.data(), (int32_t)ids.size(), vScores.data()) == HFB_OK;
    }

private:
    hfb_kfdb *mDb;
};

/* Frame's two calibration-dependent steps.  SetCamera once per context (Frame's static mK / mDistCoef, src/Frame.cc:
 * 358-372); from then on the fused extraction also leaves mvKeysUn resident for the windowed searches. */
class HFNetB200Frame
{
public:
    static bool SetCamera(hfb_ctx *ctx, float fx, float fy, float cx, float cy, const cv::Mat &distCoef)
    {
        const float K[4] = {fx, fy, cx, cy};
        const int n = distCoef.rows * distCoef.cols;
        return hfb_set_camera(ctx, K, n > 0 ? distCoef.ptr<float>() : nullptr, n) == HFB_OK;
    }

    /* Frame::UndistortKeyPoints (src/Frame.cc:760-793): mvKeysUn = mvKeys with undistorted pt. */
    static bool UndistortKeyPoints(hfb_ctx *ctx, const std::vector<cv::KeyPoint> &vKeys, std::vector<cv::KeyPoint> &vKeysUn)
    {
        const size_t N = vKeys.size();
        std::vector<float> x(N), y(N), xu(N), yu(N);
        for (size_t i = 0; i < N; ++i) { x[i] = vKeys[i].pt.x; y[i] = vKeys[i].pt.y; }
        if (hfb_undistort_points(ctx, x.data(), y.data(), (int32_t)N, xu.data(), yu.data()) != HFB_OK) return false;
        vKeysUn = vKeys;
        for (size_t i = 0; i < N; ++i) { vKeysUn[i].pt.x = xu[i]; vKeysUn[i].pt.y = yu[i]; }
        return true;
    }

    /* Frame::ComputeImageBounds (src/Frame.cc:796-825). */
    static bool ComputeImageBounds(hfb_ctx *ctx, int cols, int rows, float &mnMinX, float &mnMaxX, float &mnMinY, float &mnMaxY)
    {
        float b[4];
        if (hfb_image_bounds(ctx, cols, rows, b) != HFB_OK) return false;
        mnMinX = b[0]; mnMaxX = b[1]; mnMinY = b[2]; mnMaxY = b[3];
        return true;
    }
};

class HFNetB200Optimizer
{
public:
    /* optimizer.optimize(10) of Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1411) on the flat problem the method
     * builds between :1120 and :1400 (poses qx qy qz qw tx ty tz of every local + fixed keyframe, points, mono edges sorted
     * by point).  vbOutlier = the test of :1417-1431 (chi2 > 5.991 or non-positive depth).  Returns false on error. */
    static bool LocalBundleAdjustment(hfb_ctx *ctx, std::vector<double> &vPoses, const std::vector<unsigned char> &vbFixed,
                                      std::vector<double> &vPoints, const std::vector<int> &vEdgeCam, const std::vector<int> &vEdgePoint,
                                      const std::vector<double> &vObs, const std::vector<double> &vInvSigma2, const float K[4],
                                      bool *pbStopFlag, std::vector<unsigned char> &vbOutlier, int nIterations = 10)
    {
        hfb_lba_problem p;
        p.n_cams = (int32_t)vbFixed.size();
        p.n_points = (int32_t)(vPoints.size() / 3);
        p.n_edges = (int32_t)vEdgeCam.size();
        p.poses = vPoses.data(); p.fixed = vbFixed.data(); p.points = vPoints.data();
        p.edge_cam = vEdgeCam.data(); p.edge_point = vEdgePoint.data(); p.obs = vObs.data(); p.inv_sigma2 = vInvSigma2.data();
        for (int i = 0; i < 4; ++i) p.K[i] = K[i];
        p.huber_delta = 2.4476519360399265;            /* sqrt(5.991), thHuberMono (src/Optimizer.cc:1206) */
        std::vector<double> poses(vPoses.size()), points(vPoints.size()), chi2((size_t)p.n_edges);
        std::vector<unsigned char> depth((size_t)p.n_edges);
        hfb_lba_stats st;
        const int rc = hfb_lba_optimize(ctx, &p, nIterations, 0.0, reinterpret_cast<const volatile uint8_t *>(pbStopFlag), poses.data(),
                                        points.data(), chi2.data(), depth.data(), &st);
        if (rc != HFB_OK) return false;
        vPoses.swap(poses);
        vPoints.swap(points);
        vbOutlier.resize((size_t)p.n_edges);
        for (int e = 0; e < p.n_edges; ++e) vbOutlier[e] = (chi2[e] > 5.991 || !depth[e]) ? 1 : 0;
        return true;
    }

    /* Optimizer::PoseOptimization (src/Optimizer.cc:814-1114), monocular: returns nInitialCorrespondences - nBad. */
    static int PoseOptimization(hfb_ctx *ctx, const float K[4], double pose[7], const std::vector<double> &vXw,
                                const std::vector<double> &vObs, const std::vector<double> &vInvSigma2,
                                std::vector<unsigned char> &vbOutlier)
    {
        const int n = (int)vInvSigma2.size();
        vbOutlier.assign((size_t)(n > 0 ? n : 1), 0);
        double out[7];
        int32_t nInl = 0, nTrials = 0;
        if (hfb_pose_optimize(ctx, K, pose, n, vXw.data(), vObs.data(), vInvSigma2.data(), out, vbOutlier.data(), &nInl, &nTrials) != HFB_OK)
            return -1;
        for (int i = 0; i < 7; ++i) pose[i] = out[i];
        vbOutlier.resize((size_t)n);
        return nInl;
    }
};

} // namespace ORB_SLAM3

#endif // HFNETB200BACKENDS_H
