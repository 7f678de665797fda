/*
 * HFNetB200Model.h -- the reference-side binding of libhfnet_b200.so for the extractor: a header-only C++ shim that a
 * maintainer of LiuLimingCode/HFNet_SLAM drops into include/Extractors/ next to HFNetRTModel.h.  It implements the
 * reference's own plugin interface
 *
 *     class BaseModel { virtual bool Detect(image, vKeyPoints, localDescriptors, globalDescriptors, n, thr) = 0; ... }
 *                                                                    (include/Extractors/BaseModel.h:38-54)
 *
 * on top of the plain C entry points of hfnet_b200.h, so HFextractor (src/Extractors/HFextractor.cc) and every call
 * site of GetModelVec() keep working unmodified.  See INTEGRATION.md for the three-line change in BaseModel.cc.
 *
 *   HFNetB200Engine              ONE hfb_ctx (all pyramid levels, one copy of the weights) shared by the level objects
 *   HFNetB200Model               the per-level BaseModel facade InitAllModels() puts into gvpModels
 *   InitB200Models()             InitAllModels(path, type, ImSize, nLevels, scaleFactor) for this back-end
 *   HFNetB200Model::ExtractPyramid   HFextractor::operator() as ONE fused multi-level call (one CUDA graph, one H2D,
 *                                one D2H) -- the branch a maintainer adds at the top of HFextractor::operator()
 *
 * It only needs <opencv2/core.hpp> and the reference's BaseModel.h.  To build it in an image without OpenCV, define
 * HFNET_B200_SHIM_STANDALONE and provide cv::Mat / cv::KeyPoint / BaseModel stand-ins (tests/native/).
 */
#ifndef HFNETB200MODEL_H
#define HFNETB200MODEL_H

#include <cmath>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#ifndef HFNET_B200_SHIM_STANDALONE
#include <opencv2/core.hpp>
#include "Extractors/BaseModel.h"
#endif
#include "hfnet_b200.h"

namespace ORB_SLAM3
{

/* One hfb_ctx for all levels of one extractor (the reference keeps one TensorRT engine per level,
 * src/Extractors/BaseModel.cc:35-65).  Detect on different level objects may be called concurrently from
 * cv::parallel_for_ workers (HFextractor.cc:228-243): calls into the one context are serialised here. */
struct HFNetB200Engine
{
    hfb_ctx* ctx = nullptr;
    std::mutex mutex;
    std::string error;
    int nLevels = 1;
    float scaleFactor = 1.2f;
    int height = 0, width = 0;
    bool valid = false;

    HFNetB200Engine(const std::vector<unsigned char> &weights, int imHeight, int imWidth, int levels, float scale,
                    int maxKeypointsPerLevel = 8192, int device = 0)
        : nLevels(levels), scaleFactor(scale), height(imHeight), width(imWidth)
    {
        hfb_config cfg;
        cfg.device = device;
        cfg.height = imHeight;
        cfg.width = imWidth;
        cfg.n_levels = levels;
        cfg.scale_factor = scale;
        cfg.max_keypoints = maxKeypointsPerLevel;
        cfg.max_batch = 1;
        cfg.with_global = 1;
        valid = hfb_create(&cfg, &ctx) == HFB_OK && hfb_load_weights(ctx, weights.data(), weights.size()) == HFB_OK;
        if (!valid) error = ctx ? hfb_last_error(ctx) : "hfb_create failed";
    }
    ~HFNetB200Engine() { if (ctx) hfb_destroy(ctx); }
    HFNetB200Engine(const HFNetB200Engine &) = delete;
    HFNetB200Engine &operator=(const HFNetB200Engine &) = delete;
};

class HFNetB200Model : public BaseModel
{
public:
    /* A level facade over a shared engine: level 0 is kImageToLocalAndGlobal, the others kImageToLocal
     * (BaseModel.cc:52-58). */
    HFNetB200Model(std::shared_ptr<HFNetB200Engine> engine, int level, ModelDetectionMode mode)
        : mMode(mode), mLevel(level), mEngine(std::move(engine)) {}

    /* Stand-alone single-level model, like InitRTModel(path, mode, inputShape = {1, H, W, 1}) (BaseModel.cc:117-142). */
    HFNetB200Model(const std::vector<unsigned char> &weights, ModelDetectionMode mode, cv::Vec4i inputShape, int device = 0)
        : mMode(mode), mLevel(0),
          mEngine(std::make_shared<HFNetB200Engine>(weights, inputShape(1), inputShape(2), 1, 1.2f, 8192, device)) {}

    bool Detect(const cv::Mat &image, std::vector<cv::KeyPoint> &vKeyPoints, cv::Mat &localDescriptors,
                cv::Mat &globalDescriptors, int nKeypointsNum, float threshold) override
    {
        if (mMode != kImageToLocalAndGlobal) return false;                 /* HFNetRTModel.cc:87 */
        return Run(image, vKeyPoints, localDescriptors, &globalDescriptors, nKeypointsNum, threshold);
    }

    bool Detect(const cv::Mat &image, std::vector<cv::KeyPoint> &vKeyPoints, cv::Mat &localDescriptors,
                int nKeypointsNum, float threshold) override
    {
        if (mMode == kIntermediateToGlobal) return false;                   /* HFNetRTModel.cc:103 */
        return Run(image, vKeyPoints, localDescriptors, nullptr, nKeypointsNum, threshold);
    }

    /* The split local / global execution only exists for the TensorFlow back-end (HFNetTFModelV2.cc:41-46); the
     * TensorRT back-end returns false here as well (HFNetRTModel.cc:112-120). */
    bool Detect(const cv::Mat &, cv::Mat &) override { return false; }

    bool IsValid(void) override { return mEngine && mEngine->valid; }

    /* The reference's enum has no slot for this back-end until the maintainer adds kHFNetB200Model;
     * kHFNetRTModel keeps every existing comparison (HFextractor.cc:151 only singles out VINO) valid meanwhile. */
    ModelType Type(void) override { return kHFNetRTModel; }

    const std::string &LastError() const { return mEngine->error; }
    std::shared_ptr<HFNetB200Engine> Engine() const { return mEngine; }

    /* HFextractor::operator() (HFextractor.cc:142-157) for nlevels > 1 as ONE call: pyramid (ComputePyramid :159-173),
     * every level's network + selection, octave / scale assignment and concatenation (:255-284) run inside the
     * library as one CUDA graph with one H2D of the frame and one D2H of the results.  vnFeaturesPerLevel is
     * HFextractor::mnFeaturesPerLevel.  Returns the number of keypoints, -1 on a bad image like :145. */
    int ExtractPyramid(const cv::Mat &image, const std::vector<int> &vnFeaturesPerLevel, float threshold,
                       std::vector<cv::KeyPoint> &vKeyPoints, cv::Mat &localDescriptors, cv::Mat &globalDescriptors,
                       bool bKeepDescriptorsOnDevice = false)   /* true: localDescriptors comes back empty, the rows stay in HBM */
    {
        if (image.empty() || image.type() != CV_8UC1) return -1;
        HFNetB200Engine &e = *mEngine;
        if (!e.valid || image.rows != e.height || image.cols != e.width || (int)vnFeaturesPerLevel.size() != e.nLevels)
            return -1;
        int32_t budget[HFB_MAX_LEVELS] = {0, 0, 0, 0, 0, 0, 0, 0};
        int total = 0;
        for (int l = 0; l < e.nLevels; ++l) { budget[l] = vnFeaturesPerLevel[l]; total += vnFeaturesPerLevel[l]; }
        std::vector<float> x(total), y(total), r(total);
        std::vector<int32_t> o(total);
        localDescriptors = bKeepDescriptorsOnDevice ? cv::Mat() : cv::Mat(std::max(total, 1), HFB_DESC_DIM, CV_32F);
        cv::Mat g(HFB_GLOBAL_DIM, 1, CV_32F);
        hfb_features f;
        f.x = x.data(); f.y = y.data(); f.response = r.data(); f.octave = o.data();
        f.descriptors = bKeepDescriptorsOnDevice ? nullptr : localDescriptors.ptr<float>();
        f.global_descriptor = g.ptr<float>();
        int status;
        {
            std::lock_guard<std::mutex> lock(e.mutex);
            status = hfb_extract(e.ctx, image.data, image.rows, image.cols, (int32_t)image.step, budget, threshold, &f);
            if (status != HFB_OK) e.error = hfb_last_error(e.ctx);
        }
        if (status != HFB_OK) return -1;
        Fill(f, x, y, r, o, vKeyPoints);
        if (!bKeepDescriptorsOnDevice) localDescriptors = localDescriptors.rowRange(0, f.n_total);
        globalDescriptors = g;
        return f.n_total;
    }

private:
    static void Fill(const hfb_features &f, const std::vector<float> &x, const std::vector<float> &y,
                     const std::vector<float> &r, const std::vector<int32_t> &o, std::vector<cv::KeyPoint> &vKeyPoints)
    {
        vKeyPoints.clear();
        vKeyPoints.reserve(f.n_total);
        cv::KeyPoint kp;
        kp.angle = 0;
        for (int i = 0; i < f.n_total; ++i)
        {
            kp.pt.x = x[i];
            kp.pt.y = y[i];
            kp.response = r[i];
            kp.octave = o[i];
            vKeyPoints.emplace_back(kp);
        }
    }

    bool Run(const cv::Mat &image, std::vector<cv::KeyPoint> &vKeyPoints, cv::Mat &localDescriptors,
             cv::Mat *globalDescriptors, int nKeypointsNum, float threshold)
    {
        if (!IsValid() || image.empty() || image.type() != CV_8UC1 || nKeypointsNum < 0) return false;
        HFNetB200Engine &e = *mEngine;
        std::vector<float> x(nKeypointsNum), y(nKeypointsNum), r(nKeypointsNum);
        std::vector<int32_t> o(nKeypointsNum);
        localDescriptors = cv::Mat(std::max(nKeypointsNum, 1), HFB_DESC_DIM, CV_32F);   /* continuous, Matcher.cc:843 */
        cv::Mat g(HFB_GLOBAL_DIM, 1, CV_32F);                                           /* HFNetRTModel.cc:201 */
        hfb_features f;
        f.x = x.data(); f.y = y.data(); f.response = r.data(); f.octave = o.data();
        f.descriptors = localDescriptors.ptr<float>();
        f.global_descriptor = globalDescriptors ? g.ptr<float>() : nullptr;
        int status;
        {
            std::lock_guard<std::mutex> lock(e.mutex);
            status = hfb_extract_level(e.ctx, mLevel, image.data, image.rows, image.cols, (int32_t)image.step, nKeypointsNum,
                                       threshold, &f);
            if (status != HFB_OK) e.error = hfb_last_error(e.ctx);
        }
        if (status != HFB_OK) return false;
        Fill(f, x, y, r, o, vKeyPoints);          /* level coordinates, octave 0: HFextractor.cc:272-279 does the rest */
        localDescriptors = localDescriptors.rowRange(0, f.n_total);
        if (globalDescriptors) *globalDescriptors = g;
        return true;
    }

    ModelDetectionMode mMode;
    int mLevel;
    std::shared_ptr<HFNetB200Engine> mEngine;
};

/* InitAllModels(strModelPath, modelType, ImSize, nLevels, scaleFactor) (src/Extractors/BaseModel.cc:24-93) for this
 * back-end: one engine, nLevels facade objects (level 0 local + global, the others local only); the caller stores them
 * in gvpModels and leaves gpGlobalModel null, as for TensorRT (:78-81). */
inline std::vector<BaseModel *> InitB200Models(const std::vector<unsigned char> &weights, int imWidth, int imHeight,
                                              int nLevels, float scaleFactor, int maxKeypointsPerLevel = 8192,
                                              int device = 0)
{
    auto engine = std::make_shared<HFNetB200Engine>(weights, imHeight, imWidth, nLevels, scaleFactor, maxKeypointsPerLevel,
                                                    device);
    std::vector<BaseModel *> models;
    for (int l = 0; l < nLevels; ++l)
        models.push_back(new HFNetB200Model(engine, l, l == 0 ? kImageToLocalAndGlobal : kImageToLocal));
    return models;
}

} // namespace ORB_SLAM3

#endif // HFNETB200MODEL_H
