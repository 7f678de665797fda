/*
 * HFNetB200Model.h -- the reference-side binding of libhfnet_b200.so: a header-only C++ shim that a maintainer of
 * LiuLimingCode/HFNet_SLAM drops into include/Extractors/ next to HFNetRTModel.h.  It implements the reference's own
 * plugin interface
 *
 *     class BaseModel { virtual bool Detect(image, vKeyPoints, localDescriptors, globalDescriptors, n, thr) = 0; ... }
 *                                                                    (include/Extractors/BaseModel.h:38-54)
 *
 * on top of the plain C entry points of hfnet_b200.h, so HFextractor (src/Extractors/HFextractor.cc) and every call
 * site of GetModelVec() keep working unmodified.  See INTEGRATION.md for the three-line change in BaseModel.cc.
 *
 * It only needs <opencv2/core.hpp> and the reference's BaseModel.h.  To compile-check it in an image without OpenCV,
 * define HFNET_B200_SHIM_STANDALONE and provide cv::Mat / cv::KeyPoint / BaseModel stand-ins (tests/native/).
 */
#ifndef HFNETB200MODEL_H
#define HFNETB200MODEL_H

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

/* One hfb_ctx shared by the per-level facade objects of one extractor (the reference keeps one TensorRT engine per
 * level, src/Extractors/BaseModel.cc:35-65).  Detect on different level objects may be called concurrently from
 * cv::parallel_for_ workers (HFextractor.cc:228-243): calls into one context are serialised here. */
struct HFNetB200Shared
{
    hfb_ctx* ctx = nullptr;
    std::mutex mutex;
    std::string error;
    ~HFNetB200Shared() { if (ctx) hfb_destroy(ctx); }
};

class HFNetB200Model : public BaseModel
{
public:
    /* inputShape = {1, H, W, 1} like the other back-ends (BaseModel.cc:35-65); weights = flat HFB2WTS1 blob. */
    HFNetB200Model(const std::vector<unsigned char> &weights, ModelDetectionMode mode, cv::Vec4i inputShape, int device = 0)
        : mMode(mode), mShape(inputShape), mShared(std::make_shared<HFNetB200Shared>())
    {
        hfb_config cfg;
        cfg.device = device;
        cfg.height = inputShape(1);
        cfg.width = inputShape(2);
        cfg.n_levels = 1;
        cfg.scale_factor = 1.2f;
        cfg.max_keypoints = 8192;
        cfg.max_batch = 1;
        cfg.with_global = (mode == kImageToLocalAndGlobal) ? 1 : 0;
        mbValid = hfb_create(&cfg, &mShared->ctx) == HFB_OK &&
                  hfb_load_weights(mShared->ctx, weights.data(), weights.size()) == HFB_OK;
        if (!mbValid) mShared->error = hfb_last_error(mShared->ctx);
    }

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

    bool IsValid(void) override { return mbValid; }

    ModelType Type(void) override { return kHFNetRTModel; }

    const std::string &LastError() const { return mShared->error; }

private:
    bool Run(const cv::Mat &image, std::vector<cv::KeyPoint> &vKeyPoints, cv::Mat &localDescriptors,
             cv::Mat *globalDescriptors, int nKeypointsNum, float threshold)
    {
        if (!mbValid || image.empty() || image.type() != CV_8UC1 || image.rows != mShape(1) || image.cols != mShape(2) ||
            nKeypointsNum < 0 || nKeypointsNum > 8192)
            return false;
        std::vector<float> x(nKeypointsNum), y(nKeypointsNum), r(nKeypointsNum);
        std::vector<int32_t> o(nKeypointsNum);
        localDescriptors = cv::Mat(std::max(nKeypointsNum, 1), HFB_DESC_DIM, CV_32F);   /* continuous, Matcher.cc:843 */
        cv::Mat g(HFB_GLOBAL_DIM, 1, CV_32F);                                           /* HFNetRTModel.cc:201 */
        hfb_features f;
        f.x = x.data(); f.y = y.data(); f.response = r.data(); f.octave = o.data();
        f.descriptors = localDescriptors.ptr<float>();
        f.global_descriptor = globalDescriptors ? g.ptr<float>() : nullptr;
        const int32_t budget[HFB_MAX_LEVELS] = {nKeypointsNum, 0, 0, 0, 0, 0, 0, 0};
        int status;
        {
            std::lock_guard<std::mutex> lock(mShared->mutex);
            status = hfb_extract(mShared->ctx, image.data, image.rows, image.cols, (int32_t)image.step, budget, threshold, &f);
            if (status != HFB_OK) mShared->error = hfb_last_error(mShared->ctx);
        }
        if (status != HFB_OK) return false;
        vKeyPoints.clear();
        vKeyPoints.reserve(f.n_total);
        cv::KeyPoint kp;
        kp.angle = 0;
        kp.octave = 0;
        for (int i = 0; i < f.n_total; ++i)
        {
            kp.pt.x = x[i];
            kp.pt.y = y[i];
            kp.response = r[i];
            vKeyPoints.emplace_back(kp);
        }
        localDescriptors = localDescriptors.rowRange(0, f.n_total);
        if (globalDescriptors) *globalDescriptors = g;
        return true;
    }

    ModelDetectionMode mMode;
    cv::Vec4i mShape;
    std::shared_ptr<HFNetB200Shared> mShared;
    bool mbValid = false;
};

} // namespace ORB_SLAM3

#endif // HFNETB200MODEL_H
