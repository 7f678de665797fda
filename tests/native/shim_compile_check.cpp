// Compile / link check of the reference-side shims without OpenCV / the reference tree (stand-ins in cv_standin.h).
#include "cv_standin.h"
#include "HFNetB200Model.h"
#include "HFNetB200Backends.h"

int main(int argc, char** argv) {
  std::vector<unsigned char> blob(64, 0);
  ORB_SLAM3::HFNetB200Model model(blob, ORB_SLAM3::kImageToLocalAndGlobal, cv::Vec4i{{1, 64, 64, 1}});
  cv::Mat img(64, 64, CV_8UC1), desc, g;
  std::vector<cv::KeyPoint> kps;
  ORB_SLAM3::BaseModel* base = &model;
  bool ok = base->IsValid() && base->Detect(img, kps, desc, g, 100, 0.01f);
  if (ok) {                                   // never reached without a GPU; keeps the other shims' symbols referenced
    ORB_SLAM3::HFNetB200Matcher matcher(model.Engine()->ctx);
    std::vector<int> m;
    matcher.SearchByBoW(desc, desc, m);
    ORB_SLAM3::HFNetB200KeyFrameDatabase db(model.Engine()->ctx, 16);
    db.add(1, 0, g);
    std::vector<unsigned char> outl;
    double pose[7] = {0, 0, 0, 1, 0, 0, 0};
    const float K[4] = {1, 1, 0, 0};
    ORB_SLAM3::HFNetB200Optimizer::PoseOptimization(model.Engine()->ctx, K, pose, {}, {}, {}, outl);
  }
  return ok ? 0 : (argc > 1 ? 2 : 0);   // without a GPU / real weights the model is invalid: exit 0 unless asked
}
