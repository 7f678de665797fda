// Minimal stand-ins for the few cv:: types and the BaseModel interface (include/Extractors/BaseModel.h:10-54) that the
// reference-side shims (include/HFNetB200Model.h, include/HFNetB200Backends.h) touch, so that they can be compiled AND RUN
// in an image without OpenCV / the reference tree.  Test infrastructure only.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#define HFNET_B200_SHIM_STANDALONE
#define CV_8UC1 0
#define CV_32F 5
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct Vec4i { int v[4]; int operator()(int i) const { return v[i]; } };
struct Mat {
  int rows = 0, cols = 0, type_ = 0; size_t step = 0; unsigned char* data = nullptr; std::vector<unsigned char> buf;
  Mat() {}
  Mat(int r, int c, int t) : rows(r), cols(c), type_(t), step((size_t)c * (t == CV_32F ? 4 : 1)), buf((size_t)r * c * (t == CV_32F ? 4 : 1)) { data = buf.data(); }
  Mat(const Mat& o) : rows(o.rows), cols(o.cols), type_(o.type_), step(o.step), buf(o.buf) { data = buf.data(); }
  Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; type_ = o.type_; step = o.step; buf = o.buf; data = buf.data(); return *this; }
  bool empty() const { return rows == 0 || cols == 0; }
  int type() const { return type_; }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + r * step); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + r * step); }
  Mat rowRange(int a, int b) const { Mat m(b - a, cols, type_); if (b > a) std::memcpy(m.data, data + a * step, (size_t)(b - a) * step); return m; }
};
}  // namespace cv
namespace ORB_SLAM3 {
enum ModelType { kHFNetTFModel, kHFNetRTModel, kHFNetVINOModel };
enum ModelDetectionMode { kImageToLocalAndGlobal, kImageToLocal, kImageToLocalAndIntermediate, kIntermediateToGlobal };
class BaseModel {
 public:
  virtual ~BaseModel(void) = default;
  virtual bool Detect(const cv::Mat&, std::vector<cv::KeyPoint>&, cv::Mat&, cv::Mat&, int, float) = 0;
  virtual bool Detect(const cv::Mat&, std::vector<cv::KeyPoint>&, cv::Mat&, int, float) = 0;
  virtual bool Detect(const cv::Mat&, cv::Mat&) = 0;
  virtual bool IsValid(void) = 0;
  virtual ModelType Type(void) = 0;
};
}  // namespace ORB_SLAM3
