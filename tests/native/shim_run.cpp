// Runs the reference-side shims on a GPU (tests/test_shim_gpu.py builds this with g++ against libhfnet_b200.so, feeds it
// raw input files and compares every output file byte for byte with the ctypes path):
//   shim_run <dir> <H> <W> <n_levels> <n_features_single> <threshold> <budget_0> ... <budget_{L-1}>
// inputs  (<dir>/in_*.bin):  blob, image, level images 1..L-1 (cv::resize chain), two descriptor sets, a keyframe database,
//                            a pose-optimisation and a local-BA problem
// outputs (<dir>/out_*.bin): what BaseModel::Detect / HFextractor::operator() / the Matcher, KeyFrameDatabase and
//                            Optimizer shims return
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "cv_standin.h"
#include "HFNetB200Model.h"
#include "HFNetB200Backends.h"

using namespace ORB_SLAM3;

template <class T>
static std::vector<T> rd(const std::string& p) {
  std::ifstream f(p, std::ios::binary | std::ios::ate);
  if (!f) { std::fprintf(stderr, "cannot open %s\n", p.c_str()); std::exit(3); }
  const size_t n = (size_t)f.tellg();
  std::vector<T> v(n / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <class T>
static void wr(const std::string& p, const T* d, size_t n) {
  std::ofstream f(p, std::ios::binary);
  f.write(reinterpret_cast<const char*>(d), (std::streamsize)(n * sizeof(T)));
}
static cv::Mat mat_u8(const std::vector<unsigned char>& v, int h, int w) {
  cv::Mat m(h, w, CV_8UC1);
  std::memcpy(m.data, v.data(), (size_t)h * w);
  return m;
}
static cv::Mat mat_f32(const std::vector<float>& v, int cols) {
  cv::Mat m((int)(v.size() / cols), cols, CV_32F);
  std::memcpy(m.data, v.data(), v.size() * 4);
  return m;
}
static void dump_features(const std::string& dir, const std::string& tag, const std::vector<cv::KeyPoint>& kps, const cv::Mat& desc,
                          const cv::Mat* g) {
  std::vector<float> xyr;
  std::vector<int> oct;
  for (const cv::KeyPoint& k : kps) { xyr.push_back(k.pt.x); xyr.push_back(k.pt.y); xyr.push_back(k.response); oct.push_back(k.octave); }
  wr(dir + "/out_" + tag + "_xyr.bin", xyr.data(), xyr.size());
  wr(dir + "/out_" + tag + "_oct.bin", oct.data(), oct.size());
  wr(dir + "/out_" + tag + "_desc.bin", desc.ptr<float>(), (size_t)desc.rows * desc.cols);
  if (g) wr(dir + "/out_" + tag + "_global.bin", g->ptr<float>(), (size_t)g->rows * g->cols);
}

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  const std::string dir = argv[1];
  const int H = std::atoi(argv[2]), W = std::atoi(argv[3]), L = std::atoi(argv[4]), nSingle = std::atoi(argv[5]);
  const float thr = (float)std::atof(argv[6]);
  std::vector<int> budget;
  for (int l = 0; l < L; ++l) budget.push_back(std::atoi(argv[7 + l]));
  const std::vector<unsigned char> blob = rd<unsigned char>(dir + "/in_blob.bin");
  const cv::Mat image = mat_u8(rd<unsigned char>(dir + "/in_image.bin"), H, W);

  // 1. bare Detect on a single-level model (ExtractSingleLayer, HFextractor.cc:175-182)
  {
    HFNetB200Model model(blob, kImageToLocalAndGlobal, cv::Vec4i{{1, H, W, 1}});
    BaseModel* base = &model;
    if (!base->IsValid()) { std::fprintf(stderr, "model invalid: %s\n", model.LastError().c_str()); return 4; }
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc, g;
    if (!base->Detect(image, kps, desc, g, nSingle, thr)) { std::fprintf(stderr, "Detect failed: %s\n", model.LastError().c_str()); return 5; }
    dump_features(dir, "single", kps, desc, &g);
    cv::Mat none;
    if (base->Detect(image, none)) return 6;                       // intermediate -> global is not offered (like TensorRT)
  }
  // 2. the unmodified multi-level flow: one facade object per level on ONE shared engine, every level fed its own image,
  //    octave / scale / concat done by the caller exactly as HFextractor.cc:255-284
  std::vector<BaseModel*> models = InitB200Models(blob, W, H, L, 1.2f, 1024);
  {
    std::vector<cv::KeyPoint> all;
    std::vector<float> desc_all;
    cv::Mat g;
    float scale = 1.f;
    int h = H, w = W;
    for (int l = 0; l < L; ++l) {
      if (l > 0) scale *= 1.2f;
      cv::Mat lvl = image;
      if (l > 0) {
        const std::vector<unsigned char> v = rd<unsigned char>(dir + "/in_level" + std::to_string(l) + ".bin");
        const std::vector<int> hw = rd<int>(dir + "/in_level" + std::to_string(l) + "_hw.bin");
        h = hw[0]; w = hw[1];
        lvl = mat_u8(v, h, w);
      }
      std::vector<cv::KeyPoint> kps;
      cv::Mat desc, gl;
      const bool ok = l == 0 ? models[l]->Detect(lvl, kps, desc, gl, budget[l], thr) : models[l]->Detect(lvl, kps, desc, budget[l], thr);
      if (!ok) { std::fprintf(stderr, "level %d Detect failed: %s\n", l, static_cast<HFNetB200Model*>(models[l])->LastError().c_str()); return 7; }
      if (l == 0) g = gl;
      if (l > 0 && models[l]->Detect(lvl, kps, desc, gl, budget[l], thr)) return 8;   // wrong mode -> false (HFNetRTModel.cc:87)
      for (cv::KeyPoint& k : kps) { k.octave = l; k.pt.x *= scale; k.pt.y *= scale; all.push_back(k); }
      desc_all.insert(desc_all.end(), desc.ptr<float>(), desc.ptr<float>() + (size_t)desc.rows * 256);
    }
    cv::Mat d = mat_f32(desc_all, 256);
    dump_features(dir, "levels", all, d, &g);
  }
  // 3. HFextractor::operator() as one fused call
  {
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc, g;
    const int n = static_cast<HFNetB200Model*>(models[0])->ExtractPyramid(image, budget, thr, kps, desc, g);
    if (n < 0) return 9;
    dump_features(dir, "pyramid", kps, desc, &g);
    // the same call with the local descriptors left on the device: same keypoints and global descriptor, no descriptor rows
    std::vector<cv::KeyPoint> kps2;
    cv::Mat desc2, g2;
    const int n2 = static_cast<HFNetB200Model*>(models[0])->ExtractPyramid(image, budget, thr, kps2, desc2, g2, true);
    if (n2 != n || !desc2.empty() || std::memcmp(g2.data, g.data, HFB_GLOBAL_DIM * 4) != 0) return 15;
    for (int i = 0; i < n; ++i)
      if (kps2[i].pt.x != kps[i].pt.x || kps2[i].pt.y != kps[i].pt.y || kps2[i].octave != kps[i].octave) return 16;
  }
  hfb_ctx* ctx = static_cast<HFNetB200Model*>(models[0])->Engine()->ctx;
  // 4. Matcher
  {
    const cv::Mat A = mat_f32(rd<float>(dir + "/in_descA.bin"), 256), B = mat_f32(rd<float>(dir + "/in_descB.bin"), 256);
    HFNetB200Matcher matcher(ctx);
    std::vector<int> m;
    std::vector<float> dist;
    matcher.SearchByBoW(A, B, m, &dist);
    wr(dir + "/out_bow_idx.bin", m.data(), m.size());
    wr(dir + "/out_bow_dist.bin", dist.data(), dist.size());
    std::vector<std::pair<size_t, size_t> > pairs;
    matcher.SearchForTriangulation(A, B, pairs);
    std::vector<int> flat;
    for (auto& p : pairs) { flat.push_back((int)p.first); flat.push_back((int)p.second); }
    wr(dir + "/out_tri_pairs.bin", flat.data(), flat.size());
  }
  // 5. KeyFrameDatabase (ids 100 + i, map = i % 3; map 1 is cleared again)
  {
    const std::vector<float> rows = rd<float>(dir + "/in_kfdb.bin");
    const int n = (int)(rows.size() / 4096) - 1;                   // last row = query
    HFNetB200KeyFrameDatabase db(ctx, n + 4);
    for (int i = 0; i < n; ++i) db.add(100 + i, i % 3, mat_f32(std::vector<float>(rows.begin() + (size_t)i * 4096, rows.begin() + (size_t)(i + 1) * 4096), 1));
    db.clearMap(1);
    db.erase(100);
    std::vector<long unsigned int> ids;
    std::vector<float> sc;
    float best = 0;
    db.Query(mat_f32(std::vector<float>(rows.end() - 4096, rows.end()), 1), 0.8f, 0.f, ids, sc, best);
    std::vector<long long> ids64(ids.begin(), ids.end());
    ids64.push_back(db.size());
    wr(dir + "/out_kfdb_ids.bin", ids64.data(), ids64.size());
    sc.push_back(best);
    wr(dir + "/out_kfdb_scores.bin", sc.data(), sc.size());
  }
  // 6. Optimizer
  {
    std::vector<double> pose = rd<double>(dir + "/in_pose0.bin");
    const std::vector<double> Xw = rd<double>(dir + "/in_pose_Xw.bin"), obs = rd<double>(dir + "/in_pose_obs.bin"),
                              is2 = rd<double>(dir + "/in_pose_is2.bin");
    const std::vector<float> K = rd<float>(dir + "/in_K.bin");
    std::vector<unsigned char> outl;
    const int ninl = HFNetB200Optimizer::PoseOptimization(ctx, K.data(), pose.data(), Xw, obs, is2, outl);
    pose.push_back((double)ninl);
    wr(dir + "/out_pose.bin", pose.data(), pose.size());
    wr(dir + "/out_pose_outlier.bin", outl.data(), outl.size());
    std::vector<double> poses = rd<double>(dir + "/in_lba_poses.bin"), points = rd<double>(dir + "/in_lba_points.bin");
    const std::vector<unsigned char> fixed = rd<unsigned char>(dir + "/in_lba_fixed.bin");
    const std::vector<int> ecam = rd<int>(dir + "/in_lba_cam.bin"), ept = rd<int>(dir + "/in_lba_pt.bin");
    const std::vector<double> lobs = rd<double>(dir + "/in_lba_obs.bin"), lis2 = rd<double>(dir + "/in_lba_is2.bin");
    bool stop = false;
    std::vector<unsigned char> lout;
    if (!HFNetB200Optimizer::LocalBundleAdjustment(ctx, poses, fixed, points, ecam, ept, lobs, lis2, K.data(), &stop, lout, 5)) return 10;
    wr(dir + "/out_lba_poses.bin", poses.data(), poses.size());
    wr(dir + "/out_lba_points.bin", points.data(), points.size());
    wr(dir + "/out_lba_outlier.bin", lout.data(), lout.size());
  }
  // 7. Frame: calibration-dependent steps on the single-level keypoints
  {
    const std::vector<float> cam = rd<float>(dir + "/in_cam.bin");      // fx fy cx cy k1 k2 p1 p2
    cv::Mat dist(4, 1, CV_32F);
    std::memcpy(dist.data, cam.data() + 4, 16);
    if (!HFNetB200Frame::SetCamera(ctx, cam[0], cam[1], cam[2], cam[3], dist)) return 11;
    const std::vector<float> xyr = rd<float>(dir + "/out_single_xyr.bin");   // mvKeys of the single-level Detect above
    std::vector<cv::KeyPoint> kps(xyr.size() / 3), un;
    for (size_t i = 0; i < kps.size(); ++i) { kps[i].pt.x = xyr[3 * i]; kps[i].pt.y = xyr[3 * i + 1]; kps[i].response = xyr[3 * i + 2]; }
    if (!HFNetB200Frame::UndistortKeyPoints(ctx, kps, un)) return 13;
    std::vector<float> o;
    for (size_t i = 0; i < un.size(); ++i) { o.push_back(un[i].pt.x); o.push_back(un[i].pt.y); o.push_back(un[i].response); }
    float b[4];
    if (!HFNetB200Frame::ComputeImageBounds(ctx, W, H, b[0], b[1], b[2], b[3])) return 14;
    o.insert(o.end(), b, b + 4);
    wr(dir + "/out_undistorted.bin", o.data(), o.size());
  }
  for (BaseModel* m : models) delete m;
  std::printf("SHIM_RUN_OK\n");
  return 0;
}
