// Shared declarations of libhfnet_b200.so (sm_100a only).  Host-side context, error plumbing, device buffers.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hfnet_b200.h"

#define HFB_CUDA(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                      \
      return HFB_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

#define HFB_CHECK_LAUNCH(ctx, what)                                                              \
  do {                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) {                                                                     \
      (ctx)->set_error(std::string(what) + ": launch failed: " + cudaGetErrorString(_e));        \
      return HFB_ERR_CUDA;                                                                       \
    }                                                                                            \
    (ctx)->launches++;                                                                           \
    if ((ctx)->prof_on) (ctx)->prof_mark(what);                                                  \
  } while (0)

#define HFB_REQUIRE(ctx, cond, msg)                                                              \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      (ctx)->set_error(msg);                                                                     \
      return HFB_ERR_INVALID;                                                                    \
    }                                                                                            \
  } while (0)

#define HFB_TRY(expr)                                                                            \
  do {                                                                                           \
    int _s = (expr);                                                                             \
    if (_s != HFB_OK) return _s;                                                                 \
  } while (0)

// Every extern "C" entry point runs on its context's device whatever the calling thread's current device is (the
// reference calls Detect from cv::parallel_for_ workers and three SLAM threads, src/Extractors/HFextractor.cc:228-243):
// kernel launches, cudaMalloc and cudaFuncSetAttribute all act on the thread's current device.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int device) : dev(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define HFB_ENTER(ctx)                      \
  if (!(ctx)) return HFB_ERR_INVALID;       \
  DeviceGuard _device_guard((ctx)->device)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t ceil_div_sz(size_t a, size_t b) { return (a + b - 1) / b; }

typedef unsigned long long u64;

// A GEMM-shaped layer: weights fp16 [N][Kp] (K-major rows, Kp = K rounded up to 8), bias fp32 [N].
struct GemmW {
  int K = 0, Kp = 0, N = 0;
  const __half* w = nullptr;
  const float* b = nullptr;
};

// One inverted-residual block (hfnet/models/backbones/utils/conv_blocks.py:162-312), BatchNorm folded.
struct BlockW {
  int layer = 0, cin = 0, cexp = 0, cout = 0, stride = 1;
  bool has_expand = false, residual = false;
  GemmW expand, project;
  const float *wd = nullptr, *bd = nullptr;  // depthwise [9][cexp], [cexp]
};

struct NetW {
  int c1 = 0, n_clusters = 0, c_local = 0, c_global = 0;
  const float *conv1_w = nullptr, *conv1_b = nullptr;  // [9][c1], [c1]
  std::vector<BlockW> blocks;
  GemmW head1;      // desc.conv1 (N rows 0..255) and det.conv1 (rows 256..383) concatenated: K = 9*c_local
  GemmW desc2;      // 256 -> 256
  GemmW det2;       // 128 -> 65
  const float *vlad_w = nullptr, *vlad_b = nullptr, *vlad_c = nullptr;  // [c_global][C], [C], [C][c_global]
  const __half* vlad_wt = nullptr;   // [2][C][c_global + 8]: vlad_w transposed, fp16 high part | fp16 residual (global_head.cu)
  const __half* fc_w = nullptr;   // [c_global*C][4096] fp16 in mma B-fragment order (fc_pack_host, global_head.cu)
  const float* fc_b = nullptr;
};

// Per-pyramid-level plan: shapes and device buffers.
struct LevelPlan {
  int H = 0, W = 0;    // image size of this level
  int H8 = 0, W8 = 0;  // cropped to multiples of 8 (hf_net.py:188-190)
  float scale = 1.f;   // mvScaleFactor[level]
  bool global = false;
  uint8_t* d_img = nullptr;  // [H][W]
  int n_act = 0;             // layers computed for this level (7 or 18)
  __half* act[19] = {nullptr};
  int aH[19] = {0}, aW[19] = {0}, aC[19] = {0};
  __half *d_exp = nullptr, *d_dw = nullptr;  // expand / depthwise scratch
  __half* d_head1 = nullptr;                 // [Hd*Wd][384] desc.conv1 | det.conv1 (ReLU6)
  float* d_logits = nullptr;                 // [Hd*Wd][65]
  float* d_descmap = nullptr;                // [Hd][Wd][256] unit rows
  float *d_scores = nullptr, *d_nms = nullptr;  // [H8][W8]
  u64* d_cand = nullptr;      // [max_batch][cand_cap] threshold-scan survivors
  int* d_cand_count = nullptr;
  // global head (global_head.cu): per-group NetVLAD partial sums, normalised VLAD (fp32 + the FC's packed fp16 hi/lo
  // operand), per-tile sums of squares of the FC output, self-resetting completion counters [max_batch + 1]
  float *d_vlad_part = nullptr, *d_vladn = nullptr, *d_fc_ss = nullptr;
  uint32_t* d_fc_a = nullptr;
  int* d_gh_counters = nullptr;
  int *d_xi = nullptr, *d_yi = nullptr;      // resize tables (from previous level)
  short *d_xa = nullptr, *d_ya = nullptr;
};

// Opt-in to more than 48 KB of dynamic shared memory, per kernel instantiation (one static object each) and per device
// (cudaFuncSetAttribute acts on the current device's copy of the function).  Serialised: contexts live on different host
// threads, and a second thread must not launch before the attribute is in place.
struct SmemOptIn {
  std::mutex mu;
  size_t have[64] = {};
  template <class Kernel>
  cudaError_t ensure(Kernel kernel, int device, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    size_t& h = have[device & 63];
    if (bytes <= h) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) h = bytes;
    return e;
  }
};

struct hfb_ctx {
  hfb_config cfg{};
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;   // global branch of the encoder (forked off after layer_7)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t copy_stream = nullptr;   // in-graph D2H of the local features, concurrent with the matching kernels
  cudaEvent_t ev_local = nullptr, ev_copied = nullptr;
  cudaEvent_t ev_carry_fork = nullptr, ev_carry = nullptr;   // the carry of the previous call's descriptors runs beside the encoder
  bool fork_branches = true;            // HFB_FORK=0: everything on one stream
  bool carry_side = false;              // HFB_CARRY_SIDE=1: the carry of the previous call's descriptors on the copy stream beside the encoder (ctx.cu)
  // pyramid levels >= 1 run on their own streams (fork after the resize chain, join before the sampling kernels): a
  // single frame's level grids are far smaller than the machine, so the levels fill each other's idle SMs
  cudaStream_t level_stream[HFB_MAX_LEVELS] = {nullptr};
  cudaEvent_t ev_level_fork = nullptr, ev_level_join[HFB_MAX_LEVELS] = {nullptr};
  bool fork_levels = true;              // HFB_FORK_LEVELS=0: levels one after the other on the main stream
  bool fused_stem = true;               // HFB_STEM=0: layer_1 and layer_2 as two kernels
  bool join_pending = false;            // the global branch is still running on side_stream (joined by enqueue_extract)
  // Host destinations of the extraction results when they can be written by the extraction graph itself (page-locked,
  // contiguous over the batch): the local features leave on the main stream while the global branch is still computing.
  struct D2HPlan {
    bool on = false;
    float *x = nullptr, *y = nullptr, *r = nullptr, *d = nullptr, *g = nullptr;
    int *o = nullptr, *counts = nullptr, *overflow = nullptr;
    int* match_idx = nullptr;      // with fused matching: [B][kp_cap] rows
    float* match_val = nullptr;
  } d2h;
  // Frame-to-previous-frame association enqueued inside the extraction (hfb_extract_match_batch*): it only needs the
  // local features, so it runs on the main stream while the global branch is still computing.
  struct FusedMatch {
    bool on = false;
    int mode = 0;
    float thr = 0.f;
  } fmatch;
  std::string err;
  uint64_t launches = 0;
  // weights
  void* d_wblob = nullptr;
  bool weights_loaded = false;
  NetW net;
  // levels
  int n_levels = 0;
  LevelPlan lv[HFB_MAX_LEVELS];
  int cand_cap = 0;
  u64* d_sel = nullptr;     // [n_levels][max_batch][8192] sorted top-k keys per level
  int* d_nsel = nullptr;    // [n_levels][max_batch]
  int* d_overflow = nullptr;
  bool debug = false;       // HFB_DEBUG=1: keep detector logits for hfb_debug_tensor
  // extraction outputs (device): [max_batch][n_levels*max_keypoints] SoA + global descriptors
  int kp_cap = 0;  // per frame = n_levels * max_keypoints
  float *d_kx = nullptr, *d_ky = nullptr, *d_kresp = nullptr, *d_kdesc = nullptr, *d_global = nullptr;
  int* d_koct = nullptr;
  int* d_kcount = nullptr;  // [max_batch][HFB_MAX_LEVELS]
  // Frame's calibration (hfb_set_camera): with distortion every extraction also writes the undistorted coordinates
  // (Frame::mvKeysUn, src/Frame.cc:760-793) of its keypoints; coefficients widened to double as OpenCV does
  struct Camera {
    bool on = false;            // dist[0] != 0
    double fx = 1, fy = 1, cx = 0, cy = 0;
    double k[12] = {0};
  } cam;
  float *d_kxu = nullptr, *d_kyu = nullptr;   // [max_batch][kp_cap]
  bool kun_valid = false;      // the last extraction filled them (hfb_extract_level works in level coordinates and does not)
  int last_budget[HFB_MAX_LEVELS] = {0};
  int last_batch = 0;
  float last_threshold = 0.f;
  // pinned staging
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  double* h_post = nullptr;    // 8 page-locked doubles a kernel posts its scalars into (mapped memory) + a sequence number:
  uint64_t post_seq = 0;       // the LM loop of local BA polls it instead of copying and synchronising per trial
  // generic device scratch (matcher / lba), grown on demand
  void* d_scratch = nullptr;
  size_t d_scratch_bytes = 0;
  void* d_io = nullptr;        // persistent device staging of the host-pointer matcher entry points
  size_t d_io_bytes = 0;
  bool pdl = true;             // HFB_PDL=0: plain stream-ordered launches (no programmatic dependent launch)
  int cpl_min_tiles = 1;       // HFB_CPL_MIN_TILES
  bool fused_blocks = true;    // HFB_FUSED=0: every inverted-residual block on the generic path (expand GEMM, depthwise, project GEMM): parity cross-check of the fused kernel
  bool trace = false;          // HFB_TRACE=1: host-side stage timings of the host-pointer calls on stderr
  std::vector<void*> allocs;
  bool use_graph = true;
  // captured extraction graphs keyed by (batch, budgets, threshold bits)
  struct GraphEntry { std::vector<int> key; cudaGraphExec_t exec; uint64_t kernels; };
  std::vector<GraphEntry> graphs;
  int* d_pair_tab = nullptr;  // single-pair table for the unbatched matcher entry points
  // consecutive-frame matching of the last extraction (hfb_match_consecutive_dev)
  int* d_cm_tab = nullptr;     // [4][max_batch]
  int* d_cm_idx = nullptr;     // [max_batch][kp_cap]
  float* d_cm_val = nullptr;
  void* d_cm_ws = nullptr;     // fixed matcher workspace of the association (its addresses are baked into the graphs)
  size_t d_cm_ws_bytes = 0;
  // Streaming state of the frame-to-previous-frame association (src/Tracking.cc:2030,2167: every frame is matched
  // against the previous frame of its stream, also across calls).  d_kdesc holds 2 * max_batch frame slots: slots
  // [0, max_batch) = the frames of the current call, slots [max_batch, 2 * max_batch) = carried descriptors of the
  // previous call (stream_mode 0: slot max_batch = its last frame; stream_mode 1: slot max_batch + b = its frame b).
  int cm_shift = 1;            // carried slots below frame 0 in the last association's row window
  int stream_mode = 0;         // 0: a batch is B consecutive frames of ONE stream; 1: one frame of each of B streams
  int* d_stream_state = nullptr;  // [0] = frames of the previous extraction (0 = none yet), [1 + s] = rows of carry slot s
  // per-launch profiling (hfb_profile_extract): one CUDA event after every launch, with the launch's algorithmic
  // bytes / flops as stated by the launcher
  bool prof_on = false;
  struct ProfRec { std::string name; cudaEvent_t ev; double bytes, flops; };
  std::vector<ProfRec> prof;
  std::string label;           // set by the launcher before HFB_CHECK_LAUNCH
  double next_bytes = 0, next_flops = 0;
  void note(const std::string& l, double bytes, double flops) { label = l; next_bytes = bytes; next_flops = flops; }
  void prof_mark(const char* what) {
    ProfRec r;
    r.name = label.empty() ? std::string(what) : label;
    r.bytes = next_bytes;
    r.flops = next_flops;
    cudaEventCreate(&r.ev);
    cudaEventRecord(r.ev, stream);
    prof.push_back(r);
    label.clear();
    next_bytes = next_flops = 0;
  }

  void set_error(const std::string& s) { err = s; }
  int ensure_scratch(size_t bytes);
  int ensure_io(size_t bytes);
  int ensure_stage(size_t bytes);
  template <typename T>
  int dalloc(T** p, size_t n) {
    void* q = nullptr;
    size_t bytes = n * sizeof(T);
    cudaError_t e = cudaMalloc(&q, bytes > 0 ? bytes : 16);
    if (e != cudaSuccess) {
      set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
      return HFB_ERR_CUDA;
    }
    allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return HFB_OK;
  }
};

// Kernel launch with the programmatic-dependent-launch attribute (see tc.cuh).  Every kernel launched through this
// helper calls pdl_wait() before touching data of earlier kernels.
template <typename K, typename... Args>
static inline void hfb_launch(hfb_ctx* ctx, K kernel, dim3 grid, dim3 block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = (ctx->pdl && !ctx->prof_on) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---- launchers implemented in the individual .cu files ------------------------------------------------------
// postproc.cu
int launch_nms(hfb_ctx* ctx, const float* d_scores, float* d_out, int H, int W, int B, float threshold, u64* d_cand,
               int* d_cand_count, int cand_cap);
// d_sel / d_nsel: the LEVEL's slice of the selection arrays; d_nsel_levels + l * nsel_stride = counts of level l (the
// rows of level `level` start after the selected rows of the levels below it)
int launch_select(hfb_ctx* ctx, const float* d_nms, int H, int W, u64* d_cand, int* d_cand_count, int cand_cap,
                  u64* d_sel, int* d_nsel, int n_keypoints, float threshold, int B, int* d_overflow,
                  bool candidates_ready);
int launch_sample(hfb_ctx* ctx, int H, int W, const float* d_descmap, int Hd, int Wd, const u64* d_sel,
                  const int* d_nsel_levels, int nsel_stride, int n_keypoints, float level_scale, int level, int B,
                  int kp_cap, float* d_x, float* d_y, float* d_resp, int* d_oct, float* d_desc, int* d_kcount);
int launch_select_sample(hfb_ctx* ctx, const float* d_nms, int H, int W, const float* d_descmap, int Hd, int Wd,
                         u64* d_cand, int* d_cand_count, int cand_cap, u64* d_sel, int* d_nsel, int n_keypoints,
                         float threshold, float level_scale, int B, int kp_cap, float* d_x, float* d_y,
                         float* d_resp, int* d_oct, float* d_desc, int* d_kcount, int* d_overflow,
                         bool candidates_ready);
// cv::undistortPoints with P = K on n points, or (d_kcount != nullptr) on the first sum(d_kcount[b][:]) keypoints of each of
// B frames laid out [B][kp_cap]
struct UndistortParams { double fx, fy, cx, cy, ifx, ify, k[12]; };
int launch_undistort(hfb_ctx* ctx, const float* d_x, const float* d_y, float* d_xu, float* d_yu, int n, int B, int kp_cap,
                     const int* d_kcount);
int launch_resize(hfb_ctx* ctx, const uint8_t* d_src, int sh, int sw, uint8_t* d_dst, int dh, int dw, const int* d_xi,
                  const short* d_xa, const int* d_yi, const short* d_ya, int B);
void build_resize_tables(int sn, int dn, std::vector<int>& idx, std::vector<short>& coef);
// global_head.cu
void fc_pack_host(const float* fw, int K, int N, __half* out);
int global_head_groups(int P);
int global_head_run(hfb_ctx* ctx, const __half* x, int P, int D, int B, float* d_part, int* d_counters, float* d_vladn,
                    uint32_t* d_apacked, float* d_ss_part, float* d_out);
// encoder.cu
int encoder_plan(hfb_ctx* ctx);
void encoder_forget(hfb_ctx* ctx);
int encoder_forward(hfb_ctx* ctx, int level, int B, float threshold);
// match.cu: pair_tab = device int[4][n_pairs] (a_off | a_cnt | b_off | b_cnt)
int launch_match_batch(hfb_ctx* ctx, int mode, const float* dA, const float* dB, int n_pairs, const int* d_pair_tab,
                       int max_a, int max_b, float thr, int* d_match_idx, float* d_match_val, int na_total,
                       int nb_total, int** d_n_matches_out, void* ws = nullptr, size_t ws_bytes = 0, int pad_rows = 0,
                       const __half* B_img = nullptr, const float* hnb_pre = nullptr);
int launch_match_prep(hfb_ctx* ctx, const float* d_rows, int n, __half* d_img, float* d_hn_l2);
size_t match_workspace_bytes(int na_total, int nb_total, int n_pairs, bool same);

int launch_distinctive(hfb_ctx* ctx, const float* d_desc, const int* d_offsets, int n_points, int max_n, int* d_best_idx,
                       float* d_best_med);

// order-preserving float <-> uint32 map (larger float -> larger uint)
__host__ __device__ static inline unsigned int f2ord(float f) {
#ifdef __CUDA_ARCH__
  unsigned int u = __float_as_uint(f);
#else
  unsigned int u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ static inline float ord2f(unsigned int u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
