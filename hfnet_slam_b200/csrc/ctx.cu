// libhfnet_b200.so: context, weight loading and the extraction / matching entry points of include/hfnet_b200.h.
#include <math.h>

#include <algorithm>

#include "common.cuh"

// ================================================================================================ context
static void drop_graphs(hfb_ctx* ctx);

int hfb_ctx::ensure_scratch(size_t bytes) {
  if (bytes <= d_scratch_bytes) return HFB_OK;
  // No captured graph refers to the scratch block (the in-graph association owns a fixed workspace, d_cm_ws); dropping
  // the graphs before the block moves keeps that true by construction should a later graph ever use it.
  drop_graphs(this);
  if (d_scratch) cudaFree(d_scratch);
  d_scratch = nullptr;
  d_scratch_bytes = 0;
  size_t want = bytes + bytes / 4;
  cudaError_t e = cudaMalloc(&d_scratch, want);
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(scratch): ") + cudaGetErrorString(e));
    return HFB_ERR_CUDA;
  }
  d_scratch_bytes = want;
  return HFB_OK;
}

int hfb_ctx::ensure_io(size_t bytes) {
  if (bytes <= d_io_bytes) return HFB_OK;
  if (d_io) cudaFree(d_io);
  d_io = nullptr;
  d_io_bytes = 0;
  const size_t want = bytes + bytes / 2;
  cudaError_t e = cudaMalloc(&d_io, want);
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(io): ") + cudaGetErrorString(e));
    return HFB_ERR_CUDA;
  }
  d_io_bytes = want;
  return HFB_OK;
}

#include <chrono>
struct StageTimer {
  bool on;
  const char* what;
  std::chrono::steady_clock::time_point t0;
  std::string log;
  StageTimer(bool enabled, const char* w) : on(enabled), what(w), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* stage) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof(buf), " %s=%.3fms", stage, std::chrono::duration<double, std::milli>(t1 - t0).count());
    log += buf;
    t0 = t1;
  }
  ~StageTimer() {
    if (on) fprintf(stderr, "[hfb trace] %s:%s\n", what, log.c_str());
  }
};

int hfb_ctx::ensure_stage(size_t bytes) {
  if (bytes <= h_stage_bytes) return HFB_OK;
  if (h_stage) cudaFreeHost(h_stage);
  h_stage = nullptr;
  h_stage_bytes = 0;
  cudaError_t e = cudaMallocHost(&h_stage, bytes);
  if (e != cudaSuccess) {
    set_error(std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    return HFB_ERR_CUDA;
  }
  h_stage_bytes = bytes;
  return HFB_OK;
}

static int cv_round(double v) { return (int)lrint(v); }  // cvRound: round half to even

extern "C" int hfb_version(void) { return HFB_VERSION; }

extern "C" const char* hfb_last_error(const hfb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int hfb_create(const hfb_config* cfg, hfb_ctx** out) {
  if (!cfg || !out) return HFB_ERR_INVALID;
  *out = nullptr;
  if (cfg->height < 16 || cfg->width < 16 || cfg->n_levels < 1 || cfg->n_levels > HFB_MAX_LEVELS ||
      cfg->max_keypoints < 1 || cfg->max_keypoints > 8192 || cfg->max_batch < 1 || cfg->max_batch > 64 ||
      !(cfg->scale_factor > 1.0f || cfg->n_levels == 1))
    return HFB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) return HFB_ERR_CUDA;
  DeviceGuard _device_guard(cfg->device);   // the caller's current device is restored on return
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return HFB_ERR_CUDA;
  hfb_ctx* ctx = new hfb_ctx();
  ctx->cfg = *cfg;
  ctx->device = cfg->device;
  ctx->n_sm = prop.multiProcessorCount;
  ctx->n_levels = cfg->n_levels;
  const char* dbg = getenv("HFB_DEBUG");
  ctx->debug = dbg && dbg[0] == '1';
  const char* fu = getenv("HFB_FUSED");
  ctx->fused_blocks = !(fu && fu[0] == '0');
  if (const char* mt = getenv("HFB_CPL_MIN_TILES")) ctx->cpl_min_tiles = atoi(mt);
  const char* pd = getenv("HFB_PDL");
  ctx->pdl = !(pd && pd[0] == '0');
  const char* tr = getenv("HFB_TRACE");
  ctx->trace = tr && tr[0] == '1';
  const char* ng = getenv("HFB_NO_GRAPH");
  ctx->use_graph = !(ng && ng[0] == '1');
  *out = ctx;  // returned even on failure so that hfb_last_error works; caller destroys
  if (prop.major != 10) {
    ctx->set_error(std::string("this library is built for sm_100a only; device is ") + prop.name + " (sm_" +
                   std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    return HFB_ERR_CUDA;
  }
  HFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  {
    // the side stream carries the small late-layer grids: highest priority, so that their CTAs are placed first and the
    // wide kernels of the local branch fill the SMs they leave free
    int lo = 0, hi = 0;
    HFB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    HFB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, hi));
  }
  HFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_local, cudaEventDisableTiming));
  HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
  HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_carry_fork, cudaEventDisableTiming));
  HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_carry, cudaEventDisableTiming));
  HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  if (const char* fk = getenv("HFB_FORK")) ctx->fork_branches = !(fk[0] == '0');
  if (const char* cs = getenv("HFB_CARRY_SIDE")) ctx->carry_side = !(cs[0] == '0');
  if (const char* fk = getenv("HFB_FORK_LEVELS")) ctx->fork_levels = !(fk[0] == '0');
  if (ctx->n_levels > 1) {
    HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_level_fork, cudaEventDisableTiming));
    for (int l = 1; l < ctx->n_levels; ++l) {
      HFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->level_stream[l], cudaStreamNonBlocking));
      HFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_level_join[l], cudaEventDisableTiming));
    }
  }
  if (const char* sk = getenv("HFB_STEM")) ctx->fused_stem = !(sk[0] == '0');
  // level shapes: mvScaleFactor[l] = scaleFactor^l accumulated in float (HFextractor.cc:92-103); image size of
  // level l = cvRound(size * 1/scale) (HFextractor.cc:159-173, BaseModel.cc:35-65)
  float sf = 1.f;
  for (int l = 0; l < ctx->n_levels; ++l) {
    LevelPlan& lv = ctx->lv[l];
    if (l > 0) sf = sf * cfg->scale_factor;
    lv.scale = sf;
    const float inv = 1.0f / sf;
    lv.H = l == 0 ? cfg->height : cv_round((double)((float)cfg->height * inv));
    lv.W = l == 0 ? cfg->width : cv_round((double)((float)cfg->width * inv));
    lv.H8 = lv.H / 8 * 8;
    lv.W8 = lv.W / 8 * 8;
    lv.global = (l == 0 && cfg->with_global);
    if (lv.H8 < 8 || lv.W8 < 8) {
      ctx->set_error("pyramid level too small");
      return HFB_ERR_INVALID;
    }
    HFB_TRY(ctx->dalloc(&lv.d_img, (size_t)cfg->max_batch * lv.H * lv.W));
    if (l > 0) {
      std::vector<int> xi, yi;
      std::vector<short> xa, ya;
      build_resize_tables(ctx->lv[l - 1].W, lv.W, xi, xa);
      build_resize_tables(ctx->lv[l - 1].H, lv.H, yi, ya);
      HFB_TRY(ctx->dalloc(&lv.d_xi, xi.size()));
      HFB_TRY(ctx->dalloc(&lv.d_xa, xa.size()));
      HFB_TRY(ctx->dalloc(&lv.d_yi, yi.size()));
      HFB_TRY(ctx->dalloc(&lv.d_ya, ya.size()));
      HFB_CUDA(ctx, cudaMemcpy(lv.d_xi, xi.data(), xi.size() * sizeof(int), cudaMemcpyHostToDevice));
      HFB_CUDA(ctx, cudaMemcpy(lv.d_xa, xa.data(), xa.size() * sizeof(short), cudaMemcpyHostToDevice));
      HFB_CUDA(ctx, cudaMemcpy(lv.d_yi, yi.data(), yi.size() * sizeof(int), cudaMemcpyHostToDevice));
      HFB_CUDA(ctx, cudaMemcpy(lv.d_ya, ya.data(), ya.size() * sizeof(short), cudaMemcpyHostToDevice));
    }
  }
  ctx->cand_cap = std::min(65536, std::max(4096, cfg->height * cfg->width / 8));
  ctx->kp_cap = ctx->n_levels * cfg->max_keypoints;
  const size_t nb = (size_t)cfg->max_batch;
  HFB_TRY(ctx->dalloc(&ctx->d_kx, nb * ctx->kp_cap));
  HFB_TRY(ctx->dalloc(&ctx->d_ky, nb * ctx->kp_cap));
  HFB_TRY(ctx->dalloc(&ctx->d_kresp, nb * ctx->kp_cap));
  HFB_TRY(ctx->dalloc(&ctx->d_kxu, nb * ctx->kp_cap));
  HFB_TRY(ctx->dalloc(&ctx->d_kyu, nb * ctx->kp_cap));
  HFB_TRY(ctx->dalloc(&ctx->d_koct, nb * ctx->kp_cap));
  {
    // 2 * max_batch frame slots: the carried descriptors of the previous call sit right below the current frames
    float* slots = nullptr;
    HFB_TRY(ctx->dalloc(&slots, 2 * nb * ctx->kp_cap * HFB_DESC_DIM));
    ctx->d_kdesc = slots + nb * ctx->kp_cap * HFB_DESC_DIM;
  }
  HFB_TRY(ctx->dalloc(&ctx->d_global, nb * HFB_GLOBAL_DIM));
  HFB_TRY(ctx->dalloc(&ctx->d_kcount, nb * HFB_MAX_LEVELS));
  HFB_TRY(ctx->dalloc(&ctx->d_sel, (size_t)ctx->n_levels * nb * 8192));
  HFB_TRY(ctx->dalloc(&ctx->d_nsel, (size_t)ctx->n_levels * nb));
  HFB_CUDA(ctx, cudaMemset(ctx->d_nsel, 0, (size_t)ctx->n_levels * nb * sizeof(int)));
  HFB_TRY(ctx->dalloc(&ctx->d_overflow, 1));
  HFB_TRY(ctx->dalloc(&ctx->d_pair_tab, 4));
  HFB_TRY(ctx->dalloc(&ctx->d_cm_tab, 4 * nb));
  HFB_TRY(ctx->dalloc(&ctx->d_cm_idx, 2 * nb * ctx->kp_cap));
  HFB_TRY(ctx->dalloc(&ctx->d_cm_val, 2 * nb * ctx->kp_cap));
  HFB_TRY(ctx->dalloc(&ctx->d_stream_state, 1 + nb));
  HFB_CUDA(ctx, cudaMemset(ctx->d_stream_state, 0, (1 + nb) * sizeof(int)));
  {
    const int rows = (int)(2 * nb * ctx->kp_cap);
    ctx->d_cm_ws_bytes = match_workspace_bytes(rows, rows, (int)nb, true);
    uint8_t* w = nullptr;
    HFB_TRY(ctx->dalloc(&w, ctx->d_cm_ws_bytes));
    ctx->d_cm_ws = w;
  }
  HFB_CUDA(ctx, cudaMemset(ctx->d_overflow, 0, sizeof(int)));
  HFB_CUDA(ctx, cudaMemset(ctx->d_kcount, 0, nb * HFB_MAX_LEVELS * sizeof(int)));
  return HFB_OK;
}

static void drop_graphs(hfb_ctx* ctx) {
  if (ctx->stream && !ctx->graphs.empty()) cudaStreamSynchronize(ctx->stream);
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  ctx->graphs.clear();
}

extern "C" void hfb_destroy(hfb_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard _device_guard(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  drop_graphs(ctx);
  encoder_forget(ctx);
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->d_wblob) cudaFree(ctx->d_wblob);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_io) cudaFree(ctx->d_io);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->h_post) cudaFreeHost(ctx->h_post);
  if (ctx->ev_local) cudaEventDestroy(ctx->ev_local);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->ev_carry_fork) cudaEventDestroy(ctx->ev_carry_fork);
  if (ctx->ev_carry) cudaEventDestroy(ctx->ev_carry);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_level_fork) cudaEventDestroy(ctx->ev_level_fork);
  for (int l = 0; l < HFB_MAX_LEVELS; ++l) {
    if (ctx->ev_level_join[l]) cudaEventDestroy(ctx->ev_level_join[l]);
    if (ctx->level_stream[l]) cudaStreamDestroy(ctx->level_stream[l]);
  }
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" void* hfb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes > 0 ? bytes : 16) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void hfb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

extern "C" int hfb_sync(hfb_ctx* ctx) {
  HFB_ENTER(ctx);
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}
extern "C" void* hfb_stream(hfb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t hfb_launch_count(const hfb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ================================================================================================ weights
static int make_divisible(double v, int divisor, int min_value) {
  int nv = std::max(min_value, (int)(v + divisor / 2.0) / divisor * divisor);
  if (nv < 0.9 * v) nv += divisor;
  return nv;
}

struct ArenaBuilder {
  std::vector<uint8_t> host;
  size_t add(const void* src, size_t bytes) {
    size_t off = (host.size() + 255) & ~(size_t)255;
    host.resize(off + bytes);
    memcpy(host.data() + off, src, bytes);
    return off;
  }
};

static __half h_f2h(float f) { return __float2half_rn(f); }

// [K][N] fp32 (blob layout) -> fp16 [N][Kp]
static size_t add_gemm_w(ArenaBuilder& ab, const float* w, int K, int N, int Kp) {
  std::vector<__half> t((size_t)N * Kp, h_f2h(0.f));
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) t[(size_t)n * Kp + k] = h_f2h(w[(size_t)k * N + n]);
  return ab.add(t.data(), t.size() * sizeof(__half));
}

extern "C" int hfb_load_weights(hfb_ctx* ctx, const void* blob, size_t nbytes) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, !ctx->weights_loaded, "weights already loaded (create a new context)");
  const uint8_t* p = reinterpret_cast<const uint8_t*>(blob);
  HFB_REQUIRE(ctx, blob && nbytes >= 32 && memcmp(p, "HFB2WTS1", 8) == 0, "bad weight blob magic");
  uint32_t version, n_clusters, n_tensors;
  float dm;
  uint64_t total;
  memcpy(&version, p + 8, 4);
  memcpy(&n_clusters, p + 12, 4);
  memcpy(&dm, p + 16, 4);
  memcpy(&n_tensors, p + 20, 4);
  memcpy(&total, p + 24, 8);
  HFB_REQUIRE(ctx, version == 1, "unsupported weight blob version");
  HFB_REQUIRE(ctx, n_clusters >= 1 && n_clusters <= 64, "n_clusters must be in [1,64]");
  HFB_REQUIRE(ctx, nbytes == 32 + total * 4, "weight blob size mismatch");
  const float* f = reinterpret_cast<const float*>(p + 32);
  size_t off = 0;
  auto take = [&](size_t n) -> const float* {
    const float* r = f + off;
    off += n;
    return off <= total ? r : nullptr;
  };
  // architecture (hfnet_slam_b200/weights.py:architecture == hf_net.py:13-52 with depth multiplier dm)
  static const int nominal[17][2] = {{1, 16}, {2, 24}, {1, 24}, {2, 32}, {1, 64}, {1, 128}, {2, 64}, {1, 64}, {1, 64},
                                     {1, 64}, {1, 96}, {1, 96}, {1, 96}, {2, 160}, {1, 160}, {1, 160}, {1, 320}};
  NetW& net = ctx->net;
  net.c1 = make_divisible(32.0 * dm, 8, 8);
  net.n_clusters = (int)n_clusters;
  HFB_REQUIRE(ctx, net.c1 % 8 == 0 && net.c1 <= 32, "depth multiplier gives an unsupported first-layer width");
  ArenaBuilder ab;
  struct Fix { const void** dst; size_t off; };
  std::vector<Fix> fixes;
  auto fix = [&](const void** dst, size_t o) { fixes.push_back({dst, o}); };
#define TAKE(var, n)                                                    \
  const float* var = take(n);                                           \
  HFB_REQUIRE(ctx, var != nullptr, "weight blob truncated")
  {
    TAKE(w, (size_t)9 * net.c1);
    TAKE(b, (size_t)net.c1);
    fix((const void**)&net.conv1_w, ab.add(w, (size_t)9 * net.c1 * 4));
    fix((const void**)&net.conv1_b, ab.add(b, (size_t)net.c1 * 4));
  }
  net.blocks.clear();
  net.blocks.reserve(17);
  int cin = net.c1;
  for (int i = 0; i < 17; ++i) {
    BlockW bw;
    bw.layer = i + 2;
    bw.stride = nominal[i][0];
    bw.cin = cin;
    bw.cout = make_divisible(nominal[i][1] * (double)dm, 8, 8);
    bw.cexp = bw.layer == 2 ? make_divisible(cin * 1.0, 1, 1) : make_divisible(cin * 6.0, 8, 8);
    bw.has_expand = bw.cexp > bw.cin;
    bw.residual = bw.stride == 1 && bw.cin == bw.cout;
    HFB_REQUIRE(ctx, bw.cin % 8 == 0 && bw.cexp % 8 == 0 && bw.cout % 8 == 0, "channel counts must be multiples of 8");
    net.blocks.push_back(bw);
    cin = bw.cout;
  }
  for (BlockW& bw : net.blocks) {
    if (bw.has_expand) {
      TAKE(w, (size_t)bw.cin * bw.cexp);
      TAKE(b, (size_t)bw.cexp);
      bw.expand.K = bw.cin; bw.expand.Kp = bw.cin; bw.expand.N = bw.cexp;
      fix((const void**)&bw.expand.w, add_gemm_w(ab, w, bw.cin, bw.cexp, bw.cin));
      fix((const void**)&bw.expand.b, ab.add(b, (size_t)bw.cexp * 4));
    }
    {
      TAKE(w, (size_t)9 * bw.cexp);
      TAKE(b, (size_t)bw.cexp);
      // depthwise weights are fp16 operands like every other conv weight (kept as fp32 words holding fp16-exact values:
      // the fused kernel narrows them losslessly for its mixed-precision FMA, the plain kernels use them as they are)
      std::vector<float> w16((size_t)9 * bw.cexp);
      for (size_t i = 0; i < w16.size(); ++i) w16[i] = __half2float(__float2half_rn(w[i]));
      fix((const void**)&bw.wd, ab.add(w16.data(), (size_t)9 * bw.cexp * 4));
      fix((const void**)&bw.bd, ab.add(b, (size_t)bw.cexp * 4));
    }
    {
      TAKE(w, (size_t)bw.cexp * bw.cout);
      TAKE(b, (size_t)bw.cout);
      bw.project.K = bw.cexp; bw.project.Kp = bw.cexp; bw.project.N = bw.cout;
      fix((const void**)&bw.project.w, add_gemm_w(ab, w, bw.cexp, bw.cout, bw.cexp));
      fix((const void**)&bw.project.b, ab.add(b, (size_t)bw.cout * 4));
    }
  }
  net.c_local = net.blocks[7 - 2].cout;
  net.c_global = net.blocks[18 - 2].cout;
  {
    const int Kh = 9 * net.c_local;
    TAKE(wd1, (size_t)Kh * 256);
    TAKE(bd1, 256);
    TAKE(wd2, (size_t)256 * 256);
    TAKE(bd2, 256);
    TAKE(wt1, (size_t)Kh * 128);
    TAKE(bt1, 128);
    TAKE(wt2, (size_t)128 * 65);
    TAKE(bt2, 65);
    // head1 = desc.conv1 (rows 0..255) | det.conv1 (rows 256..383), K-major
    std::vector<__half> t((size_t)384 * Kh);
    for (int k = 0; k < Kh; ++k) {
      for (int n = 0; n < 256; ++n) t[(size_t)n * Kh + k] = h_f2h(wd1[(size_t)k * 256 + n]);
      for (int n = 0; n < 128; ++n) t[(size_t)(256 + n) * Kh + k] = h_f2h(wt1[(size_t)k * 128 + n]);
    }
    std::vector<float> hb(384);
    memcpy(hb.data(), bd1, 256 * 4);
    memcpy(hb.data() + 256, bt1, 128 * 4);
    net.head1.K = net.c_local; net.head1.Kp = Kh; net.head1.N = 384;   // K = channels per tap (implicit 3x3)
    fix((const void**)&net.head1.w, ab.add(t.data(), t.size() * 2));
    fix((const void**)&net.head1.b, ab.add(hb.data(), 384 * 4));
    net.desc2.K = 256; net.desc2.Kp = 256; net.desc2.N = 256;
    fix((const void**)&net.desc2.w, add_gemm_w(ab, wd2, 256, 256, 256));
    fix((const void**)&net.desc2.b, ab.add(bd2, 256 * 4));
    net.det2.K = 128; net.det2.Kp = 128; net.det2.N = 65;
    fix((const void**)&net.det2.w, add_gemm_w(ab, wt2, 128, 65, 128));
    fix((const void**)&net.det2.b, ab.add(bt2, 65 * 4));
  }
  {
    const int D = net.c_global, C = net.n_clusters;
    TAKE(mw, (size_t)D * C);
    TAKE(mb, (size_t)C);
    TAKE(cl, (size_t)C * D);
    TAKE(fw, (size_t)D * C * HFB_GLOBAL_DIM);
    TAKE(fb, (size_t)HFB_GLOBAL_DIM);
    fix((const void**)&net.vlad_w, ab.add(mw, (size_t)D * C * 4));
    {
      const int DP = D + 8;
      std::vector<__half> wt((size_t)2 * C * DP, h_f2h(0.f));
      for (int d = 0; d < D; ++d)
        for (int c = 0; c < C; ++c) {
          const float v = mw[(size_t)d * C + c];
          const __half hi = h_f2h(v);
          wt[(size_t)c * DP + d] = hi;
          wt[(size_t)(C + c) * DP + d] = h_f2h(v - __half2float(hi));
        }
      fix((const void**)&net.vlad_wt, ab.add(wt.data(), wt.size() * 2));
    }
    fix((const void**)&net.vlad_b, ab.add(mb, (size_t)C * 4));
    fix((const void**)&net.vlad_c, ab.add(cl, (size_t)C * D * 4));
    HFB_REQUIRE(ctx, (D * C) % 16 == 0, "clusters x global channels must be a multiple of 16");
    std::vector<__half> t((size_t)D * C * HFB_GLOBAL_DIM);
    fc_pack_host(fw, D * C, HFB_GLOBAL_DIM, t.data());
    fix((const void**)&net.fc_w, ab.add(t.data(), t.size() * 2));
    fix((const void**)&net.fc_b, ab.add(fb, (size_t)HFB_GLOBAL_DIM * 4));
  }
#undef TAKE
  HFB_REQUIRE(ctx, off == total && n_tensors > 0, "weight blob has trailing data");
  HFB_CUDA(ctx, cudaMalloc(&ctx->d_wblob, ab.host.size()));
  HFB_CUDA(ctx, cudaMemcpy(ctx->d_wblob, ab.host.data(), ab.host.size(), cudaMemcpyHostToDevice));
  for (const Fix& fx : fixes) *fx.dst = reinterpret_cast<const uint8_t*>(ctx->d_wblob) + fx.off;
  HFB_TRY(encoder_plan(ctx));
  ctx->weights_loaded = true;
  return HFB_OK;
}

// ================================================================================================ extraction
static int check_overflow(hfb_ctx* ctx) {
  int ov = 0;
  HFB_CUDA(ctx, cudaMemcpyAsync(&ov, ctx->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ov) {
    cudaMemsetAsync(ctx->d_overflow, 0, sizeof(int), ctx->stream);
    ctx->set_error("more threshold-scan candidates than the context's candidate capacity (threshold too low)");
    return HFB_ERR_CAPACITY;
  }
  return HFB_OK;
}

static int enqueue_match_consecutive(hfb_ctx* ctx, int n_images, int mode, float thr);

// Streaming association: before a new extraction overwrites the frame slots, the descriptors the NEXT association needs
// from the previous call are carried into the slots right below d_kdesc (src/Tracking.cc:2030,2167 match every frame
// against mLastFrame, whichever call delivered it).  stream_mode 0 (B consecutive frames of one stream): the previous
// call's last frame -> slot -1.  stream_mode 1 (one frame of each of B streams): previous frame b -> slot b - B (only
// when the previous call had the same batch size).  state[0] = frames of the previous extraction, state[1 + s] = rows
// carried for pair s.
__global__ void carry_prev_kernel(float* __restrict__ kdesc, const int* __restrict__ kcount, int* __restrict__ state,
                                  int n_levels, int kp_cap, int mode, int B) {
  const int s = blockIdx.y;                       // carried slot: 0 (mode 0) or stream b (mode 1)
  const int last_b = state[0];
  int src = -1;
  if (mode == 0) src = last_b - 1;
  else if (last_b == B) src = s;
  int cnt = 0;
  if (src >= 0)
    for (int l = 0; l < n_levels; ++l) cnt += kcount[src * HFB_MAX_LEVELS + l];
  if (cnt > kp_cap) cnt = kp_cap;
  const int shift = mode == 0 ? 1 : B;
  const float4* from = reinterpret_cast<const float4*>(kdesc + (size_t)(src < 0 ? 0 : src) * kp_cap * HFB_DESC_DIM);
  float4* to = reinterpret_cast<float4*>(kdesc + ((long long)s - shift) * kp_cap * HFB_DESC_DIM);
  const int n4 = cnt * (HFB_DESC_DIM / 4);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) to[i] = from[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) state[1 + s] = cnt;
}

// The carry only feeds the association at the end of the call and must be done before the sampling kernels overwrite the
// frame slots it reads.  HFB_CARRY_SIDE=1 runs it on the copy stream beside the encoder (fork here, join before the
// sampling kernels) instead of ahead of the first layer: measured -10 us on the device-resident single-context call
// (0.563 -> 0.552 ms per 8 frames, 0.264 -> 0.258 ms single frame) but +19 us on the host-buffer call, whose in-graph
// D2H of the descriptors shares that branch (end to end 16.1 k -> 14.3 k frames/s) -- so it stays off by default.
static int enqueue_carry(hfb_ctx* ctx, int B) {
  const int slots = ctx->stream_mode == 0 ? 1 : B;
  dim3 grid(std::max(1, std::min(128, ctx->kp_cap / 8)), slots);   // 1 MB per slot at the default capacity: enough CTAs to move it in ~3 us
  const bool side = ctx->carry_side && ctx->fork_branches && !ctx->prof_on;
  cudaStream_t cs = side ? ctx->copy_stream : ctx->stream;
  if (side) {
    HFB_CUDA(ctx, cudaEventRecord(ctx->ev_carry_fork, ctx->stream));
    HFB_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_carry_fork, 0));
  }
  carry_prev_kernel<<<grid, 256, 0, cs>>>(ctx->d_kdesc, ctx->d_kcount, ctx->d_stream_state, ctx->n_levels, ctx->kp_cap,
                                         ctx->stream_mode, B);
  HFB_CHECK_LAUNCH(ctx, "carry_prev");
  if (side) HFB_CUDA(ctx, cudaEventRecord(ctx->ev_carry, cs));
  return HFB_OK;
}

// Enqueues pyramid + encoder + selection of every level for frames already in lv[0].d_img.
static int enqueue_extract(hfb_ctx* ctx, int B, const int32_t* n_per_level, float threshold) {
  HFB_TRY(enqueue_carry(ctx, B));
  const size_t nb = (size_t)ctx->cfg.max_batch;
  for (int l = 1; l < ctx->n_levels; ++l) {   // the pyramid first: level l = cv::resize of level l - 1
    LevelPlan &lv = ctx->lv[l], &pv = ctx->lv[l - 1];
    HFB_TRY(launch_resize(ctx, pv.d_img, pv.H, pv.W, lv.d_img, lv.H, lv.W, lv.d_xi, lv.d_xa, lv.d_yi, lv.d_ya, B));
  }
  // Network + NMS + top-k of every level; levels >= 1 on their own streams (forked here, joined before the sampling
  // kernels, whose row offsets need the selected counts of the levels below).
  const bool fork_lv = ctx->n_levels > 1 && ctx->fork_levels && ctx->fork_branches && !ctx->prof_on;
  cudaStream_t main_stream = ctx->stream;
  struct StreamGuard {   // error returns must not leave the context on a level stream
    hfb_ctx* c;
    cudaStream_t s;
    ~StreamGuard() { c->stream = s; }
  } guard{ctx, main_stream};
  if (fork_lv) HFB_CUDA(ctx, cudaEventRecord(ctx->ev_level_fork, main_stream));
  for (int l = ctx->n_levels - 1; l >= 0; --l) {   // level 0 last: its enqueue leaves the global branch pending
    LevelPlan& lv = ctx->lv[l];
    const bool own = fork_lv && l > 0;
    if (own) {
      HFB_CUDA(ctx, cudaStreamWaitEvent(ctx->level_stream[l], ctx->ev_level_fork, 0));
      ctx->stream = ctx->level_stream[l];
    }
    HFB_TRY(encoder_forward(ctx, l, B, threshold));
    HFB_TRY(launch_select(ctx, lv.d_nms, lv.H8, lv.W8, lv.d_cand, lv.d_cand_count, ctx->cand_cap,
                          ctx->d_sel + (size_t)l * nb * 8192, ctx->d_nsel + (size_t)l * nb, n_per_level[l], threshold, B,
                          ctx->d_overflow, true));
    if (own) {
      HFB_CUDA(ctx, cudaEventRecord(ctx->ev_level_join[l], ctx->level_stream[l]));
      ctx->stream = main_stream;
    }
  }
  for (int l = 1; fork_lv && l < ctx->n_levels; ++l) HFB_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_level_join[l], 0));
  if (ctx->carry_side && ctx->fork_branches && !ctx->prof_on) HFB_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->ev_carry, 0));   // join the carry
  for (int l = 0; l < ctx->n_levels; ++l) {
    LevelPlan& lv = ctx->lv[l];
    HFB_TRY(launch_sample(ctx, lv.H8, lv.W8, lv.d_descmap, lv.H8 / 8, lv.W8 / 8, ctx->d_sel + (size_t)l * nb * 8192,
                          ctx->d_nsel, (int)nb, n_per_level[l], lv.scale, l, B, ctx->kp_cap, ctx->d_kx, ctx->d_ky,
                          ctx->d_kresp, ctx->d_koct, ctx->d_kdesc, ctx->d_kcount));
  }
  if (ctx->cam.on)   // Frame::UndistortKeyPoints: mvKeysUn next to mvKeys (one launch over all frames)
    HFB_TRY(launch_undistort(ctx, ctx->d_kx, ctx->d_ky, ctx->d_kxu, ctx->d_kyu, ctx->kp_cap, B, ctx->kp_cap, ctx->d_kcount));
  // everything that does not depend on the global branch happens now (main stream), overlapping the side stream:
  // the optional frame-to-previous-frame association and the transfer of the local features
  const hfb_ctx::D2HPlan& d = ctx->d2h;
  // Only the `budget` = sum(n_per_level) rows a frame can hold are transferred (the caller's arrays are sized for that,
  // include/hfnet_b200.h): one strided copy per field, frame b's rows at b * kp_cap on both sides.
  size_t budget = 0;
  for (int l = 0; l < ctx->n_levels; ++l) budget += (size_t)n_per_level[l];
  auto rows_d2h = [&](void* dst, const void* src, size_t elem_bytes, cudaStream_t st) -> cudaError_t {
    if (budget == 0) return cudaSuccess;
    const size_t pb = (size_t)ctx->kp_cap * elem_bytes, wb = budget * elem_bytes;
    if (B == 1 || wb == pb) return cudaMemcpyAsync(dst, src, B == 1 ? wb : pb * B, cudaMemcpyDeviceToHost, st);
    return cudaMemcpy2DAsync(dst, pb, src, pb, wb, (size_t)B, cudaMemcpyDeviceToHost, st);
  };
  if (d.on) {   // feature transfer on its own stream: the copy engine works while the matching kernels run
    cudaStream_t cs = ctx->copy_stream;
    HFB_CUDA(ctx, cudaEventRecord(ctx->ev_local, ctx->stream));
    HFB_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_local, 0));
    HFB_CUDA(ctx, cudaMemcpyAsync(d.counts, ctx->d_kcount, (size_t)B * HFB_MAX_LEVELS * 4, cudaMemcpyDeviceToHost, cs));
    HFB_CUDA(ctx, cudaMemcpyAsync(d.overflow, ctx->d_overflow, 4, cudaMemcpyDeviceToHost, cs));
    HFB_CUDA(ctx, rows_d2h(d.x, ctx->d_kx, 4, cs));
    HFB_CUDA(ctx, rows_d2h(d.y, ctx->d_ky, 4, cs));
    HFB_CUDA(ctx, rows_d2h(d.r, ctx->d_kresp, 4, cs));
    HFB_CUDA(ctx, rows_d2h(d.o, ctx->d_koct, 4, cs));
    if (d.d) HFB_CUDA(ctx, rows_d2h(d.d, ctx->d_kdesc, (size_t)HFB_DESC_DIM * 4, cs));   // NULL: descriptors stay resident
    HFB_CUDA(ctx, cudaEventRecord(ctx->ev_copied, cs));
  }
  if (ctx->fmatch.on) HFB_TRY(enqueue_match_consecutive(ctx, B, ctx->fmatch.mode, ctx->fmatch.thr));
  if (d.on && d.match_idx) {
    const size_t o = (size_t)ctx->cm_shift * ctx->kp_cap;
    HFB_CUDA(ctx, rows_d2h(d.match_idx, ctx->d_cm_idx + o, 4, ctx->stream));
    HFB_CUDA(ctx, rows_d2h(d.match_val, ctx->d_cm_val + o, 4, ctx->stream));
  }
  if (d.on) HFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
  if (ctx->join_pending) {
    ctx->join_pending = false;
    HFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  }
  if (d.on && d.g)
    HFB_CUDA(ctx, cudaMemcpyAsync(d.g, ctx->d_global, (size_t)B * HFB_GLOBAL_DIM * 4, cudaMemcpyDeviceToHost, ctx->stream));
  // state[0] = B for the next call's carry (B <= 64 fits the low byte; the upper bytes stay zero)
  HFB_CUDA(ctx, cudaMemsetAsync(ctx->d_stream_state, B, 1, ctx->stream));
  return HFB_OK;
}

static int run_extract(hfb_ctx* ctx, int B, const int32_t* n_per_level, float threshold) {
  HFB_REQUIRE(ctx, ctx->weights_loaded, "weights not loaded");
  HFB_REQUIRE(ctx, B >= 1 && B <= ctx->cfg.max_batch, "batch size outside [1, max_batch]");
  for (int l = 0; l < ctx->n_levels; ++l)
    HFB_REQUIRE(ctx, n_per_level[l] >= 0 && n_per_level[l] <= ctx->cfg.max_keypoints,
                "per-level keypoint budget outside [0, max_keypoints]");
  ctx->last_batch = B;
  ctx->last_threshold = threshold;
  ctx->kun_valid = ctx->cam.on;
  for (int l = 0; l < ctx->n_levels; ++l) ctx->last_budget[l] = n_per_level[l];
  if (!ctx->use_graph) return enqueue_extract(ctx, B, n_per_level, threshold);
  std::vector<int> key;
  key.push_back(B);
  key.push_back(ctx->stream_mode);
  for (int l = 0; l < ctx->n_levels; ++l) key.push_back(n_per_level[l]);
  int tb;
  memcpy(&tb, &threshold, 4);
  key.push_back(tb);
  if (ctx->fmatch.on) {
    int tt;
    memcpy(&tt, &ctx->fmatch.thr, 4);
    key.push_back(0x4d415443 + ctx->fmatch.mode);
    key.push_back(tt);
  }
  if (ctx->d2h.on) {   // the captured copies target these host addresses
    const void* ps[10] = {ctx->d2h.x, ctx->d2h.y, ctx->d2h.r, ctx->d2h.d, ctx->d2h.g, ctx->d2h.o, ctx->d2h.counts, ctx->d2h.overflow,
                          ctx->d2h.match_idx, ctx->d2h.match_val};
    for (const void* q : ps) {
      const uint64_t v = (uint64_t)(uintptr_t)q;
      key.push_back((int)(v & 0xffffffffu));
      key.push_back((int)(v >> 32));
    }
  }
  for (auto& g : ctx->graphs)
    if (g.key == key) {
      HFB_CUDA(ctx, cudaGraphLaunch(g.exec, ctx->stream));
      ctx->launches += g.kernels;
      return HFB_OK;
    }
  // Warm (non-captured) run first: sets kernel attributes and surfaces errors outside of capture.
  HFB_TRY(enqueue_extract(ctx, B, n_per_level, threshold));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint64_t before = ctx->launches;
  cudaGraph_t graph = nullptr;
  HFB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue_extract(ctx, B, n_per_level, threshold);
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  const uint64_t kernels = ctx->launches - before;
  ctx->launches = before;  // capture launched nothing
  if (rc != HFB_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) {
    ctx->set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    return HFB_ERR_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    ctx->set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    return HFB_ERR_CUDA;
  }
  if (ctx->graphs.size() >= 16) drop_graphs(ctx);
  ctx->graphs.push_back({key, exec, kernels});
  return HFB_OK;  // results of the warm run are in place
}

extern "C" int hfb_extract_batch_dev(hfb_ctx* ctx, const uint8_t* d_images, int32_t n_images,
                                     const int32_t* n_per_level, float threshold) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, d_images && n_per_level, "null argument");
  HFB_REQUIRE(ctx, n_images >= 1 && n_images <= ctx->cfg.max_batch, "batch size outside [1, max_batch]");
  LevelPlan& l0 = ctx->lv[0];
  if (d_images != l0.d_img)
    HFB_CUDA(ctx, cudaMemcpyAsync(l0.d_img, d_images, (size_t)n_images * l0.H * l0.W, cudaMemcpyDeviceToDevice,
                                  ctx->stream));
  return run_extract(ctx, n_images, n_per_level, threshold);
}

extern "C" int hfb_fetch_features(hfb_ctx* ctx, int32_t image_index, hfb_features* out) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, out && image_index >= 0 && image_index < ctx->last_batch, "bad image index");
  int counts[HFB_MAX_LEVELS];
  HFB_CUDA(ctx, cudaMemcpyAsync(counts, ctx->d_kcount + (size_t)image_index * HFB_MAX_LEVELS, sizeof(counts),
                                cudaMemcpyDeviceToHost, ctx->stream));
  HFB_TRY(check_overflow(ctx));  // synchronises
  int total = 0;
  for (int l = 0; l < HFB_MAX_LEVELS; ++l) {
    out->n_per_level[l] = l < ctx->n_levels ? counts[l] : 0;
    total += out->n_per_level[l];
  }
  out->n_total = total;
  const size_t o = (size_t)image_index * ctx->kp_cap;
  if (total > 0) {
    HFB_CUDA(ctx, cudaMemcpyAsync(out->x, ctx->d_kx + o, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(out->y, ctx->d_ky + o, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(out->response, ctx->d_kresp + o, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(out->octave, ctx->d_koct + o, (size_t)total * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->descriptors)
      HFB_CUDA(ctx, cudaMemcpyAsync(out->descriptors, ctx->d_kdesc + o * HFB_DESC_DIM, (size_t)total * HFB_DESC_DIM * 4,
                                    cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (out->global_descriptor && ctx->cfg.with_global)
    HFB_CUDA(ctx, cudaMemcpyAsync(out->global_descriptor, ctx->d_global + (size_t)image_index * HFB_GLOBAL_DIM,
                                  HFB_GLOBAL_DIM * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

extern "C" int hfb_extract_batch(hfb_ctx* ctx, const uint8_t* const* images, int32_t n_images, int32_t stride,
                                 const int32_t* n_per_level, float threshold, hfb_features* outs) {
  return hfb_extract_match_batch(ctx, images, n_images, stride, n_per_level, threshold, outs, -1, 0.f, nullptr, nullptr);
}

// hfb_extract_batch + hfb_match_consecutive in one call (match_mode < 0: extraction only).  The association is enqueued
// inside the extraction, ahead of the global branch's join.
extern "C" int hfb_extract_match_batch(hfb_ctx* ctx, const uint8_t* const* images, int32_t n_images, int32_t stride,
                                       const int32_t* n_per_level, float threshold, hfb_features* outs,
                                       int32_t match_mode, float match_thr, int32_t* match_idx, float* match_val) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, images && n_per_level && outs, "null argument");
  const bool want_match = match_mode >= 0;
  HFB_REQUIRE(ctx, !want_match || ((match_mode == 0 || match_mode == 1) && match_idx && match_val),
              "match mode must be 0 (l2) or 1 (cos) with non-null outputs");
  HFB_REQUIRE(ctx, n_images >= 1 && n_images <= ctx->cfg.max_batch, "batch size outside [1, max_batch]");
  LevelPlan& l0 = ctx->lv[0];
  HFB_REQUIRE(ctx, stride >= l0.W, "stride smaller than the image width");
  // Pinned staging (the reference copies from pageable memory with a synchronous cudaMemcpy, TensorRTBuffers.h:417-435)
  StageTimer tm(ctx->trace, "hfb_extract_batch");
  const size_t img_bytes = (size_t)l0.H * l0.W;
  const size_t per_frame_out = (size_t)ctx->kp_cap * (4 * 4 + HFB_DESC_DIM * 4) + HFB_GLOBAL_DIM * 4 + 64;
  HFB_TRY(ctx->ensure_stage((size_t)n_images * (img_bytes + per_frame_out)));
  uint8_t* hs = reinterpret_cast<uint8_t*>(ctx->h_stage);
  // inputs: page-locked, densely packed frames are DMA'd in place (one transfer when the frames are also contiguous
  // over the batch, e.g. slots of one capture ring); anything else goes through the pinned stage
  bool one_block = stride == l0.W && images[0] != nullptr;
  for (int b = 1; b < n_images && one_block; ++b) one_block = images[b] == images[0] + (size_t)b * img_bytes;
  if (one_block && n_images > 1 && is_pinned(images[0]) && is_pinned(images[0] + (size_t)n_images * img_bytes - 1)) {
    // adjacent addresses can still belong to separate page-locked allocations: the runtime then rejects the copy
    // (synchronously, nothing enqueued) and the frames go one by one
    if (cudaMemcpyAsync(l0.d_img, images[0], (size_t)n_images * img_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
      cudaGetLastError();
      one_block = false;
    }
  } else {
    one_block = false;
  }
  if (!one_block) {
  for (int b = 0; b < n_images; ++b) {
    HFB_REQUIRE(ctx, images[b] != nullptr, "null image");
    if (stride == l0.W && is_pinned(images[b])) {
      HFB_CUDA(ctx, cudaMemcpyAsync(l0.d_img + (size_t)b * img_bytes, images[b], img_bytes, cudaMemcpyHostToDevice,
                                    ctx->stream));
    } else {
      uint8_t* dst = hs + (size_t)b * img_bytes;
      for (int y = 0; y < l0.H; ++y) memcpy(dst + (size_t)y * l0.W, images[b] + (size_t)y * stride, l0.W);
      HFB_CUDA(ctx, cudaMemcpyAsync(l0.d_img + (size_t)b * img_bytes, dst, img_bytes, cudaMemcpyHostToDevice,
                                    ctx->stream));
    }
  }
  }
  tm.mark("stage_in");
  // Fast path: page-locked outputs that are contiguous over the batch (frame b's rows start at b * kp_cap of one array
  // per field) are written by the extraction graph itself, the local features ahead of the global branch's join.
  {
    bool fast = true;
    const hfb_features& f0 = outs[0];
    for (int b = 0; b < n_images && fast; ++b) {
      const hfb_features& f = outs[b];
      const size_t o = (size_t)b * ctx->kp_cap;
      fast = f.x == f0.x + o && f.y == f0.y + o && f.response == f0.response + o && f.octave == f0.octave + o &&
             (f0.descriptors ? f.descriptors == f0.descriptors + o * HFB_DESC_DIM : f.descriptors == nullptr) &&
             (!ctx->cfg.with_global ? true
                                    : (f0.global_descriptor ? f.global_descriptor == f0.global_descriptor + (size_t)b * HFB_GLOBAL_DIM
                                                            : f.global_descriptor == nullptr));
    }
    fast = fast && is_pinned(f0.x) && is_pinned(f0.y) && is_pinned(f0.response) && is_pinned(f0.octave) &&
           (!f0.descriptors || is_pinned(f0.descriptors)) && (!f0.global_descriptor || is_pinned(f0.global_descriptor)) &&
           (!want_match || (is_pinned(match_idx) && is_pinned(match_val)));
    if (fast) {
      int* hc = reinterpret_cast<int*>(hs + (size_t)n_images * img_bytes);   // counts + overflow flag in the pinned stage
      ctx->d2h.on = true;
      ctx->d2h.x = f0.x; ctx->d2h.y = f0.y; ctx->d2h.r = f0.response; ctx->d2h.o = f0.octave;
      ctx->d2h.d = f0.descriptors;
      ctx->d2h.g = ctx->cfg.with_global ? f0.global_descriptor : nullptr;
      ctx->d2h.counts = hc;
      ctx->d2h.overflow = hc + (size_t)n_images * HFB_MAX_LEVELS;
      ctx->d2h.match_idx = want_match ? match_idx : nullptr;
      ctx->d2h.match_val = want_match ? match_val : nullptr;
      ctx->fmatch.on = want_match;
      ctx->fmatch.mode = match_mode;
      ctx->fmatch.thr = match_thr;
      const int rc = run_extract(ctx, n_images, n_per_level, threshold);
      ctx->d2h.on = false;
      ctx->fmatch.on = false;
      HFB_TRY(rc);
      tm.mark("enqueue");
      HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      tm.mark("gpu_wait");
      if (*ctx->d2h.overflow) {
        cudaMemsetAsync(ctx->d_overflow, 0, sizeof(int), ctx->stream);
        ctx->set_error("more threshold-scan candidates than the context's candidate capacity (threshold too low)");
        return HFB_ERR_CAPACITY;
      }
      for (int b = 0; b < n_images; ++b) {
        hfb_features& f = outs[b];
        int total = 0;
        for (int l = 0; l < HFB_MAX_LEVELS; ++l) {
          f.n_per_level[l] = l < ctx->n_levels ? hc[(size_t)b * HFB_MAX_LEVELS + l] : 0;
          total += f.n_per_level[l];
        }
        f.n_total = total;
      }
      return HFB_OK;
    }
  }
  HFB_TRY(run_extract(ctx, n_images, n_per_level, threshold));
  tm.mark("enqueue");
  // outputs: one D2H burst of budget-sized slices, one synchronisation.  Page-locked caller arrays receive the DMA
  // directly; pageable ones are filled from the pinned stage after the sync.
  int budget = 0;
  for (int l = 0; l < ctx->n_levels; ++l) budget += n_per_level[l];
  uint8_t* ho = hs + (size_t)n_images * img_bytes;
  struct Slot { int* counts; float *x, *y, *r; int* o; float *d, *g; bool direct; };
  std::vector<Slot> slots(n_images);
  for (int b = 0; b < n_images; ++b) {
    uint8_t* q = ho + (size_t)b * per_frame_out;
    Slot& s = slots[b];
    hfb_features& f = outs[b];
    s.direct = (!f.descriptors || is_pinned(f.descriptors)) && is_pinned(f.x) && is_pinned(f.y) && is_pinned(f.response) &&
               is_pinned(f.octave) && (!f.global_descriptor || is_pinned(f.global_descriptor));
    s.counts = reinterpret_cast<int*>(q); q += 64;
    s.x = s.direct ? f.x : reinterpret_cast<float*>(q); q += (size_t)ctx->kp_cap * 4;
    s.y = s.direct ? f.y : reinterpret_cast<float*>(q); q += (size_t)ctx->kp_cap * 4;
    s.r = s.direct ? f.response : reinterpret_cast<float*>(q); q += (size_t)ctx->kp_cap * 4;
    s.o = s.direct ? f.octave : reinterpret_cast<int*>(q); q += (size_t)ctx->kp_cap * 4;
    s.d = s.direct ? f.descriptors : reinterpret_cast<float*>(q); q += (size_t)ctx->kp_cap * HFB_DESC_DIM * 4;
    s.g = (s.direct && f.global_descriptor) ? f.global_descriptor : reinterpret_cast<float*>(q);
    const size_t o = (size_t)b * ctx->kp_cap;
    HFB_CUDA(ctx, cudaMemcpyAsync(s.counts, ctx->d_kcount + (size_t)b * HFB_MAX_LEVELS, HFB_MAX_LEVELS * 4,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    if (budget > 0) {
      HFB_CUDA(ctx, cudaMemcpyAsync(s.x, ctx->d_kx + o, (size_t)budget * 4, cudaMemcpyDeviceToHost, ctx->stream));
      HFB_CUDA(ctx, cudaMemcpyAsync(s.y, ctx->d_ky + o, (size_t)budget * 4, cudaMemcpyDeviceToHost, ctx->stream));
      HFB_CUDA(ctx, cudaMemcpyAsync(s.r, ctx->d_kresp + o, (size_t)budget * 4, cudaMemcpyDeviceToHost, ctx->stream));
      HFB_CUDA(ctx, cudaMemcpyAsync(s.o, ctx->d_koct + o, (size_t)budget * 4, cudaMemcpyDeviceToHost, ctx->stream));
      if (f.descriptors)
        HFB_CUDA(ctx, cudaMemcpyAsync(s.d, ctx->d_kdesc + o * HFB_DESC_DIM, (size_t)budget * HFB_DESC_DIM * 4,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (ctx->cfg.with_global && (f.global_descriptor || !s.direct))
      HFB_CUDA(ctx, cudaMemcpyAsync(s.g, ctx->d_global + (size_t)b * HFB_GLOBAL_DIM, HFB_GLOBAL_DIM * 4,
                                    cudaMemcpyDeviceToHost, ctx->stream));
  }
  tm.mark("enqueue_d2h");
  HFB_TRY(check_overflow(ctx));  // synchronises the stream
  tm.mark("gpu_wait");
  for (int b = 0; b < n_images; ++b) {
    const Slot& s = slots[b];
    hfb_features& f = outs[b];
    int total = 0;
    for (int l = 0; l < HFB_MAX_LEVELS; ++l) {
      f.n_per_level[l] = l < ctx->n_levels ? s.counts[l] : 0;
      total += f.n_per_level[l];
    }
    f.n_total = total;
    if (s.direct) continue;
    if (total > 0) {
      memcpy(f.x, s.x, (size_t)total * 4);
      memcpy(f.y, s.y, (size_t)total * 4);
      memcpy(f.response, s.r, (size_t)total * 4);
      memcpy(f.octave, s.o, (size_t)total * 4);
      if (f.descriptors) memcpy(f.descriptors, s.d, (size_t)total * HFB_DESC_DIM * 4);
    }
    if (f.global_descriptor && ctx->cfg.with_global) memcpy(f.global_descriptor, s.g, HFB_GLOBAL_DIM * 4);
  }
  tm.mark("copy_out");
  if (want_match) HFB_TRY(hfb_match_consecutive(ctx, n_images, match_mode, match_thr, match_idx, match_val));
  return HFB_OK;
}

extern "C" int hfb_extract(hfb_ctx* ctx, const uint8_t* image, int32_t height, int32_t width, int32_t stride,
                           const int32_t* n_per_level, float threshold, hfb_features* out) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, height == ctx->cfg.height && width == ctx->cfg.width,
              "image size differs from the context's (the reference builds one engine per fixed shape too, "
              "BaseModel.cc:35-65)");
  const uint8_t* imgs[1] = {image};
  return hfb_extract_batch(ctx, imgs, 1, stride, n_per_level, threshold, out);
}

// BaseModel::Detect of ONE pyramid level on a context that holds all levels (src/Extractors/HFNetRTModel.cc:84-110: the
// reference keeps one engine per level and HFextractor hands every level object its own pre-scaled image,
// HFextractor.cc:228-243): `image` is the level's image, results are in LEVEL coordinates with octave 0 -- the caller
// (HFextractor.cc:272-279) applies octave / scale.  The fused multi-level call (hfb_extract) is the fast path; this entry
// keeps the unmodified per-level flow working on one shared context and one copy of the weights.
extern "C" int hfb_extract_level(hfb_ctx* ctx, int32_t level, const uint8_t* image, int32_t height, int32_t width,
                                 int32_t stride, int32_t n_keypoints, float threshold, hfb_features* out) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, ctx->weights_loaded, "weights not loaded");
  HFB_REQUIRE(ctx, image && out && level >= 0 && level < ctx->n_levels, "bad argument");
  LevelPlan& lv = ctx->lv[level];
  HFB_REQUIRE(ctx, height == lv.H && width == lv.W, "image size differs from this level's (one fixed shape per level, BaseModel.cc:35-65)");
  HFB_REQUIRE(ctx, stride >= width, "stride smaller than the image width");
  HFB_REQUIRE(ctx, n_keypoints >= 0 && n_keypoints <= ctx->cfg.max_keypoints, "keypoint budget outside [0, max_keypoints]");
  const size_t img_bytes = (size_t)lv.H * lv.W;
  const size_t out_bytes = 64 + (size_t)n_keypoints * (16 + HFB_DESC_DIM * 4) + HFB_GLOBAL_DIM * 4;
  HFB_TRY(ctx->ensure_stage(img_bytes + out_bytes));
  uint8_t* hs = reinterpret_cast<uint8_t*>(ctx->h_stage);
  for (int y = 0; y < lv.H; ++y) memcpy(hs + (size_t)y * lv.W, image + (size_t)y * stride, lv.W);
  cudaStream_t st = ctx->stream;
  HFB_CUDA(ctx, cudaMemcpyAsync(lv.d_img, hs, img_bytes, cudaMemcpyHostToDevice, st));
  ctx->last_batch = 1;
  ctx->last_threshold = threshold;
  ctx->kun_valid = false;   // level coordinates: the caller scales, then undistorts (hfb_undistort_points)
  for (int l = 0; l < ctx->n_levels; ++l) ctx->last_budget[l] = l == 0 ? n_keypoints : 0;
  HFB_CUDA(ctx, cudaMemsetAsync(ctx->d_kcount, 0, HFB_MAX_LEVELS * sizeof(int), st));
  HFB_CUDA(ctx, cudaMemsetAsync(ctx->d_stream_state, 0, sizeof(int), st));   // no streaming history through this entry
  HFB_TRY(encoder_forward(ctx, level, 1, threshold));
  HFB_TRY(launch_select_sample(ctx, lv.d_nms, lv.H8, lv.W8, lv.d_descmap, lv.H8 / 8, lv.W8 / 8, lv.d_cand, lv.d_cand_count,
                               ctx->cand_cap, ctx->d_sel, ctx->d_nsel, n_keypoints, threshold, 1.0f, 1, ctx->kp_cap,
                               ctx->d_kx, ctx->d_ky, ctx->d_kresp, ctx->d_koct, ctx->d_kdesc, ctx->d_kcount,
                               ctx->d_overflow, true));
  if (ctx->join_pending) {
    ctx->join_pending = false;
    HFB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
  }
  uint8_t* ho = hs + img_bytes;
  int* h_cnt = reinterpret_cast<int*>(ho);
  float* hx = reinterpret_cast<float*>(ho + 64);
  float* hy = hx + n_keypoints;
  float* hr = hy + n_keypoints;
  int* hoct = reinterpret_cast<int*>(hr + n_keypoints);
  float* hd = reinterpret_cast<float*>(hoct + n_keypoints);
  float* hg = hd + (size_t)n_keypoints * HFB_DESC_DIM;
  HFB_CUDA(ctx, cudaMemcpyAsync(h_cnt, ctx->d_kcount, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (n_keypoints > 0) {
    HFB_CUDA(ctx, cudaMemcpyAsync(hx, ctx->d_kx, (size_t)n_keypoints * 4, cudaMemcpyDeviceToHost, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(hy, ctx->d_ky, (size_t)n_keypoints * 4, cudaMemcpyDeviceToHost, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(hr, ctx->d_kresp, (size_t)n_keypoints * 4, cudaMemcpyDeviceToHost, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(hoct, ctx->d_koct, (size_t)n_keypoints * 4, cudaMemcpyDeviceToHost, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(hd, ctx->d_kdesc, (size_t)n_keypoints * HFB_DESC_DIM * 4, cudaMemcpyDeviceToHost, st));
  }
  const bool want_g = lv.global && out->global_descriptor;
  if (want_g) HFB_CUDA(ctx, cudaMemcpyAsync(hg, ctx->d_global, HFB_GLOBAL_DIM * 4, cudaMemcpyDeviceToHost, st));
  HFB_TRY(check_overflow(ctx));   // synchronises
  const int n = std::min(h_cnt[0], n_keypoints);
  for (int l = 0; l < HFB_MAX_LEVELS; ++l) out->n_per_level[l] = l == 0 ? n : 0;
  out->n_total = n;
  if (n > 0) {
    memcpy(out->x, hx, (size_t)n * 4);
    memcpy(out->y, hy, (size_t)n * 4);
    memcpy(out->response, hr, (size_t)n * 4);
    memcpy(out->octave, hoct, (size_t)n * 4);
    memcpy(out->descriptors, hd, (size_t)n * HFB_DESC_DIM * 4);
  }
  if (want_g) memcpy(out->global_descriptor, hg, HFB_GLOBAL_DIM * 4);
  return HFB_OK;
}

// Per-launch timing of one (un-graphed) extraction of the frames already resident in the context: JSON array of
// {"name", "ms", "bytes", "flops"} with the launcher-stated algorithmic bytes / flops (bench.py's roofline source).
extern "C" int hfb_profile_extract(hfb_ctx* ctx, int32_t n_images, const int32_t* n_per_level, float threshold,
                                   char* json_out, size_t cap) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, ctx->weights_loaded && n_per_level && json_out && cap > 2, "bad argument");
  HFB_REQUIRE(ctx, n_images >= 1 && n_images <= ctx->cfg.max_batch, "batch size outside [1, max_batch]");
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof.clear();
  ctx->prof_on = true;
  ctx->label = "start";
  ctx->prof_mark("start");
  int rc = enqueue_extract(ctx, n_images, n_per_level, threshold);
  ctx->prof_on = false;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  std::string js = "[";
  for (size_t i = 1; i < ctx->prof.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->prof[i - 1].ev, ctx->prof[i].ev);
    char buf[256];
    snprintf(buf, sizeof(buf), "%s{\"name\":\"%s\",\"ms\":%.6f,\"bytes\":%.0f,\"flops\":%.0f}", i > 1 ? "," : "",
             ctx->prof[i].name.c_str(), ms, ctx->prof[i].bytes, ctx->prof[i].flops);
    js += buf;
  }
  js += "]";
  for (auto& r : ctx->prof) cudaEventDestroy(r.ev);
  ctx->prof.clear();
  if (rc != HFB_OK) return rc;
  if (e != cudaSuccess) {
    ctx->set_error(std::string("hfb_profile_extract: ") + cudaGetErrorString(e));
    return HFB_ERR_CUDA;
  }
  HFB_REQUIRE(ctx, js.size() + 1 <= cap, "profile buffer too small");
  memcpy(json_out, js.c_str(), js.size() + 1);
  return HFB_OK;
}

// ------------------------------------------------------------------------------------------------ parity hooks
extern "C" int hfb_nms(hfb_ctx* ctx, const float* scores, int32_t height, int32_t width, float* scores_nms) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, scores && scores_nms && height > 0 && width > 0, "bad argument");
  const size_t n = (size_t)height * width;
  HFB_TRY(ctx->ensure_scratch(2 * n * 4));
  float* d_in = reinterpret_cast<float*>(ctx->d_scratch);
  float* d_out = d_in + n;
  HFB_CUDA(ctx, cudaMemcpyAsync(d_in, scores, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_TRY(launch_nms(ctx, d_in, d_out, height, width, 1, 0.f, nullptr, nullptr, 0));
  HFB_CUDA(ctx, cudaMemcpyAsync(scores_nms, d_out, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

extern "C" int hfb_select_sample(hfb_ctx* ctx, const float* scores_nms, int32_t height, int32_t width,
                                 const float* desc_map, int32_t desc_h, int32_t desc_w, int32_t n_keypoints,
                                 float threshold, float* x, float* y, float* response, float* descriptors,
                                 int32_t* n_out) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, scores_nms && desc_map && n_out && height > 0 && width > 0 && desc_h > 0 && desc_w > 0,
              "bad argument");
  HFB_REQUIRE(ctx, n_keypoints >= 0 && n_keypoints <= 8192, "n_keypoints outside [0, 8192]");
  const size_t n = (size_t)height * width, nd = (size_t)desc_h * desc_w * 256;
  const int cap = (int)std::min<size_t>(n, 1 << 20);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_desc = al(n * 4), o_cand = o_desc + al(nd * 4), o_cnt = o_cand + al((size_t)cap * 8);
  const size_t o_x = o_cnt + 256, o_y = o_x + al(8192 * 4), o_r = o_y + al(8192 * 4), o_o = o_r + al(8192 * 4);
  const size_t o_d = o_o + al(8192 * 4), o_k = o_d + al((size_t)8192 * 256 * 4), total = o_k + 256;
  HFB_TRY(ctx->ensure_scratch(total));
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->d_scratch);
  float* d_nms = reinterpret_cast<float*>(base);
  float* d_dm = reinterpret_cast<float*>(base + o_desc);
  u64* d_cand = reinterpret_cast<u64*>(base + o_cand);
  int* d_cnt = reinterpret_cast<int*>(base + o_cnt);
  float *d_x = reinterpret_cast<float*>(base + o_x), *d_y = reinterpret_cast<float*>(base + o_y);
  float* d_r = reinterpret_cast<float*>(base + o_r);
  int* d_o = reinterpret_cast<int*>(base + o_o);
  float* d_d = reinterpret_cast<float*>(base + o_d);
  int* d_k = reinterpret_cast<int*>(base + o_k);
  HFB_CUDA(ctx, cudaMemcpyAsync(d_nms, scores_nms, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(d_dm, desc_map, nd * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemsetAsync(d_k, 0, HFB_MAX_LEVELS * 4, ctx->stream));
  HFB_TRY(launch_select_sample(ctx, d_nms, height, width, d_dm, desc_h, desc_w, d_cand, d_cnt, cap, ctx->d_sel,
                               ctx->d_nsel, n_keypoints, threshold, 1.0f, 1, 8192, d_x, d_y, d_r, d_o, d_d, d_k,
                               ctx->d_overflow, false));
  int cnt = 0;
  HFB_CUDA(ctx, cudaMemcpyAsync(&cnt, d_k, 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_TRY(check_overflow(ctx));
  *n_out = cnt;
  if (cnt > 0) {
    HFB_CUDA(ctx, cudaMemcpyAsync(x, d_x, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(y, d_y, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(response, d_r, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(descriptors, d_d, (size_t)cnt * 256 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return HFB_OK;
}

extern "C" int hfb_resize_linear_u8(hfb_ctx* ctx, const uint8_t* src, int32_t sh, int32_t sw, uint8_t* dst, int32_t dh,
                                    int32_t dw) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, src && dst && sh > 0 && sw > 0 && dh > 0 && dw > 0, "bad argument");
  std::vector<int> xi, yi;
  std::vector<short> xa, ya;
  build_resize_tables(sw, dw, xi, xa);
  build_resize_tables(sh, dh, yi, ya);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_dst = al((size_t)sh * sw), o_xi = o_dst + al((size_t)dh * dw), o_xa = o_xi + al(xi.size() * 4);
  const size_t o_yi = o_xa + al(xa.size() * 2), o_ya = o_yi + al(yi.size() * 4), total = o_ya + al(ya.size() * 2);
  HFB_TRY(ctx->ensure_scratch(total));
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->d_scratch);
  HFB_CUDA(ctx, cudaMemcpyAsync(base, src, (size_t)sh * sw, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(base + o_xi, xi.data(), xi.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(base + o_xa, xa.data(), xa.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(base + o_yi, yi.data(), yi.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(base + o_ya, ya.data(), ya.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
  HFB_TRY(launch_resize(ctx, base, sh, sw, base + o_dst, dh, dw, reinterpret_cast<int*>(base + o_xi),
                        reinterpret_cast<short*>(base + o_xa), reinterpret_cast<int*>(base + o_yi),
                        reinterpret_cast<short*>(base + o_ya), 1));
  HFB_CUDA(ctx, cudaMemcpyAsync(dst, base + o_dst, (size_t)dh * dw, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // also keeps the host tables alive until the copies are done
  return HFB_OK;
}

__global__ void h2f_kernel(const __half* __restrict__ in, int ld, int col0, int cols, float* __restrict__ out,
                           long long rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  out[i] = __half2float(in[r * ld + col0 + c]);
}

extern "C" int hfb_debug_tensor(hfb_ctx* ctx, const char* name, int32_t image_index, int32_t level, float* out,
                                size_t cap, size_t* n, int32_t* dims4) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, name && n && dims4, "null argument");
  HFB_REQUIRE(ctx, ctx->weights_loaded && ctx->last_batch > 0, "no extraction has run yet");
  HFB_REQUIRE(ctx, level >= 0 && level < ctx->n_levels && image_index >= 0 && image_index < ctx->last_batch,
              "bad level / image index");
  LevelPlan& lv = ctx->lv[level];
  const std::string s(name);
  const int Hd = lv.H8 / 8, Wd = lv.W8 / 8;
  const __half* hsrc = nullptr;
  const float* fsrc = nullptr;
  int ld = 0, col0 = 0;
  int d[4] = {1, 0, 0, 0};
  if (s.rfind("layer_", 0) == 0) {
    const int L = atoi(s.c_str() + 6);
    HFB_REQUIRE(ctx, L >= 1 && L <= lv.n_act, "layer not computed for this level");
    d[1] = lv.aH[L]; d[2] = lv.aW[L]; d[3] = lv.aC[L];
    hsrc = lv.act[L] + (size_t)image_index * d[1] * d[2] * d[3];
    ld = d[3];
  } else if (s == "desc_conv1" || s == "det_conv1") {
    d[1] = Hd; d[2] = Wd; d[3] = s == "desc_conv1" ? 256 : 128;
    hsrc = lv.d_head1 + (size_t)image_index * Hd * Wd * 384;
    ld = 384;
    col0 = s == "desc_conv1" ? 0 : 256;
  } else if (s == "det_logits") {
    HFB_REQUIRE(ctx, lv.d_logits != nullptr, "det_logits are only kept with HFB_DEBUG=1");
    d[1] = Hd; d[2] = Wd; d[3] = 65;
    fsrc = lv.d_logits + (size_t)image_index * Hd * Wd * 65;
  } else if (s == "scores_dense" || s == "scores_dense_nms") {
    d[1] = lv.H8; d[2] = lv.W8; d[3] = 1;
    fsrc = (s == "scores_dense" ? lv.d_scores : lv.d_nms) + (size_t)image_index * lv.H8 * lv.W8;
  } else if (s == "local_descriptor_map") {
    d[1] = Hd; d[2] = Wd; d[3] = 256;
    fsrc = lv.d_descmap + (size_t)image_index * Hd * Wd * 256;
  } else if (s == "vlad_norm") {
    HFB_REQUIRE(ctx, lv.global, "no global head on this level");
    d[1] = 1; d[2] = 1; d[3] = ctx->net.n_clusters * ctx->net.c_global;
    fsrc = lv.d_vladn + (size_t)image_index * d[3];
  } else if (s == "global_descriptor") {
    HFB_REQUIRE(ctx, lv.global, "no global head on this level");
    d[1] = 1; d[2] = 1; d[3] = HFB_GLOBAL_DIM;
    fsrc = ctx->d_global + (size_t)image_index * HFB_GLOBAL_DIM;
  } else {
    ctx->set_error("unknown tensor name: " + s);
    return HFB_ERR_INVALID;
  }
  const size_t cnt = (size_t)d[1] * d[2] * d[3];
  *n = cnt;
  for (int i = 0; i < 4; ++i) dims4[i] = d[i];
  HFB_REQUIRE(ctx, out && cap >= cnt, "output buffer too small");
  if (hsrc) {
    HFB_TRY(ctx->ensure_scratch(cnt * 4));
    float* tmp = reinterpret_cast<float*>(ctx->d_scratch);
    h2f_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(hsrc, ld, col0, d[3], tmp,
                                                                      (long long)d[1] * d[2]);
    HFB_CHECK_LAUNCH(ctx, "h2f");
    fsrc = tmp;
  }
  HFB_CUDA(ctx, cudaMemcpyAsync(out, fsrc, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

// ================================================================================================ matching
extern "C" int hfb_match_batch_dev(hfb_ctx* ctx, int32_t mode, const float* dA_all, int32_t na_total,
                                   const float* dB_all, int32_t nb_total, int32_t n_pairs, const int32_t* d_pair_tab,
                                   int32_t max_a_cnt, int32_t max_b_cnt, float thr, int32_t* d_match_idx,
                                   float* d_match_val) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, mode == 0 || mode == 1, "mode must be 0 (l2) or 1 (cos)");
  HFB_REQUIRE(ctx, dA_all && dB_all && d_pair_tab && d_match_idx && d_match_val, "null argument");
  HFB_REQUIRE(ctx, na_total >= 0 && nb_total >= 0 && n_pairs >= 0 && max_a_cnt >= 0 && max_b_cnt >= 0, "negative size");
  return launch_match_batch(ctx, mode, dA_all, dB_all, n_pairs, d_pair_tab, max_a_cnt, max_b_cnt, thr, d_match_idx,
                            d_match_val, na_total, nb_total, nullptr);
}

// Pair table of the streaming association, rows relative to the window base d_kdesc - shift * kp_cap * 256
// (shift = 1 carried slot in stream_mode 0, B in stream_mode 1): A = frame b, B = the frame before it in its stream --
// frame b-1 of this call or the descriptors carried over from the previous call (no rows on the first call / after
// hfb_reset_stream: the frame then has no matches, like the first frame of a sequence).
__global__ void consecutive_tab_kernel(const int* __restrict__ kcount, const int* __restrict__ state, int n_levels, int B,
                                       int kp_cap, int mode, int* __restrict__ tab) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int shift = mode == 0 ? 1 : B;
  int na = 0, nb = 0;
  for (int l = 0; l < n_levels; ++l) na += kcount[b * HFB_MAX_LEVELS + l];
  if (mode == 0 && b > 0) {
    for (int l = 0; l < n_levels; ++l) nb += kcount[(b - 1) * HFB_MAX_LEVELS + l];
  } else {
    nb = state[1 + (mode == 0 ? 0 : b)];
  }
  tab[b] = (shift + b) * kp_cap;
  tab[B + b] = na;
  tab[2 * B + b] = b * kp_cap;
  tab[3 * B + b] = nb;
}

// Tracking's per-frame descriptor association with the previous frame, device-resident: the descriptors of the last
// hfb_extract_batch*(n_images) never leave HBM.  Frame b is matched against frame (b-1) mod n_images.
extern "C" int hfb_match_consecutive_dev(hfb_ctx* ctx, int32_t n_images, int32_t mode, float thr) {
  HFB_ENTER(ctx);
  return enqueue_match_consecutive(ctx, n_images, mode, thr);
}

// Device-resident extraction + association in one enqueue (see hfb_extract_match_batch).  No sync.
extern "C" int hfb_extract_match_batch_dev(hfb_ctx* ctx, const uint8_t* d_images, int32_t n_images,
                                           const int32_t* n_per_level, float threshold, int32_t match_mode,
                                           float match_thr) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, match_mode == 0 || match_mode == 1, "mode must be 0 (l2) or 1 (cos)");
  ctx->fmatch.on = true;
  ctx->fmatch.mode = match_mode;
  ctx->fmatch.thr = match_thr;
  const int rc = hfb_extract_batch_dev(ctx, d_images, n_images, n_per_level, threshold);
  ctx->fmatch.on = false;
  return rc;
}

static int enqueue_match_consecutive(hfb_ctx* ctx, int n_images, int mode, float thr) {
  HFB_REQUIRE(ctx, mode == 0 || mode == 1, "mode must be 0 (l2) or 1 (cos)");
  HFB_REQUIRE(ctx, n_images >= 1 && n_images <= ctx->last_batch, "n_images exceeds the last extracted batch");
  HFB_REQUIRE(ctx, ctx->stream_mode == 0 || n_images == ctx->last_batch,
              "stream_mode 1 associates every frame of the batch with its own stream's previous frame");
  consecutive_tab_kernel<<<1, 64, 0, ctx->stream>>>(ctx->d_kcount, ctx->d_stream_state, ctx->n_levels, n_images,
                                                   ctx->kp_cap, ctx->stream_mode, ctx->d_cm_tab);
  HFB_CHECK_LAUNCH(ctx, "consecutive_tab");
  int budget = 0;
  for (int l = 0; l < ctx->n_levels; ++l) budget += ctx->last_budget[l];
  const int shift = ctx->stream_mode == 0 ? 1 : n_images;
  ctx->cm_shift = shift;
  const float* base = ctx->d_kdesc - (size_t)shift * ctx->kp_cap * HFB_DESC_DIM;
  const int rows = (shift + n_images) * ctx->kp_cap;
  return launch_match_batch(ctx, mode, base, base, n_images, ctx->d_cm_tab, budget, budget, thr, ctx->d_cm_idx,
                            ctx->d_cm_val, rows, rows, nullptr, ctx->d_cm_ws, ctx->d_cm_ws_bytes, ctx->kp_cap);
}

// Stream layout of a batch and the reset of the association's history (Tracking::Reset / a new map start from an empty
// mLastFrame, src/Tracking.cc:3256-3330).
// --------------------------------------------------------------------------------------------------- calibration
extern "C" int hfb_set_camera(hfb_ctx* ctx, const float* K, const float* dist, int32_t n_dist) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, K && K[0] != 0.f && K[1] != 0.f, "K = (fx, fy, cx, cy) with non-zero focal lengths");
  HFB_REQUIRE(ctx, n_dist == 0 || n_dist == 4 || n_dist == 5 || n_dist == 8 || n_dist == 12,
              "n_dist must be 0, 4, 5, 8 or 12 (OpenCV's coefficient vectors without the tilt terms)");
  HFB_REQUIRE(ctx, n_dist == 0 || dist, "null distortion vector");
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  drop_graphs(ctx);   // the calibration is a kernel argument of the captured extraction
  hfb_ctx::Camera& c = ctx->cam;
  c.fx = K[0]; c.fy = K[1]; c.cx = K[2]; c.cy = K[3];
  for (int i = 0; i < 12; ++i) c.k[i] = i < n_dist ? (double)dist[i] : 0.0;
  c.on = n_dist > 0 && dist[0] != 0.f;   // src/Frame.cc:762: only the first coefficient is tested
  ctx->kun_valid = false;
  return HFB_OK;
}

extern "C" int hfb_undistort_points(hfb_ctx* ctx, const float* x, const float* y, int32_t n, float* x_un, float* y_un) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, n >= 0 && (n == 0 || (x && y && x_un && y_un)), "bad argument");
  if (n == 0) return HFB_OK;
  if (!ctx->cam.on) {   // mvKeysUn = mvKeys
    if (x_un != x) memmove(x_un, x, (size_t)n * 4);
    if (y_un != y) memmove(y_un, y, (size_t)n * 4);
    return HFB_OK;
  }
  HFB_TRY(ctx->ensure_scratch((size_t)n * 16));
  float* d = reinterpret_cast<float*>(ctx->d_scratch);
  HFB_CUDA(ctx, cudaMemcpyAsync(d, x, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(d + n, y, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_TRY(launch_undistort(ctx, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, n, 1, 0, nullptr));
  HFB_CUDA(ctx, cudaMemcpyAsync(x_un, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(y_un, d + 3 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

extern "C" int hfb_fetch_undistorted(hfb_ctx* ctx, int32_t image_index, float* x_un, float* y_un, int32_t n) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, image_index >= 0 && image_index < ctx->last_batch, "bad image index");
  HFB_REQUIRE(ctx, n >= 0 && n <= ctx->kp_cap && (n == 0 || (x_un && y_un)), "bad argument");
  if (n == 0) return HFB_OK;
  HFB_REQUIRE(ctx, !ctx->cam.on || ctx->kun_valid,
              "the last extraction did not produce undistorted coordinates (hfb_extract_level, or the camera was set afterwards)");
  const size_t o = (size_t)image_index * ctx->kp_cap;
  const float *sx = ctx->cam.on ? ctx->d_kxu : ctx->d_kx, *sy = ctx->cam.on ? ctx->d_kyu : ctx->d_ky;
  HFB_CUDA(ctx, cudaMemcpyAsync(x_un, sx + o, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(y_un, sy + o, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

extern "C" int hfb_image_bounds(hfb_ctx* ctx, int32_t width, int32_t height, float* bounds) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, bounds && width > 0 && height > 0, "bad argument");
  if (!ctx->cam.on) {
    bounds[0] = 0.f; bounds[1] = (float)width; bounds[2] = 0.f; bounds[3] = (float)height;
    return HFB_OK;
  }
  const float cx[4] = {0.f, (float)width, 0.f, (float)width}, cy[4] = {0.f, 0.f, (float)height, (float)height};
  float ux[4], uy[4];
  HFB_TRY(hfb_undistort_points(ctx, cx, cy, 4, ux, uy));
  bounds[0] = std::min(ux[0], ux[2]);
  bounds[1] = std::max(ux[1], ux[3]);
  bounds[2] = std::min(uy[0], uy[1]);
  bounds[3] = std::max(uy[2], uy[3]);
  return HFB_OK;
}

extern "C" int hfb_set_stream_mode(hfb_ctx* ctx, int32_t mode) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, mode == 0 || mode == 1, "stream mode must be 0 (one stream per batch) or 1 (one stream per batch slot)");
  if (mode != ctx->stream_mode) {
    ctx->stream_mode = mode;
    return hfb_reset_stream(ctx);
  }
  return HFB_OK;
}

extern "C" int hfb_reset_stream(hfb_ctx* ctx) {
  HFB_ENTER(ctx);
  HFB_CUDA(ctx, cudaMemsetAsync(ctx->d_stream_state, 0, (size_t)(1 + ctx->cfg.max_batch) * sizeof(int), ctx->stream));
  return HFB_OK;
}

extern "C" int hfb_fetch_matches(hfb_ctx* ctx, int32_t image_index, int32_t* match_idx, float* match_val, int32_t n) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, match_idx && match_val && image_index >= 0 && image_index < ctx->last_batch && n >= 0 &&
                       n <= ctx->kp_cap, "bad argument");
  const size_t o = (size_t)(ctx->cm_shift + image_index) * ctx->kp_cap;
  HFB_CUDA(ctx, cudaMemcpyAsync(match_idx, ctx->d_cm_idx + o, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(match_val, ctx->d_cm_val + o, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

// Host-result form of hfb_match_consecutive_dev: the descriptors of the last extraction never leave HBM, only the
// match rows ([n_images][kp_cap] indices + values) come back, in one transfer each.
extern "C" int hfb_match_consecutive(hfb_ctx* ctx, int32_t n_images, int32_t mode, float thr, int32_t* match_idx,
                                     float* match_val) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, match_idx && match_val, "null output");
  HFB_TRY(hfb_match_consecutive_dev(ctx, n_images, mode, thr));
  const size_t n = (size_t)n_images * ctx->kp_cap, o = (size_t)ctx->cm_shift * ctx->kp_cap;
  HFB_CUDA(ctx, cudaMemcpyAsync(match_idx, ctx->d_cm_idx + o, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(match_val, ctx->d_cm_val + o, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

// MapPoint::ComputeDistinctiveDescriptors for a ragged batch of map points (src/MapPoint.cc:331-400).
extern "C" int hfb_distinctive_descriptors(hfb_ctx* ctx, const float* descriptors, const int32_t* offsets,
                                           int32_t n_points, int32_t* best_index, float* best_median) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, n_points >= 0 && offsets && best_index && best_median, "bad argument");
  if (n_points == 0) return HFB_OK;
  int max_n = 0;
  HFB_REQUIRE(ctx, offsets[0] == 0, "offsets must start at 0");
  for (int p = 0; p < n_points; ++p) {
    HFB_REQUIRE(ctx, offsets[p + 1] >= offsets[p], "offsets must be non-decreasing");
    max_n = std::max(max_n, offsets[p + 1] - offsets[p]);
  }
  HFB_REQUIRE(ctx, max_n <= 128, "more than 128 observations for one map point");
  const int total = offsets[n_points];
  HFB_REQUIRE(ctx, total == 0 || descriptors, "null descriptors");
  const size_t szD = ((size_t)total * HFB_DESC_DIM * 4 + 255) & ~(size_t)255;
  const size_t szO = ((size_t)(n_points + 1) * 4 + 255) & ~(size_t)255;
  const size_t szR = ((size_t)n_points * 4 + 255) & ~(size_t)255;
  HFB_TRY(ctx->ensure_io(szD + szO + 2 * szR));
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->d_io);
  float* dD = reinterpret_cast<float*>(base);
  int* dO = reinterpret_cast<int*>(base + szD);
  int* dI = reinterpret_cast<int*>(base + szD + szO);
  float* dM = reinterpret_cast<float*>(base + szD + szO + szR);
  if (total > 0)
    HFB_CUDA(ctx, cudaMemcpyAsync(dD, descriptors, (size_t)total * HFB_DESC_DIM * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(dO, offsets, (size_t)(n_points + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_TRY(launch_distinctive(ctx, dD, dO, n_points, max_n, dI, dM));
  HFB_CUDA(ctx, cudaMemcpyAsync(best_index, dI, (size_t)n_points * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(best_median, dM, (size_t)n_points * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

static int match_host(hfb_ctx* ctx, int mode, const float* A_all, int na_total, const float* B_all, int nb_total,
                      int n_pairs, const int32_t* a_off, const int32_t* a_cnt, const int32_t* b_off,
                      const int32_t* b_cnt, float thr, int32_t* match_idx, float* match_val, int32_t* n_matches) {
  HFB_REQUIRE(ctx, mode == 0 || mode == 1, "mode must be 0 (l2) or 1 (cos)");
  HFB_REQUIRE(ctx, na_total >= 0 && nb_total >= 0 && n_pairs >= 0, "negative size");
  int max_a = 0, max_b = 0;
  for (int p = 0; p < n_pairs; ++p) {
    HFB_REQUIRE(ctx, a_off[p] >= 0 && a_cnt[p] >= 0 && a_off[p] + a_cnt[p] <= na_total && b_off[p] >= 0 &&
                         b_cnt[p] >= 0 && b_off[p] + b_cnt[p] <= nb_total,
                "pair range outside the descriptor arrays");
    max_a = std::max(max_a, a_cnt[p]);
    max_b = std::max(max_b, b_cnt[p]);
  }
  for (int i = 0; i < na_total; ++i) {
    match_idx[i] = -1;
    match_val[i] = 0.f;
  }
  if (n_matches)
    for (int p = 0; p < n_pairs; ++p) n_matches[p] = 0;
  if (na_total == 0 || nb_total == 0 || n_pairs == 0 || max_a == 0 || max_b == 0) return HFB_OK;
  // persistent device staging: A | B | idx | val | tab   (separate from ctx scratch, which launch_match_batch uses)
  StageTimer tm(ctx->trace, "hfb_match");
  const bool same = (A_all == B_all && na_total == nb_total);   // self-matching set (frame vs frame of one batch)
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t szA = al((size_t)na_total * 256 * 4), szB = same ? 0 : al((size_t)nb_total * 256 * 4);
  const size_t szI = al((size_t)na_total * 4), szT = al((size_t)n_pairs * 16);
  HFB_TRY(ctx->ensure_io(szA + szB + 2 * szI + szT));
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->d_io);
  float* dA = reinterpret_cast<float*>(base);
  float* dB = same ? dA : reinterpret_cast<float*>(base + szA);
  int* dI = reinterpret_cast<int*>(base + szA + szB);
  float* dV = reinterpret_cast<float*>(base + szA + szB + szI);
  int* dT = reinterpret_cast<int*>(base + szA + szB + 2 * szI);
  std::vector<int> tab((size_t)4 * n_pairs);
  for (int p = 0; p < n_pairs; ++p) {
    tab[p] = a_off[p];
    tab[n_pairs + p] = a_cnt[p];
    tab[2 * n_pairs + p] = b_off[p];
    tab[3 * n_pairs + p] = b_cnt[p];
  }
  HFB_CUDA(ctx, cudaMemcpyAsync(dA, A_all, (size_t)na_total * 256 * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!same) HFB_CUDA(ctx, cudaMemcpyAsync(dB, B_all, (size_t)nb_total * 256 * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(dT, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_CUDA(ctx, cudaMemsetAsync(dI, 0xFF, (size_t)na_total * 4, ctx->stream));
  HFB_CUDA(ctx, cudaMemsetAsync(dV, 0, (size_t)na_total * 4, ctx->stream));
  tm.mark("h2d");
  int* d_nm = nullptr;
  HFB_TRY(launch_match_batch(ctx, mode, dA, dB, n_pairs, dT, max_a, max_b, thr, dI, dV, na_total, nb_total, &d_nm));
  HFB_CUDA(ctx, cudaMemcpyAsync(match_idx, dI, (size_t)na_total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(match_val, dV, (size_t)na_total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_matches && d_nm)
    HFB_CUDA(ctx, cudaMemcpyAsync(n_matches, d_nm, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // also keeps `tab` alive until its copy has been consumed
  tm.mark("gpu+d2h");
  return HFB_OK;
}

extern "C" int hfb_match_batch(hfb_ctx* ctx, int32_t mode, const float* A_all, int32_t na_total, const float* B_all,
                               int32_t nb_total, int32_t n_pairs, const int32_t* a_off, const int32_t* a_cnt,
                               const int32_t* b_off, const int32_t* b_cnt, float thr, int32_t* match_idx,
                               float* match_val) {
  HFB_ENTER(ctx);
  return match_host(ctx, mode, A_all, na_total, B_all, nb_total, n_pairs, a_off, a_cnt, b_off, b_cnt, thr, match_idx,
                    match_val, nullptr);
}

static int match_single(hfb_ctx* ctx, int mode, const float* A, int na, const float* B, int nb, float thr,
                        int32_t* match_idx, float* match_val, int32_t* n_matches) {
  const int32_t z = 0;
  int32_t nm = 0;
  int rc = match_host(ctx, mode, A, na, B, nb, 1, &z, &na, &z, &nb, thr, match_idx, match_val, &nm);
  if (n_matches) *n_matches = nm;
  return rc;
}
extern "C" int hfb_match_mutual_l2(hfb_ctx* ctx, const float* A, int32_t na, const float* B, int32_t nb,
                                   float max_dist, int32_t* match_idx, float* match_val, int32_t* n_matches) {
  HFB_ENTER(ctx);
  return match_single(ctx, 0, A, na, B, nb, max_dist, match_idx, match_val, n_matches);
}
extern "C" int hfb_match_mutual_cos(hfb_ctx* ctx, const float* A, int32_t na, const float* B, int32_t nb, float min_cos,
                                    int32_t* match_idx, float* match_val, int32_t* n_matches) {
  HFB_ENTER(ctx);
  return match_single(ctx, 1, A, na, B, nb, min_cos, match_idx, match_val, n_matches);
}


// ================================================================================================ keyframe descriptor store
// Local descriptors of keyframes kept RESIDENT in HBM (fp32 rows for the exact re-evaluation + the split-precision image
// the tensor-core contraction reads + half squared norms, all prepared once when the keyframe is stored), so that the
// per-keyframe matching of LocalMapping::CreateNewMapPoints / SearchInNeighbors (current keyframe against <= 30 covisible
// keyframes, src/LocalMapping.cc:513-893, Matcher::SearchForTriangulation src/Matcher.cc:845-889) moves no descriptors:
// only keyframe ids go up and match rows come back.  A store may be filled from one context (Tracking's) and read from
// another one on the same device (LocalMapping's); hfb_kfstore_put synchronises the source stream.
struct hfb_kfstore {
  int device = 0, n_slots = 0, rows_per_slot = 0;
  float* d_rows = nullptr;     // [n_slots][rows_per_slot][256]
  __half* d_img = nullptr;     // [n_slots][rows_per_slot][512] hi | lo
  float* d_hn = nullptr;       // [n_slots][rows_per_slot] 0.5 |d|^2
  float* d_zero = nullptr;     // same shape, zeros (cosine mode)
  std::unordered_map<int64_t, std::pair<int, int>> slot_of;   // id -> (slot, rows)
  std::vector<int> free_slots;
  std::mutex mu;
};

extern "C" int hfb_kfstore_create(hfb_ctx* ctx, int32_t n_slots, int32_t rows_per_slot, hfb_kfstore** out) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, out && n_slots >= 1 && rows_per_slot >= 1 && rows_per_slot <= 65535, "bad store geometry");
  hfb_kfstore* s = new hfb_kfstore();
  s->device = ctx->device;
  s->n_slots = n_slots;
  s->rows_per_slot = rows_per_slot;
  const size_t rows = (size_t)n_slots * rows_per_slot;
  cudaError_t e = cudaMalloc(&s->d_rows, rows * 1024);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_img, rows * 1024);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_hn, rows * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_zero, rows * 4);
  if (e == cudaSuccess) e = cudaMemset(s->d_zero, 0, rows * 4);
  if (e == cudaSuccess) e = cudaMemset(s->d_img, 0, rows * 1024);
  if (e != cudaSuccess) {
    ctx->set_error(std::string("hfb_kfstore_create: ") + cudaGetErrorString(e));
    cudaFree(s->d_rows); cudaFree(s->d_img); cudaFree(s->d_hn); cudaFree(s->d_zero);
    delete s;
    return HFB_ERR_CUDA;
  }
  for (int i = n_slots - 1; i >= 0; --i) s->free_slots.push_back(i);
  *out = s;
  return HFB_OK;
}

extern "C" void hfb_kfstore_destroy(hfb_kfstore* s) {
  if (!s) return;
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  cudaFree(s->d_rows); cudaFree(s->d_img); cudaFree(s->d_hn); cudaFree(s->d_zero);
  delete s;
}

static int kfstore_put_rows(hfb_ctx* ctx, hfb_kfstore* s, int64_t kf_id, const float* src, int n, cudaMemcpyKind kind) {
  HFB_REQUIRE(ctx, s && s->device == ctx->device, "store lives on another device");
  HFB_REQUIRE(ctx, n >= 0 && n <= s->rows_per_slot, "more rows than a store slot holds");
  std::lock_guard<std::mutex> lk(s->mu);
  HFB_REQUIRE(ctx, !s->slot_of.count(kf_id), "keyframe already stored");
  if (s->free_slots.empty()) {
    ctx->set_error("keyframe store is full");
    return HFB_ERR_CAPACITY;
  }
  const int slot = s->free_slots.back();
  const size_t r0 = (size_t)slot * s->rows_per_slot;
  if (n > 0) {
    HFB_CUDA(ctx, cudaMemcpyAsync(s->d_rows + r0 * 256, src, (size_t)n * 1024, kind, ctx->stream));
    HFB_TRY(launch_match_prep(ctx, s->d_rows + r0 * 256, n, s->d_img + r0 * 512, s->d_hn + r0));
  }
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // visible to every context from here on
  s->free_slots.pop_back();
  s->slot_of[kf_id] = std::make_pair(slot, n);
  return HFB_OK;
}

// Store frame `frame_index` of the context's last extraction (its first n keypoints) as keyframe kf_id: device to device.
extern "C" int hfb_kfstore_put_frame(hfb_ctx* ctx, hfb_kfstore* s, int64_t kf_id, int32_t frame_index, int32_t n) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, ctx->last_batch > 0 && frame_index >= 0 && frame_index < ctx->last_batch && n <= ctx->kp_cap,
              "no such resident frame");
  return kfstore_put_rows(ctx, s, kf_id, ctx->d_kdesc + (size_t)frame_index * ctx->kp_cap * HFB_DESC_DIM, n, cudaMemcpyDeviceToDevice);
}
// Store host descriptors (keyframes that predate the store, map loading).
extern "C" int hfb_kfstore_put(hfb_ctx* ctx, hfb_kfstore* s, int64_t kf_id, const float* descriptors, int32_t n) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, n == 0 || descriptors, "null descriptors");
  return kfstore_put_rows(ctx, s, kf_id, descriptors, n, cudaMemcpyHostToDevice);
}
extern "C" int hfb_kfstore_erase(hfb_kfstore* s, int64_t kf_id) {
  if (!s) return HFB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(s->mu);
  auto it = s->slot_of.find(kf_id);
  if (it == s->slot_of.end()) return HFB_OK;
  s->free_slots.push_back(it->second.first);
  s->slot_of.erase(it);
  return HFB_OK;
}
extern "C" int32_t hfb_kfstore_size(hfb_kfstore* s) {
  if (!s) return 0;
  std::lock_guard<std::mutex> lk(s->mu);
  return (int32_t)s->slot_of.size();
}

__global__ void kfstore_replicate_kernel(const float* __restrict__ src, int n, int copies, float* __restrict__ dst) {
  const size_t total = (size_t)n * 64 * copies;   // float4 units
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i % ((size_t)n * 64));
}

// Keyframe kf_a against n_b stored keyframes in ONE launch chain (mode 0 = SearchByBoW flavour, thr = max distance;
// mode 1 = SearchForTriangulation flavour, thr = cosine floor).  match_idx / match_val: [n_b][rows of kf_a], indices into
// the rows of the respective neighbour, -1 = unmatched.
extern "C" int hfb_match_kf_neighbours(hfb_ctx* ctx, hfb_kfstore* s, int64_t kf_a, const int64_t* kf_b, int32_t n_b, int32_t mode,
                                       float thr, int32_t* match_idx, float* match_val, int32_t* rows_a_out) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, s && s->device == ctx->device && kf_b && n_b >= 0 && (mode == 0 || mode == 1), "bad argument");
  int slot_a, na;
  std::vector<int> tab((size_t)4 * std::max(n_b, 1));
  int max_b = 0;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    auto ia = s->slot_of.find(kf_a);
    HFB_REQUIRE(ctx, ia != s->slot_of.end(), "keyframe kf_a is not in the store");
    slot_a = ia->second.first;
    na = ia->second.second;
    for (int p = 0; p < n_b; ++p) {
      auto ib = s->slot_of.find(kf_b[p]);
      HFB_REQUIRE(ctx, ib != s->slot_of.end(), "a neighbour keyframe is not in the store");
      tab[p] = p * na;
      tab[n_b + p] = na;
      tab[2 * n_b + p] = ib->second.first * s->rows_per_slot;
      tab[3 * n_b + p] = ib->second.second;
      max_b = std::max(max_b, ib->second.second);
    }
  }
  if (rows_a_out) *rows_a_out = na;
  if (n_b == 0 || na == 0) return HFB_OK;
  HFB_REQUIRE(ctx, match_idx && match_val, "null output");
  // io block: replicated fp32 rows of kf_a (one copy per pair: the per-row best arrays are indexed like A's rows) | pair
  // table | outputs
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t rows = (size_t)na * n_b;
  const size_t oA = 0, oT = oA + al(rows * 1024), oI = oT + al((size_t)n_b * 16), oV = oI + al(rows * 4), total = oV + al(rows * 4);
  HFB_TRY(ctx->ensure_io(total));
  uint8_t* io = reinterpret_cast<uint8_t*>(ctx->d_io);
  cudaStream_t st = ctx->stream;
  kfstore_replicate_kernel<<<std::min<int>(ctx->n_sm * 4, (int)ceil_div_sz(rows * 64, 256)), 256, 0, st>>>(
      s->d_rows + (size_t)slot_a * s->rows_per_slot * 256, na, n_b, reinterpret_cast<float*>(io + oA));
  HFB_CHECK_LAUNCH(ctx, "kfstore_replicate");
  HFB_CUDA(ctx, cudaMemcpyAsync(io + oT, tab.data(), (size_t)n_b * 16, cudaMemcpyHostToDevice, st));
  HFB_TRY(launch_match_batch(ctx, mode, reinterpret_cast<const float*>(io + oA), s->d_rows, n_b, reinterpret_cast<const int*>(io + oT),
                             na, max_b, thr, reinterpret_cast<int*>(io + oI), reinterpret_cast<float*>(io + oV), (int)rows,
                             s->n_slots * s->rows_per_slot, nullptr, nullptr, 0, 0, s->d_img, mode == 0 ? s->d_hn : s->d_zero));
  HFB_CUDA(ctx, cudaMemcpyAsync(match_idx, io + oI, rows * 4, cudaMemcpyDeviceToHost, st));
  HFB_CUDA(ctx, cudaMemcpyAsync(match_val, io + oV, rows * 4, cudaMemcpyDeviceToHost, st));
  HFB_CUDA(ctx, cudaStreamSynchronize(st));   // also keeps `tab` alive until its copy has been consumed
  return HFB_OK;
}
