// HF-Net encoder forward on sm_100a (replaces the TensorRT engine of src/Extractors/HFNetRTModel.cc:122-137,208-254).
// Graph specification: hfnet/models/hf_net.py:13-104,184-237 (MobileNetV2 x0.75 backbone, local head) and
// hfnet/models/utils/layers.py:57-109 (NetVLAD + dimensionality reduction); BatchNorm folded at weight-load time.
//
// Data layout in HBM: activations NHWC fp16 [B][H][W][C] (C multiple of 8 -> 16-byte channel vectors), weights
// fp16 [Cout][K] K-major (UMMA B operand), fp32 accumulation everywhere, fp32 score / descriptor maps.
// 1x1 convolutions and the 3x3 head convolutions run on the tcgen05 GEMM (gemm_core.cuh); the first 3x3/2
// convolution (1 input channel) and the depthwise 3x3 convolutions are bandwidth-bound CUDA-core kernels.
#include "common.cuh"
#include "gemm_core.cuh"

int gemm_store(hfb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmGeom& g, int B, void* out,
               int ldo, int col_off, const float* bias, const __half* residual, int ldr, int relu6, int f32);
int gemm_l2norm(hfb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmGeom& g, float* out,
                const float* bias);
int gemm_softmax_d2s(hfb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmGeom& g, float* scores,
                     float* logits, const float* bias, int Hc, int Wc);

// TensorFlow 'SAME' padding before the first element (asymmetric on stride 2, SURVEY.md section 7 "hard parts").
static inline int same_pad_before(int n, int k, int s) {
  const int out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  return total / 2;
}

// ----------------------------------------------------------------------------------------------- conv1 (layer_1)
// u8 image -> (x-128)/128 -> 3x3 stride-2 SAME conv to C1 channels + bias + ReLU6 (hf_net.py:185-190,30).
// One thread = one output pixel x 8 channels.
// One thread = one output pixel x all C1 (<= 32) channels: the 9 input bytes are read once, weights come from shared
// memory, the pixel's C1 fp16 outputs leave as contiguous 16-byte stores.
#define CONV1_MAXC 32
__global__ void __launch_bounds__(256) conv1_kernel(const uint8_t* __restrict__ img, int img_h, int img_w, int H8,
                                                    int W8, const float* __restrict__ w,
                                                    const float* __restrict__ bias, int C1,
                                                    __half* __restrict__ out, int Ho, int Wo, int pad_t, int pad_l,
                                                    long long total) {
  __shared__ float s_w[9 * CONV1_MAXC + CONV1_MAXC];
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 10 * C1; i += blockDim.x) s_w[i] = i < 9 * C1 ? __ldg(w + i) : __ldg(bias + i - 9 * C1);
  __syncthreads();
  pdl_wait();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  long long p = gid;
  const int ox = (int)(p % Wo);
  p /= Wo;
  const int oy = (int)(p % Ho);
  const int b = (int)(p / Ho);
  const uint8_t* src = img + (size_t)b * img_h * img_w;
  float px[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - pad_t;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - pad_l;
      const bool ok = iy >= 0 && iy < H8 && ix >= 0 && ix < W8;
      px[ky * 3 + kx] = ok ? ((float)src[(size_t)iy * img_w + ix] - 128.f) * (1.f / 128.f) : 0.f;
    }
  }
  __half* o = out + (((size_t)b * Ho + oy) * Wo + ox) * C1;
  for (int c0 = 0; c0 < C1; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = s_w[9 * C1 + c0 + j];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(px[t], s_w[t * C1 + c0 + j], acc[j]);
    }
    uint4 q;
    __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      hq[j] = __floats2half2_rn(fminf(fmaxf(acc[2 * j], 0.f), 6.f), fminf(fmaxf(acc[2 * j + 1], 0.f), 6.f));
    *reinterpret_cast<uint4*>(o + c0) = q;
  }
}

// ----------------------------------------------------------------------------------------------- depthwise 3x3
// NHWC fp16 -> NHWC fp16, + bias + ReLU6 (conv_blocks.py:272-286).  One thread = one output pixel x 8 channels.
__global__ void dw3x3_kernel(const __half* __restrict__ in, int Hi, int Wi, int C, const float* __restrict__ w,
                             const float* __restrict__ bias, __half* __restrict__ out, int Ho, int Wo, int stride,
                             int pad_t, int pad_l, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int groups = C >> 3;
  const int cg = (int)(gid % groups);
  long long p = gid / groups;
  const int ox = (int)(p % Wo);
  p /= Wo;
  const int oy = (int)(p % Ho);
  const int b = (int)(p / Ho);
  float acc[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8) + 1);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  }
  const __half* src = in + (size_t)b * Hi * Wi * C + cg * 8;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * stride + ky - pad_t;
    if (iy < 0 || iy >= Hi) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * stride + kx - pad_l;
      if (ix < 0 || ix >= Wi) continue;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)iy * Wi + ix) * C));
      const __half2* hq = reinterpret_cast<const __half2*>(&q);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + cg * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (ky * 3 + kx) * C + cg * 8) + 1);
      float2 f;
      f = __half22float2(hq[0]); acc[0] = fmaf(f.x, w0.x, acc[0]); acc[1] = fmaf(f.y, w0.y, acc[1]);
      f = __half22float2(hq[1]); acc[2] = fmaf(f.x, w0.z, acc[2]); acc[3] = fmaf(f.y, w0.w, acc[3]);
      f = __half22float2(hq[2]); acc[4] = fmaf(f.x, w1.x, acc[4]); acc[5] = fmaf(f.y, w1.y, acc[5]);
      f = __half22float2(hq[3]); acc[6] = fmaf(f.x, w1.z, acc[6]); acc[7] = fmaf(f.y, w1.w, acc[7]);
    }
  }
  uint4 q;
  __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    hq[j] = __floats2half2_rn(fminf(fmaxf(acc[2 * j], 0.f), 6.f), fminf(fmaxf(acc[2 * j + 1], 0.f), 6.f));
  *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + oy) * Wo + ox) * C + cg * 8) = q;
}

// ----------------------------------------------------------------------------------------------- layer_2
// Block without an expand convolution (layer_2: depthwise 3x3 on 24 channels -> 1x1 project to 16, hf_net.py:32-34):
// far too little contraction work for a tensor-core tile, so one thread does one output pixel end to end on the CUDA
// cores: depthwise (+bias, ReLU6) for all C channels in registers, then the C x Cout projection (+bias) out of
// shared-memory weights, one contiguous Cout*2-byte store.  HBM traffic = input + output only.
template <int C, int COUT>
__global__ void __launch_bounds__(128) dw_project_small_kernel(const __half* __restrict__ in, int H, int W,
                                                               const float* __restrict__ wd,
                                                               const float* __restrict__ bd,
                                                               const __half* __restrict__ wp,   // [COUT][C] fp16
                                                               const float* __restrict__ bp, __half* __restrict__ out,
                                                               long long total) {
  __shared__ float s_wd[9 * C + C];
  __shared__ float s_wp[C * COUT + COUT];   // [c][o] for conflict-free broadcast reads
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 10 * C; i += blockDim.x) s_wd[i] = i < 9 * C ? __ldg(wd + i) : __ldg(bd + i - 9 * C);
  for (int i = threadIdx.x; i < C * COUT; i += blockDim.x) {
    const int c = i / COUT, o = i - c * COUT;
    s_wp[i] = __half2float(wp[(size_t)o * C + c]);
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) s_wp[C * COUT + i] = __ldg(bp + i);
  __syncthreads();
  pdl_wait();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  long long p = gid;
  const int x = (int)(p % W);
  p /= W;
  const int y = (int)(p % H);
  const int b = (int)(p / H);
  const __half* src = in + (size_t)b * H * W * C;
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = s_wd[9 * C + c];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = y + ky - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = x + kx - 1;
      if (ix < 0 || ix >= W) continue;
      const uint4* q = reinterpret_cast<const uint4*>(src + ((size_t)iy * W + ix) * C);
#pragma unroll
      for (int u = 0; u < C / 8; ++u) {
        const uint4 v = __ldg(q + u);
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(hv[k]);
          acc[u * 8 + 2 * k] = fmaf(f.x, s_wd[(ky * 3 + kx) * C + u * 8 + 2 * k], acc[u * 8 + 2 * k]);
          acc[u * 8 + 2 * k + 1] = fmaf(f.y, s_wd[(ky * 3 + kx) * C + u * 8 + 2 * k + 1], acc[u * 8 + 2 * k + 1]);
        }
      }
    }
  }
  // the depthwise output is rounded to fp16 like the stored activation of the three-kernel path
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = __half2float(__float2half_rn(fminf(fmaxf(acc[c], 0.f), 6.f)));
  float o[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) o[j] = s_wp[C * COUT + j];
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) o[j] = fmaf(acc[c], s_wp[c * COUT + j], o[j]);
  }
  __half* dst = out + (size_t)gid * COUT;
#pragma unroll
  for (int u = 0; u < COUT / 8; ++u) {
    uint4 q;
    __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) hq[k] = __floats2half2_rn(o[u * 8 + 2 * k], o[u * 8 + 2 * k + 1]);
    *reinterpret_cast<uint4*>(dst + u * 8) = q;
  }
}

// =============================================================================================== plans
// N tile: split N into the fewest equal tiles (multiples of 16, <= 256) that still give the machine enough CTAs.
static int pick_bn(int M, int N, int n_sm) {
  const int mt = (M + 127) / 128;
  int bn = 16;
  for (int t = 1; t <= 45; ++t) {
    bn = ((N + t - 1) / t + 15) / 16 * 16;
    if (bn > 256) continue;
    if (mt * t >= n_sm || bn <= 64) break;
  }
  return bn;
}

struct GemmPlan {
  CUtensorMap tmA, tmB;
  GemmGeom g;
};
// fused_cpl.cu: one inverted-residual block as one kernel (channel-per-lane formulation)
struct CplPlan;
CplPlan* cpl_new();
void cpl_delete(CplPlan* p);
int cpl_plan(hfb_ctx* ctx, CplPlan& cp, const BlockW& bw, const __half* in, int Bmax, int Hi, int Wi, int Ho, int Wo,
             int pad_t, int pad_l);
int cpl_run(hfb_ctx* ctx, const CplPlan& cp, const BlockW& bw, const __half* in, __half* out, int B);
int cpl_tiles(const hfb_ctx* ctx, const CplPlan& cp, int stride, int B);
double cpl_bytes(const CplPlan& cp, int B);
double cpl_flops(const CplPlan& cp, int B);
struct BlockPlan {
  GemmPlan expand, project;
  int Hi, Wi, Ho, Wo, pad_t, pad_l;
  CplPlan* cpl = nullptr;       // non-null: the whole block runs as one fused kernel (fused_cpl.cu)
};
struct LevelExec {
  int H1, W1, pad_t1, pad_l1;
  std::vector<BlockPlan> blocks;
  GemmPlan head1, desc2, det2;
  int Hd, Wd;  // descriptor grid = layer_7 size
  int P, D;    // global endpoint pixels / channels
};

// Per-context execution plans.  The table is shared by every context of the process and contexts live on different host
// threads, so look-ups are serialised; std::map nodes do not move, the returned reference stays valid.
static std::vector<LevelExec>& execs(hfb_ctx* ctx) {
  static std::mutex mu;
  static std::map<hfb_ctx*, std::vector<LevelExec>> m;
  std::lock_guard<std::mutex> lk(mu);
  return m[ctx];
}
void encoder_forget(hfb_ctx* ctx) {
  for (LevelExec& le : execs(ctx))
    for (BlockPlan& bp : le.blocks)
    {
      if (bp.cpl) cpl_delete(bp.cpl);
    }
  execs(ctx).clear();
}

static int make_plain(hfb_ctx* ctx, GemmPlan& gp, const void* A, int lda, int a_k_off, long long Mmax, const GemmW& w,
                      int n_sm, int force_bn) {
  const int bn = force_bn > 0 ? force_bn : pick_bn((int)Mmax, w.N, n_sm);
  HFB_TRY(hfb_make_tmap_2d(ctx, &gp.tmA, A, (uint64_t)lda, (uint64_t)Mmax, (uint64_t)lda * 2, 128));
  HFB_TRY(hfb_make_tmap_2d(ctx, &gp.tmB, w.w, (uint64_t)w.Kp, (uint64_t)w.N, (uint64_t)w.Kp * 2, (uint32_t)bn));
  gemm_fill_geom(gp.g, (int)Mmax, w.N, w.K, bn, a_k_off);
  return HFB_OK;
}

// Allocates the activation buffers of every level and builds all tensor maps.  Called once weights are loaded.
int encoder_plan(hfb_ctx* ctx) {
  const NetW& net = ctx->net;
  const int Bm = ctx->cfg.max_batch;
  std::vector<LevelExec>& ex = execs(ctx);
  ex.clear();
  ex.resize(ctx->n_levels);
  for (int l = 0; l < ctx->n_levels; ++l) {
    LevelPlan& lv = ctx->lv[l];
    LevelExec& le = ex[l];
    const int n_layers = lv.global ? 18 : 7;
    lv.n_act = n_layers;
    // layer_1
    le.H1 = (lv.H8 + 1) / 2;
    le.W1 = (lv.W8 + 1) / 2;
    le.pad_t1 = same_pad_before(lv.H8, 3, 2);
    le.pad_l1 = same_pad_before(lv.W8, 3, 2);
    lv.aH[1] = le.H1; lv.aW[1] = le.W1; lv.aC[1] = net.c1;
    HFB_TRY(ctx->dalloc(&lv.act[1], (size_t)Bm * le.H1 * le.W1 * net.c1));
    size_t max_exp = 0, max_dw = 0;
    for (const BlockW& bw : net.blocks) {
      if (bw.layer > n_layers) break;
      const int Hi = lv.aH[bw.layer - 1], Wi = lv.aW[bw.layer - 1];
      const int Ho = (Hi + bw.stride - 1) / bw.stride, Wo = (Wi + bw.stride - 1) / bw.stride;
      lv.aH[bw.layer] = Ho; lv.aW[bw.layer] = Wo; lv.aC[bw.layer] = bw.cout;
      HFB_TRY(ctx->dalloc(&lv.act[bw.layer], (size_t)Bm * Ho * Wo * bw.cout));
      if (bw.has_expand) max_exp = std::max(max_exp, (size_t)Bm * Hi * Wi * bw.cexp);
      max_dw = std::max(max_dw, (size_t)Bm * Ho * Wo * bw.cexp);
    }
    HFB_TRY(ctx->dalloc(&lv.d_exp, max_exp));
    HFB_TRY(ctx->dalloc(&lv.d_dw, max_dw));
    for (const BlockW& bw : net.blocks) {
      if (bw.layer > n_layers) break;
      BlockPlan bp;
      bp.Hi = lv.aH[bw.layer - 1]; bp.Wi = lv.aW[bw.layer - 1];
      bp.Ho = lv.aH[bw.layer]; bp.Wo = lv.aW[bw.layer];
      bp.pad_t = same_pad_before(bp.Hi, 3, bw.stride);
      bp.pad_l = same_pad_before(bp.Wi, 3, bw.stride);
      const long long Min = (long long)Bm * bp.Hi * bp.Wi, Mout = (long long)Bm * bp.Ho * bp.Wo;
      if (bw.has_expand)
        HFB_TRY(make_plain(ctx, bp.expand, lv.act[bw.layer - 1], bw.cin, 0, Min, bw.expand, ctx->n_sm, 0));
      HFB_TRY(make_plain(ctx, bp.project, lv.d_dw, bw.cexp, 0, Mout, bw.project, ctx->n_sm, 0));
      // HFB_CPL_LAYERS=<bit mask of layers> restricts the channel-per-lane kernel to some layers (experiments)
      static const long cpl_mask = getenv("HFB_CPL_LAYERS") ? strtol(getenv("HFB_CPL_LAYERS"), nullptr, 0) : -1L;
      if (ctx->fused_blocks && ((cpl_mask >> bw.layer) & 1)) {
        bp.cpl = cpl_new();
        const int rc = cpl_plan(ctx, *bp.cpl, bw, lv.act[bw.layer - 1], Bm, bp.Hi, bp.Wi, bp.Ho, bp.Wo, bp.pad_t,
                                bp.pad_l);
        if (rc == HFB_ERR_CAPACITY) {
          cpl_delete(bp.cpl);
          bp.cpl = nullptr;
        } else if (rc != HFB_OK) {
          return rc;
        }
      }
      le.blocks.push_back(bp);
    }
    // local head on layer_7
    le.Hd = lv.aH[7]; le.Wd = lv.aW[7];
    const int Cl = lv.aC[7];
    const long long M7 = (long long)Bm * le.Hd * le.Wd;
    HFB_TRY(ctx->dalloc(&lv.d_head1, (size_t)M7 * net.head1.N));
    HFB_TRY(ctx->dalloc(&lv.d_descmap, (size_t)M7 * 256));
    HFB_TRY(ctx->dalloc(&lv.d_scores, (size_t)Bm * lv.H8 * lv.W8));
    HFB_TRY(ctx->dalloc(&lv.d_nms, (size_t)Bm * lv.H8 * lv.W8));
    if (ctx->debug) HFB_TRY(ctx->dalloc(&lv.d_logits, (size_t)M7 * 65));
    HFB_TRY(hfb_make_tmap_nhwc(ctx, &le.head1.tmA, lv.act[7], Cl, le.Wd, le.Hd, Bm));
    HFB_TRY(hfb_make_tmap_2d(ctx, &le.head1.tmB, net.head1.w, (uint64_t)net.head1.Kp, (uint64_t)net.head1.N,
                             (uint64_t)net.head1.Kp * 2, 128));
    gemm_fill_geom_conv(le.head1.g, Bm, le.Hd, le.Wd, Cl, net.head1.N, 128);
    HFB_TRY(make_plain(ctx, le.desc2, lv.d_head1, net.head1.N, 0, M7, net.desc2, ctx->n_sm, 256));
    HFB_TRY(make_plain(ctx, le.det2, lv.d_head1, net.head1.N, 256, M7, net.det2, ctx->n_sm, 80));
    if (le.Hd * 8 != lv.H8 || le.Wd * 8 != lv.W8) {
      ctx->set_error("internal: layer_7 grid is not H8/8 x W8/8");
      return HFB_ERR_STATE;
    }
    // candidates / selection
    HFB_TRY(ctx->dalloc(&lv.d_cand, (size_t)Bm * ctx->cand_cap));
    HFB_TRY(ctx->dalloc(&lv.d_cand_count, (size_t)Bm));
    // global head
    if (lv.global) {
      le.P = lv.aH[18] * lv.aW[18];
      le.D = lv.aC[18];
      const int C = net.n_clusters, K = C * le.D;
      HFB_TRY(ctx->dalloc(&lv.d_vlad_part, (size_t)Bm * global_head_groups(le.P) * (K + C)));
      HFB_TRY(ctx->dalloc(&lv.d_vladn, (size_t)Bm * K));
      HFB_TRY(ctx->dalloc(&lv.d_fc_a, (size_t)((Bm + 7) / 8) * (K / 16) * 32 * 4));
      HFB_TRY(ctx->dalloc(&lv.d_fc_ss, (size_t)Bm * (HFB_GLOBAL_DIM / 32)));
      HFB_TRY(ctx->dalloc(&lv.d_gh_counters, (size_t)Bm + 1));
      HFB_CUDA(ctx, cudaMemset(lv.d_gh_counters, 0, ((size_t)Bm + 1) * sizeof(int)));
      HFB_CUDA(ctx, cudaMemset(lv.d_fc_a, 0, (size_t)((Bm + 7) / 8) * (K / 16) * 32 * 16));
    }
  }
  return HFB_OK;
}

static int run_plain(hfb_ctx* ctx, const GemmPlan& gp, long long M, void* out, int ldo, const float* bias,
                     const __half* residual, int ldr, int relu6) {
  GemmGeom g = gp.g;
  g.M = (int)M;
  return gemm_store(ctx, gp.tmA, gp.tmB, g, 1, out, ldo, 0, bias, residual, ldr, relu6, 0);
}

// NetVLAD + dimensionality reduction (layers.py:57-109) on layer_18 (global_head.cu).
static int global_head(hfb_ctx* ctx, LevelPlan& lv, LevelExec& le, int B) {
  return global_head_run(ctx, lv.act[18], le.P, le.D, B, lv.d_vlad_part, lv.d_gh_counters, lv.d_vladn, lv.d_fc_a,
                         lv.d_fc_ss, ctx->d_global);
}

// stem.cu: layer_1 + layer_2 in one kernel
bool stem_applies(int c1, const BlockW& bw);
int stem_run(hfb_ctx* ctx, const uint8_t* d_img, int img_h, int img_w, int H8, int W8, int H1, int W1, int pad_t,
             int pad_l, const float* w1, const float* b1, const BlockW& bw, __half* l1_out, __half* out, int B);

// Forward of one pyramid level for `B` frames whose u8 images are in lv.d_img.  Produces d_scores, d_nms (no
// selection beyond the threshold scan), d_descmap and (level 0) the global descriptors.
int encoder_forward(hfb_ctx* ctx, int level, int B, float threshold) {
  const NetW& net = ctx->net;
  LevelPlan& lv = ctx->lv[level];
  LevelExec& le = execs(ctx)[level];
  // HFB_STEM=0 keeps layer_1 and layer_2 as two kernels (comparison / debugging)
  const bool stem = ctx->fused_stem && !net.blocks.empty() && stem_applies(net.c1, net.blocks[0]);
  if (stem) {
    const BlockW& b2 = net.blocks[0];
    const double M1 = (double)B * le.H1 * le.W1;
    ctx->note("stem.conv1+l2", (double)B * lv.H8 * lv.W8 + 2.0 * M1 * b2.cout, 2.0 * M1 * (9 * net.c1 + b2.cexp * (9 + b2.cout)));
    HFB_TRY(stem_run(ctx, lv.d_img, lv.H, lv.W, lv.H8, lv.W8, le.H1, le.W1, le.pad_t1, le.pad_l1, net.conv1_w, net.conv1_b,
                     b2, ctx->debug ? lv.act[1] : nullptr, lv.act[2], B));
  } else {
    const long long total = (long long)B * le.H1 * le.W1;
    ctx->note("conv1", (double)B * lv.H8 * lv.W8 + (double)total * net.c1 * 2, 2.0 * 9 * total * net.c1);
    hfb_launch(ctx, conv1_kernel, (unsigned)((total + 255) / 256), 256, 0, 
        lv.d_img, lv.H, lv.W, lv.H8, lv.W8, net.conv1_w, net.conv1_b, net.c1, lv.act[1], le.H1, le.W1, le.pad_t1,
        le.pad_l1, total);
    HFB_CHECK_LAUNCH(ctx, "conv1");
  }
  // The global branch (layer_8 .. layer_18, NetVLAD, FC) and the local branch (heads, NMS) only share layer_7: they
  // run on two streams (fork / join with events, captured into the same graph), so that the small late-layer grids
  // and the local head fill each other's idle SMs.  Profiling runs (one event per launch) stay on one stream.
  cudaStream_t main_stream = ctx->stream;
  struct StreamGuard {   // error returns must not leave the context on the side stream
    hfb_ctx* c;
    cudaStream_t s;
    ~StreamGuard() { c->stream = s; }
  } guard{ctx, main_stream};
  const bool fork = lv.global && ctx->side_stream && ctx->fork_branches && !ctx->prof_on;
  size_t bi = 0;
  for (const BlockW& bw : net.blocks) {
    if (bw.layer > lv.n_act) break;
    if (fork && bw.layer == 8) {
      HFB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
      HFB_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
      ctx->stream = ctx->side_stream;   // every launcher enqueues on ctx->stream
    }
    const BlockPlan& bp = le.blocks[bi++];
    const __half* in = lv.act[bw.layer - 1];
    const __half* dw_in = in;
    const long long Min = (long long)B * bp.Hi * bp.Wi, Mout = (long long)B * bp.Ho * bp.Wo;
    const std::string ln = "l" + std::to_string(bw.layer);
    if (stem && bw.layer == 2) continue;   // done by the stem kernel
    if (!bw.has_expand && bw.stride == 1 && !bw.residual && bw.cexp == 24 && bw.cout == 16) {
      ctx->note(ln + ".dw+project", 2.0 * Min * bw.cin + 2.0 * Mout * bw.cout, 2.0 * Mout * bw.cexp * (9 + bw.cout));
      hfb_launch(ctx, dw_project_small_kernel<24, 16>, (unsigned)((Mout + 127) / 128), 128, 0, 
          in, bp.Hi, bp.Wi, bw.wd, bw.bd, bw.project.w, bw.project.b, lv.act[bw.layer], Mout);
      HFB_CHECK_LAUNCH(ctx, "dw_project_small");
      continue;
    }
    if (bp.cpl && cpl_tiles(ctx, *bp.cpl, bw.stride, B) >= ctx->cpl_min_tiles) {
      ctx->note(ln + ".fused", cpl_bytes(*bp.cpl, B), cpl_flops(*bp.cpl, B));
      HFB_TRY(cpl_run(ctx, *bp.cpl, bw, in, lv.act[bw.layer], B));
      continue;
    }
    // block shapes outside the fused kernel's limits (more than 112 input channels, channel counts that are not
    // multiples of 8) keep the generic path: expand GEMM, depthwise kernel, project GEMM
    if (bw.has_expand) {
      ctx->note(ln + ".expand", 2.0 * Min * (bw.cin + bw.cexp) + 2.0 * bw.cin * bw.cexp, 2.0 * Min * bw.cin * bw.cexp);
      HFB_TRY(run_plain(ctx, bp.expand, Min, lv.d_exp, bw.cexp, bw.expand.b, nullptr, 0, 1));
      dw_in = lv.d_exp;
    }
    const long long total = Mout * (bw.cexp / 8);
    ctx->note(ln + ".dw", 2.0 * (Min + Mout) * bw.cexp, 2.0 * 9 * Mout * bw.cexp);
    hfb_launch(ctx, dw3x3_kernel, (unsigned)((total + 255) / 256), 256, 0, 
        dw_in, bp.Hi, bp.Wi, bw.cexp, bw.wd, bw.bd, lv.d_dw, bp.Ho, bp.Wo, bw.stride, bp.pad_t, bp.pad_l, total);
    HFB_CHECK_LAUNCH(ctx, "dw3x3");
    ctx->note(ln + ".project", 2.0 * Mout * (bw.cexp + bw.cout * (bw.residual ? 2 : 1)) + 2.0 * bw.cexp * bw.cout,
              2.0 * Mout * bw.cexp * bw.cout);
    HFB_TRY(run_plain(ctx, bp.project, Mout, lv.act[bw.layer], bw.cout, bw.project.b,
                      bw.residual ? in : nullptr, bw.cin, 0));
  }
  if (fork) {   // global head on the side stream, then back to the main stream for the local branch
    const int rc = global_head(ctx, lv, le, B);
    ctx->stream = main_stream;
    HFB_TRY(rc);
    HFB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
  }
  // local head (hf_net.py:74-93)
  {
    GemmGeom g = le.head1.g;
    g.M = B * le.Hd * le.Wd;
    const double M7 = (double)g.M, Cl = (double)net.c_local;
    ctx->note("head.conv3x3", 2.0 * M7 * (Cl + 384) + 2.0 * 384 * 9 * Cl, 2.0 * M7 * 9 * Cl * 384);
    HFB_TRY(gemm_store(ctx, le.head1.tmA, le.head1.tmB, g, B, lv.d_head1, net.head1.N, 0, net.head1.b, nullptr, 0, 1, 0));
    GemmGeom gd = le.desc2.g;
    gd.M = g.M;
    ctx->note("head.desc1x1+l2norm", M7 * (256 * 2 + 256 * 4) + 2.0 * 256 * 256, 2.0 * M7 * 256 * 256);
    HFB_TRY(gemm_l2norm(ctx, le.desc2.tmA, le.desc2.tmB, gd, lv.d_descmap, net.desc2.b));
    GemmGeom gt = le.det2.g;
    gt.M = g.M;
    ctx->note("head.det1x1+softmax+d2s", M7 * (128 * 2 + 64 * 4) + 2.0 * 128 * 65, 2.0 * M7 * 128 * 65);
    HFB_TRY(gemm_softmax_d2s(ctx, le.det2.tmA, le.det2.tmB, gt, lv.d_scores, lv.d_logits, net.det2.b, le.Hd, le.Wd));
  }
  ctx->note("nms", 8.0 * B * lv.H8 * lv.W8, 0);
  HFB_TRY(launch_nms(ctx, lv.d_scores, lv.d_nms, lv.H8, lv.W8, B, threshold, lv.d_cand, lv.d_cand_count, ctx->cand_cap));
  if (fork) ctx->join_pending = true;   // joined by the caller once the local branch has been enqueued completely
  else if (lv.global) HFB_TRY(global_head(ctx, lv, le, B));
  return HFB_OK;
}
