// Network stem as ONE kernel: u8 image -> (x-128)/128 -> conv 3x3 stride 2 SAME to 24 channels + bias + ReLU6
// (layer_1, hf_net.py:185-190,30)  ->  layer_2 = depthwise 3x3 + bias + ReLU6 -> 1x1 project to 16 + bias
// (block without expand, hf_net.py:32-34, conv_blocks.py:263-312).  The 24-channel layer_1 tensor (4.3 MB per 752x480
// frame, written and read back once by the two-kernel path) never leaves shared memory: HBM traffic is the u8 frame in
// and the 16-channel layer_2 tensor out.
//
// One CTA walks 32 x 16 output tiles (persistent).  Per tile three phases share the 256 threads, each with its own
// thread -> work mapping so that a thread's weights stay in registers for the whole phase:
//   A  (pixel, 8-channel unit): layer_1 on the 34 x 18 halo tile from a 69 x 37 image patch, 72 weights in registers
//   B  (pixel, 8-channel unit): depthwise 3x3 out of the layer_1 tile, 72 weights in registers
//   C  (pixel, 4 outputs):      projection 24 -> 16, 96 weights in registers
// Shared-memory pixels are 48 bytes (24 fp16): lanes walk (pixel, unit) with the unit fastest, i.e. contiguous 16-byte
// pieces, conflict-free.  Arithmetic (operation order, fp16 rounding points) is that of conv1_kernel +
// dw_project_small_kernel, so the result is bit-identical to the two-kernel path (and layer_1 can still be
// materialised for debugging).
#include "common.cuh"

namespace {

constexpr int ST_TW = 32, ST_TH = 16;            // output tile
constexpr int ST_CW = ST_TW + 2, ST_CH = ST_TH + 2;   // layer_1 halo tile
constexpr int ST_PW = 2 * ST_CW + 1, ST_PH = 2 * ST_CH + 1;   // image patch
constexpr int ST_C1 = 24, ST_C2 = 16;
constexpr int ST_THREADS = 256;

// acc + x.half[XH] * w.half[WH] in fp32 (sm_100 mixed-precision FMA; exact product, i.e. == fmaf(float(x), float(w), acc))
template <int XH, int WH>
__device__ __forceinline__ float fmah(uint32_t x, uint32_t w, float acc) {
  float d;
  if constexpr (XH == 0 && WH == 0)
    asm("{\n\t.reg .b16 a, b, c, d;\n\tmov.b32 {a, b}, %1;\n\tmov.b32 {c, d}, %2;\n\tfma.rn.f32.f16 %0, a, c, %3;\n\t}"
        : "=f"(d) : "r"(x), "r"(w), "f"(acc));
  else if constexpr (XH == 0 && WH == 1)
    asm("{\n\t.reg .b16 a, b, c, d;\n\tmov.b32 {a, b}, %1;\n\tmov.b32 {c, d}, %2;\n\tfma.rn.f32.f16 %0, a, d, %3;\n\t}"
        : "=f"(d) : "r"(x), "r"(w), "f"(acc));
  else if constexpr (XH == 1 && WH == 0)
    asm("{\n\t.reg .b16 a, b, c, d;\n\tmov.b32 {a, b}, %1;\n\tmov.b32 {c, d}, %2;\n\tfma.rn.f32.f16 %0, b, c, %3;\n\t}"
        : "=f"(d) : "r"(x), "r"(w), "f"(acc));
  else
    asm("{\n\t.reg .b16 a, b, c, d;\n\tmov.b32 {a, b}, %1;\n\tmov.b32 {c, d}, %2;\n\tfma.rn.f32.f16 %0, b, d, %3;\n\t}"
        : "=f"(d) : "r"(x), "r"(w), "f"(acc));
  return d;
}
// two fp32 -> packed fp16 pair (lo, hi) with ReLU6: max(., 0) folded into the conversion, then min(., 6); rounding is
// monotonic and 0 / 6 are exact in fp16, so this equals rounding the fp32 clamp
__device__ __forceinline__ uint32_t relu6_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  const __half2 six2 = __float2half2_rn(6.f);
  __half2 h = __hmin2(*reinterpret_cast<__half2*>(&r), six2);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct StemGeom {
  int B, img_h, img_w, H8, W8;      // frame buffer size, cropped size fed to the network
  int H1, W1, pad_t, pad_l;         // layer_1 (= layer_2) size, SAME padding of the stride-2 conv
  int tiles_x, tiles_y, total_tiles;
};

__global__ void __launch_bounds__(ST_THREADS, 2) stem_kernel(const StemGeom g, const uint8_t* __restrict__ img,
                                                          const float* __restrict__ w1,    // [9][24]
                                                          const float* __restrict__ b1,    // [24]
                                                          const float* __restrict__ wd,    // [9][24]
                                                          const float* __restrict__ bd,    // [24]
                                                          const __half* __restrict__ wp,   // [16][wp_ld] fp16
                                                          int wp_ld, const float* __restrict__ bp,   // [16]
                                                          __half* __restrict__ l1_out,     // optional layer_1 tensor (debug), or null
                                                          __half* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* s_patch = reinterpret_cast<float*>(smem);                               // [ST_PH][ST_PW] normalised pixels
  __half* s_l1 = reinterpret_cast<__half*>(s_patch + ST_PH * ST_PW + 3);         // [ST_CH * ST_CW][24]
  s_l1 = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(s_l1) + 15) & ~(uintptr_t)15);
  __half* s_dw = s_l1 + ST_CH * ST_CW * ST_C1;                                   // [ST_TH * ST_TW][24]
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  // per-thread weights (weights do not depend on the predecessor kernel)
  const int u = tid % 3;                  // 8-channel unit of phases A / B (threads 252..255 idle there)
  const int og = tid & 3;                 // 4-output group of phase C
  pdl_wait();
  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
    int k = tile;
    const int tx = k % g.tiles_x;
    k /= g.tiles_x;
    const int ty = k % g.tiles_y;
    const int b = k / g.tiles_y;
    const int cy0 = ty * ST_TH - 1, cx0 = tx * ST_TW - 1;             // layer_1 coordinates of the halo tile origin
    const int py0 = cy0 * 2 - g.pad_t, px0 = cx0 * 2 - g.pad_l;       // image coordinates of the patch origin
    const uint8_t* src = img + (size_t)b * g.img_h * g.img_w;
    __syncthreads();   // previous tile's phase C has read s_dw / s_l1
    for (int i = tid; i < ST_PH * ST_PW; i += ST_THREADS) {
      const int r = i / ST_PW, c = i - r * ST_PW;
      const int iy = py0 + r, ix = px0 + c;
      const bool ok = iy >= 0 && iy < g.H8 && ix >= 0 && ix < g.W8;
      s_patch[i] = ok ? ((float)src[(size_t)iy * g.img_w + ix] - 128.f) * (1.f / 128.f) : 0.f;
    }
    __syncthreads();
    // ---- phase A: layer_1 on the halo tile (zero outside the layer_1 map: the depthwise conv pads the ACTIVATION)
    if (tid < 252) {
      float w[9][8], bias[8];
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) w[t][j] = __ldg(w1 + t * ST_C1 + u * 8 + j);
#pragma unroll
      for (int j = 0; j < 8; ++j) bias[j] = __ldg(b1 + u * 8 + j);
      for (int p = tid / 3; p < ST_CH * ST_CW; p += 84) {
        const int cy = p / ST_CW, cx = p - cy * ST_CW;
        const int gy = cy0 + cy, gx = cx0 + cx;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (gy >= 0 && gy < g.H1 && gx >= 0 && gx < g.W1) {
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = bias[j];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float x = s_patch[(cy * 2 + ky) * ST_PW + cx * 2 + kx];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = fmaf(x, w[ky * 3 + kx][j], acc[j]);
            }
          q.x = relu6_pack(acc[0], acc[1]);
          q.y = relu6_pack(acc[2], acc[3]);
          q.z = relu6_pack(acc[4], acc[5]);
          q.w = relu6_pack(acc[6], acc[7]);
          if (l1_out && cy >= 1 && cy <= ST_TH && cx >= 1 && cx <= ST_TW)
            *reinterpret_cast<uint4*>(l1_out + (((size_t)b * g.H1 + gy) * g.W1 + gx) * ST_C1 + u * 8) = q;
        }
        *reinterpret_cast<uint4*>(s_l1 + (size_t)p * ST_C1 + u * 8) = q;
      }
    }
    __syncthreads();
    // ---- phase B: depthwise 3x3 + bias + ReLU6, rounded to fp16 like the stored activation of the two-kernel path
    if (tid < 252) {
      uint32_t w[9][4];   // depthwise taps as fp16 pairs (the loader stores fp16-representable depthwise weights)
      float bias[8];
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[t][j] = pack_h2(__ldg(wd + t * ST_C1 + u * 8 + 2 * j), __ldg(wd + t * ST_C1 + u * 8 + 2 * j + 1));
#pragma unroll
      for (int j = 0; j < 8; ++j) bias[j] = __ldg(bd + u * 8 + j);
      for (int p = tid / 3; p < ST_TH * ST_TW; p += 84) {
        const int y = p / ST_TW, x = p - y * ST_TW;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bias[j];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint4 v = *reinterpret_cast<const uint4*>(s_l1 + (size_t)((y + ky) * ST_CW + x + kx) * ST_C1 + u * 8);
            const uint32_t xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[2 * j] = fmah<0, 0>(xv[j], w[ky * 3 + kx][j], acc[2 * j]);
              acc[2 * j + 1] = fmah<1, 1>(xv[j], w[ky * 3 + kx][j], acc[2 * j + 1]);
            }
          }
        uint4 q;
        q.x = relu6_pack(acc[0], acc[1]);
        q.y = relu6_pack(acc[2], acc[3]);
        q.z = relu6_pack(acc[4], acc[5]);
        q.w = relu6_pack(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(s_dw + (size_t)p * ST_C1 + u * 8) = q;
      }
    }
    __syncthreads();
    // ---- phase C: projection 24 -> 16 (+bias), 4 outputs per thread, channels accumulated in order
    {
      uint32_t w[ST_C1 / 2][4];   // w[c / 2][q] = (wp[og*4 + q][c], wp[og*4 + q][c + 1]) as an fp16 pair
      float bias[4];
#pragma unroll
      for (int c = 0; c < ST_C1 / 2; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          w[c][q] = *reinterpret_cast<const uint32_t*>(wp + (size_t)(og * 4 + q) * wp_ld + 2 * c);
#pragma unroll
      for (int j = 0; j < 4; ++j) bias[j] = __ldg(bp + og * 4 + j);
      for (int p = tid >> 2; p < ST_TH * ST_TW; p += ST_THREADS / 4) {
        const int y = p / ST_TW, x = p - y * ST_TW;
        const int oy = ty * ST_TH + y, ox = tx * ST_TW + x;
        float o[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
        for (int uu = 0; uu < 3; ++uu) {
          const uint4 v = *reinterpret_cast<const uint4*>(s_dw + (size_t)p * ST_C1 + uu * 8);
          const uint32_t xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = fmah<0, 0>(xv[j], w[uu * 4 + j][q], o[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = fmah<1, 1>(xv[j], w[uu * 4 + j][q], o[q]);
          }
        }
        if (oy < g.H1 && ox < g.W1) {
          uint2 q;
          __half2* hq = reinterpret_cast<__half2*>(&q);
          hq[0] = __floats2half2_rn(o[0], o[1]);
          hq[1] = __floats2half2_rn(o[2], o[3]);
          *reinterpret_cast<uint2*>(out + (((size_t)b * g.H1 + oy) * g.W1 + ox) * ST_C2 + og * 4) = q;
        }
      }
    }
  }
}

}  // namespace

bool stem_applies(int c1, const BlockW& bw) {
  return c1 == ST_C1 && !bw.has_expand && bw.stride == 1 && !bw.residual && bw.cin == ST_C1 && bw.cexp == ST_C1 &&
         bw.cout == ST_C2;
}

// layer_1 + layer_2 of one pyramid level for B frames.  `l1_out` (may be null) additionally receives the layer_1 tensor.
int stem_run(hfb_ctx* ctx, const uint8_t* d_img, int img_h, int img_w, int H8, int W8, int H1, int W1, int pad_t,
             int pad_l, const float* w1, const float* b1, const BlockW& bw, __half* l1_out, __half* out, int B) {
  StemGeom g;
  g.B = B; g.img_h = img_h; g.img_w = img_w; g.H8 = H8; g.W8 = W8;
  g.H1 = H1; g.W1 = W1; g.pad_t = pad_t; g.pad_l = pad_l;
  g.tiles_x = (W1 + ST_TW - 1) / ST_TW;
  g.tiles_y = (H1 + ST_TH - 1) / ST_TH;
  g.total_tiles = g.tiles_x * g.tiles_y * B;
  constexpr size_t smem = sizeof(float) * (ST_PH * ST_PW + 3) + 16 + sizeof(__half) * ST_C1 * (ST_CH * ST_CW + ST_TH * ST_TW);
  static bool configured = false;
  if (!configured) {
    HFB_CUDA(ctx, cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int grid = std::min(g.total_tiles, ctx->n_sm * 2);
  hfb_launch(ctx, stem_kernel, grid, ST_THREADS, smem, g, d_img, w1, b1, bw.wd, bw.bd, bw.project.w, bw.project.Kp,
             bw.project.b, l1_out, out);
  HFB_CHECK_LAUNCH(ctx, "stem");
  return HFB_OK;
}
