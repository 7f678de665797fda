// Network stem as ONE kernel: u8 image -> (x-128)/128 -> conv 3x3 stride 2 SAME to 24 channels + bias + ReLU6
// (layer_1, hf_net.py:185-190,30)  ->  layer_2 = depthwise 3x3 + bias + ReLU6 -> 1x1 project to 16 + bias
// (block without expand, hf_net.py:32-34, conv_blocks.py:263-312).  The 24-channel layer_1 tensor (4.3 MB per 752x480
// frame, written and read back once by the two-kernel path) never leaves shared memory: HBM traffic is the u8 frame in
// and the 16-channel layer_2 tensor out.
//
// One CTA walks 32 x 16 output tiles (persistent).  Per tile three phases share the 256 threads:
//   A  layer_1 on the 34 x 18 halo tile from a 69 x 37 image patch: an im2col product [612 pixels] x [9 taps] x [24] on
//      warp tensor-core tiles (mma.sync m16n8k16).  Pixels are exact in fp16 ((u8 - 128) / 128); the fp32 weights go in as an
//      fp16 hi + lo pair (two accumulating tiles), so the product keeps ~22 weight bits and fp32 accumulation.
//   B  depthwise 3x3 (pixel, 8-channel unit) out of the layer_1 tile, mixed-precision FMA, 72 taps in registers
//   C  projection 24 -> 16 as [512 pixels] x [24] x [16] on m16n8k16 + m16n8k8 tiles, fp32 accumulate, bias in the accumulator
// The A / C weight fragments live in registers for the whole kernel.  Shared-memory pixels are 48 bytes (24 fp16): both the
// 16-byte unit accesses of phase B and the 4-byte fragment accesses (row stride 12 words, 8 rows x 4 words) are
// conflict-free.  fp16 rounding points are those of conv1_kernel + dw_project_small_kernel (layer_1 and the depthwise
// output rounded to fp16, ReLU6 before rounding).  HFB_STEM_MMA=0 selects the scalar phases A / C, whose operation order is
// that of the two-kernel path (bit-identical to it; also used when layer_1 is materialised for debugging).
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

// warp-level fp16 tensor-core tiles (fp32 accumulate) for the two dense phases; fragment layouts of the PTX ISA:
// A row g = lane / 4 (and g + 8), k pair 2 * (lane % 4) (and + 8); B column g, same k pairs; C rows g / g + 8, columns 2t, 2t + 1
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
// fp32 weight pair -> fp16 (hi pair, lo pair) with hi + lo == w to ~2^-22 relative
__device__ __forceinline__ void split_h2(float w0, float w1, uint32_t& hi, uint32_t& lo) {
  const __half h0 = __float2half_rn(w0), h1 = __float2half_rn(w1);
  const __half2 H = __halves2half2(h0, h1);
  const __half2 L = __floats2half2_rn(w0 - __half2float(h0), w1 - __half2float(h1));
  hi = *reinterpret_cast<const uint32_t*>(&H);
  lo = *reinterpret_cast<const uint32_t*>(&L);
}

struct StemGeom {
  int B, img_h, img_w, H8, W8;      // frame buffer size, cropped size fed to the network
  int H1, W1, pad_t, pad_l;         // layer_1 (= layer_2) size, SAME padding of the stride-2 conv
  int tiles_x, tiles_y, total_tiles;
};

// MODE bit 0: projection (phase C) on mma.sync tiles; bit 1: layer_1 (phase A) as an im2col tile product (K = 9 taps)
template <int MODE>
__global__ void __launch_bounds__(ST_THREADS, (MODE & 4) ? 3 : 2) stem_kernel(const StemGeom g, const uint8_t* __restrict__ img,
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
  const int lane = tid & 31, warp = tid >> 5, fg = lane >> 2, ft = lane & 3;   // mma fragment coordinates
  // phase A fragments: B[k = tap][n = channel] = w1, split hi + lo; taps 9 .. 15 are zero
  uint32_t wa_hi[3][2], wa_lo[3][2];
  float ba[3][2];
  int tap_off[2];                          // patch offsets of taps 2t, 2t + 1 (tap 8 handled by ft == 0)
  auto load_a = [&]() {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int n = j * 8 + fg;
      split_h2(__ldg(w1 + (2 * ft) * ST_C1 + n), __ldg(w1 + (2 * ft + 1) * ST_C1 + n), wa_hi[j][0], wa_lo[j][0]);
      split_h2(ft == 0 ? __ldg(w1 + 8 * ST_C1 + n) : 0.f, 0.f, wa_hi[j][1], wa_lo[j][1]);
      ba[j][0] = __ldg(b1 + j * 8 + 2 * ft);
      ba[j][1] = __ldg(b1 + j * 8 + 2 * ft + 1);
    }
    tap_off[0] = ((2 * ft) / 3) * ST_PW + (2 * ft) % 3;
    tap_off[1] = ((2 * ft + 1) / 3) * ST_PW + (2 * ft + 1) % 3;
  };
  if constexpr ((MODE & 6) == 2) load_a();   // bit 2 (three CTAs per SM, 80 registers): fragments reloaded per tile
  // phase C fragments: B[k = channel][n = output] = wp[n][k], K = 24 as one k16 and one k8 tile
  uint32_t wc[2][3];
  float bc[2][2];
  auto load_c = [&]() {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const __half* row = wp + (size_t)(j * 8 + fg) * wp_ld + 2 * ft;
      wc[j][0] = *reinterpret_cast<const uint32_t*>(row);
      wc[j][1] = *reinterpret_cast<const uint32_t*>(row + 8);
      wc[j][2] = *reinterpret_cast<const uint32_t*>(row + 16);
      bc[j][0] = __ldg(bp + j * 8 + 2 * ft);
      bc[j][1] = __ldg(bp + j * 8 + 2 * ft + 1);
    }
  };
  if constexpr ((MODE & 5) == 1) load_c();
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
    if constexpr (MODE & 2) {
      if constexpr (MODE & 4) load_a();
      constexpr int NPX = ST_CH * ST_CW;
      for (int mt = warp; mt < (NPX + 15) / 16; mt += ST_THREADS / 32) {
        const int p0 = mt * 16 + fg, p1 = p0 + 8;
        const int q0 = min(p0, NPX - 1), q1 = min(p1, NPX - 1);
        const int cy0_ = q0 / ST_CW, cx0_ = q0 - cy0_ * ST_CW;
        const int cy1_ = q1 / ST_CW, cx1_ = q1 - cy1_ * ST_CW;
        const float* s0 = s_patch + cy0_ * 2 * ST_PW + cx0_ * 2;
        const float* s1 = s_patch + cy1_ * 2 * ST_PW + cx1_ * 2;
        const uint32_t a0 = pack_h2(s0[tap_off[0]], s0[tap_off[1]]);
        const uint32_t a1 = pack_h2(s1[tap_off[0]], s1[tap_off[1]]);
        uint32_t a2 = 0, a3 = 0;
        if (ft == 0) {
          a2 = pack_h2(s0[2 * ST_PW + 2], 0.f);
          a3 = pack_h2(s1[2 * ST_PW + 2], 0.f);
        }
        const bool in0 = p0 < NPX && cy0 + cy0_ >= 0 && cy0 + cy0_ < g.H1 && cx0 + cx0_ >= 0 && cx0 + cx0_ < g.W1;
        const bool in1 = p1 < NPX && cy0 + cy1_ >= 0 && cy0 + cy1_ < g.H1 && cx0 + cx1_ >= 0 && cx0 + cx1_ < g.W1;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float c[4] = {ba[j][0], ba[j][1], ba[j][0], ba[j][1]};
          mma16816(c, a0, a1, a2, a3, wa_lo[j][0], wa_lo[j][1]);
          mma16816(c, a0, a1, a2, a3, wa_hi[j][0], wa_hi[j][1]);
          if (p0 < NPX)
            *reinterpret_cast<uint32_t*>(s_l1 + (size_t)p0 * ST_C1 + j * 8 + 2 * ft) = in0 ? relu6_pack(c[0], c[1]) : 0u;
          if (p1 < NPX)
            *reinterpret_cast<uint32_t*>(s_l1 + (size_t)p1 * ST_C1 + j * 8 + 2 * ft) = in1 ? relu6_pack(c[2], c[3]) : 0u;
        }
      }
    } else if (tid < 252) {
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
    if constexpr (MODE & 1) {
      if constexpr (MODE & 4) load_c();
      for (int mt = warp; mt < ST_TH * ST_TW / 16; mt += ST_THREADS / 32) {
        const int p0 = mt * 16 + fg, p1 = p0 + 8;               // half a tile row per m16 tile
        const __half* r0 = s_dw + (size_t)p0 * ST_C1 + 2 * ft;
        const __half* r1 = s_dw + (size_t)p1 * ST_C1 + 2 * ft;
        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(r0), a1 = *reinterpret_cast<const uint32_t*>(r1);
        const uint32_t a2 = *reinterpret_cast<const uint32_t*>(r0 + 8), a3 = *reinterpret_cast<const uint32_t*>(r1 + 8);
        const uint32_t a4 = *reinterpret_cast<const uint32_t*>(r0 + 16), a5 = *reinterpret_cast<const uint32_t*>(r1 + 16);
        const int y = p0 / ST_TW, x0 = p0 - y * ST_TW;
        const int oy = ty * ST_TH + y, ox0 = tx * ST_TW + x0, ox1 = ox0 + 8;
        __half* orow = out + ((size_t)b * g.H1 + oy) * g.W1 * ST_C2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float c[4] = {bc[j][0], bc[j][1], bc[j][0], bc[j][1]};
          mma16816(c, a0, a1, a2, a3, wc[j][0], wc[j][1]);
          mma1688(c, a4, a5, wc[j][2]);
          if (oy < g.H1 && ox0 < g.W1)
            *reinterpret_cast<uint32_t*>(orow + (size_t)ox0 * ST_C2 + j * 8 + 2 * ft) = pack_h2(c[0], c[1]);
          if (oy < g.H1 && ox1 < g.W1)
            *reinterpret_cast<uint32_t*>(orow + (size_t)ox1 * ST_C2 + j * 8 + 2 * ft) = pack_h2(c[2], c[3]);
        }
      }
    } else {
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
  // HFB_STEM_MMA (bit 0: projection, bit 1: layer_1 on warp tensor-core tiles; default both; 7: the same at 80 registers,
  // three CTAs per SM).  The layer_1 debug output is only written by the scalar phase A.
  static const int mode_env = [] {
    const char* e = getenv("HFB_STEM_MMA");
    const int m = e ? (atoi(e) & 7) : 3;
    return (m & 4) ? 7 : m;   // three CTAs per SM exists for the all-tensor-core variant only
  }();
  static SmemOptIn optin[8];
  const int mode = l1_out ? (mode_env & 1) : mode_env;
  const int grid = std::min(g.total_tiles, ctx->n_sm * ((mode & 4) ? 3 : 2));
#define STEM_LAUNCH(M)                                                                                           \
  do {                                                                                                           \
    HFB_CUDA(ctx, optin[M].ensure(stem_kernel<M>, ctx->device, smem));                                           \
    hfb_launch(ctx, stem_kernel<M>, grid, ST_THREADS, smem, g, d_img, w1, b1, bw.wd, bw.bd, bw.project.w,        \
               bw.project.Kp, bw.project.b, l1_out, out);                                                        \
  } while (0)
  switch (mode) {
    case 0: STEM_LAUNCH(0); break;
    case 1: STEM_LAUNCH(1); break;
    case 2: STEM_LAUNCH(2); break;
    case 7: STEM_LAUNCH(7); break;
    default: STEM_LAUNCH(3); break;
  }
#undef STEM_LAUNCH
  HFB_CHECK_LAUNCH(ctx, "stem");
  return HFB_OK;
}
