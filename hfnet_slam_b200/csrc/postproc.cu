// Network-tail stages that the reference runs inside the TensorRT graph (simple_nms) or on the CPU after the D2H copy
// (threshold scan, top-k, bilinear descriptor resampling, row normalisation) and the image pyramid.  All of it is
// compare / integer / fixed-expression fp32 work: bit-exact against the oracle given identical dense maps.
//
//   simple_nms(radius 4, iterations 2)   hfnet/models/utils/layers.py:10-32, hfnet/export_model.py:35-37
//   threshold scan + top-k               src/Extractors/HFNetRTModel.cc:150-179
//   Resampler + cv::normalize            src/Extractors/BaseModel.cc:491-562, src/Extractors/HFNetRTModel.cc:181-195
//   cv::resize(INTER_LINEAR) on CV_8UC1  src/Extractors/HFextractor.cc:159-173
#include <math.h>

#include "common.cuh"

// =============================================================================================== simple_nms
// simple_nms(radius 4, iterations 2):  M1 = (s == mp(s));  supp = mp(M1) > 0;  s' = supp ? 0 : s;
// out = (M1 | ((s' == mp(s')) & !supp)) ? s : 0, mp = 9x9 max-pool, -inf outside the image.
// One CTA produces a 64 x 32 tile.  The dependency radius is 12 (three chained pools), so the tile is computed from an
// 88 x 56 halo region in shared memory (2.4 halo pixels per output pixel).  The two float pools are separable (row
// pass, column pass); each thread produces a run of 8 outputs from 16 inputs with a suffix-max / prefix-max split
// (22 max ops, 2 smem loads per output instead of 9).  Row pitch 89 (odd) keeps the row pass, whose lanes walk down
// rows, bank-conflict free.  The middle pool runs on a 0/1 mask, so it is done on bit words: the column pass of the
// first pool ballots M1 into 32-pixel words, the 9x9 dilation is shifts + ORs on 3 words per row, and s' is rebuilt on
// the fly from the scores and the supp bits (two float planes instead of three: five CTAs per SM).
#define NMS_TW 64
#define NMS_TH 32
#define NMS_AW (NMS_TW + 24)
#define NMS_AH (NMS_TH + 24)
#define NMS_P 89
#define NMS_WORDS 3   // bit b of word g of a row = region column 4 + 32 g + b

__device__ __forceinline__ void run9(const float (&v)[16], float (&o)[8]) {
  float L[8], R[8];
  L[7] = v[7];
#pragma unroll
  for (int j = 6; j >= 0; --j) L[j] = fmaxf(v[j], L[j + 1]);
  R[0] = v[8];
#pragma unroll
  for (int j = 1; j < 8; ++j) R[j] = fmaxf(v[8 + j], R[j - 1]);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaxf(L[j], R[j]);
}

// With cand != null the kernel also performs the reference's threshold scan (HFNetRTModel.cc:150-168) on its own
// output: every surviving pixel with score >= threshold is appended as a (score, scan index) key.
__global__ void __launch_bounds__(256, 5) nms_kernel(const float* __restrict__ scores, float* __restrict__ out, int H,
                                                  int W, float threshold, u64* __restrict__ cand,
                                                  int* __restrict__ cand_count, int cand_cap) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  float (*sS)[NMS_P] = reinterpret_cast<float (*)[NMS_P]>(nms_smem);   // scores, -inf outside the image
  float (*sT)[NMS_P] = sS + NMS_AH;                                     // row-pass scratch
  uint32_t (*sM1)[NMS_WORDS] = reinterpret_cast<uint32_t (*)[NMS_WORDS]>(sT + NMS_AH);   // M1 bits, rows [4, AH-4)
  uint32_t (*sRow)[NMS_WORDS] = sM1 + NMS_AH;                                               // M1 dilated along the row
  uint32_t (*sSupp)[NMS_WORDS] = sRow + NMS_AH;                                             // supp bits, rows [8, AH-8)
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int x0 = blockIdx.x * NMS_TW, y0 = blockIdx.y * NMS_TH;
  const float* src = scores + (size_t)blockIdx.z * H * W;
  float* dst = out + (size_t)blockIdx.z * H * W;
  const int ax0 = x0 - 12, ay0 = y0 - 12;
  const float NEG = -INFINITY;
  auto inside = [&](int r, int c) {
    const int y = ay0 + r, x = ax0 + c;
    return y >= 0 && y < H && x >= 0 && x < W;
  };
  if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // 16-byte loads: ax0 is a multiple of 4 and so is W, so a quad is entirely inside or outside the image
    for (int i = tid; i < NMS_AH * (NMS_AW / 4); i += 256) {
      const int r = i / (NMS_AW / 4), c = (i - r * (NMS_AW / 4)) * 4;
      float4 v = make_float4(NEG, NEG, NEG, NEG);
      if (inside(r, c)) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(ay0 + r) * W + (ax0 + c)));
      sS[r][c] = v.x; sS[r][c + 1] = v.y; sS[r][c + 2] = v.z; sS[r][c + 3] = v.w;
    }
  } else {
    for (int i = tid; i < NMS_AH * NMS_AW; i += 256) {
      const int r = i / NMS_AW, c = i % NMS_AW;
      sS[r][c] = inside(r, c) ? src[(size_t)(ay0 + r) * W + (ax0 + c)] : NEG;
    }
  }
  __syncthreads();
  // ---- pool 1, row pass: sT[r][c] = max(sS[r][c-4..c+4]) on rows [0, AH) x cols [4, 84)
#pragma unroll 2
  for (int i = tid; i < NMS_AH * 10; i += 256) {
    const int r = i % NMS_AH, c = 4 + (i / NMS_AH) * 8;
    float v[16], o[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = sS[r][c - 4 + j];
    run9(v, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) sT[r][c + j] = o[j];
  }
  __syncthreads();
  // ---- pool 1, column pass + M1 = (s == mp(s)) as ballots: warp task = (8-row block, 32-column group)
  for (int t = warp; t < ((NMS_AH - 8) / 8) * NMS_WORDS; t += 8) {
    const int rb = t / NMS_WORDS, g = t - rb * NMS_WORDS;
    const int r = 4 + rb * 8, c = 4 + 32 * g + lane;
    const bool col_ok = c < NMS_AW - 4;
    float o[8];
    if (col_ok) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = sT[r - 4 + j][c];
      run9(v, o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool mk = col_ok && inside(r + j, c) && sS[r + j][c] == o[j];
      const uint32_t word = __ballot_sync(0xffffffffu, mk);
      if (lane == 0) sM1[r + j][g] = word;
    }
  }
  __syncthreads();
  // ---- pool 2 on bits: supp = 9x9 dilation of M1.  Row dilation of rows [4, AH-4), then column OR on rows [8, AH-8)
  for (int i = tid; i < (NMS_AH - 8) * NMS_WORDS; i += 256) {
    const int r = 4 + i / NMS_WORDS, g = i % NMS_WORDS;
    const uint32_t x = sM1[r][g], l = g > 0 ? sM1[r][g - 1] : 0u, rr = g < NMS_WORDS - 1 ? sM1[r][g + 1] : 0u;
    uint32_t d = x;
#pragma unroll
    for (int k = 1; k <= 4; ++k) d |= (x << k) | (l >> (32 - k)) | (x >> k) | (rr << (32 - k));
    sRow[r][g] = d;
  }
  __syncthreads();
  for (int i = tid; i < (NMS_AH - 16) * NMS_WORDS; i += 256) {
    const int r = 8 + i / NMS_WORDS, g = i % NMS_WORDS;
    uint32_t d = 0u;
#pragma unroll
    for (int dr = -4; dr <= 4; ++dr) d |= sRow[r + dr][g];
    sSupp[r][g] = d;
  }
  __syncthreads();
  // ---- pool 3, row pass on s' = supp ? 0 : s (-inf outside the image): rows [8, AH-8) x cols [12, 76)
#pragma unroll 2
  for (int i = tid; i < (NMS_AH - 16) * 8; i += 256) {
    const int r = 8 + i % (NMS_AH - 16), c = 12 + (i / (NMS_AH - 16)) * 8;
    // supp bits of columns c-4 .. c+11 = bit positions (c - 8) .. (c + 7) of the row's 96-bit string
    const int b0 = c - 8, g = b0 >> 5;
    const uint64_t two = (uint64_t)sSupp[r][g] | ((uint64_t)(g + 1 < NMS_WORDS ? sSupp[r][g + 1] : 0u) << 32);
    const uint32_t sb = (uint32_t)(two >> (b0 & 31));
    float v[16], o[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float sv = sS[r][c - 4 + j];   // -inf outside the image already
      v[j] = ((sb >> j) & 1u) ? (inside(r, c - 4 + j) ? 0.f : NEG) : sv;
    }
    run9(v, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) sT[r][c + j] = o[j];
  }
  __syncthreads();
  // ---- pool 3, column pass + output: warp task = (8-row block of the tile, 32-column half)
  {
    const int rb = warp >> 1, h = warp & 1;
    const int r = 12 + rb * 8, c = 12 + 32 * h + lane;
    float v[16], o[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = sT[r - 4 + j][c];
    run9(v, o);
    const int bit = c - 4, g = bit >> 5, sh = bit & 31;
    const int x = ax0 + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int y = ay0 + r + j;
      if (y >= H || x >= W) continue;
      const float sv = sS[r + j][c];
      const bool supp = (sSupp[r + j][g] >> sh) & 1u;
      const bool m1 = (sM1[r + j][g] >> sh) & 1u;
      const float s2 = supp ? 0.f : sv;
      const bool mx = m1 || (!supp && s2 == o[j]);
      const float val = mx ? sv : 0.f;
      dst[(size_t)y * W + x] = val;
      if (cand && val >= threshold) {
        const int pos = atomicAdd(cand_count + blockIdx.z, 1);
        if (pos < cand_cap)
          cand[(size_t)blockIdx.z * cand_cap + pos] =
              ((u64)f2ord(val) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)(x * H + y));
      }
    }
  }
}

int launch_nms(hfb_ctx* ctx, const float* d_scores, float* d_out, int H, int W, int B, float threshold, u64* d_cand,
               int* d_cand_count, int cand_cap) {
  if (d_cand) HFB_CUDA(ctx, cudaMemsetAsync(d_cand_count, 0, sizeof(int) * B, ctx->stream));
  dim3 grid(ceil_div(W, NMS_TW), ceil_div(H, NMS_TH), B);
  constexpr size_t smem = 2 * sizeof(float) * NMS_AH * NMS_P + 3 * sizeof(uint32_t) * NMS_AH * NMS_WORDS;
  static SmemOptIn optin;
  HFB_CUDA(ctx, optin.ensure(nms_kernel, ctx->device, smem));
  hfb_launch(ctx, nms_kernel, grid, 256, smem, d_scores, d_out, H, W, threshold, d_cand, d_cand_count, cand_cap);
  HFB_CHECK_LAUNCH(ctx, "nms");
  return HFB_OK;
}

// =============================================================================================== threshold scan
// key = (ordered score bits << 32) | (~scan index), scan index = col*H + row (the reference visits column-major,
// HFNetRTModel.cc:155-168): a larger key is a better keypoint, ties resolved towards the earlier visit.
__global__ void select_compact_kernel(const float* __restrict__ nms, int H, int W, float threshold, u64* __restrict__ cand,
                                      int* __restrict__ count, int cap) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const float s = nms[(size_t)b * H * W + i];
  if (s >= threshold) {
    const int row = i / W, col = i - row * W;
    const int pos = atomicAdd(count + b, 1);
    if (pos < cap) cand[(size_t)b * cap + pos] = ((u64)f2ord(s) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)(col * H + row));
  }
}

// =============================================================================================== top-k
// One CTA per frame: radix-select the k-th largest key (8 passes of 8 bits), gather the k survivors (keys are
// unique), bitonic-sort them descending in shared memory.  k <= SELECT_KMAX.
#define SELECT_KMAX 8192

__global__ void __launch_bounds__(1024) select_topk_kernel(const u64* __restrict__ cand, const int* __restrict__ count,
                                                           int cap, int k_req, u64* __restrict__ sel,
                                                           int* __restrict__ n_sel, int sel_stride,
                                                           int* __restrict__ overflow) {
  extern __shared__ u64 s_keys[];  // [P]
  __shared__ int s_hist[256];
  __shared__ u64 s_prefix;
  __shared__ int s_remaining, s_fill, s_done;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  int n = count[b];
  if (n > cap) {
    if (tid == 0) atomicExch(overflow, 1);
    n = cap;
  }
  const u64* c = cand + (size_t)b * cap;
  const int k = min(k_req, n);
  if (tid == 0) n_sel[b] = k;
  if (k == 0) return;
  u64 kth = 0ull;
  if (n > k) {
    if (tid == 0) {
      s_prefix = 0ull;
      s_remaining = k;
      s_done = 0;
    }
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      const u64 prefix = s_prefix;
      const int rem = s_remaining;   // stable until this pass's winner bin rewrites it after the scan
      const u64 himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
      for (int i = tid; i < n; i += 1024) {
        const u64 key = c[i];
        if ((key & himask) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & 0xFF)], 1);
      }
      __syncthreads();
      // find the bin where the count from the top reaches `remaining`: parallel suffix sums over the 256 bins
      // (8 warps x 32 bins) instead of a serial walk
      {
        __shared__ int s_wtot[8];
        int h = 0, incl = 0;
        if (tid < 256) {
          h = s_hist[tid];
          incl = h;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, incl, d);
            if ((tid & 31) + d < 32) incl += v;
          }
          if ((tid & 31) == 0) s_wtot[tid >> 5] = incl;   // total of this warp's 32 bins
        }
        __syncthreads();
        if (tid < 256) {
          int above = 0;
          for (int w = (tid >> 5) + 1; w < 8; ++w) above += s_wtot[w];
          const int suffix_incl = incl + above;            // elements in bins >= tid
          const int suffix_excl = suffix_incl - h;         // elements in bins >  tid
          if (suffix_incl >= rem && suffix_excl < rem) {    // exactly one bin qualifies
            s_prefix = prefix | ((u64)tid << shift);
            s_remaining = rem - suffix_excl;
            // the whole bin is wanted: every key with this prefix is selected, the threshold is the prefix itself (low
            // bits zero) and the remaining passes would only refine inside a bin that is taken completely
            if (suffix_incl == rem) s_done = 1;
          }
        }
      }
      __syncthreads();
      if (s_done) break;
    }
    kth = s_prefix;
  }
  int P = 1;
  while (P < k) P <<= 1;
  if (tid == 0) s_fill = 0;
  for (int i = tid; i < P; i += 1024) s_keys[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < n; i += 1024) {
    const u64 key = c[i];
    if (key >= kth) {
      const int pos = atomicAdd(&s_fill, 1);
      if (pos < P) s_keys[pos] = key;
    }
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P; i += 1024) {
        const int j = i ^ stride;
        if (j > i) {
          const u64 a = s_keys[i], bb = s_keys[j];
          const bool desc = ((i & size) == 0);
          if (desc ? (a < bb) : (a > bb)) {
            s_keys[i] = bb;
            s_keys[j] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += 1024) sel[(size_t)b * sel_stride + i] = s_keys[i];
}

// =============================================================================================== sample
// One warp per keypoint: warp the coordinates to the descriptor grid, bilinear-sample 256 channels with the
// reference's expression order (BaseModel.cc:540-550; every product / sum individually rounded, no FMA contraction),
// then cv::normalize(NORM_L2): norm accumulated in double, scale rounded to fp32, fp32 multiply.
__global__ void sample_kernel(const u64* __restrict__ sel, const int* __restrict__ n_sel_levels, int nsel_stride,
                              int sel_stride, const float* __restrict__ descmap, int H, int W, int Hd, int Wd,
                              float level_scale, int level, int kp_cap, float* __restrict__ ox,
                              float* __restrict__ oy, float* __restrict__ oresp, int* __restrict__ ooct,
                              float* __restrict__ odesc, int* __restrict__ kcount_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int kp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = n_sel_levels[level * nsel_stride + b];
  int base = 0;   // rows of the levels below (HFextractor.cc:259-283 concatenates the levels in order)
  for (int l = 0; l < level; ++l) base += n_sel_levels[l * nsel_stride + b];
  if (kp == 0 && lane == 0) kcount_out[b * HFB_MAX_LEVELS + level] = n;
  if (kp >= n) return;
  const u64 key = sel[(size_t)b * sel_stride + kp];
  const float score = ord2f((unsigned int)(key >> 32));
  const unsigned int scan = 0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull);
  const int col = (int)(scan / (unsigned int)H), row = (int)(scan % (unsigned int)H);
  // HFNetRTModel.cc:147-148,181-188
  const float scale_w = __fdiv_rn(__fsub_rn((float)Wd, 1.f), __fsub_rn((float)W, 1.f));
  const float scale_h = __fdiv_rn(__fsub_rn((float)Hd, 1.f), __fsub_rn((float)H, 1.f));
  const float x = __fmul_rn(scale_w, (float)col), y = __fmul_rn(scale_h, (float)row);
  const size_t o = (size_t)b * kp_cap + base + kp;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const float* dm = descmap + (size_t)b * Hd * Wd * 256;
  if (x > -1.f && y > -1.f && x < (float)Wd && y < (float)Hd) {
    const int fx = (int)floorf(x), fy = (int)floorf(y);
    const int cx = fx + 1, cy = fy + 1;
    const float dx = __fsub_rn((float)cx, x), dy = __fsub_rn((float)cy, y);
    const float w00 = __fmul_rn(dx, dy);
    const float w11 = __fmul_rn(__fsub_rn(1.f, dx), __fsub_rn(1.f, dy));
    const float w01 = __fmul_rn(dx, __fsub_rn(1.f, dy));
    const float w10 = __fmul_rn(__fsub_rn(1.f, dx), dy);
    float f00[8], f11[8], f01[8], f10[8];
    auto fetch = [&](int xx, int yy, float (&f)[8]) {
      if (xx >= 0 && yy >= 0 && xx <= Wd - 1 && yy <= Hd - 1) {
        const float4* p = reinterpret_cast<const float4*>(dm + ((size_t)yy * Wd + xx) * 256) + lane * 2;
        const float4 a = __ldg(p), c4 = __ldg(p + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = c4.x; f[5] = c4.y; f[6] = c4.z; f[7] = c4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
    };
    fetch(fx, fy, f00);
    fetch(cx, cy, f11);
    fetch(fx, cy, f01);
    fetch(cx, fy, f10);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = __fmul_rn(w00, f00[j]), bq = __fmul_rn(w11, f11[j]);
      const float c = __fmul_rn(w01, f01[j]), d = __fmul_rn(w10, f10[j]);
      v[j] = __fadd_rn(__fadd_rn(__fadd_rn(a, bq), c), d);
    }
  }
  double ss = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) ss += (double)v[j] * (double)v[j];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  const double nrm = sqrt(ss);
  const float sc = nrm > 2.220446049250313e-16 ? (float)(1.0 / nrm) : 0.f;
  float4* od = reinterpret_cast<float4*>(odesc + o * 256) + lane * 2;
  od[0] = make_float4(__fmul_rn(v[0], sc), __fmul_rn(v[1], sc), __fmul_rn(v[2], sc), __fmul_rn(v[3], sc));
  od[1] = make_float4(__fmul_rn(v[4], sc), __fmul_rn(v[5], sc), __fmul_rn(v[6], sc), __fmul_rn(v[7], sc));
  if (lane == 0) {
    ox[o] = __fmul_rn((float)col, level_scale);   // HFextractor.cc:272-279: pt *= mvScaleFactor[level]
    oy[o] = __fmul_rn((float)row, level_scale);
    oresp[o] = score;
    ooct[o] = level;
  }
}

int launch_select(hfb_ctx* ctx, const float* d_nms, int H, int W, u64* d_cand, int* d_cand_count, int cand_cap,
                  u64* d_sel, int* d_nsel, int n_keypoints, float threshold, int B, int* d_overflow,
                  bool candidates_ready) {
  HFB_REQUIRE(ctx, n_keypoints >= 0 && n_keypoints <= SELECT_KMAX, "keypoint budget exceeds SELECT_KMAX (8192)");
  if (!candidates_ready) {   // the in-graph NMS kernel already scanned its own output otherwise
    HFB_CUDA(ctx, cudaMemsetAsync(d_cand_count, 0, sizeof(int) * B, ctx->stream));
    dim3 g1(ceil_div(H * W, 256), B);
    select_compact_kernel<<<g1, 256, 0, ctx->stream>>>(d_nms, H, W, threshold, d_cand, d_cand_count, cand_cap);
    HFB_CHECK_LAUNCH(ctx, "select_compact");
  }
  int P = 1;
  while (P < n_keypoints) P <<= 1;
  const size_t smem = (size_t)P * 8;
  static SmemOptIn optin;
  if (smem > 48 * 1024) HFB_CUDA(ctx, optin.ensure(select_topk_kernel, ctx->device, smem));
  hfb_launch(ctx, select_topk_kernel, B, 1024, smem, d_cand, d_cand_count, cand_cap, n_keypoints, d_sel, d_nsel,
                                                     SELECT_KMAX, d_overflow);
  HFB_CHECK_LAUNCH(ctx, "select_topk");
  return HFB_OK;
}

int launch_sample(hfb_ctx* ctx, int H, int W, const float* d_descmap, int Hd, int Wd, const u64* d_sel,
                  const int* d_nsel_levels, int nsel_stride, int n_keypoints, float level_scale, int level, int B,
                  int kp_cap, float* d_x, float* d_y, float* d_resp, int* d_oct, float* d_desc, int* d_kcount) {
  // budget 0: one warp per frame still publishes the count
  dim3 g2(n_keypoints > 0 ? ceil_div(n_keypoints, 8) : 1, B);
  hfb_launch(ctx, sample_kernel, g2, n_keypoints > 0 ? 256 : 32, 0, d_sel, d_nsel_levels, nsel_stride, SELECT_KMAX,
             d_descmap, H, W, Hd, Wd, level_scale, level, kp_cap, d_x, d_y, d_resp, d_oct, d_desc, d_kcount);
  HFB_CHECK_LAUNCH(ctx, "sample");
  return HFB_OK;
}

// single-level form (level 0 of its own selection arrays)
int launch_select_sample(hfb_ctx* ctx, const float* d_nms, int H, int W, const float* d_descmap, int Hd, int Wd,
                         u64* d_cand, int* d_cand_count, int cand_cap, u64* d_sel, int* d_nsel, int n_keypoints,
                         float threshold, float level_scale, int B, int kp_cap, float* d_x, float* d_y,
                         float* d_resp, int* d_oct, float* d_desc, int* d_kcount, int* d_overflow,
                         bool candidates_ready) {
  HFB_TRY(launch_select(ctx, d_nms, H, W, d_cand, d_cand_count, cand_cap, d_sel, d_nsel, n_keypoints, threshold, B,
                        d_overflow, candidates_ready));
  return launch_sample(ctx, H, W, d_descmap, Hd, Wd, d_sel, d_nsel, 0, n_keypoints, level_scale, 0, B, kp_cap, d_x, d_y,
                       d_resp, d_oct, d_desc, d_kcount);
}

// =============================================================================================== pyramid
// cv::resize(INTER_LINEAR) for CV_8UC1: 11-bit fixed-point coefficients, horizontal pass in int32, vertical pass
// (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.  Tables come from build_resize_tables (host).
__global__ void resize_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst, int dh,
                              int dw, const int* __restrict__ xi, const short* __restrict__ xa,
                              const int* __restrict__ yi, const short* __restrict__ ya) {
  pdl_launch_dependents();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw) return;
  const uint8_t* s = src + (size_t)blockIdx.z * sh * sw;
  const int sx = xi[x], sx1 = min(sx + 1, sw - 1);
  const int a0 = xa[2 * x], a1 = xa[2 * x + 1];
  const int sy = yi[y], sy1 = min(sy + 1, sh - 1);
  const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
  const int r0 = (int)s[(size_t)sy * sw + sx] * a0 + (int)s[(size_t)sy * sw + sx1] * a1;
  const int r1 = (int)s[(size_t)sy1 * sw + sx] * a0 + (int)s[(size_t)sy1 * sw + sx1] * a1;
  int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  v = v < 0 ? 0 : (v > 255 ? 255 : v);
  dst[(size_t)blockIdx.z * dh * dw + (size_t)y * dw + x] = (uint8_t)v;
}

void build_resize_tables(int sn, int dn, std::vector<int>& idx, std::vector<short>& coef) {
  idx.resize(dn);
  coef.resize(2 * (size_t)dn);
  const double scale = (double)sn / (double)dn;
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    idx[d] = s;
    coef[2 * d] = (short)lrintf((1.f - f) * 2048.f);      // cvRound: round half to even
    coef[2 * d + 1] = (short)lrintf(f * 2048.f);
  }
}

int launch_resize(hfb_ctx* ctx, const uint8_t* d_src, int sh, int sw, uint8_t* d_dst, int dh, int dw, const int* d_xi,
                  const short* d_xa, const int* d_yi, const short* d_ya, int B) {
  dim3 grid(ceil_div(dw, 128), dh, B);
  hfb_launch(ctx, resize_kernel, grid, 128, 0, d_src, sh, sw, d_dst, dh, dw, d_xi, d_xa, d_yi, d_ya);
  HFB_CHECK_LAUNCH(ctx, "resize");
  return HFB_OK;
}

// ======================================================================================================= undistortion
// Frame::UndistortKeyPoints (src/Frame.cc:760-793) = cv::undistortPoints(pts, pts, K, dist, noArray(), K): normalise with K,
// five fixed-point iterations of the inverse Brown-Conrady model (OpenCV 4 calib3d, criteria COUNT = 5), re-project with
// P = K, round to float.  All in double, every operation rounded on its own (no contraction) in OpenCV's expression
// order, so the result equals cv2.undistortPoints bit for bit (tests/test_undistort_gpu.py).
__global__ void undistort_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ xu,
                                 float* __restrict__ yu, int n, int kp_cap, const int* __restrict__ kcount,
                                 UndistortParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int lim = n;
  if (kcount) {   // frame blockIdx.y of a batch: only its selected keypoints
    lim = 0;
    for (int l = 0; l < HFB_MAX_LEVELS; ++l) lim += kcount[blockIdx.y * HFB_MAX_LEVELS + l];
    lim = min(lim, kp_cap);
  }
  if (i >= lim) return;
  const size_t o = (size_t)blockIdx.y * kp_cap + i;
#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dsub_rn((a), (b))
  double xn = MUL(SUB((double)x[o], p.cx), p.ifx), yn = MUL(SUB((double)y[o], p.cy), p.ify);
  const double x0 = xn, y0 = yn;
  const double* k = p.k;
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    const double r2 = ADD(MUL(xn, xn), MUL(yn, yn));
    const double num = ADD(1.0, MUL(ADD(MUL(ADD(MUL(k[7], r2), k[6]), r2), k[5]), r2));
    const double den = ADD(1.0, MUL(ADD(MUL(ADD(MUL(k[4], r2), k[1]), r2), k[0]), r2));
    const double icdist = __ddiv_rn(num, den);
    if (icdist < 0) {   // OpenCV gives up on the point and keeps the normalised coordinates
      xn = x0;
      yn = y0;
      break;
    }
    const double dx = ADD(ADD(ADD(MUL(MUL(MUL(2.0, k[2]), xn), yn), MUL(k[3], ADD(r2, MUL(MUL(2.0, xn), xn)))),
                              MUL(k[8], r2)), MUL(MUL(k[9], r2), r2));
    const double dy = ADD(ADD(ADD(MUL(k[2], ADD(r2, MUL(MUL(2.0, yn), yn))), MUL(MUL(MUL(2.0, k[3]), xn), yn)),
                              MUL(k[10], r2)), MUL(MUL(k[11], r2), r2));
    xn = MUL(SUB(x0, dx), icdist);
    yn = MUL(SUB(y0, dy), icdist);
  }
  // P = K as the full 3 x 3 product (the zero entries only matter for diverged points: 0 * inf)
  const double xx = ADD(ADD(MUL(p.fx, xn), MUL(0.0, yn)), p.cx), yy = ADD(ADD(MUL(0.0, xn), MUL(p.fy, yn)), p.cy);
  const double ww = __ddiv_rn(1.0, ADD(ADD(MUL(0.0, xn), MUL(0.0, yn)), 1.0));
  xu[o] = __double2float_rn(MUL(xx, ww));
  yu[o] = __double2float_rn(MUL(yy, ww));
#undef MUL
#undef ADD
#undef SUB
}

int launch_undistort(hfb_ctx* ctx, const float* d_x, const float* d_y, float* d_xu, float* d_yu, int n, int B, int kp_cap,
                     const int* d_kcount) {
  const hfb_ctx::Camera& c = ctx->cam;
  UndistortParams p;
  p.fx = c.fx; p.fy = c.fy; p.cx = c.cx; p.cy = c.cy;
  p.ifx = 1.0 / c.fx; p.ify = 1.0 / c.fy;
  for (int i = 0; i < 12; ++i) p.k[i] = c.k[i];
  if (n <= 0) return HFB_OK;
  dim3 grid(ceil_div(n, 128), B);
  hfb_launch(ctx, undistort_kernel, grid, 128, 0, d_x, d_y, d_xu, d_yu, n, kp_cap, d_kcount, p);
  HFB_CHECK_LAUNCH(ctx, "undistort");
  return HFB_OK;
}
