// One MobileNetV2 inverted-residual block (hfnet/models/backbones/utils/conv_blocks.py:162-312) as ONE kernel:
//   1x1 expand (+bias, ReLU6)  ->  3x3 depthwise stride 1|2, TF-SAME (+bias, ReLU6)  ->  1x1 project (+bias, +residual)
// The 6x-expanded tensor never touches HBM.  Persistent CTAs keep ALL weights of the block resident in shared memory
// (one TMA burst at start) and walk TH x 16 output tiles; per tile
//   (0) the input halo tile (IH x IW pixels, zero outside the image) is loaded with plain 16-byte loads into K-major
//       128B-swizzled rows: the A operand of the expand GEMM (pixels are only 32..240 bytes, too small for TMA rows),
//   then per 64-wide (32 for stride 2) chunk of the expanded channels
//   (1) tcgen05.mma  halo pixels x chunk  (fp32 in TMEM), read back with tcgen05.ld, +bias, ReLU6, zeroed outside the
//       image (SAME padding applies to the EXPANDED activation), stored fp16 in shared memory,
//   (2) depthwise 3x3 on CUDA cores out of shared memory, written as the 128B-swizzled K-major A tile of
//   (3) tcgen05.mma  128 output pixels x Cout, accumulated over the chunks in a second TMEM region,
//   and finally (4) +bias (+residual) -> fp16 NHWC.
// The MMAs are software-pipelined against the CUDA-core phases: expand(j+1) runs under depthwise(j), project(j) under
// the TMEM read-back of chunk j+1.  HBM traffic per block = input (with halo) + output.
#include <algorithm>

#include "common.cuh"
#include "tc.cuh"

struct FusedGeom {
  int B, Hi, Wi, Ho, Wo;
  int Cin, Cexp, Cout;
  int stride, pad_t, pad_l;
  int has_expand, residual;
  int tiles_x, tiles_y, total_tiles;
  int TH;                 // output tile = TH x 16 pixels
  int IH, IW_, R, MT;     // input halo tile, R = IH*IW rows, MT = ceil(R/128) expand M-tiles
  int CW;                 // chunk of expanded channels (multiple of 16, <= 64)
  int n_chunks;
  int cexp_pad;           // n_chunks * CW
  int kb_in;              // 64-channel k-blocks of Cin
  int cout_pad;           // Cout rounded up to 16
  int e_pitch;            // bytes per row of the expanded tile in smem
  int xrb;                // bytes per row of the input tile: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B, Cin <= 32)
  uint32_t tmem_cols;
  uint32_t we_chunk_bytes, wp_chunk_bytes, w_total_bytes;
  uint32_t off_X, off_A2, off_WE, off_WP, off_E, off_wd, off_bars, smem_bytes;
};

// packed (lo, hi) fp16 pair x fp16 pair -> two fp32 accumulators: one FHFMA each (sm_100 mixed-precision FMA; the
// fp16 x fp16 product is exact in fp32, so this is bit-identical to fmaf(float(x), float(w), acc))
__device__ __forceinline__ void fma2_f16(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.f16 %0, xl, wl, %0;\n\tfma.rn.f32.f16 %1, xh, wh, %1;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}
// round two fp32 to a packed fp16 pair (lo, hi) with max(., 0) folded into the conversion, then min(., 6): ReLU6.
// Rounding is monotonic and 0 / 6 are exact in fp16, so this equals rounding the fp32 clamp.
__device__ __forceinline__ uint32_t relu6_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  const __half2 six2 = __float2half2_rn(6.f);
  __half2 h = __hmin2(*reinterpret_cast<__half2*>(&r), six2);
  return *reinterpret_cast<uint32_t*>(&h);
}

// NT = threads per CTA: 256 (two CTAs per SM hide each other's barrier / MMA round-trip stalls) or 512 (one CTA per SM
// for the blocks whose resident weights leave no room for a second CTA).
template <int S, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1) fused_block_kernel(const __grid_constant__ CUtensorMap tmWE,
                                                                 const __grid_constant__ CUtensorMap tmWP,
                                                                 const FusedGeom g, const __half* __restrict__ in,
                                                                 const float* __restrict__ be,   // expand bias [Cexp]
                                                                 const float* __restrict__ wd,   // dw weights [9][Cexp]
                                                                 const float* __restrict__ bd,   // dw bias [Cexp]
                                                                 const float* __restrict__ bp,   // project bias [Cout]
                                                                 __half* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic (keeps the shared-memory address space visible to the compiler: LDS/STS
  // instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int IW = 15 * S + 3;   // halo tile width
  uint8_t* sX = smem + g.off_X;      // [kb_in][MT*128 rows][xrb B] swizzled
  uint8_t* sA2 = smem + g.off_A2;    // [128 rows][128 B] swizzled (written by the depthwise phase)
  uint8_t* sWE = smem + g.off_WE;    // [n_chunks][kb_in][CW rows][128 B] swizzled (TMA, resident)
  uint8_t* sWP = smem + g.off_WP;    // [n_chunks][cout_pad rows][128 B] swizzled (TMA, resident)
  uint8_t* sE = smem + g.off_E;      // [R][e_pitch] expanded activations of the current chunk, fp16
  __half* s_wd = reinterpret_cast<__half*>(smem + g.off_wd);  // [9][cexp_pad] dw weights (fp16-exact values)
  float* s_bd = reinterpret_cast<float*>(smem + g.off_wd + 18 * g.cexp_pad);   // [cexp_pad] dw bias
  float* s_be = s_bd + g.cexp_pad;                                             // [cexp_pad] expand bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bars);
  uint64_t* bar_w = bars;      // all weights landed (once)
  uint64_t* bar_e = bars + 1;  // [2] expand MMAs retired (alternating)
  uint64_t* bar_p = bars + 3;  // [2] project MMAs retired (alternating)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tc::pdl_launch_dependents();

  if (tid == 0) {
    tc::prefetch_tmap(&tmWE);
    tc::prefetch_tmap(&tmWP);
    for (int i = 0; i < 5; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_barrier_init();
    // resident weights: one burst
    tc::mbar_expect_tx(bar_w, g.w_total_bytes);
    for (int j = 0; j < g.n_chunks; ++j) {
      if (g.has_expand)
        for (int kb = 0; kb < g.kb_in; ++kb)
          tc::tma_load_2d(sWE + (size_t)j * g.we_chunk_bytes + (size_t)kb * g.CW * 128, &tmWE, bar_w, kb * 64, j * g.CW);
      tc::tma_load_2d(sWP + (size_t)j * g.wp_chunk_bytes, &tmWP, bar_w, j * g.CW, 0);
    }
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, g.tmem_cols);
  for (int i = tid; i < 11 * g.cexp_pad; i += NT) {
    const int row = i / g.cexp_pad, c = i - row * g.cexp_pad;
    float v = 0.f;
    if (c < g.Cexp) {
      if (row < 9) v = __ldg(wd + (size_t)row * g.Cexp + c);
      else if (row == 9) v = __ldg(bd + c);
      else if (g.has_expand) v = __ldg(be + c);
    }
    if (row < 9) s_wd[i] = __float2half_rn(v);   // exact: the loader stores fp16-representable depthwise weights
    else s_bd[i - 9 * g.cexp_pad] = v;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  tc::pdl_wait();   // everything above touched only weights / on-chip state; the input tensor is the predecessor's output
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base + (uint32_t)(g.MT * g.CW);
  const uint32_t idesc_p = tc::make_idesc_f16(g.cout_pad);

  uint32_t e_cnt = 0, p_cnt = 0;   // expand / project commits issued so far (tracked identically by every thread)
  bool first = true;

  auto issue_expand = [&](int j) {   // thread 0 only; uses commit slot e_cnt
    const int cvalid = min(g.CW, g.Cexp - j * g.CW);
    const uint32_t idesc = tc::make_idesc_f16((cvalid + 15) & ~15);
    for (int mt = 0; mt < g.MT; ++mt) {
      for (int kb = 0; kb < g.kb_in; ++kb) {
        const uint64_t da = g.xrb == 128 ? tc::make_sdesc_sw128(tc::smem_u32(sX + ((size_t)kb * g.MT + mt) * 128 * 128))
                                         : tc::make_sdesc_sw64(tc::smem_u32(sX + (size_t)mt * 128 * 64));
        const uint64_t db =
            tc::make_sdesc_sw128(tc::smem_u32(sWE + (size_t)j * g.we_chunk_bytes + (size_t)kb * g.CW * 128));
        const int krem = g.Cin - kb * 64;
        const int nk = krem >= 64 ? 4 : (krem + 15) >> 4;
        for (int k = 0; k < nk; ++k)
          tc::umma_f16(tmem_base + (uint32_t)(mt * g.CW), tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k),
                       idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
    }
    tc::umma_commit(&bar_e[e_cnt & 1u]);
  };

  // input halo tile -> swizzled K-major rows via cp.async (zero fill outside the image and in the K padding)
  auto load_x = [&](int tile_idx) {
    int q = tile_idx;
    const int ltx = q % g.tiles_x;
    q /= g.tiles_x;
    const int lty = q % g.tiles_y;
    const int limg = q / g.tiles_y;
    const int liy0 = lty * g.TH * S - g.pad_t, lix0 = ltx * 16 * S - g.pad_l;
    const int units = ((g.Cin + 15) & ~15) >> 3;   // 16-byte units per pixel incl. K padding
    const int vunits = g.Cin >> 3;                 // units holding real channels
    const __half* src = in + (size_t)limg * g.Hi * g.Wi * g.Cin;
    for (int r = tid; r < g.R; r += NT) {
      const int ry = r / IW, rx = r - ry * IW;
      const int iy = liy0 + ry, ix = lix0 + rx;
      const bool inb = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
      const __half* gp = inb ? src + ((size_t)iy * g.Wi + ix) * g.Cin : in;
      const uint32_t row_dst = tc::smem_u32(sX) + (uint32_t)r * (uint32_t)g.xrb;
      for (int u = 0; u < units; ++u) {
        const bool ok = inb && u < vunits;
        // 16-byte chunk u of row r under the operand swizzle: SWIZZLE_128B = chunk ^ (row & 7) in 128-byte rows,
        // SWIZZLE_64B = chunk ^ ((row >> 1) & 3) in 64-byte rows (address bits [4,6) ^= bits [7,9))
        const uint32_t dst = g.xrb == 128 ? row_dst + (uint32_t)(u >> 3) * (uint32_t)(g.MT * 128 * 128) +
                                                (uint32_t)(((u & 7) ^ (r & 7)) << 4)
                                          : row_dst + (uint32_t)((u ^ ((r >> 1) & 3)) << 4);
        const int nbytes = ok ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? gp + u * 8 : in), "r"(nbytes)
                     : "memory");
      }
    }
  };
  if ((int)blockIdx.x < g.total_tiles) load_x(blockIdx.x);

  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int img = t / g.tiles_y;
    const int oy0 = ty * g.TH, ox0 = tx * 16;
    const int iy0 = oy0 * S - g.pad_t, ix0 = ox0 * S - g.pad_l;

    // (0) input halo tile: already in flight (cp.async issued during the previous tile, or just above for the first)
    asm volatile("cp.async.wait_all;" ::: "memory");
    tc::fence_proxy_async();
    if (first) {
      tc::mbar_wait(bar_w, 0);
      first = false;
    }
    __syncthreads();   // S0
    if (g.has_expand) {
      if (tid == 0) {
        tc::fence_after_sync();
        issue_expand(0);
      }
      ++e_cnt;
    }

    for (int j = 0; j < g.n_chunks; ++j) {
      const int c0 = j * g.CW;                                   // first expanded channel of the chunk
      const int cvalid = min(g.CW, g.Cexp - c0);                 // real channels in the chunk (multiple of 8)
      const int cw16 = (cvalid + 15) & ~15;                      // K of the project step / N of the expand step
      if (g.has_expand) {
        // (1) wait for expand(j) (commit number e_cnt-1), TMEM -> +bias, ReLU6, zero outside the image -> fp16 rows
        const uint32_t k = e_cnt - 1;
        tc::mbar_wait(&bar_e[k & 1u], (k >> 1) & 1u);
        __syncwarp();
        tc::fence_after_sync();
        // work units = (M-tile, column half) spread over the four warp quads; warp w reads TMEM lane group w % 4
        const int nch = cw16 >> 4, ch_half = (nch + 1) >> 1;
        for (int wu = warp >> 2; wu < g.MT * 2; wu += NT / 128) {
          const int mt = wu >> 1, half = wu & 1;
          const int cbeg = half ? ch_half * 16 : 0, cend = half ? cw16 : ch_half * 16;
          const int r = mt * 128 + (warp & 3) * 32 + lane;
          const int iy = iy0 + r / IW, ix = ix0 + r % IW;
          const bool in_img = r < g.R && iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
          const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * g.CW);
          for (int cc = cbeg; cc < cend; cc += 16) {
            uint32_t v[16];
            tc::tmem_ld16(taddr + (uint32_t)cc, v);
            tc::tmem_ld_wait();
            if (r < g.R) {
              uint4 q[2];
              uint32_t* hq = reinterpret_cast<uint32_t*>(q);
              if (in_img) {
                // bias add in fp32, one rounding to fp16 with the ReLU6 clamp folded into the conversion
                const float4* bq = reinterpret_cast<const float4*>(s_be + c0 + cc);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 b4 = bq[i];
                  hq[2 * i] = relu6_pack(__uint_as_float(v[4 * i]) + b4.x, __uint_as_float(v[4 * i + 1]) + b4.y);
                  hq[2 * i + 1] = relu6_pack(__uint_as_float(v[4 * i + 2]) + b4.z, __uint_as_float(v[4 * i + 3]) + b4.w);
                }
              } else {
                q[0] = make_uint4(0, 0, 0, 0);
                q[1] = make_uint4(0, 0, 0, 0);
              }
              uint4* d = reinterpret_cast<uint4*>(sE + (size_t)r * g.e_pitch + (size_t)cc * 2);
              d[0] = q[0];
              d[1] = q[1];
            }
          }
        }
        tc::fence_before_sync();
      } else {
        // no expand conv (layer_2): the "expanded" activation is the input tile itself
        const int units = cw16 >> 3;
        for (int i = tid; i < g.R * units; i += NT) {
          const int r = i / units, u = i - r * units;
          const int cu = (c0 >> 3) + u;   // 16-byte unit inside the 128-byte swizzled row
          uint4 q = make_uint4(0, 0, 0, 0);
          if (cu * 8 < g.Cin) q = *reinterpret_cast<const uint4*>(sX + (size_t)r * 128 + (size_t)((cu ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(sE + (size_t)r * g.e_pitch + (size_t)u * 16) = q;
        }
      }
      __syncthreads();   // S1: sE complete, D1 drained
      // every expand of this tile has retired and (layer_2) the copy out of sX is done: sX is free, so the next
      // tile's halo load flies under the remaining depthwise / project / store work
      if (j + 1 == g.n_chunks && tile + (int)gridDim.x < g.total_tiles) load_x(tile + gridDim.x);

      // expand(j+1) runs on the tensor core while the CUDA cores do the depthwise of chunk j
      if (g.has_expand && j + 1 < g.n_chunks) {
        if (tid == 0) {
          tc::fence_after_sync();
          issue_expand(j + 1);
        }
        ++e_cnt;
      }
      // project(j-1) (or the previous tile's last one) must have retired before A2 is overwritten
      if (p_cnt > 0) {
        const uint32_t k = p_cnt - 1;
        tc::mbar_wait(&bar_p[k & 1u], (k >> 1) & 1u);
      }

      // (2) depthwise 3x3 (+bias, ReLU6) -> swizzled A tile of the project GEMM.  Thread = (8-channel unit, pixel
      // pair p, p+64): the unit's 72 weights + 8 biases sit in registers for both pixels.
      {
        const int units = cw16 >> 3;
        // threads of one unit (8 channels) = NT / (units rounded up to 2, 4 or 8); each walks the tile's pixels with the
        // unit's 72 weights (36 packed registers) + 8 biases resident
        const int upow = units > 4 ? 8 : (units > 2 ? 4 : 2);
        const int pstep = NT / upow;
        const int u = tid / pstep, pb = tid - u * pstep;
        if (u < units) {
          uint4 w[9];
          float bias8[8];
          {
            const float4* bq = reinterpret_cast<const float4*>(s_bd + c0 + u * 8);
            const float4 b0 = bq[0], b1 = bq[1];
            bias8[0] = b0.x; bias8[1] = b0.y; bias8[2] = b0.z; bias8[3] = b0.w;
            bias8[4] = b1.x; bias8[5] = b1.y; bias8[6] = b1.z; bias8[7] = b1.w;
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) w[tp] = *reinterpret_cast<const uint4*>(s_wd + tp * g.cexp_pad + c0 + u * 8);
          }
          const int npix = g.TH * 16;
          for (int p = pb; p < npix; p += pstep) {
            const int oy = p >> 4, ox = p & 15;
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = bias8[c];
            const uint8_t* e0 = sE + (size_t)((oy * S) * IW + ox * S) * g.e_pitch + (size_t)u * 16;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const uint4 q = *reinterpret_cast<const uint4*>(e0 + (size_t)(ky * IW + kx) * g.e_pitch);
                const uint4 wq = w[ky * 3 + kx];
                fma2_f16(acc[0], acc[1], q.x, wq.x);
                fma2_f16(acc[2], acc[3], q.y, wq.y);
                fma2_f16(acc[4], acc[5], q.z, wq.z);
                fma2_f16(acc[6], acc[7], q.w, wq.w);
              }
            }
            uint4 o;   // channels beyond Cexp (zero weights, zero bias) come out as exact zeros
            o.x = relu6_pack(acc[0], acc[1]);
            o.y = relu6_pack(acc[2], acc[3]);
            o.z = relu6_pack(acc[4], acc[5]);
            o.w = relu6_pack(acc[6], acc[7]);
            *reinterpret_cast<uint4*>(sA2 + (size_t)p * 128 + (size_t)((u ^ (p & 7)) << 4)) = o;
          }
        }
      }
      tc::fence_proxy_async();
      __syncthreads();   // S2: A2 complete, sE free

      // (3) project: D2 (+)= A2 * WP_j^T, runs under the next chunk's read-back
      if (tid == 0) {
        tc::fence_after_sync();
        const uint64_t da = tc::make_sdesc_sw128(tc::smem_u32(sA2));
        const uint64_t db = tc::make_sdesc_sw128(tc::smem_u32(sWP + (size_t)j * g.wp_chunk_bytes));
        for (int k = 0; k < (cw16 >> 4); ++k)
          tc::umma_f16(tmem_d2, tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k), idesc_p,
                       (j > 0 || k > 0) ? 1u : 0u);
        tc::umma_commit(&bar_p[p_cnt & 1u]);
      }
      ++p_cnt;
    }
    // (4) epilogue: wait for the last project, +bias (+residual) -> fp16 NHWC; the four warp quads split the channels
    {
      const uint32_t k = p_cnt - 1;
      tc::mbar_wait(&bar_p[k & 1u], (k >> 1) & 1u);
      __syncwarp();
      tc::fence_after_sync();
      const int p = (warp & 3) * 32 + lane;
      const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
      const bool valid = (p >> 4) < g.TH && oy < g.Ho && ox < g.Wo;
      const long long opix = ((long long)img * g.Ho + oy) * g.Wo + ox;
      const uint32_t taddr = tmem_d2 + ((uint32_t)((warp & 3) * 32) << 16);
      for (int cc = (warp >> 2) * 16; cc < g.cout_pad; cc += NT / 8) {   // 16-column pieces round-robin over the quads
        uint32_t v[16];
        tc::tmem_ld16(taddr + (uint32_t)cc, v);
        tc::tmem_ld_wait();
        if (!valid) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = cc + 8 * h;
          if (n >= g.Cout) break;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * h + i]) + __ldg(bp + n + i);
          if (g.residual) {
            const uint4 q = *reinterpret_cast<const uint4*>(in + opix * g.Cin + n);
            const __half2* hq = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 r2 = __half22float2(hq[i]);
              f[2 * i] += r2.x;
              f[2 * i + 1] += r2.y;
            }
          }
          uint4 o;
          __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int i = 0; i < 4; ++i) ho[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
          *reinterpret_cast<uint4*>(out + opix * g.Cout + n) = o;
        }
      }
      tc::fence_before_sync();   // D2 reads are ordered before the next tile's first project (issued after S0..S2)
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, g.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host side
int hfb_make_tmap_2d(hfb_ctx* ctx, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t row_stride_bytes, uint32_t box_outer);

struct FusedPlan {
  CUtensorMap tmWE, tmWP;
  FusedGeom g;
  int ctas_per_sm;
  int nt;   // threads per CTA (256 or 512)
};

FusedPlan* fused_block_new() { return new FusedPlan(); }
void fused_block_delete(FusedPlan* p) { delete p; }

// Returns HFB_ERR_CAPACITY when the block does not fit on chip (the caller keeps the three-kernel path).
// Configuration search, in order of preference: 256-thread CTAs, two per SM (shared memory <= 112 KB and <= 256 TMEM
// columns each), then one 512-thread CTA per SM; 8 x 16 output tiles unless the halo tile does not fit or the taller
// tiles cannot give every resident CTA two tiles (small late layers), then 4 x 16.
// HFB_FUSED_NT / HFB_FUSED_TH (environment) pin the choice for experiments.
int fused_block_plan(hfb_ctx* ctx, FusedPlan& fp, const BlockW& bw, const __half* in, int Bmax, int Hi, int Wi, int Ho,
                     int Wo, int pad_t, int pad_l) {
  (void)in;
  FusedGeom& g = fp.g;
  g.B = Bmax; g.Hi = Hi; g.Wi = Wi; g.Ho = Ho; g.Wo = Wo;
  g.Cin = bw.cin; g.Cexp = bw.cexp; g.Cout = bw.cout;
  g.stride = bw.stride; g.pad_t = pad_t; g.pad_l = pad_l;
  g.has_expand = bw.has_expand ? 1 : 0;
  g.residual = bw.residual ? 1 : 0;
  g.tiles_x = (Wo + 15) / 16;
  g.CW = bw.stride == 1 ? 64 : 32;
  if (g.CW > ((bw.cexp + 15) & ~15)) g.CW = (bw.cexp + 15) & ~15;
  g.n_chunks = (bw.cexp + g.CW - 1) / g.CW;
  g.cexp_pad = g.n_chunks * g.CW;
  g.kb_in = (bw.cin + 63) / 64;
  g.xrb = (bw.has_expand && bw.cin <= 32) ? 64 : 128;
  g.cout_pad = (bw.cout + 15) & ~15;
  g.e_pitch = g.CW * 2 + 16;
  g.we_chunk_bytes = g.has_expand ? (uint32_t)(g.kb_in * g.CW * 128) : 0u;
  g.wp_chunk_bytes = (uint32_t)(g.cout_pad * 128);
  g.w_total_bytes = (uint32_t)g.n_chunks * (g.we_chunk_bytes + g.wp_chunk_bytes);
  if (g.w_total_bytes >= (1u << 20)) return HFB_ERR_CAPACITY;   // mbarrier tx-count limit
  auto al = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  auto layout = [&](int th) {
    g.TH = th;
    g.tiles_y = (Ho + g.TH - 1) / g.TH;
    g.IH = (g.TH - 1) * bw.stride + 3;
    g.IW_ = 15 * bw.stride + 3;
    g.R = g.IH * g.IW_;
    g.MT = (g.R + 127) / 128;
    uint32_t off = 0;
    g.off_X = off;  off += al((uint32_t)(g.kb_in * g.MT * 128 * g.xrb));
    g.off_A2 = off; off += 128 * 128;
    g.off_WE = off; off += al((uint32_t)g.n_chunks * g.we_chunk_bytes);
    g.off_WP = off; off += al((uint32_t)g.n_chunks * g.wp_chunk_bytes);
    g.off_E = off;  off += al((uint32_t)(g.R * g.e_pitch));
    g.off_wd = off; off += al((uint32_t)(26 * g.cexp_pad));
    g.off_bars = off; off += 64;
    g.smem_bytes = off + 1024;
    uint32_t cols = 32;
    while ((int)cols < g.MT * g.CW + g.cout_pad) cols <<= 1;
    g.tmem_cols = cols;
  };
  const char* e_nt = getenv("HFB_FUSED_NT");
  const char* e_th = getenv("HFB_FUSED_TH");
  const int pin_nt = e_nt ? atoi(e_nt) : 0, pin_th = e_th ? atoi(e_th) : 0;
  bool found = false;
  for (int nt : {256, 512}) {
    if (pin_nt && nt != pin_nt) continue;
    const int ctas = nt == 256 ? 2 : 1;
    const uint32_t budget = nt == 256 ? 112u * 1024u : 200u * 1024u;
    const uint32_t max_cols = nt == 256 ? 256u : 512u;
    const int th0 = pin_th ? pin_th : ((g.tiles_x * ((Ho + 7) / 8) * Bmax < 2 * ctas * ctx->n_sm) ? 4 : 8);
    for (int th = th0; th >= 4 && !found; th >>= 1) {
      layout(th);
      if (g.smem_bytes <= budget && g.tmem_cols <= max_cols) {
        found = true;
        fp.nt = nt;
        fp.ctas_per_sm = ctas;
      }
      if (pin_th) break;
    }
    if (found) break;
  }
  if (!found) return HFB_ERR_CAPACITY;
  g.total_tiles = g.tiles_x * g.tiles_y * Bmax;
  if (bw.has_expand)
    HFB_TRY(hfb_make_tmap_2d(ctx, &fp.tmWE, bw.expand.w, (uint64_t)bw.expand.Kp, (uint64_t)bw.expand.N,
                             (uint64_t)bw.expand.Kp * 2, (uint32_t)g.CW));
  HFB_TRY(hfb_make_tmap_2d(ctx, &fp.tmWP, bw.project.w, (uint64_t)bw.project.Kp, (uint64_t)bw.project.N,
                           (uint64_t)bw.project.Kp * 2, (uint32_t)g.cout_pad));
  if (!bw.has_expand) fp.tmWE = fp.tmWP;   // never dereferenced by the kernel
  if (ctx->trace)
    fprintf(stderr, "hfnet_b200: fused layer_%d: NT=%d x%d TH=%d MT=%d CW=%d xrb=%d smem=%u tmem=%u tiles=%d\n", bw.layer,
            fp.nt, fp.ctas_per_sm, g.TH, g.MT, g.CW, g.xrb, g.smem_bytes, g.tmem_cols, g.total_tiles);
  return HFB_OK;
}

int fused_block_tiles(const FusedPlan& fp, int B) { return fp.g.tiles_x * fp.g.tiles_y * B; }

template <int S, int NT>
static int fused_launch(hfb_ctx* ctx, const FusedPlan& fp, const FusedGeom& g, const BlockW& bw, const __half* in,
                        __half* out, int grid) {
  static size_t configured = 0;   // per instantiation
  if (g.smem_bytes > configured) {
    HFB_CUDA(ctx, cudaFuncSetAttribute(fused_block_kernel<S, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)g.smem_bytes));
    configured = g.smem_bytes;
  }
  hfb_launch(ctx, fused_block_kernel<S, NT>, grid, NT, g.smem_bytes, fp.tmWE, fp.tmWP, g, in, bw.expand.b, bw.wd, bw.bd,
             bw.project.b, out);
  HFB_CHECK_LAUNCH(ctx, "fused_block");
  return HFB_OK;
}

int fused_block_run(hfb_ctx* ctx, const FusedPlan& fp, const BlockW& bw, const __half* in, __half* out, int B) {
  FusedGeom g = fp.g;
  g.B = B;
  g.total_tiles = g.tiles_x * g.tiles_y * B;
  const int grid = std::min(g.total_tiles, ctx->n_sm * fp.ctas_per_sm);
  if (g.stride == 1) return fp.nt == 256 ? fused_launch<1, 256>(ctx, fp, g, bw, in, out, grid)
                                         : fused_launch<1, 512>(ctx, fp, g, bw, in, out, grid);
  return fp.nt == 256 ? fused_launch<2, 256>(ctx, fp, g, bw, in, out, grid)
                      : fused_launch<2, 512>(ctx, fp, g, bw, in, out, grid);
}

double fused_block_bytes(const FusedPlan& fp, int B) {   // algorithmic: input once + output once + weights
  const FusedGeom& g = fp.g;
  return 2.0 * B * ((double)g.Hi * g.Wi * g.Cin + (double)g.Ho * g.Wo * g.Cout * (g.residual ? 2 : 1)) +
         2.0 * ((double)g.Cin * g.Cexp * g.has_expand + (double)g.Cexp * g.Cout) + 4.0 * 10 * g.Cexp;
}
double fused_block_flops(const FusedPlan& fp, int B) {
  const FusedGeom& g = fp.g;
  return 2.0 * B * ((double)g.Hi * g.Wi * g.Cin * g.Cexp * g.has_expand + (double)g.Ho * g.Wo * g.Cexp * (9 + g.Cout));
}
