// One MobileNetV2 inverted-residual block (hfnet/models/backbones/utils/conv_blocks.py:162-312) as ONE kernel:
//   1x1 expand (+bias, ReLU6)  ->  3x3 depthwise stride 1|2, TF-SAME (+bias, ReLU6)  ->  1x1 project (+bias, +residual)
// The 6x-expanded tensor never touches HBM.  Persistent CTAs walk TH x 16 output tiles; the expanded channels are
// processed in chunks of CW (64, or 32 for stride 2), and the per-chunk stages of ALL tiles of a CTA form one software
// pipeline (chunk counter c runs across tile boundaries):
//   expand(c)    tcgen05.mma  halo pixels x chunk  -> TMEM D1[c % nd]
//   readback(c)  tcgen05.ld D1[c % nd], +bias, ReLU6, zero outside the image (SAME padding applies to the EXPANDED
//                activation) -> fp16 rows in shared memory E[c & 1]
//   dw(c)        depthwise 3x3 on CUDA cores out of E[c & 1] (mixed-precision FHFMA, weights in registers)
//                -> 128B-swizzled K-major A tile A2[c & 1]
//   project(c)   tcgen05.mma  128 output pixels x Cout, accumulated over the tile's chunks in TMEM D2
//   epilogue     after the tile's last chunk: +bias (+residual) -> fp16 NHWC
// Phase q (one __syncthreads each) runs dw(q), readback(q+1) and the epilogue of the tile that ended at chunk q-1 on
// all compute threads; after the barrier the issuer warp issues project(q) and every expand whose TMEM buffer has been
// drained and whose input tile has landed (up to chunk q+3 with two D1 buffers).  Every MMA therefore has at least a
// full phase of CUDA-core work to hide behind, and no stage waits on a tensor-core round trip.
// Input halo tiles (IH x IW pixels, zero outside the image) are loaded with 16-byte cp.async into K-major swizzled
// rows (pixels are only 32..240 bytes, too small for TMA rows) one tile ahead (two buffers) when shared memory allows.
// Weights are resident in shared memory (one TMA burst) when they fit, otherwise streamed per chunk through small
// TMA rings (late layers: 120 -> 720 -> 240 channels).  HBM traffic per block = input (with halo) + output.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tc.cuh"

struct FusedGeom {
  int B, Hi, Wi, Ho, Wo;
  int Cin, Cexp, Cout;
  int stride, pad_t, pad_l;
  int residual;
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
  int nx;                 // input tile buffers (1 or 2)
  int nd;                 // expand accumulator buffers in TMEM (1 or 2)
  int se, sp;             // expand / project weight slots; == n_chunks: resident, smaller: streamed ring
  uint32_t tmem_cols;
  uint32_t we_chunk_bytes, wp_chunk_bytes, x_buf_bytes, e_buf_bytes;
  uint32_t off_X, off_A2, off_WE, off_WP, off_E, off_wd, off_bars, smem_bytes;
  long long* dbg;         // HFB_FUSED_DBG: per-phase clock stamps of CTA 0 ([phase][warp][8]), else null
};

#define FB_MAX_RING 4

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

// NT = compute threads per CTA (512, one CTA per SM); one extra warp only issues TMA / MMA work, so that no compute warp
// ever idles behind the single issuing lane.  S (stride) and TH (tile height) are compile-time: every shared-memory
// offset of the depthwise taps is an immediate and the per-chunk loops have constant trip counts -- at ~300 useful
// instructions per thread and phase, address arithmetic on run-time geometry would otherwise double the instruction count.
template <int S, int NT, int TH>
__global__ void __launch_bounds__(NT + 32, 1) fused_block_kernel(const __grid_constant__ CUtensorMap tmWE,
                                                                 const __grid_constant__ CUtensorMap tmWP,
                                                                 const FusedGeom g, const __half* __restrict__ in,
                                                                 const float* __restrict__ be,   // expand bias [Cexp]
                                                                 const float* __restrict__ wd,   // dw weights [9][Cexp]
                                                                 const float* __restrict__ bd,   // dw bias [Cexp]
                                                                 const float* __restrict__ bp,   // project bias [Cout]
                                                                 __half* __restrict__ out) {
  constexpr int IW = 15 * S + 3;            // halo tile width
  constexpr int IH = (TH - 1) * S + 3;      // halo tile height
  constexpr int R = IH * IW;                // halo pixels = rows of the expand GEMM
  constexpr int MT = (R + 127) / 128;       // expand M-tiles
  constexpr int CW = S == 1 ? 64 : 32;      // expanded channels per chunk
  constexpr int E_PITCH = CW * 2 + 16;      // bytes per row of E
  constexpr int NPIX = TH * 16;             // output pixels per tile
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic (keeps the shared-memory address space visible to the compiler: LDS/STS
  // instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem + g.off_X;      // [nx][kb_in][MT*128 rows][xrb B] swizzled
  uint8_t* sA2 = smem + g.off_A2;    // [2][128 rows][128 B] swizzled (written by the depthwise stage)
  uint8_t* sWE = smem + g.off_WE;    // [se][kb_in][CW rows][128 B] swizzled (TMA)
  uint8_t* sWP = smem + g.off_WP;    // [sp][cout_pad rows][128 B] swizzled (TMA)
  uint8_t* sE = smem + g.off_E;      // [2][R][E_PITCH] expanded activations, fp16
  __half* s_wd = reinterpret_cast<__half*>(smem + g.off_wd);                    // [9][cexp_pad] dw weights (fp16-exact)
  float* s_bd = reinterpret_cast<float*>(smem + g.off_wd + 18 * g.cexp_pad);    // [cexp_pad] dw bias
  float* s_be = s_bd + g.cexp_pad;                                              // [cexp_pad] expand bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bars);
  uint64_t* bar_e = bars;                     // [2] expand MMAs of chunk c retired (slot c & 1)
  uint64_t* bar_p = bars + 2;                 // [2] project MMAs of chunk c retired (slot c & 1)
  uint64_t* bar_we = bars + 4;                // [FB_MAX_RING] expand weights landed (slot 0 only when resident)
  uint64_t* bar_wp = bars + 4 + FB_MAX_RING;  // [FB_MAX_RING] project weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * FB_MAX_RING);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_issuer = warp == NT / 32;           // the extra warp
  const bool issue_lane = is_issuer && lane == 0;
  const int n_chunks = g.n_chunks;
  const int my_tiles =
      ((int)blockIdx.x < g.total_tiles) ? (g.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int Ctot = my_tiles * n_chunks;
  const bool stream_e = g.se < n_chunks, stream_p = g.sp < n_chunks;
  tc::pdl_launch_dependents();

  // weight loads by chunk (issuer lane only)
  auto load_we = [&](int c) {   // expand weights of global chunk c into its slot
    const int j = c % n_chunks, slot = stream_e ? c % g.se : j;
    uint64_t* bar = &bar_we[stream_e ? slot : 0];
    if (stream_e) tc::mbar_expect_tx(bar, g.we_chunk_bytes);
    for (int kb = 0; kb < g.kb_in; ++kb)
      tc::tma_load_2d(sWE + (size_t)slot * g.we_chunk_bytes + (size_t)kb * CW * 128, &tmWE, bar, kb * 64, j * CW);
  };
  auto load_wp = [&](int c) {
    const int j = c % n_chunks, slot = stream_p ? c % g.sp : j;
    uint64_t* bar = &bar_wp[stream_p ? slot : 0];
    if (stream_p) tc::mbar_expect_tx(bar, g.wp_chunk_bytes);
    tc::tma_load_2d(sWP + (size_t)slot * g.wp_chunk_bytes, &tmWP, bar, j * CW, 0);
  };

  if (issue_lane) {
    tc::prefetch_tmap(&tmWE);
    tc::prefetch_tmap(&tmWP);
    for (int i = 0; i < 4 + 2 * FB_MAX_RING; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_barrier_init();
    if (!stream_e) tc::mbar_expect_tx(&bar_we[0], (uint32_t)n_chunks * g.we_chunk_bytes);
    if (!stream_p) tc::mbar_expect_tx(&bar_wp[0], (uint32_t)n_chunks * g.wp_chunk_bytes);
    const int ne = stream_e ? min(g.se, Ctot) : n_chunks, np = stream_p ? min(g.sp, Ctot) : n_chunks;
    for (int c = 0; c < ne; ++c) load_we(c);
    for (int c = 0; c < np; ++c) load_wp(c);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, g.tmem_cols);
  for (int i = tid; i < 11 * g.cexp_pad; i += NT + 32) {
    const int row = i / g.cexp_pad, c = i - row * g.cexp_pad;
    float v = 0.f;
    if (c < g.Cexp) {
      if (row < 9) v = __ldg(wd + (size_t)row * g.Cexp + c);
      else if (row == 9) v = __ldg(bd + c);
      else v = __ldg(be + c);
    }
    if (row < 9) s_wd[i] = __float2half_rn(v);   // exact: the loader stores fp16-representable depthwise weights
    else s_bd[i - 9 * g.cexp_pad] = v;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  tc::pdl_wait();   // everything above touched only weights / on-chip state; the input tensor is the predecessor's output
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base + (uint32_t)(g.nd * MT * CW);
  const uint32_t idesc_p = tc::make_idesc_f16(g.cout_pad);

  // Tile cursors advance by gridDim.x tiles without divisions: (tx, ty, img) += (dx, dy, dimg) with carries.
  struct TilePos { int tx, ty, img; };
  TilePos pos0;
  {
    int t = (int)blockIdx.x;
    pos0.tx = t % g.tiles_x;
    t /= g.tiles_x;
    pos0.ty = t % g.tiles_y;
    pos0.img = t / g.tiles_y;
  }
  int step_x, step_y, step_img;
  {
    int t = (int)gridDim.x;
    step_x = t % g.tiles_x;
    t /= g.tiles_x;
    step_y = t % g.tiles_y;
    step_img = t / g.tiles_y;
  }
  auto advance = [&](TilePos& p) {
    p.tx += step_x;
    int carry = 0;
    if (p.tx >= g.tiles_x) { p.tx -= g.tiles_x; carry = 1; }
    p.ty += step_y + carry;
    carry = 0;
    if (p.ty >= g.tiles_y) { p.ty -= g.tiles_y; carry = 1; }
    p.img += step_img + carry;
  };

  // input halo tile -> swizzled K-major rows via cp.async (zero fill outside the image and in the K padding)
  auto load_x = [&](const TilePos& tp, int xbuf) {
    const int iy0 = tp.ty * TH * S - g.pad_t, ix0 = tp.tx * 16 * S - g.pad_l;
    const int units = ((g.Cin + 15) & ~15) >> 3;   // 16-byte units per pixel incl. K padding
    const int vunits = g.Cin >> 3;                 // units holding real channels
    const __half* src = in + (size_t)tp.img * g.Hi * g.Wi * g.Cin;
    const uint32_t xbase = tc::smem_u32(sX) + (uint32_t)xbuf * g.x_buf_bytes;
    for (int r = tid; r < R; r += NT) {
      const int ry = r / IW, rx = r - ry * IW;
      const int iy = iy0 + ry, ix = ix0 + rx;
      const bool inb = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
      const __half* gp = inb ? src + ((size_t)iy * g.Wi + ix) * g.Cin : in;
      const uint32_t row_dst = xbase + (uint32_t)r * (uint32_t)g.xrb;
      for (int u = 0; u < units; ++u) {
        const bool ok = inb && u < vunits;
        // 16-byte chunk u of row r under the operand swizzle: SWIZZLE_128B = chunk ^ (row & 7) in 128-byte rows,
        // SWIZZLE_64B = chunk ^ ((row >> 1) & 3) in 64-byte rows (address bits [4,6) ^= bits [7,9))
        const uint32_t dst = g.xrb == 128 ? row_dst + (uint32_t)(u >> 3) * (uint32_t)(MT * 128 * 128) +
                                                (uint32_t)(((u & 7) ^ (r & 7)) << 4)
                                          : row_dst + (uint32_t)((u ^ ((r >> 1) & 3)) << 4);
        const int nbytes = ok ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? gp + u * 8 : in), "r"(nbytes)
                     : "memory");
      }
    }
  };

  // ---- MMA issue (issuer lane only)
  auto issue_expand = [&](int c, int j, int ring_slot, uint32_t ring_par, int xbuf) {
    const int slot = stream_e ? ring_slot : j;
    tc::mbar_wait(&bar_we[stream_e ? slot : 0], stream_e ? ring_par : 0u);
    const int cvalid = min(CW, g.Cexp - j * CW);
    const uint32_t idesc = tc::make_idesc_f16((cvalid + 15) & ~15);
    const uint8_t* xb = sX + (size_t)xbuf * g.x_buf_bytes;
    const uint32_t d1 = tmem_base + (uint32_t)((c & (g.nd - 1)) * MT * CW);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      for (int kb = 0; kb < g.kb_in; ++kb) {
        const uint64_t da = g.xrb == 128 ? tc::make_sdesc_sw128(tc::smem_u32(xb + ((size_t)kb * MT + mt) * 128 * 128))
                                         : tc::make_sdesc_sw64(tc::smem_u32(xb + (size_t)mt * 128 * 64));
        const uint64_t db =
            tc::make_sdesc_sw128(tc::smem_u32(sWE + (size_t)slot * g.we_chunk_bytes + (size_t)kb * CW * 128));
        const int krem = g.Cin - kb * 64;
        const int nk = krem >= 64 ? 4 : (krem + 15) >> 4;
        for (int k = 0; k < nk; ++k)
          tc::umma_f16(d1 + (uint32_t)(mt * CW), tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k), idesc,
                       (kb > 0 || k > 0) ? 1u : 0u);
      }
    }
    tc::umma_commit(&bar_e[c & 1]);
  };
  auto issue_project = [&](int c, int j, int ring_slot, uint32_t ring_par) {
    const int slot = stream_p ? ring_slot : j;
    tc::mbar_wait(&bar_wp[stream_p ? slot : 0], stream_p ? ring_par : 0u);
    const int cw16 = (min(CW, g.Cexp - j * CW) + 15) & ~15;
    const uint64_t da = tc::make_sdesc_sw128(tc::smem_u32(sA2 + (size_t)(c & 1) * 128 * 128));
    const uint64_t db = tc::make_sdesc_sw128(tc::smem_u32(sWP + (size_t)slot * g.wp_chunk_bytes));
    for (int k = 0; k < (cw16 >> 4); ++k)
      tc::umma_f16(tmem_d2, tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k), idesc_p, (j > 0 || k > 0) ? 1u : 0u);
    tc::umma_commit(&bar_p[c & 1]);
  };

  // ---- depthwise 3x3 of one chunk, UPOW = units rounded up to 2, 4 or 8: PSTEP = NT / UPOW threads walk the tile's
  // pixels of one unit (8 channels) with its 72 weights (36 packed registers) + 8 biases resident.  Stride 2: adjacent
  // lanes take adjacent units so that a quarter-warp's 16-byte loads (pixel stride = 2 rows of E) hit distinct banks.
  auto dw_chunk = [&](auto upow_c, int units, int c0, const uint8_t* ebuf, uint8_t* a2) {
    constexpr int UPOW = decltype(upow_c)::value;
    constexpr int PSTEP = NT / UPOW;
    int u, pb;
    if (S == 2) {
      u = 2 * (tid / (2 * PSTEP)) + (tid & 1);
      pb = (tid >> 1) % PSTEP;
    } else {
      u = tid / PSTEP;
      pb = tid % PSTEP;
    }
    if (u >= units || pb >= NPIX) return;
    uint4 w[9];
    float bias8[8];
    {
      const float4* bq = reinterpret_cast<const float4*>(s_bd + c0 + u * 8);
      const float4 b0 = bq[0], b1 = bq[1];
      bias8[0] = b0.x; bias8[1] = b0.y; bias8[2] = b0.z; bias8[3] = b0.w;
      bias8[4] = b1.x; bias8[5] = b1.y; bias8[6] = b1.z; bias8[7] = b1.w;
      const __half* wp = s_wd + c0 + u * 8;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) w[tp] = *reinterpret_cast<const uint4*>(wp + tp * g.cexp_pad);
    }
    const uint8_t* e0 = ebuf + (size_t)(((pb >> 4) * S) * IW + (pb & 15) * S) * E_PITCH + (size_t)u * 16;
    uint8_t* o0 = a2 + (size_t)pb * 128 + (size_t)((u ^ (pb & 7)) << 4);
    constexpr int ITERS = NPIX > PSTEP ? NPIX / PSTEP : 1;   // PSTEP is a multiple of 16 pixel columns and of 8 rows of A2
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = bias8[c];
      const uint8_t* e = e0 + (size_t)it * ((PSTEP / 16) * S * IW * E_PITCH);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint4 x = *reinterpret_cast<const uint4*>(e + (ky * IW + kx) * E_PITCH);
          const uint4 wq = w[ky * 3 + kx];
          fma2_f16(acc[0], acc[1], x.x, wq.x);
          fma2_f16(acc[2], acc[3], x.y, wq.y);
          fma2_f16(acc[4], acc[5], x.z, wq.z);
          fma2_f16(acc[6], acc[7], x.w, wq.w);
        }
      }
      uint4 o;   // channels beyond Cexp (zero weights, zero bias) come out as exact zeros
      o.x = relu6_pack(acc[0], acc[1]);
      o.y = relu6_pack(acc[2], acc[3]);
      o.z = relu6_pack(acc[4], acc[5]);
      o.w = relu6_pack(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(o0 + (size_t)it * (PSTEP * 128)) = o;
    }
  };

  // ---- prologue: first input tile(s), first expand(s)
  TilePos lx_pos = pos0;   // next tile to load
  int lx_iter = 0;         // tiles whose load has been issued (tracked by every thread, issued by the compute threads)
  for (int i = 0; i < g.nx && lx_iter < my_tiles; ++i) {
    if (!is_issuer) load_x(lx_pos, i);
    advance(lx_pos);
    ++lx_iter;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  tc::fence_proxy_async();
  __syncthreads();
  int x_ready = lx_iter;   // tiles whose input has landed (as of the last barrier)

  // Incremental chunk cursors (no divisions in the phase loop): chunk q (depthwise / project), chunk q+1 (read-back),
  // next chunk whose expand is to be issued: j = chunk within its tile, tile = tile iteration of this CTA.
  int jq = -1;                  // chunk q; starts at q = -1
  int jr = 0;                   // chunk q+1
  int ex_c = 0, ex_j = 0, ex_tile = 0, ex_slot = 0;   // expand issue cursor (issuer lane)
  uint32_t ex_par = 0u;
  int ps_slot = 0;                                 // project-weight ring slot / fill parity of chunk q (valid for q >= 0)
  uint32_t ps_par = 0u;
  TilePos rb_pos = pos0;                           // tile of chunk q+1
  TilePos ep_pos = pos0;                           // tile of the next epilogue
  // expand(c) may be issued once D1[c % nd] has been drained (read-back of chunk c-nd ran in phase c-nd-1, so after
  // the barrier of phase q every c <= q+1+nd qualifies) and its input tile has landed
  auto issue_expands = [&](int q) {
    while (ex_c < Ctot && ex_c <= q + 1 + g.nd && ex_tile < x_ready) {
      issue_expand(ex_c, ex_j, ex_slot, ex_par, g.nx == 2 ? (ex_tile & 1) : 0);
      ++ex_c;
      if (++ex_j == n_chunks) { ex_j = 0; ++ex_tile; }
      if (stream_e && ++ex_slot == g.se) { ex_slot = 0; ex_par ^= 1u; }
    }
  };
  if (issue_lane) {
    tc::fence_after_sync();
    issue_expands(-2);
  }

#define FB_STAMP(k)                                                                             \
  if (g.dbg && blockIdx.x == 0 && lane == 0 && q + 1 < 64) g.dbg[((q + 1) * 17 + warp) * 8 + (k)] = clock64()
  for (int q = -1; q <= Ctot; ++q) {
    FB_STAMP(0);
    // ---- A) depthwise 3x3 of chunk q: E[q & 1] -> A2[q & 1]
    if (!is_issuer && q >= 0 && q < Ctot) {
      if (q >= 2) {   // project(q-2) read A2[q & 1]
        const uint32_t k = (uint32_t)(q - 2);
        tc::mbar_wait(&bar_p[k & 1u], (k >> 1) & 1u);
      }
      const int c0 = jq * CW;
      const int units = ((min(CW, g.Cexp - c0) + 15) & ~15) >> 3;
      const uint8_t* ebuf = sE + (size_t)(q & 1) * g.e_buf_bytes;
      uint8_t* a2 = sA2 + (size_t)(q & 1) * 128 * 128;
      if (units > 4) dw_chunk(std::integral_constant<int, 8>{}, units, c0, ebuf, a2);
      else if (units > 2) dw_chunk(std::integral_constant<int, 4>{}, units, c0, ebuf, a2);
      else dw_chunk(std::integral_constant<int, 2>{}, units, c0, ebuf, a2);
    }

    FB_STAMP(1);
    // ---- B) read-back of chunk q+1: TMEM D1 -> +bias, ReLU6, zero outside the image -> E[(q+1) & 1]
    if (!is_issuer && q + 1 < Ctot) {
      const int c = q + 1;
      const int j = jr;
      tc::mbar_wait(&bar_e[c & 1], (uint32_t)((c >> 1) & 1));
      FB_STAMP(2);
      __syncwarp();
      tc::fence_after_sync();
      // the tile's last expand has retired: its input buffer is free for the tile nx ahead
      if (j == n_chunks - 1 && lx_iter < my_tiles) load_x(lx_pos, g.nx == 2 ? (lx_iter & 1) : 0);
      const int iy0 = rb_pos.ty * TH * S - g.pad_t, ix0 = rb_pos.tx * 16 * S - g.pad_l;
      const int c0 = j * CW;
      const int cw16 = (min(CW, g.Cexp - c0) + 15) & ~15;
      uint8_t* eb = sE + (size_t)(c & 1) * g.e_buf_bytes;
      // work units = (M-tile, column half) spread over the warp quads; warp w reads TMEM lane group w % 4
      const int nch = cw16 >> 4, ch_half = (nch + 1) >> 1;
#pragma unroll 1
      for (int wu = warp >> 2; wu < MT * 2; wu += NT / 128) {
        const int mt = wu >> 1, half = wu & 1;
        const int cbeg = half ? ch_half * 16 : 0, cend = half ? cw16 : ch_half * 16;
        const int r = mt * 128 + (warp & 3) * 32 + lane;
        const int iy = iy0 + r / IW, ix = ix0 + r % IW;
        const bool in_img = r < R && iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((c & (g.nd - 1)) * MT * CW + mt * CW);
        uint8_t* erow = eb + (size_t)r * E_PITCH;
#pragma unroll 1
        for (int cc = cbeg; cc < cend; cc += 16) {
          uint32_t v[16];
          tc::tmem_ld16(taddr + (uint32_t)cc, v);
          tc::tmem_ld_wait();
          if (r < R) {
            uint4 o[2];
            uint32_t* ho = reinterpret_cast<uint32_t*>(o);
            if (in_img) {
              // bias add in fp32, one rounding to fp16 with the ReLU6 clamp folded into the conversion
              const float4* bq = reinterpret_cast<const float4*>(s_be + c0 + cc);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b4 = bq[i];
                ho[2 * i] = relu6_pack(__uint_as_float(v[4 * i]) + b4.x, __uint_as_float(v[4 * i + 1]) + b4.y);
                ho[2 * i + 1] = relu6_pack(__uint_as_float(v[4 * i + 2]) + b4.z, __uint_as_float(v[4 * i + 3]) + b4.w);
              }
            } else {
              o[0] = make_uint4(0, 0, 0, 0);
              o[1] = make_uint4(0, 0, 0, 0);
            }
            uint4* d = reinterpret_cast<uint4*>(erow + (size_t)cc * 2);
            d[0] = o[0];
            d[1] = o[1];
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // one (possibly empty) group per phase
    FB_STAMP(3);

    // ---- C) epilogue of the tile whose last chunk was q-1: D2 + bias (+residual) -> fp16 NHWC
    if (!is_issuer && q >= 1 && jq == 0) {   // chunk q opens a new tile (or q == Ctot): chunk q-1 closed the tile at ep_pos
      const uint32_t k = (uint32_t)(q - 1);
      tc::mbar_wait(&bar_p[k & 1u], (k >> 1) & 1u);
      __syncwarp();
      tc::fence_after_sync();
      const int p = (warp & 3) * 32 + lane;
      const int oy = ep_pos.ty * TH + (p >> 4), ox = ep_pos.tx * 16 + (p & 15);
      const bool valid = p < NPIX && oy < g.Ho && ox < g.Wo;
      const long long opix = ((long long)ep_pos.img * g.Ho + oy) * g.Wo + ox;
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
          const float4 bb0 = __ldg(reinterpret_cast<const float4*>(bp + n)), bb1 = __ldg(reinterpret_cast<const float4*>(bp + n) + 1);
          float f[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += __uint_as_float(v[8 * h + i]);
          if (g.residual) {
            const uint4 rq = *reinterpret_cast<const uint4*>(in + opix * g.Cin + n);
            const __half2* hq = reinterpret_cast<const __half2*>(&rq);
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
      advance(ep_pos);
    }

    FB_STAMP(4);
    // ---- phase boundary: A2 / E / X writes visible to the async proxy, TMEM reads ordered, then the MMA issues
    const int lx_before = lx_iter;
    if (q + 1 < Ctot && jr == n_chunks - 1 && lx_iter < my_tiles) {   // a load was issued in this phase (all threads track)
      advance(lx_pos);
      ++lx_iter;
    }
    if (g.nx == 2) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // all but this phase's load have landed
      x_ready = lx_before;
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      x_ready = lx_iter;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    FB_STAMP(5);
    __syncthreads();
    FB_STAMP(6);
    if (issue_lane) {
      tc::fence_after_sync();
      if (q >= 0 && q < Ctot) issue_project(q, jq, ps_slot, ps_par);
      // ring refills (before the expand issues, which may already need them): expand weights of chunk q+1 are free once
      // its MMAs (issued at least a phase ago) have retired, project weights of chunk q-1 once project(q-1) has
      if (stream_e && q + 1 >= 0 && q + 1 + g.se < Ctot) {
        const uint32_t k = (uint32_t)(q + 1);
        tc::mbar_wait(&bar_e[k & 1u], (k >> 1) & 1u);
        load_we(q + 1 + g.se);
      }
      if (stream_p && q - 1 >= 0 && q - 1 + g.sp < Ctot) {
        const uint32_t k = (uint32_t)(q - 1);
        tc::mbar_wait(&bar_p[k & 1u], (k >> 1) & 1u);
        load_wp(q - 1 + g.sp);
      }
      issue_expands(q);
    }
    FB_STAMP(7);
    // advance the cursors to phase q+1
    if (q >= 0 && stream_p && ++ps_slot == g.sp) { ps_slot = 0; ps_par ^= 1u; }
    if (++jq == n_chunks) jq = 0;
    if (++jr == n_chunks) {
      jr = 0;
      advance(rb_pos);
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
// Configuration search, in order of preference: the tallest of 8 / 4 / 2 x 16 output tiles that still gives every CTA a
// tile (small late layers) and fits;
// resident weights, else 3- then 2-slot rings; two input buffers, else one.
// HFB_FUSED_TH / HFB_FUSED_NX (environment) pin a choice for experiments.
int fused_block_plan(hfb_ctx* ctx, FusedPlan& fp, const BlockW& bw, const __half* in, int Bmax, int Hi, int Wi, int Ho,
                     int Wo, int pad_t, int pad_l) {
  (void)in;
  if (!bw.has_expand) return HFB_ERR_CAPACITY;   // layer_2 has its own kernel
  FusedGeom& g = fp.g;
  g.B = Bmax; g.Hi = Hi; g.Wi = Wi; g.Ho = Ho; g.Wo = Wo;
  g.Cin = bw.cin; g.Cexp = bw.cexp; g.Cout = bw.cout;
  g.stride = bw.stride; g.pad_t = pad_t; g.pad_l = pad_l;
  g.residual = bw.residual ? 1 : 0;
  g.dbg = nullptr;
  g.tiles_x = (Wo + 15) / 16;
  g.CW = bw.stride == 1 ? 64 : 32;                // compile-time constant of the kernel
  if (bw.cexp < g.CW) return HFB_ERR_CAPACITY;
  g.n_chunks = (bw.cexp + g.CW - 1) / g.CW;
  g.cexp_pad = g.n_chunks * g.CW;
  g.kb_in = (bw.cin + 63) / 64;
  g.xrb = bw.cin <= 32 ? 64 : 128;
  g.cout_pad = (bw.cout + 15) & ~15;
  g.e_pitch = g.CW * 2 + 16;
  g.we_chunk_bytes = (uint32_t)(g.kb_in * g.CW * 128);
  g.wp_chunk_bytes = (uint32_t)(g.cout_pad * 128);
  if (g.cout_pad > 256) return HFB_ERR_CAPACITY;   // one UMMA N
  auto al = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  auto layout = [&](int th, int nx, int se, int sp) {
    g.TH = th; g.nx = nx; g.se = se; g.sp = sp;
    g.tiles_y = (Ho + g.TH - 1) / g.TH;
    g.IH = (g.TH - 1) * bw.stride + 3;
    g.IW_ = 15 * bw.stride + 3;
    g.R = g.IH * g.IW_;
    g.MT = (g.R + 127) / 128;
    g.x_buf_bytes = al((uint32_t)(g.kb_in * g.MT * 128 * g.xrb));
    g.e_buf_bytes = al((uint32_t)(g.R * g.e_pitch));
    uint32_t off = 0;
    g.off_X = off;  off += (uint32_t)nx * g.x_buf_bytes;
    g.off_A2 = off; off += 2 * 128 * 128;
    g.off_WE = off; off += al((uint32_t)se * g.we_chunk_bytes);
    g.off_WP = off; off += al((uint32_t)sp * g.wp_chunk_bytes);
    g.off_E = off;  off += 2 * g.e_buf_bytes;
    g.off_wd = off; off += al((uint32_t)(26 * g.cexp_pad));
    g.off_bars = off; off += 128;
    g.smem_bytes = off + 1024;
    g.nd = (2 * g.MT * g.CW + g.cout_pad <= 512) ? 2 : 1;
    uint32_t cols = 32;
    while ((int)cols < g.nd * g.MT * g.CW + g.cout_pad) cols <<= 1;
    g.tmem_cols = cols;
  };
  auto env_int = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
  const int pin_nt = env_int("HFB_FUSED_NT"), pin_th = env_int("HFB_FUSED_TH"), pin_nx = env_int("HFB_FUSED_NX");
  bool found = false;
  for (int nt : {512}) {
    if (found) break;
    if (pin_nt && nt != pin_nt) continue;
    const int ctas = nt == 256 ? 2 : 1;
    const uint32_t budget = nt == 256 ? 112u * 1024u : 224u * 1024u;
    const uint32_t max_cols = nt == 256 ? 256u : 512u;
    // tallest tile that still gives every resident CTA a tile
    int th0 = 8;
    while (th0 > 2 && g.tiles_x * ((Ho + th0 - 1) / th0) * Bmax < ctas * ctx->n_sm) th0 >>= 1;
    if (pin_th) th0 = pin_th;
    for (int th = th0; th >= 2 && !found; th >>= 1) {
      // resident weights (one mbarrier transaction count each: < 1 MiB), else rings
      const int ring[3] = {g.n_chunks, 3, 2};
      for (int ri = 0; ri < 3 && !found; ++ri) {
        const int s = std::min(ring[ri], g.n_chunks);
        if (ri > 0 && s >= g.n_chunks) continue;
        if (s == g.n_chunks && (uint64_t)s * std::max(g.we_chunk_bytes, g.wp_chunk_bytes) >= (1u << 20)) continue;
        for (int nx : {2, 1}) {
          if (pin_nx && nx != pin_nx) continue;
          layout(th, nx, s, s);
          if (g.smem_bytes <= budget && g.tmem_cols <= max_cols) {
            found = true;
            fp.nt = nt;
            fp.ctas_per_sm = ctas;
            break;
          }
        }
      }
      if (pin_th) break;
    }
  }
  if (!found) return HFB_ERR_CAPACITY;
  g.total_tiles = g.tiles_x * g.tiles_y * Bmax;
  HFB_TRY(hfb_make_tmap_2d(ctx, &fp.tmWE, bw.expand.w, (uint64_t)bw.expand.Kp, (uint64_t)bw.expand.N,
                           (uint64_t)bw.expand.Kp * 2, (uint32_t)g.CW));
  HFB_TRY(hfb_make_tmap_2d(ctx, &fp.tmWP, bw.project.w, (uint64_t)bw.project.Kp, (uint64_t)bw.project.N,
                           (uint64_t)bw.project.Kp * 2, (uint32_t)g.cout_pad));
  if (ctx->trace)
    fprintf(stderr,
            "hfnet_b200: fused layer_%d: NT=%d x%d TH=%d MT=%d CW=%d chunks=%d xrb=%d nx=%d nd=%d ring=%d smem=%u tmem=%u tiles=%d\n",
            bw.layer, fp.nt, fp.ctas_per_sm, g.TH, g.MT, g.CW, g.n_chunks, g.xrb, g.nx, g.nd, g.se, g.smem_bytes, g.tmem_cols,
            g.total_tiles);
  return HFB_OK;
}

int fused_block_tiles(const FusedPlan& fp, int B) { return fp.g.tiles_x * fp.g.tiles_y * B; }

template <int S, int TH>
static int fused_launch(hfb_ctx* ctx, const FusedPlan& fp, const FusedGeom& g, const BlockW& bw, const __half* in,
                        __half* out, int grid) {
  static SmemOptIn optin;   // per instantiation
  HFB_CUDA(ctx, optin.ensure(fused_block_kernel<S, 512, TH>, ctx->device, g.smem_bytes));
  hfb_launch(ctx, fused_block_kernel<S, 512, TH>, grid, 512 + 32, g.smem_bytes, fp.tmWE, fp.tmWP, g, in, bw.expand.b, bw.wd,
             bw.bd, bw.project.b, out);
  HFB_CHECK_LAUNCH(ctx, "fused_block");
  return HFB_OK;
}

// HFB_FUSED_DBG=<layer>: stamp the phase sections of CTA 0 with clock64() and print the per-section means (cycles).
static int fused_debug_report(hfb_ctx* ctx, long long* d_dbg, int layer) {
  std::vector<long long> h(64 * 17 * 8);
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  HFB_CUDA(ctx, cudaMemcpy(h.data(), d_dbg, h.size() * 8, cudaMemcpyDeviceToHost));
  const char* names[7] = {"dw", "wait_e", "readback", "epilogue", "fences", "barrier", "issue"};
  for (int w : {0, 5, 16}) {
    double sum[7] = {0};
    int n = 0;
    for (int ph = 6; ph < 26; ++ph) {
      const long long* r = &h[(size_t)(ph * 17 + w) * 8];
      if (!r[0] || !r[7]) continue;
      for (int k = 0; k < 7; ++k) sum[k] += (double)(r[k + 1] - r[k]);
      ++n;
    }
    if (!n) continue;
    fprintf(stderr, "hfnet_b200: fused layer_%d warp %2d (%d phases):", layer, w, n);
    for (int k = 0; k < 7; ++k) fprintf(stderr, " %s=%.0f", names[k], sum[k] / n);
    const long long* a = &h[(size_t)(6 * 17 + w) * 8];
    const long long* b = &h[(size_t)(26 * 17 + w) * 8];
    if (a[0] && b[0]) fprintf(stderr, " | phase=%.0f", (double)(b[0] - a[0]) / 20.0);
    fprintf(stderr, "\n");
  }
  return HFB_OK;
}

int fused_block_run(hfb_ctx* ctx, const FusedPlan& fp, const BlockW& bw, const __half* in, __half* out, int B) {
  FusedGeom g = fp.g;
  g.B = B;
  g.dbg = nullptr;
  static int dbg_layer = getenv("HFB_FUSED_DBG") ? atoi(getenv("HFB_FUSED_DBG")) : 0;
  static long long* d_dbg = nullptr;
  static int dbg_runs = 0;
  const bool dbg = dbg_layer == bw.layer && dbg_runs < 3;
  if (dbg) {
    if (!d_dbg) HFB_CUDA(ctx, cudaMalloc(&d_dbg, 64 * 17 * 8 * 8));
    HFB_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, 64 * 17 * 8 * 8, ctx->stream));
    g.dbg = d_dbg;
    ++dbg_runs;
  }
  g.total_tiles = g.tiles_x * g.tiles_y * B;
  const int grid = std::min(g.total_tiles, ctx->n_sm * fp.ctas_per_sm);
#define FB_CASE(s, th)                                                    \
  if (g.stride == s && g.TH == th) {                                      \
    HFB_TRY((fused_launch<s, th>(ctx, fp, g, bw, in, out, grid)));        \
    return dbg ? fused_debug_report(ctx, d_dbg, bw.layer) : HFB_OK;       \
  }
  FB_CASE(1, 8); FB_CASE(1, 4); FB_CASE(1, 2);
  FB_CASE(2, 8); FB_CASE(2, 4); FB_CASE(2, 2);
#undef FB_CASE
  ctx->set_error("internal: no fused block kernel for this tile shape");
  return HFB_ERR_STATE;
}

double fused_block_bytes(const FusedPlan& fp, int B) {   // algorithmic: input once + output once + weights
  const FusedGeom& g = fp.g;
  return 2.0 * B * ((double)g.Hi * g.Wi * g.Cin + (double)g.Ho * g.Wo * g.Cout * (g.residual ? 2 : 1)) +
         2.0 * ((double)g.Cin * g.Cexp + (double)g.Cexp * g.Cout) + 4.0 * 10 * g.Cexp;
}
double fused_block_flops(const FusedPlan& fp, int B) {
  const FusedGeom& g = fp.g;
  return 2.0 * B * ((double)g.Hi * g.Wi * g.Cin * g.Cexp + (double)g.Ho * g.Wo * g.Cexp * (9 + g.Cout));
}
