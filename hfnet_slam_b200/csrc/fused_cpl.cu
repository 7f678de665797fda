// One MobileNetV2 inverted-residual block (hfnet/models/backbones/utils/conv_blocks.py:162-312) as ONE kernel,
// "channel per lane" formulation:
//   1x1 expand (+bias, ReLU6)  ->  3x3 depthwise stride 1|2, TF-SAME (+bias, ReLU6)  ->  1x1 project (+bias, +residual)
// The expand GEMM is issued TRANSPOSED: A = 128 expanded channels x K (weights, K-major), B = halo pixels x K (input
// tile, K-major), so the fp32 accumulator in tensor memory holds one expanded CHANNEL per TMEM lane and one halo PIXEL
// per column.  A thread (= one channel) then pulls whole pixel rows of its channel into registers with tcgen05.ld and
// runs the depthwise 3x3 as a register sliding window: no shared-memory round trip for the 6x-expanded tensor, scalar
// per-thread depthwise weights, 9 FMAs per output and nothing else.  The depthwise output goes to shared memory as the
// MN-major (pixel-contiguous) A operand of the project GEMM (32 contiguous bytes per thread and row), which accumulates
// over the 128-channel chunks in a second TMEM accumulator with pixels in lanes, ready for a row-per-thread epilogue.
//   * expand bias and SAME zero padding cost nothing: the input tile carries two extra "ones" channels (1 inside the
//     image, 0 outside) and the weight image carries the bias in those K columns (fp16 high part + fp16 residual, i.e.
//     fp32-accurate), so out-of-image pixels expand to exactly 0 = the zero padding the depthwise conv must see
//     (padding applies to the EXPANDED activation).
//   * 16 compute warps (4 TMEM lane groups x 4 row groups) never meet at a CTA barrier: mbarriers pair them with one
//     issuing thread that keeps expand(c+1) and project(c-1) in flight while chunk c is in the CUDA cores.
//   * weights arrive as pre-swizzled per-chunk images (one cp.async.bulk each) that stay resident when they fit and
//     stream through a ring otherwise; input halo tiles are prefetched NX-1 tiles ahead with cp.async.
// HBM traffic per block = input (with halo) + output.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tc.cuh"

namespace {

struct CplGeom {
  int B, Hi, Wi, Ho, Wo;
  int Cin, Cexp, Cout;
  int pad_t, pad_l, residual;
  int tiles_x, tiles_y, total_tiles;
  int n_chunks;      // ceil(Cexp / 128)
  int units;         // 16-byte units per input-tile row: Cin/8 data units, one "ones" unit, zero units up to K % 16 == 0
  int ones_unit;     // = Cin / 8
  int xrb;           // bytes per operand row of a k-block: 64 (K = 32, SWIZZLE_64B) or 128 (SWIZZLE_128B)
  int kb_in;         // k-blocks of 64 channels (1 when xrb == 64)
  int nk16;          // K steps of the expand MMA
  int cout_pad;      // Cout rounded up to 16 (project N)
  int NSe, NSp, NX, ND2;   // expand / project weight ring slots (== n_chunks: resident), input tile buffers, project accumulators
  int rem_vp;        // last chunk holds <= 64 channels: they are replicated every rem_vp (32 | 64) lanes, 0 = no
  uint32_t we_bytes, wp_bytes, blob_bytes;
  uint32_t x_buf_bytes, a2_buf_bytes;
  uint32_t off_X, off_A2, off_WE, off_WP, off_DW, off_bars, off_bias, smem_bytes;
  uint32_t tmem_cols;
  long long* dbg;    // HFB_CPL_DBG=<layer>: clock stamps of CTA 0, [role 0..4][chunk or tile 0..63][8]
};

constexpr int CPL_NT = 512;          // compute threads
constexpr int CPL_THREADS = 768;     // 16 depthwise warps + 2 issuer warps + 2 loader warps + 4 epilogue warps
constexpr int CPL_DW_BYTES = 6 * 128 * 4;   // per chunk: 5 words of packed fp16 taps + fp32 bias per lane

__device__ __forceinline__ uint32_t relu6_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  const __half2 six2 = __float2half2_rn(6.f);
  __half2 h = __hmin2(*reinterpret_cast<__half2*>(&r), six2);
  return *reinterpret_cast<uint32_t*>(&h);
}

// acc + x.half[XH] * w.half[WH] with an fp32 accumulator (sm_100 mixed-precision FMA; the fp16 x fp16 product is exact
// in fp32, i.e. bit-identical to fmaf(float(x), float(w), acc))
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

// selectors are compile-time constants after unrolling
__device__ __forceinline__ float fmah_sel(uint32_t x, int xh, uint32_t w, int wh, float acc) {
  if (xh) return wh ? fmah<1, 1>(x, w, acc) : fmah<1, 0>(x, w, acc);
  return wh ? fmah<0, 1>(x, w, acc) : fmah<0, 0>(x, w, acc);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& a) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
// MN-major operand (64 MN-contiguous elements per 128-byte line, 8 K lines per 1024-byte swizzle atom), 128B swizzle:
// LBO = byte distance between 64-element MN groups, SBO = byte distance between groups of 8 K
// (cute::UMMA::make_umma_desc<Major::MN>: LayoutType::B128 ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)))
__device__ __forceinline__ uint64_t make_sdesc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

#define CPL_STAMP(role, idx, k)                                                             \
  if (g.dbg && blockIdx.x == 0 && lane == 0 && (idx) < 64) g.dbg[(((role) * 64 + (idx)) * 8) + (k)] = clock64()

// S = stride, TH x TW = output tile (TW = 16 for stride 1, 8 for stride 2).  Warp roles (24 warps):
//   0..15  depthwise: TMEM lane group (warp & 3) = 32 channels of the chunk, row group (warp >> 2) = TH / 4 output rows
//   16     expand issuer (lane 0): expand(c) goes out as soon as its accumulator buffer has been drained
//   17     project issuer (lane 0) + weight ring refills
//   18     input-tile TMA issuer (lane 0)
//   19     patches the ones unit of landed tiles and hands them to the expand issuer
//   20..23 epilogue: one TMEM lane group each, D2 + bias (+residual) -> fp16 NHWC; the only warps that touch global
//          memory besides the loaders, so the proxy fences of the depthwise warps never wait on global traffic
template <int S, int TH>
__global__ void __launch_bounds__(CPL_THREADS, 1)
fused_block_cpl_kernel(const __grid_constant__ CUtensorMap tmX, const CplGeom g, const __half* __restrict__ in,
                       const uint8_t* __restrict__ wblob, const float* __restrict__ bp, __half* __restrict__ out) {
  constexpr int TW = S == 1 ? 16 : 8;
  constexpr int IW = (TW - 1) * S + 3;
  constexpr int IH = (TH - 1) * S + 3;
  constexpr int R = IH * IW;               // halo pixels = N of the expand MMA = TMEM columns per accumulator
  constexpr int RP = (R + 15) & ~15;
  constexpr int NPIX = TH * TW;
  constexpr int MG = NPIX > 64 ? 2 : 1;    // 64-pixel groups of the project A operand
  constexpr int NRO = TH / 4;              // output rows per warp
  constexpr int NRI = (NRO - 1) * S + 3;   // input rows per warp
  static_assert(RP <= 256 && NPIX <= 128 && TH % 4 == 0, "tile shape");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem + g.off_X;     // [NX][kb_in][RP rows][xrb B] swizzled K-major (B operand of expand)
  uint8_t* sA2 = smem + g.off_A2;   // [2][16 k-groups][MG][1024 B] swizzled MN-major (A operand of project)
  uint8_t* sWE = smem + g.off_WE;   // [NSe] expand weight images (freed when the chunk's expand has retired)
  uint8_t* sWP = smem + g.off_WP;   // [NSp] project weight images (freed when the chunk's project has retired)
  uint8_t* sDW = smem + g.off_DW;   // [n_chunks] depthwise taps + bias, resident
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.off_bars);
  uint64_t* bar_e = bars;            // [2] expand(c) retired              -> D1[c & 1] full
  uint64_t* bar_p = bars + 2;        // [2] project(c) retired             -> A2[c & 1] and the chunk's weight slot free
  uint64_t* bar_d1 = bars + 4;       // [2] D1[c & 1] drained by the depthwise warps (16 arrivals)
  uint64_t* bar_a2 = bars + 6;       // [2] A2[c & 1] written (16 arrivals)
  uint64_t* bar_d2 = bars + 8;       // [2] D2[t % ND2] drained by the epilogue warps (4 arrivals)
  uint64_t* bar_f = bars + 10;       // [2] last project of the tile retired -> D2[t % ND2] full
  uint64_t* bar_we = bars + 12;      // [8] expand weight image landed (tx)
  uint64_t* bar_x = bars + 20;       // [8] input tile patched (32 arrivals of warp 19) -> expand may read it
  uint64_t* bar_xf = bars + 28;      // [8] every expand reading the input tile has retired -> buffer free
  uint64_t* bar_t = bars + 36;       // [8] TMA of the input tile landed (tx) -> the loaders patch the ones unit
  uint64_t* bar_wp = bars + 44;      // [8] project weight image landed (tx)
  uint64_t* bar_dw = bars + 52;      // depthwise taps of every chunk landed (tx)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 53);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (g.dbg && blockIdx.x == 0 && tid == 0) g.dbg[(4 * 64 + 63) * 8 + 0] = clock64();   // kernel entry
  const int n_chunks = g.n_chunks, NSe = g.NSe, NSp = g.NSp, NX = g.NX, ND2 = g.ND2;
  const int my_tiles = (g.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int Ctot = my_tiles * n_chunks;
  const bool stream_e = NSe < n_chunks, stream_p = NSp < n_chunks;
  tc::pdl_launch_dependents();

  if (tid < 53) {   // one barrier per thread: a single thread initialising all of them costs ~100 cycles apiece
    const int i = tid;
    tc::mbar_init(&bars[i], (i >= 4 && i < 8) ? CPL_NT / 32 : (i >= 8 && i < 10) ? 4 : (i >= 20 && i < 28) ? 32 :
                                (i >= 28 && i < 36 && g.residual) ? 5 : 1);
    tc::fence_barrier_init();
  }
  if (tid == 64) tc::prefetch_tmap(&tmX);
  if (warp == 16) tc::tmem_alloc(tmem_slot, g.tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base + 2u * RP;   // D2[b] at + b * cout_pad
  if (g.dbg && blockIdx.x == 0 && tid == 0) g.dbg[(4 * 64 + 63) * 8 + 1] = clock64();   // prologue done

  // Tile cursors advance by gridDim.x tiles without divisions: (tx, ty, img) += (dx, dy, dimg) with carries.
  struct TilePos { int tx, ty, img; };
  TilePos pos0;
  int step_x, step_y, step_img;
  {
    int k = (int)blockIdx.x;
    pos0.tx = k % g.tiles_x;
    k /= g.tiles_x;
    pos0.ty = k % g.tiles_y;
    pos0.img = k / g.tiles_y;
    k = (int)gridDim.x;
    step_x = k % g.tiles_x;
    k /= g.tiles_x;
    step_y = k % g.tiles_y;
    step_img = k / g.tiles_y;
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

  // Register re-balancing between warpgroups: the 8 service warps give registers back, the depthwise warps take them.
  if (warp >= 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
  if (warp == 16) {
    // =========================================================================================== expand issuer
    if (lane == 0) {
      const uint32_t idesc_e = tc::make_idesc_f16(RP);
      auto load_we = [&](int slot, int chunk) {   // weights do not depend on the predecessor kernel
        tc::mbar_expect_tx(&bar_we[slot], g.we_bytes);
        bulk_load(sWE + (size_t)slot * g.we_bytes, wblob + (size_t)chunk * g.blob_bytes, g.we_bytes, &bar_we[slot]);
      };
      const uint64_t desc_w0 = g.xrb == 128 ? tc::make_sdesc_sw128(tc::smem_u32(sWE)) : tc::make_sdesc_sw64(tc::smem_u32(sWE));
      const uint64_t desc_x0 = g.xrb == 128 ? tc::make_sdesc_sw128(tc::smem_u32(sX)) : tc::make_sdesc_sw64(tc::smem_u32(sX));
      int t = 0, j = 0;
      for (int c = 0; c < Ctot; ++c) {
        const int slot = stream_e ? c % NSe : j;
        CPL_STAMP(1, c, 0);
        if (stream_e || c < n_chunks) tc::mbar_wait(&bar_we[slot], stream_e ? (uint32_t)((c / NSe) & 1) : 0u);
        CPL_STAMP(1, c, 1);
        if (j == 0) tc::mbar_wait(&bar_x[t % NX], (uint32_t)((t / NX) & 1));
        CPL_STAMP(1, c, 2);
        if (c >= 2) tc::mbar_wait(&bar_d1[c & 1], (uint32_t)(((c >> 1) - 1) & 1));
        CPL_STAMP(1, c, 3);
        tc::fence_after_sync();
        // descriptors: start-address field += bytes >> 4 (k-blocks of 64 channels are 128 rows x 128 B (weights) and
        // RP rows x 128 B (input tile) apart; a K step of 16 is 32 bytes inside the swizzled row)
        const uint64_t da = desc_w0 + (uint64_t)((slot * g.we_bytes) >> 4);
        const uint64_t db = desc_x0 + (uint64_t)(((t % NX) * g.x_buf_bytes) >> 4);
        const uint32_t d1 = tmem_base + (uint32_t)((c & 1) * RP);
        for (int k16 = 0; k16 < g.nk16; ++k16) {
          const int kb = k16 >> 2, k = g.xrb == 128 ? (k16 & 3) : k16;
          tc::umma_f16(d1, da + (uint64_t)(kb * (128 * 128 >> 4) + 2 * k), db + (uint64_t)(kb * (RP * 128 >> 4) + 2 * k),
                       idesc_e, k16 > 0 ? 1u : 0u);
        }
        tc::umma_commit(&bar_e[c & 1]);
        if (j == n_chunks - 1) tc::umma_commit(&bar_xf[t % NX]);   // the tile's input buffer may be refilled
        CPL_STAMP(1, c, 4);
        // the expand image of chunk c-1 is dead once that expand has retired (it has by now: the tensor pipe runs in
        // order and expand(c) sits behind it): its slot takes the image NSe chunks ahead -- long before the project of
        // chunk c-1 retires, which is what used to release the whole chunk blob
        if (stream_e && c >= 1 && c - 1 + NSe < Ctot) {
          const int k = c - 1;
          tc::mbar_wait(&bar_e[k & 1], (uint32_t)((k >> 1) & 1));
          load_we(k % NSe, (k + NSe) % n_chunks);
        }
        if (++j == n_chunks) { j = 0; ++t; }
      }
    }
  } else if (warp == 17) {
    // =========================================================================================== project issuer
    if (lane == 0) {
      auto load_wp = [&](int slot, int chunk) {
        tc::mbar_expect_tx(&bar_wp[slot], g.wp_bytes);
        bulk_load(sWP + (size_t)slot * g.wp_bytes, wblob + (size_t)chunk * g.blob_bytes + g.we_bytes, g.wp_bytes, &bar_wp[slot]);
      };
      // initial fills in the order they are needed (the copy engine serves them in order): expand image of chunk 0, the
      // depthwise taps, project image of chunk 0, then the following chunks; refills: WE by the expand issuer, WP here
      const int pre_e = stream_e ? min(NSe, Ctot) : n_chunks, pre_p = stream_p ? min(NSp, Ctot) : n_chunks;
      for (int c = 0; c < max(pre_e, pre_p); ++c) {
        if (c < pre_e) {
          tc::mbar_expect_tx(&bar_we[c], g.we_bytes);
          bulk_load(sWE + (size_t)c * g.we_bytes, wblob + (size_t)(c % n_chunks) * g.blob_bytes, g.we_bytes, &bar_we[c]);
        }
        if (c == 0) {
          tc::mbar_expect_tx(bar_dw, (uint32_t)(n_chunks * CPL_DW_BYTES));
          for (int q = 0; q < n_chunks; ++q)
            bulk_load(sDW + (size_t)q * CPL_DW_BYTES, wblob + (size_t)q * g.blob_bytes + g.we_bytes + g.wp_bytes, CPL_DW_BYTES, bar_dw);
        }
        if (c < pre_p) load_wp(c, c % n_chunks);
      }
      const uint32_t idesc_p = tc::make_idesc_f16(g.cout_pad) | (1u << 15);   // A operand MN-major
      const uint64_t desc_a0 = make_sdesc_mn_sw128(tc::smem_u32(sA2), 1024u, MG * 1024u);
      const uint64_t desc_p0 = tc::make_sdesc_sw128(tc::smem_u32(sWP));
      int t = 0, j = 0;
      for (int c = 0; c < Ctot; ++c) {
        if (stream_p && c >= 1 && c - 1 + NSp < Ctot) {   // the slot of chunk c-1 is free once project(c-1) has retired
          const int k = c - 1;
          tc::mbar_wait(&bar_p[k & 1], (uint32_t)((k >> 1) & 1));
          load_wp(k % NSp, (k + NSp) % n_chunks);
        }
        CPL_STAMP(2, c, 0);
        tc::mbar_wait(&bar_a2[c & 1], (uint32_t)((c >> 1) & 1));
        CPL_STAMP(2, c, 1);
        const int b2 = t % ND2;
        if (j == 0 && t >= ND2) tc::mbar_wait(&bar_d2[b2], (uint32_t)(((t / ND2) - 1) & 1));
        const int slot = stream_p ? c % NSp : j;
        if (stream_p || c < n_chunks) tc::mbar_wait(&bar_wp[slot], stream_p ? (uint32_t)((c / NSp) & 1) : 0u);
        CPL_STAMP(2, c, 2);
        tc::fence_after_sync();
        const uint64_t da = desc_a0 + (uint64_t)(((c & 1) * g.a2_buf_bytes) >> 4);
        const uint64_t db = desc_p0 + (uint64_t)((slot * g.wp_bytes) >> 4);
        const int valid = min(128, g.Cexp - j * 128);
        const int nk = (valid + 15) >> 4;
        const uint32_t d2 = tmem_d2 + (uint32_t)(b2 * g.cout_pad);
        for (int k = 0; k < nk; ++k)   // A: 2 k-groups of MG KB per K step; B: k-blocks of 64 channels cout_pad rows apart
          tc::umma_f16(d2, da + (uint64_t)(k * (2 * MG * 1024 >> 4)), db + (uint64_t)((k >> 2) * ((g.cout_pad * 128) >> 4) + 2 * (k & 3)),
                       idesc_p, (j > 0 || k > 0) ? 1u : 0u);
        tc::umma_commit(&bar_p[c & 1]);
        if (j == n_chunks - 1) tc::umma_commit(&bar_f[b2]);
        CPL_STAMP(2, c, 3);
        if (++j == n_chunks) { j = 0; ++t; }
      }
    }
  } else if (warp == 18) {
    // =========================================================================================== input-tile TMA issuer
    // One thread issues the TMA loads of the halo tiles (NHWC box, zero fill outside the image and beyond Cin), NX tiles
    // ahead, as soon as the ring slot has been released by the last expand that read it.
    if (lane == 0) {
      tc::pdl_wait();   // the input tensor is the predecessor's output
      TilePos tp = pos0;
      const uint32_t tx_bytes = (uint32_t)(g.kb_in * R * g.xrb);
      for (int t = 0; t < my_tiles; ++t) {
        const int xbuf = t % NX;
        CPL_STAMP(3, t, 0);
        if (t >= NX) tc::mbar_wait(&bar_xf[xbuf], (uint32_t)(((t / NX) - 1) & 1));
        CPL_STAMP(3, t, 1);
        tc::mbar_expect_tx(&bar_t[xbuf], tx_bytes);
        for (int kb = 0; kb < g.kb_in; ++kb)
          tc::tma_load_4d(sX + (size_t)xbuf * g.x_buf_bytes + (size_t)kb * RP * 128, &tmX, &bar_t[xbuf], kb * 64,
                          tp.tx * TW * S - g.pad_l, tp.ty * TH * S - g.pad_t, tp.img);
        CPL_STAMP(3, t, 2);
        advance(tp);
      }
    }
  } else if (warp == 19) {
    // =========================================================================================== ones-unit patcher
    // Once a tile has landed, this warp writes the "ones" unit of its in-image pixels (TMA zero-filled it) and hands the
    // tile to the expand issuer.  Each lane owns rows lane, lane + 32, ... of every tile: their (row, column) inside the
    // halo and their shared-memory offsets are computed once.
    constexpr int NROWS = (R + 31) / 32;
    int ry[NROWS], rx[NROWS];
    uint32_t roff[NROWS];
#pragma unroll
    for (int i = 0; i < NROWS; ++i) {
      const int r = lane + 32 * i;
      ry[i] = r / IW;
      rx[i] = r - ry[i] * IW;
      roff[i] = g.xrb == 128 ? (uint32_t)(g.ones_unit >> 3) * (uint32_t)(RP * 128) + (uint32_t)r * 128u +
                                   (uint32_t)(((g.ones_unit & 7) ^ (r & 7)) << 4)
                             : (uint32_t)r * 64u + (uint32_t)((g.ones_unit ^ ((r >> 1) & 3)) << 4);
    }
    TilePos lp = pos0;
    for (int t = 0; t < my_tiles; ++t) {
      const int xbuf = t % NX;
      const int iy0 = lp.ty * TH * S - g.pad_t, ix0 = lp.tx * TW * S - g.pad_l;
      const uint32_t xbase = tc::smem_u32(sX) + (uint32_t)xbuf * g.x_buf_bytes;
      tc::mbar_wait(&bar_t[xbuf], (uint32_t)((t / NX) & 1));
#pragma unroll
      for (int i = 0; i < NROWS; ++i) {
        if (lane + 32 * i < R) {
          const int iy = iy0 + ry[i], ix = ix0 + rx[i];
          const bool inb = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
          const uint32_t one2 = inb ? 0x3C003C00u : 0u;   // {1, 1, 0, 0, 0, 0, 0, 0} in fp16
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(xbase + roff[i]), "r"(one2), "r"(0u) : "memory");
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&bar_x[xbuf]);
      advance(lp);
    }
  } else if (warp >= 20) {
    // =========================================================================================== epilogue warps
    const int lg = warp - 20;
    const uint32_t tm_lane = (uint32_t)(lg * 32) << 16;
    // the project bias goes to shared memory once (it does not depend on the predecessor kernel); named barrier 1 is
    // private to the four epilogue warps
    float* s_bias = reinterpret_cast<float*>(smem + g.off_bias);
    for (int i = tid - 20 * 32; i < g.cout_pad; i += 128) s_bias[i] = i < g.Cout ? __ldg(bp + i) : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    tc::pdl_wait();   // output writes order after the predecessor
    TilePos ep = pos0;
    const int p = lg * 32 + lane;
    // residual: the block input at the output pixel is row (y + 1) * IW + x + 1 of the halo tile that is still in shared
    // memory (stride 1, SAME: pad 1), so it is read from there; the tile's ring slot is released after the last read
    const int rr = (p / TW + 1) * IW + (p % TW) + 1;
    for (int t = 0; t < my_tiles; ++t) {
      const int b2 = t % ND2;
      if (warp == 20) { CPL_STAMP(4, t, 0); }
      tc::mbar_wait(&bar_f[b2], (uint32_t)((t / ND2) & 1));
      if (warp == 20) { CPL_STAMP(4, t, 1); }
      tc::fence_after_sync();
      if (lg * 32 < NPIX) {
        const int oy = ep.ty * TH + p / TW, ox = ep.tx * TW + p % TW;
        const bool valid = p < NPIX && oy < g.Ho && ox < g.Wo;
        const long long opix = ((long long)ep.img * g.Ho + oy) * g.Wo + ox;
        const uint32_t taddr = tmem_d2 + tm_lane + (uint32_t)(b2 * g.cout_pad);
        const uint32_t xrow = tc::smem_u32(sX) + (uint32_t)(t % NX) * g.x_buf_bytes +
                              (g.xrb == 128 ? (uint32_t)rr * 128u : (uint32_t)rr * 64u);
        for (int cc = 0; cc < g.cout_pad; cc += 16) {
          uint32_t v[16];
          tc::tmem_ld16(taddr + (uint32_t)cc, v);
          tc::tmem_ld_wait();
          if (!valid) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = cc + 8 * h;
            if (n >= g.Cout) break;
            const float4 bb0 = *reinterpret_cast<const float4*>(s_bias + n), bb1 = *reinterpret_cast<const float4*>(s_bias + n + 4);
            float f[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += __uint_as_float(v[8 * h + i]);
            if (g.residual) {
              const int u = n >> 3;
              const uint32_t xa = g.xrb == 128 ? xrow + (uint32_t)(u >> 3) * (uint32_t)(RP * 128) + (uint32_t)(((u & 7) ^ (rr & 7)) << 4)
                                               : xrow + (uint32_t)((u ^ ((rr >> 1) & 3)) << 4);
              uint4 rq;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rq.x), "=r"(rq.y), "=r"(rq.z), "=r"(rq.w) : "r"(xa));
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
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&bar_d2[b2]);
        if (g.residual) tc::mbar_arrive(&bar_xf[t % NX]);   // the input tile has been read for the last time
      }
      if (warp == 20) { CPL_STAMP(4, t, 2); }
      advance(ep);
    }
  } else {
    // =========================================================================================== depthwise warps
    const int lg = warp & 3, rg = warp >> 2;
    const uint32_t tm_lane = (uint32_t)(lg * 32) << 16;
    int c = 0;
    tc::mbar_wait(bar_dw, 0u);   // depthwise taps of every chunk (resident)
    for (int t = 0; t < my_tiles; ++t) {
      for (int j = 0; j < n_chunks; ++j, ++c) {
        if (warp == 0) { CPL_STAMP(0, c, 0); }
        const uint32_t* dwp = reinterpret_cast<const uint32_t*>(sDW + (size_t)j * CPL_DW_BYTES) + lg * 32 + lane;
        uint32_t wv[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) wv[i] = dwp[i * 128];
        const float bias = __uint_as_float(dwp[5 * 128]);
        // A short last chunk (<= 64 channels) is replicated across the lane groups by the weight image: every replica
        // sees the same channels and takes a column slice of the tile, so the chunk costs 1/4 or 1/2 of a full one.
        const bool rem = g.rem_vp != 0 && j == n_chunks - 1;
        const bool active = rem || j * 128 + lg * 32 < g.Cexp;   // warp-uniform: this lane group holds real channels
        tc::mbar_wait(&bar_e[c & 1], (uint32_t)((c >> 1) & 1));
        if (warp == 0) { CPL_STAMP(0, c, 1); }
        tc::fence_after_sync();
        const uint32_t tbase = tmem_base + tm_lane + (uint32_t)((c & 1) * RP) + (uint32_t)(rg * NRO * S * IW);
        uint8_t* a2buf = sA2 + (size_t)(c & 1) * g.a2_buf_bytes;

        // NOUT output columns starting at x0 of this warp's NRO rows, channel chl of the chunk (A2 position)
        auto dw_cols = [&](auto nout_c, int x0, int chl) {
          constexpr int NOUT = decltype(nout_c)::value;
          constexpr int NIN = (NOUT - 1) * S + 3;          // input columns, starting at x0 * S
          constexpr int NPKN = (NIN + 1) / 2;
          constexpr int NV = NIN >= 16 ? NIN + 1 : 16;      // staging words per row (loads come in x16 / x8 pieces)
          uint32_t hrow[NRI][NPKN];
          const uint32_t tcol = tbase + (uint32_t)(x0 * S);
          // rows come out of TMEM in groups of at most 3 (register pressure: fp32 words before packing)
          constexpr int GRP = NRI <= 3 ? NRI : 2;
#pragma unroll
          for (int i0 = 0; i0 < NRI; i0 += GRP) {
            uint32_t v[GRP][NV];
#pragma unroll
            for (int i = 0; i < GRP; ++i) {
              if (i0 + i < NRI) {
                const uint32_t ta = tcol + (uint32_t)((i0 + i) * IW);
                if constexpr (NIN >= 16) {
                  uint32_t(&vr)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[i][0]);
                  tc::tmem_ld16(ta, vr);
                  if constexpr (NIN == 18) tmem_ld2(ta + 16, v[i][16], v[i][17]);
                  else tmem_ld1(ta + 16, v[i][16]);
                } else {
                  uint32_t(&vr)[8] = *reinterpret_cast<uint32_t(*)[8]>(&v[i][0]);
                  tmem_ld8(ta, vr);
                  if constexpr (NIN == 10) tmem_ld2(ta + 8, v[i][8], v[i][9]);
                  else if constexpr (NIN == 9) tmem_ld1(ta + 8, v[i][8]);
                }
              }
            }
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < GRP; ++i) {
              if (i0 + i < NRI) {
#pragma unroll
                for (int k = 0; k < NPKN; ++k) {
                  const float lo = __uint_as_float(v[i][2 * k]);
                  const float hi = 2 * k + 1 < NIN ? __uint_as_float(v[i][2 * k + 1]) : 0.f;
                  hrow[i0 + i][k] = relu6_pack(lo, hi);
                }
              }
            }
          }
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bar_d1[c & 1]);
          if (warp == 0) { CPL_STAMP(0, c, 2); }
          if (c >= 2) tc::mbar_wait(&bar_p[c & 1], (uint32_t)(((c - 2) >> 1) & 1));   // project(c-2) has read A2[c & 1]
          if (warp == 0) { CPL_STAMP(0, c, 3); }
          const int kg = chl >> 3, jj = chl & 7;
          uint8_t* a2 = a2buf + (size_t)kg * (MG * 1024) + jj * 128;
#pragma unroll
          for (int ro = 0; ro < NRO; ++ro) {
            float acc[NOUT];
#pragma unroll
            for (int x = 0; x < NOUT; ++x) acc[x] = bias;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t* hr = hrow[ro * S + ky];
#pragma unroll
              for (int x = 0; x < NOUT; ++x) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {   // tap (ky, kx) reads input column x * S + kx (relative to x0 * S)
                  const int ix = x * S + kx, tap = ky * 3 + kx;
                  acc[x] = fmah_sel(hr[ix >> 1], ix & 1, wv[tap >> 1], tap & 1, acc[x]);
                }
              }
            }
            // output row r = rg * NRO + ro of the tile: pixels p = r * TW + x0 + x live in atom p / 64, 16-byte chunk
            // (p % 64) / 8 (swizzled with the channel's line), 2 bytes per pixel
            const int p0 = (rg * NRO + ro) * TW + x0;
            uint8_t* arow = a2 + (size_t)(p0 >> 6) * 1024;
            const int pb = (p0 & 63) * 2;
            if constexpr (NOUT >= 8) {
#pragma unroll
              for (int h = 0; h < NOUT / 8; ++h) {
                uint4 o;
                o.x = relu6_pack(acc[8 * h + 0], acc[8 * h + 1]);
                o.y = relu6_pack(acc[8 * h + 2], acc[8 * h + 3]);
                o.z = relu6_pack(acc[8 * h + 4], acc[8 * h + 5]);
                o.w = relu6_pack(acc[8 * h + 6], acc[8 * h + 7]);
                *reinterpret_cast<uint4*>(arow + ((((pb >> 4) + h) ^ jj) << 4)) = o;
              }
            } else if constexpr (NOUT == 4) {
              uint2 o;
              o.x = relu6_pack(acc[0], acc[1]);
              o.y = relu6_pack(acc[2], acc[3]);
              *reinterpret_cast<uint2*>(arow + (((pb >> 4) ^ jj) << 4) + (pb & 15)) = o;
            } else {
              *reinterpret_cast<uint32_t*>(arow + (((pb >> 4) ^ jj) << 4) + (pb & 15)) = relu6_pack(acc[0], acc[1]);
            }
          }
        };
        if (!active) {
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bar_d1[c & 1]);
          if (c >= 2) tc::mbar_wait(&bar_p[c & 1], (uint32_t)(((c - 2) >> 1) & 1));
        } else if (!rem) {
          dw_cols(std::integral_constant<int, TW>{}, 0, lg * 32 + lane);
        } else if (g.rem_vp == 32) {
          dw_cols(std::integral_constant<int, TW / 4>{}, lg * (TW / 4), lane);
        } else {
          dw_cols(std::integral_constant<int, TW / 2>{}, (lg >> 1) * (TW / 2), (lg & 1) * 32 + lane);
        }
        if (warp == 0) { CPL_STAMP(0, c, 4); }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_a2[c & 1]);
        if (warp == 0) { CPL_STAMP(0, c, 5); }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (g.dbg && blockIdx.x == 0 && tid == 0) g.dbg[(4 * 64 + 63) * 8 + 2] = clock64();   // every role done
  if (warp == 16) tc::tmem_dealloc(tmem_base, g.tmem_cols);
}

// Builds the per-chunk weight images: WE (A operand of expand: 128 channel rows x K, K-major, swizzled, bias in the
// "ones" column), WP (B operand of project: cout_pad rows x 128 channels as two 64-channel k-blocks, K-major, 128B
// swizzle), depthwise taps (5 words of packed fp16 pairs per lane) + depthwise bias.
__global__ void cpl_pack_kernel(CplGeom g, const __half* __restrict__ we, int we_ld, const float* __restrict__ be,
                                const __half* __restrict__ wp, int wp_ld, const float* __restrict__ wd,
                                const float* __restrict__ bd, uint8_t* __restrict__ blob) {
  const int K = g.units * 8;
  const long long per_chunk = 128LL * K + (long long)g.cout_pad * 128 + 6 * 128;
  const long long total = per_chunk * g.n_chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / per_chunk);
    long long e = i - (long long)j * per_chunk;
    uint8_t* base = blob + (size_t)j * g.blob_bytes;
    const bool repl = g.rem_vp != 0 && j == g.n_chunks - 1;   // lane l of the short last chunk holds channel l % rem_vp
    if (e < 128LL * K) {
      const int r = (int)(e / K), k = (int)(e % K);
      const int ch = repl ? j * 128 + (r % g.rem_vp) : j * 128 + r;
      __half v = __float2half_rn(0.f);
      if (ch < g.Cexp) {
        if (k < g.Cin) v = we[(size_t)ch * we_ld + k];
        else if (k == g.Cin) v = __float2half_rn(be[ch]);                                      // bias, high part
        else if (k == g.Cin + 1) v = __float2half_rn(be[ch] - __half2float(__float2half_rn(be[ch])));   // low part
      }
      size_t off;
      if (g.xrb == 128) {
        const int kb = k >> 6, kk = k & 63;
        off = (size_t)kb * 128 * 128 + (size_t)r * 128 + (size_t)(((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2;
      } else {
        off = (size_t)r * 64 + (size_t)(((k >> 3) ^ ((r >> 1) & 3)) << 4) + (k & 7) * 2;
      }
      *reinterpret_cast<__half*>(base + off) = v;
      continue;
    }
    e -= 128LL * K;
    if (e < (long long)g.cout_pad * 128) {
      const int n = (int)(e / 128), k = (int)(e % 128);
      const int ch = j * 128 + k;
      __half v = __float2half_rn(0.f);
      if (n < g.Cout && ch < g.Cexp) v = wp[(size_t)n * wp_ld + ch];
      const int kb = k >> 6, kk = k & 63;
      const size_t off = (size_t)kb * g.cout_pad * 128 + (size_t)n * 128 + (size_t)(((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
      *reinterpret_cast<__half*>(base + g.we_bytes + off) = v;
      continue;
    }
    e -= (long long)g.cout_pad * 128;
    const int word = (int)(e / 128), l = (int)(e % 128);
    const int ch = repl ? j * 128 + (l % g.rem_vp) : j * 128 + l;
    uint32_t v = 0;
    if (ch < g.Cexp) {
      if (word < 5) {
        const __half lo = __float2half_rn(wd[(size_t)(2 * word) * g.Cexp + ch]);
        const __half hi = 2 * word + 1 < 9 ? __float2half_rn(wd[(size_t)(2 * word + 1) * g.Cexp + ch]) : __float2half_rn(0.f);
        v = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
      } else {
        v = __float_as_uint(bd[ch]);
      }
    }
    reinterpret_cast<uint32_t*>(base + g.we_bytes + g.wp_bytes)[word * 128 + l] = v;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
struct CplPlan {
  CplGeom g;          // everything that does not depend on the tile shape / batch
  uint8_t* d_blob = nullptr;
  CUtensorMap tmX[2]; // input halo tile boxes: [0] tall tile (TH = 8), [1] TH = 4
  int pin_th = 0, pin_ns = 0, pin_nx = 0;
};

int hfb_make_tmap_nhwc_box(hfb_ctx* ctx, CUtensorMap* out, const void* base, int C, int W, int H, int B, int box_w,
                           int box_h, int box_c);

CplPlan* cpl_new() { return new CplPlan(); }
void cpl_delete(CplPlan* p) { delete p; }

static bool cpl_layout(CplGeom& g, int S, int TH, int NSe, int NSp, int NX) {
  const int TW = S == 1 ? 16 : 8;
  const int IW = (TW - 1) * S + 3, IH = (TH - 1) * S + 3;
  const int R = IH * IW, RP = (R + 15) & ~15;
  const int NPIX = TH * TW, MG = NPIX > 64 ? 2 : 1;
  if (RP > 256 || 2 * RP + g.cout_pad > 512) return false;
  auto al = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  g.NSe = NSe;
  g.NSp = NSp;
  g.NX = NX;
  g.x_buf_bytes = al((uint32_t)(g.kb_in * RP * g.xrb));
  g.a2_buf_bytes = (uint32_t)(16 * MG * 1024);
  uint32_t off = 0;
  g.off_X = off;  off += (uint32_t)NX * g.x_buf_bytes;
  g.off_A2 = off; off += 2 * g.a2_buf_bytes;
  g.off_WE = off; off += al((uint32_t)NSe * g.we_bytes);
  g.off_WP = off; off += al((uint32_t)NSp * g.wp_bytes);
  g.off_DW = off; off += al((uint32_t)g.n_chunks * CPL_DW_BYTES);
  g.off_bars = off; off += 512;
  g.off_bias = off; off += 1024;   // project bias, fp32 [cout_pad <= 256]
  g.smem_bytes = off + 1024;
  g.ND2 = 2 * RP + 2 * g.cout_pad <= 512 ? 2 : 1;
  uint32_t cols = 32;
  while ((int)cols < 2 * RP + g.ND2 * g.cout_pad) cols <<= 1;
  g.tmem_cols = cols;
  return g.smem_bytes <= 227u * 1024u;
}

// Returns HFB_ERR_CAPACITY when the block cannot run in this formulation (the caller keeps another path).
int cpl_plan(hfb_ctx* ctx, CplPlan& cp, const BlockW& bw, const __half* in, int Bmax, int Hi, int Wi, int Ho, int Wo,
             int pad_t, int pad_l) {
  if (!bw.has_expand || bw.cin % 8 != 0 || bw.cout % 8 != 0) return HFB_ERR_CAPACITY;
  CplGeom& g = cp.g;
  memset(&g, 0, sizeof(g));
  g.Hi = Hi; g.Wi = Wi; g.Ho = Ho; g.Wo = Wo;
  g.Cin = bw.cin; g.Cexp = bw.cexp; g.Cout = bw.cout;
  g.pad_t = pad_t; g.pad_l = pad_l;
  g.residual = bw.residual ? 1 : 0;
  g.n_chunks = (bw.cexp + 127) / 128;
  {
    const int valid = bw.cexp - 128 * (g.n_chunks - 1);
    static const bool no_rep = getenv("HFB_CPL_NOREP") != nullptr;   // experiments: plain (unreplicated) last chunk
    g.rem_vp = no_rep ? 0 : (valid <= 32 ? 32 : (valid <= 64 ? 64 : 0));
  }
  g.ones_unit = bw.cin / 8;
  g.units = (g.ones_unit + 1 + 1) & ~1;        // K = units * 8, multiple of 16
  if (g.units * 8 > 128) return HFB_ERR_CAPACITY;
  g.xrb = g.units <= 4 ? 64 : 128;
  g.kb_in = g.xrb == 64 ? 1 : (g.units + 7) / 8;
  g.nk16 = g.units / 2;
  g.cout_pad = (bw.cout + 15) & ~15;
  if (g.cout_pad > 256) return HFB_ERR_CAPACITY;
  g.we_bytes = (uint32_t)(g.kb_in * 128 * g.xrb);
  g.wp_bytes = (uint32_t)(2 * g.cout_pad * 128);
  g.blob_bytes = g.we_bytes + g.wp_bytes + CPL_DW_BYTES;
  auto env_int = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
  cp.pin_th = env_int("HFB_CPL_TH");
  cp.pin_ns = env_int("HFB_CPL_NS");
  cp.pin_nx = env_int("HFB_CPL_NX");
  // at least one configuration must fit
  bool ok = false;
  for (int th : {8, 4})
    for (int ns : {g.n_chunks, 2, 1})
      if (cpl_layout(g, bw.stride, bw.stride == 2 ? 4 : th, std::min(ns, g.n_chunks), std::min(ns, g.n_chunks), 1)) ok = true;
  if (!ok) return HFB_ERR_CAPACITY;
  {
    const int S = bw.stride, TW = S == 1 ? 16 : 8, IW = (TW - 1) * S + 3;
    const int box_c = g.xrb == 128 ? 64 : 32;
    if (S == 1) HFB_TRY(hfb_make_tmap_nhwc_box(ctx, &cp.tmX[0], in, bw.cin, Wi, Hi, Bmax, IW, 7 * S + 3, box_c));
    HFB_TRY(hfb_make_tmap_nhwc_box(ctx, &cp.tmX[1], in, bw.cin, Wi, Hi, Bmax, IW, 3 * S + 3, box_c));
  }
  HFB_TRY(ctx->dalloc(&cp.d_blob, (size_t)g.n_chunks * g.blob_bytes));
  HFB_CUDA(ctx, cudaMemsetAsync(cp.d_blob, 0, (size_t)g.n_chunks * g.blob_bytes, ctx->stream));
  cpl_pack_kernel<<<64, 256, 0, ctx->stream>>>(g, bw.expand.w, bw.expand.Kp, bw.expand.b, bw.project.w, bw.project.Kp,
                                               bw.wd, bw.bd, cp.d_blob);
  HFB_CHECK_LAUNCH(ctx, "cpl_pack");
  return HFB_OK;
}

// Tile shape / ring depth for a batch: stride 2 -> 4 x 8 tiles; stride 1 -> 8 x 16 when that still gives most SMs a
// tile (or the weights would not stay resident otherwise), else 4 x 16.  Weights resident when they fit, else the
// deepest ring that fits; NX = 3 for single-chunk layers (the next expand must not wait for its input tile), else 2,
// 1 when every CTA has a single tile.
static bool cpl_configure(const hfb_ctx* ctx, const CplPlan& cp, int stride, int B, CplGeom& g, int& TH) {
  g = cp.g;
  g.B = B;
  const int TW = stride == 1 ? 16 : 8;
  g.tiles_x = (g.Wo + TW - 1) / TW;
  auto try_cfg = [&](int th, bool need_resident) -> bool {
    g.tiles_y = (g.Ho + th - 1) / th;
    g.total_tiles = g.tiles_x * g.tiles_y * B;
    const int per_cta = (g.total_tiles + ctx->n_sm - 1) / ctx->n_sm;
    // input tiles in flight: a halo tile is many short rows (32..240 B per pixel), which TMA moves slowly (thousands
    // of cycles per tile), so tiles with few chunks need a deep ring to cover it
    int nx_want = std::min(std::min(per_cta, 8), std::max(2, 6 / g.n_chunks + 1));
    if (cp.pin_nx) nx_want = cp.pin_nx;
    // resident weights first, then the deepest weight rings (expand / project) and input ring that fit
    const int rings[6][2] = {{g.n_chunks, g.n_chunks}, {3, 3}, {3, 2}, {2, 2}, {2, 1}, {1, 1}};
    for (const auto& r : rings) {
      const int nse = r[0], nsp = r[1];
      if (nse > g.n_chunks || nsp > g.n_chunks || nse > 8) continue;
      if ((nse == g.n_chunks) != (nsp == g.n_chunks)) continue;
      if (cp.pin_ns && nse != std::min(cp.pin_ns, g.n_chunks)) continue;
      if (need_resident && nse < g.n_chunks) continue;
      for (int nx = nx_want; nx >= std::min(per_cta, 2); --nx)
        if (cpl_layout(g, stride, th, nse, nsp, nx)) return true;
    }
    return false;
  };
  if (stride == 2) {
    TH = 4;
    return try_cfg(4, false);
  }
  if (cp.pin_th) {
    TH = cp.pin_th;
    return try_cfg(TH, false);
  }
  const int tiles8 = g.tiles_x * ((g.Ho + 7) / 8) * B;
  const bool prefer8 = tiles8 * 10 >= ctx->n_sm * 6;
  for (bool need_res : {true, false}) {
    for (int th : {prefer8 ? 8 : 4, prefer8 ? 4 : 8}) {
      if (try_cfg(th, need_res)) {
        TH = th;
        return true;
      }
    }
  }
  return false;
}

int cpl_tiles(const hfb_ctx* ctx, const CplPlan& cp, int stride, int B) {
  CplGeom g;
  int th;
  if (!cpl_configure(ctx, cp, stride, B, g, th)) return 0;
  return g.total_tiles;
}

template <int S, int TH>
static int cpl_launch(hfb_ctx* ctx, const CplPlan& cp, const CplGeom& g, const BlockW& bw, const __half* in, __half* out) {
  static SmemOptIn optin;   // per instantiation
  HFB_CUDA(ctx, optin.ensure(fused_block_cpl_kernel<S, TH>, ctx->device, g.smem_bytes));
  const int grid = std::min(g.total_tiles, ctx->n_sm);
  hfb_launch(ctx, fused_block_cpl_kernel<S, TH>, grid, CPL_THREADS, g.smem_bytes, cp.tmX[TH == 8 ? 0 : 1], g, in,
             (const uint8_t*)cp.d_blob, bw.project.b, out);
  HFB_CHECK_LAUNCH(ctx, "fused_block_cpl");
  return HFB_OK;
}

int cpl_run(hfb_ctx* ctx, const CplPlan& cp, const BlockW& bw, const __half* in, __half* out, int B) {
  CplGeom g;
  int TH = 0;
  if (!cpl_configure(ctx, cp, bw.stride, B, g, TH)) {
    ctx->set_error("internal: fused block (cpl) has no configuration for this batch");
    return HFB_ERR_STATE;
  }
  static bool traced[32] = {false};
  if (ctx->trace && bw.layer < 32 && !traced[bw.layer]) {
    traced[bw.layer] = true;
    fprintf(stderr, "hfnet_b200: cpl layer_%d: S=%d TH=%d chunks=%d units=%d xrb=%d NSe=%d NSp=%d NX=%d smem=%u tmem=%u tiles=%d\n",
            bw.layer, bw.stride, TH, g.n_chunks, g.units, g.xrb, g.NSe, g.NSp, g.NX, g.smem_bytes, g.tmem_cols, g.total_tiles);
  }
  g.dbg = nullptr;
  static const int dbg_layer = getenv("HFB_CPL_DBG") ? atoi(getenv("HFB_CPL_DBG")) : 0;   // needs HFB_NO_GRAPH=1
  static int dbg_runs = 0;
  static long long* d_dbg = nullptr;
  const bool dbg = dbg_layer == bw.layer && dbg_runs < 2;
  if (dbg) {
    ++dbg_runs;
    if (!d_dbg) HFB_CUDA(ctx, cudaMalloc(&d_dbg, 5 * 64 * 8 * 8));
    HFB_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, 5 * 64 * 8 * 8, ctx->stream));
    g.dbg = d_dbg;
  }
  int rc = HFB_ERR_STATE;
  if (bw.stride == 1 && TH == 8) rc = cpl_launch<1, 8>(ctx, cp, g, bw, in, out);
  else if (bw.stride == 1 && TH == 4) rc = cpl_launch<1, 4>(ctx, cp, g, bw, in, out);
  else if (bw.stride == 2 && TH == 4) rc = cpl_launch<2, 4>(ctx, cp, g, bw, in, out);
  else ctx->set_error("internal: no fused block (cpl) kernel for this tile shape");
  if (rc == HFB_OK && dbg) {
    std::vector<long long> h(5 * 64 * 8);
    HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    HFB_CUDA(ctx, cudaMemcpy(h.data(), d_dbg, h.size() * 8, cudaMemcpyDeviceToHost));
    const char* roles[5] = {"depthwise", "expand", "project", "loader", "epilogue"};
    long long t0 = 0;
    for (int r = 0; r < 5; ++r)
      for (int i = 0; i < 64; ++i)
        for (int k = 0; k < 8; ++k) {
          const long long v = h[(r * 64 + i) * 8 + k];
          if (v && (!t0 || v < t0)) t0 = v;
        }
    fprintf(stderr, "hfnet_b200: cpl layer_%d CTA 0: entry %lld prologue done %lld exit %lld\n", bw.layer,
            h[(4 * 64 + 63) * 8 + 0] - t0, h[(4 * 64 + 63) * 8 + 1] - t0, h[(4 * 64 + 63) * 8 + 2] - t0);
    for (int r = 0; r < 5; ++r) {
      fprintf(stderr, "hfnet_b200: cpl layer_%d %s stamps (cycles since first):\n", bw.layer, roles[r]);
      for (int i = 0; i < 14; ++i) {
        fprintf(stderr, "   [%2d]", i);
        for (int k = 0; k < 6; ++k) {
          const long long v = h[(r * 64 + i) * 8 + k];
          fprintf(stderr, " %8lld", v ? v - t0 : -1LL);
        }
        fprintf(stderr, "\n");
      }
    }
  }
  return rc;
}

double cpl_bytes(const CplPlan& cp, int B) {   // algorithmic: input once + output once (+ residual re-read) + weights
  const CplGeom& g = cp.g;
  return 2.0 * B * ((double)g.Hi * g.Wi * g.Cin + (double)g.Ho * g.Wo * g.Cout * (g.residual ? 2 : 1)) +
         2.0 * ((double)g.Cin * g.Cexp + (double)g.Cexp * g.Cout) + 4.0 * 10 * g.Cexp;
}
double cpl_flops(const CplPlan& cp, int B) {
  const CplGeom& g = cp.g;
  return 2.0 * B * ((double)g.Hi * g.Wi * g.Cin * g.Cexp + (double)g.Ho * g.Wo * g.Cexp * (9 + g.Cout));
}
