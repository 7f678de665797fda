// The one tensor-core main loop of this library (sm_100a): C[M][N] = A[M][K] * W[N][K]^T with fp16 operands staged by
// TMA into 128B-swizzled shared-memory tiles, tcgen05.mma (UMMA 128 x BN x 16) accumulating fp32 in TMEM, and a
// pluggable epilogue that reads the accumulator tile back with tcgen05.ld (thread <-> tile row).
//
// A comes either from a 2-D map [rows][K] (1x1 convolutions, descriptor matrices) or, for the 3x3 SAME stride-1 head
// convolutions, from a 4-D map (C, W, H, B) of an NHWC tensor: one tile row = one pixel of an 8 x 16 patch, one
// k-block = (tap, 64-channel slice), the tap shift is a TMA coordinate offset and TMA's out-of-bounds zero fill is the
// SAME padding (hfnet/models/hf_net.py:63,70).
//
// Persistent, warp-specialised CTA of 192 threads that walks tiles blockIdx.x, +gridDim.x, ...:
//   warp 0      TMA producer (one elected lane) feeding a `stages`-deep smem ring that runs ahead across tiles
//   warp 1      UMMA issuer (one lane) + TMEM owner; two accumulator stages of BN columns ping-pong, so the
//               epilogue of tile i overlaps the loads and MMAs of tile i+1
//   warps 2..5  epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (the hardware's lane-group rule) = 32 tile rows
#pragma once
#include "tc.cuh"

struct GemmGeom {
  int M, N, K;        // logical sizes; conv: K = channels per tap
  int BN;             // N tile: multiple of 16, <= 256
  int a_k_off;        // first K coordinate inside A's map
  int conv;           // 1 = implicit 3x3
  int halo;           // conv: 1 = A comes as 10-row halo boxes, one per (64-channel block, dx); the three dy taps of a box are
                      // sub-views 16 pixel rows (= 2 KB = two swizzle atoms) apart -- 2.4x less A traffic than a box per tap
  int H, W;           // conv: image size; tiles are 8 rows x 16 cols
  int tiles_x, tiles_y;
  int kb_per_row;     // k-blocks per tap (conv) / in total (plain) = ceil(K/64)
  int num_kb;         // plain: kb_per_row, conv: 9 * kb_per_row
  int stages;         // smem ring depth (<= 8)
  int m_tiles, n_tiles, total_tiles;  // per problem (pair) m/n tiling; total over all pairs
  uint32_t idesc;
  uint32_t tmem_cols;   // power of two >= max(32, 2*BN)
  uint32_t ring_bytes;  // operand ring size (multiple of 1024)
  uint32_t epi_warp_bytes;  // per-epilogue-warp staging bytes behind the ring
  uint32_t bias_bytes;      // staged bias vector behind the warp staging areas (0 = none), multiple of 16
  // Batched independent problems (descriptor matching): pair_tab = device int[4][n_pairs] holding
  // a_off | a_cnt | b_off | b_cnt (row ranges inside A's and W's maps).  null = one problem of M x N.
  const int* pair_tab;
  int n_pairs;
  // Split-precision contraction (descriptor matching): operand rows are stored once as [hi(256) | lo(256)] fp16 and the
  // K' = 768 loop reads A's k-blocks as hi|hi|lo and B's as hi|lo|hi, i.e. ah.bh + ah.bl + al.bh.
  int split3;
};

struct TileRow {
  bool valid;
  long long row;      // output row (pixel index / global A row) of this thread
  int n0;             // first output column of the tile (problem-local)
  uint32_t taddr;     // TMEM address of (this warp's lane base, accumulator column 0)
  int row_local;      // row inside the problem (== row when not batched)
  int n_cnt;          // columns of the problem (== N when not batched)
  int a_off, b_off;   // batched: first global row of the pair's A / B block
  uint8_t* stage;     // this warp's private staging area (epi_warp_bytes)
  int ewarp;          // TMEM lane group 0..3 of this warp
  int sub;            // 0 .. kWarps/4-1: which of the warps sharing the lane group (they split the columns)
  const float* s_bias;  // bias vector staged in shared memory (indexed by problem column), or null
};

#define GEMM_TILE_A_BYTES 16384  // 128 rows x 128 B
#define GEMM_THREADS(EW) (64 + 32 * (EW))
#define GEMM_EPI_BAR 1           // named barrier of the 128 epilogue threads

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(GEMM_EPI_BAR), "n"(128) : "memory"); }
template <int kThreads>
__device__ __forceinline__ void epi_bar_sync_n() { asm volatile("bar.sync %0, %1;" ::"n"(GEMM_EPI_BAR), "n"(kThreads) : "memory"); }

struct TileCoord {
  bool valid;
  int m0, n0, img, y0, x0, a_off, a_cnt, b_off, b_cnt, mt;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmGeom& g, int tile) {
  TileCoord t;
  const int per = g.m_tiles * g.n_tiles;
  const int pr = tile / per;
  const int rem = tile - pr * per;
  t.mt = rem / g.n_tiles;
  const int nt = rem - t.mt * g.n_tiles;
  t.n0 = nt * g.BN;
  t.a_off = 0; t.a_cnt = g.M; t.b_off = 0; t.b_cnt = g.N;
  t.valid = true;
  if (g.pair_tab) {
    t.a_off = g.pair_tab[pr];
    t.a_cnt = g.pair_tab[g.n_pairs + pr];
    t.b_off = g.pair_tab[2 * g.n_pairs + pr];
    t.b_cnt = g.pair_tab[3 * g.n_pairs + pr];
    t.valid = (t.mt * 128 < t.a_cnt) && (t.n0 < t.b_cnt);
  }
  t.m0 = 0; t.img = 0; t.y0 = 0; t.x0 = 0;
  if (g.conv) {
    int q = t.mt;
    const int tx = q % g.tiles_x;
    q /= g.tiles_x;
    const int ty = q % g.tiles_y;
    t.img = q / g.tiles_y;
    t.y0 = ty * 8;
    t.x0 = tx * 16;
  } else {
    t.m0 = t.a_off + t.mt * 128;
  }
  return t;
}

#define GEMM_HALO_A_BYTES 20480   // 10 x 16 pixels x 128 B
static inline size_t gemm_ring_bytes(int BN, int stages, int halo = 0) {
  if (halo) return 2 * (size_t)GEMM_HALO_A_BYTES + (size_t)stages * ((size_t)BN * 128);   // two A boxes + a ring of B tiles
  return (size_t)stages * (GEMM_TILE_A_BYTES + (size_t)BN * 128);
}
static inline size_t gemm_smem_bytes(int BN, int stages, size_t epi_bytes, int halo = 0) {
  return 1024 + gemm_ring_bytes(BN, stages, halo) + epi_bytes + 256;
}

template <class Epi>
__global__ void __launch_bounds__(GEMM_THREADS(Epi::kWarps)) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const GemmGeom g, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic: keeps the shared address space visible to the compiler (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t stage_bytes = GEMM_TILE_A_BYTES + (uint32_t)g.BN * 128u;
  uint8_t* epi_smem = smem + g.ring_bytes;
  float* s_bias = reinterpret_cast<float*>(epi_smem + (size_t)Epi::kWarps * g.epi_warp_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_bias) + g.bias_bytes);
  uint64_t* full = bars;            // [8]
  uint64_t* empty = bars + 8;       // [8]
  uint64_t* acc_full = bars + 16;   // [2]
  uint64_t* acc_empty = bars + 18;  // [2]
  uint64_t* a_full = bars + 20;     // [2] halo mode: the two A boxes
  uint64_t* a_empty = bars + 22;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tc::pdl_launch_dependents();

  if (tid == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < g.stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], Epi::kWarps);   // one arrival per epilogue warp
      tc::mbar_init(&a_full[s], 1);
      tc::mbar_init(&a_empty[s], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, g.tmem_cols);
  if (g.bias_bytes) {
    const float* gb = Epi::bias(ep);
    for (int i = tid; i < g.N; i += blockDim.x) s_bias[i] = __ldg(gb + i);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  tc::pdl_wait();   // prologue done (barriers, TMEM, bias = weights only); operands / residuals come from predecessors
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && g.halo) {
    // ------------------------------------------------------------------ TMA producer, halo boxes
    // Groups (64-channel block cb, dx) in order, flattened over this CTA's tiles.  Per group one A box (64 ch x 16 x 10
    // pixels from (x0 + dx - 1, y0 - 1): out-of-range pixels / channels arrive as zeros = SAME padding) and three B
    // tiles (taps dy = 0..2).  Issue order B(g,0) B(g,1) A(g+1) B(g,2): each load is issued when its slot frees up.
    if (lane == 0) {
      const int groups = 3 * g.kb_per_row;
      uint8_t* sB = smem + 2 * GEMM_HALO_A_BYTES;
      const uint32_t b_bytes = (uint32_t)g.BN * 128u;
      uint32_t ia = 0, ib = 0;
      auto load_a = [&](int tile, int grp) {
        const TileCoord t = decode_tile(g, tile);
        const int cb = grp / 3, dx = grp - cb * 3;
        const uint32_t sl = ia & 1u;
        tc::mbar_wait(&a_empty[sl], ((ia >> 1) & 1u) ^ 1u);
        tc::mbar_expect_tx(&a_full[sl], GEMM_HALO_A_BYTES);
        tc::tma_load_4d(smem + (size_t)sl * GEMM_HALO_A_BYTES, &tmA, &a_full[sl], cb * 64, t.x0 + dx - 1, t.y0 - 1, t.img);
        ++ia;
      };
      auto load_b = [&](const TileCoord& t, int grp, int dy) {
        const int cb = grp / 3, dx = grp - cb * 3;
        const int s = ib % g.stages;
        tc::mbar_wait(&empty[s], ((ib / g.stages) & 1u) ^ 1u);
        tc::mbar_expect_tx(&full[s], b_bytes);
        tc::tma_load_2d(sB + (size_t)s * b_bytes, &tmB, &full[s], (dy * 3 + dx) * g.K + cb * 64, t.n0);
        ++ib;
      };
      if ((int)blockIdx.x < g.total_tiles) load_a(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(g, tile);
        for (int grp = 0; grp < groups; ++grp) {
          load_b(t, grp, 0);
          load_b(t, grp, 1);
          if (grp + 1 < groups) load_a(tile, grp + 1);
          else if (tile + (int)gridDim.x < g.total_tiles) load_a(tile + gridDim.x, 0);
          load_b(t, grp, 2);
        }
      }
    }
  } else if (warp == 1 && g.halo) {
    // ------------------------------------------------------------------ UMMA issuer, halo boxes
    if (lane == 0) {
      const int groups = 3 * g.kb_per_row;
      const uint32_t sA0 = tc::smem_u32(smem), sB0 = sA0 + 2 * GEMM_HALO_A_BYTES;
      const uint32_t b_bytes = (uint32_t)g.BN * 128u;
      uint32_t ia = 0, ib = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        tc::mbar_wait(&acc_empty[as], aph ^ 1u);       // epilogue has drained this accumulator stage
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)g.BN;
        for (int grp = 0; grp < groups; ++grp, ++ia) {
          const int cb = grp / 3;
          const int krem = g.K - cb * 64;
          const int nk = krem >= 64 ? 4 : (krem + 15) >> 4;     // zero-filled channel tail needs no MMA
          const uint32_t sl = ia & 1u;
          tc::mbar_wait(&a_full[sl], (ia >> 1) & 1u);
          for (int dy = 0; dy < 3; ++dy, ++ib) {
            const int s = ib % g.stages;
            tc::mbar_wait(&full[s], (ib / g.stages) & 1u);
            tc::fence_after_sync();
            // tap dy = rows dy*16 .. dy*16 + 127 of the box: 2 KB = two whole swizzle atoms further
            const uint64_t da = tc::make_sdesc_sw128(sA0 + sl * GEMM_HALO_A_BYTES + (uint32_t)dy * 2048u);
            const uint64_t db = tc::make_sdesc_sw128(sB0 + (uint32_t)s * b_bytes);
            for (int k = 0; k < nk; ++k)
              tc::umma_f16(d_tmem, tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k), g.idesc,
                           (grp > 0 || dy > 0 || k > 0) ? 1u : 0u);
            tc::umma_commit(&empty[s]);
          }
          tc::umma_commit(&a_empty[sl]);                          // the box is free once all three taps have retired
        }
        tc::umma_commit(&acc_full[as]);
        ++acc_it;
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(g, tile);
        if (!t.valid) continue;
        for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
          const int s = it % g.stages;
          const uint32_t ph = (it / g.stages) & 1u;
          tc::mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          uint8_t* sb = sa + GEMM_TILE_A_BYTES;
          tc::mbar_expect_tx(&full[s], stage_bytes);
          if (g.conv) {
            const int tap = kb / g.kb_per_row, cb = kb - tap * g.kb_per_row;
            const int dy = tap / 3, dx = tap - dy * 3;
            tc::tma_load_4d(sa, &tmA, &full[s], cb * 64, t.x0 + dx - 1, t.y0 + dy - 1, t.img);
            tc::tma_load_2d(sb, &tmB, &full[s], tap * g.K + cb * 64, t.n0);
          } else {
            int ka = kb, kbb = kb;
            if (g.split3) {
              ka = kb < 8 ? (kb & 3) : 4 + (kb & 3);
              kbb = kb < 8 ? kb : kb - 8;
            }
            tc::tma_load_2d(sa, &tmA, &full[s], g.a_k_off + ka * 64, t.m0);
            tc::tma_load_2d(sb, &tmB, &full[s], kbb * 64, t.b_off + t.n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(g, tile);
        if (!t.valid) continue;
        const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
        tc::mbar_wait(&acc_empty[as], aph ^ 1u);       // epilogue has drained this accumulator stage
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)g.BN;
        for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
          const int s = it % g.stages;
          const uint32_t ph = (it / g.stages) & 1u;
          tc::mbar_wait(&full[s], ph);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t da = tc::make_sdesc_sw128(sa);
          const uint64_t db = tc::make_sdesc_sw128(sa + GEMM_TILE_A_BYTES);
          const int cb = kb % g.kb_per_row;
          const int krem = g.K - cb * 64;                       // valid K elements in this block
          const int nk = krem >= 64 ? 4 : (krem + 15) >> 4;     // zero-filled tail needs no MMA
          for (int k = 0; k < nk; ++k)
            tc::umma_f16(d_tmem, tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k), g.idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          tc::umma_commit(&empty[s]);                            // frees the smem slot when these MMAs retire
        }
        tc::umma_commit(&acc_full[as]);
        ++acc_it;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (2 ..)
    const int ewarp = warp & 3;   // TMEM lane group this warp may access
    const int eidx = warp - 2;    // 0 .. kWarps-1
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(g, tile);
      if (!t.valid) continue;
      const uint32_t as = acc_it & 1u, aph = (acc_it >> 1) & 1u;
      tc::mbar_wait(&acc_full[as], aph);
      __syncwarp();
      tc::fence_after_sync();
      TileRow tr;
      tr.n0 = t.n0;
      tr.taddr = tmem_base + ((uint32_t)(ewarp * 32) << 16) + as * (uint32_t)g.BN;
      const int r = ewarp * 32 + lane;
      if (g.conv) {
        const int y = t.y0 + (r >> 4), x = t.x0 + (r & 15);
        tr.valid = (y < g.H) && (x < g.W);
        tr.row = ((long long)t.img * g.H + y) * g.W + x;
        tr.row_local = (int)tr.row;
      } else {
        tr.row_local = t.mt * 128 + r;
        tr.row = (long long)t.m0 + r;
        tr.valid = tr.row_local < t.a_cnt;
      }
      tr.n_cnt = t.b_cnt;
      tr.a_off = t.a_off;
      tr.b_off = t.b_off;
      tr.stage = epi_smem + (size_t)eidx * g.epi_warp_bytes;
      tr.ewarp = ewarp;
      tr.sub = eidx >> 2;
      tr.s_bias = g.bias_bytes ? s_bias : nullptr;
      Epi::run(ep, g, tr);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[as]);
      ++acc_it;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, g.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host-side launch
struct hfb_ctx;
int hfb_make_tmap_2d(hfb_ctx* ctx, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t row_stride_bytes, uint32_t box_outer);
int hfb_make_tmap_2d_f32(hfb_ctx* ctx, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                         uint64_t row_stride_bytes, uint32_t box_outer);
int hfb_make_tmap_nhwc(hfb_ctx* ctx, CUtensorMap* out, const void* base, int C, int W, int H, int B);
int hfb_conv_halo();   // 1 (default): 3x3 convolutions load 10-row halo boxes (HFB_CONV_HALO=0: one 8-row box per tap)

static inline uint32_t tmem_cols_for(int BN) {
  uint32_t c = 32;
  while ((int)c < 2 * BN) c <<= 1;
  return c;
}

static inline void gemm_finish_geom(GemmGeom& g, int m_tiles) {
  g.m_tiles = m_tiles;
  g.n_tiles = (g.N + g.BN - 1) / g.BN;
  g.total_tiles = g.m_tiles * g.n_tiles * (g.pair_tab ? g.n_pairs : 1);
  g.idesc = tc::make_idesc_f16(g.BN);
  g.tmem_cols = tmem_cols_for(g.BN);
  g.ring_bytes = (uint32_t)gemm_ring_bytes(g.BN, g.stages, g.halo);
}

static inline void gemm_fill_geom(GemmGeom& g, int M, int N, int K, int BN, int a_k_off) {
  g.M = M; g.N = N; g.K = K; g.BN = BN; g.a_k_off = a_k_off;
  g.conv = 0; g.halo = 0; g.H = g.W = 0; g.tiles_x = g.tiles_y = 0;
  g.kb_per_row = (K + 63) / 64;
  g.num_kb = g.kb_per_row;
  g.stages = 4;
  g.epi_warp_bytes = 0;
  g.bias_bytes = 0;
  g.pair_tab = nullptr; g.n_pairs = 0;
  g.split3 = 0;
  gemm_finish_geom(g, (M + 127) / 128);
}
static inline void gemm_fill_geom_conv(GemmGeom& g, int B, int H, int W, int C, int N, int BN) {
  g.M = B * H * W; g.N = N; g.K = C; g.BN = BN; g.a_k_off = 0;
  g.conv = 1; g.halo = hfb_conv_halo(); g.H = H; g.W = W; g.tiles_x = (W + 15) / 16; g.tiles_y = (H + 7) / 8;
  g.kb_per_row = (C + 63) / 64;
  g.num_kb = 9 * g.kb_per_row;
  g.stages = 4;
  g.epi_warp_bytes = 0;
  g.bias_bytes = 0;
  g.pair_tab = nullptr; g.n_pairs = 0;
  g.split3 = 0;
  gemm_finish_geom(g, g.tiles_x * g.tiles_y * B);
}
// Re-derive the tile counts for the actual batch / row count of a launch (plans are built for max_batch).
static inline void gemm_set_rows(GemmGeom& g, int M, int B) {
  g.M = M;
  gemm_finish_geom(g, g.conv ? g.tiles_x * g.tiles_y * B : (M + 127) / 128);
}
// Persistent grid: as many CTAs as fit per SM by shared memory and TMEM columns, never more than tiles.
static inline int gemm_grid(const GemmGeom& g, int n_sm, size_t smem_bytes) {
  int per_sm = (int)(227 * 1024 / (smem_bytes + 1024));
  const int by_tmem = 512 / (int)g.tmem_cols;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const long long want = (long long)n_sm * per_sm;
  return (int)(g.total_tiles < want ? g.total_tiles : want);
}
