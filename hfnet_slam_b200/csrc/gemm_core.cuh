// The one tensor-core main loop of this library (sm_100a): C[M][N] = A[M][K] * W[N][K]^T with fp16 operands staged by
// TMA into 128B-swizzled shared-memory tiles, tcgen05.mma (UMMA 128 x BN x 16) accumulating fp32 in TMEM, and a
// pluggable epilogue that reads the accumulator tile back with tcgen05.ld (thread t <-> tile row t).
//
// A comes either from a 2-D map [rows][K] (1x1 convolutions, descriptor / database matrices) or, for the 3x3 SAME
// stride-1 head convolutions, from a 4-D map (C, W, H, B) of an NHWC tensor: one tile row = one pixel of an 8 x 16
// patch, one k-block = (tap, 64-channel slice), the tap shift is a TMA coordinate offset and TMA's out-of-bounds zero
// fill is the SAME padding (hfnet/models/hf_net.py:63,70).
//
// CTA = 128 threads, one 128 x BN output tile.  Thread 0 is the TMA producer, thread 32 the UMMA issuer, then all four
// warps run the epilogue (warp w owns TMEM lanes 32w..32w+31).
#pragma once
#include "tc.cuh"

struct GemmGeom {
  int M, N, K;        // logical sizes; conv: K = channels per tap
  int BN;             // N tile: multiple of 16, <= 256
  int a_k_off;        // first K coordinate inside A's map
  int conv;           // 1 = implicit 3x3
  int H, W;           // conv: image size; tiles are 8 rows x 16 cols
  int tiles_x, tiles_y;
  int kb_per_row;     // k-blocks per tap (conv) / in total (plain) = ceil(K/64)
  int num_kb;         // plain: kb_per_row, conv: 9 * kb_per_row
  int stages;         // smem ring depth (<= 8)
  uint32_t idesc;
  uint32_t tmem_cols; // power of two >= max(32, BN)
  uint32_t ring_bytes; // operand ring size (>= stages * stage bytes); barriers live right behind it
  // Batched independent problems (descriptor matching): blockIdx.z = pair, pair_tab = device int[4][n_pairs] holding
  // a_off | a_cnt | b_off | b_cnt (row ranges inside A's and W's maps).  null = one problem of M x N.
  const int* pair_tab;
  int n_pairs;
};

struct TileRow {
  bool valid;
  long long row;      // output row (pixel index / global A row) of this thread
  int n0;             // first output column of the tile (problem-local)
  uint32_t taddr;     // TMEM address of (this warp's lane base, column 0)
  int row_local;      // row inside the problem (== row when not batched)
  int n_cnt;          // columns of the problem (== N when not batched)
  int a_off, b_off;   // batched: first global row of the pair's A / B block
  uint8_t* stage;     // operand ring (1024-aligned), free for epilogue staging once the accumulator is complete
};

#define GEMM_TILE_A_BYTES 16384  // 128 rows x 128 B

// epi_bytes: shared memory the epilogue wants to reuse from the operand ring (the ring is grown if smaller)
static inline size_t gemm_ring_bytes(int BN, int stages, size_t epi_bytes = 0) {
  const size_t ring = (size_t)stages * (GEMM_TILE_A_BYTES + (size_t)BN * 128);
  return (ring > epi_bytes ? ring : epi_bytes + 1023) & ~(size_t)1023;
}
static inline size_t gemm_smem_bytes(int BN, int stages, size_t epi_bytes = 0) {
  return 1024 + gemm_ring_bytes(BN, stages, epi_bytes) + 256;
}

template <class Epi>
__global__ void __launch_bounds__(128) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const GemmGeom g,
                                                      const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = GEMM_TILE_A_BYTES + (uint32_t)g.BN * 128u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.ring_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* acc_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.y * g.BN;
  int a_off = 0, a_cnt = g.M, b_off = 0, b_cnt = g.N;
  if (g.pair_tab) {
    const int pr = blockIdx.z;
    a_off = g.pair_tab[pr];
    a_cnt = g.pair_tab[g.n_pairs + pr];
    b_off = g.pair_tab[2 * g.n_pairs + pr];
    b_cnt = g.pair_tab[3 * g.n_pairs + pr];
    if ((int)blockIdx.x * 128 >= a_cnt || n0 >= b_cnt) return;  // uniform: whole CTA leaves before any setup
  }
  // tile origin
  int m0 = 0, img = 0, y0 = 0, x0 = 0;
  if (g.conv) {
    int t = blockIdx.x;
    int tx = t % g.tiles_x;
    t /= g.tiles_x;
    int ty = t % g.tiles_y;
    img = t / g.tiles_y;
    y0 = ty * 8;
    x0 = tx * 16;
  } else {
    m0 = a_off + blockIdx.x * 128;
  }

  if (tid == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < g.stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, g.tmem_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int kb = 0; kb < g.num_kb; ++kb) {
      const int s = kb % g.stages;
      const uint32_t ph = (uint32_t)(kb / g.stages) & 1u;
      tc::mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* sa = smem + (size_t)s * stage_bytes;
      uint8_t* sb = sa + GEMM_TILE_A_BYTES;
      tc::mbar_expect_tx(&full[s], stage_bytes);
      if (g.conv) {
        const int tap = kb / g.kb_per_row, cb = kb - tap * g.kb_per_row;
        const int dy = tap / 3, dx = tap - dy * 3;
        tc::tma_load_4d(sa, &tmA, &full[s], cb * 64, x0 + dx - 1, y0 + dy - 1, img);
        tc::tma_load_2d(sb, &tmB, &full[s], tap * g.K + cb * 64, n0);
      } else {
        tc::tma_load_2d(sa, &tmA, &full[s], g.a_k_off + kb * 64, m0);
        tc::tma_load_2d(sb, &tmB, &full[s], kb * 64, b_off + n0);
      }
    }
  } else if (tid == 32) {
    // ------------------------------------------------------------------ UMMA issuer
    for (int kb = 0; kb < g.num_kb; ++kb) {
      const int s = kb % g.stages;
      const uint32_t ph = (uint32_t)(kb / g.stages) & 1u;
      tc::mbar_wait(&full[s], ph);
      tc::fence_after_sync();
      const uint32_t sa = tc::smem_u32(smem + (size_t)s * stage_bytes);
      const uint64_t da = tc::make_sdesc_sw128(sa);
      const uint64_t db = tc::make_sdesc_sw128(sa + GEMM_TILE_A_BYTES);
      const int cb = kb % g.kb_per_row;
      const int krem = g.K - cb * 64;                       // valid K elements in this block
      const int nk = krem >= 64 ? 4 : (krem + 15) >> 4;     // zero-filled tail needs no MMA
      for (int k = 0; k < nk; ++k)
        tc::umma_f16(tmem_base, tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k), g.idesc,
                     (kb > 0 || k > 0) ? 1u : 0u);
      tc::umma_commit(&empty[s]);                            // frees the smem slot when these MMAs retire
    }
    tc::umma_commit(acc_full);
  }
  __syncwarp();
  // ------------------------------------------------------------------ epilogue (all 128 threads)
  tc::mbar_wait(acc_full, 0);
  __syncwarp();
  tc::fence_after_sync();

  TileRow tr;
  tr.n0 = n0;
  tr.taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  if (g.conv) {
    const int y = y0 + (tid >> 4), x = x0 + (tid & 15);
    tr.valid = (y < g.H) && (x < g.W);
    tr.row = ((long long)img * g.H + y) * g.W + x;
    tr.row_local = (int)tr.row;
  } else {
    tr.row_local = blockIdx.x * 128 + tid;
    tr.row = (long long)m0 + tid;
    tr.valid = tr.row_local < a_cnt;
  }
  tr.n_cnt = b_cnt;
  tr.a_off = a_off;
  tr.b_off = b_off;
  tr.stage = smem;
  Epi::run(ep, g, tr);

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, g.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host-side launch
struct hfb_ctx;
int hfb_make_tmap_2d(hfb_ctx* ctx, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t row_stride_bytes, uint32_t box_outer);
int hfb_make_tmap_nhwc(hfb_ctx* ctx, CUtensorMap* out, const void* base, int C, int W, int H, int B);

static inline uint32_t tmem_cols_for(int BN) {
  uint32_t c = 32;
  while ((int)c < BN) c <<= 1;
  return c;
}

static inline void gemm_fill_geom(GemmGeom& g, int M, int N, int K, int BN, int a_k_off) {
  g.M = M; g.N = N; g.K = K; g.BN = BN; g.a_k_off = a_k_off;
  g.conv = 0; g.H = g.W = 0; g.tiles_x = g.tiles_y = 0;
  g.kb_per_row = (K + 63) / 64;
  g.num_kb = g.kb_per_row;
  g.stages = g.num_kb < 4 ? g.num_kb : 4;
  g.idesc = tc::make_idesc_f16(BN);
  g.tmem_cols = tmem_cols_for(BN);
  g.ring_bytes = (uint32_t)gemm_ring_bytes(BN, g.stages);
  g.pair_tab = nullptr; g.n_pairs = 0;
}
static inline void gemm_fill_geom_conv(GemmGeom& g, int B, int H, int W, int C, int N, int BN) {
  g.M = B * H * W; g.N = N; g.K = C; g.BN = BN; g.a_k_off = 0;
  g.conv = 1; g.H = H; g.W = W; g.tiles_x = (W + 15) / 16; g.tiles_y = (H + 7) / 8;
  g.kb_per_row = (C + 63) / 64;
  g.num_kb = 9 * g.kb_per_row;
  g.stages = 4;
  g.idesc = tc::make_idesc_f16(BN);
  g.tmem_cols = tmem_cols_for(BN);
  g.ring_bytes = (uint32_t)gemm_ring_bytes(BN, g.stages);
  g.pair_tab = nullptr; g.n_pairs = 0;
}
static inline dim3 gemm_grid(const GemmGeom& g, int B) {
  int mt = g.conv ? g.tiles_x * g.tiles_y * B : (g.M + 127) / 128;
  return dim3((unsigned)mt, (unsigned)((g.N + g.BN - 1) / g.BN), 1);
}
