// 256-d local-descriptor matching (replaces the brute-force flavours of src/Matcher.cc): mutual nearest neighbours of
// two descriptor sets as ONE fused tensor-core contraction + arg-max per pair.
//
//   SearchForTriangulation  (src/Matcher.cc:845-889)  S = D1 * D2^T, row arg-max above 1 - 0.5*TH_HIGH^2, column
//                                                     cross-check, strict '>' (lowest index wins ties)
//   SearchByBoW             (src/Matcher.cc:220-263, :561-621)  cv::BFMatcher(NORM_L2, crossCheck).match, dist < TH_LOW
//
// fp32 descriptors are split into fp16 hi + lo parts (a = ah + al, |al| <= 2^-11 |a|), stored once per row as
// [hi(256) | lo(256)], and the contraction runs over K' = 768: the TMA producer reads A's k-blocks as hi|hi|lo and B's as
// hi|lo|hi (GemmGeom::split3), i.e. ah.bh + ah.bl + al.bh: the dropped al.bl term is ~1e-8, so the scores that drive the
// arg-max carry fp32-level accuracy while using kind::f16 UMMA.  The epilogue (8 warps: two per TMEM lane group, each
// taking 64 of the tile's 128 columns) reduces every 128 x 128 accumulator tile to per-row and per-column (key, index)
// maxima straight out of TMEM: the row scan is thread-local; the column maxima of a warp's 32 rows come from a
// register butterfly (transpose-reduce over 32 columns: 31 shuffles instead of 32 warp reductions), the winning row of
// each column from one ballot.  Both are merged with packed 64-bit atomicMax.  The accepted value (cosine or L2
// distance) is recomputed in plain fp32 from the original descriptors before thresholding, like
// Matcher::DescriptorDistance (src/Matcher.cc:1893-1900).
#include "common.cuh"
#include "gemm_core.cuh"

#define MATCH_K 768      // contraction length of the split product
#define MATCH_LD 512     // stored row: hi(256) | lo(256) fp16
#define MATCH_BN 128

// ---- prep: fp32 [n][256] -> fp16 [n][512] (hi | lo) and half squared norms ------------------------------------------
__global__ void match_prep_kernel(const float* __restrict__ X, int n, __half* __restrict__ out, float* __restrict__ hn,
                                  int l2_mode) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4* src = reinterpret_cast<const float4*>(X + (size_t)row * 256) + lane * 2;
  const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
  const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint4 hi, lo;
  __half2* hh = reinterpret_cast<__half2*>(&hi);
  __half2* hl = reinterpret_cast<__half2*>(&lo);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half a0 = __float2half_rn(v[2 * j]), a1 = __float2half_rn(v[2 * j + 1]);
    hh[j] = __halves2half2(a0, a1);
    hl[j] = __floats2half2_rn(v[2 * j] - __half2float(a0), v[2 * j + 1] - __half2float(a1));
    ss = fmaf(v[2 * j], v[2 * j], ss);
    ss = fmaf(v[2 * j + 1], v[2 * j + 1], ss);
  }
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)row * MATCH_LD) + lane;
  o[0] = hi;      // k in [0,256)
  o[32] = lo;     // k in [256,512)
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  if (lane == 0) hn[row] = l2_mode ? 0.5f * ss : 0.f;
}

__device__ __forceinline__ u64 pack_best(float key, int idx) {
  return ((u64)f2ord(key) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)idx);
}

// ---- the matcher's own tensor-core kernel -------------------------------------------------------------------------------
// One CTA per SM walks work units (pair, 128-row block of A, chunk of B's 128-column tiles).  The A block lives in shared
// memory for the whole unit (8 swizzled k-blocks: hi 0..3 | lo 0..3 = 128 KB) and only B streams through a 5-slot ring,
// so a 128 x 128 tile costs 128 KB of L2 traffic instead of the 384 KB of a both-operands-streamed split product (which
// pins the kernel to the L2 -> SM path, not the tensor pipe).  Per B k-block pair the issuer runs
//     hi_k: D += A_hi[k] . B_hi[k],  D += A_lo[k] . B_hi[k]          lo_k: D += A_hi[k] . B_lo[k]
// (12 k-block products from 8 loads).  Roles: warp 0 TMA producer, warp 1 UMMA issuer + TMEM owner (two accumulator
// stages), warps 2..9 epilogue: warp (lane group g, half h) owns rows 32 g .. 32 g + 31 and columns 64 h .. 64 h + 63 of
// the tile.  Row arg-max: thread-local scan.  Column arg-max: every warp transposes 32 x 8 key blocks through a private
// padded shared-memory patch (2 x 16-byte stores per thread, conflict-free), then lane (c, h) scans rows 8 h .. 8 h + 7
// of column c and two shuffles join the four row groups -- no warp reductions.  Both land in packed 64-bit atomicMax
// (key << 32 | ~index: lowest index wins ties).  The patches are small on purpose: what bounds the kernel is the bytes
// of B in flight per SM (ring capacity / TMA latency), so shared memory goes to the ring first.
#define MAR_THREADS 320
#define MAR_A_BYTES (8 * 16384)
#define MAR_B_SLOTS 5
#define MAR_PATCH_WORDS (32 * 12 + 24)
#define MAR_SMEM (1024 + MAR_A_BYTES + MAR_B_SLOTS * 16384 + 8 * MAR_PATCH_WORDS * 4 + 2 * MATCH_BN * 4 + 256)

struct MatchGeom {
  int n_pairs, m_tiles, n_tiles, chunk_tiles, n_chunks, total_units;
  const int* pair_tab;   // device int[4][n_pairs]: a_off | a_cnt | b_off | b_cnt
  uint32_t idesc;
};

struct MatchUnit {
  bool valid;
  int a_off, a_cnt, b_off, b_cnt, m0, nt0, nt1;   // rows of A: a_off + m0 .. ; B tiles nt0 .. nt1-1
};

__device__ __forceinline__ MatchUnit match_decode(const MatchGeom& g, int unit) {
  MatchUnit u;
  const int per = g.m_tiles * g.n_chunks;
  const int pr = unit / per;
  const int rem = unit - pr * per;
  const int mt = rem / g.n_chunks, ch = rem - mt * g.n_chunks;
  u.a_off = g.pair_tab[pr];
  u.a_cnt = g.pair_tab[g.n_pairs + pr];
  u.b_off = g.pair_tab[2 * g.n_pairs + pr];
  u.b_cnt = g.pair_tab[3 * g.n_pairs + pr];
  u.m0 = mt * 128;
  u.nt0 = ch * g.chunk_tiles;
  const int nt_valid = (u.b_cnt + MATCH_BN - 1) / MATCH_BN;
  u.nt1 = min(min(u.nt0 + g.chunk_tiles, g.n_tiles), nt_valid);
  u.valid = u.m0 < u.a_cnt && u.nt0 < u.nt1;
  return u;
}

// Epilogue work of one warp on a 32-row x 32-column block of the accumulator (r = this thread's row): one common key
// k = s - 0.5|b_j|^2 - 0.5|a_i|^2 (= -0.5 dist^2; the norm terms are 0 in cosine mode) serves both arg-maxes -- the
// per-row term does not change a row's ordering, the per-column term not a column's.  Rows / columns outside the pair
// carry +inf in their norm term, so their keys are -inf (or NaN) and never win a strict '>' -- no masks in the loops.
// Row pass: thread-local, ascending columns, strict '>' (lowest index on ties).  Column pass: 8 columns at a time through
// the warp's shared-memory patch (32 x 8 transposition), lane (c, h) scans rows 8 h .. 8 h + 7 of column c, two shuffles
// join the four row groups.
__device__ __forceinline__ void match_epi_block(const uint32_t (&r)[32], const float* hb, float my_hna, int ncols,
                                                int col0, float& best, int& best_j, uint32_t* my_patch, int lane,
                                                u64* colbest, int row0) {
  float k[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 h4 = *reinterpret_cast<const float4*>(hb + 4 * q);
    k[4 * q + 0] = (__uint_as_float(r[4 * q + 0]) - h4.x) - my_hna;
    k[4 * q + 1] = (__uint_as_float(r[4 * q + 1]) - h4.y) - my_hna;
    k[4 * q + 2] = (__uint_as_float(r[4 * q + 2]) - h4.z) - my_hna;
    k[4 * q + 3] = (__uint_as_float(r[4 * q + 3]) - h4.w) - my_hna;
  }
  // four independent compare chains of eight columns (instruction-level parallelism), joined in ascending order with
  // strict '>' so that the lowest column still wins ties
  float cb[4];
  int cj[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    cb[q] = k[8 * q];
    cj[q] = 8 * q;
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      const bool p = k[8 * q + j] > cb[q];
      cb[q] = p ? k[8 * q + j] : cb[q];
      cj[q] = p ? 8 * q + j : cj[q];
    }
  }
  float lb = cb[0];
  int lj = cj[0];
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    const bool p = cb[q] > lb;
    lb = p ? cb[q] : lb;
    lj = p ? cj[q] : lj;
  }
  if (lb > best) {
    best = lb;
    best_j = col0 + lj;
  }
  const int c = lane & 7, h = lane >> 3;
  // patch row of tile row `lane`: 12 words apart, every group of 8 rows shifted by 8 more words (the 16-byte stores of
  // 8 lanes and the column reads each hit 32 distinct banks)
  float4* prow = reinterpret_cast<float4*>(my_patch + lane * 12 + (lane & 24));
  const float* pcol = reinterpret_cast<const float*>(my_patch) + (8 * h) * 12 + 8 * h + c;
#pragma unroll
  for (int g8 = 0; g8 < 4; ++g8) {
    if (8 * g8 >= ncols) break;   // warp-uniform
    prow[0] = make_float4(k[8 * g8 + 0], k[8 * g8 + 1], k[8 * g8 + 2], k[8 * g8 + 3]);
    prow[1] = make_float4(k[8 * g8 + 4], k[8 * g8 + 5], k[8 * g8 + 6], k[8 * g8 + 7]);
    __syncwarp();
    float m = -INFINITY;
    int mi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {       // column c over rows 8 h + i, ascending: lowest row on ties
      const float v = pcol[i * 12];
      const bool p = v > m;
      m = p ? v : m;
      mi = p ? i : mi;
    }
    mi += 8 * h;
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {  // join the four row groups: a higher group wins only when strictly greater
      const float m2 = __shfl_down_sync(0xffffffffu, m, o);
      const int mi2 = __shfl_down_sync(0xffffffffu, mi, o);
      const bool p = m2 > m;
      m = p ? m2 : m;
      mi = p ? mi2 : mi;
    }
    __syncwarp();
    if (h == 0 && m > -INFINITY)
      atomicMax(colbest + 8 * g8 + c, ((u64)f2ord(m) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)(row0 + mi)));
  }
}

__global__ void __launch_bounds__(MAR_THREADS, 1)
match_ar_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MatchGeom g,
                const float* __restrict__ hna, const float* __restrict__ hnb, u64* __restrict__ rowbest,
                u64* __restrict__ colbest) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                  // 8 k-blocks x 16 KB
  uint8_t* sB = smem + MAR_A_BYTES;                    // ring
  uint32_t* patch = reinterpret_cast<uint32_t*>(sB + MAR_B_SLOTS * 16384);   // [8 warps][32][36]
  float* s_hnb = reinterpret_cast<float*>(patch + 8 * MAR_PATCH_WORDS);      // [2 acc stages][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_hnb + 2 * MATCH_BN);
  uint64_t* a_full = bars;            // [1]
  uint64_t* a_empty = bars + 1;       // [1]
  uint64_t* b_full = bars + 2;        // [MAR_B_SLOTS]
  uint64_t* b_empty = bars + 2 + MAR_B_SLOTS;
  uint64_t* acc_full = bars + 2 + 2 * MAR_B_SLOTS;   // [2]
  uint64_t* acc_empty = acc_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  tc::pdl_launch_dependents();
  if (tid == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::mbar_init(a_full, 1);
    tc::mbar_init(a_empty, 1);
    for (int s = 0; s < MAR_B_SLOTS; ++s) {
      tc::mbar_init(&b_full[s], 1);
      tc::mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], 8);   // one arrival per epilogue warp
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  tc::pdl_wait();   // descriptor images / norms / zeroed best arrays come from the predecessors
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ua = 0, bi = 0;
      for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
        const MatchUnit u = match_decode(g, unit);
        if (!u.valid) continue;
        tc::mbar_wait(a_empty, (ua & 1u) ^ 1u);          // the previous unit's products have retired
        tc::mbar_expect_tx(a_full, MAR_A_BYTES);
        for (int kb = 0; kb < 8; ++kb) tc::tma_load_2d(sA + kb * 16384, &tmA, a_full, kb * 64, u.a_off + u.m0);
        ++ua;
        for (int nt = u.nt0; nt < u.nt1; ++nt)
          for (int j = 0; j < 8; ++j, ++bi) {            // hi0 lo0 hi1 lo1 ...
            const int s = bi % MAR_B_SLOTS;
            tc::mbar_wait(&b_empty[s], ((bi / MAR_B_SLOTS) & 1u) ^ 1u);
            tc::mbar_expect_tx(&b_full[s], 16384);
            const int kq = j >> 1, lo = j & 1;
            tc::tma_load_2d(sB + s * 16384, &tmB, &b_full[s], (lo * 4 + kq) * 64, u.b_off + nt * MATCH_BN);
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      uint32_t ua = 0, bi = 0, acc_it = 0;
      const uint32_t sa0 = tc::smem_u32(sA), sb0 = tc::smem_u32(sB);
      for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
        const MatchUnit u = match_decode(g, unit);
        if (!u.valid) continue;
        tc::mbar_wait(a_full, ua & 1u);
        ++ua;
        for (int nt = u.nt0; nt < u.nt1; ++nt) {
          const uint32_t as = acc_it & 1u;
          tc::mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1u) ^ 1u);
          tc::fence_after_sync();
          const uint32_t d_tmem = tmem_base + as * MATCH_BN;
          for (int j = 0; j < 8; ++j, ++bi) {
            const int s = bi % MAR_B_SLOTS;
            tc::mbar_wait(&b_full[s], (bi / MAR_B_SLOTS) & 1u);
            tc::fence_after_sync();
            const int kq = j >> 1, lo = j & 1;
            const uint64_t db = tc::make_sdesc_sw128(sb0 + s * 16384);
            const uint64_t da_hi = tc::make_sdesc_sw128(sa0 + kq * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::umma_f16(d_tmem, tc::sdesc_advance_k16(da_hi, k), tc::sdesc_advance_k16(db, k), g.idesc,
                           (j > 0 || k > 0) ? 1u : 0u);
            if (!lo) {
              const uint64_t da_lo = tc::make_sdesc_sw128(sa0 + (4 + kq) * 16384);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc::umma_f16(d_tmem, tc::sdesc_advance_k16(da_lo, k), tc::sdesc_advance_k16(db, k), g.idesc, 1u);
            }
            tc::umma_commit(&b_empty[s]);
          }
          tc::umma_commit(&acc_full[as]);
          ++acc_it;
        }
        tc::umma_commit(a_empty);     // the A block may be replaced once every product of this unit has retired
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9
    const int lg = warp & 3;          // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2; // which 64 of the tile's 128 columns
    uint32_t* my_patch = patch + (warp - 2) * MAR_PATCH_WORDS;
    const int et = tid - 64;
    uint32_t acc_it = 0;
    for (int unit = blockIdx.x; unit < g.total_units; unit += gridDim.x) {
      const MatchUnit u = match_decode(g, unit);
      if (!u.valid) continue;
      const int row_local = u.m0 + lg * 32 + lane;
      const bool rvalid = row_local < u.a_cnt;
      const float my_hna = rvalid ? __ldg(hna + u.a_off + row_local) : INFINITY;   // rows outside the pair never win
      float best = -INFINITY;
      int best_j = 0;
      for (int nt = u.nt0; nt < u.nt1; ++nt) {
        const uint32_t as = acc_it & 1u;
        const int n0 = nt * MATCH_BN;
        const int ncols = min(MATCH_BN, u.b_cnt - n0);
        float* hb = s_hnb + as * MATCH_BN;
        // all eight warps left the tile that used this stage two tiles ago before the issuer let this one start
        if (et < MATCH_BN) hb[et] = et < ncols ? __ldg(hnb + u.b_off + n0 + et) : INFINITY;   // nor do such columns
        tc::mbar_wait(&acc_full[as], (acc_it >> 1) & 1u);
        __syncwarp();
        tc::fence_after_sync();
        epi_bar_sync_n<256>();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + as * MATCH_BN;
#pragma unroll 1
        for (int gi = 0; gi < 2; ++gi) {
          const int c0 = half * 64 + gi * 32;
          if (c0 >= ncols) break;   // warp-uniform
          uint32_t r[32];
          tc::tmem_ld32(taddr + (uint32_t)c0, r);
          tc::tmem_ld_wait();
          match_epi_block(r, hb + c0, my_hna, ncols - c0, n0 + c0, best, best_j, my_patch, lane,
                          colbest + u.b_off + n0 + c0, u.m0 + lg * 32);
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[as]);
        ++acc_it;
      }
      if (rvalid && best > -INFINITY) atomicMax(rowbest + u.a_off + row_local, pack_best(best, best_j));
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 256);
}

// Split-precision images of `n` fp32 rows at row offset `row0` of a resident block (keyframe store: prepared once).
int launch_match_prep(hfb_ctx* ctx, const float* d_rows, int n, __half* d_img, float* d_hn_l2) {
  if (n <= 0) return HFB_OK;
  hfb_launch(ctx, match_prep_kernel, ceil_div(n, 8), 256, 0, d_rows, n, d_img, d_hn_l2, 1);
  HFB_CHECK_LAUNCH(ctx, "match_prep(store)");
  return HFB_OK;
}

// ---- finalize: mutual check + exact fp32 value + threshold ----------------------------------------------------------
__global__ void match_finalize_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                      const u64* __restrict__ rowbest, const u64* __restrict__ colbest,
                                      const int* __restrict__ pair_tab, int n_pairs, int mode, float thr,
                                      int* __restrict__ match_idx, float* __restrict__ match_val,
                                      int* __restrict__ n_matches, int pad_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int pr = blockIdx.y;
  const int a_off = pair_tab[pr], a_cnt = pair_tab[n_pairs + pr], b_off = pair_tab[2 * n_pairs + pr];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= a_cnt) {
    if (i < pad_rows && lane == 0) {   // fixed-pitch rows (frame association): unmatched beyond the frame's keypoints
      match_idx[a_off + i] = -1;
      match_val[a_off + i] = 0.f;
    }
    return;
  }
  const u64 rb = rowbest[a_off + i];
  int j = -1;
  float val = 0.f;
  if (rb != 0ull) {
    const int jj = (int)(0xFFFFFFFFu - (uint32_t)(rb & 0xFFFFFFFFull));
    const u64 cb = colbest[b_off + jj];
    const int ii = (int)(0xFFFFFFFFu - (uint32_t)(cb & 0xFFFFFFFFull));
    if (cb != 0ull && ii == i) {
      const float4* a = reinterpret_cast<const float4*>(A + (size_t)(a_off + i) * 256) + lane * 2;
      const float4* b = reinterpret_cast<const float4*>(Bm + (size_t)(b_off + jj) * 256) + lane * 2;
      const float4 a0 = __ldg(a), a1 = __ldg(a + 1), b0 = __ldg(b), b1 = __ldg(b + 1);
      float acc;
      if (mode == 0) {
        float d;
        d = a0.x - b0.x; acc = d * d;
        d = a0.y - b0.y; acc = fmaf(d, d, acc);
        d = a0.z - b0.z; acc = fmaf(d, d, acc);
        d = a0.w - b0.w; acc = fmaf(d, d, acc);
        d = a1.x - b1.x; acc = fmaf(d, d, acc);
        d = a1.y - b1.y; acc = fmaf(d, d, acc);
        d = a1.z - b1.z; acc = fmaf(d, d, acc);
        d = a1.w - b1.w; acc = fmaf(d, d, acc);
      } else {
        acc = a0.x * b0.x;
        acc = fmaf(a0.y, b0.y, acc); acc = fmaf(a0.z, b0.z, acc); acc = fmaf(a0.w, b0.w, acc);
        acc = fmaf(a1.x, b1.x, acc); acc = fmaf(a1.y, b1.y, acc); acc = fmaf(a1.z, b1.z, acc);
        acc = fmaf(a1.w, b1.w, acc);
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (mode == 0) {
        val = sqrtf(acc);
        if (val < thr) j = jj;      // dist < TH_LOW, src/Matcher.cc:253
      } else {
        val = acc;
        if (val > thr) j = jj;      // strict '>', src/Matcher.cc:851,868
      }
    }
  }
  if (lane == 0) {
    match_idx[a_off + i] = j;
    match_val[a_off + i] = j >= 0 ? val : 0.f;
    if (j >= 0 && n_matches) atomicAdd(n_matches + pr, 1);
  }
}

// Workspace layout: X' (A rows) | X' (B rows, absent when A and B are the same array) | hna | hnb | rowbest | colbest |
// n_matches[n_pairs].  ws == nullptr: the context's grow-on-demand scratch; the in-graph association passes its own
// fixed block (captured graphs bake these addresses).
size_t match_workspace_bytes(int na_total, int nb_total, int n_pairs, bool same) {
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  const int nmax = same ? (na_total > nb_total ? na_total : nb_total) : na_total;
  return al((size_t)nmax * MATCH_LD * 2) + (same ? 0 : al((size_t)nb_total * MATCH_LD * 2)) + al((size_t)nmax * 4) +
         (same ? 0 : al((size_t)nb_total * 4)) + al((size_t)na_total * 8) + al((size_t)nb_total * 8) +
         al((size_t)n_pairs * 4);
}

int launch_match_batch(hfb_ctx* ctx, int mode, const float* dA, const float* dB, int n_pairs, const int* d_pair_tab,
                       int max_a, int max_b, float thr, int* d_match_idx, float* d_match_val, int na_total,
                       int nb_total, int** d_n_matches_out, void* ws, size_t ws_bytes, int pad_rows,
                       const __half* B_img, const float* hnb_pre) {
  if (na_total <= 0 || nb_total <= 0 || n_pairs <= 0 || max_a <= 0 || max_b <= 0) {
    if (na_total > 0) {
      HFB_CUDA(ctx, cudaMemsetAsync(d_match_idx, 0xFF, (size_t)na_total * 4, ctx->stream));
      HFB_CUDA(ctx, cudaMemsetAsync(d_match_val, 0, (size_t)na_total * 4, ctx->stream));
    }
    if (d_n_matches_out) *d_n_matches_out = nullptr;
    return HFB_OK;
  }
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  const bool same = dA == dB && !B_img;
  const int nmax = same ? std::max(na_total, nb_total) : na_total;
  // B_img: the B rows already exist as split-precision images (keyframe store): nothing to prepare or to hold for B
  const size_t szA = al((size_t)nmax * MATCH_LD * 2), szB = (same || B_img) ? 0 : al((size_t)nb_total * MATCH_LD * 2);
  const size_t szna = al((size_t)nmax * 4), sznb = (same || B_img) ? 0 : al((size_t)nb_total * 4);
  const size_t szrb = al((size_t)na_total * 8), szcb = al((size_t)nb_total * 8), sznm = al((size_t)n_pairs * 4);
  const size_t need = szA + szB + szna + sznb + szrb + szcb + sznm;
  uint8_t* base;
  if (ws) {
    HFB_REQUIRE(ctx, need <= ws_bytes, "matcher workspace too small");
    base = reinterpret_cast<uint8_t*>(ws);
  } else {
    HFB_TRY(ctx->ensure_scratch(need));
    base = reinterpret_cast<uint8_t*>(ctx->d_scratch);
  }
  __half* A2 = reinterpret_cast<__half*>(base);
  const __half* B2 = B_img ? B_img : (same ? A2 : reinterpret_cast<__half*>(base + szA));
  float* hna = reinterpret_cast<float*>(base + szA + szB);
  const float* hnb = B_img ? hnb_pre : (same ? hna : reinterpret_cast<float*>(base + szA + szB + szna));
  u64* rowbest = reinterpret_cast<u64*>(base + szA + szB + szna + sznb);
  u64* colbest = reinterpret_cast<u64*>(base + szA + szB + szna + sznb + szrb);
  int* nm = reinterpret_cast<int*>(base + szA + szB + szna + sznb + szrb + szcb);
  HFB_CUDA(ctx, cudaMemsetAsync(rowbest, 0, szrb + szcb + sznm, ctx->stream));

  const int l2 = (mode == 0);
  hfb_launch(ctx, match_prep_kernel, ceil_div(nmax, 8), 256, 0, dA, nmax, A2, hna, l2);
  HFB_CHECK_LAUNCH(ctx, "match_prep(A)");
  if (!same && !B_img) {
    hfb_launch(ctx, match_prep_kernel, ceil_div(nb_total, 8), 256, 0, dB, nb_total, const_cast<__half*>(B2),
               const_cast<float*>(hnb), l2);
    HFB_CHECK_LAUNCH(ctx, "match_prep(B)");
  }

  CUtensorMap tmA, tmB;
  HFB_TRY(hfb_make_tmap_2d(ctx, &tmA, A2, MATCH_LD, (uint64_t)nmax, MATCH_LD * 2, 128));
  HFB_TRY(hfb_make_tmap_2d(ctx, &tmB, B2, MATCH_LD, (uint64_t)(same ? nmax : nb_total), MATCH_LD * 2, MATCH_BN));
  MatchGeom g;
  g.n_pairs = n_pairs;
  g.m_tiles = ceil_div(max_a, 128);
  g.n_tiles = ceil_div(max_b, MATCH_BN);
  // B tiles per unit: the split that minimises the makespan of the static unit walk, counting the reload of the A block
  // (128 KB, about one tile's worth of L2 traffic) once per unit
  {
    const int base_units = n_pairs * g.m_tiles;
    long long best_cost = -1;
    g.chunk_tiles = g.n_tiles;
    g.n_chunks = 1;
    for (int chunks = 1; chunks <= g.n_tiles; ++chunks) {
      const int ct = ceil_div(g.n_tiles, chunks), nc = ceil_div(g.n_tiles, ct);
      const long long cost = (long long)ceil_div(base_units * nc, ctx->n_sm) * (ct + 1);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        g.chunk_tiles = ct;
        g.n_chunks = nc;
      }
    }
  }
  g.total_units = n_pairs * g.m_tiles * g.n_chunks;
  g.pair_tab = d_pair_tab;
  g.idesc = tc::make_idesc_f16(MATCH_BN);
  static SmemOptIn optin;
  HFB_CUDA(ctx, optin.ensure(match_ar_kernel, ctx->device, MAR_SMEM));
  ctx->note("match_gemm_argmax", (double)(na_total + nb_total) * 1024.0, 2.0 * n_pairs * (double)max_a * max_b * 256.0);
  hfb_launch(ctx, match_ar_kernel, std::min(ctx->n_sm, g.total_units), MAR_THREADS, MAR_SMEM, tmA, tmB, g, hna, hnb,
             rowbest, colbest);
  HFB_CHECK_LAUNCH(ctx, "match_gemm_argmax");

  dim3 fgrid(ceil_div(std::max(max_a, pad_rows), 8), n_pairs);
  hfb_launch(ctx, match_finalize_kernel, fgrid, 256, 0, dA, dB, rowbest, colbest, d_pair_tab, n_pairs, mode, thr,
                                                        d_match_idx, d_match_val, nm, pad_rows);
  HFB_CHECK_LAUNCH(ctx, "match_finalize");
  if (d_n_matches_out) *d_n_matches_out = nm;
  return HFB_OK;
}


// =====================================================================================================================
// Windowed matching = the descriptor stage of Matcher::SearchByProjection (src/Matcher.cc:40-210, 1574-1721) and
// Frame::GetFeaturesInArea (src/Frame.cc:659-725): for every projected map point, the K best frame features inside
// its search window |x - u| < r, |y - v| < r with octave in [minLevel, maxLevel].  The full nq x nf contraction runs on
// the tensor cores; the window is a predicate in the epilogue (no grid cells: the reference's cells only accelerate the
// same predicate).  K = 4 sorted candidates per query let the caller replay the reference's sequential bookkeeping
// (features claimed by earlier map points are skipped, best / second-best + level-aware ratio test) without any
// descriptor arithmetic on the host.
#define PROJ_K 4

struct ProjRec {
  float d[PROJ_K];   // half squared distance proxy (tile-local ranking key), ascending
  int j[PROJ_K];     // feature index or -1
};

__device__ __forceinline__ void proj_insert(ProjRec& r, float d, int j) {
  if (!(d < r.d[PROJ_K - 1])) return;   // strict '<' keeps the earlier feature on ties (Matcher.cc:98,109)
  int pos = PROJ_K - 1;
#pragma unroll
  for (int k = PROJ_K - 1; k > 0; --k) {
    if (d < r.d[k - 1]) {
      r.d[k] = r.d[k - 1];
      r.j[k] = r.j[k - 1];
      pos = k - 1;
    }
  }
  r.d[pos] = d;
  r.j[pos] = j;
}

struct EpiProjTopK {
  static constexpr int kWarps = 4;
  struct Params {
    const float* hnq;      // [nq] 0.5*|q|^2
    const float* hnf;      // [nf]
    const float4* qwin;    // [nq] (u, v, r, unused)
    const int2* qlev;      // [nq] (minLevel, maxLevel)
    const float2* fxy;     // [nf]
    const int* flevel;     // [nf]
    const unsigned char* fskip;  // [nf] or null
    const float* finv;     // [nf] inverse level sigma^2 of every feature, or null: chi^2 gate (dx^2+dy^2)*finv <= chi2_max
    float chi2_max;
    ProjRec* rec;          // [n_tiles][nq]
    int nq;
  };
  static __device__ __forceinline__ const float* bias(const Params&) { return nullptr; }
  static __device__ __forceinline__ void run(const Params& p, const GemmGeom& g, const TileRow& tr) {
    __shared__ float s_fx[MATCH_BN], s_fy[MATCH_BN], s_hn[MATCH_BN], s_inv[MATCH_BN];
    __shared__ int s_lv[MATCH_BN];
    const int lane = threadIdx.x & 31, et = tr.ewarp * 32 + lane;
    const int ncols = min(g.BN, tr.n_cnt - tr.n0);
    if (et < ncols) {
      const int j = tr.n0 + et;
      const float2 xy = __ldg(p.fxy + j);
      s_fx[et] = xy.x;
      s_fy[et] = xy.y;
      s_hn[et] = __ldg(p.hnf + j);
      s_inv[et] = p.finv ? __ldg(p.finv + j) : 0.f;   // 0: the gate below never rejects
      s_lv[et] = (p.fskip && p.fskip[j]) ? -1000000 : __ldg(p.flevel + j);   // skipped features fail every level test
    }
    epi_bar_sync();
    ProjRec best;
#pragma unroll
    for (int k = 0; k < PROJ_K; ++k) {
      best.d[k] = 3.0e38f;
      best.j[k] = -1;
    }
    float4 win = make_float4(0.f, 0.f, -1.f, 0.f);
    int2 lev = make_int2(0, -1);
    float hq = 0.f;
    if (tr.valid) {
      win = __ldg(p.qwin + tr.row);
      lev = __ldg(p.qlev + tr.row);
      hq = __ldg(p.hnq + tr.row);
    }
    for (int c0 = 0; c0 < g.BN; c0 += 16) {
      uint32_t r[16];
      tc::tmem_ld16(tr.taddr + (uint32_t)c0, r);
      tc::tmem_ld_wait();
      if (c0 >= ncols || !tr.valid) continue;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int c = c0 + jj;
        if (c >= ncols) break;
        const int lv = s_lv[c];
        const float ex = win.x - s_fx[c], ey = win.y - s_fy[c];
        // Matcher::Fuse's reprojection gate (Matcher.cc:1180-1188): skip iff e2 * invSigma2 > chi2_max, same expression order
        const bool gate_ok = !(__fmul_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), s_inv[c]) > p.chi2_max);
        const bool in_win = fabsf(s_fx[c] - win.x) < win.z && fabsf(s_fy[c] - win.y) < win.z && lv >= lev.x &&
                            (lev.y < 0 || lv <= lev.y) && lv > -1000000 && gate_ok;
        if (in_win) proj_insert(best, hq + s_hn[c] - __uint_as_float(r[jj]), tr.n0 + c);
      }
    }
    if (tr.valid) p.rec[(size_t)(tr.n0 / g.BN) * p.nq + tr.row] = best;
    epi_bar_sync();
  }
};

// merge the per-tile lists (ascending tile = ascending feature index), then exact fp32 distances of the K survivors
__global__ void proj_finalize_kernel(const float* __restrict__ Q, const float* __restrict__ F, const ProjRec* __restrict__ rec,
                                     int n_tiles, int nq, const int* __restrict__ flevel, int* __restrict__ cand_idx,
                                     float* __restrict__ cand_dist, int* __restrict__ cand_level) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= nq) return;
  ProjRec best;
#pragma unroll
  for (int k = 0; k < PROJ_K; ++k) {
    best.d[k] = 3.0e38f;
    best.j[k] = -1;
  }
  for (int t = 0; t < n_tiles; ++t) {
    const ProjRec r = rec[(size_t)t * nq + i];
#pragma unroll
    for (int k = 0; k < PROJ_K; ++k)
      if (r.j[k] >= 0) proj_insert(best, r.d[k], r.j[k]);
  }
  const float4* q = reinterpret_cast<const float4*>(Q + (size_t)i * 256) + lane * 2;
  const float4 q0 = __ldg(q), q1 = __ldg(q + 1);
  float dist[PROJ_K];
#pragma unroll
  for (int k = 0; k < PROJ_K; ++k) {
    float acc = 0.f;
    if (best.j[k] >= 0) {
      const float4* f = reinterpret_cast<const float4*>(F + (size_t)best.j[k] * 256) + lane * 2;
      const float4 f0 = __ldg(f), f1 = __ldg(f + 1);
      float d;
      d = q0.x - f0.x; acc = d * d;
      d = q0.y - f0.y; acc = fmaf(d, d, acc);
      d = q0.z - f0.z; acc = fmaf(d, d, acc);
      d = q0.w - f0.w; acc = fmaf(d, d, acc);
      d = q1.x - f1.x; acc = fmaf(d, d, acc);
      d = q1.y - f1.y; acc = fmaf(d, d, acc);
      d = q1.z - f1.z; acc = fmaf(d, d, acc);
      d = q1.w - f1.w; acc = fmaf(d, d, acc);
    }
#pragma unroll
    for (int s2 = 16; s2 > 0; s2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s2);
    dist[k] = sqrtf(acc);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < PROJ_K; ++k) {
      const int j = best.j[k];
      cand_idx[(size_t)i * PROJ_K + k] = j;
      cand_dist[(size_t)i * PROJ_K + k] = j >= 0 ? dist[k] : 3.402823466e38f;
      cand_level[(size_t)i * PROJ_K + k] = j >= 0 ? __ldg(flevel + j) : -1;
    }
  }
}

__global__ void proj_pack_kernel(const float* __restrict__ uv, const float* __restrict__ radius, const int* __restrict__ minl,
                                 const int* __restrict__ maxl, int nq, float4* __restrict__ qwin, int2* __restrict__ qlev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  qwin[i] = make_float4(uv[2 * i], uv[2 * i + 1], radius[i], 0.f);
  qlev[i] = make_int2(minl[i], maxl[i]);
}

__global__ void proj_gather_rows_kernel(const float* __restrict__ base, const int* __restrict__ index, int n,
                                        float* __restrict__ out) {
  const int i = blockIdx.x, t = threadIdx.x;     // one CTA of 64 threads per 1 KB row
  if (i >= n) return;
  reinterpret_cast<float4*>(out + (size_t)i * 256)[t] = __ldg(reinterpret_cast<const float4*>(base + (size_t)index[i] * 256) + t);
}

__global__ void proj_pack_xy_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, float2* __restrict__ xy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) xy[i] = make_float2(x[i], y[i]);
}

// Shared body of the windowed searches.  Queries: host rows (Q) or rows q_index[i] of a resident descriptor block
// (dQ_base).  Features: host arrays (F, f_xy, f_level) or resident arrays (dF, d_fx, d_fy, d_flevel).
static int proj_common(hfb_ctx* ctx, const float* Q, const int32_t* q_index, const float* dQ_base, int32_t nq,
                       const float* q_uv, const float* q_radius, const int32_t* q_min_level, const int32_t* q_max_level,
                       const float* F, const float* dF_res, const float* d_fx, const float* d_fy, const int* d_flevel,
                       int32_t nf, const float* f_xy, const int32_t* f_level, const uint8_t* f_skip,
                       const float* f_inv_sigma2, float chi2_max, int32_t* cand_idx, float* cand_dist,
                       int32_t* cand_level) {
  HFB_REQUIRE(ctx, nq >= 0 && nf >= 0, "negative size");
  HFB_REQUIRE(ctx, cand_idx && cand_dist && cand_level, "null output");
  for (long long i = 0; i < (long long)nq * PROJ_K; ++i) {
    cand_idx[i] = -1;
    cand_dist[i] = 3.402823466e38f;
    cand_level[i] = -1;
  }
  if (nq == 0 || nf == 0) return HFB_OK;
  const bool q_res = Q == nullptr, f_res = F == nullptr;
  HFB_REQUIRE(ctx, (q_res ? (q_index && dQ_base) : true) && q_uv && q_radius && q_min_level && q_max_level &&
                       (f_res ? (dF_res && d_fx && d_fy && d_flevel) : (f_xy && f_level)), "null input");
  auto al = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const int n_tiles = ceil_div(nf, MATCH_BN);
  // io block: Q | F | [uv | r | minl | maxl | q_index] | fxy | flevel | fskip | finv | [cand_idx | cand_dist | cand_level]
  // The bracketed groups are contiguous: the per-query windows go up as ONE copy out of a page-locked staging block and
  // the candidate lists come back as ONE copy into it (eight small pageable copies cost ~50 us of a ~170 us call).
  const size_t oQ = 0, oF = oQ + al((size_t)nq * 1024), o_uv = oF + (f_res ? 0 : al((size_t)nf * 1024)),
               o_r = o_uv + al((size_t)nq * 8), o_mn = o_r + al((size_t)nq * 4), o_mx = o_mn + al((size_t)nq * 4),
               o_qi = o_mx + al((size_t)nq * 4), o_fxy = o_qi + al((size_t)nq * 4), o_fl = o_fxy + al((size_t)nf * 8),
               o_fs = o_fl + al((size_t)nf * 4), o_fi = o_fs + al((size_t)nf), o_ci = o_fi + al((size_t)nf * 4),
               o_cd = o_ci + al((size_t)nq * PROJ_K * 4), o_cl = o_cd + al((size_t)nq * PROJ_K * 4),
               io_total = o_cl + al((size_t)nq * PROJ_K * 4);
  const size_t q_blk = o_fxy - o_uv, out_blk = io_total - o_ci;
  HFB_TRY(ctx->ensure_io(io_total));
  HFB_TRY(ctx->ensure_stage(q_blk + out_blk));
  uint8_t* io = reinterpret_cast<uint8_t*>(ctx->d_io);
  uint8_t* hs = reinterpret_cast<uint8_t*>(ctx->h_stage);
  cudaStream_t st = ctx->stream;
  memcpy(hs, q_uv, (size_t)nq * 8);
  memcpy(hs + (o_r - o_uv), q_radius, (size_t)nq * 4);
  memcpy(hs + (o_mn - o_uv), q_min_level, (size_t)nq * 4);
  memcpy(hs + (o_mx - o_uv), q_max_level, (size_t)nq * 4);
  if (q_res) memcpy(hs + (o_qi - o_uv), q_index, (size_t)nq * 4);
  HFB_CUDA(ctx, cudaMemcpyAsync(io + o_uv, hs, q_blk, cudaMemcpyHostToDevice, st));
  if (q_res) {
    proj_gather_rows_kernel<<<nq, 64, 0, st>>>(dQ_base, reinterpret_cast<const int*>(io + o_qi), nq,
                                               reinterpret_cast<float*>(io + oQ));
    HFB_CHECK_LAUNCH(ctx, "proj_gather");
  } else {
    HFB_CUDA(ctx, cudaMemcpyAsync(io + oQ, Q, (size_t)nq * 1024, cudaMemcpyHostToDevice, st));
  }
  const float* dF;
  const int* dFlev;
  if (f_res) {
    dF = dF_res;
    dFlev = d_flevel;
    proj_pack_xy_kernel<<<ceil_div(nf, 256), 256, 0, st>>>(d_fx, d_fy, nf, reinterpret_cast<float2*>(io + o_fxy));
    HFB_CHECK_LAUNCH(ctx, "proj_pack_xy");
  } else {
    HFB_CUDA(ctx, cudaMemcpyAsync(io + oF, F, (size_t)nf * 1024, cudaMemcpyHostToDevice, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(io + o_fxy, f_xy, (size_t)nf * 8, cudaMemcpyHostToDevice, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(io + o_fl, f_level, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    dF = reinterpret_cast<const float*>(io + oF);
    dFlev = reinterpret_cast<const int*>(io + o_fl);
  }
  if (f_skip) HFB_CUDA(ctx, cudaMemcpyAsync(io + o_fs, f_skip, (size_t)nf, cudaMemcpyHostToDevice, st));
  if (f_inv_sigma2) HFB_CUDA(ctx, cudaMemcpyAsync(io + o_fi, f_inv_sigma2, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
  // scratch: Q' | F' | hnq | hnf | qwin | qlev | records
  const size_t sQ = 0, sF = sQ + al((size_t)nq * MATCH_LD * 2), s_hq = sF + al((size_t)nf * MATCH_LD * 2),
               s_hf = s_hq + al((size_t)nq * 4), s_qw = s_hf + al((size_t)nf * 4), s_ql = s_qw + al((size_t)nq * 16),
               s_rec = s_ql + al((size_t)nq * 8), s_total = s_rec + al((size_t)n_tiles * nq * sizeof(ProjRec));
  HFB_TRY(ctx->ensure_scratch(s_total));
  uint8_t* sc = reinterpret_cast<uint8_t*>(ctx->d_scratch);
  const float* dQ = reinterpret_cast<const float*>(io + oQ);
  __half* Q2 = reinterpret_cast<__half*>(sc + sQ);
  __half* F2 = reinterpret_cast<__half*>(sc + sF);
  float* hnq = reinterpret_cast<float*>(sc + s_hq);
  float* hnf = reinterpret_cast<float*>(sc + s_hf);
  float4* qwin = reinterpret_cast<float4*>(sc + s_qw);
  int2* qlev = reinterpret_cast<int2*>(sc + s_ql);
  ProjRec* rec = reinterpret_cast<ProjRec*>(sc + s_rec);
  hfb_launch(ctx, match_prep_kernel, ceil_div(nq, 8), 256, 0, dQ, nq, Q2, hnq, 1);
  HFB_CHECK_LAUNCH(ctx, "match_prep(Q)");
  hfb_launch(ctx, match_prep_kernel, ceil_div(nf, 8), 256, 0, dF, nf, F2, hnf, 1);
  HFB_CHECK_LAUNCH(ctx, "match_prep(F)");
  proj_pack_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(reinterpret_cast<const float*>(io + o_uv),
                                                    reinterpret_cast<const float*>(io + o_r),
                                                    reinterpret_cast<const int*>(io + o_mn),
                                                    reinterpret_cast<const int*>(io + o_mx), nq, qwin, qlev);
  HFB_CHECK_LAUNCH(ctx, "proj_pack");
  CUtensorMap tmA, tmB;
  HFB_TRY(hfb_make_tmap_2d(ctx, &tmA, Q2, MATCH_LD, (uint64_t)nq, MATCH_LD * 2, 128));
  HFB_TRY(hfb_make_tmap_2d(ctx, &tmB, F2, MATCH_LD, (uint64_t)nf, MATCH_LD * 2, MATCH_BN));
  GemmGeom g;
  gemm_fill_geom(g, nq, nf, MATCH_K, MATCH_BN, 0);
  g.split3 = 1;
  g.stages = 3;
  g.epi_warp_bytes = 0;
  g.bias_bytes = 0;
  gemm_finish_geom(g, ceil_div(nq, 128));
  const size_t smem = gemm_smem_bytes(g.BN, g.stages, 0);
  static SmemOptIn optin;
  HFB_CUDA(ctx, optin.ensure(gemm_tc_kernel<EpiProjTopK>, ctx->device, smem));
  EpiProjTopK::Params ep{hnq, hnf, qwin, qlev, reinterpret_cast<const float2*>(io + o_fxy), dFlev,
                         f_skip ? io + o_fs : nullptr,
                         f_inv_sigma2 ? reinterpret_cast<const float*>(io + o_fi) : nullptr,
                         f_inv_sigma2 ? chi2_max : 3.402823466e38f, rec, nq};
  gemm_tc_kernel<EpiProjTopK><<<gemm_grid(g, ctx->n_sm, smem), GEMM_THREADS(4), smem, st>>>(tmA, tmB, g, ep);
  HFB_CHECK_LAUNCH(ctx, "proj_gemm_topk");
  hfb_launch(ctx, proj_finalize_kernel, ceil_div(nq, 8), 256, 0, dQ, dF, rec, n_tiles, nq, dFlev,
             reinterpret_cast<int*>(io + o_ci), reinterpret_cast<float*>(io + o_cd), reinterpret_cast<int*>(io + o_cl));
  HFB_CHECK_LAUNCH(ctx, "proj_finalize");
  HFB_CUDA(ctx, cudaMemcpyAsync(hs + q_blk, io + o_ci, out_blk, cudaMemcpyDeviceToHost, st));
  HFB_CUDA(ctx, cudaStreamSynchronize(st));
  memcpy(cand_idx, hs + q_blk, (size_t)nq * PROJ_K * 4);
  memcpy(cand_dist, hs + q_blk + (o_cd - o_ci), (size_t)nq * PROJ_K * 4);
  memcpy(cand_level, hs + q_blk + (o_cl - o_ci), (size_t)nq * PROJ_K * 4);
  return HFB_OK;
}

extern "C" int hfb_match_projection(hfb_ctx* ctx, const float* Q, int32_t nq, const float* q_uv, const float* q_radius,
                                    const int32_t* q_min_level, const int32_t* q_max_level, const float* F, int32_t nf,
                                    const float* f_xy, const int32_t* f_level, const uint8_t* f_skip, int32_t* cand_idx,
                                    float* cand_dist, int32_t* cand_level) {
  return hfb_match_projection_gated(ctx, Q, nq, q_uv, q_radius, q_min_level, q_max_level, F, nf, f_xy, f_level, f_skip,
                                    nullptr, 0.f, cand_idx, cand_dist, cand_level);
}

extern "C" int hfb_match_projection_gated(hfb_ctx* ctx, const float* Q, int32_t nq, const float* q_uv,
                                          const float* q_radius, const int32_t* q_min_level, const int32_t* q_max_level,
                                          const float* F, int32_t nf, const float* f_xy, const int32_t* f_level,
                                          const uint8_t* f_skip, const float* f_inv_sigma2, float chi2_max,
                                          int32_t* cand_idx, float* cand_dist, int32_t* cand_level) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, nq == 0 || nf == 0 || (Q && F), "null input");
  return proj_common(ctx, Q, nullptr, nullptr, nq, q_uv, q_radius, q_min_level, q_max_level, F, nullptr, nullptr, nullptr,
                     nullptr, nf, f_xy, f_level, f_skip, f_inv_sigma2, chi2_max, cand_idx, cand_dist, cand_level);
}

// SearchByProjection(CurrentFrame, LastFrame) on RESIDENT descriptors: the features are frame
// `frame_index` of the last extraction (descriptors, positions and octaves as the extraction left them in HBM), the
// queries are rows q_prev_index[i] of the previous frame of its stream (frame_index - 1 of the same call, or the
// descriptors carried over from the previous call).  Only the per-query window (24 bytes) goes up.
extern "C" int hfb_match_projection_frame(hfb_ctx* ctx, int32_t frame_index, const int32_t* q_prev_index, int32_t nq,
                                          const float* q_uv, const float* q_radius, const int32_t* q_min_level,
                                          const int32_t* q_max_level, int32_t nf, const uint8_t* f_skip,
                                          const float* f_inv_sigma2, float chi2_max, int32_t* cand_idx, float* cand_dist,
                                          int32_t* cand_level) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, ctx->last_batch > 0 && frame_index >= 0 && frame_index < ctx->last_batch, "no such resident frame");
  HFB_REQUIRE(ctx, nf >= 0 && nf <= ctx->kp_cap, "nf exceeds the frame's keypoint capacity");
  HFB_REQUIRE(ctx, !ctx->cam.on || ctx->kun_valid, "the resident frame has no undistorted coordinates (camera set after its extraction)");
  HFB_REQUIRE(ctx, nq == 0 || q_prev_index, "null query index");
  for (int i = 0; i < nq; ++i)
    HFB_REQUIRE(ctx, q_prev_index[i] >= 0 && q_prev_index[i] < ctx->kp_cap, "query index outside the previous frame");
  const long long prev_slot = ctx->stream_mode == 0 ? (long long)frame_index - 1 : (long long)frame_index - ctx->last_batch;
  const float* dQ_base = ctx->d_kdesc + prev_slot * ctx->kp_cap * HFB_DESC_DIM;
  const size_t fo = (size_t)frame_index * ctx->kp_cap;
  return proj_common(ctx, nullptr, q_prev_index, dQ_base, nq, q_uv, q_radius, q_min_level, q_max_level, nullptr,
                     ctx->d_kdesc + fo * HFB_DESC_DIM, (ctx->cam.on ? ctx->d_kxu : ctx->d_kx) + fo,
                     (ctx->cam.on ? ctx->d_kyu : ctx->d_ky) + fo, ctx->d_koct + fo, nf, nullptr, nullptr,
                     f_skip, f_inv_sigma2, chi2_max, cand_idx, cand_dist, cand_level);
}

// =============================================================================================== distinctive descriptors
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:331-400) for a ragged batch of map points: all-pairs
// Matcher::DescriptorDistance over a point's N observed descriptors, per-row median (sorted[int(0.5 * (N-1))]), first row
// with the smallest median.  One CTA per map point: a warp evaluates one pair at a time in the literal difference form
// (lane = 8 of the 256 components, butterfly sum), the N x N table lives in shared memory, thread i then finds row i's
// median by rank counting (no sort: rank(j) = #{k : d[k] < d[j]} + #{k < j : d[k] == d[j]}) and the CTA takes the first
// minimum.  N <= DD_MAXN per point (more observations than any keyframe window holds); larger points report -2.
#define DD_MAXN 128
__global__ void __launch_bounds__(256) distinctive_kernel(const float* __restrict__ desc, const int* __restrict__ offsets,
                                                          int* __restrict__ best_idx, float* __restrict__ best_med) {
  extern __shared__ float s_d[];   // [N][N]
  __shared__ float s_med[DD_MAXN];
  const int p = blockIdx.x;
  const int r0 = offsets[p], N = offsets[p + 1] - r0;
  if (N <= 0 || N > DD_MAXN) {
    if (threadIdx.x == 0) {
      best_idx[p] = N <= 0 ? -1 : -2;
      best_med[p] = 3.402823466e+38f;
    }
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* D = desc + (size_t)r0 * HFB_DESC_DIM;
  const int n_pairs = N * (N - 1) / 2;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_d[i * N + i] = 0.f;
  for (int q = warp; q < n_pairs; q += nw) {
    // pair index -> (i, j), i < j, rows enumerated i-major
    int i = 0, rem = q;
    while (rem >= N - 1 - i) {
      rem -= N - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    const float4* a = reinterpret_cast<const float4*>(D + (size_t)i * HFB_DESC_DIM) + lane * 2;
    const float4* b = reinterpret_cast<const float4*>(D + (size_t)j * HFB_DESC_DIM) + lane * 2;
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 x = __ldg(a + h), y = __ldg(b + h);
      float t = x.x - y.x; s = fmaf(t, t, s);
      t = x.y - y.y; s = fmaf(t, t, s);
      t = x.z - y.z; s = fmaf(t, t, s);
      t = x.w - y.w; s = fmaf(t, t, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float d = sqrtf(s);
      s_d[i * N + j] = d;
      s_d[j * N + i] = d;
    }
  }
  __syncthreads();
  const int target = (int)(0.5 * (N - 1));
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float* row = s_d + i * N;
    float med = 0.f;
    for (int j = 0; j < N; ++j) {
      const float v = row[j];
      int rank = 0;
      for (int k = 0; k < N; ++k) rank += (row[k] < v) || (row[k] == v && k < j);
      if (rank == target) med = v;
    }
    s_med[i] = med;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bm = 3.402823466e+38f;
    int bi = 0;
    for (int i = 0; i < N; ++i)
      if (s_med[i] < bm) {
        bm = s_med[i];
        bi = i;
      }
    best_idx[p] = bi;
    best_med[p] = bm;
  }
}

int launch_distinctive(hfb_ctx* ctx, const float* d_desc, const int* d_offsets, int n_points, int max_n, int* d_best_idx,
                       float* d_best_med) {
  if (n_points <= 0) return HFB_OK;
  const int n = std::min(std::max(max_n, 1), DD_MAXN);
  const size_t smem = (size_t)n * n * sizeof(float);
  static SmemOptIn optin;
  if (smem > 48 * 1024) HFB_CUDA(ctx, optin.ensure(distinctive_kernel, ctx->device, smem));
  distinctive_kernel<<<n_points, 256, smem, ctx->stream>>>(d_desc, d_offsets, d_best_idx, d_best_med);
  HFB_CHECK_LAUNCH(ctx, "distinctive");
  return HFB_OK;
}
