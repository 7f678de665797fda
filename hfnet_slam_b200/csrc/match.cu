// 256-d local-descriptor matching (replaces the brute-force flavours of src/Matcher.cc): mutual nearest neighbours of
// two descriptor sets as ONE fused tensor-core contraction + arg-max per pair.
//
//   SearchForTriangulation  (src/Matcher.cc:845-889)  S = D1 * D2^T, row arg-max above 1 - 0.5*TH_HIGH^2, column
//                                                     cross-check, strict '>' (lowest index wins ties)
//   SearchByBoW             (src/Matcher.cc:220-263, :561-621)  cv::BFMatcher(NORM_L2, crossCheck).match, dist < TH_LOW
//
// fp32 descriptors are split into fp16 hi + lo parts (a = ah + al, |al| <= 2^-11 |a|) and the contraction runs over
// K' = 768 = [ah|ah|al] . [bh|bl|bh], i.e. ah.bh + ah.bl + al.bh: the dropped al.bl term is ~1e-8, so the scores
// that drive the arg-max carry fp32-level accuracy while using kind::f16 UMMA.  The epilogue reduces every 128 x 128
// accumulator tile to per-row and per-column (key, index) maxima straight out of TMEM (row: thread-local scan;
// column: redux.sync max + ballot inside each warp) and merges them with packed 64-bit atomicMax.  The accepted
// value (cosine or L2 distance) is recomputed in plain fp32 from the original descriptors before thresholding,
// like Matcher::DescriptorDistance (src/Matcher.cc:1893-1900).
#include "common.cuh"
#include "gemm_core.cuh"

#define MATCH_K 768
#define MATCH_BN 128

// ---- prep: fp32 [n][256] -> fp16 [n][768] (A: hi|hi|lo, B: hi|lo|hi) and half squared norms -------------------
__global__ void match_prep_kernel(const float* __restrict__ X, int n, __half* __restrict__ out, float* __restrict__ hn,
                                  int is_b, int l2_mode) {
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
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)row * MATCH_K) + lane;
  o[0] = hi;                     // k in [0,256)
  o[32] = is_b ? lo : hi;        // k in [256,512)
  o[64] = is_b ? hi : lo;        // k in [512,768)
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  if (lane == 0) hn[row] = l2_mode ? 0.5f * ss : 0.f;
}

__device__ __forceinline__ u64 pack_best(float key, int idx) {
  return ((u64)f2ord(key) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)idx);
}

// ---- epilogue: per-row and per-column arg-max of key = s - 0.5*|other|^2 (L2 mode) or s (cosine mode) ------------
struct EpiArgmax {
  static constexpr int kWarps = 4;
  struct Params {
    const float* hna;  // [na_total] 0.5*|a|^2 (0 in cosine mode)
    const float* hnb;  // [nb_total]
    u64* rowbest;      // [na_total], zero-initialised
    u64* colbest;      // [nb_total], zero-initialised
  };
  static __device__ __forceinline__ const float* bias(const Params&) { return nullptr; }
  static __device__ __forceinline__ void run(const Params& p, const GemmGeom& g, const TileRow& tr) {
    __shared__ u64 s_col[4][MATCH_BN];   // per-warp column maxima (no shared-memory atomics)
    __shared__ float s_hnb[MATCH_BN];
    const int lane = threadIdx.x & 31, warp = tr.ewarp, et = warp * 32 + lane;   // et: 0..127 among epilogue threads
    const int ncols = min(g.BN, tr.n_cnt - tr.n0);  // valid columns of this tile
    s_hnb[et] = et < ncols ? __ldg(p.hnb + tr.b_off + tr.n0 + et) : 0.f;
    epi_bar_sync();
    const float my_hna = tr.valid ? __ldg(p.hna + tr.row) : 0.f;
    const int warp_row0 = tr.row_local - lane;      // problem-local row of this warp's lane 0
    float best = -INFINITY;
    int best_j = 0;
    for (int c0 = 0; c0 < g.BN; c0 += 16) {
      uint32_t r[16];
      tc::tmem_ld16(tr.taddr + (uint32_t)c0, r);
      tc::tmem_ld_wait();
      if (c0 >= ncols) continue;  // warp-uniform
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s = __uint_as_float(r[j]);
        const bool cv = (c0 + j) < ncols;
        // row pass (this thread's row, ascending j: strict '>' keeps the lowest index on ties)
        const float kr = s - s_hnb[c0 + j];
        if (cv && tr.valid && kr > best) {
          best = kr;
          best_j = tr.n0 + c0 + j;
        }
        // column pass: max over the warp's 32 rows, lowest lane among equals
        const uint32_t kc = (cv && tr.valid) ? f2ord(s - my_hna) : 0u;
        const uint32_t mx = __reduce_max_sync(0xffffffffu, kc);
        const uint32_t who = __ballot_sync(0xffffffffu, kc == mx);
        if (lane == 0)
          s_col[warp][c0 + j] =
              mx != 0u ? (((u64)mx << 32) | (u64)(0xFFFFFFFFu - (uint32_t)(warp_row0 + __ffs(who) - 1))) : 0ull;
      }
    }
    if (tr.valid && best > -INFINITY) atomicMax(p.rowbest + tr.row, pack_best(best, best_j));
    epi_bar_sync();
    if (et < ncols) {
      u64 m = s_col[0][et];
      m = m > s_col[1][et] ? m : s_col[1][et];
      m = m > s_col[2][et] ? m : s_col[2][et];
      m = m > s_col[3][et] ? m : s_col[3][et];
      if (m != 0ull) atomicMax(p.colbest + tr.b_off + tr.n0 + et, m);
    }
    epi_bar_sync();   // s_col / s_hnb are reused by the next tile
  }
};

// ---- finalize: mutual check + exact fp32 value + threshold ----------------------------------------------------------
__global__ void match_finalize_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                      const u64* __restrict__ rowbest, const u64* __restrict__ colbest,
                                      const int* __restrict__ pair_tab, int n_pairs, int mode, float thr,
                                      int* __restrict__ match_idx, float* __restrict__ match_val,
                                      int* __restrict__ n_matches) {
  pdl_launch_dependents();
  pdl_wait();
  const int pr = blockIdx.y;
  const int a_off = pair_tab[pr], a_cnt = pair_tab[n_pairs + pr], b_off = pair_tab[2 * n_pairs + pr];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= a_cnt) return;
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

// Scratch layout (ctx->d_scratch): A' | B' | hna | hnb | rowbest | colbest | n_matches[n_pairs]
int launch_match_batch(hfb_ctx* ctx, int mode, const float* dA, const float* dB, int n_pairs, const int* d_pair_tab,
                       int max_a, int max_b, float thr, int* d_match_idx, float* d_match_val, int na_total,
                       int nb_total, int** d_n_matches_out) {
  if (na_total <= 0 || nb_total <= 0 || n_pairs <= 0 || max_a <= 0 || max_b <= 0) {
    if (na_total > 0) {
      HFB_CUDA(ctx, cudaMemsetAsync(d_match_idx, 0xFF, (size_t)na_total * 4, ctx->stream));
      HFB_CUDA(ctx, cudaMemsetAsync(d_match_val, 0, (size_t)na_total * 4, ctx->stream));
    }
    if (d_n_matches_out) *d_n_matches_out = nullptr;
    return HFB_OK;
  }
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  const size_t szA = al((size_t)na_total * MATCH_K * 2), szB = al((size_t)nb_total * MATCH_K * 2);
  const size_t szna = al((size_t)na_total * 4), sznb = al((size_t)nb_total * 4);
  const size_t szrb = al((size_t)na_total * 8), szcb = al((size_t)nb_total * 8), sznm = al((size_t)n_pairs * 4);
  HFB_TRY(ctx->ensure_scratch(szA + szB + szna + sznb + szrb + szcb + sznm));
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->d_scratch);
  __half* A2 = reinterpret_cast<__half*>(base);
  __half* B2 = reinterpret_cast<__half*>(base + szA);
  float* hna = reinterpret_cast<float*>(base + szA + szB);
  float* hnb = reinterpret_cast<float*>(base + szA + szB + szna);
  u64* rowbest = reinterpret_cast<u64*>(base + szA + szB + szna + sznb);
  u64* colbest = reinterpret_cast<u64*>(base + szA + szB + szna + sznb + szrb);
  int* nm = reinterpret_cast<int*>(base + szA + szB + szna + sznb + szrb + szcb);
  HFB_CUDA(ctx, cudaMemsetAsync(rowbest, 0, szrb + szcb + sznm, ctx->stream));

  const int l2 = (mode == 0);
  hfb_launch(ctx, match_prep_kernel, ceil_div(na_total, 8), 256, 0, dA, na_total, A2, hna, 0, l2);
  HFB_CHECK_LAUNCH(ctx, "match_prep(A)");
  hfb_launch(ctx, match_prep_kernel, ceil_div(nb_total, 8), 256, 0, dB, nb_total, B2, hnb, 1, l2);
  HFB_CHECK_LAUNCH(ctx, "match_prep(B)");

  CUtensorMap tmA, tmB;
  HFB_TRY(hfb_make_tmap_2d(ctx, &tmA, A2, MATCH_K, (uint64_t)na_total, MATCH_K * 2, 128));
  HFB_TRY(hfb_make_tmap_2d(ctx, &tmB, B2, MATCH_K, (uint64_t)nb_total, MATCH_K * 2, MATCH_BN));
  GemmGeom g;
  gemm_fill_geom(g, max_a, max_b, MATCH_K, MATCH_BN, 0);
  g.pair_tab = d_pair_tab;
  g.n_pairs = n_pairs;
  g.stages = 3;   // 96 KB ring: two persistent CTAs per SM (256 TMEM columns each)
  g.epi_warp_bytes = 0;
  g.bias_bytes = 0;
  gemm_finish_geom(g, ceil_div(max_a, 128));
  const size_t smem = gemm_smem_bytes(g.BN, g.stages, 0);
  static bool configured = false;
  if (!configured) {
    HFB_CUDA(ctx, cudaFuncSetAttribute(gemm_tc_kernel<EpiArgmax>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    configured = true;
  }
  EpiArgmax::Params ep{hna, hnb, rowbest, colbest};
  hfb_launch(ctx, gemm_tc_kernel<EpiArgmax>, gemm_grid(g, ctx->n_sm, smem), GEMM_THREADS(4), smem, tmA, tmB, g, ep);
  HFB_CHECK_LAUNCH(ctx, "match_gemm_argmax");

  dim3 fgrid(ceil_div(max_a, 8), n_pairs);
  hfb_launch(ctx, match_finalize_kernel, fgrid, 256, 0, dA, dB, rowbest, colbest, d_pair_tab, n_pairs, mode, thr,
                                                        d_match_idx, d_match_val, nm);
  HFB_CHECK_LAUNCH(ctx, "match_finalize");
  if (d_n_matches_out) *d_n_matches_out = nm;
  return HFB_OK;
}
