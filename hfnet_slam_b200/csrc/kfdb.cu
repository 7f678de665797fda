// Keyframe database place-recognition scan (replaces the per-KeyFrame loop of src/KeyFrameDatabase.cc:85-104, 177-192):
// score_i = max(0, 1 - ||q - d_i||_2) for EVERY row, best = max score, candidates = { i : score_i > max(floor,
// rel * best) }.  Rows live contiguously in HBM as fp32 [capacity][dim] (the reference scatters them over per-KeyFrame
// cv::Mat objects).  The scan is a pure HBM stream: dim*4 bytes per keyframe per query batch, the literal
// difference form (q - d).norm() of the reference in fp32, one warp per row, 8 x 16-byte loads in flight per lane.
#include <algorithm>
#include <unordered_set>

#include "common.cuh"
#include "gemm_core.cuh"

struct hfb_kfdb {
  hfb_ctx* ctx = nullptr;
  int dim = 0, capacity = 0, size = 0;
  float* d_rows = nullptr;
  float* d_scores = nullptr;     // [KFDB_QB][capacity] scores of the last scan
  float* d_query = nullptr;      // [KFDB_QB][dim]
  unsigned int* d_best = nullptr;  // [KFDB_QB] ordered-uint max score
  int* d_ncand = nullptr;        // [0] candidate count, [1] completion ticket of the single-query tail
  // d_best[0] and d_ncand[0..1] are ZERO at rest: the last kernel of a query (compact tail / shard finish) resets them, so
  // a query costs no memset launches
  int* d_cand_slot = nullptr;    // [capacity]
  float* d_cand_score = nullptr; // [capacity]
  std::vector<int64_t> ids;      // slot -> id
  std::vector<int64_t> tags;     // slot -> map id (KeyFrame::GetMap(), for clearMap)
  std::unordered_map<int64_t, int> slot_of;
  std::vector<float> h_scores;   // lazily fetched scores of the last query
  bool h_scores_valid = false;
  long long* d_ids = nullptr;    // [capacity] slot -> id on the device (shard records carry ids)
  float* h_query = nullptr;      // pinned staging of the host query
  int* h_head = nullptr;         // pinned: ncand | best bits | first KFDB_HEAD (slot, score) pairs of the candidate list
  // multi-query scan on the tensor cores (hfb_kfdb_query_batch)
  float* d_norm2 = nullptr;      // [capacity] |d|^2 of every row
  float* d_dots = nullptr;       // [capacity][KTC_QN] q . d of the current pass
  float* d_qbatch = nullptr;     // [KTC_QN][dim] padded query block
  float* d_qnorm2 = nullptr;     // [KTC_QN]
  unsigned int* d_min_d2 = nullptr;   // [KTC_QN] ordered-uint min approximate squared distance
  int* d_pair_n = nullptr;       // [2]: pairs marked for the exact re-evaluation, overflow flag
  int2* d_pairs = nullptr;       // [KTC_PAIR_CAP] (query, slot)
  float* d_pair_score = nullptr; // [KTC_PAIR_CAP]
  unsigned int* d_qbest = nullptr;    // [KTC_QN] exact best score bits
  int* d_qncand = nullptr;       // [KTC_QN]
  int* d_qcand_slot = nullptr;   // [KTC_QN][batch_cap]
  float* d_qcand_score = nullptr;
  int batch_cap = 0;
  // sharded database (row-shard by id % world): peer inboxes for the one-shot record exchange
  int rank = 0, world = 1, shard_k = 0;
  size_t rec_bytes = 0;
  uint8_t* d_inbox = nullptr;            // [2 parities][world][rec_bytes] + flags
  uint8_t** d_peer_tab = nullptr;        // device copy of peer_inbox
  uint8_t* peer_inbox[64] = {nullptr};   // this rank's view of every rank's inbox (own included)
  bool peer_opened[64] = {false};
  unsigned int epoch = 0;
  uint8_t* d_shard_out = nullptr;        // merged result: {best, count, overflow, pad, {score, pad, id}[world * k]}
  uint8_t* h_shard_out = nullptr;        // pinned
};

#define KFDB_QB 4  // queries scanned per pass over the rows
#define KTC_QN 64            // queries per tensor-core pass (UMMA N)
#define KTC_PAIR_CAP (1 << 18)
#define KFDB_HEAD 64         // candidates fetched together with the count (one synchronisation in the common case)

template <int QB>
__global__ void __launch_bounds__(256) kfdb_scan_kernel(const float* __restrict__ rows, int n, int dim,
                                                        const float* __restrict__ q, int nq, float* __restrict__ scores,
                                                        int score_stride, unsigned int* __restrict__ best) {
  extern __shared__ float s_q[];  // [QB][dim]
  pdl_launch_dependents();        // the successors wait (griddepcontrol.wait) for this grid to finish: only their launch overlaps
  pdl_wait();
  for (int i = threadIdx.x; i < QB * dim; i += blockDim.x) s_q[i] = (i / dim) < nq ? q[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * warps_per_block;
  const int nvec = dim >> 2;  // float4 per row
  float lbest[QB];
#pragma unroll
  for (int j = 0; j < QB; ++j) lbest[j] = 0.f;
  for (int r = gw; r < n; r += nwarps) {
    const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)r * dim);
    float acc[QB];
#pragma unroll
    for (int j = 0; j < QB; ++j) acc[j] = 0.f;
    for (int v0 = lane; v0 < nvec; v0 += 32 * 8) {
      float4 d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int v = v0 + 32 * u;
        d[u] = v < nvec ? __ldcs(rp + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int v = v0 + 32 * u;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < QB; ++j) {
            const float4 qq = *reinterpret_cast<const float4*>(s_q + j * dim + 4 * v);
            float t;
            t = qq.x - d[u].x; acc[j] = fmaf(t, t, acc[j]);
            t = qq.y - d[u].y; acc[j] = fmaf(t, t, acc[j]);
            t = qq.z - d[u].z; acc[j] = fmaf(t, t, acc[j]);
            t = qq.w - d[u].w; acc[j] = fmaf(t, t, acc[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < QB; ++j) {
      float a = acc[j];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
      const float sc = fmaxf(0.f, 1.f - sqrtf(a));
      if (lane == 0 && j < nq) scores[(size_t)j * score_stride + r] = sc;
      lbest[j] = fmaxf(lbest[j], sc);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < QB; ++j)
      if (j < nq) atomicMax(best + j, __float_as_uint(lbest[j]));  // scores >= 0: uint order == float order
  }
}

// The same score for one (query, row) pair with the query in global memory: identical accumulation order to
// kfdb_scan_kernel (lane-strided float4 columns, eight in flight, fma chain per lane, xor-shuffle tree), hence identical
// bits.  Whole warp.
__device__ __forceinline__ float kfdb_pair_score(const float4* __restrict__ rp, const float4* __restrict__ qp, int lane,
                                                 int nvec) {
  float acc = 0.f;
  for (int v0 = lane; v0 < nvec; v0 += 32 * 8) {
    float4 d[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int v = v0 + 32 * u;
      d[u] = v < nvec ? __ldg(rp + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int v = v0 + 32 * u;
      if (v < nvec) {
        const float4 qq = __ldg(qp + v);
        float t;
        t = qq.x - d[u].x; acc = fmaf(t, t, acc);
        t = qq.y - d[u].y; acc = fmaf(t, t, acc);
        t = qq.z - d[u].z; acc = fmaf(t, t, acc);
        t = qq.w - d[u].w; acc = fmaf(t, t, acc);
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  return fmaxf(0.f, 1.f - sqrtf(acc));
}

// =====================================================================================================================
// Multi-query scan (hfb_kfdb_query_batch): Q >= 2 queries against every row as ONE pass over the rows.  The 4-query
// CUDA-core kernel above is compute-bound beyond a handful of queries (64 queries = 29 row passes' worth of time), so
// the contraction q . d moves to the tensor cores: kind::tf32 UMMA, M = 128 rows x N = 64 queries, K = 4096 streamed as
// 128 k-blocks of 32 floats (TMA, 128B swizzle; the rows are read ONCE from HBM, the 1 MB query block re-read from L2).
// Work is split stream-K style over one persistent CTA per SM -- CTA c owns k-block units [c U, (c+1) U) of the
// (row tile, k-block) sequence, so the HBM stream is balanced to the k-block whatever the row count -- and partial tiles
// are combined with vector atomics into dots[row][64].
// tf32 reads 10 mantissa bits, so the scores from this pass only SELECT: |q.d - tf32(q).tf32(d)| <= 2^-9 |q||d| bounds
// the squared-distance error by E = 4.9e-3 |q||d|; every (query, row) pair that could be a candidate under that bound,
// or could be the best row, is re-scored exactly (kfdb_pair_score, the bits of the single-query path) and the
// candidate set / best score are decided on the exact values -- identical to Q single queries.
#define KTC_STAGES 8
#define KTC_THREADS 192
#define KTC_A_BYTES 16384
#define KTC_STAGE_BYTES (KTC_A_BYTES + KTC_QN * 128)
#define KTC_SMEM (1024 + KTC_STAGES * KTC_STAGE_BYTES + 256)
#define KTC_ERR 4.9e-3f

struct KtcGeom {
  int n_rows, kb_per_tile, n_tiles;
  long long total_units, units_per_cta;
  uint32_t idesc;
};

__global__ void __launch_bounds__(KTC_THREADS, 1)
kfdb_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KtcGeom g,
               float* __restrict__ dots) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KTC_STAGES * KTC_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + KTC_STAGES;
  uint64_t* acc_full = bars + 2 * KTC_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < KTC_STAGES; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 128);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const long long u_begin = (long long)blockIdx.x * g.units_per_cta;
  const long long u_end = u_begin + g.units_per_cta < g.total_units ? u_begin + g.units_per_cta : g.total_units;
  const int KB = g.kb_per_tile;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long long u = u_begin; u < u_end;) {
        const int tile = (int)(u / KB), kb0 = (int)(u - (long long)tile * KB);
        const int len = (int)((KB - kb0) < (u_end - u) ? (KB - kb0) : (u_end - u));
        for (int kb = kb0; kb < kb0 + len; ++kb, ++it) {
          const int s = it % KTC_STAGES;
          tc::mbar_wait(&empty[s], ((it / KTC_STAGES) & 1u) ^ 1u);
          uint8_t* sa = smem + (size_t)s * KTC_STAGE_BYTES;
          tc::mbar_expect_tx(&full[s], KTC_STAGE_BYTES);
          tc::tma_load_2d(sa, &tmA, &full[s], kb * 32, tile * 128);
          tc::tma_load_2d(sa + KTC_A_BYTES, &tmB, &full[s], kb * 32, 0);
        }
        u += len;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (long long u = u_begin; u < u_end;) {
        const int tile = (int)(u / KB), kb0 = (int)(u - (long long)tile * KB);
        const int len = (int)((KB - kb0) < (u_end - u) ? (KB - kb0) : (u_end - u));
        const uint32_t as = acc_it & 1u;
        tc::mbar_wait(&acc_empty[as], ((acc_it >> 1) & 1u) ^ 1u);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * KTC_QN;
        for (int kb = 0; kb < len; ++kb, ++it) {
          const int s = it % KTC_STAGES;
          tc::mbar_wait(&full[s], (it / KTC_STAGES) & 1u);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + (size_t)s * KTC_STAGE_BYTES);
          const uint64_t da = tc::make_sdesc_sw128(sa), db = tc::make_sdesc_sw128(sa + KTC_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_tf32(d_tmem, tc::sdesc_advance_k16(da, k), tc::sdesc_advance_k16(db, k), g.idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          tc::umma_commit(&empty[s]);
        }
        tc::umma_commit(&acc_full[as]);
        ++acc_it;
        u += len;
      }
    }
  } else {
    const int lg = warp & 3;
    uint32_t acc_it = 0;
    for (long long u = u_begin; u < u_end;) {
      const int tile = (int)(u / KB), kb0 = (int)(u - (long long)tile * KB);
      const int len = (int)((KB - kb0) < (u_end - u) ? (KB - kb0) : (u_end - u));
      const uint32_t as = acc_it & 1u;
      tc::mbar_wait(&acc_full[as], (acc_it >> 1) & 1u);
      __syncwarp();
      tc::fence_after_sync();
      const int row = tile * 128 + lg * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + as * KTC_QN;
      float4* out = reinterpret_cast<float4*>(dots + (size_t)row * KTC_QN);
      const bool whole = len == KB;       // the whole K range of this tile was ours: plain stores
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + 32u * h, r);
        tc::tmem_ld_wait();
        if (row < g.n_rows) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                         __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            if (whole) out[8 * h + q] = v;
            else atomicAdd(out + 8 * h + q, v);
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[as]);
      ++acc_it;
      u += len;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 128);
}

__global__ void kfdb_norm_kernel(const float* __restrict__ rows, int n, int dim, float* __restrict__ norm2) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  const float4* rp = reinterpret_cast<const float4*>(rows + (size_t)r * dim);
  float s = 0.f;
  for (int v = lane; v < (dim >> 2); v += 32) {
    const float4 d = __ldg(rp + v);
    s = fmaf(d.x, d.x, s); s = fmaf(d.y, d.y, s); s = fmaf(d.z, d.z, s); s = fmaf(d.w, d.w, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) norm2[r] = s;
}

// One launch prepares a pass: the padded query block (rows nq .. 63 zero), the query norms and the per-query state.
__global__ void __launch_bounds__(256) kfdb_batch_init_kernel(const float* __restrict__ queries, int nq, int dim,
                                                              float* __restrict__ qbatch, float* __restrict__ qn,
                                                              unsigned int* __restrict__ min_d2, int* __restrict__ pair_n,
                                                              unsigned int* __restrict__ qbest, int* __restrict__ qncand) {
  const int q = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __shared__ float s_part[8];
  const float4* src = reinterpret_cast<const float4*>(queries + (size_t)q * dim);
  float4* dst = reinterpret_cast<float4*>(qbatch + (size_t)q * dim);
  float s = 0.f;
  for (int v = threadIdx.x; v < (dim >> 2); v += blockDim.x) {
    const float4 d = q < nq ? __ldg(src + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    dst[v] = d;
    s = fmaf(d.x, d.x, s); s = fmaf(d.y, d.y, s); s = fmaf(d.z, d.z, s); s = fmaf(d.w, d.w, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) s_part[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_part[i];
    qn[q] = t;
    min_d2[q] = 0xFFFFFFFFu;
    qbest[q] = 0u;
    qncand[q] = 0;
    if (q == 0) {
      min_d2[KTC_QN] = 0u;
      pair_n[0] = 0;
      pair_n[1] = 0;
    }
  }
}

// approximate squared distance of (row, query) from the tensor-core pass
__device__ __forceinline__ float ktc_d2(const float* dots, const float* dn, const float* qn, int row, int q) {
  return qn[q] + dn[row] - 2.f * dots[(size_t)row * KTC_QN + q];
}

// per query: minimum approximate squared distance, and the largest row norm.  Block = 4 row lanes x 64 queries over 64
// rows: a warp reads 32 consecutive queries of one row (coalesced), 782 blocks at 50 k rows.
__global__ void __launch_bounds__(256) kfdb_tc_min_kernel(const float* __restrict__ dots, const float* __restrict__ dn,
                                                          const float* __restrict__ qn, int n, int nq,
                                                          unsigned int* __restrict__ min_d2) {
  __shared__ float s_m[4][KTC_QN];
  const int q = threadIdx.x & (KTC_QN - 1), rl = threadIdx.x >> 6;
  const int r0 = blockIdx.x * 64;
  const float myq = qn[q];
  float m = 3.0e38f, mx = 0.f;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int r = r0 + 4 * k + rl;
    if (r < n) {
      const float d = dn[r];
      m = fminf(m, myq + d - 2.f * dots[(size_t)r * KTC_QN + q]);
      mx = fmaxf(mx, d);
    }
  }
  s_m[rl][q] = m;
  if (q == 0) atomicMax(min_d2 + KTC_QN, __float_as_uint(mx));
  __syncthreads();
  if (rl == 0 && q < nq)
    atomicMin(min_d2 + q, f2ord(fminf(fminf(s_m[0][q], s_m[1][q]), fminf(s_m[2][q], s_m[3][q]))));
}

// pairs that may be candidates, or the best row, under the tf32 error bound -> exact re-evaluation list
__global__ void __launch_bounds__(256) kfdb_tc_mark_kernel(const float* __restrict__ dots, const float* __restrict__ dn,
                                                           const float* __restrict__ qn, int n, int nq,
                                                           const unsigned int* __restrict__ min_d2, float rel, float floor_,
                                                           int2* __restrict__ pairs, int* __restrict__ pair_n) {
  const int q = threadIdx.x & (KTC_QN - 1), rl = threadIdx.x >> 6;
  if (q >= nq) return;
  const int r0 = blockIdx.x * 64;
  const float myq = qn[q], qnorm = sqrtf(myq);
  const float dmin2 = ord2f(min_d2[q]);
  const float emax = KTC_ERR * qnorm * sqrtf(__uint_as_float(min_d2[KTC_QN]));   // bound for whichever row is the best
  // lowest the true best score can be, hence the most permissive threshold and the largest candidate distance
  const float best_lo = fmaxf(0.f, 1.f - sqrtf(fmaxf(dmin2 + emax, 0.f)));
  const float cut = 1.f - fmaxf(floor_, rel * best_lo);
  const float cut2 = cut * cut * 1.0001f, best_hi = dmin2 + emax;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int r = r0 + 4 * k + rl;
    if (r >= n) break;
    const float d = dn[r];
    const float lo = (myq + d - 2.f * dots[(size_t)r * KTC_QN + q]) - KTC_ERR * qnorm * sqrtf(d);   // least possible d^2
    if (lo < cut2 || lo <= best_hi) {
      const int p = atomicAdd(pair_n, 1);
      if (p < KTC_PAIR_CAP) pairs[p] = make_int2(q, r);
      else pair_n[1] = 1;
    }
  }
}

__global__ void kfdb_pair_score_kernel(const float* __restrict__ rows, int dim, const float* __restrict__ queries,
                                       const int2* __restrict__ pairs, const int* __restrict__ pair_n,
                                       float* __restrict__ pair_score, unsigned int* __restrict__ qbest) {
  const int np = min(pair_n[0], KTC_PAIR_CAP);
  const int lane = threadIdx.x & 31;
  for (int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < np; p += gridDim.x * (blockDim.x >> 5)) {
    const int2 pr = pairs[p];
    const float sc = kfdb_pair_score(reinterpret_cast<const float4*>(rows + (size_t)pr.y * dim),
                                     reinterpret_cast<const float4*>(queries + (size_t)pr.x * dim), lane, dim >> 2);
    if (lane == 0) {
      pair_score[p] = sc;
      atomicMax(qbest + pr.x, __float_as_uint(sc));
    }
  }
}

__global__ void kfdb_pair_select_kernel(const int2* __restrict__ pairs, const int* __restrict__ pair_n,
                                        const float* __restrict__ pair_score, const unsigned int* __restrict__ qbest,
                                        float rel, float floor_, int cap, int* __restrict__ ncand, int* __restrict__ cslot,
                                        float* __restrict__ cscore) {
  const int np = min(pair_n[0], KTC_PAIR_CAP);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const int2 pr = pairs[p];
    const float thr = fmaxf(floor_, __fmul_rn(__uint_as_float(qbest[pr.x]), rel));
    const float s = pair_score[p];
    if (s > thr) {
      const int k = atomicAdd(ncand + pr.x, 1);
      if (k < cap) {
        cslot[(size_t)pr.x * cap + k] = pr.y;
        cscore[(size_t)pr.x * cap + k] = s;
      }
    }
  }
}

// candidates of query 0: score > max(floor, rel*best), strict (KeyFrameDatabase.cc:98-104, 190-192)
// head != null (single-GPU query): the block that finishes last writes {count, best bits, first KFDB_HEAD slots, their
// scores} straight into the caller's page-locked host block (mapped memory: no copy operation) and puts best / count /
// ticket back to zero for the next query.
__global__ void kfdb_compact_kernel(const float* __restrict__ scores, int n, unsigned int* __restrict__ best,
                                    float rel, float floor_, int* __restrict__ ncand, int* __restrict__ slot,
                                    float* __restrict__ sc, int* __restrict__ head) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int bb = best[0];
  if (i < n) {
    const float thr = fmaxf(floor_, __fmul_rn(__uint_as_float(bb), rel));
    const float s = scores[i];
    if (s > thr) {
      const int p = atomicAdd(ncand, 1);
      slot[p] = i;
      sc[p] = s;
    }
  }
  if (!head) return;
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ncand + 1, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nc = *reinterpret_cast<volatile int*>(ncand);
  for (int j = threadIdx.x; j < min(nc, KFDB_HEAD); j += blockDim.x) {
    head[2 + j] = __ldcg(slot + j);
    head[2 + KFDB_HEAD + j] = __float_as_int(__ldcg(sc + j));
  }
  if (threadIdx.x == 0) {
    head[0] = nc;
    head[1] = (int)bb;
    best[0] = 0u;
    ncand[0] = 0;
    ncand[1] = 0;
  }
  __threadfence_system();
}

static int kfdb_scan(hfb_kfdb* db, const float* d_query, int nq, float* d_scores, int stride, unsigned int* d_best,
                     bool zero_best = true) {
  hfb_ctx* ctx = db->ctx;
  if (zero_best) HFB_CUDA(ctx, cudaMemsetAsync(d_best, 0, sizeof(unsigned int) * nq, ctx->stream));
  if (db->size == 0) return HFB_OK;
  const size_t smem = (size_t)KFDB_QB * db->dim * sizeof(float);
  static SmemOptIn optin_batch, optin_one;
  HFB_CUDA(ctx, optin_batch.ensure(kfdb_scan_kernel<KFDB_QB>, ctx->device, 200 * 1024));
  HFB_CUDA(ctx, optin_one.ensure(kfdb_scan_kernel<1>, ctx->device, 200 * 1024));
  for (int q0 = 0; q0 < nq; q0 += KFDB_QB) {
    const int nb = std::min(KFDB_QB, nq - q0);
    // grid: a multiple of the SM count, 8 warps per CTA, at most one warp per row
    int blocks = std::min(ctx->n_sm * 4, ceil_div(db->size, 8));
    if (blocks >= ctx->n_sm) blocks = blocks / ctx->n_sm * ctx->n_sm;
    if (nb == 1)
      hfb_launch(ctx, kfdb_scan_kernel<1>, dim3(blocks), dim3(256), (size_t)db->dim * 4, (const float*)db->d_rows, db->size,
                 db->dim, d_query + (size_t)q0 * db->dim, 1, d_scores + (size_t)q0 * stride, stride, d_best + q0);
    else
      hfb_launch(ctx, kfdb_scan_kernel<KFDB_QB>, dim3(blocks), dim3(256), smem, (const float*)db->d_rows, db->size, db->dim,
                 d_query + (size_t)q0 * db->dim, nb, d_scores + (size_t)q0 * stride, stride, d_best + q0);
    HFB_CHECK_LAUNCH(ctx, "kfdb_scan");
  }
  return HFB_OK;
}

static void kfdb_free(hfb_kfdb* db) {
  cudaFree(db->d_rows); cudaFree(db->d_scores); cudaFree(db->d_query); cudaFree(db->d_best);
  cudaFree(db->d_ncand); cudaFree(db->d_cand_slot); cudaFree(db->d_cand_score); cudaFree(db->d_ids);
  cudaFree(db->d_norm2); cudaFree(db->d_dots); cudaFree(db->d_qbatch); cudaFree(db->d_qnorm2); cudaFree(db->d_min_d2);
  cudaFree(db->d_pair_n); cudaFree(db->d_pairs); cudaFree(db->d_pair_score); cudaFree(db->d_qbest);
  cudaFree(db->d_qncand); cudaFree(db->d_qcand_slot); cudaFree(db->d_qcand_score);
  for (int r = 0; r < db->world; ++r)
    if (db->peer_opened[r] && db->peer_inbox[r]) cudaIpcCloseMemHandle(db->peer_inbox[r]);
  cudaFree(db->d_inbox); cudaFree(db->d_shard_out); cudaFree(db->d_peer_tab);
  if (db->h_shard_out) cudaFreeHost(db->h_shard_out);
  if (db->h_query) cudaFreeHost(db->h_query);
  if (db->h_head) cudaFreeHost(db->h_head);
  delete db;
}

extern "C" int hfb_kfdb_create(hfb_ctx* ctx, int32_t dim, int32_t capacity, hfb_kfdb** out) {
  if (!ctx || !out) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(ctx->device);
  *out = nullptr;
  HFB_REQUIRE(ctx, dim >= 4 && dim % 4 == 0 && dim <= 8192, "dim must be a multiple of 4 in [4, 8192]");
  HFB_REQUIRE(ctx, capacity >= 1, "capacity must be positive");
  hfb_kfdb* db = new hfb_kfdb();
  db->ctx = ctx;
  db->dim = dim;
  db->capacity = capacity;
  cudaError_t e = cudaMalloc(&db->d_rows, (size_t)capacity * dim * 4);
  if (e == cudaSuccess) e = cudaMalloc(&db->d_scores, (size_t)KFDB_QB * capacity * 4);
  if (e == cudaSuccess) e = cudaMalloc(&db->d_query, (size_t)KFDB_QB * dim * 4);
  if (e == cudaSuccess) e = cudaMalloc(&db->d_best, KFDB_QB * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc(&db->d_ncand, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(db->d_best, 0, KFDB_QB * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(db->d_ncand, 0, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&db->d_cand_slot, (size_t)capacity * 4);
  if (e == cudaSuccess) e = cudaMalloc(&db->d_cand_score, (size_t)capacity * 4);
  if (e == cudaSuccess) e = cudaMalloc(&db->d_ids, (size_t)capacity * 8);
  if (e == cudaSuccess) e = cudaMalloc(&db->d_norm2, (size_t)capacity * 4);
  if (e == cudaSuccess) e = cudaMallocHost(&db->h_query, (size_t)dim * 4);
  if (e == cudaSuccess) e = cudaMallocHost(&db->h_head, (2 + 2 * KFDB_HEAD) * 4);
  if (e != cudaSuccess) {
    ctx->set_error(std::string("hfb_kfdb_create: ") + cudaGetErrorString(e));
    kfdb_free(db);
    return HFB_ERR_CUDA;
  }
  db->ids.reserve(capacity);
  *out = db;
  return HFB_OK;
}

extern "C" void hfb_kfdb_destroy(hfb_kfdb* db) {
  if (!db) return;
  DeviceGuard _device_guard(db->ctx->device);
  cudaStreamSynchronize(db->ctx->stream);
  kfdb_free(db);
}

static int kfdb_add_common(hfb_kfdb* db, const int64_t* ids, const float* src, int n, cudaMemcpyKind kind,
                           const int64_t* map_ids = nullptr) {
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, ids && src && n >= 0, "bad argument");
  if (db->size + n > db->capacity) {
    ctx->set_error("keyframe database capacity exceeded");
    return HFB_ERR_CAPACITY;
  }
  {
    std::unordered_set<int64_t> seen;
    seen.reserve((size_t)n * 2);
    for (int i = 0; i < n; ++i) {
      if (db->slot_of.count(ids[i])) {
        ctx->set_error("keyframe id already in the database: " + std::to_string(ids[i]));
        return HFB_ERR_INVALID;
      }
      if (!seen.insert(ids[i]).second) {
        ctx->set_error("duplicate keyframe id in one add call");
        return HFB_ERR_INVALID;
      }
    }
  }
  if (n == 0) return HFB_OK;
  HFB_CUDA(ctx, cudaMemcpyAsync(db->d_rows + (size_t)db->size * db->dim, src, (size_t)n * db->dim * 4, kind, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(db->d_ids + db->size, ids, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  kfdb_norm_kernel<<<ceil_div(n, 8), 256, 0, ctx->stream>>>(db->d_rows + (size_t)db->size * db->dim, n, db->dim,
                                                           db->d_norm2 + db->size);
  HFB_CHECK_LAUNCH(ctx, "kfdb_norm");
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller may free src / ids
  for (int i = 0; i < n; ++i) {
    db->slot_of[ids[i]] = db->size + i;
    db->ids.push_back(ids[i]);
    db->tags.push_back(map_ids ? map_ids[i] : 0);
  }
  db->size += n;
  db->h_scores_valid = false;
  return HFB_OK;
}

extern "C" int hfb_kfdb_add(hfb_kfdb* db, const int64_t* ids, const float* descriptors, int32_t n) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  return kfdb_add_common(db, ids, descriptors, n, cudaMemcpyHostToDevice);
}
extern "C" int hfb_kfdb_add_dev(hfb_kfdb* db, const int64_t* ids, const float* d_descriptors, int32_t n) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  return kfdb_add_common(db, ids, d_descriptors, n, cudaMemcpyDeviceToDevice);
}

// KeyFrameDatabase::erase (src/KeyFrameDatabase.cc:38-43): the last row moves into the freed slot.
extern "C" int hfb_kfdb_erase(hfb_kfdb* db, int64_t id) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  auto it = db->slot_of.find(id);
  if (it == db->slot_of.end()) return HFB_OK;  // std::set::erase of a missing key is a no-op
  const int slot = it->second, last = db->size - 1;
  if (slot != last) {
    HFB_CUDA(ctx, cudaMemcpyAsync(db->d_rows + (size_t)slot * db->dim, db->d_rows + (size_t)last * db->dim,
                                  (size_t)db->dim * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(db->d_ids + slot, db->d_ids + last, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(db->d_norm2 + slot, db->d_norm2 + last, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    db->ids[slot] = db->ids[last];
    db->tags[slot] = db->tags[last];
    db->slot_of[db->ids[slot]] = slot;
  }
  db->ids.pop_back();
  db->tags.pop_back();
  db->slot_of.erase(it);
  db->size--;
  db->h_scores_valid = false;
  return HFB_OK;
}

extern "C" int hfb_kfdb_clear(hfb_kfdb* db) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  db->ids.clear();
  db->tags.clear();
  db->slot_of.clear();
  db->size = 0;
  db->h_scores_valid = false;
  return HFB_OK;
}

// KeyFrameDatabase::add with the keyframe's map recorded, and KeyFrameDatabase::clearMap (src/KeyFrameDatabase.cc:54-68):
// every keyframe whose map is `map_id` leaves the database.
extern "C" int hfb_kfdb_add_tagged(hfb_kfdb* db, const int64_t* ids, const int64_t* map_ids, const float* descriptors,
                                   int32_t n) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  HFB_REQUIRE(db->ctx, map_ids != nullptr, "null map ids");
  return kfdb_add_common(db, ids, descriptors, n, cudaMemcpyHostToDevice, map_ids);
}

extern "C" int hfb_kfdb_clear_map(hfb_kfdb* db, int64_t map_id) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  for (int slot = db->size - 1; slot >= 0; --slot)     // descending: the row moved into a freed slot was already visited
    if (db->tags[slot] == map_id) HFB_TRY(hfb_kfdb_erase(db, db->ids[slot]));
  return HFB_OK;
}

extern "C" int32_t hfb_kfdb_size(const hfb_kfdb* db) { return db ? db->size : 0; }

// Runs scan + candidate compaction for one host query; leaves candidates on the host sorted by ascending id.
static int kfdb_query_common(hfb_kfdb* db, const float* query, float rel, float floor_, std::vector<int64_t>& cid,
                             std::vector<float>& csc, float* best_score) {
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, query != nullptr, "null query");
  cid.clear();
  csc.clear();
  *best_score = 0.f;
  db->h_scores_valid = false;
  if (db->size == 0) return HFB_OK;
  memcpy(db->h_query, query, (size_t)db->dim * 4);      // page-locked staging: the H2D below is a plain DMA
  HFB_CUDA(ctx, cudaMemcpyAsync(db->d_query, db->h_query, (size_t)db->dim * 4, cudaMemcpyHostToDevice, ctx->stream));
  // three stream operations per query: the scan finds best[0] zeroed by the previous query's tail, the compaction's last
  // block writes count, best and the head of the candidate list into the page-locked block itself and re-zeroes the state
  HFB_TRY(kfdb_scan(db, db->d_query, 1, db->d_scores, db->capacity, db->d_best, false));
  hfb_launch(ctx, kfdb_compact_kernel, dim3(ceil_div(db->size, 256)), dim3(256), 0, (const float*)db->d_scores, db->size,
             db->d_best, rel, floor_, db->d_ncand, db->d_cand_slot, db->d_cand_score, db->h_head);
  HFB_CHECK_LAUNCH(ctx, "kfdb_compact");
  const int head = std::min(KFDB_HEAD, db->size);
  {
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {   // the zero-at-rest invariant may be broken: restore it before reporting
      cudaMemsetAsync(db->d_best, 0, KFDB_QB * sizeof(unsigned int), ctx->stream);
      cudaMemsetAsync(db->d_ncand, 0, 2 * sizeof(int), ctx->stream);
      ctx->set_error(std::string("kfdb query: ") + cudaGetErrorString(e));
      return HFB_ERR_CUDA;
    }
  }
  const int nc = db->h_head[0];
  memcpy(best_score, db->h_head + 1, 4);
  if (nc > 0) {
    std::vector<int> slots(nc);
    std::vector<float> sc(nc);
    if (nc <= head) {
      memcpy(slots.data(), db->h_head + 2, (size_t)nc * 4);
      memcpy(sc.data(), db->h_head + 2 + KFDB_HEAD, (size_t)nc * 4);
    } else {
      HFB_CUDA(ctx, cudaMemcpyAsync(slots.data(), db->d_cand_slot, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
      HFB_CUDA(ctx, cudaMemcpyAsync(sc.data(), db->d_cand_score, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
      HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    std::vector<int> order(nc);
    for (int i = 0; i < nc; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return db->ids[slots[a]] < db->ids[slots[b]]; });
    cid.resize(nc);
    csc.resize(nc);
    for (int i = 0; i < nc; ++i) {
      cid[i] = db->ids[slots[order[i]]];
      csc[i] = sc[order[i]];
    }
  }
  return HFB_OK;
}

extern "C" int hfb_kfdb_query(hfb_kfdb* db, const float* query, float rel, float floor_, int64_t* cand_ids,
                              float* cand_scores, int32_t cap, int32_t* n_cand, float* best_score) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, n_cand && best_score && cap >= 0, "bad argument");
  std::vector<int64_t> cid;
  std::vector<float> csc;
  HFB_TRY(kfdb_query_common(db, query, rel, floor_, cid, csc, best_score));
  *n_cand = (int)cid.size();
  const int w = std::min<int>(cap, (int)cid.size());
  for (int i = 0; i < w; ++i) {
    if (cand_ids) cand_ids[i] = cid[i];
    if (cand_scores) cand_scores[i] = csc[i];
  }
  if ((int)cid.size() > cap) {
    ctx->set_error("more candidates than the caller's capacity");
    return HFB_ERR_CAPACITY;
  }
  return HFB_OK;
}

extern "C" int hfb_kfdb_scores_of(hfb_kfdb* db, const int64_t* ids, int32_t n, float* scores) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, ids && scores && n >= 0, "bad argument");
  if (!db->h_scores_valid) {
    db->h_scores.resize(db->size);
    if (db->size > 0) {
      HFB_CUDA(ctx, cudaMemcpyAsync(db->h_scores.data(), db->d_scores, (size_t)db->size * 4, cudaMemcpyDeviceToHost,
                                    ctx->stream));
      HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    db->h_scores_valid = true;
  }
  for (int i = 0; i < n; ++i) {
    auto it = db->slot_of.find(ids[i]);
    scores[i] = it == db->slot_of.end() ? -1.f : db->h_scores[it->second];
  }
  return HFB_OK;
}

__global__ void kfdb_best_to_float_kernel(const unsigned int* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __uint_as_float(b[i]);
}

extern "C" int hfb_kfdb_scan_dev(hfb_kfdb* db, const float* d_query, int32_t n_queries, float* d_scores, float* d_best) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, d_query && d_scores && d_best && n_queries >= 1, "bad argument");
  // d_best doubles as the ordered-uint accumulator (scores >= 0, so the bit patterns are already floats)
  HFB_TRY(kfdb_scan(db, d_query, n_queries, d_scores, db->size, reinterpret_cast<unsigned int*>(d_best)));
  db->h_scores_valid = false;
  return HFB_OK;
}

// Fixed-size shard record for ONE all-gather; layout documented in include/hfnet_b200.h.
extern "C" int hfb_kfdb_query_shard(hfb_kfdb* db, const float* query, float rel, float floor_, int32_t k, void* record) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, record && k >= 1, "bad argument");
  std::vector<int64_t> cid;
  std::vector<float> csc;
  float best = 0.f;
  HFB_TRY(kfdb_query_common(db, query, rel, floor_, cid, csc, &best));
  std::vector<int> order(cid.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (csc[a] != csc[b]) return csc[a] > csc[b];
    return cid[a] < cid[b];
  });
  uint8_t* r = reinterpret_cast<uint8_t*>(record);
  memset(r, 0, 16 + 16 * (size_t)k);
  const int32_t count = (int32_t)std::min<size_t>(cid.size(), (size_t)k);
  const int32_t overflow = cid.size() > (size_t)k ? 1 : 0;
  memcpy(r, &best, 4);
  memcpy(r + 4, &count, 4);
  memcpy(r + 8, &overflow, 4);
  for (int i = 0; i < count; ++i) {
    uint8_t* e = r + 16 + 16 * (size_t)i;
    memcpy(e, &csc[order[i]], 4);
    memcpy(e + 8, &cid[order[i]], 8);
  }
  return HFB_OK;
}


// ---------------------------------------------------------------------------------------------------- multi-query
static int kfdb_batch_alloc(hfb_kfdb* db, int cap) {
  hfb_ctx* ctx = db->ctx;
  if (!db->d_dots) {
    HFB_CUDA(ctx, cudaMalloc(&db->d_dots, (size_t)(db->capacity + 128) * KTC_QN * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_qbatch, (size_t)KTC_QN * db->dim * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_qnorm2, KTC_QN * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_min_d2, (KTC_QN + 1) * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_pair_n, 2 * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_pairs, (size_t)KTC_PAIR_CAP * sizeof(int2)));
    HFB_CUDA(ctx, cudaMalloc(&db->d_pair_score, (size_t)KTC_PAIR_CAP * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_qbest, KTC_QN * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_qncand, KTC_QN * 4));
  }
  if (cap > db->batch_cap) {
    cudaFree(db->d_qcand_slot);
    cudaFree(db->d_qcand_score);
    db->d_qcand_slot = nullptr;
    db->d_qcand_score = nullptr;
    db->batch_cap = 0;
    HFB_CUDA(ctx, cudaMalloc(&db->d_qcand_slot, (size_t)KTC_QN * cap * 4));
    HFB_CUDA(ctx, cudaMalloc(&db->d_qcand_score, (size_t)KTC_QN * cap * 4));
    db->batch_cap = cap;
  }
  return HFB_OK;
}

// One pass of <= KTC_QN queries (device pointer, [nq][dim]): tensor-core scan, error-bounded marking, exact re-scoring,
// selection.  Leaves d_qbest / d_qncand / d_qcand_* filled.  No sync.
static int kfdb_batch_enqueue(hfb_kfdb* db, const float* d_queries, int nq, float rel, float floor_, int cap) {
  hfb_ctx* ctx = db->ctx;
  cudaStream_t st = ctx->stream;
  const int n = db->size, dim = db->dim;
  HFB_REQUIRE(ctx, dim % 32 == 0, "the multi-query scan needs dim to be a multiple of 32");
  HFB_TRY(kfdb_batch_alloc(db, cap));
  kfdb_batch_init_kernel<<<KTC_QN, 256, 0, st>>>(d_queries, nq, dim, db->d_qbatch, db->d_qnorm2, db->d_min_d2, db->d_pair_n,
                                                 db->d_qbest, db->d_qncand);
  HFB_CHECK_LAUNCH(ctx, "kfdb_batch_init");
  HFB_CUDA(ctx, cudaMemsetAsync(db->d_dots, 0, (size_t)n * KTC_QN * 4, st));
  CUtensorMap tmA, tmB;
  HFB_TRY(hfb_make_tmap_2d_f32(ctx, &tmA, db->d_rows, (uint64_t)dim, (uint64_t)n, (uint64_t)dim * 4, 128));
  HFB_TRY(hfb_make_tmap_2d_f32(ctx, &tmB, db->d_qbatch, (uint64_t)dim, (uint64_t)KTC_QN, (uint64_t)dim * 4, KTC_QN));
  KtcGeom g;
  g.n_rows = n;
  g.kb_per_tile = dim / 32;
  g.n_tiles = ceil_div(n, 128);
  g.total_units = (long long)g.n_tiles * g.kb_per_tile;
  const int grid = (int)std::min<long long>(ctx->n_sm, g.total_units);
  g.units_per_cta = (g.total_units + grid - 1) / grid;
  g.idesc = tc::make_idesc_tf32(KTC_QN);
  static SmemOptIn optin;
  HFB_CUDA(ctx, optin.ensure(kfdb_tc_kernel, ctx->device, KTC_SMEM));
  ctx->note("kfdb_tc_scan", (double)n * dim * 4, 2.0 * n * (double)dim * nq);
  kfdb_tc_kernel<<<grid, KTC_THREADS, KTC_SMEM, st>>>(tmA, tmB, g, db->d_dots);
  HFB_CHECK_LAUNCH(ctx, "kfdb_tc_scan");
  const int chunks = ceil_div(n, 64);
  kfdb_tc_min_kernel<<<chunks, 256, 0, st>>>(db->d_dots, db->d_norm2, db->d_qnorm2, n, nq, db->d_min_d2);
  HFB_CHECK_LAUNCH(ctx, "kfdb_tc_min");
  kfdb_tc_mark_kernel<<<chunks, 256, 0, st>>>(db->d_dots, db->d_norm2, db->d_qnorm2, n, nq, db->d_min_d2, rel, floor_,
                                                 db->d_pairs, db->d_pair_n);
  HFB_CHECK_LAUNCH(ctx, "kfdb_tc_mark");
  kfdb_pair_score_kernel<<<ctx->n_sm * 2, 256, 0, st>>>(db->d_rows, dim, db->d_qbatch, db->d_pairs, db->d_pair_n,
                                                       db->d_pair_score, db->d_qbest);
  HFB_CHECK_LAUNCH(ctx, "kfdb_pair_score");
  kfdb_pair_select_kernel<<<64, 256, 0, st>>>(db->d_pairs, db->d_pair_n, db->d_pair_score,
                                                                      db->d_qbest, rel, floor_, cap, db->d_qncand,
                                                                      db->d_qcand_slot, db->d_qcand_score);
  HFB_CHECK_LAUNCH(ctx, "kfdb_pair_select");
  return HFB_OK;
}

extern "C" int hfb_kfdb_query_batch_dev(hfb_kfdb* db, const float* d_queries, int32_t n_queries, float rel, float floor_) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, d_queries && n_queries >= 1, "bad argument");
  if (db->size == 0) return HFB_OK;
  for (int q0 = 0; q0 < n_queries; q0 += KTC_QN)
    HFB_TRY(kfdb_batch_enqueue(db, d_queries + (size_t)q0 * db->dim, std::min(KTC_QN, n_queries - q0), rel, floor_,
                               std::max(db->batch_cap, 64)));
  return HFB_OK;
}

extern "C" int hfb_kfdb_query_batch(hfb_kfdb* db, const float* queries, int32_t n_queries, float rel, float floor_,
                                    int32_t cap, int64_t* cand_ids, float* cand_scores, int32_t* n_cand,
                                    float* best_scores) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, queries && n_queries >= 1 && cap >= 1 && cand_ids && cand_scores && n_cand && best_scores, "bad argument");
  db->h_scores_valid = false;
  for (int q = 0; q < n_queries; ++q) {
    n_cand[q] = 0;
    best_scores[q] = 0.f;
  }
  if (db->size == 0) return HFB_OK;
  int rc = HFB_OK;
  const size_t qbytes = (size_t)KTC_QN * db->dim * 4;
  HFB_TRY(ctx->ensure_io(qbytes));
  std::vector<int> ncand(KTC_QN), slots;
  std::vector<unsigned int> best(KTC_QN);
  std::vector<float> sc;
  for (int q0 = 0; q0 < n_queries; q0 += KTC_QN) {
    const int nq = std::min(KTC_QN, n_queries - q0);
    HFB_CUDA(ctx, cudaMemcpyAsync(ctx->d_io, queries + (size_t)q0 * db->dim, (size_t)nq * db->dim * 4, cudaMemcpyHostToDevice,
                                  ctx->stream));
    HFB_TRY(kfdb_batch_enqueue(db, reinterpret_cast<const float*>(ctx->d_io), nq, rel, floor_, cap));
    int pn[2] = {0, 0};
    HFB_CUDA(ctx, cudaMemcpyAsync(ncand.data(), db->d_qncand, KTC_QN * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(best.data(), db->d_qbest, KTC_QN * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(pn, db->d_pair_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (pn[1]) {
      // more ambiguous pairs than the re-scoring list holds (a database of near-identical rows): the exact single-query
      // path answers this block
      for (int q = 0; q < nq; ++q) {
        int32_t nc = 0;
        int r = hfb_kfdb_query(db, queries + (size_t)(q0 + q) * db->dim, rel, floor_, cand_ids + (size_t)(q0 + q) * cap,
                               cand_scores + (size_t)(q0 + q) * cap, cap, &nc, best_scores + q0 + q);
        n_cand[q0 + q] = nc;
        if (r == HFB_ERR_CAPACITY) rc = r;
        else if (r != HFB_OK) return r;
      }
      continue;
    }
    slots.resize((size_t)KTC_QN * cap);
    sc.resize((size_t)KTC_QN * cap);
    HFB_CUDA(ctx, cudaMemcpyAsync(slots.data(), db->d_qcand_slot, (size_t)nq * cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(sc.data(), db->d_qcand_score, (size_t)nq * cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < nq; ++q) {
      memcpy(best_scores + q0 + q, &best[q], 4);
      n_cand[q0 + q] = ncand[q];
      const int w = std::min(ncand[q], cap);
      std::vector<int> order(w);
      for (int i = 0; i < w; ++i) order[i] = i;
      const int* sl = slots.data() + (size_t)q * cap;
      std::sort(order.begin(), order.end(), [&](int a, int b) { return db->ids[sl[a]] < db->ids[sl[b]]; });
      for (int i = 0; i < w; ++i) {
        cand_ids[(size_t)(q0 + q) * cap + i] = db->ids[sl[order[i]]];
        cand_scores[(size_t)(q0 + q) * cap + i] = sc[(size_t)q * cap + order[i]];
      }
      if (ncand[q] > cap) rc = HFB_ERR_CAPACITY;
    }
  }
  if (rc == HFB_ERR_CAPACITY) ctx->set_error("more candidates than the caller's per-query capacity");
  return rc;
}


// ---------------------------------------------------------------------------------------------------- sharded database
// Rows are sharded by id % world.  One query = every rank scans its shard, builds ONE fixed-size record
// (local best + its top-k rows above max(floor, rel * local best): a superset of its share of the global candidate set
// because local best <= global best), pushes it into every peer's inbox over NVLink (plain stores to peer memory mapped
// with CUDA IPC, then a system-scope flag), waits for the peers' flags and merges -- all inside one single-CTA kernel
// behind the scan, so the host sees one enqueue and one small D2H.  Inboxes are double-buffered by query parity and the
// flag carries the query epoch, so nothing is ever reset: a rank can only reach epoch e + 2 after every peer has
// published e + 1, i.e. after every peer has finished reading the e inbox.  All ranks must issue the same queries in
// the same order (collective semantics, like the all-gather it replaces).
struct ShardRecHdr { float best; int count; int overflow; int pad; };
struct ShardEntry { float score; int pad; long long id; };

__global__ void __launch_bounds__(256) kfdb_shard_finish_kernel(unsigned int* __restrict__ best, int* __restrict__ ncand,
                                                                const int* __restrict__ cand_slot, const float* __restrict__ cand_score,
                                                                const long long* __restrict__ ids, int k, float rel, float floor_,
                                                                int rank, int world, unsigned int epoch, uint8_t* const* peers,
                                                                uint8_t* my_inbox, int rec_bytes, uint8_t* out) {
  extern __shared__ uint8_t s_rec[];          // my record, then scratch
  __shared__ int s_count, s_timeout;
  __shared__ float s_best;
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  ShardRecHdr* hdr = reinterpret_cast<ShardRecHdr*>(s_rec);
  ShardEntry* ent = reinterpret_cast<ShardEntry*>(s_rec + 16);
  const int n = ncand[0];
  const unsigned int my_best = best[0];
  for (int i = tid; i < rec_bytes / 4; i += blockDim.x) reinterpret_cast<int*>(s_rec)[i] = 0;
  if (tid == 0) { s_count = 0; s_timeout = 0; }
  __syncthreads();
  if (n <= k) {
    for (int i = tid; i < n; i += blockDim.x) {
      ent[i].score = cand_score[i];
      ent[i].id = ids[cand_slot[i]];
    }
  } else {
    // top-k by (score descending, id ascending): rank counting over the shard's candidate list
    for (int i = tid; i < n; i += blockDim.x) {
      const float si = cand_score[i];
      const long long idi = ids[cand_slot[i]];
      int rk = 0;
      for (int j = 0; j < n; ++j) {
        const float sj = cand_score[j];
        rk += (sj > si) || (sj == si && ids[cand_slot[j]] < idi);
      }
      if (rk < k) {
        ent[rk].score = si;
        ent[rk].id = idi;
      }
    }
  }
  if (tid == 0) {
    hdr->best = __uint_as_float(my_best);
    hdr->count = min(n, k);
    hdr->overflow = n > k;
  }
  __syncthreads();
  if (tid == 0) {   // every read of the scan's state is done: zero at rest for the next query
    best[0] = 0u;
    ncand[0] = 0;
  }
  // push to every rank's inbox slot [parity][rank] (own inbox included), then publish the epoch
  const int parity = epoch & 1u;
  const size_t slot_off = ((size_t)parity * world + rank) * rec_bytes;
  const size_t flag_off = (size_t)2 * world * rec_bytes + ((size_t)parity * world + rank) * 4;
  for (int p = 0; p < world; ++p) {
    uint4* dst = reinterpret_cast<uint4*>(peers[p] + slot_off);
    const uint4* src = reinterpret_cast<const uint4*>(s_rec);
    for (int i = tid; i < rec_bytes / 16; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    volatile unsigned int* f = reinterpret_cast<volatile unsigned int*>(peers[tid] + flag_off);
    *f = epoch;
    __threadfence_system();
    // wait for rank `tid`'s record in MY inbox (bounded: a lost peer becomes an error, never a hung GPU)
    volatile unsigned int* mine = reinterpret_cast<volatile unsigned int*>(my_inbox + (size_t)2 * world * rec_bytes +
                                                                          ((size_t)parity * world + tid) * 4);
    unsigned int spins = 0;
    while (*mine != epoch) {
      __nanosleep(100);
      if (++spins > 20000000u) {
        atomicExch(&s_timeout, 1);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  // merge: global best, global threshold, the union of the records filtered by it
  const uint8_t* box = my_inbox + (size_t)parity * world * rec_bytes;
  if (tid == 0) {
    float b = 0.f;
    for (int p = 0; p < world; ++p) b = fmaxf(b, __int_as_float(__ldcg(reinterpret_cast<const int4*>(box + (size_t)p * rec_bytes)).x));
    s_best = b;
  }
  __syncthreads();
  const float thr = fmaxf(floor_, __fmul_rn(s_best, rel));
  ShardEntry* oent = reinterpret_cast<ShardEntry*>(out + 16);
  int overflow = 0;
  for (int i = tid; i < world * k; i += blockDim.x) {
    const int p = i / k, e = i - p * k;
    // peer-written memory: read through L2 (ld.cg), never through this SM's L1
    const int4 hw = __ldcg(reinterpret_cast<const int4*>(box + (size_t)p * rec_bytes));   // best | count | overflow | pad
    const int4* pe = reinterpret_cast<const int4*>(box + (size_t)p * rec_bytes + 16);
    if (e < hw.y) {
      const int4 ew = __ldcg(pe + e);
      if (__int_as_float(ew.x) > thr) {
        const int o = atomicAdd(&s_count, 1);
        reinterpret_cast<int4*>(oent)[o] = ew;
      }
    }
    // an overflowing shard matters only if its weakest listed row is still above the global bar
    if (e == 0 && hw.z) {
      float mn = 3.0e38f;
      for (int j = 0; j < hw.y; ++j) mn = fminf(mn, __int_as_float(__ldcg(pe + j).x));
      if (mn > thr) overflow = 1;
    }
  }
  overflow = __syncthreads_or(overflow);
  if (tid == 0) {
    ShardRecHdr* oh = reinterpret_cast<ShardRecHdr*>(out);
    oh->best = s_best;
    oh->count = s_count;
    oh->overflow = overflow;
    oh->pad = s_timeout;
  }
  __threadfence_system();   // `out` is page-locked host memory
}

extern "C" int hfb_kfdb_shard_setup(hfb_kfdb* db, int32_t rank, int32_t world, int32_t k, void* ipc_handle_out) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, world >= 1 && world <= 64 && rank >= 0 && rank < world && k >= 1 && k <= 1024, "bad shard geometry");
  HFB_REQUIRE(ctx, db->d_inbox == nullptr, "shard exchange already set up");
  db->rank = rank;
  db->world = world;
  db->shard_k = k;
  db->rec_bytes = 16 + 16 * (size_t)k;
  const size_t inbox_bytes = 2 * (size_t)world * db->rec_bytes + 2 * (size_t)world * 4 + 64;
  HFB_CUDA(ctx, cudaMalloc(&db->d_inbox, inbox_bytes));
  HFB_CUDA(ctx, cudaMemset(db->d_inbox, 0, inbox_bytes));
  const size_t out_bytes = 16 + 16 * (size_t)world * k;
  HFB_CUDA(ctx, cudaMalloc(&db->d_shard_out, out_bytes));
  HFB_CUDA(ctx, cudaMallocHost(&db->h_shard_out, out_bytes));
  HFB_CUDA(ctx, cudaMalloc(&db->d_peer_tab, 64 * sizeof(uint8_t*)));
  db->epoch = 0;
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    HFB_CUDA(ctx, cudaIpcGetMemHandle(&h, db->d_inbox));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out, &h, 64);
  }
  return HFB_OK;
}

static int kfdb_shard_publish_table(hfb_kfdb* db) {
  hfb_ctx* ctx = db->ctx;
  HFB_CUDA(ctx, cudaMemcpy(db->d_peer_tab, db->peer_inbox, 64 * sizeof(uint8_t*), cudaMemcpyHostToDevice));
  return HFB_OK;
}

// Multi-process: all_handles = the 64-byte handles of ranks 0 .. world-1 (gathered by the caller, e.g. torch.distributed).
extern "C" int hfb_kfdb_shard_connect(hfb_kfdb* db, const void* all_handles) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, db->d_inbox && all_handles, "hfb_kfdb_shard_setup first");
  for (int r = 0; r < db->world; ++r) {
    if (r == db->rank) {
      db->peer_inbox[r] = db->d_inbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const uint8_t*>(all_handles) + 64 * (size_t)r, 64);
    void* p = nullptr;
    HFB_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    db->peer_inbox[r] = reinterpret_cast<uint8_t*>(p);
    db->peer_opened[r] = true;
  }
  return kfdb_shard_publish_table(db);
}

// Same process (several shards on one device, or devices with peer access enabled by the caller): direct pointers.
extern "C" int hfb_kfdb_shard_connect_local(hfb_kfdb* db, hfb_kfdb* const* shards) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, db->d_inbox && shards, "hfb_kfdb_shard_setup first");
  for (int r = 0; r < db->world; ++r) {
    HFB_REQUIRE(ctx, shards[r] && shards[r]->d_inbox && shards[r]->world == db->world && shards[r]->shard_k == db->shard_k &&
                         shards[r]->rank == r, "peer shard not set up with the same geometry");
    db->peer_inbox[r] = shards[r]->d_inbox;
  }
  return kfdb_shard_publish_table(db);
}

extern "C" int hfb_kfdb_query_sharded_begin(hfb_kfdb* db, const float* query, float rel, float floor_) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, db->d_peer_tab && db->peer_inbox[db->rank], "shard exchange not connected");
  HFB_REQUIRE(ctx, query != nullptr, "null query");
  db->h_scores_valid = false;
  memcpy(db->h_query, query, (size_t)db->dim * 4);
  HFB_CUDA(ctx, cudaMemcpyAsync(db->d_query, db->h_query, (size_t)db->dim * 4, cudaMemcpyHostToDevice, ctx->stream));
  // four stream operations: H2D of the query, scan, compaction, finish (record exchange + merge; it writes the merged
  // list straight into the page-locked result block and puts best / count back to zero for the next query)
  HFB_TRY(kfdb_scan(db, db->d_query, 1, db->d_scores, db->capacity, db->d_best, false));
  if (db->size > 0) {
    hfb_launch(ctx, kfdb_compact_kernel, dim3(ceil_div(db->size, 256)), dim3(256), 0, (const float*)db->d_scores, db->size,
               db->d_best, rel, floor_, db->d_ncand, db->d_cand_slot, db->d_cand_score, (int*)nullptr);
    HFB_CHECK_LAUNCH(ctx, "kfdb_compact");
  }
  ++db->epoch;
  hfb_launch(ctx, kfdb_shard_finish_kernel, dim3(1), dim3(256), db->rec_bytes, db->d_best, db->d_ncand,
             (const int*)db->d_cand_slot, (const float*)db->d_cand_score, (const long long*)db->d_ids, db->shard_k, rel, floor_,
             db->rank, db->world, db->epoch, (uint8_t* const*)db->d_peer_tab, db->d_inbox, (int)db->rec_bytes, db->h_shard_out);
  HFB_CHECK_LAUNCH(ctx, "kfdb_shard_finish");
  return HFB_OK;
}

extern "C" int hfb_kfdb_query_sharded_end(hfb_kfdb* db, int64_t* cand_ids, float* cand_scores, int32_t cap, int32_t* n_cand,
                                          float* best_score, int32_t* overflow) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  hfb_ctx* ctx = db->ctx;
  HFB_REQUIRE(ctx, db->h_shard_out && n_cand && best_score && cap >= 0, "bad argument");
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const ShardRecHdr* h = reinterpret_cast<const ShardRecHdr*>(db->h_shard_out);
  if (h->pad) {
    ctx->set_error("sharded query: a peer rank did not publish its record in time");
    return HFB_ERR_STATE;
  }
  const ShardEntry* e = reinterpret_cast<const ShardEntry*>(db->h_shard_out + 16);
  std::vector<int> order(h->count);
  for (int i = 0; i < h->count; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return e[a].id < e[b].id; });
  *n_cand = h->count;
  *best_score = h->best;
  if (overflow) *overflow = h->overflow;
  const int w = std::min<int>(cap, h->count);
  for (int i = 0; i < w; ++i) {
    if (cand_ids) cand_ids[i] = e[order[i]].id;
    if (cand_scores) cand_scores[i] = e[order[i]].score;
  }
  if (h->count > cap) {
    ctx->set_error("more candidates than the caller's capacity");
    return HFB_ERR_CAPACITY;
  }
  return HFB_OK;
}

extern "C" int hfb_kfdb_query_sharded(hfb_kfdb* db, const float* query, float rel, float floor_, int64_t* cand_ids,
                                      float* cand_scores, int32_t cap, int32_t* n_cand, float* best_score, int32_t* overflow) {
  HFB_TRY(hfb_kfdb_query_sharded_begin(db, query, rel, floor_));
  return hfb_kfdb_query_sharded_end(db, cand_ids, cand_scores, cap, n_cand, best_score, overflow);
}
