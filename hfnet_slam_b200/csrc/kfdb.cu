// Keyframe database place-recognition scan (replaces the per-KeyFrame loop of src/KeyFrameDatabase.cc:85-104, 177-192):
// score_i = max(0, 1 - ||q - d_i||_2) for EVERY row, best = max score, candidates = { i : score_i > max(floor,
// rel * best) }.  Rows live contiguously in HBM as fp32 [capacity][dim] (the reference scatters them over per-KeyFrame
// cv::Mat objects).  The scan is a pure HBM stream: dim*4 bytes per keyframe per query batch, the literal
// difference form (q - d).norm() of the reference in fp32, one warp per row, 8 x 16-byte loads in flight per lane.
#include <algorithm>
#include <unordered_set>

#include "common.cuh"

struct hfb_kfdb {
  hfb_ctx* ctx = nullptr;
  int dim = 0, capacity = 0, size = 0;
  float* d_rows = nullptr;
  float* d_scores = nullptr;     // [KFDB_QB][capacity] scores of the last scan
  float* d_query = nullptr;      // [KFDB_QB][dim]
  unsigned int* d_best = nullptr;  // [KFDB_QB] ordered-uint max score
  int* d_ncand = nullptr;
  int* d_cand_slot = nullptr;    // [capacity]
  float* d_cand_score = nullptr; // [capacity]
  std::vector<int64_t> ids;      // slot -> id
  std::unordered_map<int64_t, int> slot_of;
  std::vector<float> h_scores;   // lazily fetched scores of the last query
  bool h_scores_valid = false;
};

#define KFDB_QB 4  // queries scanned per pass over the rows

template <int QB>
__global__ void __launch_bounds__(256) kfdb_scan_kernel(const float* __restrict__ rows, int n, int dim,
                                                        const float* __restrict__ q, int nq, float* __restrict__ scores,
                                                        int score_stride, unsigned int* __restrict__ best) {
  extern __shared__ float s_q[];  // [QB][dim]
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

// candidates of query 0: score > max(floor, rel*best), strict (KeyFrameDatabase.cc:98-104, 190-192)
__global__ void kfdb_compact_kernel(const float* __restrict__ scores, int n, const unsigned int* __restrict__ best,
                                    float rel, float floor_, int* __restrict__ ncand, int* __restrict__ slot,
                                    float* __restrict__ sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float b = __uint_as_float(best[0]);
  const float thr = fmaxf(floor_, __fmul_rn(b, rel));
  const float s = scores[i];
  if (s > thr) {
    const int p = atomicAdd(ncand, 1);
    slot[p] = i;
    sc[p] = s;
  }
}

static int kfdb_scan(hfb_kfdb* db, const float* d_query, int nq, float* d_scores, int stride, unsigned int* d_best) {
  hfb_ctx* ctx = db->ctx;
  HFB_CUDA(ctx, cudaMemsetAsync(d_best, 0, sizeof(unsigned int) * nq, ctx->stream));
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
      kfdb_scan_kernel<1><<<blocks, 256, (size_t)db->dim * 4, ctx->stream>>>(
          db->d_rows, db->size, db->dim, d_query + (size_t)q0 * db->dim, 1, d_scores + (size_t)q0 * stride, stride,
          d_best + q0);
    else
      kfdb_scan_kernel<KFDB_QB><<<blocks, 256, smem, ctx->stream>>>(db->d_rows, db->size, db->dim,
                                                                   d_query + (size_t)q0 * db->dim, nb,
                                                                   d_scores + (size_t)q0 * stride, stride, d_best + q0);
    HFB_CHECK_LAUNCH(ctx, "kfdb_scan");
  }
  return HFB_OK;
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
  if (e == cudaSuccess) e = cudaMalloc(&db->d_ncand, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&db->d_cand_slot, (size_t)capacity * 4);
  if (e == cudaSuccess) e = cudaMalloc(&db->d_cand_score, (size_t)capacity * 4);
  if (e != cudaSuccess) {
    ctx->set_error(std::string("hfb_kfdb_create: ") + cudaGetErrorString(e));
    cudaFree(db->d_rows); cudaFree(db->d_scores); cudaFree(db->d_query); cudaFree(db->d_best);
    cudaFree(db->d_ncand); cudaFree(db->d_cand_slot); cudaFree(db->d_cand_score);
    delete db;
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
  cudaFree(db->d_rows); cudaFree(db->d_scores); cudaFree(db->d_query); cudaFree(db->d_best);
  cudaFree(db->d_ncand); cudaFree(db->d_cand_slot); cudaFree(db->d_cand_score);
  delete db;
}

static int kfdb_add_common(hfb_kfdb* db, const int64_t* ids, const float* src, int n, cudaMemcpyKind kind) {
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
  if (kind == cudaMemcpyHostToDevice) HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller may free src
  for (int i = 0; i < n; ++i) {
    db->slot_of[ids[i]] = db->size + i;
    db->ids.push_back(ids[i]);
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
    db->ids[slot] = db->ids[last];
    db->slot_of[db->ids[slot]] = slot;
  }
  db->ids.pop_back();
  db->slot_of.erase(it);
  db->size--;
  db->h_scores_valid = false;
  return HFB_OK;
}

extern "C" int hfb_kfdb_clear(hfb_kfdb* db) {
  if (!db) return HFB_ERR_INVALID;
  DeviceGuard _device_guard(db->ctx->device);
  db->ids.clear();
  db->slot_of.clear();
  db->size = 0;
  db->h_scores_valid = false;
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
  HFB_CUDA(ctx, cudaMemcpyAsync(db->d_query, query, (size_t)db->dim * 4, cudaMemcpyHostToDevice, ctx->stream));
  HFB_TRY(kfdb_scan(db, db->d_query, 1, db->d_scores, db->capacity, db->d_best));
  HFB_CUDA(ctx, cudaMemsetAsync(db->d_ncand, 0, sizeof(int), ctx->stream));
  kfdb_compact_kernel<<<ceil_div(db->size, 256), 256, 0, ctx->stream>>>(db->d_scores, db->size, db->d_best, rel, floor_,
                                                                       db->d_ncand, db->d_cand_slot, db->d_cand_score);
  HFB_CHECK_LAUNCH(ctx, "kfdb_compact");
  int nc = 0;
  unsigned int bu = 0;
  HFB_CUDA(ctx, cudaMemcpyAsync(&nc, db->d_ncand, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaMemcpyAsync(&bu, db->d_best, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(best_score, &bu, 4);
  if (nc > 0) {
    std::vector<int> slots(nc);
    std::vector<float> sc(nc);
    HFB_CUDA(ctx, cudaMemcpyAsync(slots.data(), db->d_cand_slot, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaMemcpyAsync(sc.data(), db->d_cand_score, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
