// Tensor-map construction, the store / L2-normalise / softmax+depth_to_space epilogues of the HF-Net encoder GEMMs
// and their launchers (kernel template in gemm_core.cuh).
#include <atomic>
#include "common.cuh"
#include "gemm_core.cuh"

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode(hfb_ctx* ctx) {
  static std::atomic<PFN_encodeTiled> cached{nullptr};   // same value from every thread; atomic so the publication is defined
  if (PFN_encodeTiled fn = cached.load(std::memory_order_acquire)) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    ctx->set_error("cuTensorMapEncodeTiled entry point not available");
    return nullptr;
  }
  cached.store(reinterpret_cast<PFN_encodeTiled>(p), std::memory_order_release);
  return reinterpret_cast<PFN_encodeTiled>(p);
}

// fp16 [outer][inner] matrix, rows `row_stride_bytes` apart; box = 64 (inner) x box_outer, 128B swizzle, OOB -> 0.
int hfb_make_tmap_2d(hfb_ctx* ctx, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                     uint64_t row_stride_bytes, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode(ctx);
  if (!enc) return HFB_ERR_CUDA;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled(2d) failed: code " + std::to_string((int)r) + " inner " +
                   std::to_string(inner) + " outer " + std::to_string(outer) + " stride " +
                   std::to_string(row_stride_bytes) + " box_outer " + std::to_string(box_outer));
    return HFB_ERR_CUDA;
  }
  return HFB_OK;
}

// fp32 [outer][inner] matrix (kind::tf32 operands read fp32 words); box = 32 (inner, 128 bytes) x box_outer, 128B swizzle.
int hfb_make_tmap_2d_f32(hfb_ctx* ctx, CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                         uint64_t row_stride_bytes, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode(ctx);
  if (!enc) return HFB_ERR_CUDA;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled(2d, f32) failed: code " + std::to_string((int)r) + " inner " +
                   std::to_string(inner) + " outer " + std::to_string(outer));
    return HFB_ERR_CUDA;
  }
  return HFB_OK;
}

int hfb_conv_halo() {
  static const int v = [] {
    const char* e = getenv("HFB_CONV_HALO");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v;
}

// fp16 NHWC tensor viewed as (C, W, H, B); box = 64 channels x 16 x 8 x 1 (one 128-pixel patch), or x 10 rows (the patch
// with its upper and lower halo row) in halo mode.
int hfb_make_tmap_nhwc(hfb_ctx* ctx, CUtensorMap* out, const void* base, int C, int W, int H, int B) {
  PFN_encodeTiled enc = get_encode(ctx);
  if (!enc) return HFB_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  cuuint32_t box[4] = {64, 16, (cuuint32_t)(hfb_conv_halo() ? 10 : 8), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled(4d) failed: code " + std::to_string((int)r));
    return HFB_ERR_CUDA;
  }
  return HFB_OK;
}

// fp16 NHWC tensor viewed as (C, W, H, B); box = box_c channels (64: 128B swizzle, 32: 64B swizzle) x box_w x box_h x 1
// (input halo tile of a fused block; out-of-range coordinates, negative ones included, are zero-filled).
int hfb_make_tmap_nhwc_box(hfb_ctx* ctx, CUtensorMap* out, const void* base, int C, int W, int H, int B, int box_w,
                           int box_h, int box_c) {
  PFN_encodeTiled enc = get_encode(ctx);
  if (!enc) return HFB_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled(4d box) failed: code " + std::to_string((int)r) + " box " +
                   std::to_string(box_w) + "x" + std::to_string(box_h));
    return HFB_ERR_CUDA;
  }
  return HFB_OK;
}

// ------------------------------------------------------------------------------------------------ epilogues
// bias (+ReLU6) (+residual) (+L2 normalisation of the whole row) -> fp16 or fp32 rows.  Each epilogue warp owns 32 tile
// rows; it stages 64-column slabs in its private shared-memory area and writes them out with lanes running along the
// output row, so global stores are full, contiguous 16-byte pieces instead of one piece per thread-row.
#define EPI_SLAB 32
#define EPI_PITCH (EPI_SLAB * 4 + 16)              // bytes per staged row: odd multiple of 16 -> conflict-free
#define EPI_WARP_BYTES (32 * EPI_PITCH)

struct EpiStore {
  static constexpr int kWarps = 8;   // two warps per TMEM lane group alternate over the 32-column slabs
  struct Params {
    void* out;
    int ldo;        // elements per output row
    int col_off;    // first output column of this GEMM inside the row
    const float* bias;
    const __half* residual;  // [row][ldr] or null
    int ldr;
    int relu6;
    int f32;
    int l2norm;     // tf.nn.l2_normalize over the N columns (needs BN == N): descriptor head, hf_net.py:78-80
  };
  static __device__ __forceinline__ const float* bias(const Params& p) { return p.bias; }

  static __device__ __forceinline__ void run(const Params& p, const GemmGeom& g, const TileRow& tr) {
    const int lane = threadIdx.x & 31;
    const int esz = (p.f32 || p.residual) ? 4 : 2;   // staging element size
    const int pitch = (int)(g.epi_warp_bytes >> 5);   // bytes per staged row: 144 (fp32 slabs) or 80 (fp16 slabs), odd multiples of 16
    uint8_t* my = tr.stage + (size_t)lane * pitch;
    const int orow = tr.valid ? (int)tr.row : -1;     // rows < 2^31 (B*H*W pixels)
    float inv = 1.f;
    if (p.l2norm) {
      float ss = 0.f;
      for (int c0 = 0; c0 < g.N; c0 += 16) {
        uint32_t r[16];
        tc::tmem_ld16(tr.taddr + (uint32_t)c0, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = __uint_as_float(r[j]) + tr.s_bias[c0 + j];
          ss = fmaf(v, v, ss);
        }
      }
      inv = rsqrtf(fmaxf(ss, 1e-12f));
    }
    const int ncols = min(g.BN, g.N - tr.n0);                 // valid columns of this tile (multiple of 8)
    // Fast path of the common case -- fp16 output, staged bias, whole 32-column slabs (the 3x3 head conv and the 1x1
    // convolutions of the generic block path): one TMEM wait per slab, vector bias loads, ReLU6 folded into the fp16 pack
    // (cvt.rn.relu.f16x2 + one half2 min: rounding is monotonic and 6 is exact in fp16, so the result equals clamping in
    // fp32 first), destination rows fetched once.  ~110 instead of ~880 warp instructions per slab.
    const bool fast = esz == 2 && !p.l2norm && tr.s_bias != nullptr && (ncols & (EPI_SLAB - 1)) == 0;
    if (fast) {
      const int ch = lane & 3;
      int drow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) drow[i] = __shfl_sync(0xffffffffu, orow, (lane >> 2) + 8 * i);
      const __half2 six2 = __float2half2_rn(6.f);
      for (int s0 = tr.sub * EPI_SLAB; s0 < ncols; s0 += 2 * EPI_SLAB) {
        uint32_t r[32];
        uint32_t(&lo)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
        uint32_t(&hi)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[16]);
        tc::tmem_ld16(tr.taddr + (uint32_t)s0, lo);
        tc::tmem_ld16(tr.taddr + (uint32_t)s0 + 16, hi);
        tc::tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(tr.s_bias + tr.n0 + s0);
        uint4* d = reinterpret_cast<uint4*>(my);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 b = b4[2 * q + h];
            const float x0 = __uint_as_float(r[8 * q + 4 * h]) + b.x, x1 = __uint_as_float(r[8 * q + 4 * h + 1]) + b.y;
            const float x2 = __uint_as_float(r[8 * q + 4 * h + 2]) + b.z, x3 = __uint_as_float(r[8 * q + 4 * h + 3]) + b.w;
            if (p.relu6) {                                   // warp-uniform
              uint32_t u0, u1;
              asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(u0) : "f"(x1), "f"(x0));
              asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(u1) : "f"(x3), "f"(x2));
              const __half2 m0 = __hmin2(*reinterpret_cast<__half2*>(&u0), six2), m1 = __hmin2(*reinterpret_cast<__half2*>(&u1), six2);
              w[2 * h] = *reinterpret_cast<const uint32_t*>(&m0);
              w[2 * h + 1] = *reinterpret_cast<const uint32_t*>(&m1);
            } else {
              const __half2 m0 = __floats2half2_rn(x0, x1), m1 = __floats2half2_rn(x2, x3);
              w[2 * h] = *reinterpret_cast<const uint32_t*>(&m0);
              w[2 * h + 1] = *reinterpret_cast<const uint32_t*>(&m1);
            }
          }
          d[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
        __half* obase = reinterpret_cast<__half*>(p.out) + (long long)p.col_off + tr.n0 + s0 + ch * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {                       // row (lane >> 2) + 8 i, 16-byte chunk ch of its 64 bytes
          if (drow[i] < 0) continue;
          const uint4 q4 = *reinterpret_cast<const uint4*>(tr.stage + (size_t)((lane >> 2) + 8 * i) * pitch + (size_t)ch * 16);
          *reinterpret_cast<uint4*>(obase + (long long)drow[i] * p.ldo) = q4;
        }
        __syncwarp();
      }
      return;
    }
    for (int s0 = tr.sub * EPI_SLAB; s0 < ncols; s0 += 2 * EPI_SLAB) {
      const int scols = min(EPI_SLAB, ncols - s0);             // 8, 16, 24 or 32
#pragma unroll
      for (int c0 = 0; c0 < EPI_SLAB; c0 += 16) {
        if (c0 < scols) {
          uint32_t r[16];
          tc::tmem_ld16(tr.taddr + (uint32_t)(s0 + c0), r);
          tc::tmem_ld_wait();
          const int n = tr.n0 + s0 + c0;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float b = (tr.s_bias && n + j < g.N) ? tr.s_bias[n + j] : 0.f;
            v[j] = (__uint_as_float(r[j]) + b) * inv;
            if (p.relu6) v[j] = fminf(fmaxf(v[j], 0.f), 6.f);
          }
          if (esz == 4) {
            float4* d = reinterpret_cast<float4*>(my + (size_t)c0 * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
            uint4* d = reinterpret_cast<uint4*>(my + (size_t)c0 * 2);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint4 q;
              __half2* hq = reinterpret_cast<__half2*>(&q);
#pragma unroll
              for (int j = 0; j < 4; ++j) hq[j] = __floats2half2_rn(v[8 * h + 2 * j], v[8 * h + 2 * j + 1]);
              d[h] = q;
            }
          }
        }
      }
      __syncwarp();
      // write-out: lanes run along the row.  cpr 16-byte chunks per row, padded to a power of two so that row / chunk
      // come from shifts (tail slabs leave a few lanes idle)
      const int cpr = scols * esz >> 4;                        // 1 .. 8
      const int sh = cpr > 4 ? 3 : (cpr > 2 ? 2 : (cpr > 1 ? 1 : 0));
      const int ch = lane & ((1 << sh) - 1);
      const int rstep = 32 >> sh;
      const long long col = (long long)p.col_off + tr.n0 + s0;
      for (int row = lane >> sh; row < 32; row += rstep) {
        const int dst_row = __shfl_sync(0xffffffffu, orow, row);
        if (dst_row < 0 || ch >= cpr) continue;
        const uint4 q = *reinterpret_cast<const uint4*>(tr.stage + (size_t)row * pitch + (size_t)ch * 16);
        if (esz == 2) {
          *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + (long long)dst_row * p.ldo + col + ch * 8) = q;
        } else if (p.f32) {
          *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.out) + (long long)dst_row * p.ldo + col + ch * 4) = q;
        } else {  // fp32 staging + fp16 residual -> fp16 (single rounding)
          const float* f = reinterpret_cast<const float*>(&q);
          const uint2 rr =
              *reinterpret_cast<const uint2*>(p.residual + (long long)dst_row * p.ldr + tr.n0 + s0 + ch * 4);
          const __half2* hr = reinterpret_cast<const __half2*>(&rr);
          const float2 r0 = __half22float2(hr[0]), r1 = __half22float2(hr[1]);
          uint2 o;
          __half2* ho = reinterpret_cast<__half2*>(&o);
          ho[0] = __floats2half2_rn(f[0] + r0.x, f[1] + r0.y);
          ho[1] = __floats2half2_rn(f[2] + r1.x, f[3] + r1.y);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + (long long)dst_row * p.ldo + col + ch * 4) = o;
        }
      }
      __syncwarp();
    }
  }
};

// Detector head tail (hfnet/models/hf_net.py:88-93): + bias, softmax over 65 logits, drop the dustbin channel,
// depth_to_space(8): scores[8h+i][8w+j] = prob[h][w][8i+j].  Optionally keeps the raw logits (parity hook).
struct EpiSoftmaxD2S {
  static constexpr int kWarps = 4;
  struct Params {
    float* scores;   // [B][Hc*8][Wc*8]
    float* logits;   // [row][65] or null
    const float* bias;
    int Hc, Wc;
  };
  static __device__ __forceinline__ const float* bias(const Params&) { return nullptr; }
  static __device__ __forceinline__ void run(const Params& p, const GemmGeom& g, const TileRow& tr) {
    float v[80];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      uint32_t r[16];
      tc::tmem_ld16(tr.taddr + (uint32_t)(16 * c), r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[16 * c + j] = __uint_as_float(r[j]);
    }
    if (!tr.valid) return;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 65; ++j) {
      v[j] += __ldg(p.bias + j);
      mx = fmaxf(mx, v[j]);
    }
    if (p.logits) {
      float* lo = p.logits + tr.row * 65;
#pragma unroll
      for (int j = 0; j < 65; ++j) lo[j] = v[j];
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 65; ++j) {
      v[j] = expf(v[j] - mx);
      sum += v[j];
    }
    const float inv = 1.f / sum;
    const int w = (int)(tr.row % p.Wc);
    const long long t = tr.row / p.Wc;
    const int h = (int)(t % p.Hc);
    const long long b = t / p.Hc;
    const int W8 = p.Wc * 8;
    float* o = p.scores + (b * p.Hc * 8 + (long long)h * 8) * W8 + (long long)w * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      *reinterpret_cast<float4*>(o + (long long)i * W8) =
          make_float4(v[8 * i] * inv, v[8 * i + 1] * inv, v[8 * i + 2] * inv, v[8 * i + 3] * inv);
      *reinterpret_cast<float4*>(o + (long long)i * W8 + 4) =
          make_float4(v[8 * i + 4] * inv, v[8 * i + 5] * inv, v[8 * i + 6] * inv, v[8 * i + 7] * inv);
    }
  }
};

// ------------------------------------------------------------------------------------------------ launchers
template <class Epi>
static int launch_tc(hfb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmGeom& g_in, int B,
                     const typename Epi::Params& ep, const char* what, uint32_t epi_warp_bytes, bool stage_bias) {
  GemmGeom g = g_in;
  gemm_set_rows(g, g.M, B);
  g.epi_warp_bytes = epi_warp_bytes;
  g.bias_bytes = stage_bias ? (uint32_t)((g.N * 4 + 15) & ~15) : 0u;
  const size_t epi_bytes = (size_t)Epi::kWarps * epi_warp_bytes + g.bias_bytes;
  // ring depth: keep the CTA near 110 KB so that two persistent CTAs (8 epilogue warps) share an SM when TMEM allows
  const int halo = g.conv && g.halo;
  const size_t stage = halo ? (size_t)g.BN * 128 : GEMM_TILE_A_BYTES + (size_t)g.BN * 128;
  int st = (int)((110 * 1024 - epi_bytes - (halo ? 2 * GEMM_HALO_A_BYTES : 0)) / stage);
  g.stages = st < 2 ? 2 : (st > 4 ? 4 : st);
  g.ring_bytes = (uint32_t)gemm_ring_bytes(g.BN, g.stages, halo);
  const size_t smem = gemm_smem_bytes(g.BN, g.stages, epi_bytes, halo);
  static SmemOptIn optin;  // per instantiation
  HFB_CUDA(ctx, optin.ensure(gemm_tc_kernel<Epi>, ctx->device, smem));
  if (g.total_tiles <= 0) return HFB_OK;
  const int grid = gemm_grid(g, ctx->n_sm, smem);
  hfb_launch(ctx, gemm_tc_kernel<Epi>, grid, GEMM_THREADS(Epi::kWarps), smem, tmA, tmB, g, ep);
  HFB_CHECK_LAUNCH(ctx, what);
  return HFB_OK;
}

// Descriptor head (hf_net.py:78-80): bias + tf.nn.l2_normalize over the N = BN = 256 columns -> fp32 rows.  Its own
// epilogue type (= its own kernel instantiation: the 128 values a thread keeps in registers must not raise the register
// count of the 3x3 head conv, which runs two CTAs per SM).  The two warps of a TMEM lane group take alternate 32-column
// slabs; each reads its 128 accumulator columns ONCE (eight tcgen05.ld, one wait), adds the bias, sums the squares of
// its half, the halves meet through the 16-byte pad of the staging rows (double-buffered by tile parity, one 64-thread
// named barrier), then the scaled values go out through the staging area with lanes running along the output row.
struct EpiL2Norm {
  static constexpr int kWarps = 8;
  struct Params {
    float* out;
    int ldo;
    const float* bias;
  };
  static __device__ __forceinline__ const float* bias(const Params& p) { return p.bias; }

  static __device__ __forceinline__ void run(const Params& p, const GemmGeom& g, const TileRow& tr) {
    const int lane = threadIdx.x & 31;
    uint8_t* my = tr.stage + (size_t)lane * EPI_PITCH;
    uint32_t v[4][32];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t c = (uint32_t)((tr.sub + 2 * k) * EPI_SLAB);
      uint32_t(&lo)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[k][0]);
      uint32_t(&hi)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[k][16]);
      tc::tmem_ld16(tr.taddr + c, lo);
      tc::tmem_ld16(tr.taddr + c + 16, hi);
    }
    tc::tmem_ld_wait();
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4* b4 = reinterpret_cast<const float4*>(tr.s_bias + (tr.sub + 2 * k) * EPI_SLAB);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b = b4[q];
        const float x0 = __uint_as_float(v[k][4 * q]) + b.x, x1 = __uint_as_float(v[k][4 * q + 1]) + b.y;
        const float x2 = __uint_as_float(v[k][4 * q + 2]) + b.z, x3 = __uint_as_float(v[k][4 * q + 3]) + b.w;
        ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
        v[k][4 * q] = __float_as_uint(x0); v[k][4 * q + 1] = __float_as_uint(x1);
        v[k][4 * q + 2] = __float_as_uint(x2); v[k][4 * q + 3] = __float_as_uint(x3);
      }
    }
    // halves of the row meet: pad word (tile parity) of my staging row, partner = the other warp of this lane group
    const int par = (int)(((tr.taddr & 0xFFFFu) / (uint32_t)g.BN) & 1u);   // accumulator stage 0 / 1 alternates with the tiles
    *reinterpret_cast<float*>(my + EPI_SLAB * 4 + 4 * par) = ss;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + tr.ewarp) : "memory");
    const uint8_t* partner = my + (tr.sub ? -4 : 4) * (long long)g.epi_warp_bytes;
    const float so = *reinterpret_cast<const float*>(partner + EPI_SLAB * 4 + 4 * par);
    const float s0 = tr.sub ? so : ss, s1 = tr.sub ? ss : so;      // same order in both warps
    const float inv = rsqrtf(fmaxf(s0 + s1, 1e-12f));
    // rows this lane writes out: chunk ch of rows (lane >> 3) + 4 i
    const int ch = lane & 7;
    const int orow = tr.valid ? (int)tr.row : -1;
    int drow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) drow[i] = __shfl_sync(0xffffffffu, orow, (lane >> 3) + 4 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4* d = reinterpret_cast<float4*>(my);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        d[q] = make_float4(__uint_as_float(v[k][4 * q]) * inv, __uint_as_float(v[k][4 * q + 1]) * inv,
                           __uint_as_float(v[k][4 * q + 2]) * inv, __uint_as_float(v[k][4 * q + 3]) * inv);
      __syncwarp();
      const int col = tr.n0 + (tr.sub + 2 * k) * EPI_SLAB + ch * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (lane >> 3) + 4 * i;
        if (drow[i] < 0) continue;
        const uint4 q4 = *reinterpret_cast<const uint4*>(tr.stage + (size_t)row * EPI_PITCH + (size_t)ch * 16);
        *reinterpret_cast<uint4*>(p.out + (long long)drow[i] * p.ldo + col) = q4;
      }
      __syncwarp();
    }
  }
};

int gemm_store(hfb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmGeom& g, int B, void* out,
               int ldo, int col_off, const float* bias, const __half* residual, int ldr, int relu6, int f32) {
  EpiStore::Params p{out, ldo, col_off, bias, residual, ldr, relu6, f32, 0};
  // fp16 slabs need 64 B per staged row: the smaller staging area buys the 3x3 head conv a third B stage beside a second CTA
  const uint32_t epi_warp_bytes = (f32 || residual) ? EPI_WARP_BYTES : 32 * (EPI_SLAB * 2 + 16);
  return launch_tc<EpiStore>(ctx, tmA, tmB, g, B, p, "gemm_store", epi_warp_bytes, bias != nullptr);
}
int gemm_l2norm(hfb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmGeom& g, float* out,
                const float* bias) {
  if (g.BN != g.N) {
    ctx->set_error("gemm_l2norm needs BN == N");
    return HFB_ERR_INVALID;
  }
  if (g.N != 8 * EPI_SLAB || !bias) {   // general shapes: the two-pass path of EpiStore
    EpiStore::Params p{out, g.N, 0, bias, nullptr, 0, 0, 1, 1};
    return launch_tc<EpiStore>(ctx, tmA, tmB, g, 1, p, "gemm_l2norm", EPI_WARP_BYTES, true);
  }
  EpiL2Norm::Params p{out, g.N, bias};
  return launch_tc<EpiL2Norm>(ctx, tmA, tmB, g, 1, p, "gemm_l2norm", EPI_WARP_BYTES, true);
}
int gemm_softmax_d2s(hfb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmGeom& g, float* scores,
                     float* logits, const float* bias, int Hc, int Wc) {
  if (g.N != 65 || g.BN != 80) {
    ctx->set_error("gemm_softmax_d2s needs N == 65, BN == 80");
    return HFB_ERR_INVALID;
  }
  EpiSoftmaxD2S::Params p{scores, logits, bias, Hc, Wc};
  return launch_tc<EpiSoftmaxD2S>(ctx, tmA, tmB, g, 1, p, "gemm_softmax_d2s", 0, false);
}

// ------------------------------------------------------------------------------------------------ debug / parity
// CUDA-core restatement of the same contraction, used only by hfb_debug_gemm to cross-check the tensor-core path on
// the device (never on the product path).
__global__ void simt_gemm_kernel(const __half* __restrict__ A, int lda, const __half* __restrict__ Wt, int ldw, int M,
                                 int N, int K, const float* __restrict__ bias, int relu6, float* __restrict__ out,
                                 int conv, int H, int Wd) {
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  const long long m = blockIdx.x;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  if (!conv) {
    for (int k = 0; k < K; ++k) acc = fmaf(__half2float(A[m * lda + k]), __half2float(Wt[(long long)n * ldw + k]), acc);
  } else {
    const int x = (int)(m % Wd);
    const long long t = m / Wd;
    const int y = (int)(t % H);
    const long long b = t / H;
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= Wd) continue;
      const __half* a = A + ((b * H + yy) * Wd + xx) * (long long)K;
      const __half* w = Wt + (long long)n * ldw + tap * K;
      for (int k = 0; k < K; ++k) acc = fmaf(__half2float(a[k]), __half2float(w[k]), acc);
    }
  }
  acc += bias ? bias[n] : 0.f;
  if (relu6) acc = fminf(fmaxf(acc, 0.f), 6.f);
  out[m * N + n] = acc;
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}

extern "C" int hfb_debug_gemm(hfb_ctx* ctx, const float* A, int B, int H, int W, int K, const float* Wt, int N,
                              const float* bias, int relu6, int conv3x3, int use_tc, int BN, float* out) {
  // A: [B*H*W][K] (NHWC when conv3x3), Wt: [N][Kw] with Kw = conv3x3 ? 9*K : K.  out: [B*H*W][N] fp32.
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, K % 8 == 0 && N % 8 == 0, "hfb_debug_gemm: K and N must be multiples of 8");
  const long long M = (long long)B * H * W;
  const int Kw = conv3x3 ? 9 * K : K;
  float *dAf = nullptr, *dWf = nullptr, *dB = nullptr, *dO = nullptr;
  __half *dA = nullptr, *dW = nullptr;
  int rc = HFB_OK;
  auto cleanup = [&]() {
    cudaFree(dAf); cudaFree(dWf); cudaFree(dB); cudaFree(dO); cudaFree(dA); cudaFree(dW);
  };
#define DBG_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ctx->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));           \
      cleanup();                                                                    \
      return HFB_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)
  DBG_CUDA(cudaMalloc(&dAf, M * K * 4));
  DBG_CUDA(cudaMalloc(&dWf, (size_t)N * Kw * 4));
  DBG_CUDA(cudaMalloc(&dB, (size_t)N * 4));
  DBG_CUDA(cudaMalloc(&dO, M * N * 4));
  DBG_CUDA(cudaMalloc(&dA, M * K * 2));
  DBG_CUDA(cudaMalloc(&dW, (size_t)N * Kw * 2));
  DBG_CUDA(cudaMemcpyAsync(dAf, A, M * K * 4, cudaMemcpyHostToDevice, ctx->stream));
  DBG_CUDA(cudaMemcpyAsync(dWf, Wt, (size_t)N * Kw * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (bias) DBG_CUDA(cudaMemcpyAsync(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
  f32_to_f16_kernel<<<(unsigned)((M * K + 255) / 256), 256, 0, ctx->stream>>>(dAf, dA, M * K);
  f32_to_f16_kernel<<<(unsigned)(((long long)N * Kw + 255) / 256), 256, 0, ctx->stream>>>(dWf, dW, (long long)N * Kw);
  DBG_CUDA(cudaMemsetAsync(dO, 0xFF, M * N * 4, ctx->stream));
  if (use_tc) {
    CUtensorMap tmA, tmB;
    GemmGeom g;
    if (BN <= 0) BN = ((N + 15) / 16) * 16 > 256 ? 128 : ((N + 15) / 16) * 16;
    if (conv3x3) {
      rc = hfb_make_tmap_nhwc(ctx, &tmA, dA, K, W, H, B);
      gemm_fill_geom_conv(g, B, H, W, K, N, BN);
    } else {
      rc = hfb_make_tmap_2d(ctx, &tmA, dA, K, M, (uint64_t)K * 2, 128);
      gemm_fill_geom(g, (int)M, N, K, BN, 0);
    }
    if (rc == HFB_OK) rc = hfb_make_tmap_2d(ctx, &tmB, dW, Kw, N, (uint64_t)Kw * 2, BN);
    if (rc == HFB_OK) rc = gemm_store(ctx, tmA, tmB, g, B, dO, N, 0, bias ? dB : nullptr, nullptr, 0, relu6, 1);
  } else {
    dim3 grid((unsigned)M, (N + 127) / 128);
    simt_gemm_kernel<<<grid, 128, 0, ctx->stream>>>(dA, K, dW, Kw, (int)M, N, K, bias ? dB : nullptr, relu6, dO,
                                                    conv3x3, H, W);
  }
  if (rc == HFB_OK) {
    cudaError_t e = cudaMemcpyAsync(out, dO, M * N * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->set_error(std::string("hfb_debug_gemm: ") + cudaGetErrorString(e));
      rc = HFB_ERR_CUDA;
    }
  }
  cleanup();
  return rc;
#undef DBG_CUDA
}
