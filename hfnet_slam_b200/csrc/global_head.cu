// Global head of HF-Net: NetVLAD (hfnet/models/layers.py:57-95) + dimensionality reduction FC (layers.py:97-109) on
// the layer_18 endpoint, as TWO kernels:
//   vlad_kernel   grid (pixel groups, frames): soft-assignment (1x1 conv D -> C, softmax over clusters) and the
//                 per-cluster aggregation of a group of 64 pixels out of shared memory; the partial sums of a frame's
//                 groups are combined in fixed order by the group that finishes last, which also runs the two
//                 normalisations and emits the FC operand (fp32 -> fp16 high + low parts, already in the register layout
//                 of the warp MMA A fragment).
//   fc_mma_kernel grid (4096 / 32 output-column tiles): the K x 4096 fp16 weight is stored at load time in the order the
//                 B fragments of mma.m16n8k16 are consumed, so every warp streams its K range with contiguous 512-byte
//                 loads straight into registers (no shared-memory staging: the layer is a skinny, HBM-bound product --
//                 at most 16 frames against 63 MB of weights -- and what matters is bytes in flight).  fp32 accumulate,
//                 activations enter as fp16 high + low (fp32-accurate), fixed-order cross-warp sum, bias, and the final
//                 tf.nn.l2_normalize by the CTA that finishes last.
// Warp-level mma.sync rather than tcgen05 on purpose: M is the number of frames (<= 16 per pass) and the operand that
// matters is streamed exactly once, so there is nothing for a shared-memory operand pipeline + TMEM round trip to reuse.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int VP = 64;          // pixels per group
constexpr int VLAD_THREADS = 256;
constexpr int FC_THREADS = 512; // 16 warps split K
constexpr int FC_NT = 32;       // output columns per CTA

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Both NetVLAD products of a 64-pixel group run on warp MMA tiles (fp32 accumulate): the soft-assignment logits
// X[64 x D] . W[D x C] with the fp32 weights as an fp16 high + low pair (wt: [2][C][D + 8], built at load time), and the
// aggregation M^T[C x 64] . X[64 x D] with the fp32 memberships split the same way; X is exact in fp16.
// part: [B][n_groups][C * D + C] (sum_p m x | sum_p m); counters[b] self-resetting.
// a_packed: [ceil(B / 8)][K / 16][32 lanes][4 words] = {hi a0, hi a2, lo a0, lo a2} of mma.m16n8k16's A fragment for the
// 8 frames of a group (row = frame % 8): word a0 holds k = 2 * (lane % 4) + {0, 1}, a2 the same + 8.
__global__ void __launch_bounds__(VLAD_THREADS) vlad_kernel(const __half* __restrict__ x, int P, int D, int C,
                                                            const __half* __restrict__ wt, const float* __restrict__ bias,
                                                            const float* __restrict__ clusters, float* __restrict__ part,
                                                            int* __restrict__ counters, float* __restrict__ vladn,
                                                            uint32_t* __restrict__ a_packed) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int DP = D + 8;                        // halves per row: an odd number of 16-byte units -> conflict-free ldmatrix
  constexpr int MP = VP + 8;
  __half* s_xh = reinterpret_cast<__half*>(s_raw);                 // [VP][DP]
  __half* s_wt = s_xh + (size_t)VP * DP;                           // [2][C][DP]
  float* s_m = reinterpret_cast<float*>(s_wt + (size_t)2 * C * DP);   // [VP][C]
  __half* s_mT = reinterpret_cast<__half*>(s_m + (size_t)VP * C);  // [2][C][MP]
  float* s_ms = reinterpret_cast<float*>(s_mT + (size_t)2 * C * MP);  // [C]
  __shared__ float red[32];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, grp = blockIdx.x, n_groups = gridDim.x;
  const int K = C * D;
  pdl_launch_dependents();
  {   // weights do not depend on the predecessor; 16-byte copies, all of a thread's loads in flight together
    const uint4* w4 = reinterpret_cast<const uint4*>(wt);
    const int n4 = 2 * C * DP / 8;
#pragma unroll 8
    for (int i = tid; i < n4; i += VLAD_THREADS) reinterpret_cast<uint4*>(s_wt)[i] = __ldg(w4 + i);
  }
  pdl_wait();
  const int pix0 = grp * VP;
  const int npix = min(VP, P - pix0);
  {
    const uint4* src = reinterpret_cast<const uint4*>(x + ((size_t)b * P + pix0) * D);
    const int upr = D / 8;   // 16-byte units per pixel row
#pragma unroll 8
    for (int i = tid; i < VP * upr; i += VLAD_THREADS) {
      const int p = i / upr, u = i - p * upr;
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (p < npix) q = __ldg(src + i);
      *reinterpret_cast<uint4*>(s_xh + (size_t)p * DP + u * 8) = q;
    }
  }
  __syncthreads();
  const int n_ct = C / 8;   // n8 tiles of the cluster axis (<= 8)
  // ---- soft assignment: warps 0..3 own 16 pixels each, all clusters
  if (warp < VP / 16) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const uint32_t xa = (uint32_t)__cvta_generic_to_shared(s_xh + (size_t)(warp * 16 + (lane & 15)) * DP + (lane >> 4) * 8);
    const uint32_t wa = (uint32_t)__cvta_generic_to_shared(s_wt + (size_t)((lane & 7) + (lane >> 4) * 8) * DP + ((lane >> 3) & 1) * 8);
    const uint32_t lo_off = (uint32_t)((size_t)C * DP * 2);
    for (int k0 = 0; k0 < D; k0 += 16) {
      uint32_t a0, a1, a2, a3;
      ldsm_x4(xa + k0 * 2, a0, a1, a2, a3);
#pragma unroll
      for (int np = 0; np < 4; ++np) {   // pairs of n8 tiles
        if (2 * np < n_ct) {
          uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
          const uint32_t wb = wa + (uint32_t)(np * 16 * DP + k0) * 2u;
          ldsm_x4(wb, h0, h1, h2, h3);
          ldsm_x4(wb + lo_off, l0, l1, l2, l3);
          mma16816(acc[2 * np], a0, a1, a2, a3, h0, h1);
          mma16816(acc[2 * np + 1], a0, a1, a2, a3, h2, h3);
          mma16816(acc[2 * np], a0, a1, a2, a3, l0, l1);
          mma16816(acc[2 * np + 1], a0, a1, a2, a3, l2, l3);
        }
      }
    }
    // softmax over the clusters of a pixel row: a row lives in the four lanes of a quad
    const int cq = 2 * (lane & 3);
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int p = warp * 16 + (lane >> 2) + 8 * hrow;
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < n_ct) {
          acc[nt][2 * hrow] += __ldg(bias + nt * 8 + cq);
          acc[nt][2 * hrow + 1] += __ldg(bias + nt * 8 + cq + 1);
          mx = fmaxf(mx, fmaxf(acc[nt][2 * hrow], acc[nt][2 * hrow + 1]));
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < n_ct) {
          acc[nt][2 * hrow] = expf(acc[nt][2 * hrow] - mx);
          acc[nt][2 * hrow + 1] = expf(acc[nt][2 * hrow + 1] - mx);
          sum += acc[nt][2 * hrow] + acc[nt][2 * hrow + 1];
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = p < npix ? 1.f / sum : 0.f;   // padded pixels carry no weight
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < n_ct) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = nt * 8 + cq + j;
            const float m = acc[nt][2 * hrow + j] * inv;
            s_m[(size_t)p * C + c] = m;
            const __half hi = __float2half_rn(m);
            s_mT[(size_t)c * MP + p] = hi;
            s_mT[(size_t)(C + c) * MP + p] = __float2half_rn(m - __half2float(hi));
          }
        }
      }
    }
  }
  __syncthreads();
  float* pp = part + ((size_t)b * n_groups + grp) * (size_t)(K + C);
  if (tid < C) {
    float s = 0.f;
    for (int p = 0; p < VP; ++p) s += s_m[(size_t)p * C + tid];
    pp[K + tid] = s;
  }
  // ---- aggregation over this group's pixels: warp = 32 channels (four n8 tiles), 16 clusters per pass
  for (int n0 = warp * 32; n0 < D; n0 += (VLAD_THREADS / 32) * 32) {
    const uint32_t ma = (uint32_t)__cvta_generic_to_shared(s_mT + (size_t)(lane & 15) * MP + (lane >> 4) * 8);
    const uint32_t xb = (uint32_t)__cvta_generic_to_shared(s_xh + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * DP + n0 + (lane >> 4) * 8);
    const uint32_t mlo = (uint32_t)((size_t)C * MP * 2);
    for (int mt = 0; mt < C / 16; ++mt) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
      for (int k0 = 0; k0 < VP; k0 += 16) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        ldsm_x4(ma + (uint32_t)(mt * 16 * MP + k0) * 2u, h0, h1, h2, h3);
        ldsm_x4(ma + (uint32_t)(mt * 16 * MP + k0) * 2u + mlo, l0, l1, l2, l3);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          if (n0 + np * 16 < D) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(xb + (uint32_t)(k0 * DP + np * 16) * 2u, b0, b1, b2, b3);
            mma16816(acc[2 * np], h0, h1, h2, h3, b0, b1);
            mma16816(acc[2 * np + 1], h0, h1, h2, h3, b2, b3);
            mma16816(acc[2 * np], l0, l1, l2, l3, b0, b1);
            mma16816(acc[2 * np + 1], l0, l1, l2, l3, b2, b3);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int d = n0 + nt * 8 + 2 * (lane & 3);
        if (d < D) {
          const int c = mt * 16 + (lane >> 2);
          *reinterpret_cast<float2*>(pp + (size_t)c * D + d) = make_float2(acc[nt][0], acc[nt][1]);
          *reinterpret_cast<float2*>(pp + (size_t)(c + 8) * D + d) = make_float2(acc[nt][2], acc[nt][3]);
        }
      }
    }
  }
  // ---- the group that finishes last combines the frame's partial sums (fixed order) and normalises
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int old = atomicAdd(&counters[b], 1);
    s_last = old == n_groups - 1;
    if (s_last) counters[b] = 0;   // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* pb = part + (size_t)b * n_groups * (size_t)(K + C);
  if (tid < C) {
    float s = 0.f;
    for (int g = 0; g < n_groups; ++g) s += __ldcg(pb + (size_t)g * (K + C) + K + tid);
    s_ms[tid] = s;
  }
  __syncthreads();
  float* s_V = reinterpret_cast<float*>(s_raw);   // [C][D], over the pixel / weight staging area
  for (int i0 = tid * 4; i0 < K; i0 += VLAD_THREADS * 4) {   // K + C and D are multiples of 4: rows stay 16-byte aligned
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 6
    for (int g = 0; g < n_groups; ++g) {
      const float4 q = __ldcg(reinterpret_cast<const float4*>(pb + (size_t)g * (K + C) + i0));
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    // layers.py:81-86: sum_p m (centroid - x)
    const float4 cl = __ldg(reinterpret_cast<const float4*>(clusters + i0));
    const float ms = s_ms[i0 / D];
    *reinterpret_cast<float4*>(s_V + i0) = make_float4(ms * cl.x - a.x, ms * cl.y - a.y, ms * cl.z - a.z, ms * cl.w - a.w);
  }
  __syncthreads();
  // intra-normalisation over the CLUSTER axis (layers.py:87-88, restated literally), flatten, L2 (layers.py:89-90) and
  // the L2 at the top of the dimensionality reduction (layers.py:97)
  float ss = 0.f;
  for (int d = tid; d < D; d += VLAD_THREADS) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(s_V[(size_t)c * D + d], s_V[(size_t)c * D + d], s);
    const float inv = rsqrtf(fmaxf(s, 1e-12f));
    for (int c = 0; c < C; ++c) {
      const float t = s_V[(size_t)c * D + d] * inv;
      s_V[(size_t)c * D + d] = t;
      ss = fmaf(t, t, ss);
    }
  }
  float tot = block_sum_256(ss, red);
  const float inv1 = rsqrtf(fmaxf(tot, 1e-12f));
  float ss2 = 0.f;
  for (int i = tid; i < K; i += VLAD_THREADS) {
    const float t = s_V[i] * inv1;
    s_V[i] = t;
    ss2 = fmaf(t, t, ss2);
  }
  tot = block_sum_256(ss2, red);
  const float inv2 = rsqrtf(fmaxf(tot, 1e-12f));
  float* vo = vladn + (size_t)b * K;
  __half* ap = reinterpret_cast<__half*>(a_packed) + (size_t)(b >> 3) * (K / 16) * 32 * 8;
  const int row = b & 7;
  for (int i = tid; i < K; i += VLAD_THREADS) {
    const float t = s_V[i] * inv2;
    vo[i] = t;
    const __half hi = __float2half_rn(t);
    const __half lo = __float2half_rn(t - __half2float(hi));
    const int k16 = i >> 4, kk = i & 15;
    const int ln = row * 4 + ((kk & 7) >> 1);
    const size_t o = ((size_t)k16 * 32 + ln) * 8 + (size_t)(kk >> 3) * 2 + (kk & 1);
    ap[o] = hi;
    ap[o + 4] = lo;
  }
}

// The weight stream goes through a per-warp cp.async ring in shared memory: with plain loads the compiler sinks every load
// next to its MMAs (one K step = 1 KB in flight per warp), which caps the stream at a third of the HBM rate; with the ring
// a warp keeps FC_STAGES - 1 KB in flight and nothing in the register scoreboard.  Every lane reads back exactly the
// 16-byte slots it copied itself, so cp.async.wait_group is the only synchronisation.
constexpr int FC_STAGES = 5;
constexpr int FC_STAGE_U4 = 128;   // uint4 per stage and warp: 64 weights | 32 A (frames 0..7) | 32 A (frames 8..15)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
// The product is issued TRANSPOSED: the weight is the MMA's A operand (16 output columns x 16 K per tile), the frames are
// the 8 columns of the B operand, so a batch of <= 8 frames wastes no MMA rows -- legacy HMMA runs at a fraction of the
// tcgen05 rate on this part and the first (frames-as-rows) version was bound by it, not by the HBM stream.
// wp: [K / 16][4096 / 32 tiles][2][32 lanes][4 words]: A fragment (a0..a3) of the 16-column tile h (fc_pack_host).
// K steps are dealt round-robin to the warps and the tiles of a K step are adjacent in memory, so at any
// moment the whole grid reads one contiguous window of the weight (DRAM pages are read out while they are open, like a
// streaming copy) instead of 2048 separate streams.  out: [B][4096]; ss_part: [B][gridDim.x]; counter self-resetting.
__global__ void __launch_bounds__(FC_THREADS) fc_mma_kernel(const uint4* __restrict__ wp, const uint4* __restrict__ a_packed,
                                                            int KT, int B, const float* __restrict__ bias, int N,
                                                            float* __restrict__ out, float* __restrict__ ss_part,
                                                            int* __restrict__ counter, unsigned long long* __restrict__ dbg) {
  extern __shared__ uint4 fc_ring[];               // [warp][FC_STAGES][FC_STAGE_U4]
  __shared__ float red[FC_THREADS / 32][32][17];   // [warp][lane][acc], padded
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntile = blockIdx.x, n_tiles = gridDim.x;
  constexpr int NW = FC_THREADS / 32;
  const int k_begin = warp, k_end = KT;   // K steps warp, warp + NW, ...
  auto stamp = [&](int i) {   // HFB_FC_DBG=1: globaltimer stamps of thread 0 of every CTA
    if (dbg && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[blockIdx.x * 8 + i] = t;
    }
  };
  stamp(0);
  pdl_launch_dependents();
  pdl_wait();
  stamp(1);
  const uint4* wt = wp + (size_t)ntile * 64 + lane;
  const size_t wstep = (size_t)n_tiles * 64;   // uint4 per K step
  const int n_groups = (B + 7) >> 3;
  for (int g0 = 0; g0 < n_groups; g0 += 2) {   // 16 frames per pass: two groups of 8 = two B operands per weight tile
    const bool two = g0 + 1 < n_groups;
    const uint4* a0p = a_packed + (size_t)g0 * KT * 32 + lane;
    const uint4* a1p = a_packed + (size_t)(g0 + 1) * KT * 32 + lane;
    float acc[4][4];   // [frame group][column tile]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(fc_ring) + (uint32_t)(warp * FC_STAGES * FC_STAGE_U4 + lane) * 16u;
    auto issue = [&](int k, int slot) {
      if (k < k_end) {
        const uint32_t dst = ring + (uint32_t)(slot * FC_STAGE_U4) * 16u;
        cp_async16(dst, wt + (size_t)k * wstep);
        cp_async16(dst + 32 * 16, wt + (size_t)k * wstep + 32);
        cp_async16(dst + 64 * 16, a0p + (size_t)k * 32);
        if (two) cp_async16(dst + 96 * 16, a1p + (size_t)k * 32);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty) group: the wait count stays uniform
    };
#pragma unroll
    for (int s0 = 0; s0 < FC_STAGES - 1; ++s0) issue(k_begin + s0 * NW, s0);
    int slot = 0;
    for (int k = k_begin; k < k_end; k += NW) {
      asm volatile("cp.async.wait_group %0;" ::"n"(FC_STAGES - 2) : "memory");
      const uint32_t src = ring + (uint32_t)(slot * FC_STAGE_U4) * 16u;
      const uint4 w0 = lds_v4(src), w1 = lds_v4(src + 32 * 16), a = lds_v4(src + 64 * 16);
      const uint4 c = two ? lds_v4(src + 96 * 16) : make_uint4(0u, 0u, 0u, 0u);
      issue(k + (FC_STAGES - 1) * NW, slot == 0 ? FC_STAGES - 1 : slot - 1);   // refills the slot consumed one step ago
      // weights = A (rows = output columns), frames = B: hi + lo parts of the activations
      mma16816(acc[0], w0.x, w0.y, w0.z, w0.w, a.x, a.y);
      mma16816(acc[1], w1.x, w1.y, w1.z, w1.w, a.x, a.y);
      mma16816(acc[0], w0.x, w0.y, w0.z, w0.w, a.z, a.w);
      mma16816(acc[1], w1.x, w1.y, w1.z, w1.w, a.z, a.w);
      if (two) {
        mma16816(acc[2], w0.x, w0.y, w0.z, w0.w, c.x, c.y);
        mma16816(acc[3], w1.x, w1.y, w1.z, w1.w, c.x, c.y);
        mma16816(acc[2], w0.x, w0.y, w0.z, w0.w, c.z, c.w);
        mma16816(acc[3], w1.x, w1.y, w1.z, w1.w, c.z, c.w);
      }
      slot = slot == FC_STAGES - 1 ? 0 : slot + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    stamp(2);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][lane][i * 4 + j] = acc[i][j];
    __syncthreads();
    // thread = (MMA row, column of the tile): fixed-order sum over the K ranges of the warps
    {
      // accumulator (group, tile) element c[j]: column = tile * 16 + lane / 4 + 8 * (j / 2), frame = 2 * (lane % 4) + j % 2
      const int row = tid >> 5, col = tid & 31;
      const int f = row & 7, r = col & 15;
      const int ln = (r & 7) * 4 + (f >> 1), ai = ((row >> 3) * 2 + (col >> 4)) * 4 + (r >> 3) * 2 + (f & 1);
      float y = 0.f;
#pragma unroll
      for (int wq = 0; wq < NW; ++wq) y += red[wq][ln][ai];
      const int frame = g0 * 8 + row;
      const int n = ntile * FC_NT + col;
      y += __ldg(bias + n);
      const bool live = frame < B;
      if (live) out[(size_t)frame * N + n] = y;
      float sq = live ? y * y : 0.f;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, s);
      if (live && col == 0) ss_part[(size_t)frame * n_tiles + ntile] = sq;
    }
  }
  // ---- final tf.nn.l2_normalize (layers.py:108) by the CTA that finishes last
  stamp(3);
  __threadfence();
  __syncthreads();
  stamp(4);
  if (tid == 0) {
    const int old = atomicAdd(counter, 1);
    s_last = old == n_tiles - 1;
    if (s_last) *counter = 0;
  }
  __syncthreads();
  stamp(5);
  if (!s_last) return;
  __threadfence();
  float* s_inv = &red[0][0][0];
  for (int f = warp; f < B; f += NW) {
    float t = 0.f;
    for (int i = lane; i < n_tiles; i += 32) t += __ldcg(ss_part + (size_t)f * n_tiles + i);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
    if (lane == 0) s_inv[f] = rsqrtf(fmaxf(t, 1e-12f));
  }
  __syncthreads();
  {   // out *= inv[frame]: loads of a batch of 8 vectors issued together (a plain loop serialises on the possible alias)
    const int n4 = N / 4, total = B * n4;
    float4* o = reinterpret_cast<float4*>(out);
    for (int i0 = tid; i0 < total; i0 += FC_THREADS * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * FC_THREADS;
        if (i < total) v[u] = __ldcg(o + i);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * FC_THREADS;
        if (i < total) {
          const float inv = s_inv[i / n4];
          o[i] = make_float4(v[u].x * inv, v[u].y * inv, v[u].z * inv, v[u].w * inv);
        }
      }
    }
  }
  __syncthreads();
  stamp(6);
}

}  // namespace

// fw: [K][N] fp32 (blob layout) -> fp16 in the order fc_mma_kernel streams it.
void fc_pack_host(const float* fw, int K, int N, __half* out) {
  for (int k = 0; k < K; ++k) {
    const int k16 = k >> 4, kk = k & 15;
    const float* row = fw + (size_t)k * N;
    for (int n = 0; n < N; ++n) {
      const int ntile = n >> 5, mt = (n & 31) >> 4, r = n & 15;
      const int ln = (r & 7) * 4 + ((kk & 7) >> 1);
      const size_t o = ((((size_t)k16 * (N / 32) + ntile) * 2 + mt) * 32 + ln) * 8 + (size_t)((r >> 3) + 2 * (kk >> 3)) * 2 +
                       (kk & 1);
      out[o] = __float2half_rn(row[n]);
    }
  }
}

int global_head_groups(int P) { return (P + VP - 1) / VP; }

int global_head_run(hfb_ctx* ctx, const __half* x, int P, int D, int B, float* d_part, int* d_counters, float* d_vladn,
                    uint32_t* d_apacked, float* d_ss_part, float* d_out) {
  const NetW& net = ctx->net;
  const int C = net.n_clusters, K = C * D;
  HFB_REQUIRE(ctx, D % 16 == 0 && C % 16 == 0 && C <= 64, "global head: needs channels % 16 == 0 and clusters in {16, 32, 48, 64}");
  const size_t DP = (size_t)D + 8, MP = VP + 8;
  const size_t smem = (VP + 2 * (size_t)C) * DP * 2 + (size_t)VP * C * 4 + 2 * (size_t)C * MP * 2 + (size_t)C * 4 + 64;
  HFB_REQUIRE(ctx, smem <= 200 * 1024, "global head: endpoint too wide for the NetVLAD kernel");
  static SmemOptIn optin;
  if (smem > 48 * 1024) HFB_CUDA(ctx, optin.ensure(vlad_kernel, ctx->device, smem));
  const int groups = global_head_groups(P);
  ctx->note("global.netvlad", 2.0 * B * P * D, 4.0 * B * P * D * C);
  hfb_launch(ctx, vlad_kernel, dim3(groups, B), VLAD_THREADS, smem, x, P, D, C, net.vlad_wt, net.vlad_b, net.vlad_c, d_part,
             d_counters, d_vladn, d_apacked);
  HFB_CHECK_LAUNCH(ctx, "vlad");
  ctx->note("global.fc", 2.0 * K * HFB_GLOBAL_DIM + 4.0 * B * K, 2.0 * B * K * HFB_GLOBAL_DIM);
  const size_t fc_smem = (size_t)(FC_THREADS / 32) * FC_STAGES * FC_STAGE_U4 * 16;
  static unsigned long long* d_dbg = nullptr;
  static int dbg_runs = 0;
  if (getenv("HFB_FC_DBG") && !d_dbg) {
    cudaMalloc(&d_dbg, 128 * 8 * 8);
    cudaMemset(d_dbg, 0, 128 * 8 * 8);
  }
  static SmemOptIn fc_optin;
  HFB_CUDA(ctx, fc_optin.ensure(fc_mma_kernel, ctx->device, fc_smem));
  hfb_launch(ctx, fc_mma_kernel, HFB_GLOBAL_DIM / FC_NT, FC_THREADS, fc_smem, reinterpret_cast<const uint4*>(net.fc_w),
             reinterpret_cast<const uint4*>(d_apacked), K / 16, B, net.fc_b, HFB_GLOBAL_DIM, d_out, d_ss_part,
             d_counters + ctx->cfg.max_batch, d_dbg);
  HFB_CHECK_LAUNCH(ctx, "fc_mma");
  if (d_dbg && dbg_runs++ == 3) {
    cudaStreamSynchronize(ctx->stream);
    std::vector<unsigned long long> h(128 * 8);
    cudaMemcpy(h.data(), d_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < 128; ++c) t0 = std::min(t0, h[c * 8]);
    for (int i = 0; i < 7; ++i) {
      unsigned long long lo = ~0ull, hi = 0;
      for (int c = 0; c < 128; ++c)
        if (h[c * 8 + i]) { lo = std::min(lo, h[c * 8 + i]); hi = std::max(hi, h[c * 8 + i]); }
      fprintf(stderr, "hfnet_b200: fc stamp %d: min %lld max %lld ns\n", i, (long long)(lo - t0), (long long)(hi - t0));
    }
  }
  return HFB_OK;
}
