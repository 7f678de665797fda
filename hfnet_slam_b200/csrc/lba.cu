// Local bundle adjustment numeric core on the GPU (fp64), replacing g2o's per-edge linearisation, Hessian assembly,
// Schur complement and landmark back-substitution inside Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1116-1498):
//
//   residual / depth      include/OptimizableTypes.h:99-110, src/CameraModels/Pinhole.cpp:35-41
//   Jacobians             src/OptimizableTypes.cpp:139-159, src/CameraModels/Pinhole.cpp:71-81
//   Huber + quadratic form Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91, base_edge.h:96-102,
//                          base_binary_edge.hpp:55-121
//   Schur / back-subst.   Thirdparty/g2o/g2o/core/block_solver.hpp:354-486, setLambda :564-589
//   Levenberg control     Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-185 (host loop below)
//
// Layout: edges sorted by point (SoA), so one thread owns one landmark and walks its contiguous edges: Hll / bl are
// accumulated in registers in edge order; camera blocks Hpp / bp are summed per camera in a fixed edge order from a
// camera-sorted edge list; the Schur complement is accumulated per CTA into private upper-triangular block buffers
// (plain read-modify-write, one point at a time, entries of one point never collide) and reduced over CTAs in a fixed
// order.  No floating-point atomics anywhere: results are bit-reproducible run to run.  Only the reduced 6n x 6n
// camera system is factorised on the host (dense Cholesky; the reference uses sparse LDLT on the same system).
#include <math.h>

#include <algorithm>

#include <atomic>

#include "common.cuh"

struct LbaDev {
  int n_cams = 0, n_points = 0, n_edges = 0, n_opt = 0, n_blk = 0, G = 0;
  double *poses = nullptr, *poses_t = nullptr, *points = nullptr, *points_t = nullptr;
  int *edge_cam = nullptr, *pt_ptr = nullptr, *cam_ptr = nullptr, *cam_edges = nullptr, *cam_slot = nullptr,
      *opt_cams = nullptr, *edge_slot = nullptr;   // edge_slot[e] = cam_slot[edge_cam[e]] (one load instead of a dependent pair)
  double *obs = nullptr, *invs2 = nullptr;
  double *Jc = nullptr, *wo = nullptr, *r = nullptr, *Hpl = nullptr, *chi2_a = nullptr, *chi2_b = nullptr;
  double *Hll = nullptr, *bl = nullptr, *Dinv = nullptr, *rho_pt = nullptr, *scale_pt = nullptr;
  double *Hpp = nullptr, *bp = nullptr, *partial = nullptr, *Hs = nullptr, *bs = nullptr, *xp = nullptr;
  double* scal = nullptr;  // [4]: chi, max diag, trial chi, scale_l
  unsigned char* depth_ok = nullptr;
  double K[4];
  double delta;
};

__device__ __forceinline__ void quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ double lba_shfl_xor(double v, int o) {
  return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), o),
                          __shfl_xor_sync(0xffffffffu, __double2loint(v), o));
}

// LBA_PL lanes per landmark: errors, robust weights, Jacobians, per-edge camera terms; lane j of a landmark takes its edges
// j, j + LBA_PL, ... and the landmark's sums (Hll, bl, rho) are joined by a fixed xor butterfly (bit-reproducible).
// mode 0 = full linearisation; mode 1 = error evaluation only (chi2, rho).
#define LBA_PL 8
// tail != 0 (LM trial, mode 1): the block that finishes last also sums the trial's robust chi2 (rho_pt) and the landmark
// part of the gain-ratio denominator (scale_pt, from the back-substitution before) in a fixed order into scal[2] / scal[3]
// and posts scal[0..5] + the sequence number into page-locked host memory -- the trial needs no reduction launch.
__global__ void lba_linearize_kernel(LbaDev d, const double* __restrict__ poses, const double* __restrict__ points,
                                     double* __restrict__ chi2_out, int mode, int tail, volatile double* post, double seq) {
  pdl_launch_dependents();
  pdl_wait();
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = gt / LBA_PL, sub = gt % LBA_PL;
  const bool live = p < d.n_points;            // whole groups of LBA_PL lanes are live or not; dead ones still shuffle
  const int pp = live ? p : 0;
  const double X0 = points[3 * pp], X1 = points[3 * pp + 1], X2 = points[3 * pp + 2];
  const double fx = d.K[0], fy = d.K[1], cx = d.K[2], cy = d.K[3];
  const double dsqr = d.delta * d.delta;
  double H0 = 0, H1 = 0, H2 = 0, H3 = 0, H4 = 0, H5 = 0, b0 = 0, b1 = 0, b2 = 0, rho_sum = 0;
  const int e_end = live ? d.pt_ptr[pp + 1] : 0;
  for (int e = (live ? d.pt_ptr[pp] : 0) + sub; e < e_end; e += LBA_PL) {
    const int c = d.edge_cam[e];
    const double* ps = poses + 7 * c;
    double R[9];
    quat_to_R(ps, R);
    const double x = R[0] * X0 + R[1] * X1 + R[2] * X2 + ps[4];
    const double y = R[3] * X0 + R[4] * X1 + R[5] * X2 + ps[5];
    const double z = R[6] * X0 + R[7] * X1 + R[8] * X2 + ps[6];
    const double e0 = d.obs[2 * e] - (fx * x / z + cx);
    const double e1 = d.obs[2 * e + 1] - (fy * y / z + cy);
    const double is2 = d.invs2[e];
    const double chi2 = is2 * (e0 * e0 + e1 * e1);
    chi2_out[e] = chi2;
    const bool inl = chi2 <= dsqr;
    const double sq = sqrt(fmax(chi2, 1e-300));
    rho_sum += inl ? chi2 : 2 * sq * d.delta - dsqr;
    if (mode == 1) {
      if (d.depth_ok) d.depth_ok[e] = z > 0.0 ? 1 : 0;
      continue;
    }
    const double w = inl ? 1.0 : d.delta / sq;
    const double wo = w * is2;
    // -Jproj (2x3): [[-fx/z, 0, fx x/z^2], [0, -fy/z, fy y/z^2]]
    const double iz = 1.0 / z, iz2 = 1.0 / (z * z);
    const double a00 = -fx * iz, a02 = fx * x * iz2, a11 = -fy * iz, a12 = fy * y * iz2;
    // Jp = (-Jproj) * R
    double Jp[6];
    Jp[0] = a00 * R[0] + a02 * R[6]; Jp[1] = a00 * R[1] + a02 * R[7]; Jp[2] = a00 * R[2] + a02 * R[8];
    Jp[3] = a11 * R[3] + a12 * R[6]; Jp[4] = a11 * R[4] + a12 * R[7]; Jp[5] = a11 * R[5] + a12 * R[8];
    // Jc = (-Jproj) * [ -[X]x | I ],  -[X]x = [[0, z, -y], [-z, 0, x], [y, -x, 0]]
    double Jc[12];
    Jc[0] = a02 * y;            Jc[1] = a00 * z - a02 * x;  Jc[2] = -a00 * y;  Jc[3] = a00; Jc[4] = 0;   Jc[5] = a02;
    Jc[6] = -a11 * z + a12 * y; Jc[7] = -a12 * x;           Jc[8] = a11 * x;   Jc[9] = 0;   Jc[10] = a11; Jc[11] = a12;
    const double r0 = -wo * e0, r1 = -wo * e1;  // -(rho' Omega e)
    H0 += wo * (Jp[0] * Jp[0] + Jp[3] * Jp[3]);
    H1 += wo * (Jp[0] * Jp[1] + Jp[3] * Jp[4]);
    H2 += wo * (Jp[0] * Jp[2] + Jp[3] * Jp[5]);
    H3 += wo * (Jp[1] * Jp[1] + Jp[4] * Jp[4]);
    H4 += wo * (Jp[1] * Jp[2] + Jp[4] * Jp[5]);
    H5 += wo * (Jp[2] * Jp[2] + Jp[5] * Jp[5]);
    b0 += Jp[0] * r0 + Jp[3] * r1;
    b1 += Jp[1] * r0 + Jp[4] * r1;
    b2 += Jp[2] * r0 + Jp[5] * r1;
    if (d.cam_slot[c] >= 0) {
      double* jc = d.Jc + 12 * (size_t)e;
#pragma unroll
      for (int i = 0; i < 12; ++i) jc[i] = Jc[i];
      d.wo[e] = wo;
      d.r[2 * e] = r0;
      d.r[2 * e + 1] = r1;
      double* hpl = d.Hpl + 18 * (size_t)e;  // 6x3 = wo * Jc^T Jp
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) hpl[3 * i + j] = wo * (Jc[i] * Jp[j] + Jc[6 + i] * Jp[3 + j]);
    }
  }
#pragma unroll
  for (int o = LBA_PL / 2; o > 0; o >>= 1) {
    rho_sum += lba_shfl_xor(rho_sum, o);
    if (mode == 0) {                             // warp-uniform
      H0 += lba_shfl_xor(H0, o); H1 += lba_shfl_xor(H1, o); H2 += lba_shfl_xor(H2, o);
      H3 += lba_shfl_xor(H3, o); H4 += lba_shfl_xor(H4, o); H5 += lba_shfl_xor(H5, o);
      b0 += lba_shfl_xor(b0, o); b1 += lba_shfl_xor(b1, o); b2 += lba_shfl_xor(b2, o);
    }
  }
  if (live && sub == 0) {
    d.rho_pt[p] = rho_sum;
    if (mode == 0) {
      double* h = d.Hll + 6 * (size_t)p;
      h[0] = H0; h[1] = H1; h[2] = H2; h[3] = H3; h[4] = H4; h[5] = H5;
      d.bl[3 * p] = b0; d.bl[3 * p + 1] = b1; d.bl[3 * p + 2] = b2;
    }
  }
  if (!tail) return;
  __shared__ int s_last;
  __shared__ double s_red[2][128];
  int* ticket = reinterpret_cast<int*>(d.scal + 8);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int t = threadIdx.x;                     // 128 threads: strided partial sums, then a fixed tree
  double a0 = 0, a1 = 0;
  for (int i = t; i < d.n_points; i += 128) {
    a0 += __ldcg(d.rho_pt + i);
    a1 += __ldcg(d.scale_pt + i);
  }
  s_red[0][t] = a0;
  s_red[1][t] = a1;
  __syncthreads();
  for (int s2 = 64; s2 > 0; s2 >>= 1) {
    if (t < s2) {
      s_red[0][t] += s_red[0][t + s2];
      s_red[1][t] += s_red[1][t + s2];
    }
    __syncthreads();
  }
  if (t == 0) {
    d.scal[2] = s_red[0][0];
    d.scal[3] = s_red[1][0];
    *ticket = 0;
    if (post) {
      post[0] = d.scal[0]; post[1] = d.scal[1]; post[2] = s_red[0][0]; post[3] = s_red[1][0];
      post[4] = __ldcg(d.scal + 4); post[5] = __ldcg(d.scal + 5);
      __threadfence_system();
      post[6] = seq;
      __threadfence_system();
    }
  }
}

// One CTA per optimisable camera: Hpp (21 upper-triangular entries, mirrored) and bp (6) summed over the camera's edges.
// Thread t accumulates edges t, t + 256, ... of the camera's (ascending) edge list; the 27 sums are then combined in a
// fixed order -- register butterfly inside each warp (lane i ends with entry i), warps added in warp order -- so the
// result is bit-reproducible without a serial walk over the ~10^3 edges of a keyframe.
__global__ void __launch_bounds__(256) lba_camera_kernel(LbaDev d) {
  __shared__ double sh[8][32];
  pdl_launch_dependents();
  pdl_wait();
  const int s = blockIdx.x, c = d.opt_cams[s], t = threadIdx.x, lane = t & 31, warp = t >> 5;
  double v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0;
  for (int k = d.cam_ptr[c] + t; k < d.cam_ptr[c + 1]; k += 256) {
    const int e = d.cam_edges[k];
    const double* jc = d.Jc + 12 * (size_t)e;
    double J[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) J[i] = jc[i];
    const double wo = d.wo[e], r0 = d.r[2 * e], r1 = d.r[2 * e + 1];
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
      for (int b2 = a; b2 < 6; ++b2) v[q++] += wo * (J[a] * J[b2] + J[6 + a] * J[6 + b2]);
      v[21 + a] += J[a] * r0 + J[6 + a] * r1;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const double send = up ? v[i] : v[i + o];
      const double keep = up ? v[i + o] : v[i];
      v[i] = keep + lba_shfl_xor(send, o);
    }
  }
  sh[warp][lane] = v[0];
  __syncthreads();
  if (t < 27) {
    double tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += sh[w][t];
    if (t < 21) {
      int a = 0, rem = t;
      while (rem >= 6 - a) {
        rem -= 6 - a;
        ++a;
      }
      const int b2 = a + rem;
      d.Hpp[36 * (size_t)s + a * 6 + b2] = tot;
      d.Hpp[36 * (size_t)s + b2 * 6 + a] = tot;
    } else {
      d.bp[6 * (size_t)s + (t - 21)] = tot;
    }
  }
}

// Single-CTA fixed-order reductions: out[0] = sum(v[0..n)), optionally out[1] = max |diag| of Hll and Hpp.
__global__ void lba_reduce_kernel(const double* __restrict__ v, int n, double* __restrict__ out, int out_idx,
                                  LbaDev d, int want_maxdiag) {
  __shared__ double sh[1024];
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x;
  double acc = 0;
  for (int i = t; i < n; i += 1024) acc += v[i];
  sh[t] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (t < s) sh[t] += sh[t + s];
    __syncthreads();
  }
  if (t == 0) out[out_idx] = sh[0];
  if (want_maxdiag) {
    double m = 0;
    for (int p = t; p < d.n_points; p += 1024) {
      const double* h = d.Hll + 6 * (size_t)p;
      m = fmax(m, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
    }
    for (int i = t; i < d.n_opt * 6; i += 1024) m = fmax(m, fabs(d.Hpp[36 * (size_t)(i / 6) + 7 * (i % 6)]));
    __syncthreads();
    sh[t] = m;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (t < s) sh[t] = fmax(sh[t], sh[t + s]);
      __syncthreads();
    }
    if (t == 0) out[1] = sh[0];
  }
}

__device__ __forceinline__ int blk_index(int a, int b, int n) { return a * n - a * (a - 1) / 2 + (b - a); }  // a <= b

#define LBA_MAX_OBS 64

// Schur complement: CTA g walks points g, g+G, ...; for every point adds Hpl_a Dinv Hpl_b^T to block (a,b) of its
// private upper-triangular buffer and Hpl_a Dinv bl to its private right-hand side.
// acc_smem != 0: the CTA's accumulators live in (dynamic) shared memory for the whole walk and are written out once at the
// end -- the read-modify-write of every block entry costs a shared-memory round trip instead of an L2 one; otherwise
// (more cameras than fit) they are accumulated in the CTA's slice of `partial` directly.  Same order either way.
__global__ void __launch_bounds__(256) lba_schur_kernel(LbaDev d, double lambda, int acc_smem) {
  extern __shared__ double s_acc[];
  __shared__ double sBD[LBA_MAX_OBS][18];
  __shared__ double sH[LBA_MAX_OBS][18];
  __shared__ int sSlot[LBA_MAX_OBS];
  __shared__ double sDinv[6], sDb[3];
  __shared__ int sK;
  const int t = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const int acc_n = d.n_blk * 36 + d.n_opt * 6;
  double* gpart = d.partial + (size_t)blockIdx.x * acc_n;
  double* part = acc_smem ? s_acc : gpart;
  double* part_b = part + (size_t)d.n_blk * 36;
  for (int i = t; i < acc_n; i += 256) part[i] = 0.0;   // this CTA's private accumulators
  __syncthreads();
  for (int p = blockIdx.x; p < d.n_points; p += gridDim.x) {
    const int e_begin = d.pt_ptr[p], e_end = d.pt_ptr[p + 1];
    if (t < 32) {
      // warp 0: the point's optimisable observations, in edge order (ballot compaction, 32 edges per round)
      int k = 0;
      for (int e0 = e_begin; e0 < e_end; e0 += 32) {
        const int e2 = e0 + t;
        const int sl = e2 < e_end ? d.edge_slot[e2] : -1;
        const unsigned m = __ballot_sync(0xffffffffu, sl >= 0);
        const int pos = k + __popc(m & ((1u << t) - 1u));
        if (sl >= 0 && pos < LBA_MAX_OBS) sSlot[pos] = sl | (e2 - e_begin) << 16;
        k += __popc(m);
      }
      if (t == 0) sK = k < LBA_MAX_OBS ? k : LBA_MAX_OBS;
    } else if (t == 32) {
      // meanwhile one thread of warp 1: Dinv = (Hll + lambda I)^-1 and Dinv bl
      const double* h = d.Hll + 6 * (size_t)p;
      const double a = h[0] + lambda, b = h[1], c = h[2], e = h[3] + lambda, f = h[4], i = h[5] + lambda;
      const double A = e * i - f * f, B = -(b * i - c * f), C = b * f - c * e;
      const double det = a * A + b * B + c * C;
      const double id = 1.0 / det;
      const double D0 = A * id, D1 = B * id, D2 = C * id;
      const double D3 = (a * i - c * c) * id, D4 = -(a * f - b * c) * id, D5 = (a * e - b * b) * id;
      sDinv[0] = D0; sDinv[1] = D1; sDinv[2] = D2; sDinv[3] = D3; sDinv[4] = D4; sDinv[5] = D5;
      double* dv = d.Dinv + 6 * (size_t)p;
      dv[0] = D0; dv[1] = D1; dv[2] = D2; dv[3] = D3; dv[4] = D4; dv[5] = D5;
      const double l0 = d.bl[3 * p], l1 = d.bl[3 * p + 1], l2 = d.bl[3 * p + 2];
      sDb[0] = D0 * l0 + D1 * l1 + D2 * l2;
      sDb[1] = D1 * l0 + D3 * l1 + D4 * l2;
      sDb[2] = D2 * l0 + D4 * l1 + D5 * l2;
    }
    __syncthreads();
    const int k = sK;
    for (int i = t; i < k * 18; i += 256) {
      const int a = i / 18, ij = i % 18;
      const int e = e_begin + (sSlot[a] >> 16);
      sH[a][ij] = d.Hpl[18 * (size_t)e + ij];
    }
    __syncthreads();
    for (int i = t; i < k * 18; i += 256) {
      const int a = i / 18, ij = i % 18, r = ij / 3, m = ij % 3;
      // BD = Hpl * Dinv (Dinv symmetric: rows (0,1,2),(1,3,4),(2,4,5))
      const double h0 = sH[a][3 * r], h1 = sH[a][3 * r + 1], h2 = sH[a][3 * r + 2];
      const double c0 = m == 0 ? sDinv[0] : (m == 1 ? sDinv[1] : sDinv[2]);
      const double c1 = m == 0 ? sDinv[1] : (m == 1 ? sDinv[3] : sDinv[4]);
      const double c2 = m == 0 ? sDinv[2] : (m == 1 ? sDinv[4] : sDinv[5]);
      sBD[a][ij] = h0 * c0 + h1 * c1 + h2 * c2;
    }
    __syncthreads();
    const int npairs = k * (k + 1) / 2;
    for (int i = t; i < npairs * 36; i += 256) {
      const int pr = i / 36, ij = i % 36, r = ij / 6, c = ij % 6;
      // unrank pr -> (a <= b) in list order
      int a = 0, rem = pr;
      while (rem >= k - a) {
        rem -= k - a;
        ++a;
      }
      const int b = a + rem;
      const int sa = sSlot[a] & 0xFFFF, sb = sSlot[b] & 0xFFFF;
      const double v = sBD[a][3 * r] * sH[b][3 * c] + sBD[a][3 * r + 1] * sH[b][3 * c + 1] +
                       sBD[a][3 * r + 2] * sH[b][3 * c + 2];
      if (sa <= sb) part[(size_t)blk_index(sa, sb, d.n_opt) * 36 + r * 6 + c] += v;
      else part[(size_t)blk_index(sb, sa, d.n_opt) * 36 + c * 6 + r] += v;
    }
    for (int i = t; i < k * 6; i += 256) {
      const int a = i / 6, r = i % 6;
      part_b[(size_t)(sSlot[a] & 0xFFFF) * 6 + r] +=
          sH[a][3 * r] * sDb[0] + sH[a][3 * r + 1] * sDb[1] + sH[a][3 * r + 2] * sDb[2];
    }
    __syncthreads();
  }
  if (acc_smem)
    for (int i = t; i < acc_n; i += 256) gpart[i] = s_acc[i];
}

// Hschur = diag(Hpp + lambda I) - sum_g partial_g (mirrored to the full symmetric matrix), bschur = bp - sum_g.
// Four lanes per output entry: lane q sums the CTA partials g = q, q + 4, ... (eight loads in flight), then the four
// sums are combined in lane order -- fixed order, so the result stays bit-reproducible.
__global__ void lba_schur_reduce_kernel(LbaDev d, double lambda, int G) {
  const int N = 6 * d.n_opt;
  pdl_launch_dependents();
  pdl_wait();
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gi >> 2, q = gi & 3;
  const size_t stride = (size_t)d.n_blk * 36 + (size_t)d.n_opt * 6;
  const bool is_h = i < N * N, is_b = !is_h && i < N * N + N;
  size_t idx = 0;
  int r = 0, c = 0;
  if (is_h) {
    r = i / N;
    c = i % N;
    const int br = r / 6, bc = c / 6, ri = r % 6, ci = c % 6;
    // upper triangle is authoritative (also inside diagonal blocks) so that Hschur is bitwise symmetric
    const bool upper = br < bc || (br == bc && ri <= ci);
    idx = upper ? (size_t)blk_index(br, bc, d.n_opt) * 36 + ri * 6 + ci : (size_t)blk_index(bc, br, d.n_opt) * 36 + ci * 6 + ri;
  } else if (is_b) {
    idx = (size_t)d.n_blk * 36 + (i - N * N);
  }
  double s = 0;
  if (is_h || is_b) {
    int g = q;
    for (; g + 28 < G; g += 32) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = d.partial[(size_t)(g + 4 * u) * stride + idx];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; g < G; g += 4) s += d.partial[(size_t)g * stride + idx];
  }
  // lanes 4k .. 4k+3 hold the four strided sums of one entry (the whole warp takes part in the shuffles)
  const double s1 = __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(s), 1), __shfl_down_sync(0xffffffffu, __double2loint(s), 1));
  const double s2 = __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(s), 2), __shfl_down_sync(0xffffffffu, __double2loint(s), 2));
  const double s3 = __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(s), 3), __shfl_down_sync(0xffffffffu, __double2loint(s), 3));
  if (q != 0) return;
  const double tot = ((s + s1) + s2) + s3;
  if (is_h) {
    const int br = r / 6, bc = c / 6, ri = r % 6, ci = c % 6;
    double base = 0;
    if (br == bc) base = d.Hpp[36 * (size_t)br + ri * 6 + ci] + (ri == ci ? lambda : 0.0);
    d.Hs[i] = base - tot;
  } else if (is_b) {
    d.bs[i - N * N] = d.bp[i - N * N] - tot;
  }
}

// Landmark back-substitution xl = Dinv (bl - sum_a Hpl_a^T xp_a), trial point = point + xl, and the landmark part of
// the gain-ratio denominator sum x (lambda x + b)  (optimization_algorithm_levenberg.cpp:186-199).
__global__ void lba_backsub_kernel(LbaDev d, double lambda) {
  pdl_launch_dependents();
  pdl_wait();
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = gt / LBA_PL, sub = gt % LBA_PL;
  const bool live = p < d.n_points;
  double c0 = 0, c1 = 0, c2 = 0;                 // lane j: edges j, j + LBA_PL, ... of the landmark
  const int e_end = live ? d.pt_ptr[p + 1] : 0;
  for (int e = (live ? d.pt_ptr[p] : 0) + sub; e < e_end; e += LBA_PL) {
    const int s = d.edge_slot[e];
    if (s < 0) continue;
    const double* h = d.Hpl + 18 * (size_t)e;
    const double* x = d.xp + 6 * (size_t)s;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      c0 -= h[3 * i] * x[i];
      c1 -= h[3 * i + 1] * x[i];
      c2 -= h[3 * i + 2] * x[i];
    }
  }
#pragma unroll
  for (int o = LBA_PL / 2; o > 0; o >>= 1) {     // fixed butterfly: bit-reproducible
    c0 += lba_shfl_xor(c0, o);
    c1 += lba_shfl_xor(c1, o);
    c2 += lba_shfl_xor(c2, o);
  }
  if (!live || sub != 0) return;
  c0 += d.bl[3 * p]; c1 += d.bl[3 * p + 1]; c2 += d.bl[3 * p + 2];
  const double* D = d.Dinv + 6 * (size_t)p;
  const double x0 = D[0] * c0 + D[1] * c1 + D[2] * c2;
  const double x1 = D[1] * c0 + D[3] * c1 + D[4] * c2;
  const double x2 = D[2] * c0 + D[4] * c1 + D[5] * c2;
  d.points_t[3 * p] = d.points[3 * p] + x0;
  d.points_t[3 * p + 1] = d.points[3 * p + 1] + x1;
  d.points_t[3 * p + 2] = d.points[3 * p + 2] + x2;
  d.scale_pt[p] = x0 * (lambda * x0 + d.bl[3 * p]) + x1 * (lambda * x1 + d.bl[3 * p + 1]) +
                  x2 * (lambda * x2 + d.bl[3 * p + 2]);
}

__device__ void po_pose_oplus(const double* pose, const double* u, double* out);

// rho and scale sums of one LM trial in one launch (same strided + tree order as lba_reduce_kernel)
// post != null: thread 0 then posts out[0..5] and the sequence number into page-locked host memory (the host polls it).
__global__ void lba_reduce2_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ out,
                                   int ia, int ib, volatile double* post, double seq) {
  __shared__ double sh[1024];
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x;
  for (int which = 0; which < 2; ++which) {
    const double* v = which ? b : a;
    double acc = 0;
    for (int i = t; i < n; i += 1024) acc += v[i];
    __syncthreads();
    sh[t] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (t < s) sh[t] += sh[t + s];
      __syncthreads();
    }
    if (t == 0) out[which ? ib : ia] = sh[0];
  }
  if (t == 0 && post) {
#pragma unroll
    for (int i = 0; i < 6; ++i) post[i] = out[i];
    __threadfence_system();
    post[6] = seq;
    __threadfence_system();
  }
}

// The reduced camera system on the device (g2o: LinearSolverEigen::solve, Thirdparty/g2o/g2o/solvers/
// linear_solver_eigen.h:94-124, a sparse LDLT of the same matrix): one CTA, Hschur as a packed lower triangle in shared
// memory (N <= 192: 148 KB), blocked right-looking Cholesky with the forward substitution carried along, backward
// substitution by block inverses, then the pose update
// exp(xp) * T of every optimisable camera (types_six_dof_expmap.h:73-76) and the camera part of the gain-ratio
// denominator.  scal[4] = sum xp (lambda xp + bp), scal[5] = 1 if the factorisation succeeded (else xp = 0).
#define LBA_SOLVE_MAX_N 192
__device__ __forceinline__ int lcol(int k, int N) { return k * N - k * (k - 1) / 2; }   // start of packed column k (rows k..N-1)

#define T6(c, r) ((c) * 6 - (c) * ((c) - 1) / 2 + (r) - (c))   // 6 x 6 lower triangle packed by columns, r >= c

__global__ void __launch_bounds__(256) lba_solve_kernel(LbaDev d, double lambda) {
  extern __shared__ double sL[];                 // lower triangle packed BY COLUMNS: L[i][k] at lcol(k) + i - k, so the
                                                 // rows i = j + tid of one column step read consecutive words
  __shared__ double s_b[LBA_SOLVE_MAX_N], s_y[LBA_SOLVE_MAX_N], s_invd[LBA_SOLVE_MAX_N];
  __shared__ double s_linv[LBA_SOLVE_MAX_N / 6][21];   // inverses of the factored diagonal blocks (backward substitution)
  __shared__ int s_fail;
  const int N = 6 * d.n_opt, t = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = t; i < N; i += 256) s_b[i] = d.bs[i];
  if (t == 0) s_fail = 0;
  // Hs is bitwise symmetric: the lower triangle is read row by row (coalesced), eight loads in flight per thread
#pragma unroll 8
  for (int q = t; q < N * N; q += 256) {
    const int r = q / N, c = q - r * N;
    if (c <= r) sL[lcol(c, N) + r - c] = __ldg(d.Hs + q);
  }
  __syncthreads();
  // Blocked right-looking Cholesky on the 6 x 6 camera blocks (N = 6 n_opt) with one block of look-ahead.  Per block
  // column: (i) one thread per row below solves its 1 x 6 panel row while thread 255 runs the forward substitution of
  // this block (b is carried along like one more matrix row, so L y = b costs no pass of its own); (ii) warps 1..7 apply
  // the rank-6 update to the trailing triangle -- a CONTIGUOUS range of the column-packed array -- and to b, while warp
  // 0 updates just the NEXT diagonal block and one of its lanes factors it in registers (the only serial part: six
  // dependent pivots), so the pivots hide behind the update.
  const int lane = t & 31;
  auto factor_diag = [&](int j0) {                 // ONE thread: in-place Cholesky of the 6 x 6 block at (j0, j0)
    double a[21];
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = c; r < 6; ++r) a[T6(c, r)] = sL[lcol(j0 + c, N) + r - c];
    bool fail = false;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double dj = a[T6(c, c)];
      const bool bad = !(dj > 0.0) || !isfinite(dj);
      const double inv = bad ? 0.0 : rsqrt(dj);
      fail |= bad;
      s_invd[j0 + c] = inv;
      a[T6(c, c)] = bad ? 1.0 : dj * inv;
#pragma unroll
      for (int r = c + 1; r < 6; ++r) a[T6(c, r)] *= inv;
#pragma unroll
      for (int c2 = c + 1; c2 < 6; ++c2)
#pragma unroll
        for (int r2 = c2; r2 < 6; ++r2) a[T6(c2, r2)] -= a[T6(c, r2)] * a[T6(c, c2)];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
      for (int r = c; r < 6; ++r) sL[lcol(j0 + c, N) + r - c] = a[T6(c, r)];
    if (fail) s_fail = 1;
  };
  if (t == 0 && N > 0) factor_diag(0);
  __syncthreads();
  for (int j0 = 0; j0 < N; j0 += 6) {
    if (s_fail) break;                             // uniform: written before the last barrier
    {
      const int i = j0 + 6 + t;                    // panel: row i of L[:, j0 .. j0+5]
      if (i < N) {
        double l[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double v = sL[lcol(j0 + c, N) + i - (j0 + c)];
#pragma unroll
          for (int c1 = 0; c1 < c; ++c1) v -= l[c1] * sL[lcol(j0 + c1, N) + (c - c1)];   // L[j0+c][j0+c1]
          l[c] = v * s_invd[j0 + c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) sL[lcol(j0 + c, N) + i - (j0 + c)] = l[c];
      } else if (t == 255) {                       // forward substitution of this block: y = L_jj^-1 b_j (N - 6 < 255)
        double y[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double v = s_b[j0 + c];
#pragma unroll
          for (int c1 = 0; c1 < c; ++c1) v -= sL[lcol(j0 + c1, N) + (c - c1)] * y[c1];
          y[c] = v * s_invd[j0 + c];
          s_y[j0 + c] = y[c];
        }
      }
    }
    __syncthreads();
    if (j0 + 6 < N) {
      const double* col[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) col[c] = sL + lcol(j0 + c, N) - (j0 + c);   // col[c][r] = L[r][j0 + c]
      if (t < 32) {
        // next diagonal block (rows and columns j0+6 .. j0+11): 21 entries, then its factorisation
        if (lane < 21) {
          int c2 = 0, e = lane;
          while (e >= 6 - c2) {
            e -= 6 - c2;
            ++c2;
          }
          const int cc = j0 + 6 + c2, r = cc + e;
          double acc = sL[lcol(cc, N) + r - cc];
#pragma unroll
          for (int c = 0; c < 6; ++c) acc -= col[c][r] * col[c][cc];
          sL[lcol(cc, N) + r - cc] = acc;
        }
        __syncwarp();
        if (lane == 0) factor_diag(j0 + 6);
      } else {
        {
          const int i = j0 + 6 + (t - 32);         // b below this block (N - 6 <= 186 rows: one pass of the 224 threads)
          if (i < N) {
            double v = s_b[i];
#pragma unroll
            for (int c = 0; c < 6; ++c) v -= col[c][i] * s_y[j0 + c];
            s_b[i] = v;
          }
        }
        // trailing triangle by 6-wide block columns: warps 1..7 deal them round-robin; a warp keeps the block column's
        // 6 x 6 factor rows W[k][c] = L[6 bj + k][j0 + c] in registers and its lanes walk down the rows below, so an
        // element costs 1 + 1/6 shared-memory operand loads instead of 12 (the rank-6 update was LSU-bound)
        const int jb = j0 / 6, nbk = N / 6, w = (t >> 5) - 1;
        for (int bj = jb + 1 + w; bj < nbk; bj += 7) {
          double W[6][6];
#pragma unroll
          for (int k = 0; k < 6; ++k)
#pragma unroll
            for (int c = 0; c < 6; ++c) W[k][c] = col[c][6 * bj + k];
          const int cbase = 6 * bj;
          double* dcol[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) dcol[k] = sL + lcol(cbase + k, N) - (cbase + k);   // dcol[k][r] = A[r][cbase + k]
          for (int r = (bj == jb + 1 ? j0 + 12 : cbase) + lane; r < N; r += 32) {   // the next diagonal block is warp 0's
            // all six outputs of the row at once (six independent accumulation chains; a branch per output would
            // serialise them): entries above the diagonal (inside this column's diagonal block) are computed on a
            // dummy and not stored
            double a[6], acc[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) a[c] = col[c][r];
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[k] = cbase + k <= r ? dcol[k][r] : 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c)
#pragma unroll
              for (int k = 0; k < 6; ++k) acc[k] -= a[c] * W[k][c];
#pragma unroll
            for (int k = 0; k < 6; ++k)
              if (cbase + k <= r) dcol[k][r] = acc[k];
          }
        }
      }
    }
    __syncthreads();
  }
  if (!s_fail) {
    // inverses of the diagonal blocks, one thread per camera (lower triangular, column by column)
    if (t < d.n_opt) {
      const int j0 = 6 * t;
      double l[21], m[21];
#pragma unroll
      for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int r = c; r < 6; ++r) l[T6(c, r)] = sL[lcol(j0 + c, N) + r - c];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        m[T6(c, c)] = s_invd[j0 + c];
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
          double acc = 0.0;
#pragma unroll
          for (int k = c; k < r; ++k) acc += l[T6(k, r)] * m[T6(c, k)];
          m[T6(c, r)] = -acc * s_invd[j0 + r];
        }
      }
#pragma unroll
      for (int e = 0; e < 21; ++e) s_linv[t][e] = m[e];
    }
    __syncthreads();
    // backward L^T x = y by blocks: six threads apply the block inverse (x_c = sum_{r >= c} Linv[r][c] v_r), then every
    // row above takes the rank-6 correction
    for (int j0 = N - 6; j0 >= 0; j0 -= 6) {
      if (t < 6) {
        const double* m = s_linv[j0 / 6];
        double x = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
          if (r >= t) x += m[T6(t, r)] * s_y[j0 + r];
        s_b[j0 + t] = x;
      }
      __syncthreads();
      if (t < j0) {
        double v = s_y[t];
#pragma unroll
        for (int c = 0; c < 6; ++c) v -= sL[lcol(t, N) + (j0 + c) - t] * s_b[j0 + c];        // L[j0+c][t]
        s_y[t] = v;
      }
      __syncthreads();
    }
  }
  for (int i = t; i < N; i += 256) d.xp[i] = s_fail ? 0.0 : s_b[i];
  // trial poses: every camera copied, optimisable ones updated
  for (int i = t; i < d.n_cams * 7; i += 256) d.poses_t[i] = d.poses[i];
  __syncthreads();
  if (t < d.n_opt) {
    double u[6];
    for (int i = 0; i < 6; ++i) u[i] = s_fail ? 0.0 : s_b[6 * t + i];
    const int c = d.opt_cams[t];
    po_pose_oplus(d.poses + 7 * (size_t)c, u, d.poses_t + 7 * (size_t)c);
  } else if (t >= 224) {                           // the camera part of the gain-ratio denominator: fixed-order warp sum
    double sc = 0;
    if (!s_fail)
      for (int i = lane; i < N; i += 32) sc += s_b[i] * (lambda * s_b[i] + d.bp[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      sc += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(sc), o), __shfl_xor_sync(0xffffffffu, __double2loint(sc), o));
    if (lane == 0) {
      d.scal[4] = sc;
      d.scal[5] = s_fail ? 0.0 : 1.0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host helpers
static bool cholesky_solve(std::vector<double>& A, std::vector<double>& b, int n) {  // A row-major SPD, in place
  for (int j = 0; j < n; ++j) {
    double s = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) s -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(s > 0.0) || !std::isfinite(s)) return false;
    const double l = sqrt(s);
    A[(size_t)j * n + j] = l;
    for (int i = j + 1; i < n; ++i) {
      double t = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) t -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = t / l;
    }
  }
  for (int i = 0; i < n; ++i) {
    double t = b[i];
    for (int k = 0; k < i; ++k) t -= A[(size_t)i * n + k] * b[k];
    b[i] = t / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double t = b[i];
    for (int k = i + 1; k < n; ++k) t -= A[(size_t)k * n + i] * b[k];
    b[i] = t / A[(size_t)i * n + i];
  }
  return true;
}

static void h_quat_to_R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen Quaternion(Matrix3) (Shepperd), (x,y,z,w)
static void h_R_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}
static void h_normalize_rot(double* q) {
  if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}
static void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// estimate <- exp(update) * estimate  (types_six_dof_expmap.h:73-76, se3quat.h:223-257); update = [omega, upsilon]
static void h_pose_oplus(const double* pose, const double* u, double* out) {
  const double w0 = u[0], w1 = u[1], w2 = u[2];
  const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
  const double Om[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
  double Om2[9], R[9], V[9];
  mat3_mul(Om, Om, Om2);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta < 0.00001) {
    for (int i = 0; i < 9; ++i) R[i] = I[i] + Om[i] + Om2[i];
    for (int i = 0; i < 9; ++i) V[i] = R[i];
  } else {
    double sn, cs;
    sincos(theta, &sn, &cs);   // one argument reduction for both
    const double a = sn / theta, b = (1 - cs) / (theta * theta);
    const double c = (theta - sn) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) {
      R[i] = I[i] + a * Om[i] + b * Om2[i];
      V[i] = I[i] + b * Om[i] + c * Om2[i];
    }
  }
  double qe[4];
  h_R_to_quat(R, qe);
  h_normalize_rot(qe);
  const double te[3] = {V[0] * u[3] + V[1] * u[4] + V[2] * u[5], V[3] * u[3] + V[4] * u[4] + V[5] * u[5],
                        V[6] * u[3] + V[7] * u[4] + V[8] * u[5]};
  // q = qe * q_pose (Hamilton, xyzw)
  const double ax = qe[0], ay = qe[1], az = qe[2], aw = qe[3];
  const double bx = pose[0], by = pose[1], bz = pose[2], bw = pose[3];
  double q[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                 aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
  double Re[9];
  h_quat_to_R(qe, Re);
  const double* t = pose + 4;
  out[4] = te[0] + Re[0] * t[0] + Re[1] * t[1] + Re[2] * t[2];
  out[5] = te[1] + Re[3] * t[0] + Re[4] * t[1] + Re[5] * t[2];
  out[6] = te[2] + Re[6] * t[0] + Re[7] * t[1] + Re[8] * t[2];
  h_normalize_rot(q);
  for (int i = 0; i < 4; ++i) out[i] = q[i];
}

struct LbaHost {
  std::vector<int> pt_ptr, cam_ptr, cam_edges, cam_slot, opt_cams, edge_slot;
};

static int lba_setup(hfb_ctx* ctx, const hfb_lba_problem* pr, LbaDev& d, LbaHost& h, uint8_t** arena_out) {
  HFB_REQUIRE(ctx, pr && pr->n_cams > 0 && pr->n_points >= 0 && pr->n_edges >= 0, "bad problem sizes");
  HFB_REQUIRE(ctx, pr->poses && pr->fixed && (pr->n_points == 0 || pr->points) &&
                       (pr->n_edges == 0 || (pr->edge_cam && pr->edge_point && pr->obs && pr->inv_sigma2)),
              "null problem array");
  const int nc = pr->n_cams, np = pr->n_points, ne = pr->n_edges;
  h.pt_ptr.assign(np + 1, 0);
  h.cam_ptr.assign(nc + 1, 0);
  for (int e = 0; e < ne; ++e) {
    const int p = pr->edge_point[e], c = pr->edge_cam[e];
    HFB_REQUIRE(ctx, p >= 0 && p < np && c >= 0 && c < nc, "edge index out of range");
    HFB_REQUIRE(ctx, e == 0 || p >= pr->edge_point[e - 1], "edges must be sorted by point");
    h.pt_ptr[p + 1]++;
    h.cam_ptr[c + 1]++;
  }
  for (int p = 0; p < np; ++p) h.pt_ptr[p + 1] += h.pt_ptr[p];
  for (int c = 0; c < nc; ++c) h.cam_ptr[c + 1] += h.cam_ptr[c];
  h.cam_edges.resize(ne);
  {
    std::vector<int> fill(h.cam_ptr.begin(), h.cam_ptr.end() - 1);
    for (int e = 0; e < ne; ++e) h.cam_edges[fill[pr->edge_cam[e]]++] = e;
  }
  h.cam_slot.assign(nc, -1);
  h.opt_cams.clear();
  for (int c = 0; c < nc; ++c)
    if (!pr->fixed[c]) {
      h.cam_slot[c] = (int)h.opt_cams.size();
      h.opt_cams.push_back(c);
    }
  const int no = (int)h.opt_cams.size();
  HFB_REQUIRE(ctx, no <= 4096, "too many optimisable cameras");
  for (int p = 0; p < np; ++p) {
    int k = 0;
    for (int e = h.pt_ptr[p]; e < h.pt_ptr[p + 1]; ++e) k += h.cam_slot[pr->edge_cam[e]] >= 0;
    HFB_REQUIRE(ctx, k <= LBA_MAX_OBS, "a landmark has more than 64 observations from optimisable keyframes");
  }
  d.n_cams = nc; d.n_points = np; d.n_edges = ne; d.n_opt = no;
  d.n_blk = no * (no + 1) / 2;
  d.G = std::max(1, std::min(2 * ctx->n_sm, np));
  if (const char* ge = getenv("HFB_LBA_G")) d.G = std::max(1, std::min(atoi(ge), np));   // experiment: CTAs (= partial buffers) of the Schur kernel
  h.edge_slot.resize(ne);
  for (int e = 0; e < ne; ++e) h.edge_slot[e] = h.cam_slot[pr->edge_cam[e]];
  for (int i = 0; i < 4; ++i) d.K[i] = (double)pr->K[i];
  d.delta = pr->huber_delta;
  // one arena
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t N = (size_t)6 * no;
  const size_t part_stride = (size_t)d.n_blk * 36 + (size_t)no * 6;
  const size_t o_poses = take(nc * 7 * 8), o_poses_t = take(nc * 7 * 8), o_points = take((size_t)np * 24),
               o_points_t = take((size_t)np * 24), o_ecam = take((size_t)ne * 4), o_ptptr = take((size_t)(np + 1) * 4),
               o_camptr = take((size_t)(nc + 1) * 4), o_camedges = take((size_t)ne * 4), o_slot = take((size_t)nc * 4),
               o_opt = take((size_t)std::max(no, 1) * 4), o_eslot = take((size_t)ne * 4 + 4), o_obs = take((size_t)ne * 16), o_is2 = take((size_t)ne * 8),
               o_Jc = take((size_t)ne * 96), o_wo = take((size_t)ne * 8), o_r = take((size_t)ne * 16),
               o_Hpl = take((size_t)ne * 144), o_chia = take((size_t)ne * 8), o_chib = take((size_t)ne * 8),
               o_Hll = take((size_t)np * 48), o_bl = take((size_t)np * 24), o_Dinv = take((size_t)np * 48),
               o_rho = take((size_t)np * 8), o_scale = take((size_t)np * 8), o_Hpp = take((size_t)no * 288 + 8),
               o_bp = take((size_t)no * 48 + 8), o_part = take((size_t)d.G * part_stride * 8 + 8),
               o_Hs = take(N * N * 8 + 8), o_bs = take(N * 8 + 8), o_xp = take(N * 8 + 8), o_scal = take(256),
               o_depth = take((size_t)ne + 8);
  HFB_TRY(ctx->ensure_scratch(off));
  uint8_t* a = reinterpret_cast<uint8_t*>(ctx->d_scratch);
  *arena_out = a;
  d.poses = (double*)(a + o_poses); d.poses_t = (double*)(a + o_poses_t);
  d.points = (double*)(a + o_points); d.points_t = (double*)(a + o_points_t);
  d.edge_cam = (int*)(a + o_ecam); d.pt_ptr = (int*)(a + o_ptptr); d.cam_ptr = (int*)(a + o_camptr);
  d.cam_edges = (int*)(a + o_camedges); d.cam_slot = (int*)(a + o_slot); d.opt_cams = (int*)(a + o_opt);
  d.edge_slot = (int*)(a + o_eslot);
  d.obs = (double*)(a + o_obs); d.invs2 = (double*)(a + o_is2);
  d.Jc = (double*)(a + o_Jc); d.wo = (double*)(a + o_wo); d.r = (double*)(a + o_r); d.Hpl = (double*)(a + o_Hpl);
  d.chi2_a = (double*)(a + o_chia); d.chi2_b = (double*)(a + o_chib);
  d.Hll = (double*)(a + o_Hll); d.bl = (double*)(a + o_bl); d.Dinv = (double*)(a + o_Dinv);
  d.rho_pt = (double*)(a + o_rho); d.scale_pt = (double*)(a + o_scale);
  d.Hpp = (double*)(a + o_Hpp); d.bp = (double*)(a + o_bp); d.partial = (double*)(a + o_part);
  d.Hs = (double*)(a + o_Hs); d.bs = (double*)(a + o_bs); d.xp = (double*)(a + o_xp); d.scal = (double*)(a + o_scal);
  d.depth_ok = (unsigned char*)(a + o_depth);
  cudaStream_t st = ctx->stream;
  HFB_CUDA(ctx, cudaMemsetAsync(d.scal, 0, 256, st));   // scalars + the completion ticket of the trial's last kernel (scal + 8)
  HFB_CUDA(ctx, cudaMemcpyAsync(d.poses, pr->poses, (size_t)nc * 56, cudaMemcpyHostToDevice, st));
  if (np) HFB_CUDA(ctx, cudaMemcpyAsync(d.points, pr->points, (size_t)np * 24, cudaMemcpyHostToDevice, st));
  if (ne) {
    HFB_CUDA(ctx, cudaMemcpyAsync(d.edge_cam, pr->edge_cam, (size_t)ne * 4, cudaMemcpyHostToDevice, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(d.cam_edges, h.cam_edges.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(d.edge_slot, h.edge_slot.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(d.obs, pr->obs, (size_t)ne * 16, cudaMemcpyHostToDevice, st));
    HFB_CUDA(ctx, cudaMemcpyAsync(d.invs2, pr->inv_sigma2, (size_t)ne * 8, cudaMemcpyHostToDevice, st));
  }
  HFB_CUDA(ctx, cudaMemcpyAsync(d.pt_ptr, h.pt_ptr.data(), (size_t)(np + 1) * 4, cudaMemcpyHostToDevice, st));
  HFB_CUDA(ctx, cudaMemcpyAsync(d.cam_ptr, h.cam_ptr.data(), (size_t)(nc + 1) * 4, cudaMemcpyHostToDevice, st));
  HFB_CUDA(ctx, cudaMemcpyAsync(d.cam_slot, h.cam_slot.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, st));
  if (no) HFB_CUDA(ctx, cudaMemcpyAsync(d.opt_cams, h.opt_cams.data(), (size_t)no * 4, cudaMemcpyHostToDevice, st));
  HFB_CUDA(ctx, cudaStreamSynchronize(st));  // host staging vectors may now go out of scope
  return HFB_OK;
}

// computeActiveErrors + buildSystem at (d.poses, d.points): fills Hll/bl/Hpl/Hpp/bp, chi2 into chi2_buf,
// scal[0] = robust chi2, scal[1] = max |diag|.
static int lba_build(hfb_ctx* ctx, LbaDev& d, double* chi2_buf, int want_maxdiag, bool need_scalars = true) {
  if (d.n_points > 0) {
    hfb_launch(ctx, lba_linearize_kernel, dim3(ceil_div(d.n_points * LBA_PL, 128)), dim3(128), 0, d, (const double*)d.poses, (const double*)d.points, chi2_buf, 0, 0, (volatile double*)nullptr, 0.0);
    HFB_CHECK_LAUNCH(ctx, "lba_linearize");
  }
  if (d.n_opt > 0) {
    hfb_launch(ctx, lba_camera_kernel, dim3(d.n_opt), dim3(256), 0, d);
    HFB_CHECK_LAUNCH(ctx, "lba_camera");
  }
  if (need_scalars) {   // robust chi2 / largest diagonal entry of the linearisation: nobody reads them after iteration 0
    hfb_launch(ctx, lba_reduce_kernel, dim3(1), dim3(1024), 0, (const double*)d.rho_pt, d.n_points, d.scal, 0, d, want_maxdiag);
    HFB_CHECK_LAUNCH(ctx, "lba_reduce");
  }
  return HFB_OK;
}

static int lba_schur(hfb_ctx* ctx, LbaDev& d, double lambda) {
  if (d.n_points > 0) {
    // zeroes its CTA-private accumulators itself; they sit in shared memory when they fit beside a second CTA
    const size_t acc_bytes = ((size_t)d.n_blk * 36 + (size_t)d.n_opt * 6) * 8;
    const int acc_smem = acc_bytes <= 90 * 1024 ? 1 : 0;
    if (acc_smem) {
      static SmemOptIn optin;
      HFB_CUDA(ctx, optin.ensure(lba_schur_kernel, ctx->device, 90 * 1024));
    }
    hfb_launch(ctx, lba_schur_kernel, dim3(d.G), dim3(256), acc_smem ? acc_bytes : 0, d, lambda, acc_smem);
    HFB_CHECK_LAUNCH(ctx, "lba_schur");
  } else {
    const size_t part_stride = (size_t)d.n_blk * 36 + (size_t)d.n_opt * 6;
    HFB_CUDA(ctx, cudaMemsetAsync(d.partial, 0, (size_t)d.G * part_stride * 8, ctx->stream));
  }
  const int N = 6 * d.n_opt;
  if (N > 0) {
    hfb_launch(ctx, lba_schur_reduce_kernel, dim3(ceil_div(4 * (N * N + N), 256)), dim3(256), 0, d, lambda, d.G);
    HFB_CHECK_LAUNCH(ctx, "lba_schur_reduce");
  }
  return HFB_OK;
}

extern "C" int hfb_lba_build_schur(hfb_ctx* ctx, const hfb_lba_problem* problem, double lambda, double* Hschur,
                                   double* bschur, double* robust_chi2, int32_t* n_opt_cams) {
  HFB_ENTER(ctx);
  LbaDev d;
  LbaHost h;
  uint8_t* arena = nullptr;
  HFB_TRY(lba_setup(ctx, problem, d, h, &arena));
  HFB_TRY(lba_build(ctx, d, d.chi2_a, 0));
  HFB_TRY(lba_schur(ctx, d, lambda));
  const size_t N = (size_t)6 * d.n_opt;
  if (n_opt_cams) *n_opt_cams = d.n_opt;
  if (N && Hschur) HFB_CUDA(ctx, cudaMemcpyAsync(Hschur, d.Hs, N * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (N && bschur) HFB_CUDA(ctx, cudaMemcpyAsync(bschur, d.bs, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (robust_chi2) HFB_CUDA(ctx, cudaMemcpyAsync(robust_chi2, d.scal, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HFB_OK;
}

extern "C" int hfb_lba_optimize(hfb_ctx* ctx, const hfb_lba_problem* problem, int32_t iterations,
                                double user_lambda_init, const volatile uint8_t* stop_flag, double* poses_out,
                                double* points_out, double* chi2_out, uint8_t* depth_positive_out,
                                hfb_lba_stats* stats) {
  HFB_ENTER(ctx);
  LbaDev d;
  LbaHost h;
  uint8_t* arena = nullptr;
  const uint64_t launches0 = ctx->launches;
  HFB_TRY(lba_setup(ctx, problem, d, h, &arena));
  cudaStream_t st = ctx->stream;
  if (!ctx->h_post) {
    HFB_CUDA(ctx, cudaMallocHost(&ctx->h_post, 8 * sizeof(double)));
    memset(ctx->h_post, 0, 8 * sizeof(double));
  }
  const int no = d.n_opt, N = 6 * no, nc = d.n_cams;
  // The reduced system is solved on the device (lba_solve_kernel) unless it does not fit one CTA's shared memory or the
  // host-solve mode is requested (HFB_LBA_HOST_SOLVE=1: the round-1 path, kept as the cross-check of the device solver).
  const char* hs_env = getenv("HFB_LBA_HOST_SOLVE");
  const size_t solve_smem = (size_t)N * (N + 1) / 2 * 8 + 16;   // packed triangle (doubles)
  const bool dev_solve = N > 0 && N <= LBA_SOLVE_MAX_N && !(hs_env && hs_env[0] == '1');
  if (dev_solve) {
    static SmemOptIn optin;
    HFB_CUDA(ctx, optin.ensure(lba_solve_kernel, ctx->device, (size_t)LBA_SOLVE_MAX_N * (LBA_SOLVE_MAX_N + 1) / 2 * 8 + 16));
  }
  std::vector<double> poses(problem->poses, problem->poses + (size_t)nc * 7), poses_t(poses);
  std::vector<double> Hs((size_t)N * N), bs(N), bp(N), xp(N);
  auto terminate = [&]() { return stop_flag && *stop_flag; };
  const double tau = 1e-5, good_lo = 1.0 / 3.0, good_hi = 2.0 / 3.0;
  const int max_trials = 10;
  double lambda = 0.0, ni = 2.0;
  int n_bad = 0, it_done = 0, trials = 0;
  double* chi2_last = d.chi2_a;   // buffer holding the most recent computeActiveErrors result
  double* chi2_other = d.chi2_b;
  double initial_chi = 0.0, current_chi = 0.0;
  double scal[6];
  // chi2 at the initial estimate: the first linearisation computes it too, so the separate evaluation only runs when
  // there is no first iteration (iterations == 0 or an immediate stop still report it)
  const bool first_build_has_chi = iterations > 0 && !terminate();
  if (!first_build_has_chi) {
    if (d.n_points > 0) {
      hfb_launch(ctx, lba_linearize_kernel, dim3(ceil_div(d.n_points * LBA_PL, 128)), dim3(128), 0, d, (const double*)d.poses, (const double*)d.points, chi2_last, 1, 0, (volatile double*)nullptr, 0.0);
      HFB_CHECK_LAUNCH(ctx, "lba_errors");
    }
    hfb_launch(ctx, lba_reduce_kernel, dim3(1), dim3(1024), 0, (const double*)d.rho_pt, d.n_points, d.scal, 0, d, 0);
    HFB_CHECK_LAUNCH(ctx, "lba_reduce");
    HFB_CUDA(ctx, cudaMemcpyAsync(scal, d.scal, 8, cudaMemcpyDeviceToHost, st));
    HFB_CUDA(ctx, cudaStreamSynchronize(st));
    initial_chi = current_chi = scal[0];
  }
  for (int it = 0; it < iterations; ++it) {
    if (!(it == 0 && first_build_has_chi) && terminate()) break;
    HFB_TRY(lba_build(ctx, d, chi2_last, it == 0, it == 0 || !dev_solve));
    // The host needs the linearisation's scalars only at the first iteration (lambda from the largest diagonal entry) or
    // when it solves the reduced system itself: afterwards the robust chi2 of the estimate IS the accepted trial's
    // (same kernel, same summation order: bitwise equal), so later iterations enqueue straight through to the trial.
    if (it == 0 || !dev_solve) {
      HFB_CUDA(ctx, cudaMemcpyAsync(scal, d.scal, 16, cudaMemcpyDeviceToHost, st));
      if (N && !dev_solve) HFB_CUDA(ctx, cudaMemcpyAsync(bp.data(), d.bp, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
      HFB_CUDA(ctx, cudaStreamSynchronize(st));
      current_chi = scal[0];
      if (it == 0 && first_build_has_chi) initial_chi = current_chi;
    }
    const double ini_chi = current_chi;
    if (it == 0) {
      lambda = user_lambda_init > 0 ? user_lambda_init : tau * scal[1];
      ni = 2.0;
      n_bad = 0;
    }
    double rho = 0.0;
    int qmax = 0;
    do {
      HFB_TRY(lba_schur(ctx, d, lambda));
      bool ok2 = true;
      if (dev_solve) {
        hfb_launch(ctx, lba_solve_kernel, dim3(1), dim3(256), solve_smem, d, lambda);
        HFB_CHECK_LAUNCH(ctx, "lba_solve");
      } else {
        if (N) {
          HFB_CUDA(ctx, cudaMemcpyAsync(Hs.data(), d.Hs, (size_t)N * N * 8, cudaMemcpyDeviceToHost, st));
          HFB_CUDA(ctx, cudaMemcpyAsync(bs.data(), d.bs, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
          HFB_CUDA(ctx, cudaStreamSynchronize(st));
          xp = bs;
          ok2 = cholesky_solve(Hs, xp, N);
          if (!ok2) std::fill(xp.begin(), xp.end(), 0.0);
          HFB_CUDA(ctx, cudaMemcpyAsync(d.xp, xp.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
        }
        poses_t = poses;
        for (int s = 0; s < no; ++s) h_pose_oplus(&poses[(size_t)h.opt_cams[s] * 7], &xp[(size_t)6 * s], &poses_t[(size_t)h.opt_cams[s] * 7]);
        HFB_CUDA(ctx, cudaMemcpyAsync(d.poses_t, poses_t.data(), (size_t)nc * 56, cudaMemcpyHostToDevice, st));
      }
      // the trial's scalars come back through page-locked memory the last kernel writes itself; the host polls the
      // sequence number (no copy operation, no stream synchronisation, no reduction launch per trial)
      const double seq = (double)(++ctx->post_seq);
      if (d.n_points > 0) {
        hfb_launch(ctx, lba_backsub_kernel, dim3(ceil_div(d.n_points * LBA_PL, 128)), dim3(128), 0, d, lambda);
        HFB_CHECK_LAUNCH(ctx, "lba_backsub");
        hfb_launch(ctx, lba_linearize_kernel, dim3(ceil_div(d.n_points * LBA_PL, 128)), dim3(128), 0, d, (const double*)d.poses_t,
                   (const double*)d.points_t, chi2_other, 1, 1, (volatile double*)ctx->h_post, seq);
        HFB_CHECK_LAUNCH(ctx, "lba_errors");
      } else {
        hfb_launch(ctx, lba_reduce2_kernel, dim3(1), dim3(1024), 0, (const double*)d.rho_pt, (const double*)d.scale_pt, d.n_points, d.scal, 2, 3,
                   (volatile double*)ctx->h_post, seq);
        HFB_CHECK_LAUNCH(ctx, "lba_reduce2");
      }
      std::swap(chi2_last, chi2_other);
      {
        volatile double* hp = ctx->h_post;
        uint32_t spins = 0;
        while (hp[6] != seq) {
          if ((++spins & 0x3ff) == 0) {          // every ~1000 polls: has the stream died or finished without posting?
            const cudaError_t q = cudaStreamQuery(st);
            if (q == cudaSuccess) {
              if (hp[6] == seq) break;
              ctx->set_error("local BA: the trial finished without posting its result");
              return HFB_ERR_CUDA;
            }
            if (q != cudaErrorNotReady) {
              ctx->set_error(std::string("local BA trial: ") + cudaGetErrorString(q));
              return HFB_ERR_CUDA;
            }
          }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        for (int i = 0; i < 6; ++i) scal[i] = hp[i];
      }
      if (dev_solve) ok2 = scal[5] != 0.0;
      ++trials;
      const double temp_chi = ok2 ? scal[2] : 1.7976931348623157e308;
      double scale = scal[3] + 1e-3;
      if (dev_solve) scale += scal[4];
      else
        for (int i = 0; i < N; ++i) scale += xp[i] * (lambda * xp[i] + bp[i]);
      rho = (current_chi - temp_chi) / scale;
      if (rho > 0 && std::isfinite(temp_chi)) {
        double alpha = 1.0 - pow(2 * rho - 1, 3);
        alpha = std::min(alpha, good_hi);
        lambda *= std::max(good_lo, alpha);
        ni = 2.0;
        current_chi = temp_chi;
        if (!dev_solve) poses = poses_t;
        std::swap(d.poses, d.poses_t);
        std::swap(d.points, d.points_t);
      } else {
        lambda *= ni;
        ni *= 2;
      }
      ++qmax;
    } while (rho < 0 && qmax < max_trials && !terminate());
    ++it_done;
    if (qmax == max_trials || rho == 0) break;
    if ((ini_chi - current_chi) * 1e3 < ini_chi) ++n_bad;
    else n_bad = 0;
    if (n_bad >= 3) break;
  }
  // outputs: final estimate, cached chi2 of the LAST error evaluation (src/Optimizer.cc:1425 quirk), depth test
  if (d.n_points > 0 && depth_positive_out) {
    hfb_launch(ctx, lba_linearize_kernel, dim3(ceil_div(d.n_points * LBA_PL, 128)), dim3(128), 0, d, (const double*)d.poses, (const double*)d.points, chi2_other, 1, 0, (volatile double*)nullptr, 0.0);
    HFB_CHECK_LAUNCH(ctx, "lba_errors");
    HFB_CUDA(ctx, cudaMemcpyAsync(depth_positive_out, d.depth_ok, (size_t)d.n_edges, cudaMemcpyDeviceToHost, st));
  }
  if (poses_out) {
    if (dev_solve) HFB_CUDA(ctx, cudaMemcpyAsync(poses_out, d.poses, (size_t)nc * 56, cudaMemcpyDeviceToHost, st));
    else memcpy(poses_out, poses.data(), (size_t)nc * 56);
  }
  if (points_out && d.n_points)
    HFB_CUDA(ctx, cudaMemcpyAsync(points_out, d.points, (size_t)d.n_points * 24, cudaMemcpyDeviceToHost, st));
  if (chi2_out && d.n_edges)
    HFB_CUDA(ctx, cudaMemcpyAsync(chi2_out, chi2_last, (size_t)d.n_edges * 8, cudaMemcpyDeviceToHost, st));
  HFB_CUDA(ctx, cudaStreamSynchronize(st));
  if (stats) {
    stats->iterations = it_done;
    stats->trials = trials;
    stats->initial_chi2 = initial_chi;
    stats->final_chi2 = current_chi;
    stats->lambda = lambda;
    stats->n_opt_cams = no;
    stats->gpu_launches = (int)(ctx->launches - launches0);
  }
  return HFB_OK;
}


// =====================================================================================================================
// Optimizer::PoseOptimization (src/Optimizer.cc:814-1114), monocular branch, as ONE persistent single-CTA kernel: the
// whole schedule -- 4 rounds x optimize(10) of Levenberg-Marquardt on one SE3 vertex with unary
// EdgeSE3ProjectXYZOnlyPose edges (src/OptimizableTypes.cpp:49-64), Huber kernel dropped after the third round,
// inlier re-classification with chi2 > 5.991 after every round -- runs on the device without a host round trip.
// The problem is a few hundred edges: latency, not bandwidth, so everything (6x6 Cholesky, exp map) stays on chip.
#define PO_THREADS 256
#define PO_WARPS (PO_THREADS / 32)
#define PO_NV 28   // 21 upper-triangular H entries + 6 b entries + 1 robust chi2

__device__ __forceinline__ void po_edge(const double* R, const double* t, const double* K, const double* Xw,
                                        const double* obs, double is2, double& e0, double& e1, double& chi2, double& x,
                                        double& y, double& z) {
  x = R[0] * Xw[0] + R[1] * Xw[1] + R[2] * Xw[2] + t[0];
  y = R[3] * Xw[0] + R[4] * Xw[1] + R[5] * Xw[2] + t[1];
  z = R[6] * Xw[0] + R[7] * Xw[1] + R[8] * Xw[2] + t[2];
  e0 = obs[0] - (K[0] * x / z + K[2]);
  e1 = obs[1] - (K[1] * y / z + K[3]);
  chi2 = is2 * (e0 * e0 + e1 * e1);
}

__device__ __forceinline__ double po_shfl_xor(double v, int o) {
  return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), o),
                          __shfl_xor_sync(0xffffffffu, __double2loint(v), o));
}

// Fixed-order block sum of 32 doubles per thread (PO_NV used, the rest zero): a register butterfly inside each warp
// (recursive halving: 31 exchanges instead of 32 x 5 shuffle reductions; lane i ends up with entry i summed over the
// warp), then entry i is summed over the warps in warp order.  out[i] valid for every thread after the final barrier.
__device__ __forceinline__ void po_reduce32(double (&v)[32], double (*sh)[32], double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const double send = up ? v[i] : v[i + o];
      const double keep = up ? v[i + o] : v[i];
      v[i] = keep + po_shfl_xor(send, o);
    }
  }
  sh[warp][lane] = v[0];
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < PO_WARPS; ++w) t += sh[w][threadIdx.x];
    out[threadIdx.x] = t;
  }
  __syncthreads();
}

// fixed-order block sum of one double per thread
__device__ __forceinline__ double po_reduce1(double v, double* sh, double* out) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += po_shfl_xor(v, s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < PO_WARPS; ++w) t += sh[w];
    *out = t;
  }
  __syncthreads();
  return *out;
}

__device__ void po_pose_oplus(const double* pose, const double* u, double* out) {
  const double w0 = u[0], w1 = u[1], w2 = u[2];
  const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
  const double Om[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
  double Om2[9], R[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Om2[3 * i + j] = Om[3 * i] * Om[j] + Om[3 * i + 1] * Om[3 + j] + Om[3 * i + 2] * Om[6 + j];
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta < 0.00001) {
    for (int i = 0; i < 9; ++i) R[i] = I[i] + Om[i] + Om2[i];
    for (int i = 0; i < 9; ++i) V[i] = R[i];
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta);
    const double c = (theta - sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) {
      R[i] = I[i] + a * Om[i] + b * Om2[i];
      V[i] = I[i] + b * Om[i] + c * Om2[i];
    }
  }
  // Eigen Quaternion(Matrix3)
  double qe[4];
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double t = sqrt(tr + 1.0);
    qe[3] = 0.5 * t;
    t = 0.5 / t;
    qe[0] = (R[7] - R[5]) * t;
    qe[1] = (R[2] - R[6]) * t;
    qe[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    qe[i] = 0.5 * t;
    t = 0.5 / t;
    qe[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    qe[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    qe[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  {
    if (qe[3] < 0) for (int i = 0; i < 4; ++i) qe[i] = -qe[i];
    const double nn = sqrt(qe[0] * qe[0] + qe[1] * qe[1] + qe[2] * qe[2] + qe[3] * qe[3]);
    for (int i = 0; i < 4; ++i) qe[i] /= nn;
  }
  const double te[3] = {V[0] * u[3] + V[1] * u[4] + V[2] * u[5], V[3] * u[3] + V[4] * u[4] + V[5] * u[5],
                        V[6] * u[3] + V[7] * u[4] + V[8] * u[5]};
  const double ax = qe[0], ay = qe[1], az = qe[2], aw = qe[3];
  const double bx = pose[0], by = pose[1], bz = pose[2], bw = pose[3];
  double q[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                 aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
  double Re[9];
  quat_to_R(qe, Re);
  const double* t = pose + 4;
  out[4] = te[0] + Re[0] * t[0] + Re[1] * t[1] + Re[2] * t[2];
  out[5] = te[1] + Re[3] * t[0] + Re[4] * t[1] + Re[5] * t[2];
  out[6] = te[2] + Re[6] * t[0] + Re[7] * t[1] + Re[8] * t[2];
  if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
  const double nn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) out[i] = q[i] / nn;
}

// 6 x 6 Cholesky solve on the one thread everything else waits for: every index is a compile-time constant (the factor
// lives in registers; a runtime-indexed A[36] sits in local memory) and each column costs one square root and ONE
// division (the reciprocal of the pivot; the column scaling and both substitutions multiply by it) instead of up to
// seven dependent fp64 divisions of ~20 instructions each.
__device__ __forceinline__ bool po_chol6(const double* Hu, double lambda, const double* b, double* x) {   // Hu: 21 upper-tri entries
  double A[36], inv[6];
  {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = i; j < 6; ++j) {
        A[6 * i + j] = A[6 * j + i] = Hu[k++];
      }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) A[7 * i] += lambda;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[7 * j];
#pragma unroll
    for (int q = 0; q < j; ++q) s -= A[6 * j + q] * A[6 * j + q];
    if (!(s > 0.0) || !isfinite(s)) ok = false;
    inv[j] = 1.0 / sqrt(s);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = A[6 * i + j];
#pragma unroll
      for (int q = 0; q < j; ++q) t -= A[6 * i + q] * A[6 * j + q];
      A[6 * i + j] = t * inv[j];
    }
  }
  if (!ok) return false;
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = b[i];
#pragma unroll
    for (int q = 0; q < i; ++q) t -= A[6 * i + q] * y[q];
    y[i] = t * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double t = y[i];
#pragma unroll
    for (int q = i + 1; q < 6; ++q) t -= A[6 * q + i] * y[q];
    y[i] = t * inv[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = y[i];
  return true;
}

// io: [n*3 Xw | n*2 obs | n invs2 | 4 K | 7 pose0] doubles in, [7 pose | 3 stats (as doubles)] + n outlier bytes out.  The
// edges are staged in shared memory once (they are re-read ~50 times: every LM trial is one pass over them).
__global__ void __launch_bounds__(PO_THREADS) pose_opt_kernel(int n, const double* __restrict__ in, double delta,
                                                              int edges_in_smem, double* __restrict__ cached_g,
                                                              unsigned char* __restrict__ outlier_g, double* __restrict__ out,
                                                              int dbg) {
  long long t_lin = 0, t_red = 0, t_solve = 0, t_trial = 0, t_mark = 0, t_all = clock64();   // HFB_PO_DBG: thread 0's cycles
  extern __shared__ double s_dyn[];   // edges_in_smem: [n*3 | n*2 | n | n cached] doubles + n outlier bytes
  __shared__ double sh[PO_WARPS][32], s_v[32], sh1[PO_WARPS], s_one;
  __shared__ double s_pose[7], s_trial[7], s_x[6];
  __shared__ double s_lambda, s_ni, s_cur, s_ini, s_rho;
  __shared__ int s_ok, s_cont, s_stop, s_nbad_lm, s_qmax, s_trials, s_iters;
  const int tid = threadIdx.x;
  const double* gX = in;
  const double* gO = in + (size_t)3 * n;
  const double* gS = in + (size_t)5 * n;
  const double* gK = in + (size_t)6 * n;
  __shared__ double s_pose0[7];   // the inputs may sit in page-locked host memory: read them once
  if (threadIdx.x < 7) s_pose0[threadIdx.x] = gK[4 + threadIdx.x];
  const double* pose0 = s_pose0;
  const double K[4] = {gK[0], gK[1], gK[2], gK[3]};
  const double dsqr = delta * delta;
  const double *Xw = gX, *obs = gO, *invs2 = gS;
  double* cached = cached_g;
  unsigned char* outlier = outlier_g;
  if (edges_in_smem) {
    for (int i = tid; i < 6 * n; i += PO_THREADS) s_dyn[i] = in[i];
    Xw = s_dyn;
    obs = s_dyn + (size_t)3 * n;
    invs2 = s_dyn + (size_t)5 * n;
    cached = s_dyn + (size_t)6 * n;
    outlier = reinterpret_cast<unsigned char*>(s_dyn + (size_t)7 * n);
  }
  if (tid == 0) {
    s_trials = 0;
    s_iters = 0;
  }
  for (int e = tid; e < n; e += PO_THREADS) outlier[e] = 0;
  __syncthreads();
  int n_bad = 0;
  for (int rnd = 0; rnd < 4; ++rnd) {
    const bool robust = rnd < 3;
    if (tid < 7) s_pose[tid] = pose0[tid];
    if (tid == 0) {
      s_lambda = 0;
      s_ni = 2;
      s_nbad_lm = 0;
      s_stop = 0;
    }
    __syncthreads();
    for (int it = 0; it < 10; ++it) {
      // ---- computeActiveErrors + buildSystem at s_pose
      t_mark = clock64();
      double R[9];
      quat_to_R(s_pose, R);
      const double t3[3] = {s_pose[4], s_pose[5], s_pose[6]};
      double v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0;
      for (int e = tid; e < n; e += PO_THREADS) {
        if (outlier[e]) continue;
        v[28] += 1.0;                       // active edges
        double e0, e1, chi2, x, y, z;
        po_edge(R, t3, K, Xw + 3 * e, obs + 2 * e, invs2[e], e0, e1, chi2, x, y, z);
        cached[e] = chi2;
        const bool inl = !robust || chi2 <= dsqr;
        const double sq = sqrt(fmax(chi2, 1e-300));
        v[27] += inl ? chi2 : 2 * sq * delta - dsqr;
        const double wo = (inl ? 1.0 : delta / sq) * invs2[e];
        const double iz = 1.0 / z, iz2 = 1.0 / (z * z);
        const double a00 = -K[0] * iz, a02 = K[0] * x * iz2, a11 = -K[1] * iz, a12 = K[1] * y * iz2;
        const double J0[6] = {a02 * y, a00 * z - a02 * x, -a00 * y, a00, 0, a02};
        const double J1[6] = {-a11 * z + a12 * y, -a12 * x, a11 * x, 0, a11, a12};
        const double r0 = -wo * e0, r1 = -wo * e1;
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
          for (int j = i; j < 6; ++j) v[k++] += wo * (J0[i] * J0[j] + J1[i] * J1[j]);
          v[21 + i] += J0[i] * r0 + J1[i] * r1;
        }
      }
      t_lin += clock64() - t_mark;
      t_mark = clock64();
      po_reduce32(v, sh, s_v);
      t_red += clock64() - t_mark;
      if (tid == 0) {
        s_cur = s_v[27];
        s_ini = s_v[27];
        if (it == 0) {
          double md = 0;
          int k = 0;
          for (int i = 0; i < 6; ++i) {
            md = fmax(md, fabs(s_v[k]));
            k += 6 - i;
          }
          s_lambda = 1e-5 * md;
          s_ni = 2;
          s_nbad_lm = 0;
        }
        s_qmax = 0;
        s_rho = 0;
      }
      __syncthreads();
      if (s_v[28] == 0.0) break;
      // ---- Levenberg trials
      while (true) {
        t_mark = clock64();
        if (tid == 0) {
          double x[6];
          s_ok = po_chol6(s_v, s_lambda, s_v + 21, x) ? 1 : 0;
          if (!s_ok)
            for (int i = 0; i < 6; ++i) x[i] = 0;
          for (int i = 0; i < 6; ++i) s_x[i] = x[i];
          po_pose_oplus(s_pose, x, s_trial);
        }
        __syncthreads();
        t_solve += clock64() - t_mark;
        t_mark = clock64();
        double Rt[9];
        quat_to_R(s_trial, Rt);
        const double tt[3] = {s_trial[4], s_trial[5], s_trial[6]};
        double c = 0;
        for (int e = tid; e < n; e += PO_THREADS) {
          if (outlier[e]) continue;
          double e0, e1, chi2, x, y, z;
          po_edge(Rt, tt, K, Xw + 3 * e, obs + 2 * e, invs2[e], e0, e1, chi2, x, y, z);
          cached[e] = chi2;
          const bool inl = !robust || chi2 <= dsqr;
          c += inl ? chi2 : 2 * sqrt(fmax(chi2, 1e-300)) * delta - dsqr;
        }
        c = po_reduce1(c, sh1, &s_one);
        t_trial += clock64() - t_mark;
        if (tid == 0) {
          const double temp = s_ok ? c : 1.7976931348623157e308;
          double scale = 1e-3;
          for (int i = 0; i < 6; ++i) scale += s_x[i] * (s_lambda * s_x[i] + s_v[21 + i]);
          const double rho = (s_cur - temp) / scale;
          s_rho = rho;
          ++s_trials;
          if (rho > 0 && isfinite(temp)) {
            double alpha = 1.0 - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
            alpha = fmin(alpha, 2.0 / 3.0);
            s_lambda *= fmax(1.0 / 3.0, alpha);
            s_ni = 2;
            s_cur = temp;
            for (int i = 0; i < 7; ++i) s_pose[i] = s_trial[i];
          } else {
            s_lambda *= s_ni;
            s_ni *= 2;
          }
          ++s_qmax;
          s_cont = (rho < 0 && s_qmax < 10) ? 1 : 0;
        }
        __syncthreads();
        if (!s_cont) break;
      }
      if (tid == 0) {
        ++s_iters;
        if (s_qmax == 10 || s_rho == 0) s_stop = 1;
        else {
          if ((s_ini - s_cur) * 1e3 < s_ini) ++s_nbad_lm;
          else s_nbad_lm = 0;
          if (s_nbad_lm >= 3) s_stop = 1;
        }
      }
      __syncthreads();
      if (s_stop) break;
    }
    // ---- classification: outliers are re-evaluated at the round's final pose, inliers keep the cached chi2
    double R[9];
    quat_to_R(s_pose, R);
    const double t3[3] = {s_pose[4], s_pose[5], s_pose[6]};
    double bad = 0;
    for (int e = tid; e < n; e += PO_THREADS) {
      double chi2 = cached[e];
      if (outlier[e]) {
        double e0, e1, x, y, z;
        po_edge(R, t3, K, Xw + 3 * e, obs + 2 * e, invs2[e], e0, e1, chi2, x, y, z);
      }
      const bool o = (float)chi2 > 5.991f;
      outlier[e] = o ? 1 : 0;
      bad += o ? 1.0 : 0.0;
    }
    n_bad = (int)po_reduce1(bad, sh1, &s_one);
    if (n < 10) break;
  }
  if (tid < 7) out[tid] = s_pose[tid];
  if (tid == 0) {
    out[7] = n - n_bad;
    out[8] = s_trials;
    out[9] = s_iters;
  }
  if (edges_in_smem)
    for (int e = tid; e < n; e += PO_THREADS) outlier_g[e] = outlier[e];
  if (dbg && tid == 0)
    printf("pose_opt cycles: total %lld linearize %lld reduce %lld solve+oplus %lld trial %lld (trials %d iters %d)\n",
           clock64() - t_all, t_lin, t_red, t_solve, t_trial, s_trials, s_iters);
}

extern "C" int hfb_pose_optimize(hfb_ctx* ctx, const float* K, const double* pose_in, int32_t n, const double* Xw,
                                 const double* obs, const double* inv_sigma2, double* pose_out, uint8_t* outlier_out,
                                 int32_t* n_inliers, int32_t* n_trials) {
  HFB_ENTER(ctx);
  HFB_REQUIRE(ctx, K && pose_in && pose_out && n >= 0 && (n == 0 || (Xw && obs && inv_sigma2)), "bad argument");
  if (n == 0) {
    memcpy(pose_out, pose_in, 56);
    if (n_inliers) *n_inliers = 0;
    if (n_trials) *n_trials = 0;
    return HFB_OK;
  }
  // one page-locked block in, one out: [Xw | obs | invs2 | K | pose0] -> [pose | stats] + outlier flags
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t in_d = (size_t)6 * n + 4 + 7, in_bytes = al(in_d * 8), out_bytes = al(10 * 8 + (size_t)n);
  const size_t oIn = 0, oOut = in_bytes, oC = oOut + out_bytes, total = oC + al((size_t)n * 8);
  HFB_TRY(ctx->ensure_scratch(total));
  HFB_TRY(ctx->ensure_stage(in_bytes + out_bytes));
  uint8_t* a = reinterpret_cast<uint8_t*>(ctx->d_scratch);
  double* hin = reinterpret_cast<double*>(ctx->h_stage);
  uint8_t* hout = reinterpret_cast<uint8_t*>(ctx->h_stage) + in_bytes;
  memcpy(hin, Xw, (size_t)n * 24);
  memcpy(hin + (size_t)3 * n, obs, (size_t)n * 16);
  memcpy(hin + (size_t)5 * n, inv_sigma2, (size_t)n * 8);
  for (int i = 0; i < 4; ++i) hin[(size_t)6 * n + i] = (double)K[i];
  memcpy(hin + (size_t)6 * n + 4, pose_in, 56);
  cudaStream_t st = ctx->stream;
  const size_t smem = (size_t)7 * n * 8 + (size_t)n + 16;
  const int in_smem = smem <= 200 * 1024;
  static SmemOptIn optin;
  if (in_smem && smem > 48 * 1024) HFB_CUDA(ctx, optin.ensure(pose_opt_kernel, ctx->device, 200 * 1024));
  const int dbg = getenv("HFB_PO_DBG") ? 1 : 0;
  if (in_smem) {
    // The edges are read exactly once (into shared memory) and the result is 80 + n bytes: the kernel reads the
    // page-locked block and writes the result block in place (unified addressing), which saves the two copy operations
    // -- a third of the wall time of a call on a problem this small.
    pose_opt_kernel<<<1, PO_THREADS, smem, st>>>(n, hin, sqrt(5.991), 1, (double*)(a + oC), hout + 80,
                                                 reinterpret_cast<double*>(hout), dbg);
    HFB_CHECK_LAUNCH(ctx, "pose_opt");
  } else {
    HFB_CUDA(ctx, cudaMemcpyAsync(a + oIn, hin, in_d * 8, cudaMemcpyHostToDevice, st));
    pose_opt_kernel<<<1, PO_THREADS, 0, st>>>(n, (const double*)(a + oIn), sqrt(5.991), 0, (double*)(a + oC), a + oOut + 80,
                                              (double*)(a + oOut), dbg);
    HFB_CHECK_LAUNCH(ctx, "pose_opt");
    HFB_CUDA(ctx, cudaMemcpyAsync(hout, a + oOut, 80 + (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  HFB_CUDA(ctx, cudaStreamSynchronize(st));
  const double* ho = reinterpret_cast<const double*>(hout);
  memcpy(pose_out, ho, 56);
  if (outlier_out) memcpy(outlier_out, hout + 80, (size_t)n);
  if (n_inliers) *n_inliers = (int)ho[7];
  if (n_trials) *n_trials = (int)ho[8];
  return HFB_OK;
}
