/* Oracle (TEST INFRASTRUCTURE ONLY): plain-C double-precision restatement of the numeric core of
 * Optimizer::LocalBundleAdjustment, edge by edge and single-threaded like the reference (g2o is built without OpenMP,
 * Thirdparty/g2o/CMakeLists.txt:48).  It mirrors oracle/lba_ref.py function for function (same update order, same LM
 * control) and is checked against it in tests/test_oracle_pins.py; bench.py times it as the CPU baseline of the
 * local-BA part of the metric.  Nothing under hfnet_slam_b200/ links or loads this file.
 *
 * Reference lines (relative to the reference repository):
 *   edge error / depth test        include/OptimizableTypes.h:99-110, src/CameraModels/Pinhole.cpp:35-49
 *   Jacobians                      src/OptimizableTypes.cpp:139-159, src/CameraModels/Pinhole.cpp:71-81
 *   Huber + quadratic form         Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91, base_edge.h:96-102,
 *                                  base_binary_edge.hpp:55-121
 *   Schur complement / back-subst  Thirdparty/g2o/g2o/core/block_solver.hpp:354-486, setLambda :564-589
 *   Levenberg-Marquardt            Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-185
 *   pose / point updates           Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:73-76, se3quat.h:223-257, types_sba.h:52-56
 * The reduced camera system is solved by a dense Cholesky (g2o: sparse LDLT of the same matrix,
 * Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int n_cam, n_pt, n_edge;
  const double* poses0;
  const uint8_t* fixed;
  const double* points0;
  const int32_t *cam, *pt;
  const double *obs, *is2;
  double K[4], delta;
  int n_opt;
  int* slot;     /* [n_cam] */
  int* opt_cam;  /* [n_opt] */
  /* system */
  double *Hpp, *bp, *Hll, *bl, *Hpl, *chi2, *rho0, *err, *Xc;
} Lba;

static void quat_to_rot(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
               tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

static void rot_to_quat(const double* R, double* q) { /* Eigen's Quaternion(Matrix3) */
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
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}

static void normalize_rotation(double* q) {
  if (q[3] < 0)
    for (int i = 0; i < 4; ++i) q[i] = -q[i];
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

static void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

/* se3quat.h:223-257: update = [omega, upsilon] */
static void se3_exp(const double* u, double* q, double* t) {
  const double* om = u;
  const double* up = u + 3;
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double Om2[9], R[9], V[9];
  mat3_mul(Om, Om, Om2);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta < 0.00001) {
    for (int i = 0; i < 9; ++i) R[i] = I[i] + Om[i] + Om2[i];
    memcpy(V, R, sizeof(R));
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta),
                 c = (theta - sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) {
      R[i] = I[i] + a * Om[i] + b * Om2[i];
      V[i] = I[i] + b * Om[i] + c * Om2[i];
    }
  }
  rot_to_quat(R, q);
  normalize_rotation(q);
  for (int i = 0; i < 3; ++i) t[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
}

static void pose_oplus(const double* pose, const double* u, double* out) { /* exp(u) * pose */
  double qe[4], te[3], Re[9];
  se3_exp(u, qe, te);
  const double ax = qe[0], ay = qe[1], az = qe[2], aw = qe[3], bx = pose[0], by = pose[1], bz = pose[2], bw = pose[3];
  out[0] = aw * bx + ax * bw + ay * bz - az * by;
  out[1] = aw * by + ay * bw + az * bx - ax * bz;
  out[2] = aw * bz + az * bw + ax * by - ay * bx;
  out[3] = aw * bw - ax * bx - ay * by - az * bz;
  normalize_rotation(out);
  quat_to_rot(qe, Re);
  for (int i = 0; i < 3; ++i) out[4 + i] = te[i] + Re[i * 3] * pose[4] + Re[i * 3 + 1] * pose[5] + Re[i * 3 + 2] * pose[6];
}

/* computeActiveErrors: err, chi2 = e^T Omega e, robustified rho(chi2), camera-frame point */
static double edge_errors(const Lba* L, const double* poses, const double* points, double* err, double* chi2,
                          double* rho0, double* Xc, const double* Rs) {
  const double fx = L->K[0], fy = L->K[1], cx = L->K[2], cy = L->K[3], dsqr = L->delta * L->delta;
  double sum = 0.0;
  for (int e = 0; e < L->n_edge; ++e) {
    const int c = L->cam[e], p = L->pt[e];
    const double* R = Rs + 9 * c;
    const double* X = points + 3 * p;
    const double* t = poses + 7 * c + 4;
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double ex = L->obs[2 * e] - (fx * x / z + cx), ey = L->obs[2 * e + 1] - (fy * y / z + cy);
    const double c2 = L->is2[e] * (ex * ex + ey * ey);
    if (err) {
      err[2 * e] = ex;
      err[2 * e + 1] = ey;
    }
    if (Xc) {
      Xc[3 * e] = x;
      Xc[3 * e + 1] = y;
      Xc[3 * e + 2] = z;
    }
    chi2[e] = c2;
    const double r = c2 <= dsqr ? c2 : 2 * sqrt(c2) * L->delta - dsqr;
    if (rho0) rho0[e] = r;
    sum += r;
  }
  return sum;
}

static void all_rotations(const Lba* L, const double* poses, double* Rs) {
  for (int c = 0; c < L->n_cam; ++c) quat_to_rot(poses + 7 * c, Rs + 9 * c);
}

/* computeActiveErrors + buildSystem (block_solver.hpp:502-560), robust branch of constructQuadraticForm */
static double build_system(Lba* L, const double* poses, const double* points, double* Rs) {
  all_rotations(L, poses, Rs);
  const double cur = edge_errors(L, poses, points, L->err, L->chi2, L->rho0, L->Xc, Rs);
  memset(L->Hpp, 0, sizeof(double) * 36 * (size_t)L->n_opt);
  memset(L->bp, 0, sizeof(double) * 6 * (size_t)L->n_opt);
  memset(L->Hll, 0, sizeof(double) * 9 * (size_t)L->n_pt);
  memset(L->bl, 0, sizeof(double) * 3 * (size_t)L->n_pt);
  const double fx = L->K[0], fy = L->K[1], dsqr = L->delta * L->delta;
  for (int e = 0; e < L->n_edge; ++e) {
    const int c = L->cam[e], p = L->pt[e], s = L->slot[c];
    const double x = L->Xc[3 * e], y = L->Xc[3 * e + 1], z = L->Xc[3 * e + 2];
    /* -Jproj (2x3) */
    const double P[6] = {-fx / z, 0, fx * x / (z * z), 0, -fy / z, fy * y / (z * z)};
    const double* R = Rs + 9 * c;
    double Jx[6], Jc[12];
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 3; ++j) Jx[i * 3 + j] = P[i * 3] * R[j] + P[i * 3 + 1] * R[3 + j] + P[i * 3 + 2] * R[6 + j];
    const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 6; ++j) Jc[i * 6 + j] = P[i * 3] * D[j] + P[i * 3 + 1] * D[6 + j] + P[i * 3 + 2] * D[12 + j];
    const double w = L->chi2[e] <= dsqr ? 1.0 : L->delta / sqrt(L->chi2[e]);
    const double wo = w * L->is2[e];
    const double r0 = -(wo * L->err[2 * e]), r1 = -(wo * L->err[2 * e + 1]);
    double* Hll = L->Hll + 9 * (size_t)p;
    double* bl = L->bl + 3 * (size_t)p;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Hll[i * 3 + j] += wo * (Jx[i] * Jx[j] + Jx[3 + i] * Jx[3 + j]);
      bl[i] += Jx[i] * r0 + Jx[3 + i] * r1;
    }
    double* Hpl = L->Hpl + 18 * (size_t)e;
    if (s >= 0) {
      double* Hpp = L->Hpp + 36 * (size_t)s;
      double* bp = L->bp + 6 * (size_t)s;
      for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) Hpp[i * 6 + j] += wo * (Jc[i] * Jc[j] + Jc[6 + i] * Jc[6 + j]);
        bp[i] += Jc[i] * r0 + Jc[6 + i] * r1;
        for (int j = 0; j < 3; ++j) Hpl[i * 3 + j] = wo * (Jc[i] * Jx[j] + Jc[6 + i] * Jx[3 + j]);
      }
    } else {
      memset(Hpl, 0, sizeof(double) * 18);
    }
  }
  return cur;
}

static void inv3(const double* A, double* B) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02, id = 1.0 / det;
  B[0] = c00 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  B[3] = c01 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  B[6] = c02 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

/* BlockSolver::solve marginalisation (block_solver.hpp:380-439) with lambda on every diagonal (setLambda) */
static void schur(const Lba* L, double lam, double* Hs, double* bs, double* Dinv) {
  const int n = 6 * L->n_opt;
  memset(Hs, 0, sizeof(double) * (size_t)n * n);
  for (int s = 0; s < L->n_opt; ++s)
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j) Hs[(size_t)(6 * s + i) * n + 6 * s + j] = L->Hpp[36 * (size_t)s + i * 6 + j] + (i == j ? lam : 0.0);
      bs[6 * s + i] = L->bp[6 * (size_t)s + i];
    }
  int e0 = 0;
  while (e0 < L->n_edge) {
    const int p = L->pt[e0];
    int e1 = e0;
    while (e1 < L->n_edge && L->pt[e1] == p) ++e1;
    double Dm[9];
    memcpy(Dm, L->Hll + 9 * (size_t)p, sizeof(Dm));
    Dm[0] += lam; Dm[4] += lam; Dm[8] += lam;
    double* Di = Dinv + 9 * (size_t)p;
    inv3(Dm, Di);
    const double* bl = L->bl + 3 * (size_t)p;
    const double db[3] = {Di[0] * bl[0] + Di[1] * bl[1] + Di[2] * bl[2], Di[3] * bl[0] + Di[4] * bl[1] + Di[5] * bl[2],
                          Di[6] * bl[0] + Di[7] * bl[1] + Di[8] * bl[2]};
    for (int a = e0; a < e1; ++a) {
      const int ia = L->slot[L->cam[a]];
      if (ia < 0) continue;
      const double* Ba = L->Hpl + 18 * (size_t)a;
      double BD[18];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j) BD[i * 3 + j] = Ba[i * 3] * Di[j] + Ba[i * 3 + 1] * Di[3 + j] + Ba[i * 3 + 2] * Di[6 + j];
      for (int i = 0; i < 6; ++i) bs[6 * ia + i] -= Ba[i * 3] * db[0] + Ba[i * 3 + 1] * db[1] + Ba[i * 3 + 2] * db[2];
      for (int b = e0; b < e1; ++b) {
        const int ib = L->slot[L->cam[b]];
        if (ib < 0) continue;
        const double* Bb = L->Hpl + 18 * (size_t)b;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j)
            Hs[(size_t)(6 * ia + i) * n + 6 * ib + j] -= BD[i * 3] * Bb[j * 3] + BD[i * 3 + 1] * Bb[j * 3 + 1] + BD[i * 3 + 2] * Bb[j * 3 + 2];
      }
    }
    e0 = e1;
  }
}

/* dense Cholesky A = L L^T (lower, in place), then solve; returns 0 when A is not positive definite */
static int chol_solve(double* A, int n, const double* b, double* x) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !isfinite(d)) return 0;
    d = sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * x[k];
    x[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * x[k];
    x[i] = s / A[(size_t)i * n + i];
  }
  return 1;
}

/* block_solver.hpp:461-481: xl = Dinv (bl - Hpl^T xp) */
static void back_substitute(const Lba* L, const double* Dinv, const double* xp, double* xl, double* cl) {
  memcpy(cl, L->bl, sizeof(double) * 3 * (size_t)L->n_pt);
  for (int e = 0; e < L->n_edge; ++e) {
    const int s = L->slot[L->cam[e]];
    if (s < 0) continue;
    const double* B = L->Hpl + 18 * (size_t)e;
    const double* x = xp + 6 * s;
    double* c = cl + 3 * (size_t)L->pt[e];
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int i = 0; i < 6; ++i) acc += B[i * 3 + j] * x[i];
      c[j] -= acc;
    }
  }
  for (int p = 0; p < L->n_pt; ++p) {
    const double* Di = Dinv + 9 * (size_t)p;
    const double* c = cl + 3 * (size_t)p;
    for (int i = 0; i < 3; ++i) xl[3 * p + i] = Di[i * 3] * c[0] + Di[i * 3 + 1] * c[1] + Di[i * 3 + 2] * c[2];
  }
}

/* g2o SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg + BlockSolver_6_3.
 * stats: [0] iterations, [1] trials, [2] initial robust chi2, [3] final robust chi2, [4] final lambda */
int ref_lba_optimize(int n_cam, int n_pt, int n_edge, const double* poses0, const uint8_t* fixed, const double* points0,
                     const int32_t* cam, const int32_t* pt, const double* obs, const double* is2, const float* K,
                     double huber_delta, int iterations, double user_lambda_init, double* poses_out, double* points_out,
                     double* chi2_out, uint8_t* depth_positive_out, double* stats) {
  Lba L;
  memset(&L, 0, sizeof(L));
  L.n_cam = n_cam; L.n_pt = n_pt; L.n_edge = n_edge;
  L.poses0 = poses0; L.fixed = fixed; L.points0 = points0; L.cam = cam; L.pt = pt; L.obs = obs; L.is2 = is2;
  for (int i = 0; i < 4; ++i) L.K[i] = (double)K[i];
  L.delta = huber_delta;
  L.slot = (int*)malloc(sizeof(int) * (size_t)(n_cam + 1));
  L.opt_cam = (int*)malloc(sizeof(int) * (size_t)(n_cam + 1));
  for (int c = 0; c < n_cam; ++c) {
    L.slot[c] = fixed[c] ? -1 : L.n_opt;
    if (!fixed[c]) L.opt_cam[L.n_opt++] = c;
  }
  /* points without edges must still get a finite Dinv: every point is visited by schur() only through its edges, so
   * their Dinv stays at the initial zero and their update is zero (g2o: such vertices are not active). */
  const int n = 6 * L.n_opt;
  const size_t ne = (size_t)(n_edge > 0 ? n_edge : 1), np_ = (size_t)(n_pt > 0 ? n_pt : 1), no = (size_t)(L.n_opt > 0 ? L.n_opt : 1);
  L.Hpp = (double*)calloc(36 * no, sizeof(double)); L.bp = (double*)calloc(6 * no, sizeof(double));
  L.Hll = (double*)calloc(9 * np_, sizeof(double)); L.bl = (double*)calloc(3 * np_, sizeof(double));
  L.Hpl = (double*)calloc(18 * ne, sizeof(double)); L.chi2 = (double*)calloc(ne, sizeof(double));
  L.rho0 = (double*)calloc(ne, sizeof(double)); L.err = (double*)calloc(2 * ne, sizeof(double));
  L.Xc = (double*)calloc(3 * ne, sizeof(double));
  double* poses = (double*)malloc(sizeof(double) * 7 * (size_t)(n_cam + 1));
  double* points = (double*)malloc(sizeof(double) * 3 * np_);
  double* new_poses = (double*)malloc(sizeof(double) * 7 * (size_t)(n_cam + 1));
  double* new_points = (double*)malloc(sizeof(double) * 3 * np_);
  double* Rs = (double*)malloc(sizeof(double) * 9 * (size_t)(n_cam + 1));
  double* Hs = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (size_t)(n > 0 ? n : 1));
  double* bs = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double* xp = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double* Dinv = (double*)calloc(9 * np_, sizeof(double));
  double* xl = (double*)calloc(3 * np_, sizeof(double));
  double* cl = (double*)calloc(3 * np_, sizeof(double));
  double* chi2_t = (double*)calloc(ne, sizeof(double));
  double* last_chi2 = (double*)calloc(ne, sizeof(double));
  memcpy(poses, poses0, sizeof(double) * 7 * (size_t)n_cam);
  memcpy(points, points0, sizeof(double) * 3 * (size_t)n_pt);

  const double tau = 1e-5, good_lo = 1.0 / 3.0, good_hi = 2.0 / 3.0;
  const int max_trials = 10;
  double lam = 0.0, ni = 2.0, initial_chi = 0.0, final_chi = 0.0;
  int n_bad = 0, it_done = 0, trials = 0;
  all_rotations(&L, poses, Rs);
  edge_errors(&L, poses, points, NULL, last_chi2, NULL, NULL, Rs);
  for (int it = 0; it < iterations; ++it) {
    double current_chi = build_system(&L, poses, points, Rs);
    memcpy(last_chi2, L.chi2, sizeof(double) * (size_t)n_edge);
    const double ini_chi = current_chi;
    if (it == 0) {
      initial_chi = current_chi;
      if (user_lambda_init > 0) {
        lam = user_lambda_init;
      } else {
        double max_diag = 0.0;
        for (int s = 0; s < L.n_opt; ++s)
          for (int i = 0; i < 6; ++i) max_diag = fmax(max_diag, fabs(L.Hpp[36 * (size_t)s + i * 7]));
        for (int p = 0; p < n_pt; ++p)
          for (int i = 0; i < 3; ++i) max_diag = fmax(max_diag, fabs(L.Hll[9 * (size_t)p + i * 4]));
        lam = tau * max_diag;
      }
      ni = 2.0;
      n_bad = 0;
    }
    double rho = 0.0;
    int qmax = 0;
    do {
      schur(&L, lam, Hs, bs, Dinv);
      const int ok2 = n > 0 ? chol_solve(Hs, n, bs, xp) : 1;
      if (!ok2) memset(xp, 0, sizeof(double) * (size_t)n);
      back_substitute(&L, Dinv, xp, xl, cl);
      memcpy(new_poses, poses, sizeof(double) * 7 * (size_t)n_cam);
      for (int s = 0; s < L.n_opt; ++s) pose_oplus(poses + 7 * L.opt_cam[s], xp + 6 * s, new_poses + 7 * L.opt_cam[s]);
      for (int i = 0; i < 3 * n_pt; ++i) new_points[i] = points[i] + xl[i];
      all_rotations(&L, new_poses, Rs);
      double temp_chi = edge_errors(&L, new_poses, new_points, NULL, chi2_t, NULL, NULL, Rs);
      memcpy(last_chi2, chi2_t, sizeof(double) * (size_t)n_edge);
      if (!ok2) temp_chi = DBL_MAX;
      double scale = 0.0;
      for (int i = 0; i < n; ++i) scale += xp[i] * (lam * xp[i] + L.bp[i]);
      for (int i = 0; i < 3 * n_pt; ++i) scale += xl[i] * (lam * xl[i] + L.bl[i]);
      scale += 1e-3;
      rho = (current_chi - temp_chi) / scale;
      ++trials;
      if (rho > 0 && isfinite(temp_chi)) {
        double alpha = 1.0 - pow(2 * rho - 1, 3);
        alpha = fmin(alpha, good_hi);
        lam *= fmax(good_lo, alpha);
        ni = 2.0;
        current_chi = temp_chi;
        memcpy(poses, new_poses, sizeof(double) * 7 * (size_t)n_cam);
        memcpy(points, new_points, sizeof(double) * 3 * (size_t)n_pt);
      } else {
        lam *= ni;
        ni *= 2;
      }
      ++qmax;
    } while (rho < 0 && qmax < max_trials);
    final_chi = current_chi;
    ++it_done;
    if (qmax == max_trials || rho == 0) break;
    if ((ini_chi - current_chi) * 1e3 < ini_chi) ++n_bad; else n_bad = 0;
    if (n_bad >= 3) break;
  }
  all_rotations(&L, poses, Rs);
  for (int e = 0; e < n_edge; ++e) {
    const double* R = Rs + 9 * cam[e];
    const double* X = points + 3 * pt[e];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + poses[7 * cam[e] + 6];
    depth_positive_out[e] = z > 0.0;
  }
  memcpy(poses_out, poses, sizeof(double) * 7 * (size_t)n_cam);
  memcpy(points_out, points, sizeof(double) * 3 * (size_t)n_pt);
  memcpy(chi2_out, last_chi2, sizeof(double) * (size_t)n_edge);
  stats[0] = it_done; stats[1] = trials; stats[2] = initial_chi; stats[3] = final_chi; stats[4] = lam;
  free(L.slot); free(L.opt_cam); free(L.Hpp); free(L.bp); free(L.Hll); free(L.bl); free(L.Hpl); free(L.chi2);
  free(L.rho0); free(L.err); free(L.Xc); free(poses); free(points); free(new_poses); free(new_points); free(Rs);
  free(Hs); free(bs); free(xp); free(Dinv); free(xl); free(cl); free(chi2_t); free(last_chi2);
  return 0;
}

/* ===================================================================================================================
 * Optimizer::PoseOptimization (src/Optimizer.cc:814-1114), monocular branch, mirroring oracle/lba_ref.py:
 * pose_optimization: one VertexSE3Expmap, unary EdgeSE3ProjectXYZOnlyPose edges (include/OptimizableTypes.h:30-56,
 * src/OptimizableTypes.cpp:49-64); four rounds of optimize(10), each restarted from the frame's pose, inliers only,
 * re-classification with chi2 > 5.991 (float compare) after every round, Huber dropped after the third. */
static double pose_edges(const double* K, const double* pose, int n, const double* Xw, const double* obs,
                         const double* is2, const uint8_t* skip, double* err, double* chi2, double* Xc, int robust,
                         double delta) {
  double R[9];
  quat_to_rot(pose, R);
  const double dsqr = delta * delta;
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    if (skip && skip[i]) continue;
    const double* X = Xw + 3 * i;
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + pose[4];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + pose[5];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + pose[6];
    const double ex = obs[2 * i] - (K[0] * x / z + K[2]), ey = obs[2 * i + 1] - (K[1] * y / z + K[3]);
    const double c2 = is2[i] * (ex * ex + ey * ey);
    if (err) { err[2 * i] = ex; err[2 * i + 1] = ey; }
    if (Xc) { Xc[3 * i] = x; Xc[3 * i + 1] = y; Xc[3 * i + 2] = z; }
    chi2[i] = c2;
    sum += (!robust || c2 <= dsqr) ? c2 : 2 * sqrt(c2) * delta - dsqr;
  }
  return sum;
}

int ref_pose_optimize(const float* Kf, const double* pose0, int n, const double* Xw, const double* obs, const double* is2,
                      double huber_delta, double* pose_out, uint8_t* outlier_out, int* n_inliers, int* n_trials) {
  const double K[4] = {Kf[0], Kf[1], Kf[2], Kf[3]};
  const double tau = 1e-5, good_lo = 1.0 / 3.0, good_hi = 2.0 / 3.0, dsqr = huber_delta * huber_delta;
  const int max_trials = 10;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  uint8_t* outlier = (uint8_t*)calloc(nn, 1);
  double* err = (double*)calloc(2 * nn, sizeof(double));
  double* chi2 = (double*)calloc(nn, sizeof(double));
  double* chi2_t = (double*)calloc(nn, sizeof(double));
  double* cached = (double*)calloc(nn, sizeof(double));
  double* Xc = (double*)calloc(3 * nn, sizeof(double));
  double pose[7], new_pose[7];
  int trials = 0, n_bad = 0;
  memcpy(pose, pose0, sizeof(pose));
  for (int rnd = 0; rnd < 4; ++rnd) {
    const int robust = rnd < 3;
    memcpy(pose, pose0, sizeof(pose));
    memset(cached, 0, sizeof(double) * nn);
    int n_act = 0;
    for (int i = 0; i < n; ++i) n_act += !outlier[i];
    double lam = 0.0, ni = 2.0;
    int nbad_lm = 0;
    for (int it = 0; it < 10 && n_act > 0; ++it) {
      double current_chi = pose_edges(K, pose, n, Xw, obs, is2, outlier, err, chi2, Xc, robust, huber_delta);
      const double ini_chi = current_chi;
      double H[36] = {0}, b[6] = {0};
      for (int i = 0; i < n; ++i) {
        if (outlier[i]) continue;
        cached[i] = chi2[i];
        const double x = Xc[3 * i], y = Xc[3 * i + 1], z = Xc[3 * i + 2];
        const double P[6] = {-K[0] / z, 0, K[0] * x / (z * z), 0, -K[1] / z, K[1] * y / (z * z)};
        const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
        double J[12];
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 6; ++c) J[r * 6 + c] = P[r * 3] * D[c] + P[r * 3 + 1] * D[6 + c] + P[r * 3 + 2] * D[12 + c];
        const double w = (!robust || chi2[i] <= dsqr) ? 1.0 : huber_delta / sqrt(chi2[i]);
        const double wo = w * is2[i];
        const double r0 = -(wo * err[2 * i]), r1 = -(wo * err[2 * i + 1]);
        for (int r = 0; r < 6; ++r) {
          for (int c = 0; c < 6; ++c) H[r * 6 + c] += wo * (J[r] * J[c] + J[6 + r] * J[6 + c]);
          b[r] += J[r] * r0 + J[6 + r] * r1;
        }
      }
      if (it == 0) {
        double md = 0.0;
        for (int r = 0; r < 6; ++r) md = fmax(md, fabs(H[r * 7]));
        lam = tau * md;
        ni = 2.0;
        nbad_lm = 0;
      }
      double rho = 0.0;
      int qmax = 0;
      do {
        double A[36], x6[6];
        memcpy(A, H, sizeof(A));
        for (int r = 0; r < 6; ++r) A[r * 7] += lam;
        const int ok2 = chol_solve(A, 6, b, x6);
        if (!ok2) memset(x6, 0, sizeof(x6));
        pose_oplus(pose, x6, new_pose);
        double temp_chi = pose_edges(K, new_pose, n, Xw, obs, is2, outlier, NULL, chi2_t, NULL, robust, huber_delta);
        for (int i = 0; i < n; ++i)
          if (!outlier[i]) cached[i] = chi2_t[i];
        if (!ok2) temp_chi = DBL_MAX;
        double scale = 1e-3;
        for (int r = 0; r < 6; ++r) scale += x6[r] * (lam * x6[r] + b[r]);
        rho = (current_chi - temp_chi) / scale;
        ++trials;
        if (rho > 0 && isfinite(temp_chi)) {
          const double alpha = fmin(1.0 - pow(2 * rho - 1, 3), good_hi);
          lam *= fmax(good_lo, alpha);
          ni = 2.0;
          current_chi = temp_chi;
          memcpy(pose, new_pose, sizeof(pose));
        } else {
          lam *= ni;
          ni *= 2;
        }
        ++qmax;
      } while (rho < 0 && qmax < max_trials);
      if (qmax == max_trials || rho == 0) break;
      if ((ini_chi - current_chi) * 1e3 < ini_chi) ++nbad_lm; else nbad_lm = 0;
      if (nbad_lm >= 3) break;
    }
    /* classification: outlier edges are re-evaluated at the round's final pose, inliers keep the cached chi2 */
    pose_edges(K, pose, n, Xw, obs, is2, NULL, NULL, chi2, NULL, 0, huber_delta);
    n_bad = 0;
    for (int i = 0; i < n; ++i) {
      const double used = outlier[i] ? chi2[i] : cached[i];
      outlier[i] = (float)used > 5.991f;
      n_bad += outlier[i];
    }
    if (n < 10) break;
  }
  memcpy(pose_out, pose, sizeof(pose));
  memcpy(outlier_out, outlier, (size_t)n);
  *n_inliers = n - n_bad;
  *n_trials = trials;
  free(outlier); free(err); free(chi2); free(chi2_t); free(cached); free(Xc);
  return 0;
}
