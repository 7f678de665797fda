/* Oracle (TEST INFRASTRUCTURE ONLY): plain-C restatements of the reference's CPU loops for descriptor matching and the
 * keyframe-database scan, used (a) to cross-check the numpy oracle and (b) as the honest multi-core CPU baseline of
 * bench.py.  Nothing under hfnet_slam_b200/ links or loads this file.  Paths are relative to the reference repository.
 *
 *   ref_match_cos_mutual   src/Matcher.cc:845-889   S = D1 * D2^T (Eigen sgemm), per-row arg-max with S > floor (strict),
 *                                                   per-column arg-max, mutual check; lowest index wins ties
 *   ref_match_bf_l2        src/Matcher.cc:229-253   cv::BFMatcher(NORM_L2, crossCheck = true).match + dist < max_dist:
 *                                                   nearest train row per query row (L2 of the difference), kept iff the
 *                                                   query row is also the nearest of that train row
 *   ref_kfdb_scores        src/KeyFrameDatabase.cc:86-96   score_i = max(0, 1 - (q - d_i).norm()) for EVERY keyframe, the
 *                                                   literal per-keyframe loop (fp32)
 * OpenMP parallelises the outer loops (the reference gives Eigen half the cores, src/System.cc:45-48; its database scan
 * holds the database mutex and is serial -- bench.py reports the thread count it used).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static float dotf(const float* a, const float* b, int dim) {
  float s = 0.f;
#pragma omp simd reduction(+ : s)   /* vectorised like Eigen's kernels: the summation order is not the reference's contract */
  for (int k = 0; k < dim; ++k) s += a[k] * b[k];
  return s;
}

/* S is materialised like the reference's MatrixXf (n1 x n2 floats). */
void ref_match_cos_mutual(const float* D1, int n1, const float* D2, int n2, int dim, float floor_, int32_t* match12,
                          float* score12) {
  float* S = (float*)malloc((size_t)(n1 > 0 ? n1 : 1) * (size_t)(n2 > 0 ? n2 : 1) * sizeof(float));
  int32_t* best1 = (int32_t*)malloc((size_t)(n1 > 0 ? n1 : 1) * sizeof(int32_t));
  int32_t* best2 = (int32_t*)malloc((size_t)(n2 > 0 ? n2 : 1) * sizeof(int32_t));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n1; ++i)
    for (int j = 0; j < n2; ++j) S[(size_t)i * n2 + j] = dotf(D1 + (size_t)i * dim, D2 + (size_t)j * dim, dim);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n1; ++i) {            /* :853-872 row pass */
    float bs = floor_;
    int bj = -1;
    for (int j = 0; j < n2; ++j) {
      const float s = S[(size_t)i * n2 + j];
      if (s > bs) {
        bs = s;
        bj = j;
      }
    }
    best1[i] = bj;
  }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < n2; ++j) {            /* :875-889 column pass */
    float bs = floor_;
    int bi = -1;
    for (int i = 0; i < n1; ++i) {
      const float s = S[(size_t)i * n2 + j];
      if (s > bs) {
        bs = s;
        bi = i;
      }
    }
    best2[j] = bi;
  }
  for (int i = 0; i < n1; ++i) {
    const int j = best1[i];
    if (j >= 0 && best2[j] == i) {
      match12[i] = j;
      score12[i] = S[(size_t)i * n2 + j];
    } else {
      match12[i] = -1;
      score12[i] = 0.f;
    }
  }
  free(S);
  free(best1);
  free(best2);
}

static float l2f(const float* a, const float* b, int dim) {
  float s = 0.f;
#pragma omp simd reduction(+ : s)
  for (int k = 0; k < dim; ++k) {
    const float d = a[k] - b[k];
    s += d * d;
  }
  return sqrtf(s);
}

void ref_match_bf_l2(const float* A, int na, const float* B, int nb, int dim, float max_dist, int32_t* match_ab,
                     float* dist_ab) {
  int32_t* nn_a = (int32_t*)malloc((size_t)(na > 0 ? na : 1) * sizeof(int32_t));
  float* d_a = (float*)malloc((size_t)(na > 0 ? na : 1) * sizeof(float));
  int32_t* nn_b = (int32_t*)malloc((size_t)(nb > 0 ? nb : 1) * sizeof(int32_t));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < na; ++i) {
    float bd = INFINITY;
    int bj = -1;
    for (int j = 0; j < nb; ++j) {
      const float d = l2f(A + (size_t)i * dim, B + (size_t)j * dim, dim);
      if (d < bd) {
        bd = d;
        bj = j;
      }
    }
    nn_a[i] = bj;
    d_a[i] = bd;
  }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < nb; ++j) {
    float bd = INFINITY;
    int bi = -1;
    for (int i = 0; i < na; ++i) {
      const float d = l2f(A + (size_t)i * dim, B + (size_t)j * dim, dim);
      if (d < bd) {
        bd = d;
        bi = i;
      }
    }
    nn_b[j] = bi;
  }
  for (int i = 0; i < na; ++i) {
    const int j = nn_a[i];
    if (j >= 0 && nn_b[j] == i && d_a[i] < max_dist) {
      match_ab[i] = j;
      dist_ab[i] = d_a[i];
    } else {
      match_ab[i] = -1;
      dist_ab[i] = 0.f;
    }
  }
  free(nn_a);
  free(d_a);
  free(nn_b);
}

void ref_kfdb_scores(const float* q, const float* db, int n, int dim, float* scores) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    const float* d = db + (size_t)i * dim;
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int k = 0; k < dim; ++k) {
      const float t = q[k] - d[k];
      s += t * t;
    }
    const float sc = 1.f - sqrtf(s);
    scores[i] = sc > 0.f ? sc : 0.f;
  }
}
