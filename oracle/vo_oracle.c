/*
 * vo_oracle.c — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see vo_oracle.h).  Build: gcc -O3 -std=c11 -ffp-contract=off -fPIC -shared (oracle/Makefile);
 * -ffp-contract=off + x86-64 baseline (SSE2) means every + - * / below is one IEEE FP32
 * rounding, exactly like the reference's "-O3 -DNDEBUG" build without -march (CMakeLists.txt:7).
 */
#include "vo_oracle.h"

#include <math.h>
#include <stddef.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * NN distance.  brute_force_search.h:14,34: (p-query).tail(Dim-1).squaredNorm().
 * tail(n) takes a run-time n, so the expression is a dynamic-size block and Eigen reduces it
 * with its linear-vectorised traversal over Packet4f (SSE2):
 *     s_i = fl(fl(p_i-q_i)^2)
 *     two packet accumulators P0 = s[0..3], P1 = s[4..7], advanced 8 lanes at a time,
 *     P0 += P1, (+ one trailing packet if n%8 >= 4),
 *     horizontal add  r = (P0[0]+P0[2]) + (P0[1]+P0[3]),
 *     scalar tail     r = r + s_j  for the n%4 remaining coefficients,
 * and a plain left-to-right sum when n < 4.
 * ---------------------------------------------------------------------------------------- */
float oracle_sqdist(const float* p, const float* q, int dim) {
  float s[64];
  if (dim <= 0) return 0.f;
  for (int i = 0; i < dim; ++i) {
    float d = p[i] - q[i];
    s[i] = d * d;
  }
  const int n4 = (dim / 4) * 4; /* lanes covered by whole packets            */
  const int n8 = (dim / 8) * 8; /* lanes covered by the 2-packet unrolled loop */
  float r;
  if (n4 > 0) {
    float a[4] = {s[0], s[1], s[2], s[3]};
    if (n4 > 4) {
      float c[4] = {s[4], s[5], s[6], s[7]};
      for (int i = 8; i < n8; i += 8)
        for (int l = 0; l < 4; ++l) {
          a[l] = a[l] + s[i + l];
          c[l] = c[l] + s[i + 4 + l];
        }
      for (int l = 0; l < 4; ++l) a[l] = a[l] + c[l];
      if (n4 > n8)
        for (int l = 0; l < 4; ++l) a[l] = a[l] + s[n8 + l];
    }
    r = (a[0] + a[2]) + (a[1] + a[3]);
    for (int i = n4; i < dim; ++i) r = r + s[i];
  } else {
    r = s[0];
    for (int i = 1; i < dim; ++i) r = r + s[i];
  }
  return r;
}

/* brute_force_search.h:22-41 — strict '<' against norm*norm, then against the running best:
 * the first (lowest-index) minimum wins; "no match" is a null pointer -> -1 here. */
void oracle_nn_best_match(const float* map, int64_t n_rows, int row_stride, int skip_cols,
                          const float* queries, int64_t n_queries, int query_stride, float norm,
                          int32_t* best_idx, float* best_d2) {
  const int dim = row_stride - skip_cols;
  for (int64_t qi = 0; qi < n_queries; ++qi) {
    const float* q = queries + qi * (int64_t)query_stride + skip_cols;
    int32_t best = -1;
    float best_sq = norm * norm; /* :31 */
    for (int64_t r = 0; r < n_rows; ++r) {
      float d2 = oracle_sqdist(map + r * (int64_t)row_stride + skip_cols, q, dim); /* :34 */
      if (d2 < best_sq) { /* :35 */
        best = (int32_t)r;
        best_sq = d2;
      }
    }
    best_idx[qi] = best;
    if (best_d2) best_d2[qi] = best_sq;
  }
}

/* brute_force_search.h:3-20 */
void oracle_nn_radius_search(const float* map, int64_t n_rows, int row_stride, int skip_cols,
                             const float* queries, int64_t n_queries, int query_stride,
                             float norm, int32_t* counts, int32_t* idx_out,
                             int32_t max_per_query) {
  const int dim = row_stride - skip_cols;
  const float sq = norm * norm; /* :10 */
  for (int64_t qi = 0; qi < n_queries; ++qi) {
    const float* q = queries + qi * (int64_t)query_stride + skip_cols;
    int32_t m = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
      if (oracle_sqdist(map + r * (int64_t)row_stride + skip_cols, q, dim) < sq) { /* :14 */
        if (idx_out && m < max_per_query) idx_out[qi * (int64_t)max_per_query + m] = (int32_t)r;
        ++m;
      }
    }
    counts[qi] = m;
  }
}

/* ------------------------------------------------------------------------------------------
 * small fixed-size algebra, column-major, one rounding per operation
 * ---------------------------------------------------------------------------------------- */
/* Fixed-size products are coefficient-based in Eigen: each coefficient is
 * lhs.row(i).cwiseProduct(rhs.col(j)).sum(), and a 3-term sum is unrolled by halves
 * (redux_novec_unroller):  s0 + (s1 + s2). */
static void mat3_vec(const float* M, const float v[3], float out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = M[i] * v[0] + (M[3 + i] * v[1] + M[6 + i] * v[2]);
}
/* Isometry3f * Vector3f: res = translation; res += linear*v  (Transform.h,
 * transform_right_product_impl) -> t + (s0 + (s1 + s2)).  T is a column-major 4x4. */
static void iso_point(const float* T, const float p[3], float out[3]) {
  for (int i = 0; i < 3; ++i)
    out[i] = T[12 + i] + (T[i] * p[0] + (T[4 + i] * p[1] + T[8 + i] * p[2]));
}

/* camera.h:25-37 */
int oracle_project_point(const oracle_camera* cam, const float wp[3], float out[2]) {
  float cp[3];
  iso_point(cam->T, wp, cp); /* :27 */
  if (cp[2] > (float)cam->z_far || cp[2] < (float)cam->z_near) return 0; /* :28 */
  float pp[3];
  mat3_vec(cam->K, cp, pp);                        /* :30 */
  const float inv = (float)(1. / (double)pp[2]);   /* :31, 1./z is a double, demoted */
  out[0] = pp[0] * inv;
  out[1] = pp[1] * inv;
  if (out[0] < 0 || out[0] > (float)(cam->cols - 1)) return 0; /* :32 */
  if (out[1] < 0 || out[1] > (float)(cam->rows - 1)) return 0; /* :34 */
  return 1;
}

/* camera.cpp:16-37 */
void oracle_project_points(const oracle_camera* cam, const float* world, int64_t n,
                           int keep_indices, float* out_image, int64_t* n_out,
                           int64_t* n_inside) {
  int64_t w = 0, inside = 0;
  for (int64_t i = 0; i < n; ++i) {
    float uv[2];
    int ok = oracle_project_point(cam, world + 3 * i, uv);
    if (ok) ++inside;
    else uv[0] = uv[1] = -1.f; /* :21,30 */
    if (keep_indices || ok) {  /* :31 */
      out_image[2 * w] = uv[0];
      out_image[2 * w + 1] = uv[1];
      ++w;
    }
  }
  *n_out = w;
  *n_inside = inside;
}

/* ------------------------------------------------------------------------------------------
 * LDLT with diagonal pivoting, the algorithm behind Eigen's ldlt().solve() that
 * picp_solver.cpp:109 and utils.cpp:40 call.  A is symmetric, column-major, n <= 6.
 * ---------------------------------------------------------------------------------------- */
void oracle_ldlt_solve(int n, const float* A_in, const float* rhs, float* x) {
  float A[36];
  int tr[6];
  float tmp[6];
  memcpy(A, A_in, sizeof(float) * (size_t)(n * n));
#define AT(i, j) A[(j) * n + (i)]
  for (int k = 0; k < n; ++k) {
    /* pivot = largest |diagonal| of the trailing block */
    int piv = k;
    float big = fabsf(AT(k, k));
    for (int i = k + 1; i < n; ++i)
      if (fabsf(AT(i, i)) > big) {
        big = fabsf(AT(i, i));
        piv = i;
      }
    tr[k] = piv;
    if (piv != k) { /* symmetric swap touching only the lower triangle */
      for (int j = 0; j < k; ++j) {
        float t = AT(k, j);
        AT(k, j) = AT(piv, j);
        AT(piv, j) = t;
      }
      for (int i = piv + 1; i < n; ++i) {
        float t = AT(i, k);
        AT(i, k) = AT(i, piv);
        AT(i, piv) = t;
      }
      {
        float t = AT(k, k);
        AT(k, k) = AT(piv, piv);
        AT(piv, piv) = t;
      }
      for (int i = k + 1; i < piv; ++i) {
        float t = AT(i, k);
        AT(i, k) = AT(piv, i);
        AT(piv, i) = t;
      }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; ++j) tmp[j] = AT(j, j) * AT(k, j);
      float acc = 0.f;
      for (int j = 0; j < k; ++j) acc = acc + AT(k, j) * tmp[j];
      AT(k, k) = AT(k, k) - acc;
      for (int i = 0; i < rs; ++i) {
        float a2 = 0.f;
        for (int j = 0; j < k; ++j) a2 = a2 + AT(k + 1 + i, j) * tmp[j];
        AT(k + 1 + i, k) = AT(k + 1 + i, k) - a2;
      }
    }
    const float akk = AT(k, k);
    if (rs > 0 && fabsf(akk) > 0.f)
      for (int i = 0; i < rs; ++i) AT(k + 1 + i, k) = AT(k + 1 + i, k) / akk;
  }
  /* solve: x = P^T L^-T D^+ L^-1 P rhs */
  float y[6];
  for (int i = 0; i < n; ++i) y[i] = rhs[i];
  for (int k = 0; k < n; ++k)
    if (tr[k] != k) {
      float t = y[k];
      y[k] = y[tr[k]];
      y[tr[k]] = t;
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) y[i] = y[i] - AT(i, j) * y[j];
  for (int i = 0; i < n; ++i) {
    if (fabsf(AT(i, i)) > 1.17549435e-38f) y[i] = y[i] / AT(i, i);
    else y[i] = 0.f;
  }
  for (int i = n - 1; i >= 0; --i)
    for (int j = i + 1; j < n; ++j) y[i] = y[i] - AT(j, i) * y[j];
  for (int k = n - 1; k >= 0; --k)
    if (tr[k] != k) {
      float t = y[k];
      y[k] = y[tr[k]];
      y[tr[k]] = t;
    }
  for (int i = 0; i < n; ++i) x[i] = y[i];
#undef AT
}

/* utils.h:16-78 — R = Rx(v3)*Ry(v4)*Rz(v5), t = v[0..2]. */
static void mat3_mul(const float* A, const float* B, float* C) {
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i)
      C[j * 3 + i] = A[i] * B[j * 3] + (A[3 + i] * B[j * 3 + 1] + A[6 + i] * B[j * 3 + 2]);
}
void oracle_v2t_euler(const float v[6], float T[16]) {
  const float sx = sinf(v[3]), cx = cosf(v[3]);
  const float sy = sinf(v[4]), cy = cosf(v[4]);
  const float sz = sinf(v[5]), cz = cosf(v[5]);
  /* column-major */
  const float Rx[9] = {1, 0, 0, 0, cx, sx, 0, -sx, cx};  /* utils.h:23-26 */
  const float Ry[9] = {cy, 0, -sy, 0, 1, 0, sy, 0, cy};  /* utils.h:39-42 */
  const float Rz[9] = {cz, sz, 0, -sz, cz, 0, 0, 0, 1};  /* utils.h:55-58 */
  float Rxy[9], R[9];
  mat3_mul(Rx, Ry, Rxy);
  mat3_mul(Rxy, Rz, R); /* utils.h:66 */
  memset(T, 0, sizeof(float) * 16);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) T[j * 4 + i] = R[j * 3 + i];
  T[12] = v[0];
  T[13] = v[1];
  T[14] = v[2];
  T[15] = 1.f;
}

/* picp_solver.cpp:25-53 */
static int picp_error_and_jacobian(const oracle_camera* cam, const float wp[3],
                                   const float meas[2], float e[2], float J[12] /*2x6 col-major*/) {
  float pred[2];
  if (!oracle_project_point(cam, wp, pred)) return 0; /* :32-34 */
  e[0] = pred[0] - meas[0];                           /* :35 */
  e[1] = pred[1] - meas[1];
  float cp[3];
  iso_point(cam->T, wp, cp); /* :38 */
  /* Jr = [I | skew(-cp)] :39-41, skew per utils.h:96-102 */
  const float n0 = -cp[0], n1 = -cp[1], n2 = -cp[2];
  const float S[9] = {0.f, n2, -n1, -n2, 0.f, n0, n1, -n0, 0.f}; /* column-major skew(-cp) */
  float ph[3];
  mat3_vec(cam->K, cp, ph);                     /* :43 */
  const float iz = (float)(1. / (double)ph[2]); /* :44 */
  const float iz2 = iz * iz;                    /* :45 */
  const float Jp[6] = {iz, 0.f, 0.f, iz, -ph[0] * iz2, -ph[1] * iz2}; /* 2x3 col-major :47-49 */
  /* A = Jp*K (2x3), J = A*Jr (2x6)  :51 */
  float A[6];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 2; ++i)
      A[j * 2 + i] = Jp[i] * cam->K[j * 3] +
                     (Jp[2 + i] * cam->K[j * 3 + 1] + Jp[4 + i] * cam->K[j * 3 + 2]);
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 2; ++i) {
      J[j * 2 + i] = A[j * 2 + i];
      J[(3 + j) * 2 + i] = A[i] * S[j * 3] + (A[2 + i] * S[j * 3 + 1] + A[4 + i] * S[j * 3 + 2]);
    }
  return 1;
}

int oracle_picp_one_round(oracle_picp_state* st, const oracle_camera* cam_params,
                          const float* world, const float* image, const int32_t* pairs,
                          int64_t n_pairs, int keep_outliers, float kernel_threshold,
                          float damping, int32_t min_num_inliers) {
  oracle_camera cam = *cam_params;
  memcpy(cam.T, st->T, sizeof(cam.T));
  float* H = st->H;
  float* b = st->b;
  /* linearize, picp_solver.cpp:55-96 */
  memset(H, 0, sizeof(float) * 36);
  memset(b, 0, sizeof(float) * 6);
  st->num_inliers = 0;
  st->chi_inliers = 0.f;
  st->chi_outliers = 0.f;
  for (int64_t c = 0; c < n_pairs; ++c) {
    const int32_t ref_idx = pairs[2 * c];      /* :66  image/measurement index */
    const int32_t cur_idx = pairs[2 * c + 1];  /* :67  world index             */
    float e[2], J[12];
    if (!picp_error_and_jacobian(&cam, world + 3 * (int64_t)cur_idx,
                                 image + 2 * (int64_t)ref_idx, e, J))
      continue; /* :72 */
    const float chi = e[0] * e[0] + e[1] * e[1]; /* :75 */
    float lambda = 1.f;
    int inlier = 1;
    if (chi > kernel_threshold) {                              /* :78 */
      lambda = (float)sqrt((double)(kernel_threshold / chi));  /* :80 */
      inlier = 0;
      st->chi_outliers = st->chi_outliers + chi; /* :82 */
    } else {
      st->chi_inliers = st->chi_inliers + chi; /* :86 */
      st->num_inliers++;
    }
    if (inlier || keep_outliers) { /* :90-94 */
      for (int cc = 0; cc < 6; ++cc)
        for (int r = 0; r < 6; ++r) {
          const float jtj = J[r * 2] * J[cc * 2] + J[r * 2 + 1] * J[cc * 2 + 1];
          H[cc * 6 + r] = H[cc * 6 + r] + jtj * lambda;
        }
      for (int r = 0; r < 6; ++r) {
        const float jte = J[r * 2] * e[0] + J[r * 2 + 1] * e[1];
        b[r] = b[r] + jte * lambda;
      }
    }
  }
  /* oneRound, :98-112 */
  for (int i = 0; i < 6; ++i) H[i * 6 + i] = H[i * 6 + i] + damping; /* :102 */
  st->rounds_done++;
  if (st->num_inliers < min_num_inliers) { /* :103-107 */
    st->last_ok = 0;
    return 0;
  }
  float nb[6], dx[6];
  for (int i = 0; i < 6; ++i) nb[i] = -b[i];
  oracle_ldlt_solve(6, H, nb, dx); /* :109 */
  float D[16], Tn[16];
  oracle_v2t_euler(dx, D);
  /* :110  pose <- D * pose (isometry product: R = Rd*R, t = Rd*t + td) */
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 3; ++i) {
      float acc = D[i] * st->T[j * 4] + (D[4 + i] * st->T[j * 4 + 1] + D[8 + i] * st->T[j * 4 + 2]);
      if (j == 3) acc = acc + D[12 + i];
      Tn[j * 4 + i] = acc;
    }
  Tn[3] = Tn[7] = Tn[11] = 0.f;
  Tn[15] = 1.f;
  memcpy(st->T, Tn, sizeof(Tn));
  st->last_ok = 1;
  return 1;
}

/* ---- float64 truth of the same round (not the reference's arithmetic) -------------------- */
static void solve6_f64(const double* Hin, const double* rhs, double* x) {
  /* Gaussian elimination with partial pivoting */
  double A[6][7];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) A[i][j] = Hin[j * 6 + i];
    A[i][6] = rhs[i];
  }
  for (int k = 0; k < 6; ++k) {
    int p = k;
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[i][k]) > fabs(A[p][k])) p = i;
    if (p != k)
      for (int j = 0; j < 7; ++j) {
        double t = A[k][j];
        A[k][j] = A[p][j];
        A[p][j] = t;
      }
    for (int i = k + 1; i < 6; ++i) {
      double f = A[i][k] / A[k][k];
      for (int j = k; j < 7; ++j) A[i][j] -= f * A[k][j];
    }
  }
  for (int i = 5; i >= 0; --i) {
    double s = A[i][6];
    for (int j = i + 1; j < 6; ++j) s -= A[i][j] * x[j];
    x[i] = s / A[i][i];
  }
}

int oracle_picp_one_round_f64(double T[16], double H[36], double b[6], double stats[3],
                              const oracle_camera* cam, const float* world, const float* image,
                              const int32_t* pairs, int64_t n_pairs, int keep_outliers,
                              double thr, double damping) {
  memset(H, 0, sizeof(double) * 36);
  memset(b, 0, sizeof(double) * 6);
  stats[0] = stats[1] = stats[2] = 0.0; /* chi_in, chi_out, n_in */
  double K[9];
  for (int i = 0; i < 9; ++i) K[i] = cam->K[i];
  for (int64_t c = 0; c < n_pairs; ++c) {
    const float* wp = world + 3 * (int64_t)pairs[2 * c + 1];
    const float* ms = image + 2 * (int64_t)pairs[2 * c];
    double cp[3], ph[3];
    for (int i = 0; i < 3; ++i) cp[i] = T[i] * wp[0] + T[4 + i] * wp[1] + T[8 + i] * wp[2] + T[12 + i];
    if (cp[2] > cam->z_far || cp[2] < cam->z_near) continue;
    for (int i = 0; i < 3; ++i) ph[i] = K[i] * cp[0] + K[3 + i] * cp[1] + K[6 + i] * cp[2];
    const double iz = 1.0 / ph[2];
    const double u = ph[0] * iz, v = ph[1] * iz;
    if (u < 0 || u > cam->cols - 1 || v < 0 || v > cam->rows - 1) continue;
    const double e[2] = {u - ms[0], v - ms[1]};
    const double iz2 = iz * iz;
    double A[6], J[12];
    for (int j = 0; j < 3; ++j) {
      A[j * 2] = iz * K[j * 3] - ph[0] * iz2 * K[j * 3 + 2];
      A[j * 2 + 1] = iz * K[j * 3 + 1] - ph[1] * iz2 * K[j * 3 + 2];
    }
    const double S[9] = {0, -cp[2], cp[1], cp[2], 0, -cp[0], -cp[1], cp[0], 0};
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 2; ++i) {
        J[j * 2 + i] = A[j * 2 + i];
        J[(3 + j) * 2 + i] = A[i] * S[j * 3] + A[2 + i] * S[j * 3 + 1] + A[4 + i] * S[j * 3 + 2];
      }
    const double chi = e[0] * e[0] + e[1] * e[1];
    double lambda = 1.0;
    int inlier = 1;
    if (chi > thr) {
      lambda = sqrt(thr / chi);
      inlier = 0;
      stats[1] += chi;
    } else {
      stats[0] += chi;
      stats[2] += 1.0;
    }
    if (inlier || keep_outliers) {
      for (int cc = 0; cc < 6; ++cc)
        for (int r = 0; r < 6; ++r)
          H[cc * 6 + r] += (J[r * 2] * J[cc * 2] + J[r * 2 + 1] * J[cc * 2 + 1]) * lambda;
      for (int r = 0; r < 6; ++r) b[r] += (J[r * 2] * e[0] + J[r * 2 + 1] * e[1]) * lambda;
    }
  }
  for (int i = 0; i < 6; ++i) H[i * 6 + i] += damping;
  double nb[6], dx[6];
  for (int i = 0; i < 6; ++i) nb[i] = -b[i];
  solve6_f64(H, nb, dx);
  const double sx = sin(dx[3]), cx = cos(dx[3]), sy = sin(dx[4]), cy = cos(dx[4]),
               sz = sin(dx[5]), cz = cos(dx[5]);
  /* R = Rx*Ry*Rz, row-major entries */
  const double R[3][3] = {{cy * cz, -cy * sz, sy},
                          {sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy},
                          {-cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy}};
  double Tn[16];
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 3; ++i) {
      double acc = R[i][0] * T[j * 4] + R[i][1] * T[j * 4 + 1] + R[i][2] * T[j * 4 + 2];
      if (j == 3) acc += dx[i];
      Tn[j * 4 + i] = acc;
    }
  Tn[3] = Tn[7] = Tn[11] = 0.0;
  Tn[15] = 1.0;
  memcpy(T, Tn, sizeof(Tn));
  return 1;
}

/* ------------------------------------------------------------------------------------------
 * triangulation
 * ---------------------------------------------------------------------------------------- */
/* Matrix3f::inverse(): Eigen's compute_inverse_size3 — cofactor_3x3<i,j> with cyclic indices,
 * determinant = cofactors_col0 . matrix.col(0) (3-term sum by halves), result(i,j) =
 * cofactor<j,i> * (1/det). */
static float cof3(const float* M, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
#define m(r, c) M[(c) * 3 + (r)]
  return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
}
static void mat3_inverse(const float* M, float* out) {
  const float c0 = cof3(M, 0, 0), c1 = cof3(M, 1, 0), c2 = cof3(M, 2, 0);
  const float det = c0 * m(0, 0) + (c1 * m(1, 0) + c2 * m(2, 0));
  const float id = 1.f / det;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      out[j * 3 + i] = (i == 0 ? (j == 0 ? c0 : (j == 1 ? c1 : c2)) : cof3(M, j, i)) * id;
#undef m
}

/* utils.cpp:36-49 */
static int triangulate_point(const float d1[3], const float d2[3], const float p2[3], float p[3]) {
  /* D = [-d1 d2];  ss = -(D^T D).ldlt().solve(D^T p2)   :37-40 */
  float nd1[3] = {-d1[0], -d1[1], -d1[2]};
  float A[4], rhs[2], ss[2];
  A[0] = nd1[0] * nd1[0] + (nd1[1] * nd1[1] + nd1[2] * nd1[2]);
  A[1] = d2[0] * nd1[0] + (d2[1] * nd1[1] + d2[2] * nd1[2]);
  A[2] = nd1[0] * d2[0] + (nd1[1] * d2[1] + nd1[2] * d2[2]);
  A[3] = d2[0] * d2[0] + (d2[1] * d2[1] + d2[2] * d2[2]);
  rhs[0] = nd1[0] * p2[0] + (nd1[1] * p2[1] + nd1[2] * p2[2]);
  rhs[1] = d2[0] * p2[0] + (d2[1] * p2[1] + d2[2] * p2[2]);
  oracle_ldlt_solve(2, A, rhs, ss);
  ss[0] = -ss[0];
  ss[1] = -ss[1];
  if (ss[0] < 0 || ss[1] < 0) return 0; /* :41 */
  for (int i = 0; i < 3; ++i) {
    const float a = ss[0] * d1[i];          /* :44 */
    const float c = p2[i] + ss[1] * d2[i];  /* :45 */
    p[i] = 0.5f * (a + c);                  /* :47 */
  }
  return 1;
}

int64_t oracle_triangulate_points(const float K[9], const float X[16], const int32_t* corr,
                                  int64_t n_corr, const float* p1, const float* p2,
                                  const float* app2, float* out_points, int32_t* out_corr_new,
                                  float* out_app, int32_t* out_src) {
  /* iX = X.inverse() (isometry: R^T, -R^T t), iK = k.inverse(), iRiK = iX.linear()*iK,
   * t = iX.translation()    utils.cpp:53-56 / :79-82 / :108-111 */
  float iR[9], t[3], iK[9], iRiK[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) iR[j * 3 + i] = X[i * 4 + j];
  {
    const float tx[3] = {X[12], X[13], X[14]};
    float r[3];
    mat3_vec(iR, tx, r);
    t[0] = -r[0];
    t[1] = -r[1];
    t[2] = -r[2];
  }
  mat3_inverse(K, iK);
  mat3_mul(iR, iK, iRiK);
  int64_t ns = 0;
  for (int64_t c = 0; c < n_corr; ++c) {
    const int32_t i1 = corr[2 * c], i2 = corr[2 * c + 1]; /* :87-88 */
    const float h1[3] = {p1[2 * (int64_t)i1], p1[2 * (int64_t)i1 + 1], 1.f};
    const float h2[3] = {p2[2 * (int64_t)i2], p2[2 * (int64_t)i2 + 1], 1.f};
    float d1[3], d2[3], p[3];
    mat3_vec(iK, h1, d1);   /* :91 */
    mat3_vec(iRiK, h2, d2); /* :94 */
    if (triangulate_point(d1, d2, t, p)) {
      if (out_corr_new) { /* :97 */
        out_corr_new[2 * ns] = i2;
        out_corr_new[2 * ns + 1] = (int32_t)ns;
      }
      out_points[3 * ns] = p[0];
      out_points[3 * ns + 1] = p[1];
      out_points[3 * ns + 2] = p[2];
      if (out_app && app2) memcpy(out_app + 10 * ns, app2 + 10 * (int64_t)i2, 10 * sizeof(float)); /* :127 */
      if (out_src) out_src[ns] = (int32_t)c;
      ++ns;
    }
  }
  return ns;
}
