// linalg.cuh — tiny fixed-size device linear algebra shared by picp.cu and triangulate.cu.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#include "common.cuh"
#endif

namespace vo {

// ---- pivoted LDL^T solve, same algorithm as Eigen's LDLT (picp_solver.cpp:109) -------------
#if defined(__CUDACC__)
#define VO_HD __host__ __device__ __forceinline__
#else
#define VO_HD inline
#endif

// Pivoted LDL^T solve with the operation sequence of Eigen's LDLT (ldlt_inplace<Lower>::unblocked
// + LDLT::_solve_impl; picp_solver.cpp:109, utils.cpp:40) — the same sequence as
// oracle_ldlt_solve.  Everything lives in registers: all loops have compile-time bounds after
// unrolling, and the only run-time quantity, the pivot row, is applied through conditional swaps
// with static indices, so no local-memory array is ever indexed dynamically (the single thread
// that solves the 6x6 system sits on the critical path of every Gauss-Newton round).
template <int N>
VO_HD void ldlt_solve_dev(float* A /* N*N col-major, destroyed */, const float* rhs, float* x) {
  int tr[N];
#define AT(i, j) A[(j) * N + (i)]
#define VO_CSWAP(c, u, v)       \
  {                             \
    const float _a = (u), _b = (v); \
    (u) = (c) ? _b : _a;        \
    (v) = (c) ? _a : _b;        \
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    float big = fabsf(AT(k, k));
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const float d = fabsf(AT(i, i));
      const bool g = d > big;
      big = g ? d : big;
      piv = g ? i : piv;
    }
    tr[k] = piv;
#pragma unroll
    for (int c = k + 1; c < N; ++c) {  // symmetric swap k <-> c inside the lower triangle
      const bool sw = (piv == c);
#pragma unroll
      for (int j = 0; j < k; ++j) VO_CSWAP(sw, AT(k, j), AT(c, j));
#pragma unroll
      for (int i = c + 1; i < N; ++i) VO_CSWAP(sw, AT(i, k), AT(i, c));
      VO_CSWAP(sw, AT(k, k), AT(c, c));
#pragma unroll
      for (int i = k + 1; i < c; ++i) VO_CSWAP(sw, AT(i, k), AT(c, i));
    }
    if (k > 0) {
      float tmp[N];
#pragma unroll
      for (int j = 0; j < k; ++j) tmp[j] = AT(j, j) * AT(k, j);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < k; ++j) acc = acc + AT(k, j) * tmp[j];
      AT(k, k) = AT(k, k) - acc;
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        float a2 = 0.f;
#pragma unroll
        for (int j = 0; j < k; ++j) a2 = a2 + AT(i, j) * tmp[j];
        AT(i, k) = AT(i, k) - a2;
      }
    }
    const float akk = AT(k, k);
    const bool nz = fabsf(akk) > 0.f;
#pragma unroll
    for (int i = k + 1; i < N; ++i) AT(i, k) = nz ? AT(i, k) / akk : AT(i, k);
  }
  float y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = rhs[i];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int c = k + 1; c < N; ++c) VO_CSWAP(tr[k] == c, y[k], y[c]);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) y[i] = y[i] - AT(i, j) * y[j];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = (fabsf(AT(i, i)) > 1.17549435e-38f) ? y[i] / AT(i, i) : 0.f;
#pragma unroll
  for (int i = N - 1; i >= 0; --i)
#pragma unroll
    for (int j = i + 1; j < N; ++j) y[i] = y[i] - AT(j, i) * y[j];
#pragma unroll
  for (int k = N - 1; k >= 0; --k)
#pragma unroll
    for (int c = k + 1; c < N; ++c) VO_CSWAP(tr[k] == c, y[k], y[c]);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = y[i];
#undef VO_CSWAP
#undef AT
}

// The same pivoted LDL^T with every division replaced by a multiplication with the pivot's
// reciprocal (one reciprocal per column instead of N-k divisions, none in the diagonal solve): the
// PICP kernels' variant, where the solve is a serial section every other thread waits for.  Same
// pivots, results within a few ulp of ldlt_solve_dev (which stays bit-pinned to the oracle).
VO_HD float vo_rcp_seed(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.f / x;
#endif
}
template <int N>
VO_HD void ldlt_solve_recip_dev(float* A /* N*N col-major, destroyed */, const float* rhs, float* x) {
  int tr[N];
  float rinv[N];  // 1 / D(k), or 0 for a vanishing pivot
#define AT(i, j) A[(j) * N + (i)]
#define VO_CSWAP(c, u, v)       \
  {                             \
    const float _a = (u), _b = (v); \
    (u) = (c) ? _b : _a;        \
    (v) = (c) ? _a : _b;        \
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    float big = fabsf(AT(k, k));
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const float d = fabsf(AT(i, i));
      const bool g = d > big;
      big = g ? d : big;
      piv = g ? i : piv;
    }
    tr[k] = piv;
#pragma unroll
    for (int c = k + 1; c < N; ++c) {  // symmetric swap k <-> c inside the lower triangle
      const bool sw = (piv == c);
#pragma unroll
      for (int j = 0; j < k; ++j) VO_CSWAP(sw, AT(k, j), AT(c, j));
#pragma unroll
      for (int i = c + 1; i < N; ++i) VO_CSWAP(sw, AT(i, k), AT(i, c));
      VO_CSWAP(sw, AT(k, k), AT(c, c));
#pragma unroll
      for (int i = k + 1; i < c; ++i) VO_CSWAP(sw, AT(i, k), AT(c, i));
    }
    if (k > 0) {
      float tmp[N];
#pragma unroll
      for (int j = 0; j < k; ++j) tmp[j] = AT(j, j) * AT(k, j);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < k; ++j) acc = acc + AT(k, j) * tmp[j];
      AT(k, k) = AT(k, k) - acc;
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        float a2 = 0.f;
#pragma unroll
        for (int j = 0; j < k; ++j) a2 = a2 + AT(i, j) * tmp[j];
        AT(i, k) = AT(i, k) - a2;
      }
    }
    const float akk = AT(k, k);
    const bool nz = fabsf(akk) > 1.17549435e-38f;
    const float r0 = vo_rcp_seed(akk);
    rinv[k] = nz ? r0 + r0 * (1.f - akk * r0) : 0.f;  // one Newton step: <= 1 ulp
#pragma unroll
    for (int i = k + 1; i < N; ++i) AT(i, k) = nz ? AT(i, k) * rinv[k] : AT(i, k);
  }
  float y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = rhs[i];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int c = k + 1; c < N; ++c) VO_CSWAP(tr[k] == c, y[k], y[c]);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) y[i] = y[i] - AT(i, j) * y[j];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = y[i] * rinv[i];
#pragma unroll
  for (int i = N - 1; i >= 0; --i)
#pragma unroll
    for (int j = i + 1; j < N; ++j) y[i] = y[i] - AT(j, i) * y[j];
#pragma unroll
  for (int k = N - 1; k >= 0; --k)
#pragma unroll
    for (int c = k + 1; c < N; ++c) VO_CSWAP(tr[k] == c, y[k], y[c]);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = y[i];
#undef VO_CSWAP
#undef AT
}

VO_HD void mat3_mul_dev(const float* A, const float* B, float* C) {
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i)
      C[j * 3 + i] = (A[i] * B[j * 3] + A[3 + i] * B[j * 3 + 1]) + A[6 + i] * B[j * 3 + 2];
}

}  // namespace vo
