// linalg.cuh — tiny fixed-size device linear algebra shared by picp.cu and triangulate.cu.
#pragma once
#include "common.cuh"

namespace vo {

// ---- pivoted LDL^T solve, same algorithm as Eigen's LDLT (picp_solver.cpp:109) -------------
template <int N>
__device__ void ldlt_solve_dev(float* A /* N*N col-major, destroyed */, const float* rhs, float* x) {
  int tr[N];
  float tmp[N];
#define AT(i, j) A[(j) * N + (i)]
  for (int k = 0; k < N; ++k) {
    int piv = k;
    float big = fabsf(AT(k, k));
    for (int i = k + 1; i < N; ++i)
      if (fabsf(AT(i, i)) > big) {
        big = fabsf(AT(i, i));
        piv = i;
      }
    tr[k] = piv;
    if (piv != k) {
      for (int j = 0; j < k; ++j) {
        const float t = AT(k, j);
        AT(k, j) = AT(piv, j);
        AT(piv, j) = t;
      }
      for (int i = piv + 1; i < N; ++i) {
        const float t = AT(i, k);
        AT(i, k) = AT(i, piv);
        AT(i, piv) = t;
      }
      {
        const float t = AT(k, k);
        AT(k, k) = AT(piv, piv);
        AT(piv, piv) = t;
      }
      for (int i = k + 1; i < piv; ++i) {
        const float t = AT(i, k);
        AT(i, k) = AT(piv, i);
        AT(piv, i) = t;
      }
    }
    const int rs = N - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; ++j) tmp[j] = AT(j, j) * AT(k, j);
      float acc = 0.f;
      for (int j = 0; j < k; ++j) acc += AT(k, j) * tmp[j];
      AT(k, k) -= acc;
      for (int i = 0; i < rs; ++i) {
        float a2 = 0.f;
        for (int j = 0; j < k; ++j) a2 += AT(k + 1 + i, j) * tmp[j];
        AT(k + 1 + i, k) -= a2;
      }
    }
    const float akk = AT(k, k);
    if (rs > 0 && fabsf(akk) > 0.f)
      for (int i = 0; i < rs; ++i) AT(k + 1 + i, k) = AT(k + 1 + i, k) / akk;
  }
  float y[N];
  for (int i = 0; i < N; ++i) y[i] = rhs[i];
  for (int k = 0; k < N; ++k)
    if (tr[k] != k) {
      const float t = y[k];
      y[k] = y[tr[k]];
      y[tr[k]] = t;
    }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < i; ++j) y[i] -= AT(i, j) * y[j];
  for (int i = 0; i < N; ++i) {
    if (fabsf(AT(i, i)) > 1.17549435e-38f) y[i] = y[i] / AT(i, i);
    else y[i] = 0.f;
  }
  for (int i = N - 1; i >= 0; --i)
    for (int j = i + 1; j < N; ++j) y[i] -= AT(j, i) * y[j];
  for (int k = N - 1; k >= 0; --k)
    if (tr[k] != k) {
      const float t = y[k];
      y[k] = y[tr[k]];
      y[tr[k]] = t;
    }
  for (int i = 0; i < N; ++i) x[i] = y[i];
#undef AT
}

__device__ __forceinline__ void mat3_mul_dev(const float* A, const float* B, float* C) {
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i)
      C[j * 3 + i] = (A[i] * B[j * 3] + A[3 + i] * B[j * 3 + 1]) + A[6 + i] * B[j * 3 + 2];
}

}  // namespace vo
