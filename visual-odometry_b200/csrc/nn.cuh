// nn.cuh — pieces shared by the two nearest-neighbour filters (nn.cu: FP32 FFMA2 partial-distance
// filter; nn_tc.cu: tcgen05 f16 tensor-core filter): the reference's distance in its own rounding
// order, the result keys, the soundness margin of the FMA filters and the handle.
#pragma once
#include <float.h>
#include <math.h>

#include <map>
#include <vector>

#include "common.cuh"

namespace vo {

constexpr int NN_DIM = 10;            // fast-path dimension (Vector11f minus the id column)
constexpr int NN_FDIM = 5;            // leading dimensions the streaming filter looks at (<= 6)
constexpr int NN_TM = 128;            // map rows per shared-memory tile
constexpr int NN_STAGES = 4;          // TMA stages in flight
constexpr int NN_ROW_BYTES = 48;      // packed row: 3 x float4
constexpr uint32_t NN_TILE_BYTES = NN_TM * NN_ROW_BYTES;
constexpr unsigned long long NN_KEY_NONE = 0xFFFFFFFFFFFFFFFFull;

// ---- the reference's distance, one rounding per operation ---------------------------------
// (p-q).tail(n).squaredNorm() under Eigen's SSE2 linear-vectorised redux; see oracle_sqdist.
template <int DIM>
__device__ __forceinline__ float ref_sqdist(const float (&m)[DIM], const float (&q)[DIM]) {
  float s[DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const float d = __fsub_rn(m[i], q[i]);
    s[i] = __fmul_rn(d, d);
  }
  constexpr int n4 = (DIM / 4) * 4, n8 = (DIM / 8) * 8;
  float r;
  if (n4 > 0) {
    float a[4] = {s[0], s[1], s[2], s[3]};
    if (n4 > 4) {
      float c[4] = {s[4], s[5], s[6], s[7]};
#pragma unroll
      for (int i = 8; i < n8; i += 8)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          a[l] = __fadd_rn(a[l], s[i + l]);
          c[l] = __fadd_rn(c[l], s[i + 4 + l]);
        }
#pragma unroll
      for (int l = 0; l < 4; ++l) a[l] = __fadd_rn(a[l], c[l]);
      if (n4 > n8)
#pragma unroll
        for (int l = 0; l < 4; ++l) a[l] = __fadd_rn(a[l], s[n8 + l]);
    }
    r = __fadd_rn(__fadd_rn(a[0], a[2]), __fadd_rn(a[1], a[3]));
#pragma unroll
    for (int i = n4; i < DIM; ++i) r = __fadd_rn(r, s[i]);
  } else {
    r = s[0];
#pragma unroll
    for (int i = 1; i < DIM; ++i) r = __fadd_rn(r, s[i]);
  }
  return r;
}

// run-time dimension version (general kernel)
__device__ __forceinline__ float ref_sqdist_dyn(const float* __restrict__ m,
                                                const float* __restrict__ q, int dim) {
  const int n4 = (dim / 4) * 4, n8 = (dim / 8) * 8;
  auto sq = [&](int i) {
    const float d = __fsub_rn(m[i], q[i]);
    return __fmul_rn(d, d);
  };
  float r;
  if (n4 > 0) {
    float a0 = sq(0), a1 = sq(1), a2 = sq(2), a3 = sq(3);
    if (n4 > 4) {
      float c0 = sq(4), c1 = sq(5), c2 = sq(6), c3 = sq(7);
      for (int i = 8; i < n8; i += 8) {
        a0 = __fadd_rn(a0, sq(i));
        a1 = __fadd_rn(a1, sq(i + 1));
        a2 = __fadd_rn(a2, sq(i + 2));
        a3 = __fadd_rn(a3, sq(i + 3));
        c0 = __fadd_rn(c0, sq(i + 4));
        c1 = __fadd_rn(c1, sq(i + 5));
        c2 = __fadd_rn(c2, sq(i + 6));
        c3 = __fadd_rn(c3, sq(i + 7));
      }
      a0 = __fadd_rn(a0, c0);
      a1 = __fadd_rn(a1, c1);
      a2 = __fadd_rn(a2, c2);
      a3 = __fadd_rn(a3, c3);
      if (n4 > n8) {
        a0 = __fadd_rn(a0, sq(n8));
        a1 = __fadd_rn(a1, sq(n8 + 1));
        a2 = __fadd_rn(a2, sq(n8 + 2));
        a3 = __fadd_rn(a3, sq(n8 + 3));
      }
    }
    r = __fadd_rn(__fadd_rn(a0, a2), __fadd_rn(a1, a3));
    for (int i = n4; i < dim; ++i) r = __fadd_rn(r, sq(i));
  } else {
    r = sq(0);
    for (int i = 1; i < dim; ++i) r = __fadd_rn(r, sq(i));
  }
  return r;
}

__device__ __forceinline__ unsigned long long nn_pack_key(float d2, uint32_t row) {
  // d2 >= 0, so its bit pattern is monotone as an unsigned integer
  return (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) | row;
}

// Filter thresholds for one query.  With u = 2^-24, d2_ref < bound implies
//   full:     acc10 = |m|^2   + sum_{k<10} (-2 q_k) m_k  <  (bound - |q|^2)   + eps
//   partial:  acc6  = |m|^2_6 + sum_{k<6}  (-2 q_k) m_k  <  (bound - |q|^2_6) + eps
// (the partial squared distance over the first NN_FDIM dimensions is a lower bound of the full
// one), where eps = 64u(|q|^2 + max|m|^2) + 32u|bound| covers the FMA chains and the roundings of
// |m|^2, |q|^2 and of the reference's own d2.  `qn` is the query scaled by -2.
__device__ __forceinline__ float nn_eps(float qq, float bound, float mm_max) {
  const float u64 = 64.f * 5.9604645e-8f;  // 64 * 2^-24
  return u64 * (qq + mm_max) + 0.5f * u64 * fabsf(bound);
}
__device__ __forceinline__ void nn_query_norms(const float (&qn)[NN_DIM], float* qq6, float* qq) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NN_FDIM; ++k) {
    const float q = -0.5f * qn[k];
    s = fmaf(q, q, s);
  }
  *qq6 = s;
#pragma unroll
  for (int k = NN_FDIM; k < NN_DIM; ++k) {
    const float q = -0.5f * qn[k];
    s = fmaf(q, q, s);
  }
  *qq = s;
}
__device__ __forceinline__ float nn_threshold_partial(float qq6, float qq, float bound, float mm_max) {
  return (bound - qq6) + nn_eps(qq, bound, mm_max);
}
__device__ __forceinline__ float nn_threshold_full(float qq, float bound, float mm_max) {
  return (bound - qq) + nn_eps(qq, bound, mm_max);
}

}  // namespace vo

struct vo_nn_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t n_rows = 0;
  int row_stride = 0, skip = 0, dim = 0;
  bool fast = false;          // packed layout valid (dim == NN_DIM)
  int64_t n_tiles = 0;
  vo::DevBuf raw;                 // staging copy of the caller's rows (host variant / general path)
  // rows as seen by the general + radius kernels: the packed buffer (stride 12, skip 0) on the
  // fast path, the private raw copy otherwise
  const float* rows_dev = nullptr;
  int map_stride = 0, map_skip = 0;
  vo::DevBuf packed;
  vo::DevBuf scalars;             // [0] = mm_max
  vo::DevBuf keys;
  vo::DevBuf q_stage, idx_stage, d2_stage, cnt_stage, list_stage;
  std::map<int, int> occupancy;  // (TQ, THREADS) -> resident CTAs per SM of that filter variant
  // filter launches of the last best_match call: (TQ, THREADS, query tiles, map splits) each, so a
  // test can PROVE which instantiation answered it (vo_nn_last_launches)
  std::vector<int32_t> last_launches;
  // tensor-core filter (nn_tc.cu): f16 operand tiles of the map, built at set_map when the map is
  // large enough to be worth it and its norms fit f16 arithmetic
  vo::DevBuf tiles16, tc_stats;
  int64_t n_tiles16 = 0;
  bool tc_ready = false, tc_opted_in = false, last_was_tc = false;
  // host copy of max|m|^2.  It only steers the CHOICE of the filter (both are exact for any data), so
  // it may be one map old: every set_map sends the new value to pinned memory without waiting, and
  // the first query that finds it arrived adopts it; only a handle that has never seen a value waits.
  float mm_max_host = 0.f;
  bool have_mm_max = false, mm_pending = false;
  float* mm_pinned = nullptr;
  cudaEvent_t mm_event = nullptr;
  void* tiny_host = nullptr;  // mapped pinned memory for the answers of the few-queries path
  void* tiny_dev = nullptr;
  unsigned int tiny_seq = 0;  // sequence number of the last few-queries call (flag value the host waits for)
  int force_path = 0;  // VO_NN_FORCE_PATH: 0 auto, 1 ffma (FP32 CUDA cores), 2 tc (tensor cores)
};

// nn_tc.cu
int nn_tc_pack(vo_nn_s* h);
int nn_tc_launch(vo_nn_s* h, const float* queries_dev, int64_t nq, int qstride, float bound);
// When the tensor-core filter answers (tools/nn_crossover.py, set_map + best_match, us FP32 / us
// tensor: 8192 x 4096 29 / 32, 12000 x 4096 36 / 32, 8192 x 9000 40 / 38, 16384 x 9000 63 / 43,
// 32768 x 30000 272 / 100, 2048 x 30000 52 / 46): from about 4e7 (query, row) pairs.  The f16 tiles
// are built at set_map for maps of NN_TC_EAGER_ROWS rows, otherwise by the first batch that wants them.
constexpr int64_t NN_TC_MIN_ROWS = 2048;
constexpr int64_t NN_TC_MIN_QUERIES = 512;
constexpr int64_t NN_TC_MIN_PAIRS = 40000000;
constexpr int64_t NN_TC_EAGER_ROWS = 32768;
inline bool nn_tc_worthwhile(int64_t rows, int64_t nq) {
  return rows >= NN_TC_MIN_ROWS && nq >= NN_TC_MIN_QUERIES && rows >= (NN_TC_MIN_PAIRS + nq - 1) / nq;
}

