// triangulate.cu — two-view midpoint triangulation and batch projection with order-preserving
// compaction, for sm_100a.
//
// Replaces triangulate_point / triangulate_points x3 (reference src/utils.cpp:36-134) and
// Camera::projectPoints (src/camera.cpp:16-37).
//
// Both are HBM-bound elementwise maps followed by the reference's sequential "n_success"
// compaction.  Each is ONE kernel: a tile of 1024 correspondences is loaded with coalesced 8-byte
// pair reads + gathered 8-byte image points, solved in registers, ranked with warp ballots, and
// the tile's output offset is the sum of the two-level status words of everything before it
// (scan.cuh) — 44 algorithmic bytes per correspondence (8 pair + 8 + 8 points in, 12 point +
// 8 pair out), touched once.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "linalg.cuh"
#include "scan.cuh"

namespace vo {

constexpr int TRI_THREADS = 256;
constexpr int TRI_ITEMS = 4;
constexpr int TRI_TILE = TRI_THREADS * TRI_ITEMS;

struct TriParams {
  float iK[9];    // K^-1                      (utils.cpp:54)
  float iRiK[9];  // X^-1.linear() * K^-1      (utils.cpp:55)
  float t[3];     // X^-1.translation()        (utils.cpp:56)
  const int2* __restrict__ corr;
  int64_t n_corr;
  const float2* __restrict__ p1;
  const float2* __restrict__ p2;
  const float* __restrict__ app2;  // nullable, 10 floats / point
  float* out_points;               // 3 floats / success
  int2* out_corr_new;              // nullable
  float* out_app;                  // nullable
  int32_t* out_src;                // nullable
  long long* n_success;            // device int64
  ScanWorkspace ws;
  int num_tiles;
};

// triangulate_point, utils.cpp:36-49
__device__ __forceinline__ bool triangulate_point_dev(const float (&d1)[3], const float (&d2)[3],
                                                      const float (&t)[3], float (&p)[3]) {
  const float n0 = -d1[0], n1 = -d1[1], n2 = -d1[2];  // D.col(0) = -d1
  float A[4], rhs[2], ss[2];
  // fixed-size Eigen products sum three terms as s0 + (s1 + s2) (see oracle/vo_oracle.c)
  A[0] = n0 * n0 + (n1 * n1 + n2 * n2);  // D^T D
  A[1] = d2[0] * n0 + (d2[1] * n1 + d2[2] * n2);
  A[2] = A[1];
  A[3] = d2[0] * d2[0] + (d2[1] * d2[1] + d2[2] * d2[2]);
  rhs[0] = n0 * t[0] + (n1 * t[1] + n2 * t[2]);  // D^T p2
  rhs[1] = d2[0] * t[0] + (d2[1] * t[1] + d2[2] * t[2]);
  ldlt_solve_dev<2>(A, rhs, ss);  // :40
  const float s0 = -ss[0], s1 = -ss[1];
  if (s0 < 0.f || s1 < 0.f) return false;  // :41
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = 0.5f * (s0 * d1[i] + (t[i] + s1 * d2[i]));  // :44-47
  return true;
}

// One tile of THREADS*ITEMS correspondences per block.  Item j of lane l of warp w sits at tile
// position w*32*ITEMS + j*32 + l, so every load and store of a warp is a contiguous run.
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) triangulate_kernel(const TriParams q) {
  constexpr int TILE = THREADS * ITEMS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = scan_take_ticket(q.ws);
  const int64_t warp_base = (int64_t)tile * TILE + (int64_t)warp * (32 * ITEMS);

  int2 c[ITEMS];
  bool in[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t i = warp_base + j * 32 + lane;
    in[j] = i < q.n_corr;
    c[j] = in[j] ? __ldg(q.corr + i) : make_int2(0, 0);
  }
  float2 a[ITEMS], b[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    if (in[j]) {
      a[j] = __ldg(q.p1 + c[j].x);  // .first  -> image 1   (utils.cpp:87)
      b[j] = __ldg(q.p2 + c[j].y);  // .second -> image 2   (utils.cpp:88)
    }
  float P[ITEMS][3];
  bool ok[ITEMS];
  const float t[3] = {q.t[0], q.t[1], q.t[2]};
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    ok[j] = false;
    if (in[j]) {
      float d1[3], d2[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        d1[i] = q.iK[i] * a[j].x + (q.iK[3 + i] * a[j].y + q.iK[6 + i]);        // iK*[p1;1]   :91
        d2[i] = q.iRiK[i] * b[j].x + (q.iRiK[3 + i] * b[j].y + q.iRiK[6 + i]);  // iRiK*[p2;1] :94
      }
      ok[j] = triangulate_point_dev(d1, d2, t, P[j]);
    }
  }
  int local[ITEMS];
  int total;
  scan_tile_post<THREADS, ITEMS>(q.ws, tile, ok, local, &total);
  const long long excl = scan_tile_lookback<THREADS>(q.ws, tile);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    if (ok[j]) {
      const long long k = excl + local[j];
      q.out_points[3 * k + 0] = P[j][0];
      q.out_points[3 * k + 1] = P[j][1];
      q.out_points[3 * k + 2] = P[j][2];
      if (q.out_corr_new) q.out_corr_new[k] = make_int2(c[j].y, (int)k);  // (idx_second, k) :97
      if (q.out_src) q.out_src[k] = (int32_t)(warp_base + j * 32 + lane);
      if (q.out_app) {  // :127 — the appearance travels with the point
        const float2* src = reinterpret_cast<const float2*>(q.app2 + 10 * (int64_t)c[j].y);
        float2* dst = reinterpret_cast<float2*>(q.out_app + 10 * k);
#pragma unroll
        for (int i = 0; i < 5; ++i) dst[i] = __ldg(src + i);
      }
    }
  if (tile == q.num_tiles - 1 && threadIdx.x == 0) *q.n_success = excl + total;
}

// ---- Camera::projectPoints -----------------------------------------------------------------------
struct ProjParams {
  float T[12];  // world-in-camera: 3x3 linear (col-major) then translation
  float K[9];
  float z_near, z_far, max_u, max_v;
  const float* __restrict__ world;
  int64_t n;
  int keep_indices;
  float2* out;
  long long* counts;  // [0] = n_out, [1] = n_inside
  ScanWorkspace ws;
  int num_tiles;
};

// Camera::projectPoint, camera.h:25-37
__device__ __forceinline__ bool project_point_dev(const ProjParams& q, float wx, float wy, float wz,
                                                  float2* uv) {
  const float px = q.T[9] + (q.T[0] * wx + (q.T[3] * wy + q.T[6] * wz));
  const float py = q.T[10] + (q.T[1] * wx + (q.T[4] * wy + q.T[7] * wz));
  const float pz = q.T[11] + (q.T[2] * wx + (q.T[5] * wy + q.T[8] * wz));
  if (pz > q.z_far || pz < q.z_near) return false;
  const float hx = q.K[0] * px + (q.K[3] * py + q.K[6] * pz);
  const float hy = q.K[1] * px + (q.K[4] * py + q.K[7] * pz);
  const float hz = q.K[2] * px + (q.K[5] * py + q.K[8] * pz);
  const float iz = (float)(1.0 / (double)hz);  // camera.h:31: 1./z is a double, then demoted
  uv->x = hx * iz;
  uv->y = hy * iz;
  if (uv->x < 0.f || uv->x > q.max_u) return false;
  if (uv->y < 0.f || uv->y > q.max_v) return false;
  return true;
}

__global__ void __launch_bounds__(TRI_THREADS) project_points_kernel(const ProjParams q) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = scan_take_ticket(q.ws);
  const int64_t warp_base = (int64_t)tile * TRI_TILE + (int64_t)warp * (32 * TRI_ITEMS);
  float2 uv[TRI_ITEMS];
  bool ok[TRI_ITEMS], in[TRI_ITEMS];
#pragma unroll
  for (int j = 0; j < TRI_ITEMS; ++j) {
    const int64_t i = warp_base + j * 32 + lane;
    in[j] = i < q.n;
    ok[j] = false;
    if (in[j]) {
      const float* w = q.world + 3 * i;
      ok[j] = project_point_dev(q, __ldg(w), __ldg(w + 1), __ldg(w + 2), &uv[j]);
      if (!ok[j]) uv[j] = make_float2(-1.f, -1.f);  // camera.cpp:21,30
    }
  }
  int local[TRI_ITEMS];
  int total;
  scan_tile_post<TRI_THREADS, TRI_ITEMS>(q.ws, tile, ok, local, &total);
  const long long excl = scan_tile_lookback<TRI_THREADS>(q.ws, tile);
#pragma unroll
  for (int j = 0; j < TRI_ITEMS; ++j) {
    if (!in[j]) continue;
    if (q.keep_indices) q.out[warp_base + j * 32 + lane] = uv[j];
    else if (ok[j]) q.out[excl + local[j]] = uv[j];
  }
  if (tile == q.num_tiles - 1 && threadIdx.x == 0) {
    q.counts[1] = excl + total;
    q.counts[0] = q.keep_indices ? (long long)q.n : excl + total;
  }
}

// ---- host-side precomputation (one rounding per operation, like the reference build) ----------
static void h_mat3_vec(const float* M, const float v[3], float out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = M[i] * v[0] + (M[3 + i] * v[1] + M[6 + i] * v[2]);
}
static void h_mat3_mul(const float* A, const float* B, float* C) {
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i)
      C[j * 3 + i] = A[i] * B[j * 3] + (A[3 + i] * B[j * 3 + 1] + A[6 + i] * B[j * 3 + 2]);
}
// cofactor inverse, the fixed-size path behind Matrix3f::inverse(): cyclic cofactors, determinant
// expanded along column 0, result(i,j) = cofactor(j,i) / det
static float h_cof3(const float* M, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return M[j1 * 3 + i1] * M[j2 * 3 + i2] - M[j2 * 3 + i1] * M[j1 * 3 + i2];
}
static void h_mat3_inverse(const float* M, float* out) {
  const float c0 = h_cof3(M, 0, 0), c1 = h_cof3(M, 1, 0), c2 = h_cof3(M, 2, 0);
  const float det = c0 * M[0] + (c1 * M[1] + c2 * M[2]);
  const float id = 1.f / det;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      out[j * 3 + i] = (i == 0 ? (j == 0 ? c0 : (j == 1 ? c1 : c2)) : h_cof3(M, j, i)) * id;
}

static void tri_precompute(const float K[9], const float X[16], TriParams* q) {
  // iX = X.inverse() for an isometry: (R^T, -R^T t)   utils.cpp:53,56
  float iR[9], r[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) iR[j * 3 + i] = X[i * 4 + j];
  const float tx[3] = {X[12], X[13], X[14]};
  h_mat3_vec(iR, tx, r);
  q->t[0] = -r[0];
  q->t[1] = -r[1];
  q->t[2] = -r[2];
  h_mat3_inverse(K, q->iK);
  h_mat3_mul(iR, q->iK, q->iRiK);
}

static int tri_launch(cudaStream_t stream, const float K[9], const float X[16], const int32_t* corr,
                      int64_t n, const float* p1, const float* p2, const float* app2, float* out_pts,
                      int32_t* out_corr_new, float* out_app, int32_t* out_src, int64_t* n_success,
                      void* workspace) {
  const int tile_items = TRI_TILE;
  const int64_t tiles = (n + tile_items - 1) / tile_items;
  VO_REQUIRE(tiles < (1LL << 31), VO_ERR_UNSUPPORTED, "too many correspondences");
  if (n == 0) {
    VO_CUDA(cudaMemsetAsync(n_success, 0, sizeof(int64_t), stream));
    return VO_OK;
  }
  VO_CUDA(cudaMemsetAsync(workspace, 0, (size_t)scan_workspace_bytes(tiles), stream));
  TriParams q;
  tri_precompute(K, X, &q);
  q.corr = reinterpret_cast<const int2*>(corr);
  q.n_corr = n;
  q.p1 = reinterpret_cast<const float2*>(p1);
  q.p2 = reinterpret_cast<const float2*>(p2);
  q.app2 = app2;
  q.out_points = out_pts;
  q.out_corr_new = reinterpret_cast<int2*>(out_corr_new);
  q.out_app = (app2 != nullptr) ? out_app : nullptr;
  q.out_src = out_src;
  q.n_success = reinterpret_cast<long long*>(n_success);
  q.ws = scan_workspace_at(workspace, tiles);
  q.num_tiles = (int)tiles;
  triangulate_kernel<TRI_THREADS, TRI_ITEMS><<<(unsigned)tiles, TRI_THREADS, 0, stream>>>(q);
  VO_LAUNCH_CHECK();
  return VO_OK;
}

// per-device scratch for the host-pointer entry points
struct HostCtx {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  DevBuf in_a, in_b, in_c, in_d, out_a, out_b, out_c, out_d, ws, cnt;
};
static HostCtx* host_ctx(int device) {
  static HostCtx ctx[64];
  if (device < 0 || device >= 64) return nullptr;
  return &ctx[device];
}

}  // namespace vo

using namespace vo;

extern "C" {

int64_t vo_triangulate_workspace_bytes(int64_t n_corr) {
  if (n_corr < 0) return VO_ERR_ARG;
  return scan_workspace_bytes((n_corr + TRI_TILE - 1) / TRI_TILE);
}

int vo_triangulate_device(void* cuda_stream, const float K[9], const float X[16],
                          const int32_t* corr_dev, int64_t n_corr, const float* p1_dev,
                          const float* p2_dev, const float* app2_dev, float* out_points_dev,
                          int32_t* out_corr_new_dev, float* out_app_dev, int32_t* out_src_dev,
                          int64_t* n_success_dev, void* workspace_dev) {
  VO_REQUIRE(K && X && n_success_dev, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(n_corr >= 0, VO_ERR_ARG, "negative size");
  VO_REQUIRE(n_corr == 0 || (corr_dev && p1_dev && p2_dev && out_points_dev && workspace_dev),
             VO_ERR_ARG, "null pointer");
  return tri_launch(static_cast<cudaStream_t>(cuda_stream), K, X, corr_dev, n_corr, p1_dev, p2_dev,
                    app2_dev, out_points_dev, out_corr_new_dev, out_app_dev, out_src_dev,
                    n_success_dev, workspace_dev);
}

int vo_triangulate(int device, const float K[9], const float X[16], const int32_t* corr_host,
                   int64_t n_corr, const float* p1_host, int64_t n_p1, const float* p2_host,
                   int64_t n_p2, const float* app2_host, float* out_points_host,
                   int32_t* out_corr_new_host, float* out_app_host, int32_t* out_src_host,
                   int64_t* n_success) {
  VO_REQUIRE(K && X && n_success, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(n_corr >= 0 && n_p1 >= 0 && n_p2 >= 0, VO_ERR_ARG, "negative size");
  *n_success = 0;
  if (n_corr == 0) return VO_OK;
  VO_REQUIRE(corr_host && p1_host && p2_host && out_points_host, VO_ERR_ARG, "null pointer");
  // the PointCloud overload indexes with .at() (utils.cpp:119-127): out-of-range is an error,
  // never a device fault
  for (int64_t i = 0; i < n_corr; ++i) {
    const int32_t a = corr_host[2 * i], b = corr_host[2 * i + 1];
    if (a < 0 || a >= n_p1 || b < 0 || b >= n_p2) {
      set_error("vo_triangulate: correspondence %lld = (%d,%d) out of range", (long long)i, a, b);
      return VO_ERR_ARG;
    }
  }
  int ndev = 0;
  VO_CUDA(cudaGetDeviceCount(&ndev));
  VO_REQUIRE(device >= 0 && device < ndev && device < 64, VO_ERR_ARG, "bad device ordinal");
  DeviceGuard g(device);
  HostCtx* cx = host_ctx(device);
  std::lock_guard<std::mutex> lock(cx->mu);
  if (!cx->stream) VO_CUDA(cudaStreamCreateWithFlags(&cx->stream, cudaStreamNonBlocking));
  cudaStream_t s = cx->stream;
  const bool with_app = app2_host != nullptr && out_app_host != nullptr;
  int rc;
  if ((rc = cx->in_a.reserve((size_t)n_corr * 8))) return rc;
  if ((rc = cx->in_b.reserve((size_t)n_p1 * 8 + 16))) return rc;
  if ((rc = cx->in_c.reserve((size_t)n_p2 * 8 + 16))) return rc;
  if (with_app && (rc = cx->in_d.reserve((size_t)n_p2 * 40 + 16))) return rc;
  if ((rc = cx->out_a.reserve((size_t)n_corr * 12))) return rc;
  if (out_corr_new_host && (rc = cx->out_b.reserve((size_t)n_corr * 8))) return rc;
  if (with_app && (rc = cx->out_c.reserve((size_t)n_corr * 40))) return rc;
  if (out_src_host && (rc = cx->out_d.reserve((size_t)n_corr * 4))) return rc;
  if ((rc = cx->ws.reserve((size_t)vo_triangulate_workspace_bytes(n_corr)))) return rc;
  if ((rc = cx->cnt.reserve(64))) return rc;
  VO_CUDA(cudaMemcpyAsync(cx->in_a.p, corr_host, (size_t)n_corr * 8, cudaMemcpyHostToDevice, s));
  VO_CUDA(cudaMemcpyAsync(cx->in_b.p, p1_host, (size_t)n_p1 * 8, cudaMemcpyHostToDevice, s));
  VO_CUDA(cudaMemcpyAsync(cx->in_c.p, p2_host, (size_t)n_p2 * 8, cudaMemcpyHostToDevice, s));
  if (with_app)
    VO_CUDA(cudaMemcpyAsync(cx->in_d.p, app2_host, (size_t)n_p2 * 40, cudaMemcpyHostToDevice, s));
  rc = tri_launch(s, K, X, cx->in_a.as<int32_t>(), n_corr, cx->in_b.as<float>(),
                  cx->in_c.as<float>(), with_app ? cx->in_d.as<float>() : nullptr,
                  cx->out_a.as<float>(), out_corr_new_host ? cx->out_b.as<int32_t>() : nullptr,
                  with_app ? cx->out_c.as<float>() : nullptr,
                  out_src_host ? cx->out_d.as<int32_t>() : nullptr, cx->cnt.as<int64_t>(),
                  cx->ws.p);
  if (rc) return rc;
  int64_t ns = 0;
  VO_CUDA(cudaMemcpyAsync(&ns, cx->cnt.p, sizeof(ns), cudaMemcpyDeviceToHost, s));
  VO_CUDA(cudaStreamSynchronize(s));
  if (ns > 0) {
    VO_CUDA(cudaMemcpyAsync(out_points_host, cx->out_a.p, (size_t)ns * 12, cudaMemcpyDeviceToHost, s));
    if (out_corr_new_host)
      VO_CUDA(cudaMemcpyAsync(out_corr_new_host, cx->out_b.p, (size_t)ns * 8,
                              cudaMemcpyDeviceToHost, s));
    if (with_app)
      VO_CUDA(cudaMemcpyAsync(out_app_host, cx->out_c.p, (size_t)ns * 40, cudaMemcpyDeviceToHost, s));
    if (out_src_host)
      VO_CUDA(cudaMemcpyAsync(out_src_host, cx->out_d.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
  }
  *n_success = ns;
  return VO_OK;
}

int vo_project_points(int device, const vo_camera* cam, const float* world_host, int64_t n_points,
                      int keep_indices, float* out_image_host, int64_t* n_out, int64_t* n_inside) {
  VO_REQUIRE(cam && n_out && n_inside, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(n_points >= 0, VO_ERR_ARG, "negative size");
  *n_out = 0;
  *n_inside = 0;
  if (n_points == 0) return VO_OK;
  VO_REQUIRE(world_host && out_image_host, VO_ERR_ARG, "null pointer");
  int ndev = 0;
  VO_CUDA(cudaGetDeviceCount(&ndev));
  VO_REQUIRE(device >= 0 && device < ndev && device < 64, VO_ERR_ARG, "bad device ordinal");
  DeviceGuard g(device);
  HostCtx* cx = host_ctx(device);
  std::lock_guard<std::mutex> lock(cx->mu);
  if (!cx->stream) VO_CUDA(cudaStreamCreateWithFlags(&cx->stream, cudaStreamNonBlocking));
  cudaStream_t s = cx->stream;
  const int64_t tiles = (n_points + TRI_TILE - 1) / TRI_TILE;
  VO_REQUIRE(tiles < (1LL << 31), VO_ERR_UNSUPPORTED, "too many points");
  int rc;
  if ((rc = cx->in_a.reserve((size_t)n_points * 12))) return rc;
  if ((rc = cx->out_a.reserve((size_t)n_points * 12))) return rc;
  if ((rc = cx->ws.reserve((size_t)scan_workspace_bytes(tiles)))) return rc;
  if ((rc = cx->cnt.reserve(64))) return rc;
  VO_CUDA(cudaMemcpyAsync(cx->in_a.p, world_host, (size_t)n_points * 12, cudaMemcpyHostToDevice, s));
  VO_CUDA(cudaMemsetAsync(cx->ws.p, 0, (size_t)scan_workspace_bytes(tiles), s));
  ProjParams q;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 3; ++i) q.T[j * 3 + i] = cam->T[j * 4 + i];
  memcpy(q.K, cam->K, sizeof(q.K));
  q.z_near = (float)cam->z_near;
  q.z_far = (float)cam->z_far;
  q.max_u = (float)(cam->cols - 1);
  q.max_v = (float)(cam->rows - 1);
  q.world = cx->in_a.as<float>();
  q.n = n_points;
  q.keep_indices = keep_indices ? 1 : 0;
  q.out = cx->out_a.as<float2>();
  q.counts = cx->cnt.as<long long>();
  q.ws = scan_workspace_at(cx->ws.p, tiles);
  q.num_tiles = (int)tiles;
  project_points_kernel<<<(unsigned)tiles, TRI_THREADS, 0, s>>>(q);
  VO_LAUNCH_CHECK();
  long long counts[2] = {0, 0};
  VO_CUDA(cudaMemcpyAsync(counts, cx->cnt.p, sizeof(counts), cudaMemcpyDeviceToHost, s));
  VO_CUDA(cudaStreamSynchronize(s));
  if (counts[0] > 0) {
    VO_CUDA(cudaMemcpyAsync(out_image_host, cx->out_a.p, (size_t)counts[0] * 8,
                            cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
  }
  *n_out = counts[0];
  *n_inside = counts[1];
  return VO_OK;
}

}  // extern "C"
