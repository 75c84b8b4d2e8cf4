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

#include <algorithm>
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
  int64_t n_corr;                  // count, or its upper bound when n_corr_dev is set
  const int* n_corr_dev;           // frame pipeline: the count lives on the device
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
  // bounds of the two point sets (INT_MAX when the caller vouches for the indices) and where an
  // out-of-range pair is reported: *bad = lowest offending position (all ones = none).  The reference's
  // PointCloud overload throws through .at() (utils.cpp:119-127); here the item is dropped and the
  // host-pointer entry point turns the report into VO_ERR_ARG.
  int n_p1, n_p2;
  unsigned long long* bad;
};

// triangulate_point (utils.cpp:36-49) for TWO correspondences at once: the two lanes of sm_100a's
// packed FP32 pairs (mul/add.rn.f32x2: one issue slot, two IEEE-rounded results) carry two
// independent points.  No contraction anywhere on this path — every product (f2_prod) and every sum
// is rounded separately, in the reference's order — so each lane equals the CPU oracle (IEEE ==;
// the sign of an exact zero is the only thing that may differ, see f2_prod):
//   D = [-d1 d2];  A = D^T D;  rhs = D^T t;  ss = ldlt(A).solve(rhs);  s = -ss;
//   reject if s0 < 0 || s1 < 0;  p = 0.5 * (s0*d1 + (t + s1*d2)).
// The 2x2 pivoted LDL^T is ldlt_solve_dev<2> written out (same operation sequence as Eigen's:
// pivot on the larger |diagonal|, L = A10/D0, D1 = A11 - L*(D0*L), forward / diagonal / backward
// substitution, un-permute); its three divisions stay scalar IEEE divisions per lane.
typedef unsigned long long f2_t;

__device__ __forceinline__ f2_t f2_dot3(const f2_t (&a)[3], const f2_t (&b)[3]) {
  // fixed-size Eigen products sum three terms as s0 + (s1 + s2) (see oracle/vo_oracle.c)
  return f2_add(f2_prod(a[0], b[0]), f2_add(f2_prod(a[1], b[1]), f2_prod(a[2], b[2])));
}
__device__ __forceinline__ f2_t f2_select(bool c0, bool c1, f2_t a, f2_t b) {
  float a0, a1, b0, b1;
  f2_unpack(a, a0, a1);
  f2_unpack(b, b0, b1);
  return f2_pack(c0 ? a0 : b0, c1 ? a1 : b1);
}

__device__ __forceinline__ void triangulate_pair_dev(const f2_t (&d1)[3], const f2_t (&d2)[3],
                                                     const float (&t)[3], f2_t (&p)[3], bool& ok0,
                                                     bool& ok1) {
  const f2_t neg1 = f2_bc(-1.f);
  const f2_t tt[3] = {f2_bc(t[0]), f2_bc(t[1]), f2_bc(t[2])};
  // D.col(0) = -d1:  (-a)*(-b) == a*b and (-a)*b == -(a*b) exactly, so the sign is applied to the
  // finished dot products
  const f2_t A00 = f2_dot3(d1, d1);                 // n.n
  const f2_t A10 = f2_mul(f2_dot3(d2, d1), neg1);   // d2.n
  const f2_t A11 = f2_dot3(d2, d2);
  const f2_t r0 = f2_mul(f2_dot3(d1, tt), neg1);    // n.t
  const f2_t r1 = f2_dot3(d2, tt);
  float a00[2], a10[2], a11[2];
  f2_unpack(A00, a00[0], a00[1]);
  f2_unpack(A10, a10[0], a10[1]);
  f2_unpack(A11, a11[0], a11[1]);
  // pivot: the larger |diagonal| first (strict >, as in the generic routine)
  const bool sw0 = fabsf(a11[0]) > fabsf(a00[0]), sw1 = fabsf(a11[1]) > fabsf(a00[1]);
  const f2_t D0 = f2_select(sw0, sw1, A11, A00);
  f2_t D1 = f2_select(sw0, sw1, A00, A11);
  f2_t y0 = f2_select(sw0, sw1, r1, r0);
  f2_t y1 = f2_select(sw0, sw1, r0, r1);
  float d0[2];
  f2_unpack(D0, d0[0], d0[1]);
  const f2_t L = f2_pack(fabsf(d0[0]) > 0.f ? a10[0] / d0[0] : a10[0],
                         fabsf(d0[1]) > 0.f ? a10[1] / d0[1] : a10[1]);
  D1 = f2_add(D1, f2_mul(f2_prod(L, f2_prod(D0, L)), neg1));  // A11 - L*(D0*L)
  y1 = f2_add(y1, f2_mul(f2_prod(L, y0), neg1));              // forward substitution
  float d1v[2], y0v[2], y1v[2];
  f2_unpack(D1, d1v[0], d1v[1]);
  f2_unpack(y0, y0v[0], y0v[1]);
  f2_unpack(y1, y1v[0], y1v[1]);
  constexpr float tiny = 1.17549435e-38f;
  y0 = f2_pack(fabsf(d0[0]) > tiny ? y0v[0] / d0[0] : 0.f, fabsf(d0[1]) > tiny ? y0v[1] / d0[1] : 0.f);
  y1 = f2_pack(fabsf(d1v[0]) > tiny ? y1v[0] / d1v[0] : 0.f, fabsf(d1v[1]) > tiny ? y1v[1] / d1v[1] : 0.f);
  y0 = f2_add(y0, f2_mul(f2_prod(L, y1), neg1));              // backward substitution
  const f2_t ss0 = f2_select(sw0, sw1, y1, y0), ss1 = f2_select(sw0, sw1, y0, y1);
  const f2_t s0 = f2_mul(ss0, neg1), s1 = f2_mul(ss1, neg1);  // :40
  float s0v[2], s1v[2];
  f2_unpack(s0, s0v[0], s0v[1]);
  f2_unpack(s1, s1v[0], s1v[1]);
  ok0 = !(s0v[0] < 0.f || s1v[0] < 0.f);  // :41
  ok1 = !(s0v[1] < 0.f || s1v[1] < 0.f);
  const f2_t half = f2_bc(0.5f);
#pragma unroll
  for (int i = 0; i < 3; ++i)  // :44-47
    p[i] = f2_mul(half, f2_add(f2_prod(s0, d1[i]), f2_add(tt[i], f2_prod(s1, d2[i]))));
}

// Persistent blocks, deferred write-out.  An ordered compaction makes every tile depend on the
// success counts of ALL earlier tiles; when a block waits for that prefix right after computing
// its tile it idles for the slowest of several hundred in-flight predecessors (ncu: half of all
// warp samples sat on that barrier).  Here a block therefore
//   1. claims a tile (atomic ticket: it can only ever depend on tiles already claimed by running
//      blocks), loads and solves it, ranks the successes with warp ballots,
//   2. parks the compacted results (points + second index) in one of two shared-memory slots and
//      publishes the tile's count,
//   3. and only then RETIRES THE PREVIOUS tile: by now its predecessors have had a whole tile time
//      to publish, so the look-back (warp 0 only) rarely spins, and the parked results stream out
//      as fully coalesced stores.
// Item j of lane l of warp w sits at tile position w*32*ITEMS + j*32 + l, so every load of a warp
// is a contiguous run.
// Tried in round 2 and dropped (profiles/r02_notes.md): fetching the NEXT tile's index pairs with
// cp.async while the current tile is solved (tickets two ahead) — 0.101 ms against 0.092 for this
// version at 8.9e6 correspondences; block shapes 128x4, 512x4, 128x8, 256x2 — all slower.  ncu: 193
// executed instructions per correspondence at 57 % issue utilisation; the kernel is co-limited by
// issue and latency, not by bytes in flight.
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) triangulate_kernel(const TriParams q) {
  static_assert(ITEMS % 2 == 0, "items are processed in packed pairs");
  constexpr int TILE = THREADS * ITEMS;
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_pts[2][TILE * 3];
  __shared__ int s_c2[2][TILE];
  __shared__ int s_src[2][TILE];
  __shared__ int s_next;
  __shared__ int s_warp_tot[WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float t[3] = {q.t[0], q.t[1], q.t[2]};
  const int64_t n_corr = q.n_corr_dev ? min((int64_t)*q.n_corr_dev, q.n_corr) : q.n_corr;
  const int num_tiles = (int)((n_corr + TILE - 1) / TILE);
  if (n_corr == 0) {
    if (blockIdx.x == 0 && tid == 0) *q.n_success = 0;
    return;
  }

  // write-out of a parked tile: prefix of everything before it, then coalesced copies
  auto retire = [&](int slot, int tile, int total) {
    const long long excl = scan_tile_lookback_warp0<THREADS>(q.ws, tile);
    float* dst = q.out_points + 3 * excl;
    const int nf = 3 * total;
#pragma unroll
    for (int k = 0; k < 3 * ITEMS; ++k) {
      const int j = k * THREADS + tid;
      if (j < nf) dst[j] = s_pts[slot][j];
    }
    if (q.out_corr_new) {
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int r = k * THREADS + tid;
        if (r < total) q.out_corr_new[excl + r] = make_int2(s_c2[slot][r], (int)(excl + r));  // (idx_second, k) :97
      }
    }
    if (q.out_src)
      for (int r = tid; r < total; r += THREADS) q.out_src[excl + r] = s_src[slot][r];
    if (q.out_app)  // :127 — the appearance travels with the point
      for (int r = tid; r < total; r += THREADS) {
        const float2* src = reinterpret_cast<const float2*>(q.app2 + 10 * (int64_t)s_c2[slot][r]);
        float2* o = reinterpret_cast<float2*>(q.out_app + 10 * (excl + r));
#pragma unroll
        for (int i = 0; i < 5; ++i) o[i] = __ldg(src + i);
      }
    if (tile == num_tiles - 1 && tid == 0) *q.n_success = excl + total;
  };

  int pend_tile = -1, pend_total = 0, pend_slot = 0, slot = 0;
  if (tid == 0) s_next = (int)atomicAdd(q.ws.ticket, 1u);
  __syncthreads();
  int tile = s_next;
  while (tile < num_tiles) {
    const int64_t warp_base = (int64_t)tile * TILE + (int64_t)warp * (32 * ITEMS);
    // out-of-range slots of the last tile re-read the last correspondence and are masked out
    int2 c[ITEMS];
    bool in[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int64_t i = warp_base + j * 32 + lane;
      in[j] = i < n_corr;
      c[j] = __ldg(q.corr + (in[j] ? i : n_corr - 1));
      if ((unsigned)c[j].x >= (unsigned)q.n_p1 || (unsigned)c[j].y >= (unsigned)q.n_p2) {
        if (in[j] && q.bad) atomicMin(q.bad, (unsigned long long)i);
        in[j] = false;
        c[j] = make_int2(0, 0);
      }
    }
    float2 a[ITEMS], b[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      a[j] = __ldg(q.p1 + c[j].x);  // .first  -> image 1   (utils.cpp:87)
      b[j] = __ldg(q.p2 + c[j].y);  // .second -> image 2   (utils.cpp:88)
    }
    int claimed = 0;  // the tile after this one, claimed early so the atomic's latency is hidden too
    if (tid == 0) claimed = (int)atomicAdd(q.ws.ticket, 1u);
    f2_t P[ITEMS / 2][3];
    bool ok[ITEMS];
#pragma unroll
    for (int jp = 0; jp < ITEMS / 2; ++jp) {
      const f2_t ax = f2_pack(a[2 * jp].x, a[2 * jp + 1].x), ay = f2_pack(a[2 * jp].y, a[2 * jp + 1].y);
      const f2_t bx = f2_pack(b[2 * jp].x, b[2 * jp + 1].x), by = f2_pack(b[2 * jp].y, b[2 * jp + 1].y);
      f2_t d1[3], d2[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        // iK*[p1;1] (:91) and iRiK*[p2;1] (:94):  m0*x + (m1*y + m2)
        d1[i] = f2_add(f2_prod(f2_bc(q.iK[i]), ax), f2_add(f2_prod(f2_bc(q.iK[3 + i]), ay), f2_bc(q.iK[6 + i])));
        d2[i] = f2_add(f2_prod(f2_bc(q.iRiK[i]), bx), f2_add(f2_prod(f2_bc(q.iRiK[3 + i]), by), f2_bc(q.iRiK[6 + i])));
      }
      bool k0, k1;
      triangulate_pair_dev(d1, d2, t, P[jp], k0, k1);
      ok[2 * jp] = k0 && in[2 * jp];
      ok[2 * jp + 1] = k1 && in[2 * jp + 1];
    }

    // rank inside the tile: warp ballots, then the warp totals through shared memory
    int local[ITEMS];
    const unsigned lt = (1u << lane) - 1u;
    int run = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const unsigned bal = __ballot_sync(0xffffffffu, ok[j]);
      local[j] = run + __popc(bal & lt);
      run += __popc(bal);
    }
    if (lane == 0) s_warp_tot[warp] = run;
    if (tid == 0) s_next = claimed;
    __syncthreads();
    int warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const int x = s_warp_tot[w];
      warp_off += (w < warp) ? x : 0;
      total += x;
    }
    const int next = s_next;
    if (tid == 0) {
      scan_st(q.ws.status + tile, SCAN_POSTED | (unsigned long long)total);
      atomicAdd(q.ws.groups + tile / q.ws.group_tiles,
                (1ull << SCAN_GROUP_SHIFT) | (unsigned long long)total);
    }
    // park the compacted results of this tile
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if (ok[j]) {
        const int r = warp_off + local[j];
        float px, py, pz, ox, oy, oz;
        f2_unpack(P[j >> 1][0], px, ox);
        f2_unpack(P[j >> 1][1], py, oy);
        f2_unpack(P[j >> 1][2], pz, oz);
        s_pts[slot][3 * r + 0] = (j & 1) ? ox : px;
        s_pts[slot][3 * r + 1] = (j & 1) ? oy : py;
        s_pts[slot][3 * r + 2] = (j & 1) ? oz : pz;
        s_c2[slot][r] = c[j].y;
        s_src[slot][r] = (int)(warp_base + j * 32 + lane);
      }
    // retire the previous tile now that this one is published: its predecessors have had a whole
    // tile time to publish (measured: retiring it earlier, under this tile's load latency, brings
    // the waiting back).  The barrier inside also orders this tile's parking before its own
    // retirement in the next iteration.
    if (pend_tile >= 0) retire(pend_slot, pend_tile, pend_total);
    else __syncthreads();
    pend_tile = tile;
    pend_total = total;
    pend_slot = slot;
    slot ^= 1;
    tile = next;
  }
  if (pend_tile >= 0) {
    __syncthreads();
    retire(pend_slot, pend_tile, pend_total);
  }
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
                      void* workspace, const int32_t* n_corr_dev = nullptr, int64_t n_p1 = INT32_MAX,
                      int64_t n_p2 = INT32_MAX, unsigned long long* bad = nullptr) {
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
  q.n_corr_dev = n_corr_dev;
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
  q.n_p1 = (int)std::min<int64_t>(n_p1, INT32_MAX);
  q.n_p2 = (int)std::min<int64_t>(n_p2, INT32_MAX);
  q.bad = bad;
  // persistent blocks: as many as can be resident, each loops over dynamically claimed tiles
  static int resident = 0;
  if (resident == 0) {
    int per_sm = 1, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, triangulate_kernel<TRI_THREADS, TRI_ITEMS>,
                                                  TRI_THREADS, 0);
    resident = (per_sm < 1 ? 1 : per_sm) * (sms < 1 ? 1 : sms);
  }
  const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
  triangulate_kernel<TRI_THREADS, TRI_ITEMS><<<grid, TRI_THREADS, 0, stream>>>(q);
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

int vo_triangulate_device_ex(void* cuda_stream, const float K[9], const float X[16],
                             const int32_t* corr_dev, int64_t n_corr_max, const int32_t* n_corr_dev,
                             const float* p1_dev, const float* p2_dev, const float* app2_dev,
                             float* out_points_dev, int32_t* out_corr_new_dev, float* out_app_dev,
                             int64_t* n_success_dev, void* workspace_dev) {
  VO_REQUIRE(K && X && n_success_dev && n_corr_dev, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(n_corr_max >= 0, VO_ERR_ARG, "negative size");
  VO_REQUIRE(n_corr_max == 0 || (corr_dev && p1_dev && p2_dev && out_points_dev && workspace_dev),
             VO_ERR_ARG, "null pointer");
  return tri_launch(static_cast<cudaStream_t>(cuda_stream), K, X, corr_dev, n_corr_max, p1_dev, p2_dev,
                    app2_dev, out_points_dev, out_corr_new_dev, out_app_dev, nullptr, n_success_dev,
                    workspace_dev, n_corr_dev);
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
  // inputs through the pinned staging ring (the host copy of chunk k+1 overlaps the DMA of chunk k)
  if ((rc = stage_h2d(device, cx->in_a.p, corr_host, (size_t)n_corr * 8, s))) return rc;
  if ((rc = stage_h2d(device, cx->in_b.p, p1_host, (size_t)n_p1 * 8, s))) return rc;
  if ((rc = stage_h2d(device, cx->in_c.p, p2_host, (size_t)n_p2 * 8, s))) return rc;
  if (with_app && (rc = stage_h2d(device, cx->in_d.p, app2_host, (size_t)n_p2 * 40, s))) return rc;
  // cnt[0] = number of successes, cnt[1] = lowest out-of-range position (all ones = none): the
  // PointCloud overload indexes with .at() (utils.cpp:119-127), so out of range is an error, never
  // a device fault — checked by the kernel itself instead of an O(n) host loop
  VO_CUDA(cudaMemsetAsync(cx->cnt.p, 0xFF, 16, s));
  rc = tri_launch(s, K, X, cx->in_a.as<int32_t>(), n_corr, cx->in_b.as<float>(),
                  cx->in_c.as<float>(), with_app ? cx->in_d.as<float>() : nullptr,
                  cx->out_a.as<float>(), out_corr_new_host ? cx->out_b.as<int32_t>() : nullptr,
                  with_app ? cx->out_c.as<float>() : nullptr,
                  out_src_host ? cx->out_d.as<int32_t>() : nullptr, cx->cnt.as<int64_t>(),
                  cx->ws.p, nullptr, n_p1, n_p2, cx->cnt.as<unsigned long long>() + 1);
  if (rc) return rc;
  long long res[2] = {0, 0};
  VO_CUDA(cudaMemcpyAsync(res, cx->cnt.p, sizeof(res), cudaMemcpyDeviceToHost, s));
  VO_CUDA(cudaStreamSynchronize(s));
  if (res[1] != -1) {
    const long long i = res[1];
    set_error("vo_triangulate: correspondence %lld = (%d,%d) out of range", i, corr_host[2 * i],
              corr_host[2 * i + 1]);
    return VO_ERR_ARG;
  }
  const int64_t ns = res[0];
  if (ns > 0) {
    if ((rc = stage_d2h(device, out_points_host, cx->out_a.p, (size_t)ns * 12, s))) return rc;
    if (out_corr_new_host && (rc = stage_d2h(device, out_corr_new_host, cx->out_b.p, (size_t)ns * 8, s))) return rc;
    if (with_app && (rc = stage_d2h(device, out_app_host, cx->out_c.p, (size_t)ns * 40, s))) return rc;
    if (out_src_host && (rc = stage_d2h(device, out_src_host, cx->out_d.p, (size_t)ns * 4, s))) return rc;
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
  if ((rc = stage_h2d(device, cx->in_a.p, world_host, (size_t)n_points * 12, s))) return rc;
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
  if (counts[0] > 0 && (rc = stage_d2h(device, out_image_host, cx->out_a.p, (size_t)counts[0] * 8, s)))
    return rc;
  *n_out = counts[0];
  *n_inside = counts[1];
  return VO_OK;
}

}  // extern "C"
