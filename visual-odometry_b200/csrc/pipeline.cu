// pipeline.cu — the frame loop of vo_complete (reference src/apps/vo_complete.cpp:150-178) with every
// per-frame array RESIDENT ON THE DEVICE.
//
// The reference's loop body is:  associate consecutive frames by appearance (:151, :12-48),
// join with the previous triangulation (:152, :51-66), move the previous cloud into the previous
// camera frame (:154), 100 PICP rounds (:156-160), re-triangulate (:168-169) and merge the new
// points into the global map (:171, PointCloud.h:52-66).  Run through the drop-in classes this
// costs a dozen host<->device copies and synchronisations per frame, and the join and the map
// merge run on the host.  Here only the new frame's measurements go up
// (44 B per measurement) and the pose comes down (one synchronisation per frame); the match list,
// the join, the correspondence counts, the triangulated cloud and the map never leave the GPU:
//   nn_filter (nn.cu)  ->  assoc_join_kernel  ->  picp_resident_kernel (counts read on device, the
//   cloud transform folded into its gather)  ->  pose to the host  ->  triangulate_kernel (count
//   read on device)  ->  map_update_kernel (device hash index over the map's appearances).
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>

#include "common.cuh"

namespace vo {

constexpr int PIPE_THREADS = 1024;
constexpr int PIPE_JOIN_CLUSTER = 8;    // CTAs of the association + join kernel (one cluster)
constexpr int PIPE_MAX_POINTS = 32768;  // per frame: the join table lives in shared memory

// ---- block-wide exclusive rank of a flag (one call = one chunk of PIPE_THREADS items) ----------
// returns the rank of this thread's item among the flagged ones of the chunk; *total = their count.
// Contains two __syncthreads().
__device__ __forceinline__ int block_rank(bool flag, int* s_warp /*[32]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(bal & ((1u << lane) - 1u));
  __syncthreads();  // previous readers of s_warp are done
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < PIPE_THREADS / 32; ++w) {
    const int x = s_warp[w];
    off += (w < warp) ? x : 0;
    tot += x;
  }
  *total = tot;
  return off + in_warp;
}

// ---- association + join -----------------------------------------------------------------------------
// nn_idx[q]: best map row of query q or -1.  `first_is_map`: the map side is the reference frame
// (vo_complete.cpp:20-33 builds the tree over the larger set).  Emits, in query order,
//   corr_imgs  = (ref_idx, cur_idx)                          vo_complete.cpp:39-43
//   pairs_picp = (cur_idx, world_idx) for every corr_imgs entry whose ref_idx appears in
//                corr_world (first entry wins)               vo_complete.cpp:52-66
// counts[0] = |corr_imgs|, counts[1] = |pairs_picp| (int64 slots).
__global__ void __launch_bounds__(PIPE_THREADS)
assoc_join_kernel(const int32_t* __restrict__ nn_idx, int n_q, int first_is_map, int n_ref,
                  const int2* __restrict__ corr_world, const int* __restrict__ n_corr_world_dev,
                  int2* __restrict__ corr_imgs, int2* __restrict__ pairs_picp,
                  long long* __restrict__ counts) {
  extern __shared__ int s_first[];  // [n_ref]: first position in corr_world of each ref index
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  const int n_cw = n_corr_world_dev ? *n_corr_world_dev : 0;
  for (int r = tid; r < n_ref; r += PIPE_THREADS) s_first[r] = INT_MAX;
  __syncthreads();
  for (int j = tid; j < n_cw; j += PIPE_THREADS) {
    const int r = corr_world[j].x;
    if (r >= 0 && r < n_ref) atomicMin(&s_first[r], j);
  }
  __syncthreads();
  int n_ci = 0, n_pp = 0;
  for (int base = 0; base < n_q; base += PIPE_THREADS) {
    const int q = base + tid;
    const int m = q < n_q ? nn_idx[q] : -1;
    const bool hit = m >= 0;
    const int ref = first_is_map ? m : q, cur = first_is_map ? q : m;
    int j = INT_MAX;
    if (hit && ref < n_ref) j = s_first[ref];
    const bool joined = hit && j != INT_MAX;
    int t1, t2;
    const int r1 = block_rank(hit, s_warp, &t1);
    const int r2 = block_rank(joined, s_warp, &t2);
    if (hit) corr_imgs[n_ci + r1] = make_int2(ref, cur);
    if (joined) pairs_picp[n_pp + r2] = make_int2(cur, corr_world[j].y);
    n_ci += t1;
    n_pp += t2;
  }
  if (tid == 0) {
    counts[0] = n_ci;
    counts[1] = n_pp;
  }
}

// The same on a thread-block cluster: CTA r takes the r-th contiguous slice of the queries, counts its
// hits and joins, hands the two counts to every peer through distributed shared memory (one cluster
// barrier), and writes its slice at the offsets of the slices before it — the output order is the
// single-CTA kernel's.  Every CTA builds the whole join table itself (|corr_world| is a few thousand).
// 29 us -> measured in profiles/r02_notes.md for a 1e4-measurement frame.
__global__ void __cluster_dims__(PIPE_JOIN_CLUSTER, 1, 1) __launch_bounds__(PIPE_THREADS)
assoc_join_cluster_kernel(const int32_t* __restrict__ nn_idx, int n_q, int first_is_map, int n_ref,
                          const int2* __restrict__ corr_world, const int* __restrict__ n_corr_world_dev,
                          int2* __restrict__ corr_imgs, int2* __restrict__ pairs_picp,
                          long long* __restrict__ counts) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  extern __shared__ int s_first[];  // [n_ref]: first position in corr_world of each ref index
  __shared__ int s_warp[32];
  __shared__ int s_cnt[PIPE_JOIN_CLUSTER][2];  // [source CTA][hits, joins], written by the peers
  const int tid = threadIdx.x;
  const int n_cw = n_corr_world_dev ? *n_corr_world_dev : 0;
  for (int r = tid; r < n_ref; r += PIPE_THREADS) s_first[r] = INT_MAX;
  __syncthreads();
  for (int j = tid; j < n_cw; j += PIPE_THREADS) {
    const int r = corr_world[j].x;
    if (r >= 0 && r < n_ref) atomicMin(&s_first[r], j);
  }
  __syncthreads();
  const int chunk = (n_q + PIPE_JOIN_CLUSTER - 1) / PIPE_JOIN_CLUSTER;
  const int q0 = min(n_q, rank * chunk), q1 = min(n_q, q0 + chunk);
  // pass 1: this slice's counts
  int c_hit = 0, c_join = 0;
  for (int q = q0 + tid; q < q1; q += PIPE_THREADS) {
    const int m = nn_idx[q];
    const bool hit = m >= 0;
    const int ref = first_is_map ? m : q;
    c_hit += hit ? 1 : 0;
    c_join += (hit && ref < n_ref && s_first[ref] != INT_MAX) ? 1 : 0;
  }
  int packed = (c_join << 16) | c_hit;  // a thread sees at most 32 queries of <= 32768
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int other = __shfl_xor_sync(0xffffffffu, packed, o);
    packed = (((packed >> 16) + (other >> 16)) << 16) | ((packed & 0xFFFF) + (other & 0xFFFF));
  }
  if ((tid & 31) == 0) s_warp[tid >> 5] = packed;
  __syncthreads();
  if (tid == 0) {
    int hits = 0, joins = 0;
    for (int w = 0; w < PIPE_THREADS / 32; ++w) {
      hits += s_warp[w] & 0xFFFF;
      joins += s_warp[w] >> 16;
    }
    for (int r = 0; r < PIPE_JOIN_CLUSTER; ++r) {
      int* dst = cluster.map_shared_rank(&s_cnt[rank][0], r);
      dst[0] = hits;
      dst[1] = joins;
    }
  }
  cluster.sync();  // all counts are visible everywhere; also orders the reuse of s_warp below
  int n_ci = 0, n_pp = 0, all_ci = 0, all_pp = 0;
  for (int r = 0; r < PIPE_JOIN_CLUSTER; ++r) {
    n_ci += r < rank ? s_cnt[r][0] : 0;
    n_pp += r < rank ? s_cnt[r][1] : 0;
    all_ci += s_cnt[r][0];
    all_pp += s_cnt[r][1];
  }
  // pass 2: the ordered writes of this slice
  for (int base = q0; base < q1; base += PIPE_THREADS) {
    const int q = base + tid;
    const int m = q < q1 ? nn_idx[q] : -1;
    const bool hit = m >= 0;
    const int ref = first_is_map ? m : q, cur = first_is_map ? q : m;
    int j = INT_MAX;
    if (hit && ref < n_ref) j = s_first[ref];
    const bool joined = hit && j != INT_MAX;
    int t1, t2;
    const int r1 = block_rank(hit, s_warp, &t1);
    const int r2 = block_rank(joined, s_warp, &t2);
    if (hit) corr_imgs[n_ci + r1] = make_int2(ref, cur);
    if (joined) pairs_picp[n_pp + r2] = make_int2(cur, corr_world[j].y);
    n_ci += t1;
    n_pp += t2;
  }
  if (rank == 0 && tid == 0) {
    counts[0] = all_ci;
    counts[1] = all_pp;
  }
}

// ---- map merge (PointCloudVector::update, PointCloud.h:52-66) on a device hash index ---------------
// Sequential semantics to reproduce: for every new point, in order — if some stored appearance is
// equal (float ==, first hit) its position is overwritten, else the point is appended; appended
// points take part in the matching of later ones.  Equivalent order-free statement used here:
//   * a key that already exists keeps its slot; a new key is appended at the rank of its FIRST
//     occurrence among the new keys; either way the stored position is the LAST occurrence's.
//   * an appearance holding a NaN equals nothing: always appended, never indexed.
// slots[]: 0xFFFFFFFF empty | map index (< 2^31) | 0x80000000+i = claimed this call by new point i.
constexpr unsigned int MAP_EMPTY = 0xFFFFFFFFu;
constexpr unsigned int MAP_CLAIM = 0x80000000u;
// a claim dropped because the map was full: probing walks over it (an EMPTY here would cut every key
// that was inserted past this slot off its probe sequence) and never claims it
constexpr unsigned int MAP_TOMB = 0x7FFFFFFFu;

__device__ __forceinline__ bool app_hash(const float* a, unsigned long long* h) {
  unsigned long long x = 0x9E3779B97F4A7C15ull;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    float v = a[i];
    ok = ok && (v == v);
    v = (v == 0.f) ? 0.f : v;  // -0.0 == +0.0
    x = (x ^ (unsigned long long)__float_as_uint(v)) * 0x100000001B3ull;
    x ^= x >> 29;
  }
  *h = x;
  return ok;
}
__device__ __forceinline__ bool app_equal(const float* a, const float* b) {
  bool eq = true;
#pragma unroll
  for (int i = 0; i < 10; ++i) eq = eq && (a[i] == b[i]);
  return eq;
}

struct MapParams {
  const float* new_pts;   // 3 floats / point, in the frame of the triangulation
  const float* new_app;   // 10 floats / point
  const long long* n_new_dev;
  float X[12];            // `history`: 3x3 linear (column-major) then translation
  float* map_pts;
  float* map_app;
  unsigned int* slots;
  int* last;              // per slot: last new index that hit it in this call (-1 between calls)
  int* key_slot;          // scratch [PIPE_MAX_POINTS]: slot of each new point, -1 = not comparable
  unsigned int cap_mask;
  long long max_map;
  long long* n_map_dev;
  long long* overflow_dev;
};

__global__ void __launch_bounds__(PIPE_THREADS) map_update_kernel(const MapParams p) {
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  const int n_new = (int)min(*p.n_new_dev, (long long)PIPE_MAX_POINTS);
  const long long n_map0 = *p.n_map_dev;
  // pass 1: find or claim the slot of every new appearance
  for (int i = tid; i < n_new; i += PIPE_THREADS) {
    float a[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) a[k] = p.new_app[10 * (long long)i + k];
    unsigned long long h;
    int found = -1;
    if (app_hash(a, &h)) {
      unsigned int s = (unsigned int)h & p.cap_mask;
      while (true) {
        unsigned int v = *reinterpret_cast<volatile unsigned int*>(p.slots + s);
        if (v == MAP_EMPTY) {
          const unsigned int old = atomicCAS(p.slots + s, MAP_EMPTY, MAP_CLAIM | (unsigned int)i);
          if (old == MAP_EMPTY) {
            found = (int)s;
            break;
          }
          v = old;
        }
        if (v == MAP_TOMB) {
          s = (s + 1) & p.cap_mask;
          continue;
        }
        const float* other = (v & MAP_CLAIM) ? p.new_app + 10 * (long long)(v & ~MAP_CLAIM)
                                             : p.map_app + 10 * (long long)v;
        float b[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) b[k] = other[k];
        if (app_equal(a, b)) {
          // same key claimed by another new point: the lowest new index is the first occurrence
          if (v & MAP_CLAIM) atomicMin(p.slots + s, MAP_CLAIM | (unsigned int)i);
          found = (int)s;
          break;
        }
        s = (s + 1) & p.cap_mask;
      }
      atomicMax(p.last + found, i);
    }
    p.key_slot[i] = found;
  }
  __syncthreads();
  // pass 2: ordered append of the first occurrence of every new key (and of every NaN point)
  int appended = 0;
  for (int base = 0; base < n_new; base += PIPE_THREADS) {
    const int i = base + tid;
    bool app = false;
    int ks = -1;
    if (i < n_new) {
      ks = p.key_slot[i];
      app = ks < 0 || p.slots[ks] == (MAP_CLAIM | (unsigned int)i);
    }
    int tot;
    const int r = block_rank(app, s_warp, &tot);
    if (app) {
      const long long pos = n_map0 + appended + r;
      if (pos < p.max_map) {
#pragma unroll
        for (int k = 0; k < 10; ++k) p.map_app[10 * pos + k] = p.new_app[10 * (long long)i + k];
        if (ks >= 0) {
          p.slots[ks] = (unsigned int)pos;
        } else {  // not comparable: nobody else writes this point
          const float x = p.new_pts[3 * (long long)i], y = p.new_pts[3 * (long long)i + 1],
                      z = p.new_pts[3 * (long long)i + 2];
          p.map_pts[3 * pos + 0] = p.X[0] * x + (p.X[3] * y + p.X[6] * z) + p.X[9];
          p.map_pts[3 * pos + 1] = p.X[1] * x + (p.X[4] * y + p.X[7] * z) + p.X[10];
          p.map_pts[3 * pos + 2] = p.X[2] * x + (p.X[5] * y + p.X[8] * z) + p.X[11];
        }
      } else if (ks >= 0) {
        p.slots[ks] = MAP_TOMB;  // map full: drop the claim (flagged below), keep the probe chain
      }
    }
    appended += tot;
  }
  __syncthreads();
  // pass 3: every key takes the position of its LAST occurrence
  for (int i = tid; i < n_new; i += PIPE_THREADS) {
    const int ks = p.key_slot[i];
    if (ks < 0 || p.last[ks] != i) continue;
    const unsigned int j = p.slots[ks];
    if (j != MAP_EMPTY && j != MAP_TOMB && !(j & MAP_CLAIM)) {
      const float x = p.new_pts[3 * (long long)i], y = p.new_pts[3 * (long long)i + 1],
                  z = p.new_pts[3 * (long long)i + 2];
      p.map_pts[3 * (long long)j + 0] = p.X[0] * x + (p.X[3] * y + p.X[6] * z) + p.X[9];
      p.map_pts[3 * (long long)j + 1] = p.X[1] * x + (p.X[4] * y + p.X[7] * z) + p.X[10];
      p.map_pts[3 * (long long)j + 2] = p.X[2] * x + (p.X[5] * y + p.X[8] * z) + p.X[11];
    }
  }
  __syncthreads();
  for (int i = tid; i < n_new; i += PIPE_THREADS) {
    const int ks = p.key_slot[i];
    if (ks >= 0) p.last[ks] = -1;
  }
  if (tid == 0) {
    long long n = n_map0 + appended;
    if (n > p.max_map) {
      *p.overflow_dev = 1;
      n = p.max_map;
    }
    *p.n_map_dev = n;
  }
}

// ---- tiny host-side 4x4 helpers (column-major isometries) -----------------------------------------
static void iso_identity(float* X) {
  memset(X, 0, 16 * sizeof(float));
  X[0] = X[5] = X[10] = X[15] = 1.f;
}
static void iso_mul(const float* A, const float* B, float* C) {  // C = A*B
  float R[16];
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc += A[k * 4 + i] * B[j * 4 + k];
      R[j * 4 + i] = acc;
    }
  memcpy(C, R, sizeof(R));
}
static void iso_inverse(const float* X, float* Y) {  // (R, t)^-1 = (R^T, -R^T t)
  float R[16];
  iso_identity(R);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[j * 4 + i] = X[i * 4 + j];
  for (int i = 0; i < 3; ++i)
    R[12 + i] = -(R[i] * X[12] + R[4 + i] * X[13] + R[8 + i] * X[14]);
  memcpy(Y, R, sizeof(R));
}

}  // namespace vo

using namespace vo;

struct vo_pipe_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  // the map merge of frame k only feeds the map: it runs on a side stream, concurrently with the
  // association and the PICP rounds of frame k+1
  cudaStream_t map_stream = nullptr;
  cudaEvent_t tri_done = nullptr, map_done = nullptr;
  vo_camera cam{};
  int64_t max_pts = 0, max_map = 0;
  vo_nn_t nn = nullptr;
  vo_picp_t picp = nullptr;
  DevBuf pts2d[2], app[2];
  int64_t n_frame[2] = {0, 0};
  int ref = 0;            // slot of the reference frame; the current frame goes to 1 - ref
  bool have_ref = false, have_cur = false, bootstrapped = false;
  DevBuf nn_idx, corr_imgs, pairs_picp, corr_world[2], tri_pts[2], tri_app[2], counts, tri_ws;
  int tri_slot = 0;       // which corr_world / tri_* holds the latest triangulation
  DevBuf map_pts, map_app, map_slots, map_last, key_slot;
  unsigned int cap = 0;
  float X_curr[16], history[16];
  unsigned char* pinned = nullptr;  // host staging for the per-frame read-back
  int64_t n_q_last = 0;   // upper bound of |corr_imgs| of the last association
};

// counts (int64 each): 0 |corr_imgs|, 1 |pairs_picp|, 2/3 |triangulated| per slot, 4 |map|, 5 overflow
enum { C_CI = 0, C_PP = 1, C_TRI0 = 2, C_MAP = 4, C_OVF = 5, C_N = 8 };

static int pipe_upload_frame(vo_pipe_s* h, int slot, const float* pts_host, const float* app_host,
                             int64_t n) {
  VO_REQUIRE(n >= 0 && n <= h->max_pts, VO_ERR_ARG, "frame larger than max_points_per_frame");
  VO_REQUIRE((pts_host && app_host) || n == 0, VO_ERR_ARG, "null frame");
  if (n) {
    VO_CUDA(cudaMemcpyAsync(h->pts2d[slot].p, pts_host, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    VO_CUDA(cudaMemcpyAsync(h->app[slot].p, app_host, (size_t)n * 40, cudaMemcpyHostToDevice, h->stream));
    if (host_source_still_in_use(pts_host) || host_source_still_in_use(app_host))
      VO_CUDA(cudaStreamSynchronize(h->stream));
  }
  h->n_frame[slot] = n;
  return VO_OK;
}

// association of the reference and the current frame + join with the latest triangulation
static int pipe_associate(vo_pipe_s* h, bool with_join) {
  const int ref = h->ref, cur = 1 - h->ref;
  const int64_t n_ref = h->n_frame[ref], n_cur = h->n_frame[cur];
  const bool first_is_map = n_ref >= n_cur;  // vo_complete.cpp:15,20
  const int map_slot = first_is_map ? ref : cur, q_slot = first_is_map ? cur : ref;
  const int64_t n_q = h->n_frame[q_slot];
  h->n_q_last = n_q;
  int rc = vo_nn_set_map_device(h->nn, h->app[map_slot].as<float>(), h->n_frame[map_slot], 10, 0);
  if (rc) return rc;
  rc = vo_nn_best_match_device(h->nn, h->app[q_slot].as<float>(), n_q, 10, 0.1f,
                               h->nn_idx.as<int32_t>(), nullptr);
  if (rc) return rc;
  const size_t smem = (size_t)(n_ref > 0 ? n_ref : 1) * sizeof(int);
  const int* n_cw_dev =
      with_join ? reinterpret_cast<const int*>(h->counts.as<long long>() + C_TRI0 + h->tri_slot) : nullptr;
  static const bool single = getenv("VO_PIPE_JOIN_SINGLE") != nullptr;  // tests: the one-CTA kernel
  if (n_q >= 2 * PIPE_THREADS && !single)  // a slice per CTA of a cluster; small frames stay on one CTA
    assoc_join_cluster_kernel<<<PIPE_JOIN_CLUSTER, PIPE_THREADS, smem, h->stream>>>(
        h->nn_idx.as<int32_t>(), (int)n_q, first_is_map ? 1 : 0, (int)n_ref,
        h->corr_world[h->tri_slot].as<int2>(), n_cw_dev, h->corr_imgs.as<int2>(), h->pairs_picp.as<int2>(),
        h->counts.as<long long>());
  else
    assoc_join_kernel<<<1, PIPE_THREADS, smem, h->stream>>>(
        h->nn_idx.as<int32_t>(), (int)n_q, first_is_map ? 1 : 0, (int)n_ref,
        h->corr_world[h->tri_slot].as<int2>(), n_cw_dev, h->corr_imgs.as<int2>(), h->pairs_picp.as<int2>(),
        h->counts.as<long long>());
  VO_LAUNCH_CHECK();
  return VO_OK;
}

static int pipe_merge(vo_pipe_s* h, cudaStream_t stream, const float* pts_dev, const float* app_dev,
                      const long long* n_new_dev, const float* X);

// triangulate (ref, cur) with pose X into the other tri slot and merge into the map with `history`
static int pipe_triangulate_and_merge(vo_pipe_s* h, const float* X) {
  const int ref = h->ref, cur = 1 - h->ref, nt = 1 - h->tri_slot;
  long long* counts = h->counts.as<long long>();
  // the slot about to be overwritten was read by the merge of two frames ago; the merges are
  // ordered on their stream, so waiting for the latest one is enough
  VO_CUDA(cudaStreamWaitEvent(h->stream, h->map_done, 0));
  int rc = vo_triangulate_device_ex(
      h->stream, h->cam.K, X, h->corr_imgs.as<int32_t>(), h->n_q_last,
      reinterpret_cast<const int32_t*>(counts + C_CI), h->pts2d[ref].as<float>(),
      h->pts2d[cur].as<float>(), h->app[cur].as<float>(), h->tri_pts[nt].as<float>(),
      h->corr_world[nt].as<int32_t>(), h->tri_app[nt].as<float>(),
      reinterpret_cast<int64_t*>(counts + C_TRI0 + nt), h->tri_ws.p);
  if (rc) return rc;
  VO_CUDA(cudaEventRecord(h->tri_done, h->stream));
  VO_CUDA(cudaStreamWaitEvent(h->map_stream, h->tri_done, 0));
  rc = pipe_merge(h, h->map_stream, h->tri_pts[nt].as<float>(), h->tri_app[nt].as<float>(),
                  counts + C_TRI0 + nt, h->history);
  if (rc) return rc;
  VO_CUDA(cudaEventRecord(h->map_done, h->map_stream));
  h->tri_slot = nt;
  return VO_OK;
}

static int pipe_merge(vo_pipe_s* h, cudaStream_t stream, const float* pts_dev, const float* app_dev,
                      const long long* n_new_dev, const float* X) {
  long long* counts = h->counts.as<long long>();
  MapParams p;
  p.new_pts = pts_dev;
  p.new_app = app_dev;
  p.n_new_dev = n_new_dev;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 3; ++i) p.X[j * 3 + i] = X[j * 4 + i];
  p.map_pts = h->map_pts.as<float>();
  p.map_app = h->map_app.as<float>();
  p.slots = h->map_slots.as<unsigned int>();
  p.last = h->map_last.as<int>();
  p.key_slot = h->key_slot.as<int>();
  p.cap_mask = h->cap - 1;
  p.max_map = h->max_map;
  p.n_map_dev = counts + C_MAP;
  p.overflow_dev = counts + C_OVF;
  map_update_kernel<<<1, PIPE_THREADS, 0, stream>>>(p);
  VO_LAUNCH_CHECK();
  return VO_OK;
}

extern "C" {

int vo_pipe_create(vo_pipe_t* out, int device, const vo_camera* cam, int64_t max_points_per_frame,
                   int64_t max_map_points) {
  VO_REQUIRE(out && cam, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(max_points_per_frame > 0 && max_points_per_frame <= PIPE_MAX_POINTS, VO_ERR_ARG,
             "max_points_per_frame must be in 1..32768");
  VO_REQUIRE(max_map_points > 0 && max_map_points < (1LL << 30), VO_ERR_ARG, "bad max_map_points");
  int n = 0;
  VO_CUDA(cudaGetDeviceCount(&n));
  VO_REQUIRE(device >= 0 && device < n, VO_ERR_ARG, "bad device ordinal");
  DeviceGuard g(device);
  vo_pipe_s* h = new vo_pipe_s();
  // a CUDA failure from here on releases everything acquired so far
#define PIPE_TRY(expr)                                                  \
  do {                                                                  \
    const cudaError_t e_ = (expr);                                      \
    if (e_ != cudaSuccess) {                                            \
      set_error("vo_pipe_create: %s -> %s", #expr, cudaGetErrorString(e_)); \
      vo_pipe_destroy(h);                                               \
      return VO_ERR_CUDA;                                               \
    }                                                                   \
  } while (0)
  h->device = device;
  h->cam = *cam;
  h->max_pts = max_points_per_frame;
  h->max_map = max_map_points;
  auto fail = [&](int rc) {
    vo_pipe_destroy(h);
    return rc;
  };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    delete h;
    return VO_ERR_CUDA;
  }
  if (cudaStreamCreateWithFlags(&h->map_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->tri_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->map_done, cudaEventDisableTiming) != cudaSuccess) {
    set_error("stream/event creation failed");
    vo_pipe_destroy(h);
    return VO_ERR_CUDA;
  }
  int rc;
  if ((rc = vo_nn_create(&h->nn, device)) || (rc = vo_nn_set_stream(h->nn, h->stream))) return fail(rc);
  if ((rc = vo_picp_create(&h->picp, device)) || (rc = vo_picp_set_stream(h->picp, h->stream)))
    return fail(rc);
  const size_t P = (size_t)h->max_pts;
  for (int s = 0; s < 2; ++s) {
    if ((rc = h->pts2d[s].reserve(P * 8 + 16)) || (rc = h->app[s].reserve(P * 40 + 16)) ||
        (rc = h->corr_world[s].reserve(P * 8 + 16)) || (rc = h->tri_pts[s].reserve(P * 12 + 16)) ||
        (rc = h->tri_app[s].reserve(P * 40 + 16)))
      return fail(rc);
  }
  if ((rc = h->nn_idx.reserve(P * 4 + 16)) || (rc = h->corr_imgs.reserve(P * 8 + 16)) ||
      (rc = h->pairs_picp.reserve(P * 8 + 16)) || (rc = h->counts.reserve(C_N * 8)) ||
      (rc = h->key_slot.reserve(P * 4 + 16)) ||
      (rc = h->tri_ws.reserve((size_t)vo_triangulate_workspace_bytes((int64_t)P))))
    return fail(rc);
  unsigned int cap = 1024;
  while ((long long)cap < 2 * h->max_map + 2 * h->max_pts) cap *= 2;
  h->cap = cap;
  if ((rc = h->map_pts.reserve((size_t)h->max_map * 12 + 16)) ||
      (rc = h->map_app.reserve((size_t)h->max_map * 40 + 16)) ||
      (rc = h->map_slots.reserve((size_t)cap * 4)) || (rc = h->map_last.reserve((size_t)cap * 4)))
    return fail(rc);
  PIPE_TRY(cudaMemsetAsync(h->counts.p, 0, C_N * 8, h->stream));
  PIPE_TRY(cudaMemsetAsync(h->map_slots.p, 0xFF, (size_t)cap * 4, h->stream));
  PIPE_TRY(cudaMemsetAsync(h->map_last.p, 0xFF, (size_t)cap * 4, h->stream));
  PIPE_TRY(cudaFuncSetAttribute(assoc_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               PIPE_MAX_POINTS * (int)sizeof(int)));
  PIPE_TRY(cudaFuncSetAttribute(assoc_join_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               PIPE_MAX_POINTS * (int)sizeof(int)));
  iso_identity(h->X_curr);
  iso_identity(h->history);
  PIPE_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h->pinned), C_N * 8 + sizeof(vo_picp_state) + 64,
                        cudaHostAllocDefault));
  PIPE_TRY(cudaStreamSynchronize(h->stream));
  PIPE_TRY(cudaEventRecord(h->map_done, h->map_stream));  // "no merge pending"
#undef PIPE_TRY
  *out = h;
  return VO_OK;
}

int vo_pipe_destroy(vo_pipe_t h) {
  if (!h) return VO_OK;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->map_stream) cudaStreamSynchronize(h->map_stream);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->tri_done) cudaEventDestroy(h->tri_done);
  if (h->map_done) cudaEventDestroy(h->map_done);
  if (h->map_stream) cudaStreamDestroy(h->map_stream);
  if (h->nn) vo_nn_destroy(h->nn);
  if (h->picp) vo_picp_destroy(h->picp);
  for (int s = 0; s < 2; ++s) {
    h->pts2d[s].release();
    h->app[s].release();
    h->corr_world[s].release();
    h->tri_pts[s].release();
    h->tri_app[s].release();
  }
  h->nn_idx.release();
  h->corr_imgs.release();
  h->pairs_picp.release();
  h->counts.release();
  h->tri_ws.release();
  h->map_pts.release();
  h->map_app.release();
  h->map_slots.release();
  h->map_last.release();
  h->key_slot.release();
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return VO_OK;
}

int vo_pipe_first_frame(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n) {
  VO_REQUIRE(h != nullptr, VO_ERR_ARG, "null handle");
  DeviceGuard g(h->device);
  // a new sequence starts from an empty map: counters, overflow flag, hash index and history are
  // reset (a second sequence used to merge into the first one's map)
  cudaStreamSynchronize(h->map_stream);
  VO_CUDA(cudaMemsetAsync(h->counts.p, 0, C_N * 8, h->stream));
  VO_CUDA(cudaMemsetAsync(h->map_slots.p, 0xFF, (size_t)h->cap * 4, h->stream));
  VO_CUDA(cudaMemsetAsync(h->map_last.p, 0xFF, (size_t)h->cap * 4, h->stream));
  iso_identity(h->X_curr);
  iso_identity(h->history);
  h->tri_slot = 0;
  h->n_q_last = 0;
  h->ref = 0;
  int rc = pipe_upload_frame(h, 0, points_host, app_host, n);
  if (rc) return rc;
  h->have_ref = true;
  h->have_cur = h->bootstrapped = false;
  return VO_OK;
}

int vo_pipe_second_frame(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n,
                         int32_t* corr_host, int64_t corr_capacity, int64_t* n_corr) {
  VO_REQUIRE(h != nullptr && n_corr != nullptr, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(corr_capacity >= 0 && (corr_host != nullptr || corr_capacity == 0), VO_ERR_ARG,
             "bad match buffer");
  VO_REQUIRE(h->have_ref && !h->bootstrapped, VO_ERR_STATE, "call vo_pipe_first_frame first");
  DeviceGuard g(h->device);
  int rc = pipe_upload_frame(h, 1 - h->ref, points_host, app_host, n);
  if (rc) return rc;
  rc = pipe_associate(h, /*with_join=*/false);
  if (rc) return rc;
  long long cnt = 0;
  VO_CUDA(cudaMemcpyAsync(&cnt, h->counts.as<long long>() + C_CI, 8, cudaMemcpyDeviceToHost, h->stream));
  VO_CUDA(cudaStreamSynchronize(h->stream));
  const long long take = cnt < corr_capacity ? cnt : corr_capacity;
  if (take > 0) {
    VO_CUDA(cudaMemcpyAsync(corr_host, h->corr_imgs.p, (size_t)take * 8, cudaMemcpyDeviceToHost, h->stream));
    VO_CUDA(cudaStreamSynchronize(h->stream));
  }
  *n_corr = cnt;
  h->have_cur = true;
  return VO_OK;
}

int vo_pipe_bootstrap(vo_pipe_t h, const float X[16]) {
  VO_REQUIRE(h != nullptr && X != nullptr, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(h->have_cur && !h->bootstrapped, VO_ERR_STATE, "call vo_pipe_second_frame first");
  DeviceGuard g(h->device);
  iso_identity(h->history);  // map.update(triangulated_pc)            vo_complete.cpp:146
  int rc = pipe_triangulate_and_merge(h, X);
  if (rc) return rc;
  iso_inverse(X, h->history);  // history = X.inverse()                vo_complete.cpp:147
  memcpy(h->X_curr, X, 16 * sizeof(float));
  h->ref = 1 - h->ref;  // reference_pc = current_pc                   vo_complete.cpp:140
  h->have_cur = false;
  h->bootstrapped = true;
  return VO_OK;
}

int vo_pipe_step(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n, int rounds,
                 float kernel_threshold, vo_pipe_result* out) {
  VO_REQUIRE(h != nullptr && out != nullptr, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(h->bootstrapped, VO_ERR_STATE, "call vo_pipe_bootstrap first");
  VO_REQUIRE(rounds >= 0, VO_ERR_ARG, "negative rounds");
  DeviceGuard g(h->device);
  const int cur = 1 - h->ref;
  int rc = pipe_upload_frame(h, cur, points_host, app_host, n);            // vo_complete.cpp:150
  if (rc) return rc;
  if ((rc = pipe_associate(h, /*with_join=*/true))) return rc;             // :151-152
  // :154-160 — world points = X_curr * triangulated (folded into the solver's gather), camera
  // reset to the identity, `rounds` Gauss-Newton iterations
  vo_camera cam = h->cam;
  iso_identity(cam.T);
  if ((rc = vo_picp_set_params(h->picp, kernel_threshold, 1.f, 0))) return rc;
  if ((rc = vo_picp_init_device(h->picp, &cam, h->tri_pts[h->tri_slot].as<float>(), h->max_pts,
                                h->pts2d[cur].as<float>(), n)))
    return rc;
  if ((rc = vo_picp_set_correspondences_device(h->picp, h->pairs_picp.as<int32_t>(), h->n_q_last)))
    return rc;
  if ((rc = vo_picp_compute_ex(h->picp, 0, rounds,
                               reinterpret_cast<const int32_t*>(h->counts.as<long long>() + C_PP),
                               h->X_curr)))
    return rc;
  // the frame's one synchronisation: counts and solver state land in pinned memory together
  long long* counts = reinterpret_cast<long long*>(h->pinned);
  vo_picp_state& st = *reinterpret_cast<vo_picp_state*>(h->pinned + C_N * 8);
  VO_CUDA(cudaMemcpyAsync(counts, h->counts.p, C_N * 8, cudaMemcpyDeviceToHost, h->stream));
  VO_CUDA(cudaMemcpyAsync(&st, picp_state_device_ptr(h->picp), sizeof(vo_picp_state),
                          cudaMemcpyDeviceToHost, h->stream));
  VO_CUDA(cudaStreamSynchronize(h->stream));
  memcpy(h->X_curr, st.T, sizeof(st.T));                                   // :161-163
  memcpy(out->T, st.T, sizeof(st.T));
  out->n_measurements = n;
  out->n_matches = counts[C_CI];
  out->n_correspondences = counts[C_PP];
  out->n_inliers = st.num_inliers;
  out->chi_inliers = st.chi_inliers;
  out->map_points = counts[C_MAP];  // as of the previous frame's merge
  out->map_overflow = counts[C_OVF] != 0;
  // :168-176 — re-triangulate with the new pose and merge (asynchronous: overlaps the caller)
  if ((rc = pipe_triangulate_and_merge(h, h->X_curr))) return rc;
  float inv[16];
  iso_inverse(h->X_curr, inv);
  iso_mul(h->history, inv, h->history);  // history = history * X_curr.inverse()
  h->ref = cur;                          // reference_pc = current_pc
  return VO_OK;
}

int vo_pipe_merge_cloud(vo_pipe_t h, const float* points_host, const float* app_host, int64_t n,
                        const float X[16]) {
  VO_REQUIRE(h != nullptr && X != nullptr, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(n >= 0 && n <= h->max_pts, VO_ERR_ARG, "cloud larger than max_points_per_frame");
  VO_REQUIRE((points_host && app_host) || n == 0, VO_ERR_ARG, "null cloud");
  if (n == 0) return VO_OK;
  DeviceGuard g(h->device);
  VO_CUDA(cudaStreamSynchronize(h->map_stream));
  // staged through the buffers of the NEXT triangulation (free between frames)
  const int nt = 1 - h->tri_slot;
  long long* counts = h->counts.as<long long>();
  const long long cnt = n;
  VO_CUDA(cudaMemcpyAsync(h->tri_pts[nt].p, points_host, (size_t)n * 12, cudaMemcpyHostToDevice, h->stream));
  VO_CUDA(cudaMemcpyAsync(h->tri_app[nt].p, app_host, (size_t)n * 40, cudaMemcpyHostToDevice, h->stream));
  VO_CUDA(cudaMemcpyAsync(counts + C_TRI0 + nt, &cnt, 8, cudaMemcpyHostToDevice, h->stream));
  int rc = pipe_merge(h, h->stream, h->tri_pts[nt].as<float>(), h->tri_app[nt].as<float>(),
                      counts + C_TRI0 + nt, X);
  if (rc) return rc;
  VO_CUDA(cudaStreamSynchronize(h->stream));
  return VO_OK;
}

int vo_pipe_get_map(vo_pipe_t h, float* points_host, float* app_host, int64_t capacity, int64_t* n) {
  VO_REQUIRE(h != nullptr && n != nullptr, VO_ERR_ARG, "null pointer");
  DeviceGuard g(h->device);
  VO_CUDA(cudaStreamSynchronize(h->map_stream));  // the last frame's merge
  long long cnt = 0;
  VO_CUDA(cudaMemcpyAsync(&cnt, h->counts.as<long long>() + C_MAP, 8, cudaMemcpyDeviceToHost, h->stream));
  VO_CUDA(cudaStreamSynchronize(h->stream));
  *n = cnt;
  const int64_t m = cnt < capacity ? cnt : capacity;
  if (m > 0 && points_host)
    VO_CUDA(cudaMemcpyAsync(points_host, h->map_pts.p, (size_t)m * 12, cudaMemcpyDeviceToHost, h->stream));
  if (m > 0 && app_host)
    VO_CUDA(cudaMemcpyAsync(app_host, h->map_app.p, (size_t)m * 40, cudaMemcpyDeviceToHost, h->stream));
  VO_CUDA(cudaStreamSynchronize(h->stream));
  return VO_OK;
}

}  // extern "C"
