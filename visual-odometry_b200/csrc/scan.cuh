// scan.cuh — single-pass, order-preserving ("stable") stream compaction support.
//
// The reference compacts successes sequentially (n_success counter, src/utils.cpp:96-100 and
// src/camera.cpp:31-33).  On the device every tile of TILE items computes its items and flags,
// ranks them with warp ballots, and obtains the number of successes in all EARLIER tiles through
// a decoupled look-back over 64-bit status words  [2-bit flag | 62-bit count]  — one pass over
// the data, no second kernel.  Tile ids are handed out by an atomic ticket so a tile can only
// ever wait on tiles that have already been claimed by a running block (forward progress without
// co-residency assumptions).  Blocks are persistent and software-pipelined: a block claims and
// starts loading its NEXT tile before it looks back for the current one, so the memory system
// stays busy while the prefix chain resolves (a block holding tile t and t' > t only ever waits
// on tiles < t, hence no cycle).
#pragma once
#include "common.cuh"

namespace vo {

constexpr unsigned long long SCAN_AGG = 1ull << 62;     // tile total published
constexpr unsigned long long SCAN_PREFIX = 2ull << 62;  // inclusive prefix published
constexpr unsigned long long SCAN_MASK = (1ull << 62) - 1;

struct ScanWorkspace {
  unsigned long long* status;  // [num_tiles], zero-initialised before the launch
  unsigned int* ticket;        // zero-initialised before the launch
};

inline int64_t scan_workspace_bytes(int64_t num_tiles) { return (num_tiles + 1) * 8 + 64; }
inline ScanWorkspace scan_workspace_at(void* base, int64_t num_tiles) {
  ScanWorkspace w;
  w.status = reinterpret_cast<unsigned long long*>(base);
  w.ticket = reinterpret_cast<unsigned int*>(w.status + num_tiles);
  return w;
}

#ifdef __CUDACC__
// The status word carries flag and count in ONE 64-bit value and nothing else is published
// through it, so relaxed (L2-coherent, non-caching) accesses are sufficient — an acquire load
// would invalidate the SM's L1 on every poll.
__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void scan_st(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Step 1 of a tile: rank the tile's items with warp ballots and publish the tile total.
// Item j of lane l of warp w sits at tile position w*32*ITEMS + j*32 + l.  On return local[j] is
// the item's rank inside the tile (meaningful where flag[j] is set) and *tile_total the number of
// selected items of the tile.  Contains one __syncthreads().
template <int THREADS, int ITEMS>
__device__ __forceinline__ void scan_tile_post(const ScanWorkspace& ws, int tile,
                                               const bool (&flag)[ITEMS], int (&local)[ITEMS],
                                               int* tile_total) {
  constexpr int WARPS = THREADS / 32;
  __shared__ int s_warp_tot[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  int run = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const unsigned bal = __ballot_sync(0xffffffffu, flag[j]);
    local[j] = run + __popc(bal & lt);
    run += __popc(bal);
  }
  if (lane == 0) s_warp_tot[warp] = run;
  __syncthreads();
  int warp_off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    const int x = s_warp_tot[w];
    warp_off += (w < warp) ? x : 0;
    total += x;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) local[j] += warp_off;
  if (threadIdx.x == 0)
    scan_st(ws.status + tile, (tile == 0 ? SCAN_PREFIX : SCAN_AGG) | (unsigned long long)total);
  *tile_total = total;
}

// Step 2: number of selected items in all EARLIER tiles.  The look-back is done by the WHOLE
// block: thread i inspects predecessor tile-1-i, so one round covers THREADS predecessors.
// Contains __syncthreads(); every thread returns the same value.
template <int THREADS>
__device__ __forceinline__ long long scan_tile_lookback(const ScanWorkspace& ws, int tile,
                                                        int total) {
  constexpr int WARPS = THREADS / 32;
  __shared__ long long s_lb_sum[WARPS];
  __shared__ int s_lb_has[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (tile == 0) return 0;
  long long excl = 0;
  int idx = tile - 1;
  while (true) {
    const int j = idx - (int)threadIdx.x;
    unsigned long long v = SCAN_PREFIX;  // "tiles before tile 0": prefix 0
    if (j >= 0) {
      do {
        v = scan_ld(ws.status + j);
      } while ((v >> 62) == 0ull);
    }
    const unsigned pm = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
    const int first = pm ? (__ffs(pm) - 1) : 32;
    long long c = (lane <= first) ? (long long)(v & SCAN_MASK) : 0ll;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __syncthreads();  // readers of the previous round are done
    if (lane == 0) {
      s_lb_sum[warp] = c;
      s_lb_has[warp] = pm != 0u;
    }
    __syncthreads();
    bool done = false;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      if (!done) {
        excl += s_lb_sum[w];
        done = s_lb_has[w] != 0;
      }
    }
    if (done) break;
    idx -= THREADS;
  }
  if (threadIdx.x == 0)
    scan_st(ws.status + tile, SCAN_PREFIX | (unsigned long long)(excl + total));
  return excl;
}

// Dynamic tile id for a persistent block (one __syncthreads()).
__device__ __forceinline__ int scan_take_ticket(const ScanWorkspace& ws) {
  __shared__ int s_ticket;
  __syncthreads();  // previous readers of s_ticket are done
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ws.ticket, 1u);
  __syncthreads();
  return s_ticket;
}
#endif

}  // namespace vo
