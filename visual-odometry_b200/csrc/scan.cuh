// scan.cuh — single-pass, order-preserving ("stable") stream compaction support.
//
// The reference compacts successes sequentially (n_success counter, src/utils.cpp:96-100 and
// src/camera.cpp:31-33).  On the device every tile of TILE items computes its items and flags,
// ranks them with warp ballots, and obtains the number of successes in all EARLIER tiles through
// two levels of 64-bit status words (per tile, and per group of tiles) — one pass over the
// data, no second kernel.  Tile ids are handed out by an atomic ticket so a tile can only
// ever wait on tiles that have already been claimed by a running block (forward progress without
// co-residency assumptions).
#pragma once
#include "common.cuh"

namespace vo {

// status word of a tile: bit 63 = "posted", low bits = number of selected items in the tile.
// group word: high 24 bits = tiles of the group that have posted, low 40 bits = their sum.
constexpr unsigned long long SCAN_POSTED = 1ull << 63;
constexpr unsigned long long SCAN_VALUE_MASK = (1ull << 40) - 1;
constexpr int SCAN_GROUP_SHIFT = 40;

struct ScanWorkspace {
  unsigned long long* status;  // [num_tiles]   zero-initialised before the launch
  unsigned long long* groups;  // [num_groups]  zero-initialised before the launch
  unsigned int* ticket;        // zero-initialised before the launch
  int group_tiles;             // tiles per group
};

// tiles per group: at least 64, and few enough groups (<= 256) that one block-wide poll covers
// all of them
inline int scan_group_tiles(int64_t num_tiles) {
  int64_t g = (num_tiles + 255) / 256;
  return (int)(g < 64 ? 64 : g);
}
inline int64_t scan_workspace_bytes(int64_t num_tiles) { return (num_tiles + 256 + 2) * 8 + 64; }
inline ScanWorkspace scan_workspace_at(void* base, int64_t num_tiles) {
  ScanWorkspace w;
  w.group_tiles = scan_group_tiles(num_tiles);
  w.status = reinterpret_cast<unsigned long long*>(base);
  w.groups = w.status + num_tiles;
  w.ticket = reinterpret_cast<unsigned int*>(w.groups + 257);
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

// Step 1 of a tile: rank the tile's items with warp ballots and publish the tile total, both as
// the tile's own status word and as one atomic add into its group's (count, sum) word.
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
  if (threadIdx.x == 0) {
    scan_st(ws.status + tile, SCAN_POSTED | (unsigned long long)total);
    atomicAdd(ws.groups + tile / ws.group_tiles,
              (1ull << SCAN_GROUP_SHIFT) | (unsigned long long)total);
  }
  *tile_total = total;
}

// Step 2: number of selected items in all EARLIER tiles
//     = sum of the COMPLETE groups before this tile's group  (one word each)
//     + sum of the earlier tiles of its own group            (one word each).
// There is no prefix to propagate from tile to tile: every tile sums independently as soon as
// its predecessors have posted, with the whole block polling (thread i takes word i), so the
// only serial dependency left is the unavoidable one — an ordered output position needs the
// counts of everything before it.  Contains __syncthreads(); every thread returns the same value.
template <int THREADS>
__device__ __forceinline__ long long scan_tile_lookback(const ScanWorkspace& ws, int tile) {
  constexpr int WARPS = THREADS / 32;
  __shared__ long long s_lb_sum[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = ws.group_tiles;
  const int g = tile / G;
  long long c = 0;
  for (int idx = (int)threadIdx.x; idx < g; idx += THREADS) {
    unsigned long long v;
    do {
      v = scan_ld(ws.groups + idx);
    } while ((int)(v >> SCAN_GROUP_SHIFT) != G);
    c += (long long)(v & SCAN_VALUE_MASK);
  }
  for (int idx = g * G + (int)threadIdx.x; idx < tile; idx += THREADS) {
    unsigned long long v;
    do {
      v = scan_ld(ws.status + idx);
    } while ((v & SCAN_POSTED) == 0ull);
    c += (long long)(v & SCAN_VALUE_MASK);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __syncthreads();  // s_lb_sum may still be read by a previous call
  if (lane == 0) s_lb_sum[warp] = c;
  __syncthreads();
  long long excl = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) excl += s_lb_sum[w];
  return excl;
}

// Step 2, single-warp flavour: the same sum, gathered by WARP 0 ONLY (its lanes poll 32 words at a
// time, every word of a chunk is requested before the first one is examined) and handed to the rest
// of the block through shared memory.  The other warps sleep on the barrier instead of burning
// issue slots in a spin loop.  Contains one __syncthreads(); every thread returns the same value.
template <int THREADS>
__device__ __forceinline__ long long scan_tile_lookback_warp0(const ScanWorkspace& ws, int tile) {
  __shared__ long long s_excl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    const int G = ws.group_tiles;
    const int g = tile / G;
    long long c = 0;
    constexpr int B = 4;  // words in flight per lane
    for (int base = lane; base < g; base += 32 * B) {
      unsigned long long v[B];
#pragma unroll
      for (int k = 0; k < B; ++k) {
        const int idx = base + 32 * k;
        v[k] = idx < g ? scan_ld(ws.groups + idx) : ((unsigned long long)G << SCAN_GROUP_SHIFT);
      }
#pragma unroll
      for (int k = 0; k < B; ++k) {
        const int idx = base + 32 * k;
        while ((int)(v[k] >> SCAN_GROUP_SHIFT) != G) v[k] = scan_ld(ws.groups + idx);
        c += (long long)(v[k] & SCAN_VALUE_MASK);
      }
    }
    for (int base = g * G + lane; base < tile; base += 32 * B) {
      unsigned long long v[B];
#pragma unroll
      for (int k = 0; k < B; ++k) {
        const int idx = base + 32 * k;
        v[k] = idx < tile ? scan_ld(ws.status + idx) : SCAN_POSTED;
      }
#pragma unroll
      for (int k = 0; k < B; ++k) {
        const int idx = base + 32 * k;
        while ((v[k] & SCAN_POSTED) == 0ull) v[k] = scan_ld(ws.status + idx);
        c += (long long)(v[k] & SCAN_VALUE_MASK);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_excl = c;
  }
  __syncthreads();
  return s_excl;
}

// Dynamic tile id (two __syncthreads()).
__device__ __forceinline__ int scan_take_ticket(const ScanWorkspace& ws) {
  __shared__ int s_ticket;
  __syncthreads();  // previous readers of s_ticket are done
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ws.ticket, 1u);
  __syncthreads();
  return s_ticket;
}
#endif

}  // namespace vo
