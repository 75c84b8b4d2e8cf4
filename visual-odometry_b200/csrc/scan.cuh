// scan.cuh — single-pass, order-preserving ("stable") stream compaction support.
//
// The reference compacts successes sequentially (n_success counter, src/utils.cpp:96-100 and
// src/camera.cpp:31-33).  On the device every tile of TILE items computes its items and flags,
// ranks them with warp ballots, and obtains the number of successes in all EARLIER tiles through
// a decoupled look-back over 64-bit status words  [2-bit flag | 62-bit count]  — one pass over
// the data, no second kernel.  Tile ids are handed out by an atomic ticket so a tile can only
// ever wait on tiles that have already started (forward progress without co-residency
// assumptions).
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
__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void scan_st(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Executed by ONE full warp of the block.  `total` = successes in this tile (warp-uniform).
// Returns the number of successes in all earlier tiles.
__device__ __forceinline__ long long scan_lookback(unsigned long long* status, int tile,
                                                   long long total) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) scan_st(status, SCAN_PREFIX | (unsigned long long)total);
    return 0;
  }
  if (lane == 0) scan_st(status + tile, SCAN_AGG | (unsigned long long)total);
  long long excl = 0;
  int idx = tile - 1;
  while (true) {
    const int j = idx - lane;
    unsigned long long v = SCAN_PREFIX;  // "tiles before tile 0": prefix 0
    if (j >= 0) {
      do {
        v = scan_ld(status + j);
      } while ((v >> 62) == 0ull);
    }
    const unsigned pm = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
    const int first = pm ? (__ffs(pm) - 1) : 32;
    long long c = (lane <= first) ? (long long)(v & SCAN_MASK) : 0ll;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    excl += c;
    if (pm) break;
    idx -= 32;
  }
  if (lane == 0) scan_st(status + tile, SCAN_PREFIX | (unsigned long long)(excl + total));
  return excl;
}

// Block-wide helper: given each thread's per-item flags (ITEMS of them, item j of lane l of warp
// w sits at tile position w*32*ITEMS + j*32 + l) computes for every item its global output rank.
// Returns the exclusive prefix of the tile and the tile total through the out-params; rank[j] is
// only meaningful where flag[j] is set.
template <int THREADS, int ITEMS>
__device__ __forceinline__ void scan_tile_ranks(const ScanWorkspace& ws, int tile,
                                                const bool (&flag)[ITEMS], long long (&rank)[ITEMS],
                                                long long* tile_excl, int* tile_total) {
  constexpr int WARPS = THREADS / 32;
  __shared__ int s_warp_tot[WARPS];
  __shared__ long long s_excl;
  __shared__ int s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  int local[ITEMS];
  int run = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const unsigned bal = __ballot_sync(0xffffffffu, flag[j]);
    local[j] = run + __popc(bal & lt);
    run += __popc(bal);
  }
  if (lane == 0) s_warp_tot[warp] = run;
  __syncthreads();
  if (warp == 0) {
    int x = (lane < WARPS) ? s_warp_tot[lane] : 0;
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane < WARPS) s_warp_tot[lane] = incl - x;  // exclusive warp offsets
    const long long excl = scan_lookback(ws.status, tile, (long long)total);
    if (lane == 0) {
      s_excl = excl;
      s_total = total;
    }
  }
  __syncthreads();
  const long long base = s_excl + s_warp_tot[warp];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) rank[j] = base + local[j];
  *tile_excl = s_excl;
  *tile_total = s_total;
}
#endif

}  // namespace vo
