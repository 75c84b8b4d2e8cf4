// nn_tc.cu — the tensor-core filter of the exact appearance nearest neighbour (sm_100a, tcgen05).
//
// Same contract as nn.cu (bruteForceBestMatch, reference include/brute_force_search.h:22-41): the
// answer is decided by an exact re-rank in the reference's own FP32 rounding order; what changes is
// the FILTER that finds the few (query,row) pairs worth re-ranking.  nn.cu tests a 5-of-10-dimension
// partial distance with FP32 FMAs (10 executed flop per pair, data-dependent pruning).  Here the
// full 10-D distance of 128 queries x 128 rows is ONE tcgen05.mma:
//
//     A (queries, K-major f16)   a_i = [ -2q_0 .. -2q_9 | 1 | 1 | |q|^2_hi | |q|^2_lo | 0 0 ]
//     B (map rows, K-major f16)  b_j = [  m_0 ..  m_9   | |m|^2_hi | |m|^2_lo | 1 | 1 | 0 0 ]
//     D = A B^T  (F16, in TMEM)   d_ij = |m_j|^2 - 2 q_i.m_j + |q_i|^2  ~  |q_i - m_j|^2
//
// (K = 16 is one kind::f16 instruction; the norms travel as hi+lo f16 pairs so that only the
// rounding of the COORDINATES to f16, u = 2^-11, matters.)  The accumulator never leaves the chip
// as data.  It is kept in F16: an epilogue thread (= TMEM lane = query) reads it back two columns
// per register (tcgen05.ld ... .pack::16b), folds its 128 columns with the packed 16-bit 3-input
// integer minimum (VIMNMX3.S16x2 — the code of a non-negative f16 orders like a signed integer and
// every negative f16 is below every threshold) and compares the result with the CODE of the query's
// threshold
//     thr_i = best_i + eps_i ,   eps_i = 4.01 u |q_i| max|m| + (accumulation / norm slack) ,
// which is an upper bound of d_ij for every row the reference could accept (derivation and the
// measurements it rests on: DESIGN.md §4.1b, tools/tc_probe2.cu).  A warp with a flagged query
// re-scans those 64 rows from the FP32 rows in global memory exactly as nn.cu does (full FMA test,
// then the reference-order distance, 64-bit atomicMin key).  In ten dimensions the margin costs
// nothing: for uniform appearances the filter passes ~3e-10 of the pairs, and clustered
// descriptors only add re-scans, never a cliff.
//
// Pipeline per CTA (one per SM, persistent, 20 warps):
//   * 256-row f16 map tiles (8 KB, 1-D bulk copies) stream through an 8-stage ring, requested 6
//     tiles ahead by the first issuing thread;
//   * all 512 TMEM columns hold four 128x128 accumulators; accumulator b = (half `b & 1` of the map
//     tile, every second resident query tile) has its own MMA-issuing thread (one N=128 tcgen05.mma
//     per use, committed to an mbarrier) and its own four epilogue warps, one per TMEM lane quadrant;
//   * an epilogue warp waits for its accumulator, reads the 128 columns with two packed loads through
//     one 32-register buffer, releases the accumulator, folds (2 x 16 VIMNMX3.S16x2) and tests.
// Measured on hardware (tools/tc_probe.cu, tools/tc_probe2.cu, in-kernel counters with
// -DNN_TC_PROFILE): the f16 MMA takes N/2 cycles; a thread needs ~100 cycles to issue one and ~60 for
// the commit whatever the shape, several threads issue in parallel; issue -> barrier-visible latency
// 288 cycles; packed TMEM read + fold with 16 warps 175 elements per clock per SM.  What binds is
// TMEM capacity x latency: an f16 accumulator still occupies a 32-bit cell, so 65536 cells are all
// that is ever in flight, and one accumulator is busy for ~760 cycles from the issue of its MMA until
// its last column has been read: 86 pairs per clock per SM, 0.397 s for 1e13 pairs (ncu: tensor pipe
// 36 %, ALU 56 %, issue slots 49 %, profiles/r02p_ncu_nn_tc.md).
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "nn.cuh"

// -DNN_TC_PROFILE: cycle counters of one epilogue warp and one MMA warp of CTA 0 in stats[1..6]
// (tools/nn_tc_time.py prints them): 1 wait accumulator, 2 tcgen05.ld, 3 fold + threshold,
// 4 MMA warp waits for the map tile, 5 waits for a free accumulator, 6 issues.
#ifdef NN_TC_PROFILE
#define TC_PROF_T(v) const long long v = clock64()
#define TC_PROF_ADD(slot, v) prof[slot] += clock64() - (v)
#else
#define TC_PROF_T(v)
#define TC_PROF_ADD(slot, v)
#endif

namespace vo {

constexpr int TC_BN = 256;                  // map rows per smem tile
constexpr int TC_SUB = 128;                 // rows per MMA (N) and per TMEM buffer
constexpr int TC_ROW_BYTES = 32;            // 16 halves
constexpr uint32_t TC_TILE_BYTES = TC_BN * TC_ROW_BYTES;   // 8 KB
constexpr int TC_RESCAN = 64;                // rows one epilogue warp answers for (its column half)
constexpr int TC_STAGES = 8;
constexpr int TC_LOOKAHEAD = 6;             // map tiles requested ahead of the one being multiplied
constexpr int TC_QT_MAX = 16;               // query tiles (of 128) resident per CTA
constexpr uint32_t TC_A_BYTES = 128 * TC_ROW_BYTES;        // 4 KB per query tile
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_PIPES = TC_BN / TC_SUB;    // 2 independent MMA->epilogue pipelines (halves of a map tile)
constexpr int TC_PIPE_WARPS = TC_EPI_WARPS / TC_PIPES;     // 8 epilogue warps per pipeline
constexpr int TC_BUFS = 4;                  // TMEM accumulator buffers of TC_SUB columns; pipeline p owns p, p+2
#ifndef TC_HALVES_N
#define TC_HALVES_N 1
#endif
// TC_HALVES = 2 fills and releases a buffer in two independent column halves (one MMA of N = 64 each,
// own full/empty barriers) so that the refill of the first half is under way while the second is
// still being read, with one issuing warp per half (24 warps; setmaxnreg moves the issuers' registers
// to the epilogue warps).  Measured SLOWER twice: 433 against 397 ms at the headline size with four
// issuers (two MMAs per accumulator make the issuers the limit again), 88.5 against 79.1 ms at
// M = 2e7 with eight.  Kept as a switch.
constexpr int TC_HALVES = TC_HALVES_N;
constexpr int TC_HCOLS = TC_SUB / TC_HALVES;

// + one MMA-issuing warp per buffer (the first also requests the map tiles).  20 warps = 5 per
// scheduler partition of the register file: 96 registers per thread (a 21st warp would cap all at 80)
constexpr int TC_ISSUERS = TC_BUFS * TC_HALVES;  // one MMA-issuing warp per separately released part
constexpr int TC_THREADS = (TC_EPI_WARPS + TC_ISSUERS) * 32;
constexpr float TC_U16 = 4.8828125e-4f;     // 2^-11, unit roundoff of f16 (round to nearest)
constexpr float TC_PAD_NORM = 60000.f;      // |m|^2 of a padding row: never under any threshold
constexpr float TC_MAX_NORM = 30000.f;      // |m|^2, |q|^2 above this do not fit f16 arithmetic

constexpr size_t TC_SMEM_A = (size_t)TC_QT_MAX * TC_A_BYTES;            // 64 KB
constexpr size_t TC_SMEM_B = (size_t)TC_STAGES * TC_TILE_BYTES;         // 32 KB
constexpr size_t TC_SMEM_THR = (size_t)TC_QT_MAX * 128 * sizeof(float); // 8 KB (x2: thr, eps)
constexpr size_t TC_SMEM_BYTES = TC_SMEM_A + TC_SMEM_B + 2 * TC_SMEM_THR + 512;

struct NNTCParams {
  const unsigned char* tiles16;  // f16 map tiles, TC_TILE_BYTES each, 8-row interleaved
  const float4* packed;          // FP32 rows (48 B) for the exact re-rank
  int64_t n_rows;
  int64_t n_rows_packed;         // rows the FP32 buffer holds (padding included)
  int64_t n_tiles16;
  const float* queries;
  int64_t n_queries;
  int query_stride;
  int skip;
  float bound;                   // norm*norm
  const float* mm_max;
  unsigned long long* keys;
  int qt;                        // query tiles per group (<= TC_QT_MAX)
  int n_groups;
  unsigned long long* stats;     // [0] = re-scans (flagged (query, 128-row block) pairs)
};

// ---- tcgen05 / TMEM wrappers ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// mbarrier arrive when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\n"  // p = false: D = A*B (no accumulation)
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
// 64 columns of an f16 accumulator: two adjacent columns (their low halves) per register
__device__ __forceinline__ void tc_ld32_pack(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// packed signed 16-bit minima: ptxas fuses the pair into ONE VIMNMX3.S16x2 (four new values per
// instruction on the ALU pipe, against two for FMNMX3)
__device__ __forceinline__ uint32_t tc_min3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("{\n.reg .b32 t;\nmin.s16x2 t, %1, %2;\nmin.s16x2 %0, t, %3;\n}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t tc_min_s16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t tc_max_s16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// minimum of 64 packed accumulator columns (even columns in the low halves, odd ones in the high
// halves), four independent chains
__device__ __forceinline__ uint32_t tc_fold16(const uint32_t (&r)[32]) {
  uint32_t mn[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) mn[c] = tc_min3_s16x2(r[c], r[4 + c], r[8 + c]);
#pragma unroll
  for (int c = 0; c < 4; ++c) mn[c] = tc_min3_s16x2(mn[c], r[12 + c], r[16 + c]);
#pragma unroll
  for (int c = 0; c < 4; ++c) mn[c] = tc_min3_s16x2(mn[c], r[20 + c], r[24 + c]);
  mn[0] = tc_min3_s16x2(mn[0], r[28], r[29]);
  mn[1] = tc_min3_s16x2(mn[1], r[30], r[31]);
  return tc_min_s16x2(tc_min3_s16x2(mn[0], mn[1], mn[2]), mn[3]);
}
// The f16 accumulator is compared as a SIGNED 16-BIT INTEGER: the code of a non-negative f16 grows
// with its value (+inf = 0x7C00 above every finite one) and every negative f16 (distances that came
// out below zero: true matches) is a negative integer, below every threshold.  tc_thr16 returns c,
// in both halves of a word, such that an accumulator is worth a re-scan iff code < c:
//   thr = +inf (norms too large for f16: always re-scan) -> 0x7FFF, above +inf's code;
//   thr = -inf (non-finite query: never)                 -> -32768, below everything;
//   otherwise code(thr rounded UP to f16) + 3.  Rounding is monotone, so a value <= thr has a code
//   <= code(ru(thr)) if the tensor core rounds its internal sum once; measured (tools/tc_probe2.cu,
//   4e6 elements) its f16 result is within ONE code of the rounded f32 result it would have
//   delivered — two codes of slack (1.5e-5 at 0.016) on top.
__device__ __forceinline__ uint32_t tc_thr16(float thr) {
  int c;
  if (thr == INFINITY) c = 0x7FFF;
  else if (!(thr > 0.f)) c = -32768;
  else c = min((int)__half_as_ushort(__float2half_ru(thr)) + 3, 0x7FFF);
  return ((uint32_t)c & 0xFFFFu) * 0x10001u;
}
// Shared-memory matrix descriptor: K-major, no swizzle.  A core matrix is 8 rows x 16 bytes stored
// as 128 contiguous bytes; the two 16-byte K chunks of one instruction are `lbo` bytes apart,
// consecutive 8-row groups `sbo` bytes (layout verified on hardware by tools/tc_probe.cu).
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
  constexpr uint64_t lbo = 128, sbo = 256;
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t tc_desc_hi() {  // everything but the start address
  return ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = F16 (bits 4-5 = 0; 1 would be F32), A = B = F16 (0), both K-major,
// N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t TC_IDESC_D16 = ((uint32_t)(TC_HCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// byte offset of (row r, K chunk c) inside an interleaved operand tile
__host__ __device__ __forceinline__ uint32_t tc_row_off(int r, int c) {
  return (uint32_t)((r >> 3) * 256 + c * 128 + (r & 7) * 16);
}

// x = hi + lo with hi, lo representable in f16 (|x - hi - lo| <= 2^-22 |x| + 2^-25)
__device__ __forceinline__ void tc_split(float x, __half* hi, __half* lo) {
  const __half h = __float2half_rn(x);
  *hi = h;
  *lo = __float2half_rn(x - __half2float(h));
}

// ---- map tiles: FP32 packed rows -> interleaved f16 operand tiles -----------------------------------
__global__ void __launch_bounds__(256)
nn_tc_pack_kernel(const float4* __restrict__ packed, int64_t n_rows, int64_t n_rows_tc,
                  unsigned char* __restrict__ tiles16) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows_tc) return;
  __align__(16) __half v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = __float2half_rn(0.f);
  if (r < n_rows) {
    const float4 a = packed[r * 3 + 0], b = packed[r * 3 + 1], c = packed[r * 3 + 2];
    const float m[NN_DIM] = {a.x, a.y, a.z, a.w, b.x, b.y, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < NN_DIM; ++k) v[k] = __float2half_rn(m[k]);
    tc_split(b.w, &v[10], &v[11]);  // |m|^2 as computed by the FP32 re-pack
  } else {
    v[10] = __float2half_rn(TC_PAD_NORM);
  }
  v[12] = v[13] = __float2half_rn(1.f);  // multiply |q|^2_hi, |q|^2_lo
  unsigned char* tile = tiles16 + (r / TC_BN) * (int64_t)TC_TILE_BYTES;
  const int rr = (int)(r % TC_BN);
  *reinterpret_cast<uint4*>(tile + tc_row_off(rr, 0)) = *reinterpret_cast<const uint4*>(&v[0]);
  *reinterpret_cast<uint4*>(tile + tc_row_off(rr, 1)) = *reinterpret_cast<const uint4*>(&v[8]);
}

// Soundness margin of the f16 tensor-core distance for one query:
//   |d_tc - d2_ref| <= 2 * (2u + u^2) |q||m|          coordinates rounded to f16, u = 2^-11
//                    + 2^-18 (|q|^2 + |m|^2 + 2|q||m|)  norms' hi/lo residual, their FP32 rounding
//                                                       and the tensor core's FP32 accumulation
//                    + 2^-21 (|q| + |m|)                 f16 underflow of tiny coordinates
//                    + nn_eps                            the reference's own rounding of d2
__device__ __forceinline__ float tc_eps(float qq, float bound, float mm_max) {
  const float qm = sqrtf(qq * mm_max) * 1.0000002f;
  return 4.01f * TC_U16 * qm + 3.8146973e-6f * (qq + mm_max + 2.f * qm) +
         4.7683716e-7f * (sqrtf(qq) + sqrtf(mm_max)) + nn_eps(qq, bound, mm_max);
}

// Exact re-rank of ONE query against TC_RESCAN consecutive rows, executed by the whole warp (lane l takes
// rows l, l+32, ...) from the FP32 rows in global memory: the full 10-D FMA test first, then the
// candidates in the REFERENCE order, merged into the query's 64-bit key (strict minimum, lowest row
// on ties — brute_force_search.h:30-40).  Returns the best exact d2 known for the query.
__device__ __forceinline__ float tc_rescan_warp(const float4* __restrict__ packed, int64_t row0,
                                                int64_t n_rows, const float* __restrict__ query,
                                                float radius2, float mm_max, unsigned long long* key) {
  const int lane = threadIdx.x & 31;
  float q[NN_DIM], qn[NN_DIM];
#pragma unroll
  for (int k = 0; k < NN_DIM; ++k) {
    q[k] = __ldg(query + k);
    qn[k] = -2.f * q[k];
  }
  float qq6, qq;
  nn_query_norms(qn, &qq6, &qq);
  const unsigned long long k0 = *reinterpret_cast<volatile unsigned long long*>(key);
  const float best = (k0 == NN_KEY_NONE) ? radius2 : __uint_as_float(static_cast<unsigned int>(k0 >> 32));
  const float tq = nn_threshold_full(qq, best, mm_max);
  float found = best;
#pragma unroll
  for (int r = lane; r < TC_RESCAN; r += 32) {
    const int64_t row = row0 + r;
    if (row < n_rows) {
      const float4 a = __ldg(packed + row * 3 + 0), b = __ldg(packed + row * 3 + 1),
                   c = __ldg(packed + row * 3 + 2);
      float acc = b.w;
      acc = fmaf(qn[0], a.x, acc);
      acc = fmaf(qn[1], a.y, acc);
      acc = fmaf(qn[2], a.z, acc);
      acc = fmaf(qn[3], a.w, acc);
      acc = fmaf(qn[4], b.x, acc);
      acc = fmaf(qn[5], b.y, acc);
      acc = fmaf(qn[6], c.x, acc);
      acc = fmaf(qn[7], c.y, acc);
      acc = fmaf(qn[8], c.z, acc);
      acc = fmaf(qn[9], c.w, acc);
      if (acc < tq) {
        const float m[NN_DIM] = {a.x, a.y, a.z, a.w, b.x, b.y, c.x, c.y, c.z, c.w};
        const float d2 = ref_sqdist<NN_DIM>(m, q);
        if (d2 < radius2 && d2 <= best) {
          found = fminf(found, d2);
          atomicMin(key, nn_pack_key(d2, static_cast<uint32_t>(row)));
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) found = fminf(found, __shfl_xor_sync(0xffffffffu, found, o));
  return found;
}

// The rare path of the epilogue (inlined: a call boundary measured 1.5 % slower): the warp takes its
// flagged queries (lane = query) one by one and re-scans the 64 rows starting at row0 for each.  q0 = index of lane 0's query; thr / eps = the 32 thresholds and
// margins of this warp's queries in shared memory.
__device__ __forceinline__ void tc_rescan_flagged(const NNTCParams& p, bool flagged, int64_t q0, int64_t row0,
                                               float mm_max, uint32_t* thr, const float* eps) {
  const int lane = threadIdx.x & 31;
  unsigned pending = __ballot_sync(0xffffffffu, flagged);
  while (pending) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1;
    const int64_t qi = q0 + src;
    const float found = tc_rescan_warp(p.packed, row0, p.n_rows, p.queries + qi * (int64_t)p.query_stride + p.skip,
                                       p.bound, mm_max, p.keys + qi);
    if (lane == 0) {
      // later rows only matter if they can reach d2 <= found.  Four warps share a query's threshold;
      // a lost update only leaves it higher than necessary (more re-scans).
      const float nt_thr = found + eps[src];
      // thresholds are kept as packed 16-bit codes (tc_thr16)
      const uint32_t c = tc_thr16(nt_thr);
      if ((short)(c & 0xFFFFu) < (short)(thr[src] & 0xFFFFu)) thr[src] = c;
      atomicAdd(p.stats, 1ull);
    }
  }
  __syncwarp();
}

// The body of the kernel for one role: ISSUER = the MMA-issuing (and TMA-requesting) warps, otherwise
// the epilogue warps.  Both roles walk the same segments and meet at the same block-wide barriers; the
// split exists so that each role's code is dominated by its own setmaxnreg (TC_HALVES = 2).
template <bool ISSUER>
__device__ __forceinline__ void tc_role(const NNTCParams& p, unsigned char* sA, unsigned char* sB, uint32_t* thr_s,
                                        float* eps_s, uint64_t* full, uint64_t* empty, uint64_t* tfull,
                                        uint64_t* tempty) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = p.qt;
  const float mm_max = __ldg(p.mm_max);

  // this CTA's share of the (query group, map tile) units, group-major
  const int64_t units = (int64_t)p.n_groups * p.n_tiles16;
  int64_t u = units * blockIdx.x / gridDim.x;
  const int64_t u_end = units * (blockIdx.x + 1) / gridDim.x;

#ifdef NN_TC_PROFILE
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  // running counters, identical in every role
  uint32_t unit_n = 0;   // map tiles streamed so far (position in the smem ring)
  uint32_t acc_n = 0;    // accumulator tiles each pipeline has produced so far (= unit_n * qt)

  while (u < u_end) {
    const int g = (int)(u / p.n_tiles16);
    const int64_t t0 = u % p.n_tiles16;
    const int64_t nt = min(p.n_tiles16 - t0, u_end - u);  // tiles of this segment
    u += nt;

    // ---- segment prologue: this group's queries -> f16 A tiles + thresholds -------------------------
    __syncthreads();  // every accumulator of the previous segment has been drained
    const int64_t qbase = (int64_t)g * qt * 128;
    if (!ISSUER) {
      for (int i = tid; i < qt * 128; i += TC_EPI_WARPS * 32) {
        const int64_t qi = qbase + i;
        __align__(16) __half v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __float2half_rn(0.f);
        float thr = -INFINITY, eps = 0.f;
        if (qi < p.n_queries) {
          const float* src = p.queries + qi * (int64_t)p.query_stride + p.skip;
          float qn[NN_DIM];
#pragma unroll
          for (int k = 0; k < NN_DIM; ++k) qn[k] = -2.f * __ldg(src + k);
          float qq6, qq;
          nn_query_norms(qn, &qq6, &qq);
          if (qq <= TC_MAX_NORM) {
#pragma unroll
            for (int k = 0; k < NN_DIM; ++k) v[k] = __float2half_rn(qn[k]);
            v[10] = v[11] = __float2half_rn(1.f);  // multiply |m|^2_hi, |m|^2_lo
            tc_split(qq, &v[12], &v[13]);
            eps = tc_eps(qq, p.bound, mm_max);
            thr = p.bound + eps;
          } else if (qq < INFINITY) {
            thr = INFINITY;  // too large for f16: every block is re-scanned exactly (A row stays 0)
          }                  // non-finite query: no row can satisfy d2 < norm^2, nothing to do
        }
        unsigned char* dst = sA + (i >> 7) * TC_A_BYTES;
        *reinterpret_cast<uint4*>(dst + tc_row_off(i & 127, 0)) = *reinterpret_cast<const uint4*>(&v[0]);
        *reinterpret_cast<uint4*>(dst + tc_row_off(i & 127, 1)) = *reinterpret_cast<const uint4*>(&v[8]);
        thr_s[i] = tc_thr16(thr);
        eps_s[i] = eps;
      }
      // the tensor core reads shared memory through the async proxy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (ISSUER) {
      // ===== MMA issuer of buffer `buf` = pipeline `pipe` (rows [pipe*128, +128) of every map tile),
      // every second accumulator of that pipeline.  ONE thread runs the whole loop.  A thread needs
      // ~100 cycles to issue a tcgen05.mma and ~60 for the commit whatever the shape (in-kernel
      // counters, make TC_PROFILE=1), several threads issue in parallel: with one issuer per pipeline
      // the issuing threads were the limit of the whole kernel (430 cycles per accumulator), hence one
      // per buffer.
      const int part = warp - TC_EPI_WARPS;  // (buffer, column half) this thread fills
      const int buf = part / TC_HALVES, hh = part % TC_HALVES;
      const int pipe = buf % TC_PIPES;
      const uint32_t bsel = (uint32_t)(buf / TC_PIPES);
      if (lane == 0) {
        const uint32_t a_desc0 = (uint32_t)((smem_u32(sA) >> 4) & 0x3FFF);
        // accumulator f of this segment (f = tile * qt + a) is number acc_n + f of the pipeline and lands
        // in buffer (acc_n + f) & 1; k counts the uses of this buffer
        int a = (int)((acc_n ^ bsel) & 1u);
        uint32_t k = (acc_n + (uint32_t)a) >> 1;
        // buffer 0's thread is also the TMA producer: 256-row map tiles, TC_LOOKAHEAD ahead of the
        // one being multiplied, into a ring of TC_STAGES (a stage is free once all four issuers have
        // committed the MMAs that read it)
        auto request = [&](int64_t i) {
          const uint32_t n = unit_n + (uint32_t)i;
          const int s = (int)(n % TC_STAGES);
          mbar_wait(&empty[s], ((n / TC_STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full[s], TC_TILE_BYTES);
          tma_load_1d(sB + s * TC_TILE_BYTES, p.tiles16 + (t0 + i) * (int64_t)TC_TILE_BYTES, TC_TILE_BYTES, &full[s]);
        };
        if (part == 0)
          for (int64_t i = 0; i < min(nt, (int64_t)TC_LOOKAHEAD); ++i) request(i);
        for (int64_t i = 0; i < nt; ++i) {
          if (part == 0 && i + TC_LOOKAHEAD < nt) request(i + TC_LOOKAHEAD);
          const uint32_t n = unit_n + (uint32_t)i;
          const int s = (int)(n % TC_STAGES);
          TC_PROF_T(pa);
          mbar_wait(&full[s], (n / TC_STAGES) & 1u);
          TC_PROF_ADD(4, pa);
          tc_fence_after();
          const uint64_t bdesc =
              tc_desc(smem_u32(sB + s * TC_TILE_BYTES) + pipe * (TC_SUB * TC_ROW_BYTES) + hh * (TC_HCOLS * TC_ROW_BYTES));
          for (; a < qt; a += 2, ++k) {
            const uint64_t adesc = tc_desc_hi() | (uint64_t)(a_desc0 + a * (TC_A_BYTES >> 4));
            TC_PROF_T(pb);
            mbar_wait(&tempty[part], (k & 1u) ^ 1u);
            TC_PROF_ADD(5, pb);
            tc_fence_after();
            TC_PROF_T(pd);
            tc_mma_f16((uint32_t)(buf * TC_SUB + hh * TC_HCOLS), adesc, bdesc, TC_IDESC_D16);
            TC_PROF_ADD(7, pd);
            TC_PROF_T(pe);
            tc_commit(&tfull[part]);
            TC_PROF_ADD(3, pe);
          }
          a -= qt;
          tc_commit(&empty[s]);  // the stage is free once every MMA above has read it
        }
      }
      __syncwarp();
    } else {
      // ===== f16-accumulator epilogue.  Of the 8 warps of pipeline `pipe`, warps 0-3 (one per lane
      // quadrant) drain buffer `pipe`, warps 4-7 buffer `pipe + 2`: a warp takes every SECOND
      // accumulator of its pipeline and reads all 128 columns of it — two packed loads of 64 columns
      // through one 32-register buffer, 16 VIMNMX3.S16x2 each.  Per 4096 pairs a warp executes about
      // as many instructions as the f32-accumulator epilogue of the first version needed for 2048.
      const int pipe = warp / TC_PIPE_WARPS, w8 = warp % TC_PIPE_WARPS;
      const int quad = w8 & 3;
      const uint32_t bsel = (uint32_t)(w8 >> 2);
      const int buf = pipe + TC_PIPES * (int)bsel;
      uint32_t t_addr = ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * TC_SUB;
      uint32_t full_addr = smem_u32(&tfull[buf * TC_HALVES]), empty_addr = smem_u32(&tempty[buf * TC_HALVES]);
      uint32_t thr_addr0 = smem_u32(thr_s + quad * 32 + lane);
      asm volatile("mov.u32 %0, %0;\nmov.u32 %1, %1;\nmov.u32 %2, %2;\nmov.u32 %3, %3;"
                   : "+r"(t_addr), "+r"(full_addr), "+r"(empty_addr), "+r"(thr_addr0));
      const uint32_t arrive_lane = lane;
      // accumulator f of this segment (f = unit * qt + a) is number acc_n + f of the pipeline and lands
      // in buffer (acc_n + f) & 1
      const uint32_t total = (uint32_t)(nt * qt);
      uint32_t f = (acc_n ^ bsel) & 1u;
      uint32_t par = ((acc_n + f) >> 1) & 1u;
      int a = (int)f;
      int64_t unit = 0;
      while (a >= qt) {
        a -= qt;
        ++unit;
      }
      auto wait_full = [&](uint32_t addr) {
        uint32_t ok;
        do {
          asm volatile(
              "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
              : "=r"(ok)
              : "r"(addr), "r"(par)
              : "memory");
        } while (!ok);
      };
      // the half may be overwritten as soon as its four warps have read it
      auto release = [&](uint32_t addr) {
        tc_fence_before();
        __syncwarp();
        asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %1, 0;\n@p mbarrier.arrive.shared::cta.b64 _, [%0];\n}" ::"r"(addr),
                     "r"(arrive_lane)
                     : "memory");
      };
      for (; f < total; f += 2) {
        TC_PROF_T(pa);
        uint32_t my_thr;  // packed code, read before the wait: its latency hides behind it
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(my_thr) : "r"(thr_addr0 + (uint32_t)a * 512u) : "memory");
        wait_full(full_addr);
        TC_PROF_ADD(1, pa);
        TC_PROF_T(pb);
        tc_fence_after();
        uint32_t mA, mB;
        {
          uint32_t r[32];
          tc_ld32_pack(t_addr, r);
          tc_wait_ld();
          if (TC_HALVES == 2) {
            release(empty_addr);
            mA = tc_fold16(r);
            wait_full(full_addr + 8u);
            tc_fence_after();
          } else {
            mA = tc_fold16(r);
          }
          tc_ld32_pack(t_addr + 64, r);
          tc_wait_ld();
          release(empty_addr + (TC_HALVES == 2 ? 8u : 0u));
          TC_PROF_ADD(2, pb);
          mB = tc_fold16(r);
        }
        TC_PROF_T(pc);
        par ^= 1u;
        // some half below its threshold code  <=>  max(m, c) != m
        const uint32_t mAB = tc_min_s16x2(mA, mB);
        if (__any_sync(0xffffffffu, tc_max_s16x2(mAB, my_thr) != mAB)) {
          const int64_t row0 = (t0 + unit) * TC_BN + pipe * TC_SUB;
          tc_rescan_flagged(p, tc_max_s16x2(mA, my_thr) != mA, qbase + a * 128 + quad * 32, row0, mm_max,
                                  thr_s + a * 128 + quad * 32, eps_s + a * 128 + quad * 32);
          tc_rescan_flagged(p, tc_max_s16x2(mB, my_thr) != mB, qbase + a * 128 + quad * 32, row0 + 64, mm_max,
                                  thr_s + a * 128 + quad * 32, eps_s + a * 128 + quad * 32);
        }
        TC_PROF_ADD(3, pc);
        a += 2;
        while (a >= qt) {
          a -= qt;
          ++unit;
        }
      }
    }
    unit_n += (uint32_t)nt;
    acc_n += (uint32_t)(nt * qt);
  }

#ifdef NN_TC_PROFILE
  if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == TC_EPI_WARPS))
    for (int i = 1; i < 8; ++i)
      if (prof[i]) atomicAdd(p.stats + (warp == 0 ? 0 : 8) + i, (unsigned long long)prof[i]);
#endif
}

__global__ void __launch_bounds__(TC_THREADS, 1) nn_tc_filter_kernel(const NNTCParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char* sA = tc_smem;
  unsigned char* sB = tc_smem + TC_SMEM_A;
  uint32_t* thr_s = reinterpret_cast<uint32_t*>(tc_smem + TC_SMEM_A + TC_SMEM_B);  // [qt][128] packed codes
  float* eps_s = reinterpret_cast<float*>(thr_s + TC_QT_MAX * 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(eps_s + TC_QT_MAX * 128);
  uint64_t* full = bars;                       // [TC_STAGES]  TMA bytes landed
  uint64_t* empty = full + TC_STAGES;          // [TC_STAGES]  every MMA reading the stage is done
  uint64_t* tfull = empty + TC_STAGES;               // [TC_BUFS * TC_HALVES]  accumulator half written
  uint64_t* tempty = tfull + TC_BUFS * TC_HALVES;    // [TC_BUFS * TC_HALVES]  drained by its four warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + TC_BUFS * TC_HALVES);

  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TC_ISSUERS);  // one commit per MMA warp
    }
#pragma unroll
    for (int b = 0; b < TC_BUFS * TC_HALVES; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4);  // one warp per lane quadrant
    }
    mbar_fence_init();
  }
  if (warp == TC_EPI_WARPS) tmem_alloc(tmem_slot, 512);  // the whole tensor memory of the SM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // 512 columns is all of it, so the allocation starts at column 0 / lane 0: addresses below are
  // plain constants (and the issuing warp keeps them in uniform registers)
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0u) __trap();
  if (warp >= TC_EPI_WARPS) {
#if TC_HALVES_N == 2
    // 24 warps: the launch gives every thread 80 registers; the issuing warpgroups hand theirs back
    // and the epilogue warpgroups take them
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
#endif
    tc_role<true>(p, sA, sB, thr_s, eps_s, full, empty, tfull, tempty);
  } else {
#if TC_HALVES_N == 2
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
#endif
    tc_role<false>(p, sA, sB, thr_s, eps_s, full, empty, tfull, tempty);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

}  // namespace vo

using namespace vo;

// ---- host side (called from nn.cu) ----------------------------------------------------------------
int nn_tc_pack(vo_nn_s* h) {
  h->n_tiles16 = (h->n_rows + TC_BN - 1) / TC_BN;
  const int64_t rows_tc = h->n_tiles16 * TC_BN;
  int rc = h->tiles16.reserve((size_t)h->n_tiles16 * TC_TILE_BYTES);
  if (rc) return rc;
  rc = h->tc_stats.reserve(128);
  if (rc) return rc;
  const int threads = 256;
  nn_tc_pack_kernel<<<(unsigned)((rows_tc + threads - 1) / threads), threads, 0, h->stream>>>(
      h->packed.as<float4>(), h->n_rows, rows_tc, h->tiles16.as<unsigned char>());
  VO_LAUNCH_CHECK();
  return VO_OK;
}

// query tiles per group.  The work of a launch is (groups x tiles per group) accumulators per map
// tile, so what matters is the padding of the last group: a sharded batch of 25 000 queries = 196
// tiles runs as 14 groups of 14 (13 groups of 16 would compute 208), 12 500 queries = 98 tiles as 7
// groups of 14.  More groups only mean more passes over the f16 map (HBM is at 6 %).  Ties go to the
// larger group.  (For huge maps groups * tiles overflows nothing: both are < 2^31.)
static int tc_pick_qt(int64_t n_qtiles, int64_t n_tiles16, int ctas) {
  int best_qt = TC_QT_MAX;
  int64_t best_cost = INT64_MAX;
  for (int qt = TC_QT_MAX; qt >= TC_QT_MAX / 2; --qt) {
    const int64_t groups = (n_qtiles + qt - 1) / qt;
    // accumulators of the busiest CTA: the (group, map tile) units are dealt out in equal contiguous
    // shares.  For a large map this is groups * qt * tiles / ctas; for a frame-sized one (40 map
    // tiles, 5 groups) the rounding of the shares decides.
    const int64_t cost = ((groups * n_tiles16 + ctas - 1) / ctas) * qt;
    if (cost < best_cost) {
      best_cost = cost;
      best_qt = qt;
    }
  }
  return (int)std::min<int64_t>(best_qt, std::max<int64_t>(n_qtiles, 1));
}

int nn_tc_launch(vo_nn_s* h, const float* queries_dev, int64_t nq, int qstride, float bound) {
  NNTCParams p;
  p.tiles16 = h->tiles16.as<unsigned char>();
  p.packed = h->packed.as<float4>();
  p.n_rows = h->n_rows;
  p.n_rows_packed = h->n_tiles * NN_TM;
  p.n_tiles16 = h->n_tiles16;
  p.queries = queries_dev;
  p.n_queries = nq;
  p.query_stride = qstride;
  p.skip = h->skip;
  p.bound = bound;
  p.mm_max = h->scalars.as<float>();
  p.keys = h->keys.as<unsigned long long>();
  const int64_t n_qtiles = (nq + 127) / 128;
  const int sms = num_sms(h->device);
  p.qt = tc_pick_qt(n_qtiles, p.n_tiles16, sms);
  p.n_groups = (int)((n_qtiles + p.qt - 1) / p.qt);
  p.stats = h->tc_stats.as<unsigned long long>();
  if (!h->tc_opted_in) {
    VO_CUDA(cudaFuncSetAttribute(nn_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)TC_SMEM_BYTES));
    h->tc_opted_in = true;
  }
  VO_CUDA(cudaMemsetAsync(h->tc_stats.p, 0, 128, h->stream));
  const int64_t units = (int64_t)p.n_groups * p.n_tiles16;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(sms, units));
  nn_tc_filter_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, h->stream>>>(p);
  VO_LAUNCH_CHECK();
  // (queries per thread = 0 marks the tensor-core filter, threads, query groups, CTAs)
  h->last_launches.insert(h->last_launches.end(), {0, TC_THREADS, (int32_t)p.n_groups, (int32_t)grid});
  return VO_OK;
}
