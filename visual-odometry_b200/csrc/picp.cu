// picp.cu — projective-ICP Gauss-Newton rounds, entirely on the device (sm_100a).
//
// Replaces PICPSolver::{init,errorAndJacobian,linearize,oneRound}  (reference
// src/picp_solver.cpp:16-112) with Camera::projectPoint inlined (include/camera.h:25-37) and
// v2tEuler / Rotation{X,Y,Z} / skew (include/utils.h:16-102).
//
// One launch == ALL the rounds of a vo_picp_compute() call; three kernels by problem size:
//   * picp_resident_kernel<.., GRID=false>  (<= 65536 correspondences, every VO frame): one
//     thread-block cluster keeps the correspondences in shared memory, DSMEM reduction;
//   * picp_resident_kernel<.., GRID=true>   (<= SMs x 8192): the same across the whole chip,
//     cooperative launch, partial rows through L2 and a grid barrier;
//   * picp_stream_kernel                    (larger): cooperative launch, the correspondences are
//     streamed from HBM through a per-thread cp.async ring every round.
// Shared arithmetic (picp_point2), per correspondence: pc = T*p, z-range / image-bounds rejection,
// e = proj - meas, J = Jp*K*[I | skew(-pc)] in the factored form  A = iz*(K_row - uv*K_row2),
// J = [A | A*S]  (K is treated as a general 3x3), robust weight, and the 21 upper-triangular
// entries of lambda*J^T J plus the 6 of lambda*J^T e accumulated in registers — two correspondences
// at a time in the lanes of packed FP32 pairs.  Reductions are fixed-order (no float atomics:
// results are run-to-run deterministic); the damping, the pivoted 6x6 LDL^T solve, v2tEuler(dx)
// and the pose update run on the device, so there is no host round-trip between rounds.
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>


#include <algorithm>

#include <cooperative_groups.h>

#include "common.cuh"
#include "linalg.cuh"

namespace vo {

#ifndef PICP_THREADS_N
#define PICP_THREADS_N 384
#endif
#ifndef PICP_DEPTH_N
#define PICP_DEPTH_N 3
#endif
constexpr int PICP_THREADS = PICP_THREADS_N;
#ifndef PICP_UNROLL_N
#define PICP_UNROLL_N 4
#endif
constexpr int PICP_UNROLL = PICP_UNROLL_N;
constexpr int PICP_NACC = 32;  // 21 H + 6 b + chi_in + chi_out + n_in (as float bits of int) + 2 pad

struct PicpDeviceState {
  vo_picp_state s;       // what vo_picp_get_state copies back
};

struct PicpParams {
  const float* __restrict__ world;   // 3 floats / point
  const float* __restrict__ image;   // 2 floats / point
  const int2* __restrict__ pairs;    // (image idx, world idx)
  int64_t n_pairs;
  float K[9];                        // column-major
  float z_near, z_far;               // ints promoted to float (camera.h:28)
  float max_u, max_v;                // cols-1, rows-1            (camera.h:32-35)
  float thr, damping;
  int min_inliers;
  int keep_outliers;
  PicpDeviceState* st;
  float* partials;                   // [2][gridDim.x][PICP_NACC]
  unsigned int* barrier;             // grid-barrier arrival counter, zeroed before every launch
  int early_out;                     // resident kernel: stop iterating once the pose sequence repeats
  // frame-pipeline extensions (resident kernel only)
  const int* n_pairs_dev;            // if set: the correspondence count lives on the device
  int has_pre;                       // if set: world points are moved by `pre` while gathered
  float pre[12];                     //   3x3 linear (column-major) then translation
};

// Two correspondences at a time: errorAndJacobian (:25-53) + the body of linearize's loop (:62-95).
// Branch-free: a rejected point runs the same instructions on zeroed operands (selected, not
// multiplied by a zero weight), so a warp never diverges on the data and nothing non-finite that
// a rejected point produced can reach the sums.
//
// Packed arithmetic.  At 28 B per point the kernel is HBM-bound only if its instruction stream
// stays well under the issue rate 6.5 TB/s implies; the scalar formulation (136 instructions per
// point, 88 of them FP32) was issue-bound (ncu: 70 % issue-active, "not selected" the top stall).
// The two lanes of sm_100a's packed FP32 pairs (fma/mul/add.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2,
// one issue slot for two operations) therefore carry TWO POINTS: every FP32 instruction of the
// per-point math is issued once per pair.  Each lane performs exactly the scalar operation
// sequence, so results do not depend on which lane a point lands in; the accumulators are
// lane-split partial sums, added together once after the loop.
//
// PINHOLE == true is selected by the host when K is exactly [fx 0 cx; 0 fy cy; 0 0 1] (the only
// form the reference's mains ever build: picp_solver_test.cpp:52-54, camera.dat).  Every product
// with a structural zero of K, and every term of J^T J that contains one, is dropped at compile
// time; the surviving operations are the same FMAs in the same order, so the result is
// bit-identical to the general-K path (tests/test_picp_gpu.py checks this).
typedef unsigned long long f2_t;

struct PicpAcc {
  f2_t h[21];  // upper triangle of sum(lambda J^T J), lanes = the two point slots
  f2_t b[6];
  f2_t chi_in, chi_out;
  int n_in;
};

// per-thread constants of a round
struct PicpConsts {
  float T[12];  // world-in-camera: 3x3 linear (column-major) then translation
};

// 1/x to within 1 ulp without the IEEE-division slow path: MUFU.RCP + one Newton step
__device__ __forceinline__ f2_t picp_rcp2(f2_t x) {
  float x0, x1, r0, r1;
  f2_unpack(x, x0, x1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(x0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(x1));
  const f2_t r = f2_pack(r0, r1);
  const f2_t e = f2_fma(x, f2_mul(r, f2_bc(-1.f)), f2_bc(1.0f));  // 1 - x*r
  return f2_fma(r, e, r);
}

__device__ __forceinline__ f2_t f2_sel(bool c0, bool c1, f2_t a, f2_t b) {
  float a0, a1, b0, b1;
  f2_unpack(a, a0, a1);
  f2_unpack(b, b0, b1);
  return f2_pack(c0 ? a0 : b0, c1 ? a1 : b1);
}

template <bool PINHOLE, bool KEEP>
__device__ __forceinline__ void picp_point2(const PicpParams& p, const PicpConsts& c, f2_t wx,
                                            f2_t wy, f2_t wz, f2_t mu, f2_t mv, bool have1,
                                            PicpAcc& a, bool have0 = true) {
#ifdef PICP_NO_MATH  // experiment: the memory pipeline alone (tools/build_variants.sh)
  a.h[0] = f2_add(a.h[0], f2_add(f2_add(wx, wy), f2_add(wz, f2_add(mu, mv))));
  return;
#endif
  const float* T = c.T;
  const f2_t neg1 = f2_bc(-1.f);
  // camera_point = world_in_camera * world_point  (camera.h:27, picp_solver.cpp:38)
  f2_t px = f2_fma(f2_bc(T[6]), wz, f2_fma(f2_bc(T[3]), wy, f2_fma(f2_bc(T[0]), wx, f2_bc(T[9]))));
  f2_t py = f2_fma(f2_bc(T[7]), wz, f2_fma(f2_bc(T[4]), wy, f2_fma(f2_bc(T[1]), wx, f2_bc(T[10]))));
  f2_t pz = f2_fma(f2_bc(T[8]), wz, f2_fma(f2_bc(T[5]), wy, f2_fma(f2_bc(T[2]), wx, f2_bc(T[11]))));
  float pz0, pz1;
  f2_unpack(pz, pz0, pz1);
  bool valid0 = have0 && !(pz0 > p.z_far || pz0 < p.z_near);  // camera.h:28
  bool valid1 = have1 && !(pz1 > p.z_far || pz1 < p.z_near);
  // phom = K * camera_point  (camera.h:30, picp_solver.cpp:43)
  f2_t hx, hy, hz;
  if (PINHOLE) {
    hx = f2_fma(f2_bc(p.K[6]), pz, f2_mul(f2_bc(p.K[0]), px));
    hy = f2_fma(f2_bc(p.K[7]), pz, f2_mul(f2_bc(p.K[4]), py));
    hz = pz;
  } else {
    hx = f2_fma(f2_bc(p.K[6]), pz, f2_fma(f2_bc(p.K[3]), py, f2_mul(f2_bc(p.K[0]), px)));
    hy = f2_fma(f2_bc(p.K[7]), pz, f2_fma(f2_bc(p.K[4]), py, f2_mul(f2_bc(p.K[1]), px)));
    hz = f2_fma(f2_bc(p.K[8]), pz, f2_fma(f2_bc(p.K[5]), py, f2_mul(f2_bc(p.K[2]), px)));
  }
  const f2_t iz_raw = picp_rcp2(hz);  // camera.h:31 / picp_solver.cpp:44
  const f2_t u_raw = f2_mul(hx, iz_raw), v_raw = f2_mul(hy, iz_raw);
  float u0, u1, v0, v1;
  f2_unpack(u_raw, u0, u1);
  f2_unpack(v_raw, v0, v1);
  // camera.h:32-35, written so that a NaN projection is rejected as well: a point ON the camera
  // plane (hz == 0, legal with z_near == 0) gives 1/0 = Inf in the reference, u = +-Inf, rejected;
  // here the Newton step of picp_rcp2 turns that Inf into NaN, which the negated form would accept.
  valid0 = valid0 && (u0 >= 0.f && u0 <= p.max_u) && (v0 >= 0.f && v0 <= p.max_v);
  valid1 = valid1 && (u1 >= 0.f && u1 <= p.max_u) && (v1 >= 0.f && v1 <= p.max_v);
  // A rejected point is masked by SELECTION, not by a zero weight: 0 * Inf = NaN would poison H and
  // b for every later round, while the reference simply skips the point (picp_solver.cpp:72).  With
  // pc = iz = u = v = e = 0 every Jacobian entry and every product below is an exact zero.
  px = f2_sel(valid0, valid1, px, 0ull);
  py = f2_sel(valid0, valid1, py, 0ull);
  pz = f2_sel(valid0, valid1, pz, 0ull);
  const f2_t iz = f2_sel(valid0, valid1, iz_raw, 0ull);
  const f2_t u = f2_sel(valid0, valid1, u_raw, 0ull), v = f2_sel(valid0, valid1, v_raw, 0ull);
  const f2_t e0 = f2_sel(valid0, valid1, f2_fma(mu, neg1, u_raw), 0ull);  // e = proj - meas (:35)
  const f2_t e1 = f2_sel(valid0, valid1, f2_fma(mv, neg1, v_raw), 0ull);
  // A = Jp*K with Jp = [iz 0 -hx*iz^2; 0 iz -hy*iz^2]  ==  iz * (K_row{0,1} - {u,v} * K_row2)
  // J = [A | A*skew(-pc)],  skew(-pc) = [0 pz -py; -pz 0 px; py -px 0]   (:39-41, utils.h:96-102)
  const f2_t npx = f2_mul(px, neg1), npy = f2_mul(py, neg1), npz = f2_mul(pz, neg1);
  f2_t J0[6], J1[6];
  if (PINHOLE) {
    J0[0] = f2_mul(iz, f2_bc(p.K[0]));
    J0[1] = 0ull;
    J0[2] = f2_mul(iz, f2_fma(u, neg1, f2_bc(p.K[6])));
    J1[0] = 0ull;
    J1[1] = f2_mul(iz, f2_bc(p.K[4]));
    J1[2] = f2_mul(iz, f2_fma(v, neg1, f2_bc(p.K[7])));
    J0[3] = f2_mul(J0[2], py);
    J0[4] = f2_fma(J0[0], pz, f2_mul(J0[2], npx));
    J0[5] = f2_mul(J0[0], npy);
    J1[3] = f2_fma(J1[2], py, f2_mul(J1[1], npz));
    J1[4] = f2_mul(J1[2], npx);
    J1[5] = f2_mul(J1[1], px);
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J0[j] = f2_mul(iz, f2_fma(u, f2_bc(-p.K[j * 3 + 2]), f2_bc(p.K[j * 3 + 0])));
      J1[j] = f2_mul(iz, f2_fma(v, f2_bc(-p.K[j * 3 + 2]), f2_bc(p.K[j * 3 + 1])));
    }
    J0[3] = f2_fma(J0[2], py, f2_mul(J0[1], npz));
    J0[4] = f2_fma(J0[0], pz, f2_mul(J0[2], npx));
    J0[5] = f2_fma(J0[1], px, f2_mul(J0[0], npy));
    J1[3] = f2_fma(J1[2], py, f2_mul(J1[1], npz));
    J1[4] = f2_fma(J1[0], pz, f2_mul(J1[2], npx));
    J1[5] = f2_fma(J1[1], px, f2_mul(J1[0], npy));
  }
  const f2_t chi = f2_fma(e1, e1, f2_mul(e0, e0));  // :75
  float chi0, chi1;
  f2_unpack(chi, chi0, chi1);
  const bool out0 = chi0 > p.thr, out1 = chi1 > p.thr;  // :78
  float l0 = out0 ? 0.f : 1.f, l1 = out1 ? 0.f : 1.f;    // dropped outliers weigh 0 (:90)
  if (KEEP) {                                            // :80
    l0 = out0 ? sqrtf(p.thr / chi0) : 1.f;
    l1 = out1 ? sqrtf(p.thr / chi1) : 1.f;
  }
  const f2_t w = f2_pack(valid0 ? l0 : 0.f, valid1 ? l1 : 0.f);
  a.chi_out = f2_add(a.chi_out, f2_sel(valid0 && out0, valid1 && out1, chi, 0ull));    // :82
  a.chi_in = f2_add(a.chi_in, f2_sel(valid0 && !out0, valid1 && !out1, chi, 0ull));   // :86
  a.n_in += ((valid0 && !out0) ? 1 : 0) + ((valid1 && !out1) ? 1 : 0);                 // :87
  // H += J^T J * lambda ; b += J^T e * lambda   (:92-93)
  f2_t L0[6], L1[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    L0[j] = (PINHOLE && j == 1) ? 0ull : f2_mul(J0[j], w);
    L1[j] = (PINHOLE && j == 0) ? 0ull : f2_mul(J1[j], w);
  }
  int k = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
#pragma unroll
    for (int cc = r; cc < 6; ++cc) {
      const bool z0 = PINHOLE && (r == 1 || cc == 1);  // J0[1] == 0
      const bool z1 = PINHOLE && (r == 0 || cc == 0);  // J1[0] == 0
      f2_t hk = a.h[k];
      if (!z1) hk = f2_fma(L1[r], J1[cc], hk);
      if (!z0) hk = f2_fma(L0[r], J0[cc], hk);
      a.h[k] = hk;
      ++k;
    }
    f2_t bk = a.b[r];
    if (!(PINHOLE && r == 0)) bk = f2_fma(L1[r], e1, bk);
    if (!(PINHOLE && r == 1)) bk = f2_fma(L0[r], e0, bk);
    a.b[r] = bk;
  }
}

// Grid-wide barrier of a cooperative launch (all CTAs co-resident), hand-written: one arrival
// counter on one L2 line that only ever grows — round r is complete when it reaches (r+1) x CTAs —
// so there is no reset, no second phase and no generation flag.  arrive() publishes the CTA's
// partial row (fence, then one relaxed atomic); wait() is one thread polling with acquire loads
// while the rest of the CTA sleeps on the block barrier.  cooperative_groups' grid.sync() cost
// 2.6 us per round here (measured in-kernel at 148 CTAs), this one ~1 us, and work can be placed
// between arrive and wait.
__device__ __forceinline__ void grid_arrive(unsigned int* counter) {
  __threadfence();
  atomicAdd(counter, 1u);
}
__device__ __forceinline__ void grid_wait(const unsigned int* counter, unsigned int target) {
  unsigned int v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
  } while (v < target);
}

// ---- per-thread asynchronous staging ring (cp.async, SASS LDGSTS) ---------------------------------
// Registers cannot hold enough loads in flight: 6.5 TB/s x ~1.5 us of loaded HBM latency is ~10 MB,
// i.e. ~70 KB per SM, while the accumulators already take 58 registers per thread.  So the
// operands travel global -> shared memory asynchronously, PICP_DEPTH batches ahead of the
// arithmetic, in a ring that is PRIVATE to each thread (a thread only ever reads back what it
// copied itself: no block-wide barrier, only cp.async.wait_group).  Two dependent streams:
//   P(b): the int2 pairs of batch b          (coalesced 8-byte copies)
//   G(b): the 5 floats gathered through them (4-byte copies: a Vector3f is only 4-byte aligned)
// In iteration b the thread waits for the group that carries G(b) and P(b+DEPTH), issues
// G(b+DEPTH) (reading P(b+DEPTH) back from shared memory) and P(b+2*DEPTH), commits them as one
// group, and then linearises batch b.  Slots are laid out so that every access of a warp is
// conflict-free, and so that the two points a thread pairs up in packed-FP32 lanes sit in one
// 8-byte word: one LDS.64 yields a ready-made packed operand.
constexpr int PICP_DEPTH = PICP_DEPTH_N;
constexpr int PICP_SLOTS = PICP_DEPTH + 1;
constexpr int PICP_PTS_WORDS = PICP_SLOTS * (PICP_UNROLL / 2) * 5;  // 8-byte words per thread
constexpr int PICP_PRS_WORDS = PICP_SLOTS * PICP_UNROLL;
constexpr size_t PICP_SMEM_BYTES = (size_t)(PICP_PTS_WORDS + PICP_PRS_WORDS) * PICP_THREADS * 8;

__device__ __forceinline__ void cp_async4(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ f2_t lds_f2(uint32_t addr) {
  f2_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int2 lds_int2(uint32_t addr) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
// byte offsets inside the ring, relative to the thread's own base (tid * 8)
__device__ __forceinline__ constexpr uint32_t picp_pts_off(int slot, int pair, int comp) {
  return (uint32_t)(((slot * (PICP_UNROLL / 2) + pair) * 5 + comp) * PICP_THREADS * 8);
}
__device__ __forceinline__ constexpr uint32_t picp_prs_off(int slot, int u) {
  return (uint32_t)((PICP_PTS_WORDS + slot * PICP_UNROLL + u) * PICP_THREADS * 8);
}

// ---- resident kernel: all rounds of a frame-sized problem in ONE launch ---------------------------
// The streaming kernel pays a fixed ~7 us per round (launch, last-block hand-off, one-thread solve
// with global round-trips), which dwarfs the arithmetic when a frame has a few thousand
// correspondences — the size of every real VO frame (vo_complete.cpp:164-166 runs 100 rounds on
// ~100 points; the synthetic sequences of config 5 on ~7e3).  Up to PICP_RES_MAX correspondences
// are therefore gathered ONCE into the shared memory of a thread-block cluster (<= 8 CTAs, each
// holding a contiguous slice already paired up for the packed-FP32 lanes), and the cluster
// iterates all the rounds without leaving the SMs:
//   linearise the slice from shared memory -> transposed warp reduction (31 shuffles for 32
//   sums) -> per-CTA partial -> pushed into EVERY CTA's shared memory through DSMEM -> one
//   cluster barrier -> every CTA adds the partials in rank order and its thread 0 solves the 6x6
//   system and updates its own copy of the pose (identical instruction stream => identical bits in
//   every CTA, so no second barrier and no broadcast).
// Global memory is touched at the start (gather) and at the end (state write-back) only.
constexpr int PICP_HIST = 2;          // rounds of history kept for the exact early-out (8 was tried: at 8e3
                                      // correspondences the pose never re-enters a cycle that short within 100
                                      // rounds, and the longer comparison cost 6 % of a frame)
constexpr int PICP_RES_THREADS = 512;
constexpr int PICP_RES_CLUSTER = 8;                       // portable maximum
constexpr int PICP_RES_SLOTS = 4096;                      // packed two-point slots per CTA
constexpr int PICP_RES_MAX = PICP_RES_CLUSTER * PICP_RES_SLOTS * 2;  // 65536 correspondences
constexpr size_t PICP_RES_SMEM = (size_t)5 * PICP_RES_SLOTS * 8;

// lane l ends up with the sum over the warp of v[l] (v has 32 entries); 31 shuffles
__device__ __forceinline__ float warp_reduce_transposed(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// oneRound's tail (picp_solver.cpp:102-110) on register/shared state: tot = the 30 sums, T = the
// 3x4 pose (column-major 3x3 then translation), updated in place.  Returns false when the round
// is skipped (too few inliers).
__device__ __noinline__ bool picp_solve_local(const PicpParams& p, const float* tot, float* T,
                                              float* H_out, float* b_out) {
  float H[36], b[6];
  {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) {
        H[c * 6 + r] = tot[k];
        H[r * 6 + c] = tot[k];
        ++k;
      }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) b[i] = tot[21 + i];
#pragma unroll
  for (int i = 0; i < 6; ++i) H[i * 6 + i] += p.damping;  // :102
#pragma unroll
  for (int i = 0; i < 36; ++i) H_out[i] = H[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) b_out[i] = b[i];
  if (__float_as_int(tot[29]) < p.min_inliers) return false;  // :103-107
  float nb[6], dx[6];
  float* A = H;
#pragma unroll
  for (int i = 0; i < 6; ++i) nb[i] = -b[i];
  ldlt_solve_recip_dev<6>(A, nb, dx);  // :109
  float sx, cx, sy, cy, sz, cz;
  sincosf(dx[3], &sx, &cx);
  sincosf(dx[4], &sy, &cy);
  sincosf(dx[5], &sz, &cz);
  const float Rx[9] = {1, 0, 0, 0, cx, sx, 0, -sx, cx};
  const float Ry[9] = {cy, 0, -sy, 0, 1, 0, sy, 0, cy};
  const float Rz[9] = {cz, sz, 0, -sz, cz, 0, 0, 0, 1};
  float Rxy[9], R[9];
  mat3_mul_dev(Rx, Ry, Rxy);
  mat3_mul_dev(Rxy, Rz, R);
  float Tn[12];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float acc = (R[i] * T[j * 3] + R[3 + i] * T[j * 3 + 1]) + R[6 + i] * T[j * 3 + 2];
      if (j == 3) acc += dx[i];
      Tn[j * 3 + i] = acc;
    }
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = Tn[i];
  return true;
}

// GRID == false: one thread-block cluster (<= 65536 correspondences), partials exchanged through
// DSMEM and a cluster barrier.  GRID == true: the whole chip — one CTA per SM, cooperative launch,
// up to SMs x 8192 correspondences (1.2e6 on a B200) resident in the SMs' shared memory; partial
// rows go through global memory (L2) and a grid-wide barrier, and every CTA sums all of them.
template <bool PINHOLE, bool KEEP, bool GRID>
__global__ void __launch_bounds__(PICP_RES_THREADS, 1)
picp_resident_kernel(const PicpParams p, const int rounds) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = GRID ? (int)gridDim.x : (int)cluster.num_blocks();
  const int rank = GRID ? (int)blockIdx.x : (int)cluster.block_rank();
  extern __shared__ __align__(16) unsigned char picp_ring[];
  f2_t* pts = reinterpret_cast<f2_t*>(picp_ring);  // [5][PICP_RES_SLOTS]
  constexpr int W = PICP_RES_THREADS / 32;
  __shared__ float s_red[W][32];
  __shared__ float s_part[2][PICP_RES_CLUSTER][32];  // [round parity][source CTA][sum]
  __shared__ float s_T[12];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = p.n_pairs_dev ? min(*p.n_pairs_dev, (int)p.n_pairs) : (int)p.n_pairs;
  const int nslots = (n + 1) >> 1;
  const int per = (nslots + csize - 1) / csize;  // <= PICP_RES_SLOTS (host guarantees)
  const int q0 = rank * per;
  const int mine = max(0, min(per, nslots - q0));

  // gather once: global slot q holds correspondences 2q (lane 0) and 2q+1 (lane 1)
  for (int ql = tid; ql < mine; ql += PICP_RES_THREADS) {
    const int q = q0 + ql;
    const bool have1 = 2 * q + 1 < n;
    const int2 pr0 = __ldg(p.pairs + 2 * q);
    const int2 pr1 = have1 ? __ldg(p.pairs + 2 * q + 1) : pr0;
    const float* w0 = p.world + 3 * (int64_t)pr0.y;  // .second -> world (:67)
    const float* w1 = p.world + 3 * (int64_t)pr1.y;
    const float2 m0 = __ldg(reinterpret_cast<const float2*>(p.image) + pr0.x);  // .first -> image
    const float2 m1 = __ldg(reinterpret_cast<const float2*>(p.image) + pr1.x);
    float x0 = __ldg(w0), y0 = __ldg(w0 + 1), z0 = __ldg(w0 + 2);
    float x1 = __ldg(w1), y1 = __ldg(w1 + 1), z1 = __ldg(w1 + 2);
    if (p.has_pre) {  // operator*(Isometry3f, PointCloudVector<3>) folded into the gather
      const float* X = p.pre;
      const float a0 = X[0] * x0 + (X[3] * y0 + X[6] * z0) + X[9];
      const float b0 = X[1] * x0 + (X[4] * y0 + X[7] * z0) + X[10];
      const float c0 = X[2] * x0 + (X[5] * y0 + X[8] * z0) + X[11];
      const float a1 = X[0] * x1 + (X[3] * y1 + X[6] * z1) + X[9];
      const float b1 = X[1] * x1 + (X[4] * y1 + X[7] * z1) + X[10];
      const float c1 = X[2] * x1 + (X[5] * y1 + X[8] * z1) + X[11];
      x0 = a0, y0 = b0, z0 = c0, x1 = a1, y1 = b1, z1 = c1;
    }
    pts[0 * PICP_RES_SLOTS + ql] = f2_pack(x0, x1);
    pts[1 * PICP_RES_SLOTS + ql] = f2_pack(y0, y1);
    pts[2 * PICP_RES_SLOTS + ql] = f2_pack(z0, z1);
    pts[3 * PICP_RES_SLOTS + ql] = f2_pack(m0.x, m1.x);
    pts[4 * PICP_RES_SLOTS + ql] = f2_pack(m0.y, m1.y);
  }
  if (tid < 12) s_T[tid] = p.st->s.T[(tid / 3) * 4 + (tid % 3)];
  // every CTA's shared memory must exist before anyone pushes into it
  if (GRID) __syncthreads();
  else cluster.sync();

  // Results of the last PICP_HIST rounds, slot [round % PICP_HIST]: pose after the round, the round's
  // linearisation (H with damping, b) and its statistics — written by thread 0.  Early-out:
  // Gauss-Newton in FP32 is a deterministic map of the pose, so once the pose after a round equals
  // (bit for bit) the pose k <= PICP_HIST rounds earlier, the iteration has entered a cycle of
  // period k and the state after all `rounds` rounds is one of the k stored ones — known without
  // running them (vo_complete runs 100 rounds on frames that settle, at FP32 resolution, in 10-30).
  __shared__ float s_Ts[PICP_HIST][12], s_H[PICP_HIST][36], s_b[PICP_HIST][6], s_keep[PICP_HIST][4];
  __shared__ int s_stop, s_final;
  if (tid < 12) s_Ts[PICP_HIST - 1][tid] = s_T[tid];  // "round -1": the initial pose
  if (tid == 0) s_stop = 0, s_final = (rounds - 1) % PICP_HIST;
  __syncthreads();
  for (int round = 0; round < rounds; ++round) {
    PicpConsts c;
#pragma unroll
    for (int i = 0; i < 12; ++i) c.T[i] = s_T[i];
    PicpAcc a;
#pragma unroll
    for (int i = 0; i < 21; ++i) a.h[i] = 0ull;
#pragma unroll
    for (int i = 0; i < 6; ++i) a.b[i] = 0ull;
    a.chi_in = a.chi_out = 0ull;
    a.n_in = 0;
    for (int ql = tid; ql < mine; ql += PICP_RES_THREADS)
      picp_point2<PINHOLE, KEEP>(p, c, pts[ql], pts[PICP_RES_SLOTS + ql], pts[2 * PICP_RES_SLOTS + ql],
                                 pts[3 * PICP_RES_SLOTS + ql], pts[4 * PICP_RES_SLOTS + ql],
                                 2 * (q0 + ql) + 1 < n, a);
    // the two point slots, then the warp (transposed: lane i gets sum i), then the CTA
    float v[32];
#pragma unroll
    for (int i = 0; i < 29; ++i) {
      float lo, hi;
      f2_unpack(i < 21 ? a.h[i] : (i < 27 ? a.b[i - 21] : (i == 27 ? a.chi_in : a.chi_out)), lo, hi);
      v[i] = lo + hi;
    }
    v[29] = (float)a.n_in;  // exact in FP32: at most 1.2e6 < 2^24 points are resident
    v[30] = v[31] = 0.f;
    s_red[warp][lane] = warp_reduce_transposed(v);
    __syncthreads();
    if (warp == 0) {
      float t = s_red[0][lane];
#pragma unroll
      for (int wv = 1; wv < W; ++wv) t += s_red[wv][lane];
      if (GRID) {
        // this CTA's partial row, in the buffer of this round's parity
        p.partials[((int64_t)(round & 1) * csize + rank) * PICP_NACC + lane] = t;
      } else {
        // push this CTA's partial into every CTA of the cluster (DSMEM)
        float* slot = &s_part[round & 1][rank][lane];
        for (int r = 0; r < csize; ++r) *cluster.map_shared_rank(slot, r) = t;
      }
    }
    if (GRID) {
      if (warp == 0) {
        __syncwarp();
        if (lane == 0) grid_arrive(p.barrier);
      }
      if (tid == 0) grid_wait(p.barrier, (unsigned int)(round + 1) * gridDim.x);
      __syncthreads();
      // warp w sums rows w, w+W, ... (all loads in flight at once), warp 0 then sums the W results
      const float* part = p.partials + (int64_t)(round & 1) * csize * PICP_NACC;
      float acc = 0.f;
      constexpr int INFLIGHT = 10;  // 148 rows / 16 warps: every row of a warp in one round-trip
      for (int r0 = warp; r0 < csize; r0 += W * INFLIGHT) {
        float x[INFLIGHT];
#pragma unroll
        for (int q = 0; q < INFLIGHT; ++q) {
          const int r = r0 + q * W;
          x[q] = r < csize ? __ldcg(part + (int64_t)r * PICP_NACC + lane) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < INFLIGHT; ++q) acc += x[q];
      }
      __syncthreads();  // s_red is about to be reused
      s_red[warp][lane] = acc;
      __syncthreads();
    } else {
      cluster.sync();  // release/acquire: all partials of this round are visible everywhere
    }
    if (warp == 0) {
      float t;
      if (GRID) {
        t = s_red[0][lane];
#pragma unroll
        for (int wv = 1; wv < W; ++wv) t += s_red[wv][lane];
      } else {
        t = s_part[round & 1][0][lane];
        for (int r = 1; r < csize; ++r) t += s_part[round & 1][r][lane];
      }
      if (lane == 29) t = __int_as_float((int)t);  // the solve reads n_in as an integer
      s_red[0][lane] = t;
      __syncwarp();
      if (lane == 0) {
        const int cur = round % PICP_HIST;
        s_keep[cur][3] = picp_solve_local(p, s_red[0], s_T, s_H[cur], s_b[cur]) ? 1.f : 0.f;
        s_keep[cur][0] = s_red[0][27];
        s_keep[cur][1] = s_red[0][28];
        s_keep[cur][2] = s_red[0][29];
        // smallest k with pose(after this round) == pose(after round - k); round - k == -1 is the
        // initial pose.  Slot `cur` still holds round - PICP_HIST, so k == PICP_HIST is checked first.
        int period = 0;
        if (p.early_out && round + 1 < rounds) {
          for (int k = PICP_HIST; k >= 1; --k) {
            if (round - k < -1) continue;
            const float* old = s_Ts[(round - k + PICP_HIST) % PICP_HIST];
            bool same = true;
#pragma unroll
            for (int i = 0; i < 12; ++i) same = same && __float_as_uint(s_T[i]) == __float_as_uint(old[i]);
            if (same) period = k;
          }
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) s_Ts[cur][i] = s_T[i];
        if (period > 0) {
          // states repeat with this period: S(round + j) == S(round + j - period)
          const int m = (rounds - 1 - round) % period;
          s_stop = 1;
          s_final = m == 0 ? cur : (round + m - period + PICP_HIST) % PICP_HIST;
        }
      }
    }
    __syncthreads();
    if (s_stop) break;  // the same decision in every CTA: they hold identical copies of the pose
  }
  // state write-back (CTA 0, thread 0): the pose and the LAST linearisation
  if (rank == 0 && tid == 0 && rounds > 0) {
    vo_picp_state& s = p.st->s;
    const int fin = s_final;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) s.T[j * 4 + i] = s_Ts[fin][j * 3 + i];
    s.T[3] = s.T[7] = s.T[11] = 0.f;
    s.T[15] = 1.f;
#pragma unroll
    for (int i = 0; i < 36; ++i) s.H[i] = s_H[fin][i];
#pragma unroll
    for (int i = 0; i < 6; ++i) s.b[i] = s_b[fin][i];
    s.chi_inliers = s_keep[fin][0];
    s.chi_outliers = s_keep[fin][1];
    s.num_inliers = __float_as_int(s_keep[fin][2]);
    s.rounds_done += rounds;
    s.last_ok = s_keep[fin][3] != 0.f ? 1 : 0;
  }
  // no CTA may exit while a peer can still push into its shared memory
  if (!GRID) cluster.sync();
}

// ---- streaming kernel: more correspondences than the cluster's shared memory holds --------------
// One cooperative launch for ALL rounds (grid = one CTA per SM).  Per round every CTA streams its
// share of the correspondences through the cp.async ring, reduces to one partial row, and after a
// single grid-wide barrier every CTA sums all rows in a fixed order and solves for its own copy of
// the pose — no kernel boundary, no last-block hand-off and no global round-trip of the pose
// between Gauss-Newton iterations.
template <bool PINHOLE, bool KEEP>
__global__ void __launch_bounds__(PICP_THREADS, 1)
picp_stream_kernel(const PicpParams p, const int rounds) {
  namespace cg = cooperative_groups;
  __shared__ float s_red[PICP_THREADS / 32][PICP_NACC];
  __shared__ float s_T[12], s_H[36], s_b[6], s_keep[4];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid < 12) s_T[tid] = p.st->s.T[(tid / 3) * 4 + (tid % 3)];
  __syncthreads();

  // ---- asynchronous stream over the correspondences ------------------------------------------
  // thread t owns items t, t+stride, t+2*stride, ...; they are consumed in batches of PICP_UNROLL.
  const int n = (int)p.n_pairs;
  const int stride = (int)gridDim.x * PICP_THREADS;
  const int i0 = (int)blockIdx.x * PICP_THREADS + tid;
  const int mine = i0 < n ? (n - i0 + stride - 1) / stride : 0;  // items of this thread
  // batches of PICP_UNROLL items; the last one is padded with repeats of the thread's first item
  // (masked out in the arithmetic) so that EVERY item goes through the asynchronous ring — a
  // separate tail loop of dependent global loads cost ~3 us of exposed latency per round
  const int nb = (mine + PICP_UNROLL - 1) / PICP_UNROLL;
  const int2* pp = p.pairs + i0;
  extern __shared__ __align__(16) unsigned char picp_ring[];
  const uint32_t ring = smem_u32(picp_ring) + (uint32_t)tid * 8u;
  auto issue_pairs = [&](int b, int slot) {  // P(b)
    if (b < nb) {
#pragma unroll
      for (int u = 0; u < PICP_UNROLL; ++u) {
        const int j = b * PICP_UNROLL + u;
        cp_async8(ring + picp_prs_off(slot, u), pp + (int64_t)(j < mine ? j : 0) * stride);
      }
    }
  };
  auto issue_gathers = [&](int b, int slot) {  // G(b); P(b) has landed
    if (b < nb) {
#pragma unroll
      for (int u = 0; u < PICP_UNROLL; ++u) {
        const int2 pr = lds_int2(ring + picp_prs_off(slot, u));
        const float* wp = p.world + 3 * (int64_t)pr.y;  // .second -> world (:67)
        const float* ip = p.image + 2 * (int64_t)pr.x;  // .first -> image (:66)
        const uint32_t half = (uint32_t)(u & 1) * 4u;   // which half of the packed pair
        cp_async4(ring + picp_pts_off(slot, u >> 1, 0) + half, wp);
        cp_async4(ring + picp_pts_off(slot, u >> 1, 1) + half, wp + 1);
        cp_async4(ring + picp_pts_off(slot, u >> 1, 2) + half, wp + 2);
        cp_async4(ring + picp_pts_off(slot, u >> 1, 3) + half, ip);
        cp_async4(ring + picp_pts_off(slot, u >> 1, 4) + half, ip + 1);
      }
    }
  };
  // The ring's fill does not depend on the pose, only the arithmetic does: the fill of round r+1
  // (three dependent memory round-trips) is issued while round r is being reduced and solved.
  auto fill_pairs = [&]() {  // P(0..DEPTH-1)
#pragma unroll
    for (int j = 0; j < PICP_DEPTH; ++j) issue_pairs(j, j % PICP_SLOTS);
    cp_async_commit();
  };
  auto fill_gathers = [&]() {  // G(j) + P(j+DEPTH) as the groups the main loop expects
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < PICP_DEPTH; ++j) {
      issue_gathers(j, j % PICP_SLOTS);
      issue_pairs(j + PICP_DEPTH, (j + PICP_DEPTH) % PICP_SLOTS);
      cp_async_commit();
    }
  };
  fill_pairs();
  fill_gathers();

#ifdef PICP_PROFILE
  long long prof[6] = {0, 0, 0, 0, 0, 0};
#define PP_T(v) const long long v = clock64()
#define PP_ADD(i, v) prof[i] += clock64() - (v)
#else
#define PP_T(v)
#define PP_ADD(i, v)
#endif
  for (int round = 0; round < rounds; ++round) {
  PP_T(t_loop);

  PicpConsts c;
#pragma unroll
  for (int i = 0; i < 12; ++i) c.T[i] = s_T[i];

  PicpAcc a;
#pragma unroll
  for (int i = 0; i < 21; ++i) a.h[i] = 0ull;
#pragma unroll
  for (int i = 0; i < 6; ++i) a.b[i] = 0ull;
  a.chi_in = a.chi_out = 0ull;
  a.n_in = 0;

  {
    for (int b0 = 0; b0 < nb; b0 += PICP_SLOTS) {
#pragma unroll
      for (int sl = 0; sl < PICP_SLOTS; ++sl) {
        const int b = b0 + sl;
        if (b < nb) {
          cp_async_wait<PICP_DEPTH - 1>();  // G(b) and P(b+DEPTH) have landed
          issue_gathers(b + PICP_DEPTH, (sl + PICP_DEPTH) % PICP_SLOTS);
          issue_pairs(b + 2 * PICP_DEPTH, (sl + 2 * PICP_DEPTH) % PICP_SLOTS);
          cp_async_commit();
#pragma unroll
          for (int pi = 0; pi < PICP_UNROLL / 2; ++pi)
            picp_point2<PINHOLE, KEEP>(p, c, lds_f2(ring + picp_pts_off(sl, pi, 0)),
                                       lds_f2(ring + picp_pts_off(sl, pi, 1)),
                                       lds_f2(ring + picp_pts_off(sl, pi, 2)),
                                       lds_f2(ring + picp_pts_off(sl, pi, 3)),
                                       lds_f2(ring + picp_pts_off(sl, pi, 4)),
                                       b * PICP_UNROLL + 2 * pi + 1 < mine, a, b * PICP_UNROLL + 2 * pi < mine);
        }
      }
    }
    cp_async_wait<0>();
    // next round's pair loads go out now; its gathers follow after the warp reduction below
    const bool more = round + 1 < rounds;
    if (more) fill_pairs();
  }
  PP_ADD(0, t_loop);
  PP_T(t_red);

  // ---- block reduction: transposed warp reduction (31 shuffles for the 30 sums), fixed-order sum
  // across warps, one partial row per CTA ---------------------------------------------------------
  {
    float v[32];
#pragma unroll
    for (int i = 0; i < 29; ++i) {  // the two point slots
      float lo, hi;
      f2_unpack(i < 21 ? a.h[i] : (i < 27 ? a.b[i - 21] : (i == 27 ? a.chi_in : a.chi_out)), lo, hi);
      v[i] = lo + hi;
    }
    v[29] = (float)a.n_in;  // exact: a thread sees far fewer than 2^24 points per round
    v[30] = v[31] = 0.f;
    float mine_sum = warp_reduce_transposed(v);  // lane i holds sum i of the warp
    if (lane == 29) mine_sum = __int_as_float((int)mine_sum);  // counts travel as integers
    s_red[warp][lane] = mine_sum;
  }
  __syncthreads();
  if (warp == 0) {
    if (lane < 30) {
      float* out = p.partials + ((int64_t)(round & 1) * gridDim.x + blockIdx.x) * PICP_NACC;
      if (lane < 29) {
        float s = s_red[0][lane];
#pragma unroll
        for (int wv = 1; wv < PICP_THREADS / 32; ++wv) s += s_red[wv][lane];
        out[lane] = s;
      } else {
        int s = 0;
#pragma unroll
        for (int wv = 0; wv < PICP_THREADS / 32; ++wv) s += __float_as_int(s_red[wv][29]);
        out[29] = __int_as_float(s);
      }
    }
    __syncwarp();
    if (lane == 0) grid_arrive(p.barrier);
  }
  PP_ADD(1, t_red);
  PP_T(t_bar);
  // the next round's gathers go out while the other CTAs arrive: they are in flight across the
  // barrier, the cross-CTA sum and the solve
  if (round + 1 < rounds) fill_gathers();

  // ---- every block: fixed-order sum over all blocks' partials, solve, own copy of the pose -------
  // One grid-wide barrier per round; the partial buffers alternate with the round's parity, so a
  // fast block's next round cannot overwrite what a slow one is still reading.  Every block runs
  // the same instruction stream on the same numbers, so all copies of the pose stay identical.
  if (tid == 0) grid_wait(p.barrier, (unsigned int)(round + 1) * gridDim.x);
  __syncthreads();
  PP_ADD(2, t_bar);
  PP_T(t_sum);
  {
    // warp w sums blocks w, w+W, ... for component `lane` (16 independent loads in flight per
    // step, combined in a fixed order); then warp 0 sums the W rows
    constexpr int W = PICP_THREADS / 32;
    constexpr int B = 16;
    float s = 0.f;
    int si = 0;
    const int nblk = (int)gridDim.x;
    const float* part = p.partials + (int64_t)(round & 1) * gridDim.x * PICP_NACC;
    for (int bk0 = warp; bk0 < nblk; bk0 += W * B) {
      float x[B];
#pragma unroll
      for (int q = 0; q < B; ++q) {
        const int bk = bk0 + q * W;
        x[q] = (bk < nblk && lane < 30) ? __ldcg(part + (int64_t)bk * PICP_NACC + lane) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < B; ++q) {
        if (lane == 29) si += __float_as_int(x[q]);
        else s += x[q];
      }
    }
    __syncthreads();
    s_red[warp][lane] = (lane == 29) ? __int_as_float(si) : s;
    __syncthreads();
    if (warp == 0) {
      float tot = 0.f;
      int toti = 0;
#pragma unroll
      for (int wv = 0; wv < W; ++wv) {
        if (lane == 29) toti += __float_as_int(s_red[wv][29]);
        else tot += s_red[wv][lane];
      }
      s_red[0][lane] = (lane == 29) ? __int_as_float(toti) : tot;
      __syncwarp();
      PP_ADD(3, t_sum);
      PP_T(t_solve);
      if (lane == 0) {
        s_keep[3] = picp_solve_local(p, s_red[0], s_T, s_H, s_b) ? 1.f : 0.f;
        s_keep[0] = s_red[0][27];
        s_keep[1] = s_red[0][28];
        s_keep[2] = s_red[0][29];
      }
      PP_ADD(4, t_solve);
    }
    __syncthreads();
  }
  }  // rounds
#ifdef PICP_PROFILE
  if (blockIdx.x == 0 && tid == 0)
    printf("picp_stream profile (cycles over %d rounds, CTA 0 thread 0): loop %lld  block-reduce %lld  grid.sync %lld  partial-sum %lld  solve %lld\n",
           rounds, prof[0], prof[1], prof[2], prof[3], prof[4]);
#endif
  // state write-back (block 0, thread 0): the pose and the LAST linearisation
  if (blockIdx.x == 0 && tid == 0 && rounds > 0) {
    vo_picp_state& st = p.st->s;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) st.T[j * 4 + i] = s_T[j * 3 + i];
    st.T[3] = st.T[7] = st.T[11] = 0.f;
    st.T[15] = 1.f;
#pragma unroll
    for (int i = 0; i < 36; ++i) st.H[i] = s_H[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) st.b[i] = s_b[i];
    st.chi_inliers = s_keep[0];
    st.chi_outliers = s_keep[1];
    st.num_inliers = __float_as_int(s_keep[2]);
    st.rounds_done += rounds;
    st.last_ok = s_keep[3] != 0.f ? 1 : 0;
  }
}

// ---- streaming kernel, second generation: TMA windows instead of per-thread gathers ----------------
// What limits picp_stream_kernel is not HBM: with its arithmetic removed it still needs 44 us per
// round for 279 MB (6.3 TB/s), and the arithmetic adds 8 us that do not overlap — 6 cp.async and
// ~15 address/LDS instructions per point compete with 80 arithmetic ones for the issue slots of 12
// warps.  Correspondence lists that reach this size are monotone or nearly so (the synthetic sweeps
// of picp_solver_test.cpp:8-26 pair point i with point i): a tile of 1536 consecutive pairs then
// references a WINDOW of the world / image arrays barely larger than the tile.  So a producer warp
//   1. GUESSES the windows from the tile's first and last pair (PW_MAX points from the smaller index
//      on; the two pairs are read with plain loads one tile earlier),
//   2. requests the tile's pairs (12 KB) and both windows with three bulk copies (cp.async.bulk) on one
//      mbarrier, three tiles ahead of the arithmetic,
// and the 12 arithmetic warps read pairs and points from shared memory, spending their issue slots
// on arithmetic.  The guess needs no proof: every thread checks that its pair's indices fall inside
// the staged windows and gathers the point from global memory itself when they do not (a shuffled
// list, the ragged end of the list, a window cut short by the end of an array) — correct for any
// correspondence list, just slower.  The producer runs ahead across round boundaries: the next
// round's first tiles land while the grid reduces and solves.
// RESULT (B200, 9.97e6 correspondences, 10 rounds): 61.3 us per round against 57.4 for
// picp_stream_kernel — the instruction count per point drops as intended, but what the arithmetic
// warps lack is not issue slots: at 128 registers only 3 of them fit per scheduler and the ~25-deep
// dependent chain of projection -> reciprocal -> Jacobian leaves the issue port idle half the time
// either way, and a tile-wide barrier marches all twelve warps in step where the per-thread rings
// let them drift apart.  Kept opt-in (VO_PICP_STREAM_V2=1) and tested, not the default.
constexpr int PW_THREADS = 384;                 // arithmetic threads
constexpr int PW_TILE = PW_THREADS * 4;         // pairs per tile
constexpr int PW_MAX = 2048;                    // points per window
constexpr int PW_STAGES = 4;
constexpr uint32_t PW_PAIR_BYTES = PW_TILE * 8;
constexpr uint32_t PW_WORLD_BYTES = PW_MAX * 12 + 32;
constexpr uint32_t PW_IMAGE_BYTES = PW_MAX * 8 + 32;
constexpr uint32_t PW_STAGE_BYTES = PW_PAIR_BYTES + PW_WORLD_BYTES + PW_IMAGE_BYTES;
constexpr size_t PW_SMEM_BYTES = (size_t)PW_STAGES * PW_STAGE_BYTES;

struct PwTile {      // what the producer tells the arithmetic warps about a staged tile
  int pairs_staged;  // 0: ragged last tile, pairs are read from global memory too
  int world_first;   // index of the first world / image point held by the windows ...
  int image_first;
  int world_count;   // ... and how many complete points each window holds (0: nothing staged)
  int image_count;
  int world_skip;    // bytes between the window's 16-byte aligned start and that first element
  int image_skip;
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PW_THREADS) : "memory"); }

template <bool PINHOLE, bool KEEP>
__global__ void __launch_bounds__(PW_THREADS + 32, 1)
picp_window_kernel(const PicpParams p, const int rounds, const int64_t n_world, const int64_t n_image) {
  extern __shared__ __align__(128) unsigned char pw_smem[];
  __shared__ uint64_t full[PW_STAGES], empty[PW_STAGES];
  __shared__ PwTile s_tile[PW_STAGES];
  __shared__ float s_red[PW_THREADS / 32][PICP_NACC];
  __shared__ float s_T[12], s_H[36], s_b[6], s_keep[4];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t n = p.n_pairs;
  const int64_t n_tiles = (n + PW_TILE - 1) / PW_TILE;
  // tiles blockIdx.x, blockIdx.x + grid, ... of every round, as one sequence over all rounds
  const int64_t mine = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total = mine * rounds;
  if (tid == 0) {
    for (int s = 0; s < PW_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], PW_THREADS / 32);
    }
    mbar_fence_init();
  }
  if (tid < 12) s_T[tid] = p.st->s.T[(tid / 3) * 4 + (tid % 3)];
  __syncthreads();

  if (warp == PW_THREADS / 32) {
    // ===== producer (one thread) =========================================================================
    // A tile's pairs and both windows are requested together, PW_STAGES - 1 tiles ahead of the
    // arithmetic: the windows are guessed from the tile's first and last pair, which are read with
    // two plain loads one iteration before they are needed.
    if (lane == 0) {
      auto tile_of = [&](int64_t k) { return (int64_t)blockIdx.x + (k % mine) * gridDim.x; };
      int2 nf = make_int2(0, 0), nl = make_int2(0, 0);
      auto peek = [&](int64_t k) {
        const int64_t t = tile_of(k);
        if ((t + 1) * PW_TILE <= n) {
          nf = __ldg(p.pairs + t * PW_TILE);
          nl = __ldg(p.pairs + t * PW_TILE + PW_TILE - 1);
        }
      };
      if (total > 0) peek(0);
      for (int64_t k = 0; k < total; ++k) {
        const int s = (int)(k % PW_STAGES);
        const int2 first = nf, last = nl;
        if (k + 1 < total) peek(k + 1);
        mbar_wait(&empty[s], (uint32_t)(((k / PW_STAGES) & 1) ^ 1));
        const int64_t t = tile_of(k);
        const bool whole = (t + 1) * PW_TILE <= n;
        PwTile d;
        d.pairs_staged = whole ? 1 : 0;
        d.world_first = d.image_first = d.world_count = d.image_count = d.world_skip = d.image_skip = 0;
        uint32_t bytes = whole ? PW_PAIR_BYTES : 0u;
        int64_t w0 = 0, w1 = 0, i0 = 0, i1 = 0;
        bool windows = false;
        if (whole) {
          const int64_t lo_w = max(0, min(first.y, last.y)), lo_i = max(0, min(first.x, last.x));
          // byte ranges of the windows: from the 16-byte boundary below the first element, PW_MAX
          // points long, cut at the last 16-byte boundary inside the array
          w0 = (lo_w * 12) & ~15LL;
          w1 = min((long long)(w0 + (int64_t)PW_MAX * 12), (long long)((n_world * 12) & ~15LL));
          i0 = (lo_i * 8) & ~15LL;
          i1 = min((long long)(i0 + (int64_t)PW_MAX * 8), (long long)((n_image * 8) & ~15LL));
          if (w1 > w0 && i1 > i0) {
            windows = true;
            d.world_first = (int)lo_w, d.image_first = (int)lo_i;
            d.world_count = (int)max(0LL, (long long)(w1 / 12 - lo_w));
            d.image_count = (int)max(0LL, (long long)(i1 / 8 - lo_i));
            d.world_skip = (int)(lo_w * 12 - w0), d.image_skip = (int)(lo_i * 8 - i0);
            bytes += (uint32_t)((w1 - w0) + (i1 - i0));
          }
        }
        s_tile[s] = d;
        unsigned char* base = pw_smem + (size_t)s * PW_STAGE_BYTES;
        if (bytes) mbar_arrive_expect_tx(&full[s], bytes);
        else mbar_arrive(&full[s]);
        if (whole) tma_load_1d(base, p.pairs + t * PW_TILE, PW_PAIR_BYTES, &full[s]);
        if (windows) {
          tma_load_1d(base + PW_PAIR_BYTES, reinterpret_cast<const unsigned char*>(p.world) + w0, (uint32_t)(w1 - w0),
                      &full[s]);
          tma_load_1d(base + PW_PAIR_BYTES + PW_WORLD_BYTES, reinterpret_cast<const unsigned char*>(p.image) + i0,
                      (uint32_t)(i1 - i0), &full[s]);
        }
      }
    }
    return;
  }

  // ===== arithmetic warps ==================================================================================
  int64_t k = 0;  // position in this CTA's tile sequence
  for (int round = 0; round < rounds; ++round) {
    PicpConsts c;
#pragma unroll
    for (int i = 0; i < 12; ++i) c.T[i] = s_T[i];
    PicpAcc a;
#pragma unroll
    for (int i = 0; i < 21; ++i) a.h[i] = 0ull;
#pragma unroll
    for (int i = 0; i < 6; ++i) a.b[i] = 0ull;
    a.chi_in = a.chi_out = 0ull;
    a.n_in = 0;

    for (int64_t m = 0; m < mine; ++m, ++k) {
      const int s = (int)(k % PW_STAGES);
      const uint32_t ph = (uint32_t)((k / PW_STAGES) & 1);
      mbar_wait(&full[s], ph);
      const unsigned char* base = pw_smem + (size_t)s * PW_STAGE_BYTES;
      const PwTile d = s_tile[s];
      const int64_t t0 = ((int64_t)blockIdx.x + m * gridDim.x) * PW_TILE;
      // thread t takes pairs t, t + 384, t + 768, t + 1152 of the tile: the first two and the last
      // two share the lanes of a packed pair
      float wx[4], wy[4], wz[4], mu[4], mv[4];
      bool have[4];
      {
        const int2* pr = reinterpret_cast<const int2*>(base);
        const unsigned char* wwin = base + PW_PAIR_BYTES + d.world_skip;
        const unsigned char* iwin = base + PW_PAIR_BYTES + PW_WORLD_BYTES + d.image_skip;
        int2 v[4];
        unsigned dw[4], di[4];
        bool inside = d.pairs_staged != 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t i = t0 + u * PW_THREADS + tid;
          have[u] = i < n;
          v[u] = d.pairs_staged ? pr[u * PW_THREADS + tid] : __ldg(p.pairs + (have[u] ? i : 0));
          dw[u] = (unsigned)(v[u].y - d.world_first), di[u] = (unsigned)(v[u].x - d.image_first);
          inside = inside && dw[u] < (unsigned)d.world_count && di[u] < (unsigned)d.image_count;
        }
        if (__all_sync(0xffffffffu, inside)) {
          // the common case, branch-free: every point of the warp is in the staged windows
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float* w = reinterpret_cast<const float*>(wwin + (size_t)dw[u] * 12);
            const float2 im = *reinterpret_cast<const float2*>(iwin + (size_t)di[u] * 8);
            wx[u] = w[0], wy[u] = w[1], wz[u] = w[2], mu[u] = im.x, mv[u] = im.y;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (dw[u] < (unsigned)d.world_count) {
              const float* w = reinterpret_cast<const float*>(wwin + (size_t)dw[u] * 12);
              wx[u] = w[0], wy[u] = w[1], wz[u] = w[2];
            } else {
              const float* w = p.world + 3 * (int64_t)v[u].y;
              wx[u] = __ldg(w), wy[u] = __ldg(w + 1), wz[u] = __ldg(w + 2);
            }
            if (di[u] < (unsigned)d.image_count) {
              const float2 im = *reinterpret_cast<const float2*>(iwin + (size_t)di[u] * 8);
              mu[u] = im.x, mv[u] = im.y;
            } else {
              const float2 im = __ldg(reinterpret_cast<const float2*>(p.image) + v[u].x);
              mu[u] = im.x, mv[u] = im.y;
            }
          }
        }
      }
      // this warp is done with the stage as soon as its operands are in registers
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
#pragma unroll
      for (int pi = 0; pi < 2; ++pi)
        picp_point2<PINHOLE, KEEP>(p, c, f2_pack(wx[2 * pi], wx[2 * pi + 1]), f2_pack(wy[2 * pi], wy[2 * pi + 1]),
                                   f2_pack(wz[2 * pi], wz[2 * pi + 1]), f2_pack(mu[2 * pi], mu[2 * pi + 1]),
                                   f2_pack(mv[2 * pi], mv[2 * pi + 1]), have[2 * pi + 1], a, have[2 * pi]);
    }

    // ---- block reduction, grid barrier, cross-CTA sum, solve: as in picp_stream_kernel ---------------
    {
      float v[32];
#pragma unroll
      for (int i = 0; i < 29; ++i) {
        float lo, hi;
        f2_unpack(i < 21 ? a.h[i] : (i < 27 ? a.b[i - 21] : (i == 27 ? a.chi_in : a.chi_out)), lo, hi);
        v[i] = lo + hi;
      }
      v[29] = (float)a.n_in;
      v[30] = v[31] = 0.f;
      float mine_sum = warp_reduce_transposed(v);
      if (lane == 29) mine_sum = __int_as_float((int)mine_sum);
      s_red[warp][lane] = mine_sum;
    }
    consumer_sync();
    if (warp == 0) {
      if (lane < 30) {
        float* out = p.partials + ((int64_t)(round & 1) * gridDim.x + blockIdx.x) * PICP_NACC;
        if (lane < 29) {
          float sum = s_red[0][lane];
#pragma unroll
          for (int wv = 1; wv < PW_THREADS / 32; ++wv) sum += s_red[wv][lane];
          out[lane] = sum;
        } else {
          int sum = 0;
#pragma unroll
          for (int wv = 0; wv < PW_THREADS / 32; ++wv) sum += __float_as_int(s_red[wv][29]);
          out[29] = __int_as_float(sum);
        }
      }
      __syncwarp();
      if (lane == 0) grid_arrive(p.barrier);
    }
    if (tid == 0) grid_wait(p.barrier, (unsigned int)(round + 1) * gridDim.x);
    consumer_sync();
    {
      constexpr int W = PW_THREADS / 32;
      constexpr int B = 16;
      float sum = 0.f;
      int si = 0;
      const int nblk = (int)gridDim.x;
      const float* part = p.partials + (int64_t)(round & 1) * gridDim.x * PICP_NACC;
      for (int bk0 = warp; bk0 < nblk; bk0 += W * B) {
        float x[B];
#pragma unroll
        for (int q = 0; q < B; ++q) {
          const int bk = bk0 + q * W;
          x[q] = (bk < nblk && lane < 30) ? __ldcg(part + (int64_t)bk * PICP_NACC + lane) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < B; ++q) {
          if (lane == 29) si += __float_as_int(x[q]);
          else sum += x[q];
        }
      }
      consumer_sync();
      s_red[warp][lane] = (lane == 29) ? __int_as_float(si) : sum;
      consumer_sync();
      if (warp == 0) {
        float tot = 0.f;
        int toti = 0;
#pragma unroll
        for (int wv = 0; wv < W; ++wv) {
          if (lane == 29) toti += __float_as_int(s_red[wv][29]);
          else tot += s_red[wv][lane];
        }
        s_red[0][lane] = (lane == 29) ? __int_as_float(toti) : tot;
        __syncwarp();
        if (lane == 0) {
          s_keep[3] = picp_solve_local(p, s_red[0], s_T, s_H, s_b) ? 1.f : 0.f;
          s_keep[0] = s_red[0][27];
          s_keep[1] = s_red[0][28];
          s_keep[2] = s_red[0][29];
        }
      }
      consumer_sync();
    }
  }
  if (blockIdx.x == 0 && tid == 0 && rounds > 0) {
    vo_picp_state& st = p.st->s;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) st.T[j * 4 + i] = s_T[j * 3 + i];
    st.T[3] = st.T[7] = st.T[11] = 0.f;
    st.T[15] = 1.f;
#pragma unroll
    for (int i = 0; i < 36; ++i) st.H[i] = s_H[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) st.b[i] = s_b[i];
    st.chi_inliers = s_keep[0];
    st.chi_outliers = s_keep[1];
    st.num_inliers = __float_as_int(s_keep[2]);
    st.rounds_done += rounds;
    st.last_ok = s_keep[3] != 0.f ? 1 : 0;
  }
}

// lowest position of a pair whose indices fall outside the two point sets (all ones = none)
__global__ void __launch_bounds__(256)
picp_check_pairs_kernel(const int2* __restrict__ pairs, int64_t n, int n_image, int n_world,
                        unsigned long long* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int2 p = __ldg(pairs + i);
    if ((unsigned)p.x >= (unsigned)n_image || (unsigned)p.y >= (unsigned)n_world)
      atomicMin(bad, (unsigned long long)i);
  }
}

}  // namespace vo

// =================================================================================================
// host side
// =================================================================================================
using namespace vo;

struct vo_picp_s {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  vo_camera cam{};
  bool have_cam = false;
  float thr = 1000.f, damping = 1.f;  // picp_solver.cpp:10-13
  bool smem_opted_in = false;
  bool force_stream = false;          // VO_PICP_FORCE_STREAM=1: never use the resident kernel
  bool force_general = false;         // VO_PICP_FORCE_GENERAL=1: never use the pinhole kernel
  bool early_out = true;              // VO_PICP_NO_EARLY_OUT=1: always run every requested round
  bool stream_v2 = false;             // VO_PICP_STREAM_V2=1: TMA-window streaming kernel instead of per-thread cp.async
  bool window_opted_in = false;
  int32_t min_inliers = 0;
  DevBuf world_buf, image_buf, pairs_buf, state_buf, partials_buf, check_buf, barrier_buf;
  const float* world = nullptr;
  const float* image = nullptr;
  const int32_t* pairs = nullptr;
  int64_t n_world = 0, n_image = 0, n_pairs = 0;
  int grid = 0;
};

static int picp_upload_state(vo_picp_s* h) {
  int rc = h->state_buf.reserve(sizeof(PicpDeviceState));
  if (rc) return rc;
  PicpDeviceState st;
  memset(&st, 0, sizeof(st));
  memcpy(st.s.T, h->cam.T, sizeof(st.s.T));
  st.s.last_ok = 1;
  // pageable -> device copies are staged by the runtime before the call returns
  VO_CUDA(cudaMemcpyAsync(h->state_buf.p, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
  return VO_OK;
}

static int picp_fill_params(vo_picp_s* h, int keep_outliers, PicpParams* p) {
  p->world = h->world;
  p->image = h->image;
  p->pairs = reinterpret_cast<const int2*>(h->pairs);
  p->n_pairs = h->n_pairs;
  memcpy(p->K, h->cam.K, sizeof(p->K));
  p->z_near = (float)h->cam.z_near;
  p->z_far = (float)h->cam.z_far;
  p->max_u = (float)(h->cam.cols - 1);
  p->max_v = (float)(h->cam.rows - 1);
  p->thr = h->thr;
  p->damping = h->damping;
  p->min_inliers = h->min_inliers;
  p->keep_outliers = keep_outliers ? 1 : 0;
  p->st = h->state_buf.as<PicpDeviceState>();
  p->partials = h->partials_buf.as<float>();
  p->barrier = h->barrier_buf.as<unsigned int>();
  p->early_out = h->early_out ? 1 : 0;
  p->n_pairs_dev = nullptr;
  p->has_pre = 0;
  memset(p->pre, 0, sizeof(p->pre));
  return VO_OK;
}

static int picp_pick_grid(vo_picp_s* h) {
  const int sms = num_sms(h->device);
  int per_sm = 2;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, picp_stream_kernel<false, true>, PICP_THREADS,
                                                PICP_SMEM_BYTES);
  if (per_sm < 1) per_sm = 1;
  const int64_t full = (int64_t)sms * per_sm;
  const int64_t need =
      (h->n_pairs + (int64_t)PICP_THREADS * PICP_UNROLL - 1) / ((int64_t)PICP_THREADS * PICP_UNROLL);
  int64_t g = need < full ? need : full;
  if (g < 1) g = 1;
  return (int)g;
}

namespace vo {
const void* picp_state_device_ptr(vo_picp_t h) { return h ? h->state_buf.p : nullptr; }
}  // namespace vo

extern "C" {

int vo_picp_create(vo_picp_t* out, int device) {
  VO_REQUIRE(out != nullptr, VO_ERR_ARG, "null handle pointer");
  int n = 0;
  VO_CUDA(cudaGetDeviceCount(&n));
  VO_REQUIRE(device >= 0 && device < n, VO_ERR_ARG, "bad device ordinal");
  DeviceGuard g(device);
  vo_picp_s* h = new vo_picp_s();
  h->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate -> %s", cudaGetErrorString(e));
    delete h;
    return VO_ERR_CUDA;
  }
  h->own_stream = true;
  const char* fg = getenv("VO_PICP_FORCE_GENERAL");
  h->force_general = fg != nullptr && fg[0] == '1';
  const char* fs = getenv("VO_PICP_FORCE_STREAM");
  h->force_stream = fs != nullptr && fs[0] == '1';
  const char* ne = getenv("VO_PICP_NO_EARLY_OUT");
  h->early_out = !(ne != nullptr && ne[0] == '1');
  const char* v2 = getenv("VO_PICP_STREAM_V2");
  h->stream_v2 = v2 != nullptr && v2[0] == '1';
  *out = h;
  return VO_OK;
}

int vo_picp_destroy(vo_picp_t h) {
  if (!h) return VO_OK;
  DeviceGuard g(h->device);
  cudaStreamSynchronize(h->stream);
  h->world_buf.release();
  h->image_buf.release();
  h->pairs_buf.release();
  h->state_buf.release();
  h->partials_buf.release();
  h->check_buf.release();
  h->barrier_buf.release();
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return VO_OK;
}

int vo_picp_set_stream(vo_picp_t h, void* cuda_stream) {
  VO_REQUIRE(h != nullptr, VO_ERR_ARG, "null handle");
  DeviceGuard g(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->own_stream) {
    cudaStreamDestroy(h->stream);
    h->own_stream = false;
  }
  h->stream = static_cast<cudaStream_t>(cuda_stream);
  return VO_OK;
}

int vo_picp_synchronize(vo_picp_t h) {
  VO_REQUIRE(h != nullptr, VO_ERR_ARG, "null handle");
  DeviceGuard g(h->device);
  VO_CUDA(cudaStreamSynchronize(h->stream));
  return VO_OK;
}

int vo_picp_set_params(vo_picp_t h, float kernel_threshold, float damping,
                       int32_t min_num_inliers) {
  VO_REQUIRE(h != nullptr, VO_ERR_ARG, "null handle");
  h->thr = kernel_threshold;
  h->damping = damping;
  h->min_inliers = min_num_inliers;
  return VO_OK;
}

static int picp_init_common(vo_picp_s* h, const vo_camera* cam, int64_t n_world, int64_t n_image) {
  h->cam = *cam;
  h->have_cam = true;
  h->n_world = n_world;
  h->n_image = n_image;
  return picp_upload_state(h);
}

int vo_picp_init(vo_picp_t h, const vo_camera* cam, const float* world_host, int64_t n_world,
                 const float* image_host, int64_t n_image) {
  VO_REQUIRE(h != nullptr && cam != nullptr, VO_ERR_ARG, "null handle/camera");
  VO_REQUIRE(n_world >= 0 && n_image >= 0, VO_ERR_ARG, "negative size");
  VO_REQUIRE((world_host || n_world == 0) && (image_host || n_image == 0), VO_ERR_ARG,
             "null points");
  DeviceGuard g(h->device);
  int rc = h->world_buf.reserve((size_t)n_world * 12 + 16);
  if (rc) return rc;
  rc = h->image_buf.reserve((size_t)n_image * 8 + 16);
  if (rc) return rc;
  // through the pinned staging ring; the caller's vectors are free again when this returns (the
  // reference borrows them, picp_solver.cpp:21-22 — we own a copy)
  if ((rc = stage_h2d(h->device, h->world_buf.p, world_host, (size_t)n_world * 12, h->stream))) return rc;
  if ((rc = stage_h2d(h->device, h->image_buf.p, image_host, (size_t)n_image * 8, h->stream))) return rc;
  h->world = h->world_buf.as<float>();
  h->image = h->image_buf.as<float>();
  return picp_init_common(h, cam, n_world, n_image);
}

int vo_picp_init_device(vo_picp_t h, const vo_camera* cam, const float* world_dev, int64_t n_world,
                        const float* image_dev, int64_t n_image) {
  VO_REQUIRE(h != nullptr && cam != nullptr, VO_ERR_ARG, "null handle/camera");
  VO_REQUIRE(n_world >= 0 && n_image >= 0, VO_ERR_ARG, "negative size");
  VO_REQUIRE((world_dev || n_world == 0) && (image_dev || n_image == 0), VO_ERR_ARG, "null points");
  VO_REQUIRE((reinterpret_cast<uintptr_t>(image_dev) & 7u) == 0, VO_ERR_ARG,
             "image points must be 8-byte aligned");
  DeviceGuard g(h->device);
  h->world = world_dev;
  h->image = image_dev;
  return picp_init_common(h, cam, n_world, n_image);
}

int vo_picp_set_correspondences(vo_picp_t h, const int32_t* pairs_host, int64_t n_pairs) {
  VO_REQUIRE(h != nullptr, VO_ERR_ARG, "null handle");
  VO_REQUIRE(n_pairs >= 0 && n_pairs < (1LL << 31) && (pairs_host || n_pairs == 0), VO_ERR_ARG,
             "bad pairs");
  DeviceGuard g(h->device);
  int rc = h->pairs_buf.reserve((size_t)n_pairs * 8 + 16);
  if (rc) return rc;
  h->n_pairs = 0;
  // The reference indexes with operator[] (UB when out of range, picp_solver.cpp:66-71); we refuse
  // instead of reading out of bounds on the device.  Frame-sized lists are checked on the host
  // before the upload; large ones by a kernel over the uploaded copy (an O(n) host loop cost as
  // much as the whole transfer).
  constexpr int64_t HOST_CHECK_MAX = 1 << 16;
  if (n_pairs <= HOST_CHECK_MAX) {
    for (int64_t i = 0; i < n_pairs; ++i) {
      const int32_t a = pairs_host[2 * i], b = pairs_host[2 * i + 1];
      if (a < 0 || a >= h->n_image || b < 0 || b >= h->n_world) {
        set_error("vo_picp_set_correspondences: pair %lld = (%d,%d) out of range", (long long)i, a, b);
        return VO_ERR_ARG;
      }
    }
  }
  if ((rc = stage_h2d(h->device, h->pairs_buf.p, pairs_host, (size_t)n_pairs * 8, h->stream))) return rc;
  if (n_pairs > HOST_CHECK_MAX) {
    rc = h->check_buf.reserve(16);
    if (rc) return rc;
    VO_CUDA(cudaMemsetAsync(h->check_buf.p, 0xFF, 8, h->stream));
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((n_pairs + threads - 1) / threads, 8LL * num_sms(h->device));
    picp_check_pairs_kernel<<<blocks, threads, 0, h->stream>>>(
        h->pairs_buf.as<int2>(), n_pairs, (int)std::min<int64_t>(h->n_image, INT32_MAX),
        (int)std::min<int64_t>(h->n_world, INT32_MAX), h->check_buf.as<unsigned long long>());
    VO_LAUNCH_CHECK();
    long long bad = -1;
    VO_CUDA(cudaMemcpyAsync(&bad, h->check_buf.p, sizeof(bad), cudaMemcpyDeviceToHost, h->stream));
    VO_CUDA(cudaStreamSynchronize(h->stream));
    if (bad != -1) {
      set_error("vo_picp_set_correspondences: pair %lld = (%d,%d) out of range", bad, pairs_host[2 * bad],
                pairs_host[2 * bad + 1]);
      return VO_ERR_ARG;
    }
  }
  h->pairs = h->pairs_buf.as<int32_t>();
  h->n_pairs = n_pairs;
  return VO_OK;
}

int vo_picp_set_correspondences_device(vo_picp_t h, const int32_t* pairs_dev, int64_t n_pairs) {
  VO_REQUIRE(h != nullptr, VO_ERR_ARG, "null handle");
  VO_REQUIRE(n_pairs >= 0 && n_pairs < (1LL << 31) && (pairs_dev || n_pairs == 0), VO_ERR_ARG,
             "bad pairs");
  VO_REQUIRE((reinterpret_cast<uintptr_t>(pairs_dev) & 7u) == 0, VO_ERR_ARG,
             "pairs must be 8-byte aligned");
  h->pairs = pairs_dev;
  h->n_pairs = n_pairs;
  return VO_OK;
}

static int picp_compute_common(vo_picp_t h, int keep_outliers, int rounds, const int32_t* n_pairs_dev,
                               const float* pre_transform) {
  VO_REQUIRE(h != nullptr, VO_ERR_ARG, "null handle");
  VO_REQUIRE(h->have_cam, VO_ERR_STATE, "init not called");
  VO_REQUIRE(rounds >= 0, VO_ERR_ARG, "negative rounds");
  const bool extended = n_pairs_dev != nullptr || pre_transform != nullptr;
  VO_REQUIRE(!extended || (h->n_pairs <= PICP_RES_MAX && !h->force_stream), VO_ERR_UNSUPPORTED,
             "device-side count / pre-transform need the resident kernel (<= 65536 correspondences)");
  if (rounds == 0) return VO_OK;
  DeviceGuard g(h->device);
  if (!h->smem_opted_in) {  // the staging ring needs the opt-in shared-memory carve-out (per device)
    VO_CUDA(cudaFuncSetAttribute(picp_stream_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_SMEM_BYTES));
    VO_CUDA(cudaFuncSetAttribute(picp_stream_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_SMEM_BYTES));
    VO_CUDA(cudaFuncSetAttribute(picp_stream_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_SMEM_BYTES));
    VO_CUDA(cudaFuncSetAttribute(picp_stream_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_SMEM_BYTES));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    VO_CUDA(cudaFuncSetAttribute(picp_resident_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PICP_RES_SMEM));
    h->smem_opted_in = true;
  }
  const int grid = picp_pick_grid(h);
  int rc = h->partials_buf.reserve((size_t)2 * grid * PICP_NACC * sizeof(float));
  if (rc) return rc;
  if ((rc = h->barrier_buf.reserve(128))) return rc;
  PicpParams p;
  picp_fill_params(h, keep_outliers, &p);
  p.n_pairs_dev = n_pairs_dev;
  if (pre_transform) {
    p.has_pre = 1;
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 3; ++i) p.pre[j * 3 + i] = pre_transform[j * 4 + i];
  }

  // K == [fx 0 cx; 0 fy cy; 0 0 1] exactly -> the structurally-sparse instantiation
  const float* K = h->cam.K;
  const bool pinhole = !h->force_general && K[1] == 0.f && K[2] == 0.f && K[3] == 0.f &&
                       K[5] == 0.f && K[8] == 1.f;
  auto kernel = pinhole ? (p.keep_outliers ? picp_stream_kernel<true, true> : picp_stream_kernel<true, false>)
                        : (p.keep_outliers ? picp_stream_kernel<false, true> : picp_stream_kernel<false, false>);
  if (h->n_pairs <= PICP_RES_MAX && !h->force_stream) {
    // a frame-sized problem: every round inside one resident thread-block cluster
    auto rk = pinhole ? (p.keep_outliers ? picp_resident_kernel<true, true, false> : picp_resident_kernel<true, false, false>)
                      : (p.keep_outliers ? picp_resident_kernel<false, true, false> : picp_resident_kernel<false, false, false>);
    // one packed slot per thread where possible; cluster sizes 1, 2, 4, 8
    const int64_t nslots = (h->n_pairs + 1) / 2;
    int csize = 1;
    while (csize < PICP_RES_CLUSTER && nslots > (int64_t)csize * PICP_RES_THREADS) csize *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)csize);
    cfg.blockDim = dim3(PICP_RES_THREADS);
    cfg.dynamicSmemBytes = PICP_RES_SMEM;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VO_CUDA(cudaLaunchKernelEx(&cfg, rk, p, rounds));
    VO_LAUNCH_CHECK();
    return VO_OK;
  }
  // mid-sized problems: resident in the shared memory of the whole chip (one CTA per SM)
  {
    const int sms = num_sms(h->device);
    const int64_t nslots = (h->n_pairs + 1) / 2;
    if (!h->force_stream && nslots <= (int64_t)sms * PICP_RES_SLOTS) {
      auto gk = pinhole ? (p.keep_outliers ? picp_resident_kernel<true, true, true> : picp_resident_kernel<true, false, true>)
                        : (p.keep_outliers ? picp_resident_kernel<false, true, true> : picp_resident_kernel<false, false, true>);
      rc = h->partials_buf.reserve((size_t)2 * sms * PICP_NACC * sizeof(float));
      if (rc) return rc;
      p.partials = h->partials_buf.as<float>();
      int rounds_arg = rounds;
      void* args[] = {(void*)&p, (void*)&rounds_arg};
      VO_CUDA(cudaMemsetAsync(h->barrier_buf.p, 0, 4, h->stream));
      VO_CUDA(cudaLaunchCooperativeKernel((const void*)gk, dim3((unsigned)sms), dim3(PICP_RES_THREADS), args,
                                          PICP_RES_SMEM, h->stream));
      VO_LAUNCH_CHECK();
      return VO_OK;
    }
  }
  // streaming, second generation: TMA windows — opt-in (VO_PICP_STREAM_V2=1; bulk copies need 16-byte
  // aligned arrays).  Measured SLOWER than the per-thread cp.async kernel below (61.3 against 57.4 us
  // per round at 9.97e6), see the comment on picp_window_kernel.
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(h->world) | reinterpret_cast<uintptr_t>(h->image) |
                           reinterpret_cast<uintptr_t>(h->pairs)) & 15u) == 0;
  if (aligned16 && h->stream_v2) {
    auto wk = pinhole ? (p.keep_outliers ? picp_window_kernel<true, true> : picp_window_kernel<true, false>)
                      : (p.keep_outliers ? picp_window_kernel<false, true> : picp_window_kernel<false, false>);
    if (!h->window_opted_in) {
      VO_CUDA(cudaFuncSetAttribute(picp_window_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PW_SMEM_BYTES));
      VO_CUDA(cudaFuncSetAttribute(picp_window_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PW_SMEM_BYTES));
      VO_CUDA(cudaFuncSetAttribute(picp_window_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PW_SMEM_BYTES));
      VO_CUDA(cudaFuncSetAttribute(picp_window_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PW_SMEM_BYTES));
      h->window_opted_in = true;
    }
    const int sms = num_sms(h->device);
    const int64_t n_tiles = (h->n_pairs + PW_TILE - 1) / PW_TILE;
    const int wgrid = (int)std::min<int64_t>(sms, n_tiles);
    rc = h->partials_buf.reserve((size_t)2 * sms * PICP_NACC * sizeof(float));
    if (rc) return rc;
    p.partials = h->partials_buf.as<float>();
    int rounds_arg = rounds;
    int64_t nw = h->n_world, ni = h->n_image;
    void* args[] = {(void*)&p, (void*)&rounds_arg, (void*)&nw, (void*)&ni};
    VO_CUDA(cudaMemsetAsync(h->barrier_buf.p, 0, 4, h->stream));
    VO_CUDA(cudaLaunchCooperativeKernel((const void*)wk, dim3((unsigned)wgrid), dim3(PW_THREADS + 32), args,
                                        PW_SMEM_BYTES, h->stream));
    VO_LAUNCH_CHECK();
    return VO_OK;
  }
  // streaming kernel: ONE cooperative launch runs every round (grid-wide barrier per round)
  {
    int rounds_arg = rounds;
    void* args[] = {(void*)&p, (void*)&rounds_arg};
    VO_CUDA(cudaMemsetAsync(h->barrier_buf.p, 0, 4, h->stream));
    VO_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)grid), dim3(PICP_THREADS),
                                        args, PICP_SMEM_BYTES, h->stream));
    VO_LAUNCH_CHECK();
  }
  return VO_OK;
}

int vo_picp_compute(vo_picp_t h, int keep_outliers, int rounds) {
  return picp_compute_common(h, keep_outliers, rounds, nullptr, nullptr);
}

int vo_picp_compute_ex(vo_picp_t h, int keep_outliers, int rounds, const int32_t* n_pairs_dev,
                       const float* pre_transform) {
  return picp_compute_common(h, keep_outliers, rounds, n_pairs_dev, pre_transform);
}

int vo_picp_one_round(vo_picp_t h, const int32_t* pairs_host, int64_t n_pairs, int keep_outliers) {
  int rc = vo_picp_set_correspondences(h, pairs_host, n_pairs);
  if (rc) return rc;
  return vo_picp_compute(h, keep_outliers, 1);
}

int vo_picp_get_state(vo_picp_t h, vo_picp_state* out) {
  VO_REQUIRE(h != nullptr && out != nullptr, VO_ERR_ARG, "null pointer");
  VO_REQUIRE(h->have_cam, VO_ERR_STATE, "init not called");
  DeviceGuard g(h->device);
  VO_CUDA(cudaMemcpyAsync(out, h->state_buf.p, sizeof(vo_picp_state), cudaMemcpyDeviceToHost,
                          h->stream));
  VO_CUDA(cudaStreamSynchronize(h->stream));
  return VO_OK;
}

}  // extern "C"
