// common.cuh — shared host/device helpers for libvo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/vo_b200.h"

namespace vo {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define VO_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::vo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return VO_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define VO_REQUIRE(cond, code, msg)                        \
  do {                                                     \
    if (!(cond)) {                                         \
      ::vo::set_error("%s: %s", __func__, msg);            \
      return code;                                         \
    }                                                      \
  } while (0)

// every kernel launch of the library goes through this so bench.py can report gpu_launches
#define VO_LAUNCH_CHECK()                         \
  do {                                            \
    ::vo::g_launches.fetch_add(1);                \
    VO_CUDA(cudaGetLastError());                  \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return VO_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
      return VO_ERR_CUDA;
    }
    cap = bytes;
    return VO_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

int num_sms(int device);
// internal (not part of the C ABI): device address of a solver's vo_picp_state, for callers inside
// the library that fetch it together with other results in one synchronisation
const void* picp_state_device_ptr(vo_picp_t h);

// After cudaMemcpyAsync FROM `host_ptr`: may the caller reuse / free the buffer without a stream
// synchronisation?  A copy from pageable memory has been staged by the runtime when the call
// returns; only a copy from pinned (or managed) memory is still reading the caller's buffer.
inline bool host_source_still_in_use(const void* host_ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) {
    cudaGetLastError();
    return true;  // unknown: be safe
  }
  return a.type != cudaMemoryTypeUnregistered;
}

inline bool host_source_is_pageable(const void* host_ptr) { return !host_source_still_in_use(host_ptr); }

// Pinned, chunked staging between pageable host memory and the device (stage.cu): a pool of
// host threads copies through a ring of pinned chunks while the DMA engine moves the previous
// one.  stage_h2d returns as soon as the caller's buffer is no longer referenced (the tail may
// still be in flight on `s`); stage_d2h returns with the data in dst_host.  Transfers below
// 1 MB, and pinned / managed host memory, take a plain cudaMemcpyAsync.
int stage_h2d(int device, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s);
int stage_d2h(int device, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s);

// ---- device-side PTX wrappers (TMA bulk copy + mbarrier) ----------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  // make the initialised barriers visible to the async (TMA) proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// packed FP32 pairs: sm_100a executes fma.rn.f32x2 as ONE FFMA2 issue slot for two FMAs, and
// ptxas folds a {x,x} pair into a scalar-broadcast operand (FFMA2 Rd, Ra.F32x2, Rb.F32, Rc.F32x2),
// so two queries share every map coefficient without any packing instruction.
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b,
                                                     unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long f2_bc(float x) { return f2_pack(x, x); }
// A product that is rounded ON ITS OWN.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one
// FFMA2 even under --fmad=false (it does not for the scalar .rn forms), which would remove the
// product's rounding.  fma(a, b, +0.0) is the same IEEE product (only a -0 result becomes +0) and
// cannot be folded into a following add, so code that must match a non-FMA CPU uses this.
__device__ __forceinline__ unsigned long long f2_prod(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("{\n.reg .b64 z;\nmov.b64 z, 0;\nfma.rn.f32x2 %0, %1, %2, z;\n}" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#endif

}  // namespace vo
