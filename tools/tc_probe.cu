// tc_probe.cu — stand-alone micro-benchmarks that decide whether a tcgen05 nearest-neighbour filter
// can beat the FFMA2 one.  Not part of libvo_b200.so.
//   (1) layout check: D = A * B^T with f16 inputs in the no-swizzle K-major "interleaved" layout the
//       NN kernel would use, read back with tcgen05.ld and compared with the host;
//   (2) tcgen05.ld (TMEM -> registers) throughput per SM with 4/8/16 warps and a min-reduction
//       consumer (the whole epilogue of the filter);
//   (3) tcgen05.mma issue rate for M=128, N=256, K=16 (kind::f16) and 2 x K=8 (kind::tf32).
// Every wait is bounded: a wrong descriptor ends in a "timeout" line, never in a hung GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tc_probe tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
// bounded wait: false on timeout
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long max_cycles) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > max_cycles) return false;
  return true;
}
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
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
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
__device__ __forceinline__ float f_min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// shared-memory matrix descriptor, K-major, no swizzle ("interleaved"): a core matrix is 8 rows x
// 16 bytes stored as 128 contiguous bytes; lbo = byte distance between the two 16-byte K chunks of
// one instruction, sbo = byte distance between consecutive 8-row groups.
__host__ __device__ inline uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor: c_format [4,6) 1=F32; a_format [7,10), b_format [10,13): 0=F16 1=BF16 2=TF32;
// a_major bit 15, b_major bit 16 (0 = K-major); n_dim [17,23) = N>>3; m_dim [24,29) = M>>4
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- (1) layout check -----------------------------------------------------------------------------
// A: 128 x 16 halves, B: 256 x 16 halves, both given row-major; the kernel stores them interleaved
// (8-row groups: [8 rows x first 16 B][8 rows x second 16 B]) and computes D = A B^T (128 x 256 f32).
__global__ void __launch_bounds__(128) layout_check_kernel(const __half* A, const __half* B, float* D,
                                                           int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* sA = reinterpret_cast<__half*>(smem);             // 128 * 32 B = 4 KB
  __half* sB = reinterpret_cast<__half*>(smem + 4096);      // 256 * 32 B = 8 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  auto put = [](__half* dst, const __half* src, int rows) {
    for (int i = threadIdx.x; i < rows * 16; i += blockDim.x) {
      const int r = i / 16, k = i % 16;
      const int off = (r / 8) * 128 + (k / 8) * 64 + (r % 8) * 8 + (k % 8);  // in halves
      dst[off] = src[i];
    }
  };
  put(sA, A, 128);
  put(sB, B, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  // generic-proxy writes to smem must be visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint64_t ad = make_desc(smem_u32(sA), 128, 256);
    const uint64_t bd = make_desc(smem_u32(sB), 128, 256);
    tc_mma_f16(tb, ad, bd, make_idesc(0, 128, 256), 0);
    tc_commit(&bar);
  }
  const bool ok = mbar_wait_bounded(&bar, 0, 200000000LL);
  if (!ok) {
    if (tid == 0) *status = -1;
  } else {
    tc_fence_after();
    for (int c0 = 0; c0 < 256; c0 += 32) {
      uint32_t r[32];
      tc_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
      tc_wait_ld();
      for (int j = 0; j < 32; ++j) D[(size_t)tid * 256 + c0 + j] = __uint_as_float(r[j]);
    }
    if (tid == 0) *status = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

// ---- (2) tcgen05.ld throughput --------------------------------------------------------------------
// Every warp reads `cols` columns of its 32-lane quadrant per iteration (32 columns per instruction)
// and folds them into a running minimum with 3-input mins, as the filter epilogue would.
template <int BATCH>  // tcgen05.ld instructions issued before one wait::ld
__global__ void __launch_bounds__(512) ldtm_kernel(int iters, int cols, float* sink, long long* cycles) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float mn = INFINITY;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c0 = 0; c0 < cols; c0 += 32 * BATCH) {
      uint32_t r[BATCH][32];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) tc_ld32(tb + ((c0 + 32 * b) & 511), r[b]);
      tc_wait_ld();
#pragma unroll
      for (int b = 0; b < BATCH; ++b)
#pragma unroll
        for (int j = 0; j < 32; j += 2) mn = f_min3(mn, __uint_as_float(r[b][j]), __uint_as_float(r[b][j + 1]));
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (mn == 123.456f) sink[0] = mn;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- (3) MMA issue rate, alone and with a concurrent epilogue ---------------------------------------
// thread 0 issues `n_mma` groups (one f16 K=16 instruction, or two tf32 K=8 instructions) alternating
// between two 256-column accumulators; with EPI, warps 4..7 keep reading the accumulators meanwhile.
template <bool TF32, bool EPI>
__global__ void __launch_bounds__(256) mma_rate_kernel(int n_mma, float* sink, long long* cycles, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    done = 0;
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  float mn = INFINITY;
  if (tid == 0) {
    // f16: rows of 32 B (2 chunks);  tf32: rows of 64 B (4 chunks, two instructions of 2 chunks each)
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 8192);
    const uint32_t idesc = make_idesc(TF32 ? 2 : 0, 128, 256);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tb + (uint32_t)(i & 1) * 256;
      if (TF32) {
        tc_mma_tf32(d, make_desc(a0, 128, 512), make_desc(b0, 128, 512), idesc, 0);
        tc_mma_tf32(d, make_desc(a0 + 256, 128, 512), make_desc(b0 + 256, 128, 512), idesc, 1);
      } else {
        tc_mma_f16(d, make_desc(a0, 128, 256), make_desc(b0, 128, 256), idesc, 0);
      }
    }
    tc_commit(&bar);
    const bool ok = mbar_wait_bounded(&bar, 0, 2000000000LL);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0) *status = ok ? 1 : -1;
    done = 1;
  } else if (EPI && warp >= 4) {
    const uint32_t tq = tb + ((uint32_t)((warp & 3) * 32) << 16);
    while (!done) {
      for (int c0 = 0; c0 < 512; c0 += 64) {
        uint32_t r[2][32];
        tc_ld32(tq + c0, r[0]);
        tc_ld32(tq + c0 + 32, r[1]);
        tc_wait_ld();
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int j = 0; j < 32; j += 2) mn = f_min3(mn, __uint_as_float(r[b][j]), __uint_as_float(r[b][j + 1]));
      }
    }
  }
  if (mn == 123.456f) sink[0] = mn;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ---- (4) dependent latency: issue one MMA, commit, wait for the barrier, repeat --------------------
template <int N>
__global__ void __launch_bounds__(128) mma_latency_kernel(int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint64_t ad = make_desc(smem_u32(smem), 128, 256), bd = make_desc(smem_u32(smem + 8192), 128, 256);
    const uint32_t idesc = make_idesc(0, 128, N);
    bool ok = true;
    const long long t0 = clock64();
    for (int i = 0; i < iters && ok; ++i) {
      tc_mma_f16(tb, ad, bd, idesc, 0);
      tc_commit(&bar);
      ok = mbar_wait_bounded(&bar, (uint32_t)(i & 1), 100000000LL);
      tc_fence_after();
    }
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0) *status = ok ? 1 : -1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ---- (5) dependent latency of one tcgen05.ld.x32 + wait::ld, single warp ---------------------------
__global__ void __launch_bounds__(32) ldtm_latency_kernel(int iters, float* sink, long long* cycles) {
  __shared__ uint32_t tmem_base;
  tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncwarp();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  float mn = INFINITY;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[32];
    tc_ld32(tb + ((it * 32) & 511), r);
    tc_wait_ld();
    mn = f_min3(mn, __uint_as_float(r[0]), __uint_as_float(r[31]));
  }
  const long long t1 = clock64();
  if (mn == 123.456f) sink[0] = mn;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  tc_fence_before();
  __syncwarp();
  tmem_dealloc(tb, 512);
}

// ---- (6) issue cost: back-to-back MMAs of width N, optionally one commit per MMA, from 1 or 2 warps ----
template <int N>
__global__ void __launch_bounds__(128) mma_issue_kernel(int n_mma, int commit_each, int n_issuers, long long* cycles,
                                                        int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint64_t done_bar[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1 << 20);  // never completes: commits just arrive
    mbar_init(&done_bar[0], 1);
    mbar_init(&done_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp < n_issuers && lane == 0) {
    const uint64_t ad = make_desc(smem_u32(smem), 128, 256), bd = make_desc(smem_u32(smem + 8192), 128, 256);
    const uint32_t idesc = make_idesc(0, 128, N);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      tc_mma_f16(tb + (uint32_t)(((i & 1) * 2 + warp) * 128) % 512, ad, bd, idesc, 0);
      if (commit_each) tc_commit(&bar[(i & 1) * 2 + warp]);
    }
    const long long t_issue = clock64() - t0;
    tc_commit(&done_bar[warp]);
    const bool ok = mbar_wait_bounded(&done_bar[warp], 0, 2000000000LL);
    const long long t_all = clock64() - t0;
    cycles[(blockIdx.x * 2 + warp) * 2] = t_issue;
    cycles[(blockIdx.x * 2 + warp) * 2 + 1] = t_all;
    if (blockIdx.x == 0 && warp == 0) *status = ok ? 1 : -1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  printf("{\"probe\": \"device\", \"sms\": %d, \"clock_khz\": %d}\n", sms, khz);

  // (1) layout check with small integers (exact in f16 and f32)
  {
    std::vector<__half> A(128 * 16), B(256 * 16);
    std::vector<float> Af(128 * 16), Bf(256 * 16);
    for (int i = 0; i < 128 * 16; ++i) {
      Af[i] = (float)((i * 7 + (i / 16) * 3) % 13 - 6);
      A[i] = __float2half(Af[i]);
    }
    for (int i = 0; i < 256 * 16; ++i) {
      Bf[i] = (float)((i * 5 + (i / 16) * 11) % 9 - 4);
      B[i] = __float2half(Bf[i]);
    }
    __half *dA, *dB;
    float* dD;
    int* dS;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 2));
    CK(cudaMalloc(&dD, 128 * 256 * 4));
    CK(cudaMalloc(&dS, 4));
    CK(cudaMemset(dS, 0, 4));
    CK(cudaMemset(dD, 0xFF, 128 * 256 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    layout_check_kernel<<<1, 128, 12288>>>(dA, dB, dD, dS);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * 256);
    int st = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    int bad = 0, first_bad = -1;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 256; ++j) {
        float ref = 0.f;
        for (int k = 0; k < 16; ++k) ref += Af[i * 16 + k] * Bf[j * 16 + k];
        if (D[i * 256 + j] != ref) {
          if (first_bad < 0) first_bad = i * 256 + j;
          ++bad;
        }
      }
    printf("{\"probe\": \"layout_check\", \"status\": %d, \"mismatches\": %d, \"first_bad\": %d", st, bad, first_bad);
    if (first_bad >= 0) {
      float ref = 0.f;
      const int i = first_bad / 256, j = first_bad % 256;
      for (int k = 0; k < 16; ++k) ref += Af[i * 16 + k] * Bf[j * 16 + k];
      printf(", \"got\": %g, \"want\": %g", D[first_bad], ref);
    }
    printf("}\n");
  }

  float* sink;
  long long* cyc;
  int* st;
  CK(cudaMalloc(&sink, 64));
  CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
  CK(cudaMalloc(&st, 4));
  std::vector<long long> h(1024);

  // (2) tcgen05.ld throughput: one CTA per SM, 4 / 8 / 16 warps
  for (int batch = 1; batch <= 4; batch *= 2)
    for (int warps = 4; warps <= 16; warps *= 2) {
      const int iters = 2000, cols = 256;
      for (int rep = 0; rep < 2; ++rep) {
        if (batch == 1) ldtm_kernel<1><<<sms, warps * 32>>>(iters, cols, sink, cyc);
        else if (batch == 2) ldtm_kernel<2><<<sms, warps * 32>>>(iters, cols, sink, cyc);
        else ldtm_kernel<4><<<sms, warps * 32>>>(iters, cols, sink, cyc);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)warps * iters * cols * 32 * 4;
      printf("{\"probe\": \"ldtm\", \"warps\": %d, \"batch\": %d, \"cycles\": %lld, \"bytes_per_clk_per_sm\": %.1f, "
             "\"elems_per_clk_per_sm\": %.1f}\n",
             warps, batch, mx, bytes / mx, bytes / 4 / mx);
    }

  // (3) MMA rate
  for (int variant = 0; variant < 4; ++variant) {
    const int n = 4000;
    CK(cudaMemset(st, 0, 4));
    auto launch = [&]() {
      switch (variant) {
        case 0: mma_rate_kernel<false, false><<<sms, 256, 24576>>>(n, sink, cyc, st); break;
        case 1: mma_rate_kernel<false, true><<<sms, 256, 24576>>>(n, sink, cyc, st); break;
        case 2: mma_rate_kernel<true, false><<<sms, 256, 24576>>>(n, sink, cyc, st); break;
        default: mma_rate_kernel<true, true><<<sms, 256, 24576>>>(n, sink, cyc, st); break;
      }
    };
    launch();
    CK(cudaDeviceSynchronize());
    launch();
    CK(cudaDeviceSynchronize());
    int s = 0;
    CK(cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("{\"probe\": \"mma_rate\", \"kind\": \"%s\", \"concurrent_epilogue\": %d, \"status\": %d, "
           "\"cycles_per_128x256x16_tile\": %.1f}\n",
           variant >= 2 ? "tf32" : "f16", variant & 1, s, (double)mx / n);
  }
  // (4) dependent MMA latency
  for (int variant = 0; variant < 3; ++variant) {
    const int iters = 2000;
    CK(cudaMemset(st, 0, 4));
    for (int rep = 0; rep < 2; ++rep) {
      if (variant == 0) mma_latency_kernel<64><<<sms, 128, 24576>>>(iters, cyc, st);
      else if (variant == 1) mma_latency_kernel<128><<<sms, 128, 24576>>>(iters, cyc, st);
      else mma_latency_kernel<256><<<sms, 128, 24576>>>(iters, cyc, st);
      CK(cudaDeviceSynchronize());
    }
    int sflag = 0;
    CK(cudaMemcpy(&sflag, st, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("{\"probe\": \"mma_dependent_latency\", \"N\": %d, \"status\": %d, \"cycles_issue_to_barrier\": %.1f}\n",
           variant == 0 ? 64 : (variant == 1 ? 128 : 256), sflag, (double)mx / iters);
  }
  // (5) dependent tcgen05.ld latency
  {
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      ldtm_latency_kernel<<<1, 32>>>(iters, sink, cyc);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("{\"probe\": \"ldtm_dependent_latency\", \"cycles_per_x32_load\": %.1f}\n", (double)h[0] / iters);
  }
  // (6) issue cost
  for (int nsel = 0; nsel < 3; ++nsel)
    for (int commit_each = 0; commit_each < 2; ++commit_each)
      for (int issuers = 1; issuers <= 2; ++issuers) {
        const int n = 2000;
        CK(cudaMemset(st, 0, 4));
        CK(cudaMemset(cyc, 0, sizeof(long long) * 1024));
        for (int rep = 0; rep < 2; ++rep) {
          if (nsel == 0) mma_issue_kernel<64><<<sms, 128, 24576>>>(n, commit_each, issuers, cyc, st);
          else if (nsel == 1) mma_issue_kernel<128><<<sms, 128, 24576>>>(n, commit_each, issuers, cyc, st);
          else mma_issue_kernel<256><<<sms, 128, 24576>>>(n, commit_each, issuers, cyc, st);
          CK(cudaDeviceSynchronize());
        }
        int sflag = 0;
        CK(cudaMemcpy(&sflag, st, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * 4, cudaMemcpyDeviceToHost));
        printf("{\"probe\": \"mma_issue\", \"N\": %d, \"commit_each\": %d, \"issuers\": %d, \"status\": %d, "
               "\"issue_cycles_per_mma\": %.1f, \"total_cycles_per_mma\": %.1f}\n",
               nsel == 0 ? 64 : (nsel == 1 ? 128 : 256), commit_each, issuers, sflag, (double)h[0] / n, (double)h[1] / n);
      }
  return 0;
}
