// hmma_probe.cu — throughput of the legacy warp-level tensor-core path (mma.sync m16n8k16, f16 inputs)
// on sm_100a, with f16 and f32 accumulators, alone and interleaved 1:1 with the packed 16-bit 3-input
// minimum (VIMNMX3.S16x2).  Question behind it: a nearest-neighbour filter built on mma.sync keeps its
// accumulators in registers (no TMEM round trip, the limit of nn_tc_filter_kernel); it needs one
// m16n8k16 and one VIMNMX3 per 128 (query, row) pairs.  Not part of libvo_b200.so.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/hmma_probe tools/hmma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void hmma16(uint32_t (&d)[2], const uint32_t (&a)[4], const uint32_t (&b)[2], const uint32_t (&c)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%8,%9};"
               : "=r"(d[0]), "=r"(d[1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(c[0]), "r"(c[1]));
}
__device__ __forceinline__ void hmma32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2], const float (&c)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]),
                 "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ uint32_t min3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("{.reg .b32 t;\nmin.s16x2 t, %1, %2;\nmin.s16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// MODE 0: f16-accumulator HMMA, D independent of C (C = 0) -> fully pipelined
// MODE 1: f32-accumulator HMMA, 8 independent accumulation chains
// MODE 2: MODE 0 + one VIMNMX3.S16x2 per HMMA (the filter's inner loop)
template <int MODE>
__global__ void __launch_bounds__(1024) hmma_kernel(int iters, const uint32_t* in, uint32_t* sink, long long* cycles) {
  uint32_t a[8][4], b[2], z[2] = {0u, 0u}, mn[8];
  float acc[8][4];
  for (int i = 0; i < 8; ++i) {
    for (int k = 0; k < 4; ++k) { a[i][k] = in[(threadIdx.x + i * 4 + k) & 63]; acc[i][k] = 0.f; }
    mn[i] = 0x7fff7fffu;
  }
  b[0] = in[threadIdx.x & 63]; b[1] = in[(threadIdx.x + 7) & 63];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 1) {
        hmma32(acc[i], a[i], b, acc[i]);
      } else {
        uint32_t d[2];
        hmma16(d, a[i], b, z);
        if (MODE == 2) mn[i] = min3_s16x2(mn[i], d[0], d[1]);
        else { mn[i] ^= d[0]; }
      }
    }
    b[0] += 0x00010001u;  // a new B fragment per sweep
  }
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= mn[i] ^ __float_as_uint(acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3]);
  if (s == 0x12345678u) sink[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- accuracy of the f16-accumulator mma.sync on the NN filter's operands --------------------------
// One warp, one m16n8k16: A = 16 query rows, B = 8 map rows (K-major), C = 0.  Fragment layout (PTX ISA):
// a0:(g, 2t..2t+1) a1:(g+8, 2t..) a2:(g, 2t+8..) a3:(g+8, 2t+8..); b0:(k=2t.., n=g) b1:(k=2t+8.., n=g);
// d0:(g, 2t..2t+1) d1:(g+8, 2t..2t+1), g = lane/4, t = lane%4.
#include <cuda_fp16.h>
__global__ void hmma_acc_kernel(const __half* A /*16x16 row-major*/, const __half* B /*8x16 row-major (n, k)*/,
                                __half* D /*16x8*/) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  auto ld2 = [](const __half* p) { return *reinterpret_cast<const uint32_t*>(p); };
  uint32_t a[4] = {ld2(A + g * 16 + 2 * t), ld2(A + (g + 8) * 16 + 2 * t), ld2(A + g * 16 + 2 * t + 8),
                   ld2(A + (g + 8) * 16 + 2 * t + 8)};
  uint32_t b[2] = {ld2(B + g * 16 + 2 * t), ld2(B + g * 16 + 2 * t + 8)};
  uint32_t z[2] = {0u, 0u}, d[2];
  hmma16(d, a, b, z);
  *reinterpret_cast<uint32_t*>(D + g * 8 + 2 * t) = d[0];
  *reinterpret_cast<uint32_t*>(D + (g + 8) * 8 + 2 * t) = d[1];
}
static unsigned rs = 777u;
static float frand() { rs = rs * 1664525u + 1013904223u; return (float)((rs >> 8) & 0xFFFFFF) / 8388608.f - 1.f; }
static void split16(float x, __half* hi, __half* lo) { const __half h = __float2half_rn(x); *hi = h; *lo = __float2half_rn(x - __half2float(h)); }
static void acc_check() {
  __half *dA, *dB, *dD;
  CK(cudaMalloc(&dA, 512)); CK(cudaMalloc(&dB, 256)); CK(cudaMalloc(&dD, 256));
  double worst = 0.0, worst_near = 0.0;
  long long n = 0;
  for (int tile = 0; tile < 2000; ++tile) {
    __half A[256], B[128], D[128];
    float q[16][10], m[8][10];
    for (auto& r : q) for (float& v : r) v = frand();
    for (int j = 0; j < 8; ++j) for (int k = 0; k < 10; ++k) m[j][k] = (j & 1) ? q[j][k] + 0.02f * frand() : frand();
    for (int i = 0; i < 16; ++i) {
      float qq = 0; for (int k = 0; k < 10; ++k) { A[i * 16 + k] = __float2half_rn(-2.f * q[i][k]); qq += q[i][k] * q[i][k]; }
      A[i * 16 + 10] = A[i * 16 + 11] = __float2half_rn(1.f); split16(qq, &A[i * 16 + 12], &A[i * 16 + 13]);
      A[i * 16 + 14] = A[i * 16 + 15] = __float2half_rn(0.f);
    }
    for (int j = 0; j < 8; ++j) {
      float mm = 0; for (int k = 0; k < 10; ++k) { B[j * 16 + k] = __float2half_rn(m[j][k]); mm += m[j][k] * m[j][k]; }
      split16(mm, &B[j * 16 + 10], &B[j * 16 + 11]); B[j * 16 + 12] = B[j * 16 + 13] = __float2half_rn(1.f);
      B[j * 16 + 14] = B[j * 16 + 15] = __float2half_rn(0.f);
    }
    CK(cudaMemcpy(dA, A, 512, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B, 256, cudaMemcpyHostToDevice));
    hmma_acc_kernel<<<1, 32>>>(dA, dB, dD);
    CK(cudaMemcpy(D, dD, 256, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) {
      double exact = 0, sum_abs = 0;
      for (int k = 0; k < 16; ++k) { const double t = (double)__half2float(A[i * 16 + k]) * (double)__half2float(B[j * 16 + k]); exact += t; sum_abs += fabs(t); }
      const double e = fabs((double)__half2float(D[i * 8 + j]) - exact) - (fabs(exact) * 4.8828125e-4 + 3e-8);
      const double r = e / sum_abs;
      if (r > worst) worst = r;
      if (fabs(exact) < 0.05 && r > worst_near) worst_near = r;
      ++n;
    }
  }
  printf("{\"probe\": \"mma.sync f16-accumulator accuracy\", \"elements\": %lld, "
         "\"max_err_beyond_half_ulp_over_sum_abs_terms\": %.3g, \"same_for_results_near_zero\": %.3g}\n", n, worst, worst_near);
}

int main() {
  acc_check();
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint32_t *sink, *din; long long* cyc;
  CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&din, 256)); CK(cudaMalloc(&cyc, 8 * 1024));
  uint32_t h[64];
  for (int i = 0; i < 64; ++i) h[i] = 0x3c003800u + i * 0x00010001u;
  CK(cudaMemcpy(din, h, 256, cudaMemcpyHostToDevice));
  std::vector<long long> hc(1024);
  const char* names[] = {"f16 acc", "f32 acc", "f16 acc + VIMNMX3.S16x2"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      const int iters = 4000;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) hmma_kernel<0><<<sms, warps * 32>>>(iters, din, sink, cyc);
        else if (mode == 1) hmma_kernel<1><<<sms, warps * 32>>>(iters, din, sink, cyc);
        else hmma_kernel<2><<<sms, warps * 32>>>(iters, din, sink, cyc);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(hc.data(), cyc, 8 * sms, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = hc[i] > mx ? hc[i] : mx;
      const double n = (double)iters * 8 * warps;
      printf("{\"probe\": \"mma.sync m16n8k16\", \"mode\": \"%s\", \"warps\": %d, \"hmma_per_clk_per_sm\": %.3f, "
             "\"mac_per_clk_per_sm\": %.0f, \"pairs_per_clk_per_sm\": %.1f}\n",
             names[mode], warps, n / mx, n / mx * 2048, n / mx * 128);
    }
  return 0;
}
