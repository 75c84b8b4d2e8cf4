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

int main() {
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
