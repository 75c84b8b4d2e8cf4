// lib.cu — library-level entry points of libvo_b200 (version, errors, launch counter, the FFMA
// micro-benchmark used as the FP32 roofline denominator).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vo {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int num_sms(int device) {
  static int cache[64];
  static bool have[64];
  if (device >= 0 && device < 64 && have[device]) return cache[device];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0)
    n = 148;
  if (device >= 0 && device < 64) {
    cache[device] = n;
    have[device] = true;
  }
  return n;
}

// 8 independent FMA chains per thread, operands in registers: issue-bound on the FMA pipe.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b);
      x1 = fmaf(x1, a, b);
      x2 = fmaf(x2, a, b);
      x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b);
      x5 = fmaf(x5, a, b);
      x6 = fmaf(x6, a, b);
      x7 = fmaf(x7, a, b);
    }
  }
  float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[0] = s;  // keep the chains alive
}

// the same loop with packed fma.rn.f32x2 (SASS FFMA2): two FMAs per lane per issue slot
__global__ void __launch_bounds__(256) ffma2_peak_kernel(float* out, int iters, float a, float b) {
  unsigned long long x[8], A, B;
  asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float v = threadIdx.x * 1e-3f + u;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x[u]) : "f"(v), "f"(v + 0.5f));
  }
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(A), "l"(B));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k]));
    s += lo + hi;
  }
  if (s == 123.456f) out[0] = s;
}

}  // namespace vo

using namespace vo;

extern "C" {

int vo_abi_version(void) { return VO_B200_ABI_VERSION; }

const char* vo_last_error(void) { return t_err; }

int vo_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
    return VO_ERR_CUDA;
  }
  return n;
}

int64_t vo_launch_count(void) { return g_launches.load(); }

static int measure_peak(int device, double* tflops_out, bool packed);

int vo_measure_ffma_peak(int device, double* tflops_out) {
  return measure_peak(device, tflops_out, false);
}
int vo_measure_ffma2_peak(int device, double* tflops_out) {
  return measure_peak(device, tflops_out, true);
}

static int measure_peak(int device, double* tflops_out, bool packed) {
  VO_REQUIRE(tflops_out != nullptr, VO_ERR_ARG, "null output");
  DeviceGuard g(device);
  float* d = nullptr;
  VO_CUDA(cudaMalloc(&d, sizeof(float)));
  const int sms = num_sms(device);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  VO_CUDA(cudaEventCreate(&e0));
  VO_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    VO_CUDA(cudaEventRecord(e0, 0));
    if (packed) ffma2_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 1e-3f);
    else ffma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 1e-3f);
    VO_LAUNCH_CHECK();
    VO_CUDA(cudaEventRecord(e1, 0));
    VO_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    VO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (packed ? 2.0 : 1.0) * 2.0 * 8 * 16 * (double)iters * threads * (double)blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return VO_OK;
}

}  // extern "C"
