// Micro-benchmark: raw FP64 pipe throughput on this GPU (DMMA.8x8x4 vs DFMA).
// Used once to pick the arithmetic instruction for the trailing-update kernels and to
// sanity-check the FP64 roofline denominator measured with cuBLAS (tools/fp64_peak.py).
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_loop(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_loop(double* out, int iters) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; cudaMalloc(&out, 8);
  int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { dmma_loop<16><<<sms, warps * 32>>>(out, iters); });
    double flop = 2.0 * 256 * 16 * (double)iters * warps * sms;
    printf("DMMA  warps/SM %2d  acc 16: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
  }
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { dmma_loop<4><<<sms, warps * 32>>>(out, iters); });
    double flop = 2.0 * 256 * 4 * (double)iters * warps * sms;
    printf("DMMA  warps/SM %2d  acc  4: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
  }
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { dfma_loop<16><<<sms, warps * 32>>>(out, iters); });
    double flop = 2.0 * 32 * 16 * (double)iters * warps * sms;
    printf("DFMA  warps/SM %2d  acc 16: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
