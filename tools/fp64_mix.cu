// Micro-benchmark: do DMMA (mma.sync.m8n8k4.f64) and DFMA share one FP64 datapath on B200, or can they overlap?
// Variants: (a) DMMA-only warps, (b) DFMA-only warps, (c) half the warps DMMA + half DFMA in one CTA,
// (d) both instruction kinds interleaved in every warp.  Reports the summed FP64 flop rate.
#include <cstdio>
#include <cuda_runtime.h>

template <int NMMA, int NFMA>
__global__ void __launch_bounds__(512) mix_loop(double* out, int iters, int split) {
  // split = 0: every warp runs NMMA DMMA + NFMA DFMA per iteration;
  // split = 1: even warps run only the DMMA part, odd warps only the DFMA part.
  double c[NMMA > 0 ? NMMA : 1][2];
  double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < (NMMA > 0 ? NMMA : 1); i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); i++) f[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const int warp = threadIdx.x >> 5;
  const bool do_mma = !split || (warp & 1) == 0, do_fma = !split || (warp & 1) == 1;
  for (int it = 0; it < iters; it++) {
    if (do_mma) {
#pragma unroll
      for (int i = 0; i < NMMA; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    if (do_fma) {
#pragma unroll
      for (int i = 0; i < NFMA; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < (NMMA > 0 ? NMMA : 1); i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < (NFMA > 0 ? NFMA : 1); i++) s += f[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

template <int NMMA, int NFMA>
void run(const char* name, int sms, int warps, int split, double* out) {
  const int iters = 20000;
  float ms = time_ms([&] { mix_loop<NMMA, NFMA><<<sms, warps * 32>>>(out, iters, split); });
  const double wm = split ? warps / 2.0 : warps, wf = split ? warps / 2.0 : warps;
  const double fm = 2.0 * 256 * NMMA * (double)iters * wm * sms, ff = 2.0 * 32 * NFMA * (double)iters * wf * sms;
  printf("%-28s warps/SM %2d split %d: %8.3f ms  DMMA %6.2f + DFMA %6.2f = %6.2f TFLOP/s\n", name, warps, split, ms,
         fm / ms * 1e-9, ff / ms * 1e-9, (fm + ff) / ms * 1e-9);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s sms %d\n", p.name, sms);
  double* out; cudaMalloc(&out, 8);
  for (int warps : {8, 16}) {
    run<8, 0>("DMMA only (8 acc)", sms, warps, 0, out);
    run<0, 16>("DFMA only (16 acc)", sms, warps, 0, out);
    run<8, 16>("DMMA8 + DFMA16 split warps", sms, warps, 1, out);
    run<8, 16>("DMMA8 + DFMA16 interleaved", sms, warps, 0, out);
    run<8, 32>("DMMA8 + DFMA32 interleaved", sms, warps, 0, out);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
