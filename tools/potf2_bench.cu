// Micro-benchmark / self-check of the diagonal-block kernels (development aid): k_potf2 vs k_potf2_v2 on one random SPD
// 128x128 block: max differences of L, L^-1, L^-T, event timing of 200 back-to-back launches and, with -DP2_TIMING,
// the clock64 stamps of the phases of k_potf2_v2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DP2_TIMING -o tools/potf2_bench tools/potf2_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../include/dgp.h"
#include "../discontinuum_b200/csrc/dgp_cov.cuh"
#include "../discontinuum_b200/csrc/dgp_gemm.cuh"
#include "../discontinuum_b200/csrc/dgp_panel.cuh"
#include "../discontinuum_b200/csrc/dgp_potf2.cuh"
using namespace dgp;

int main() {
  const int ld = 256;
  std::vector<double> A(128 * ld, 0.0), G(128 * 128);
  srand(1);
  for (auto& g : G) g = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < 128; i++)
    for (int j = 0; j <= i; j++) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < 128; k++) s += G[i * 128 + k] * G[j * 128 + k] / 16.0;
      A[i * ld + j] = s;
      A[j * ld + i] = 777.0;  // garbage above the diagonal must be ignored
      if (i == j) A[i * ld + j] = s;
    }
  double *dA, *dL[2], *dU[2], *dDI[2], *dT[2], *scal;
  cudaMalloc(&dA, 128 * ld * 8); cudaMalloc(&scal, 64 * 8);
  for (int v = 0; v < 2; v++) { cudaMalloc(&dL[v], 128 * ld * 8); cudaMalloc(&dU[v], 128 * ld * 8); cudaMalloc(&dDI[v], 128 * 128 * 8); cudaMalloc(&dT[v], 128 * ld * 8);
    cudaMemset(dL[v], 0xff, 128 * ld * 8); cudaMemset(dU[v], 0xff, 128 * ld * 8); cudaMemset(dT[v], 0xff, 128 * ld * 8); }
  cudaMemcpy(dA, A.data(), 128 * ld * 8, cudaMemcpyHostToDevice);
  cudaMemset(scal, 0, 64 * 8);
  cudaFuncSetAttribute(k_potf2, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM);
  cudaFuncSetAttribute(k_potf2_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int v = 0; v < 2; v++) {
    for (int rep = 0; rep < 2; rep++) {
      const int iters = rep ? 200 : 3;
      cudaEventRecord(e0);
      for (int it = 0; it < iters; it++) {
        if (v == 0) k_potf2<<<1, PF_THREADS, PF_SMEM>>>(dA, dL[v], dU[v], ld, dDI[v], scal, 0, dT[v]);
        else k_potf2_v2<<<1, P2_THREADS, P2_SMEM>>>(dA, dL[v], dU[v], ld, getenv("P2_FULL") ? dDI[v] : nullptr, scal, 0, dT[v], getenv("P2_FULL") ? 1 : 0, P2Batch{});
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("v%d: %.2f us per launch (%s)\n", v + 1, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
    }
  }
  std::vector<double> L[2], U[2], DI[2], T[2];
  for (int v = 0; v < 2; v++) {
    L[v].resize(128 * ld); U[v].resize(128 * ld); DI[v].resize(128 * 128); T[v].resize(128 * ld);
    cudaMemcpy(L[v].data(), dL[v], 128 * ld * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(U[v].data(), dU[v], 128 * ld * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(DI[v].data(), dDI[v], 128 * 128 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(T[v].data(), dT[v], 128 * ld * 8, cudaMemcpyDeviceToHost);
  }
  double dl = 0, du = 0, dd = 0, dt = 0;
  for (int i = 0; i < 128; i++)
    for (int j = 0; j < 128; j++) {
      dl = fmax(dl, fabs(L[0][i * ld + j] - L[1][i * ld + j]));
      du = fmax(du, fabs(U[0][i * ld + j] - U[1][i * ld + j]));
      dt = fmax(dt, fabs(T[0][i * ld + j] - T[1][i * ld + j]));
      dd = fmax(dd, fabs(DI[0][i * 128 + j] - DI[1][i * 128 + j]));
    }
  printf("max |v1 - v2|: L %.3e  U %.3e  DI %.3e  T %.3e\n", dl, du, dd, dt);
#ifdef P2_TIMING
  long long ts[64];
  cudaMemcpyFromSymbol(ts, p2_ts, sizeof(ts));
  printf("stamps (cycles from start): load %lld\n", ts[1] - ts[0]);
  for (int k = 0; k < 4; k++)
    printf(" k=%d pivots %lld  A-end %lld  (sync %lld)  B %lld  C %lld  D %lld\n", k, ts[2 + 6 * k] - (k ? ts[7 + 6 * (k - 1)] : ts[1]),
           ts[3 + 6 * k] - ts[2 + 6 * k], ts[4 + 6 * k] - ts[3 + 6 * k], ts[5 + 6 * k] - ts[4 + 6 * k], ts[6 + 6 * k] - ts[5 + 6 * k],
           ts[7 + 6 * k] - ts[6 + 6 * k]);
  printf(" inv row 3 %lld  outputs %lld  total %lld\n", ts[26] - ts[25], ts[27] - ts[26], ts[27] - ts[0]);
  long long wts[64];
  cudaMemcpyFromSymbol(wts, p2_wts, sizeof(wts));
  for (int b = 0; b < 2; b++)
    for (int k = 0; k < 4; k++) {
      printf(" barrier %d k=%d arrival of warps 0..7 (cycles after phase A start):", b, k);
      for (int w = 0; w < 8; w++) printf(" %lld", wts[b * 32 + k * 8 + w] - (k ? ts[7 + 6 * (k - 1)] : ts[1]));
      printf("\n");
    }
#endif
  return 0;
}
