// Does DMMA.8x8x4 block the SMSP issue port on B200?  Time DMMA alone vs DMMA + N independent integer
// instructions per DMMA, and DFMA (8 per "DMMA-equivalent") + the same integer work.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MATH, int NALU>  // MATH: 0 none, 1 DMMA x4, 2 DFMA x32 (same flops as 4 DMMA)
__global__ void __launch_bounds__(1024) k(double* out, int iters, double x, unsigned seed) {
  double c0[4], c1[4];
  double acc[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 4; i++) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i] = i + threadIdx.x * 1e-9; u[i] = seed + i * 77u + threadIdx.x; }
  const double a = x, b = 1.0 - x;
  for (int it = 0; it < iters; it++) {
    if (MATH == 1) {
#pragma unroll
      for (int i = 0; i < 4; i++) dmma(c0[i], c1[i], a, b);
    } else if (MATH == 2) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int j = 0; j < NALU * 4; j++) u[j & 7] = (u[j & 7] ^ (u[(j + 1) & 7] >> 3)) + 0x9e3779b9u;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MATH, int NALU>
void run(const char* name) {
  const int blocks = 148, iters = 5000;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MATH, NALU><<<blocks, 1024>>>(out, iters, 0.999, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MATH, NALU><<<blocks, 1024>>>(out, iters, 0.999, 1);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // per SMSP: 8 warps x iters x (4 DMMA or 32 DFMA) ; cycles at 1.965 GHz
  double cyc = ms * 1e-3 * 1.965e9;
  printf("%-28s %.3f ms  cycles/iter/SMSP = %.1f  (8 warps x 4 dmma-equivalents; alu instr/iter/SMSP = %d)  %s\n", name, ms,
         cyc / iters, 8 * NALU * 4 * 3, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<1, 0>("dmma only");
  run<1, 2>("dmma + 2x4 alu-triples");
  run<1, 4>("dmma + 4x4 alu-triples");
  run<1, 8>("dmma + 8x4 alu-triples");
  run<0, 2>("alu 2x4 only");
  run<0, 4>("alu 4x4 only");
  run<0, 8>("alu 8x4 only");
  run<2, 0>("dfma only");
  run<2, 2>("dfma + 2x4 alu-triples");
  run<2, 4>("dfma + 4x4 alu-triples");
  run<2, 8>("dfma + 8x4 alu-triples");
  return 0;
}
