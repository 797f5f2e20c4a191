// Microbenchmark: FP64 FMA pipe vs FP64 tensor (DMMA m8n8k4) throughput on B200, alone and mixed.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -o gpurun_out/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0 = DFMA, 1 = DMMA, 2 = mixed
__global__ void __launch_bounds__(512) k(double* out, int iters, double x) {
  double acc[8];
  double c0[4], c1[4];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) { c0[i] = i; c1[i] = -i; }
  const double a = x, b = 1.0 - x;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; i++) dmma(c0[i], c1[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int blocks, int iters) {
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 512>>>(out, iters, 0.999);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 512>>>(out, iters, 0.999);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps = (double)blocks * 16;
  double fma_flops = (MODE == 0 || MODE == 2) ? warps * 32 * 8.0 * iters * 2 : 0;
  double mma_flops = (MODE == 1 || MODE == 2) ? warps * 4.0 * iters * (8 * 8 * 4 * 2) : 0;
  printf("%-6s blocks=%d  %.3f ms  DFMA %.2f TF/s  DMMA %.2f TF/s  total %.2f TF/s  err=%s\n", name, blocks, ms,
         fma_flops / ms / 1e9, mma_flops / ms / 1e9, (fma_flops + mma_flops) / ms / 1e9,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int blocks : {148, 296, 592}) {
    run<0>("dfma", blocks, 20000);
    run<1>("dmma", blocks, 20000);
    run<2>("mixed", blocks, 20000);
  }
  return 0;
}
