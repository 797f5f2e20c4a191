// DMMA.8x8x4 (mma.sync m8n8k4 f64) on B200: dependent-issue latency and throughput per SM sub-partition as a
// function of the warps per sub-partition and of the dependency structure (the fused-block chain of
// tile_pass_kernel: k <- 2 dependent DMMA, then re and im <- 2 dependent DMMA each, both depending on k).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_c(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// MODE 0: one dependent chain; 1: two chains; 2: four chains; 3: block pattern, one group at a time;
// 4: block pattern, two groups interleaved by hand; 5: block pattern, four groups interleaved
template <int MODE>
__global__ void __launch_bounds__(1024) k(double* out, int iters, double x) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = i * 1e-3 + threadIdx.x * 1e-9;
  const double a = x, b = 1.0 - x;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 24; i++) dmma(c[0], c[1], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 12; i++) { dmma(c[0], c[1], a, b); dmma(c[2], c[3], a, b); }
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 6; i++) { dmma(c[0], c[1], a, b); dmma(c[2], c[3], a, b); dmma(c[4], c[5], a, b); dmma(c[6], c[7], a, b); }
    } else {
      constexpr int G = MODE == 3 ? 1 : MODE == 4 ? 2 : 4;
#pragma unroll
      for (int rep = 0; rep < 4 / G; rep++) {
        double k0[G], k1[G], r0[G], r1[G], i0[G], i1[G];
#pragma unroll
        for (int g = 0; g < G; g++) { k0[g] = 0; k1[g] = 0; dmma(k0[g], k1[g], c[g], b); }
#pragma unroll
        for (int g = 0; g < G; g++) dmma(k0[g], k1[g], c[4 + g], a);
#pragma unroll
        for (int g = 0; g < G; g++) { dmma_c(r0[g], r1[g], c[g], a, k0[g], k1[g]); dmma_c(i0[g], i1[g], c[4 + g], b, k0[g], k1[g]); }
#pragma unroll
        for (int g = 0; g < G; g++) { dmma(r0[g], r1[g], c[8 + g], b); dmma(i0[g], i1[g], c[12 + g], a); }
#pragma unroll
        for (int g = 0; g < G; g++) { c[g] = r0[g] * 1e-3; c[4 + g] = i1[g] * 1e-3; c[8 + g] = r1[g]; c[12 + g] = i0[g]; }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  for (int wps : {1, 2, 4, 6, 8}) {  // warps per sub-partition
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, wps * 128>>>(out, 100, 0.999);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148, wps * 128>>>(out, iters, 0.999);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dmmas = (double)iters * 24 * wps;  // per sub-partition
    printf("%-34s warps/SMSP %d: %.3f ms  %.1f ns/DMMA/SMSP = %.1f clk @1.9GHz  %.1f TF/s\n", name, wps, ms,
           ms * 1e6 / dmmas, ms * 1e6 / dmmas * 1.9, dmmas * 4 * 148 * 512 / (ms * 1e-3) / 1e12);
  }
  cudaFree(out);
}

int main() {
  run<0>("1 dependent chain");
  run<1>("2 chains");
  run<2>("4 chains");
  run<3>("block pattern, 1 group");
  run<4>("block pattern, 2 groups interleaved");
  run<5>("block pattern, 4 groups interleaved");
  return 0;
}
