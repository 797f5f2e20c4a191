// Does shared-memory traffic (LDS.128 / STS.128 in the proportion of tile_pass_kernel's fused-block path: per 64
// amplitudes 2 LDS.128 + 2 DADD + 6 DMMA + 2 STS.128 per lane) slow the FP64 tensor pipe down on B200?  24 warps per
// SM (6 per sub-partition) as in the pass kernel.  MODE 0: DMMA only; 1: + LDS/STS (conflict-free, 16-byte stride);
// 2: + LDS/STS with the data dependency of the real kernel (operands of the DMMA come from the loads, stores take the
// results); 3: mode 2 + a __syncwarp and 8 more LDS.64 (B fragments) every 4 steps (= one op on a 256-amplitude
// sub-cube); 4: LDS/STS only (no DMMA).
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

template <int MODE>
__global__ void __launch_bounds__(768) k(double* out, int iters, double x) {
  extern __shared__ double2 sm[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2* mine = sm + wid * 512;  // 8 KiB per warp: 4 steps x (2 x 32 lanes) double2
  for (int i = lane; i < 512; i += 32) mine[i] = make_double2(1e-3 * i, 1e-3);
  __syncthreads();
  double bq[6];
#pragma unroll
  for (int q = 0; q < 6; q++) bq[q] = x + q * 1e-3;
  double acc = 0.0;
  for (int it = 0; it < iters; it++) {
    if (MODE == 3) {
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 6; q++) bq[q] = reinterpret_cast<const double*>(sm + ((wid + 1) % 24) * 512)[q * 32 + lane];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double2 u0 = make_double2(x, x * 0.5), u1 = make_double2(x * 0.25, x);
      if (MODE >= 1) {
        u0 = mine[j * 64 + lane];
        u1 = mine[j * 64 + 32 + lane];
      }
      if (MODE == 1) { acc += u0.x + u1.y; u0 = make_double2(x, x * 0.5); u1 = make_double2(x * 0.25, x); }
      double k0 = 0.0, k1 = 0.0, r0 = 0, r1 = 0, i0 = 0, i1 = 0;
      if (MODE != 4) {
        dmma(k0, k1, u0.x + u0.y, bq[0]);
        dmma(k0, k1, u1.x + u1.y, bq[1]);
        dmma_c(r0, r1, u0.y, bq[2], k0, k1);
        dmma_c(i0, i1, u0.x, bq[4], k0, k1);
        dmma(r0, r1, u1.y, bq[3]);
        dmma(i0, i1, u1.x, bq[5]);
      } else {
        r0 = u0.x; i0 = u0.y; r1 = u1.x; i1 = u1.y;
      }
      if (MODE == 1) {
        acc += r0 + i1;
        mine[j * 64 + lane] = make_double2(x, acc * 1e-30);
        mine[j * 64 + 32 + lane] = make_double2(acc * 1e-30, x);
      } else if (MODE >= 2) {
        mine[j * 64 + lane] = make_double2(r0 * 1e-3, i0 * 1e-3);
        mine[j * 64 + 32 + lane] = make_double2(r1 * 1e-3, i1 * 1e-3);
      } else {
        acc += r0 + i0 + r1 + i1;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + mine[lane].x;
}

template <int MODE>
void run(const char* name) {
  double* out;
  cudaMalloc(&out, sizeof(double) * 148 * 768);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 8192);
  const int iters = 3000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 768, 24 * 8192>>>(out, 100, 0.999);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148, 768, 24 * 8192>>>(out, iters, 0.999);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double steps = (double)iters * 4 * 6;  // DMMA steps per sub-partition
  printf("%-44s %.3f ms  %.1f clk/step/SMSP @1.9GHz (6 DMMA = 96 clk at the pipe's peak; smem 4 x 4 wavefronts x 24 warps / 4 = 96 clk)  %s\n",
         name, ms, ms * 1e6 / steps * 1.9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("DMMA only (block pattern)");
  run<1>("DMMA + independent LDS.128/STS.128");
  run<2>("DMMA fed by LDS, results to STS");
  run<3>("... + syncwarp and fragment loads per op");
  run<4>("LDS/STS only");
  return 0;
}
