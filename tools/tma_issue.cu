// How fast can ONE lane issue 256-byte TMA bulk copies (cp.async.bulk, SASS UBLKCP) on B200, and what does the
// issuing warp pay per copy?  Shapes: W issuing warps per SM (1 = a dedicated producer warp, 3 = one per warp
// group, 8 / 24 = every compute warp issues its own pieces as the round-1 kernel did).  Each issuing warp moves
// batches of 128 pieces (one 11-qubit tile) between a large global buffer and its own shared-memory buffer.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// mode 0: loads only (wait for each batch), 1: stores only (wait_group.read per batch), 2: store batch, wait read, load batch
// (the producer's steady state), 3: loads, two batches in flight
__global__ void __launch_bounds__(1024) k(const char* __restrict__ src, char* __restrict__ dst, size_t span, int batches,
                                          int mode, unsigned long long* out) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t bar[32][2];
  const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  char* buf = smem + (size_t)wid * 2 * 32768 / (nw > 3 ? nw / 3 : 1);  // <= 3 warps: 64 KiB each; more: share
  const int pieces = nw > 3 ? 128 * 3 / nw : 128;                      // keep 3 tiles per SM in flight in every shape
  if (threadIdx.x == 0)
    for (int i = 0; i < 64; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar[0][0] + i)));
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  const size_t warp_id = (size_t)blockIdx.x * nw + wid;
  unsigned long long t_issue = 0, t_total = 0;
  const long long t_start = clock64();
  uint32_t par[2] = {0, 0};
  for (int b = 0; b < batches; b++) {
    // a tile-like address pattern: 128 pieces of 256 B, 64 KiB apart, tile base walks 256 B
    const size_t base = ((warp_id * batches + b) * 256) % (span - (size_t)128 * 65536);
    const int s = (mode == 3) ? (b & 1) : 0;
    if (elect_one()) {
      const long long t0 = clock64();
      if (mode == 1 || mode == 2) {
        for (int p = 0; p < pieces; p++)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;\n" ::"l"(dst + base + (size_t)p * 65536),
                       "r"(smem_u32(buf + p * 256)) : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
      long long t1 = clock64();
      if (mode == 1 || mode == 2) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      long long t2 = clock64();
      if (mode != 1) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar[wid][s])), "r"(pieces * 256) : "memory");
        for (int p = 0; p < pieces; p++)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];\n" ::"r"(
                           smem_u32(buf + s * 32768 / (nw > 3 ? nw / 3 : 1) + p * 256)), "l"(src + base + (size_t)p * 65536),
                       "r"(smem_u32(&bar[wid][s])) : "memory");
      }
      const long long t3 = clock64();
      t_issue += (unsigned long long)((t1 - t0) + (t3 - t2));
    }
    __syncwarp();
    if (mode == 0 || mode == 2) { mbar_wait(&bar[wid][0], par[0]); par[0] ^= 1; }
    if (mode == 3 && b > 0) { mbar_wait(&bar[wid][(b - 1) & 1], par[(b - 1) & 1]); par[(b - 1) & 1] ^= 1; }
  }
  if (mode == 3) mbar_wait(&bar[wid][(batches - 1) & 1], par[(batches - 1) & 1]);
  t_total = (unsigned long long)(clock64() - t_start);
  if ((threadIdx.x & 31) == 0) {
    out[warp_id * 2] = t_issue;
    out[warp_id * 2 + 1] = t_total;
  }
}

int main() {
  const size_t span = (size_t)8 << 30;
  char *src, *dst;
  unsigned long long* out;
  cudaMalloc(&src, span);
  cudaMalloc(&dst, span);
  cudaMalloc(&out, sizeof(unsigned long long) * 2 * 148 * 32);
  cudaMemset(src, 1, span);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[4] = {"load,wait", "store,wait_read", "store,wait_read,load,wait", "load x2 in flight"};
  for (int mode = 0; mode < 4; mode++)
    for (int nw : {1, 3, 6, 24}) {
      const int batches = 200;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      k<<<148, nw * 32, 200 * 1024>>>(src, dst, span, 10, mode, out);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      k<<<148, nw * 32, 200 * 1024>>>(src, dst, span, batches, mode, out);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      static unsigned long long h[2 * 148 * 32];
      cudaMemcpy(h, out, sizeof(unsigned long long) * 2 * 148 * nw, cudaMemcpyDeviceToHost);
      double iss = 0, tot = 0;
      for (int i = 0; i < 148 * nw; i++) { iss += h[2 * i]; tot += h[2 * i + 1]; }
      const int pieces = nw > 3 ? 128 * 3 / nw : 128;
      const double copies = (double)batches * pieces * ((mode == 2) ? 2 : 1);
      const double bytes = 148.0 * nw * copies * 256;
      printf("mode %-26s warps/SM %2d: %.3f ms  %.0f GB/s  issue %.1f clk/copy (issuing lane)  %.0f clk/batch/warp  %s\n",
             names[mode], nw, ms, bytes / ms / 1e6, iss / (148.0 * nw) / copies, tot / (148.0 * nw) / batches,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
