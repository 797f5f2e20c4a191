// Shared helpers for libqsv (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "qsv.h"

namespace qsv {

extern thread_local char g_err[512];
extern unsigned long long g_launches;
extern int g_p2p_unroll;
extern bool g_p2p_chunked;
extern uint32_t g_p2p_mode;

inline int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

#define QSV_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      snprintf(qsv::g_err, sizeof(qsv::g_err), "%s failed: %s (%s:%d)", #expr,               \
               cudaGetErrorString(_e), __FILE__, __LINE__);                                  \
      return (int)_e;                                                                        \
    }                                                                                        \
  } while (0)

#define QSV_LAUNCHED()                                                                       \
  do {                                                                                       \
    __sync_fetch_and_add(&qsv::g_launches, 1ULL);                                            \
    QSV_CUDA(cudaGetLastError());                                                            \
  } while (0)

constexpr int kThreads = 256;

// SM count of the current device (148 on a B200), queried once per device; grids are sized from it.
inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

typedef unsigned long long u64;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc += a * b (complex), 4 DFMA
__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// dense_k.cu: k-qubit dense gate (4 <= k <= 7, n >= 11) as a complex GEMM on the FP64 tensor path;
// dev_scratch holds dense_k_scratch_doubles(k) doubles
int launch_dense_k_dmma(double* sv, int n, const int32_t* qubits, int k, const double* host_matrix, double* dev_scratch,
                        cudaStream_t s);
size_t dense_k_scratch_doubles(int k);

}  // namespace qsv
