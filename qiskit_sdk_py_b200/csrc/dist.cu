// Shard-exchange helpers for the multi-GPU layout (K12 of SURVEY 2.5): the state is split by its
// high-order "global" qubits, one contiguous slab of 2^n_local amplitudes per GPU.  Swapping a global
// qubit with local qubit L means: the half of my slab whose bit L differs from my rank's bit of that
// global qubit goes to the partner rank, and the partner's matching half comes here.  These kernels
// gather / scatter that half to and from a contiguous staging buffer; the transport itself is NCCL
// send/recv issued by the Python host through torch.distributed.  (The reference has no multi-device
// path at all: qasm_simulator.py keeps one numpy array in host memory.)
#include "common.cuh"

namespace qsv {

__global__ void __launch_bounds__(kThreads)
pack_half_kernel(const double2* __restrict__ sv, int L, u64 bitv, double2* __restrict__ buf, u64 begin, u64 count) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 low_mask = (1ull << L) - 1ull;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
    const u64 h = begin + j;
    const u64 idx = ((h >> L) << (L + 1)) | (bitv << L) | (h & low_mask);
    buf[j] = sv[idx];
  }
}

__global__ void __launch_bounds__(kThreads)
unpack_half_kernel(double2* __restrict__ sv, int L, u64 bitv, const double2* __restrict__ buf, u64 begin, u64 count) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 low_mask = (1ull << L) - 1ull;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
    const u64 h = begin + j;
    const u64 idx = ((h >> L) << (L + 1)) | (bitv << L) | (h & low_mask);
    sv[idx] = buf[j];
  }
}

static int grid_for(u64 n) {
  u64 blocks = (n + kThreads - 1) / kThreads;
  const u64 cap = (u64)kNumSMs * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace qsv

using namespace qsv;

extern "C" {

int qsv_pack_half(const double* sv, int n_local, int L, int bitv, double* dev_buf, size_t begin, size_t count,
                  void* stream) {
  if (!sv || !dev_buf || n_local < 1 || L < 0 || L >= n_local || (bitv != 0 && bitv != 1) ||
      begin + count > ((size_t)1 << (n_local - 1)))
    return fail(QSV_EINVAL, "qsv_pack_half: bad arguments");
  if (count == 0) return 0;
  pack_half_kernel<<<grid_for(count), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(sv), L, (u64)bitv, reinterpret_cast<double2*>(dev_buf), begin, count);
  QSV_LAUNCHED();
  return 0;
}

int qsv_unpack_half(double* sv, int n_local, int L, int bitv, const double* dev_buf, size_t begin, size_t count,
                    void* stream) {
  if (!sv || !dev_buf || n_local < 1 || L < 0 || L >= n_local || (bitv != 0 && bitv != 1) ||
      begin + count > ((size_t)1 << (n_local - 1)))
    return fail(QSV_EINVAL, "qsv_unpack_half: bad arguments");
  if (count == 0) return 0;
  unpack_half_kernel<<<grid_for(count), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<double2*>(sv), L, (u64)bitv, reinterpret_cast<const double2*>(dev_buf), begin, count);
  QSV_LAUNCHED();
  return 0;
}

}  // extern "C"
