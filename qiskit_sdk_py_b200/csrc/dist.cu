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

// General form: the sub-block whose local bits `bits[0..k)` (ascending) equal the pattern `vals`.
struct SubDesc {
  int32_t k;
  int32_t bits[8];   // ascending local bit positions
  uint64_t value;    // the pattern already deposited at those positions
};

__device__ __forceinline__ u64 sub_index(const SubDesc& d, u64 h) {
  for (int i = 0; i < d.k; i++) {
    const int p = d.bits[i];
    h = ((h >> p) << (p + 1)) | (h & ((1ull << p) - 1ull));
  }
  return h | d.value;
}

__global__ void __launch_bounds__(kThreads)
pack_bits_kernel(const double2* __restrict__ sv, const __grid_constant__ SubDesc d, double2* __restrict__ buf,
                 u64 begin, u64 count) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) buf[j] = sv[sub_index(d, begin + j)];
}

__global__ void __launch_bounds__(kThreads)
unpack_bits_kernel(double2* __restrict__ sv, const __grid_constant__ SubDesc d, const double2* __restrict__ buf,
                   u64 begin, u64 count) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) sv[sub_index(d, begin + j)] = buf[j];
}

// dst[base | deposit(s)] = src[s]: the marginal of a shard over its local measured bits goes to the bins its
// rank bits select in the marginal of the whole register (the others stay zero; the shards are then all-reduced)
struct ScatterDesc {
  int32_t m;
  int32_t dest_bit[40];
  uint64_t base;
};

__global__ void __launch_bounds__(kThreads)
scatter_bits_kernel(const double* __restrict__ src, const __grid_constant__ ScatterDesc d, double* __restrict__ dst) {
  const u64 bins = 1ull << d.m;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 sidx = (u64)blockIdx.x * blockDim.x + threadIdx.x; sidx < bins; sidx += stride) {
    u64 t = d.base;
    for (int i = 0; i < d.m; i++) t |= ((sidx >> i) & 1ull) << d.dest_bit[i];
    dst[t] = src[sidx];
  }
}

static int make_sub(int n_local, const int32_t* local_qubits, int k, uint32_t pattern, SubDesc* d) {
  if (k < 1 || k > 8 || k >= n_local || !local_qubits) return -1;
  int order[8];
  for (int i = 0; i < k; i++) order[i] = i;
  for (int i = 0; i < k; i++)
    for (int j = i + 1; j < k; j++)
      if (local_qubits[order[j]] < local_qubits[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  d->k = k;
  d->value = 0;
  for (int i = 0; i < k; i++) {
    const int q = local_qubits[order[i]];
    if (q < 0 || q >= n_local || (i > 0 && q == d->bits[i - 1])) return -1;
    d->bits[i] = q;
    d->value |= (uint64_t)((pattern >> order[i]) & 1u) << q;
  }
  return 0;
}

static int grid_for(u64 n) {
  u64 blocks = (n + kThreads - 1) / kThreads;
  const u64 cap = (u64)sm_count() * 16;
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

int qsv_scatter_bits(const double* dev_src, int m, const int32_t* dest_bit, uint64_t base, double* dev_dst,
                     int log2_dst, void* stream) {
  if (!dev_src || !dev_dst || m < 0 || m > 40 || log2_dst < m || log2_dst > 40 || (m > 0 && !dest_bit) ||
      base >= (1ull << log2_dst))
    return fail(QSV_EINVAL, "qsv_scatter_bits: bad arguments");
  ScatterDesc d;
  d.m = m;
  d.base = base;
  uint64_t seen = 0;
  for (int i = 0; i < m; i++) {
    const int b = dest_bit[i];
    if (b < 0 || b >= log2_dst || ((seen >> b) & 1ull) || ((base >> b) & 1ull))
      return fail(QSV_EINVAL, "qsv_scatter_bits: dest bits must be distinct, below log2_dst and clear in base");
    seen |= 1ull << b;
    d.dest_bit[i] = b;
  }
  cudaStream_t s = (cudaStream_t)stream;
  QSV_CUDA(cudaMemsetAsync(dev_dst, 0, sizeof(double) << log2_dst, s));
  scatter_bits_kernel<<<grid_for(1ull << m), kThreads, 0, s>>>(dev_src, d, dev_dst);
  QSV_LAUNCHED();
  return 0;
}

int qsv_pack_bits(const double* sv, int n_local, const int32_t* local_qubits, int k, uint32_t pattern, double* dev_buf,
                  size_t begin, size_t count, void* stream) {
  SubDesc d;
  if (!sv || !dev_buf || n_local < 2 || make_sub(n_local, local_qubits, k, pattern, &d) ||
      begin + count > ((size_t)1 << (n_local - k)))
    return fail(QSV_EINVAL, "qsv_pack_bits: bad arguments");
  if (count == 0) return 0;
  pack_bits_kernel<<<grid_for(count), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(sv), d, reinterpret_cast<double2*>(dev_buf), begin, count);
  QSV_LAUNCHED();
  return 0;
}

int qsv_unpack_bits(double* sv, int n_local, const int32_t* local_qubits, int k, uint32_t pattern, const double* dev_buf,
                    size_t begin, size_t count, void* stream) {
  SubDesc d;
  if (!sv || !dev_buf || n_local < 2 || make_sub(n_local, local_qubits, k, pattern, &d) ||
      begin + count > ((size_t)1 << (n_local - k)))
    return fail(QSV_EINVAL, "qsv_unpack_bits: bad arguments");
  if (count == 0) return 0;
  unpack_bits_kernel<<<grid_for(count), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<double2*>(sv), d, reinterpret_cast<const double2*>(dev_buf), begin, count);
  QSV_LAUNCHED();
  return 0;
}

}  // extern "C"
