// Shard-exchange helpers for the multi-GPU layout (K12 of SURVEY 2.5): the state is split by its
// high-order "global" qubits, one contiguous slab of 2^n_local amplitudes per GPU.  Swapping a global
// qubit with local qubit L means: the half of my slab whose bit L differs from my rank's bit of that
// global qubit goes to the partner rank, and the partner's matching half comes here.  These kernels
// gather / scatter that half to and from a contiguous staging buffer; the transport itself is NCCL
// send/recv issued by the Python host through torch.distributed.  (The reference has no multi-device
// path at all: qasm_simulator.py keeps one numpy array in host memory.)
#include <cuda.h>
#include <string.h>

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

// ---- qubit swap over NVLink peer memory, in place, no staging and no NCCL on the data path -----------------
// Swapping k global qubits with the k local qubits `bits` is an all-to-all among the 2^k ranks that differ only
// in those global bits: rank beta's sub-block [local bits = lam] trades places with rank lam's sub-block [local
// bits = beta].  The peers' shards are mapped into this process (CUDA IPC, qsv_ipc_import), so ONE kernel per
// rank does the whole exchange: for every peer lam it walks HALF of the pair's sub-block (the rank with the
// smaller pattern takes the lower half of the compressed index range, its partner the upper half) and swaps
// element by element -- remote load + local load, remote store + local store, 16 bytes each, coalesced in runs
// of 2^(lowest swapped bit) amplitudes.  Every amplitude that has to move crosses NVLink exactly once in each
// direction and HBM sees exactly one read and one write per moved amplitude on each side (the staged NCCL path
// read and wrote every moved amplitude three times).  The caller brackets the kernel with a stream-ordered
// barrier among the ranks (all passes before it are complete on every rank; all swaps are complete before the
// next pass).
struct P2PDesc {
  int32_t k, n_local;
  int32_t bits[8];        // ascending local bit positions
  uint32_t beta;          // this rank's pattern of the k global bits
  uint32_t mode;          // 0 = swap; debug library only (bandwidth probes, destroy the data): 1 = push only (local
                          // load + remote store), 2 = pull only (remote load + local store)
  uint32_t perm[8];       // bit i of a pattern <-> position bits[perm[i]] ... (patterns are given in the caller's order)
  double2* peer[256];     // peer[lam]: base of the shard of the rank whose global bits read lam (beta: unused)
};

__device__ __forceinline__ u64 p2p_deposit(const P2PDesc& d, uint32_t pattern) {
  u64 v = 0;
  for (int i = 0; i < d.k; i++) v |= (u64)((pattern >> i) & 1u) << d.bits[d.perm[i]];
  return v;
}

template <int kU, bool kChunked>  // kU independent remote loads in flight per thread
__global__ void __launch_bounds__(1024)
swap_p2p_kernel(double2* __restrict__ mine, const __grid_constant__ P2PDesc d) {
  const u64 sub = 1ull << (d.n_local - d.k);  // amplitudes per sub-block
  const u64 half = sub >> 1;
  const u64 dep_beta = p2p_deposit(d, d.beta);
  // kChunked: every CTA walks ONE contiguous slice of the index range (1024 * kU amplitudes per iteration), so an SM
  // touches 1/gridDim of the pages of the two shards instead of all of them (grid-stride: consecutive CTAs take
  // consecutive 16 KiB pieces, every SM's TLB sees every page of a 128 GiB shard and of its peer mapping)
  const u64 per_iter = (u64)blockDim.x * kU;
  const u64 per_cta = kChunked ? ((half + gridDim.x - 1) / gridDim.x + per_iter - 1) / per_iter * per_iter : 0;
  const u64 lo = kChunked ? (u64)blockIdx.x * per_cta : 0;
  const u64 hi = kChunked ? (lo + per_cta < half ? lo + per_cta : half) : half;
  const u64 step = kChunked ? (u64)blockDim.x : (u64)gridDim.x * blockDim.x;  // distance between the unrolled accesses
  const u64 t0 = kChunked ? lo + threadIdx.x : (u64)blockIdx.x * blockDim.x + threadIdx.x;
  // XOR schedule: in step j every rank works on its partner beta ^ j -- a perfect matching of the 2^k ranks per step,
  // so no rank's memory or NVLink port is the target of more than one partner at a time (a common order 0, 1, 2, ...
  // has everybody hit rank 0 first: measured 280 GB/s instead of 650 on 4 GPUs)
  for (uint32_t j = 1; j < (1u << d.k); j++) {
    const uint32_t lam = d.beta ^ j;
    if (!d.peer[lam]) continue;  // a partial exchange: only the peers given
    double2* __restrict__ theirs = d.peer[lam];
    const u64 dep_lam = p2p_deposit(d, lam);
    const u64 begin = d.beta < lam ? 0 : half;
    for (u64 h0 = t0; h0 < hi; h0 += step * kU) {
      double2 a[kU], b[kU];
      u64 im[kU], ip[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const u64 h = h0 + (u64)u * step;
        u64 x = begin + (h < hi ? h : 0);
        for (int i = 0; i < d.k; i++) {
          const int p = d.bits[i];
          x = ((x >> p) << (p + 1)) | (x & ((1ull << p) - 1ull));
        }
        im[u] = x | dep_lam;   // my sub-block that leaves
        ip[u] = x | dep_beta;  // the partner's sub-block that comes here
        if (h < hi) {
#ifdef QSV_DEBUG_BUILD
          if (d.mode != 1u) b[u] = __ldcv(theirs + ip[u]);
          if (d.mode != 2u) a[u] = mine[im[u]];
#else
          b[u] = __ldcv(theirs + ip[u]);
          a[u] = mine[im[u]];
#endif
        }
      }
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const u64 h = h0 + (u64)u * step;
        if (h < hi) {
#ifdef QSV_DEBUG_BUILD
          if (d.mode != 1u) mine[im[u]] = b[u];
          if (d.mode != 2u) theirs[ip[u]] = a[u];
#else
          mine[im[u]] = b[u];
          theirs[ip[u]] = a[u];
#endif
        }
      }
    }
  }
}

bool g_p2p_chunked = false;  // qsv_debug_set_flags bit 14 (tools/p2p_swap_probe.py)
uint32_t g_p2p_mode = 0;     // ... bits 15-16, debug library only: P2PDesc::mode
int g_p2p_unroll = 4;  // tools/p2p_swap_probe.py sweeps it through qsv_debug_set_flags (bits 12-13)

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

// ---- SWAP gates on disjoint qubit pairs: one in-place permutation pass over HBM -----------------------------------
struct SwapPairsDesc {
  int32_t n_pairs;
  int32_t a[QSV_MAX_SWAP_PAIRS], b[QSV_MAX_SWAP_PAIRS];
};

// pi(i) = i with the bits of every pair exchanged, an involution: the thread of the smaller index of {i, pi(i)} swaps the
// two amplitudes.  Whether a thread is that one is decided by the pair with the highest bit that differs; with all pair
// bits >= 5 a warp's 32 consecutive indices decide alike, so active warps move 512 contiguous bytes on both sides and the
// others leave at once.  kU independent index pairs per thread keep enough loads in flight.  pi is a permutation of the
// index BITS, i.e. the OR of the images of the index's bytes: five 256-entry tables in shared memory (built per CTA from
// the pairs) replace the per-pair 64-bit shift arithmetic, which made the first version ALU-bound (83 % ALU pipe, 49 %
// DRAM: profiles/r2_swap_pairs_kernel_ncu.txt).
template <int kU>
__global__ void __launch_bounds__(kThreads)
swap_pairs_kernel(double2* __restrict__ sv, const __grid_constant__ SwapPairsDesc d, u64 n_amps) {
  __shared__ uint8_t dest[40];      // dest[s] = where index bit s goes
  __shared__ u64 tab[5][256];       // tab[c][v] = image of byte c of the index holding the value v
  if (threadIdx.x < 40) {
    int to = threadIdx.x;
    for (int q = 0; q < d.n_pairs; q++) {
      if (d.a[q] == (int)threadIdx.x) to = d.b[q];
      if (d.b[q] == (int)threadIdx.x) to = d.a[q];
    }
    dest[threadIdx.x] = (uint8_t)to;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 5 * 256; e += blockDim.x) {
    const int c = e >> 8, v = e & 255;
    u64 img = 0;
    for (int j = 0; j < 8; j++)
      if ((v >> j) & 1) img |= 1ull << dest[8 * c + j];
    tab[c][v] = img;
  }
  __syncthreads();
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_amps; i0 += stride * kU) {
    u64 ii[kU], pp[kU];
    double2 x[kU], y[kU];
    bool act[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) {
      const u64 i = i0 + (u64)u * stride;
      const uint32_t lo = (uint32_t)i, hi = (uint32_t)(i >> 32);
      const u64 p = tab[0][lo & 255u] | tab[1][(lo >> 8) & 255u] | tab[2][(lo >> 16) & 255u] | tab[3][lo >> 24] |
                    tab[4][hi & 255u];
      ii[u] = i;
      pp[u] = p;
      act[u] = i < n_amps && p > i;
      if (act[u]) {
        x[u] = sv[i];
        y[u] = sv[p];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; u++)
      if (act[u]) {
        sv[ii[u]] = y[u];
        sv[pp[u]] = x[u];
      }
  }
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

int qsv_ipc_export(const void* dev_ptr, void* handle_out, uint64_t* offset_out) {
  if (!dev_ptr || !handle_out || !offset_out) return fail(QSV_EINVAL, "qsv_ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  // the handle names the whole allocation: find its base (the caller's pointer may sit inside a pooled block)
  cudaPointerAttributes attr;
  QSV_CUDA(cudaPointerGetAttributes(&attr, dev_ptr));
  if (attr.type != cudaMemoryTypeDevice) return fail(QSV_EINVAL, "qsv_ipc_export: not a device pointer");
  void* base = nullptr;
  size_t size = 0;
  // driver entry point fetched at run time: the library must load (host-only entry points, CPU tests) on machines
  // without libcuda
  typedef CUresult (*get_range_t)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  QSV_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(QSV_EINVAL, "qsv_ipc_export: cuMemGetAddressRange unavailable");
  if (reinterpret_cast<get_range_t>(fn)(reinterpret_cast<CUdeviceptr*>(&base), &size, (CUdeviceptr)(uintptr_t)dev_ptr) !=
      CUDA_SUCCESS)
    return fail(QSV_EINVAL, "qsv_ipc_export: cuMemGetAddressRange failed");
  QSV_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), base));
  *offset_out = (uint64_t)((const char*)dev_ptr - (const char*)base);
  return 0;
}

int qsv_ipc_import(const void* handle, uint64_t offset, void** peer_ptr_out) {
  if (!handle || !peer_ptr_out) return fail(QSV_EINVAL, "qsv_ipc_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  QSV_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr_out = (char*)base + offset;
  return 0;
}

int qsv_ipc_release(void* peer_ptr, uint64_t offset) {
  if (!peer_ptr) return 0;
  QSV_CUDA(cudaIpcCloseMemHandle((char*)peer_ptr - offset));
  return 0;
}

int qsv_swap_bits_p2p(double* sv, int n_local, const int32_t* local_qubits, int k, uint32_t my_pattern,
                      double* const* peer_sv, int max_ctas, void* stream) {
  if (!sv || !local_qubits || !peer_sv || k < 1 || k > 8 || k >= n_local || my_pattern >= (1u << k))
    return fail(QSV_EINVAL, "qsv_swap_bits_p2p: bad arguments");
  P2PDesc d;
  memset(&d, 0, sizeof(d));
  d.k = k;
  d.n_local = n_local;
  d.beta = my_pattern;
  d.mode = g_p2p_mode;
  int order[8];
  for (int i = 0; i < k; i++) order[i] = i;
  for (int i = 0; i < k; i++)
    for (int j = i + 1; j < k; j++)
      if (local_qubits[order[j]] < local_qubits[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  for (int i = 0; i < k; i++) {
    const int q = local_qubits[order[i]];
    if (q < 0 || q >= n_local || (i > 0 && q == d.bits[i - 1])) return fail(QSV_EINVAL, "qsv_swap_bits_p2p: bad local qubits");
    d.bits[i] = q;
    d.perm[order[i]] = (uint32_t)i;  // pattern bit order[i] sits at the i-th smallest position
  }
  for (uint32_t lam = 0; lam < (1u << k); lam++)  // a null entry = that pair is not exchanged by this call
    d.peer[lam] = lam == my_pattern ? nullptr : reinterpret_cast<double2*>(peer_sv[lam]);
  const u64 half = (1ull << (n_local - k)) >> 1;
  if (half == 0) return fail(QSV_ERANGE, "qsv_swap_bits_p2p: sub-block too small");
  int ctas = max_ctas > 0 ? max_ctas : 2 * sm_count();
  const u64 need = (half + 1023) / 1024;
  if ((u64)ctas > need) ctas = (int)need;
#define QSV_P2P_LAUNCH(U)                                                                                         \
  if (g_p2p_chunked)                                                                                              \
    swap_p2p_kernel<U, true><<<ctas, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<double2*>(sv), d);        \
  else                                                                                                            \
    swap_p2p_kernel<U, false><<<ctas, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<double2*>(sv), d)
  switch (g_p2p_unroll) {
    case 1: QSV_P2P_LAUNCH(1); break;
    case 2: QSV_P2P_LAUNCH(2); break;
    case 8: QSV_P2P_LAUNCH(8); break;
    default: QSV_P2P_LAUNCH(4); break;
  }
#undef QSV_P2P_LAUNCH
  QSV_LAUNCHED();
  return 0;
}

int qsv_swap_qubit_pairs(double* sv, int n_qubits, const int32_t* pairs, int n_pairs, void* stream) {
  if (!sv || n_qubits < 1 || n_qubits > 40 || n_pairs < 0 || n_pairs > QSV_MAX_SWAP_PAIRS || (n_pairs > 0 && !pairs))
    return fail(QSV_EINVAL, "qsv_swap_qubit_pairs: bad arguments");
  if (n_pairs == 0) return 0;
  SwapPairsDesc d;
  memset(&d, 0, sizeof(d));
  d.n_pairs = n_pairs;
  uint64_t seen = 0;
  for (int i = 0; i < 2 * n_pairs; i++) {
    const int q = pairs[i];
    if (q < 0 || q >= n_qubits || ((seen >> q) & 1ull))
      return fail(QSV_EINVAL, "qsv_swap_qubit_pairs: the pairs must be disjoint qubits of the register");
    seen |= 1ull << q;
    (i & 1 ? d.b : d.a)[i >> 1] = q;
  }
  const u64 n_amps = 1ull << n_qubits;
  const u64 blocks = (n_amps + (u64)kThreads * 4 - 1) / ((u64)kThreads * 4);
  const u64 cap = (u64)sm_count() * 8;
  swap_pairs_kernel<4><<<(unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)), kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<double2*>(sv), d, n_amps);
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
