// k-qubit dense gates, 4 <= k <= 7, on the FP64 tensor path (north_star item (2); the `unitary` instruction with k
// qubits, /root/reference/qiskit/providers/basicaer/qasm_simulator.py:543-546 -> _add_unitary :145-161, one einsum
// over the whole vector per gate in the reference).
//
// A CTA stages a tile of 2^11 amplitudes -- the k gate qubits plus the lowest other qubits -- in shared memory, i.e.
// G = 2^(11-k) groups of 2^k amplitudes, and computes OUT = M * IN for all groups as a complex GEMM in the
// 3-multiplication form (M = P + iQ, amplitudes x + iy):
//     K = P (x + y),   R = -(P + Q) y,   I = (Q - P) x,     re = K + R,   im = K + I
// three real 2^k x 2^k products on DMMA.8x8x4 (mma.sync m8n8k4 f64): the A operand is 8 groups x 4 amplitudes of the
// tile (one LDS.128 per lane gives x, y and x + y), the B operand a 4 x 8 slice of the matrix (pre-arranged on the
// host in fragment order, read through the read-only path, re-used for 2-4 group-steps), and a lane ends up with two
// complex outputs of one group.  The result goes to a second tile buffer (every output needs all 2^k inputs of its
// group) and from there back to HBM.  Per amplitude 3 * 2^k / 256 DMMA: at k = 5 a pass needs about as long on the
// FP64 pipe as on HBM; k = 6, 7 are FP64-bound, k = 4 is HBM-bound.  Both tile buffers are padded per group (64 B on
// the load side, 16 B on the store side) so that the quarter-warp accesses are bank-conflict free.
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace qsv {

constexpr int kDkTile = 11;      // tile qubits
constexpr int kDkThreads = 256;  // 8 warps, 32 (group-step, output-block) work units per tile

struct DenseKDmmaDesc {
  int32_t n, k;
  uint32_t n_tiles;
  uint8_t tbs[16];   // tile bits ascending (physical)
  uint8_t loc[16];   // local index bit of the j-th smallest tile bit: [0, k) = matrix index bits, [k, 11) = group bits
};

__device__ __forceinline__ void dk_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int K>
__global__ void __launch_bounds__(kDkThreads)
dense_k_dmma_kernel(double2* __restrict__ sv, const double* __restrict__ frag, const __grid_constant__ DenseKDmmaDesc d) {
  constexpr int kAmps = 1 << K;             // amplitudes per group
  constexpr int kGroups = 1 << (kDkTile - K);
  constexpr int kInStride = kAmps + 4;      // slots of 16 B; +64 B: consecutive groups land in the other half of the banks
  constexpr int kOutStride = kAmps + 1;     // +16 B
  constexpr int kObs = kAmps / 8;           // output blocks of 8 amplitudes
  constexpr int kKs = kAmps / 4;            // k-slices of 4 amplitudes
  constexpr int kObCount = kObs > 8 ? kObs / 8 : 1;  // output blocks per warp
  constexpr int kObWarps = kObs > 8 ? 8 : kObs;      // warps that differ in their output block
  constexpr int kSteps = 4 / kObCount;               // group-steps (8 groups each) per warp, sharing the B fragments
  extern __shared__ double2 dk_smem[];
  double2* tin = dk_smem;
  double2* tout = dk_smem + kGroups * kInStride;
  __shared__ unsigned long long g_hi[8];
  __shared__ uint32_t s_hi[8];  // in-slot (low 16 bits) and out-slot (high 16) contributions of element bits 8..10

  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  // element e of the tile in ascending PHYSICAL order (coalesced: the lowest tile bits are physical bits 0..): its
  // offset in HBM and its slots in the two buffers are sums of per-bit terms; thread t owns e = t + 256 i
  auto terms = [&](uint32_t e, int lo, int hi, unsigned long long& g, uint32_t& si, uint32_t& so) {
    g = 0; si = 0; so = 0;
    for (int j = lo; j < hi; j++)
      if ((e >> j) & 1u) {
        g |= 1ull << d.tbs[j];
        const int l = d.loc[j];
        if (l < K) { si += 1u << l; so += 1u << l; }
        else { si += (uint32_t)kInStride << (l - K); so += (uint32_t)kOutStride << (l - K); }
      }
  };
  unsigned long long g_lo;
  uint32_t si_lo, so_lo;
  terms(tid, 0, 8, g_lo, si_lo, so_lo);
  if (tid < 8) {
    unsigned long long g;
    uint32_t si, so;
    terms(tid << 8, 8, kDkTile, g, si, so);
    g_hi[tid] = g;
    s_hi[tid] = si | (so << 16);
  }
  __syncthreads();

  const uint32_t g8 = lane >> 2, c = lane & 3u;
  const uint32_t ob0 = (wid % kObWarps) * kObCount, step0 = (wid / kObWarps) * kSteps;

  for (unsigned long long t = blockIdx.x; t < d.n_tiles; t += gridDim.x) {
    unsigned long long base = t;
    for (int j = 0; j < kDkTile; j++) {
      const int p = d.tbs[j];
      base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
    }
    // ---- one HBM read of the tile ----
#pragma unroll
    for (int i = 0; i < 8; i++) tin[si_lo + (s_hi[i] & 0xffffu)] = sv[base + g_lo + g_hi[i]];
    __syncthreads();
    // ---- OUT = M * IN ----
#pragma unroll 1
    for (int obi = 0; obi < kObCount; obi++) {
      const uint32_t ob = ob0 + obi;
      double kk0[kSteps], kk1[kSteps], rr0[kSteps], rr1[kSteps], ii0[kSteps], ii1[kSteps];
#pragma unroll
      for (int s = 0; s < kSteps; s++) kk0[s] = kk1[s] = rr0[s] = rr1[s] = ii0[s] = ii1[s] = 0.0;
      const double* fr = frag + ((size_t)ob * kKs * 3) * 32 + lane;
#pragma unroll 4
      for (int ks = 0; ks < kKs; ks++) {
        const double bp = __ldg(fr + (ks * 3 + 0) * 32), bm = __ldg(fr + (ks * 3 + 1) * 32),
                     bq = __ldg(fr + (ks * 3 + 2) * 32);
#pragma unroll
        for (int s = 0; s < kSteps; s++) {
          const double2 u = tin[((step0 + s) * 8 + g8) * kInStride + 4 * ks + c];
          dk_dmma(kk0[s], kk1[s], u.x + u.y, bp);
          dk_dmma(rr0[s], rr1[s], u.y, bm);
          dk_dmma(ii0[s], ii1[s], u.x, bq);
        }
      }
#pragma unroll
      for (int s = 0; s < kSteps; s++) {
        double2* o = tout + ((step0 + s) * 8 + g8) * kOutStride + 8 * ob + 2 * c;
        o[0] = make_double2(kk0[s] + rr0[s], kk0[s] + ii0[s]);
        o[1] = make_double2(kk1[s] + rr1[s], kk1[s] + ii1[s]);
      }
    }
    __syncthreads();
    // ---- one HBM write of the tile ----
#pragma unroll
    for (int i = 0; i < 8; i++) sv[base + g_lo + g_hi[i]] = tout[so_lo + (s_hi[i] >> 16)];
    // the next tile's loads go to `tin`, which every warp has finished reading (barrier above); its stores to
    // `tout` come after the next barrier, i.e. after this write-back
  }
}

template <int K>
static int launch_k(double* sv, const double* dev_frag, const DenseKDmmaDesc& d, cudaStream_t s) {
  constexpr size_t smem = sizeof(double2) * ((size_t)(1 << (kDkTile - K)) * ((1 << K) + 4) +
                                            (size_t)(1 << (kDkTile - K)) * ((1 << K) + 1));
  static bool attr_done[64] = {false};
  int dev = 0;
  QSV_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    QSV_CUDA(cudaFuncSetAttribute(dense_k_dmma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev] = true;
  }
  int per_sm = 1;
  QSV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_k_dmma_kernel<K>, kDkThreads, smem));
  if (per_sm < 1) return fail(QSV_ERANGE, "qsv_apply_matrix: dense-k kernel does not fit in shared memory");
  const unsigned grid = std::min<unsigned>(d.n_tiles, (unsigned)per_sm * (unsigned)sm_count());
  dense_k_dmma_kernel<K><<<grid, kDkThreads, smem, s>>>(reinterpret_cast<double2*>(sv), dev_frag, d);
  QSV_LAUNCHED();
  return 0;
}

// B fragments of the three real products in the order the kernel reads them: [ob][ks][product][lane], lane l holds
// M_p[8 ob + (l >> 2)][4 ks + (l & 3)] with M_0 = P, M_1 = -(P + Q), M_2 = Q - P  (M = P + iQ row-major)
void dense_k_fragments(const double* host_matrix, int k, double* out) {
  const int dim = 1 << k, n_ob = dim / 8, n_ks = dim / 4;
  for (int ob = 0; ob < n_ob; ob++)
    for (int ks = 0; ks < n_ks; ks++)
      for (int l = 0; l < 32; l++) {
        const int row = 8 * ob + (l >> 2), col = 4 * ks + (l & 3);
        const double p = host_matrix[2 * ((size_t)row * dim + col)], q = host_matrix[2 * ((size_t)row * dim + col) + 1];
        double* o = out + (((size_t)ob * n_ks + ks) * 3) * 32 + l;
        o[0] = p;
        o[32] = -(p + q);
        o[64] = q - p;
      }
}

size_t dense_k_scratch_doubles(int k) { return (size_t)3 << (2 * k); }

// qubits[j] = matrix index bit j.  Needs n >= 11 and 4 <= k <= 7.
int launch_dense_k_dmma(double* sv, int n, const int32_t* qubits, int k, const double* host_matrix, double* dev_scratch,
                        cudaStream_t s) {
  if (n < kDkTile || k < 4 || k > 7) return fail(QSV_ERANGE, "dense-k tensor path: needs n >= 11 and 4 <= k <= 7");
  DenseKDmmaDesc d;
  memset(&d, 0, sizeof(d));
  d.n = n;
  d.k = k;
  d.n_tiles = 1u << (n - kDkTile);
  if (n - kDkTile > 31) return fail(QSV_ERANGE, "qsv_apply_matrix: grid too large");
  bool in_tile[64] = {false};
  int local_of[64];
  for (int q = 0; q < 64; q++) local_of[q] = -1;
  for (int j = 0; j < k; j++) {
    if (qubits[j] < 0 || qubits[j] >= n || in_tile[qubits[j]]) return fail(QSV_EINVAL, "qsv_apply_matrix: bad qubit list");
    in_tile[qubits[j]] = true;
    local_of[qubits[j]] = j;
  }
  int nt = k, next_local = k;
  for (int q = 0; q < n && nt < kDkTile; q++)  // the lowest other qubits: contiguous runs in HBM
    if (!in_tile[q]) { in_tile[q] = true; local_of[q] = next_local++; nt++; }
  int j = 0;
  for (int q = 0; q < n; q++)
    if (in_tile[q]) { d.tbs[j] = (uint8_t)q; d.loc[j] = (uint8_t)local_of[q]; j++; }
  std::vector<double> frags(dense_k_scratch_doubles(k));
  dense_k_fragments(host_matrix, k, frags.data());
  QSV_CUDA(cudaMemcpyAsync(dev_scratch, frags.data(), sizeof(double) * frags.size(), cudaMemcpyHostToDevice, s));
  QSV_CUDA(cudaStreamSynchronize(s));  // `frags` is a pageable temporary
  switch (k) {
    case 4: return launch_k<4>(sv, dev_scratch, d, s);
    case 5: return launch_k<5>(sv, dev_scratch, d, s);
    case 6: return launch_k<6>(sv, dev_scratch, d, s);
    default: return launch_k<7>(sv, dev_scratch, d, s);
  }
}

}  // namespace qsv
