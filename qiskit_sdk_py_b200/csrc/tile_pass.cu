// Cache-blocked gate application (K1-K5 of SURVEY 2.5) for sm_100a.
//
// One launch = one pass over the state: CTA t stages the 2^b amplitudes of tile t (the amplitudes
// that differ only in the b "tile qubits") in shared memory, applies a run of gates there and writes
// the tile back.  HBM traffic is exactly one read + one write of the vector per pass, however many
// gates the pass fuses.  Replaces runs of QasmSimulatorPy._add_unitary calls
// (/root/reference/qiskit/providers/basicaer/qasm_simulator.py:145-161), each of which is one
// out-of-place np.einsum over the whole vector.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace qsv {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

constexpr int kMaxOps = QSV_MAX_PASS_GATES;
constexpr int kMaxMat = 1536;  // doubles of matrix payload per pass (48 dense 2-qubit gates)

struct __align__(16) PassOp {
  uint8_t kind, t0, t1, nl;   // t0/t1: tile-local bit positions; nl: number of lane bits
  uint8_t l0, l1, l2, nins;   // lane bits (bank-conflict-free choice) and #zero-insert positions
  uint8_t ins[6];             // ascending tile-local positions where zeros are inserted
  uint16_t moff;              // offset (in doubles) of this op's matrix in PassDesc::mats
};

struct PassDesc {
  int32_t n, b, nops, pad;
  uint8_t tb[16];  // ascending physical bit position of each tile-local bit
  PassOp ops[kMaxOps];
  double mats[kMaxMat];
};

// group index w -> tile-local base index with zeros at the gate bits; the low `nl` bits of w land on
// the op's lane bits so that the 8 lanes of a quarter warp hit 8 different bank groups.
__device__ __forceinline__ uint32_t map_w(const PassOp& op, uint32_t w) {
  uint32_t lane = w & ((1u << op.nl) - 1u);
  uint32_t x = w >> op.nl;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    if (i < op.nins) {
      uint32_t p = op.ins[i];
      x = ((x >> p) << (p + 1)) | (x & ((1u << p) - 1u));
    }
  }
  x |= (lane & 1u) << op.l0;
  x |= ((lane >> 1) & 1u) << op.l1;
  x |= ((lane >> 2) & 1u) << op.l2;
  return x;
}

// tile-local index -> physical offset inside the tile's span
__device__ __forceinline__ u64 deposit_tile(const PassDesc& d, uint32_t li, int from_bit, int nbits) {
  u64 off = 0;
  for (int j = 0; j < nbits; j++) off |= (u64)((li >> (from_bit + j)) & 1u) << d.tb[from_bit + j];
  return off;
}

__global__ void __launch_bounds__(kThreads, 2)
tile_pass_kernel(double2* __restrict__ sv, const __grid_constant__ PassDesc d) {
  extern __shared__ double2 tile[];
  __shared__ u64 hi_off[32];

  const uint32_t tid = threadIdx.x;
  const int b = d.b;
  const uint32_t tile_amps = 1u << b;

  // physical base of this tile: spread blockIdx over the non-tile bits
  u64 base = blockIdx.x;
  for (int j = 0; j < b; j++) {
    const int p = d.tb[j];
    base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
  }
  // per-thread low part of the address, and the table for the high part of the loop index
  const int lo_bits = b < 8 ? b : 8;
  const u64 lo_off = deposit_tile(d, tid, 0, lo_bits);
  if (b > 8 && tid < (1u << (b - 8))) hi_off[tid] = deposit_tile(d, tid << 8, 8, b - 8);
  if (b <= 8 && tid == 0) hi_off[0] = 0;
  __syncthreads();

  double2* g = sv + base + lo_off;
  const uint32_t iters = tile_amps > kThreads ? tile_amps / kThreads : 1;
  const bool active = tid < tile_amps;

  // ---- load: one HBM read of the tile ----
  if (active) {
#pragma unroll 4
    for (uint32_t it = 0; it < iters; it++) {
      const uint32_t li = tid + it * kThreads;
      tile[swz(li)] = g[hi_off[it]];
    }
  }

  bool elementwise_prev = true;  // the load phase owns element li = tid + it*256, like DIAG ops
  for (int o = 0; o < d.nops; o++) {
    const PassOp op = d.ops[o];
    const bool elementwise = (op.kind == QSV_GATE_DIAG1 || op.kind == QSV_GATE_DIAG2);
    if (!(elementwise && elementwise_prev)) __syncthreads();
    elementwise_prev = elementwise;
    const double* m = d.mats + op.moff;

    if (op.kind == QSV_GATE_DENSE2) {
      double2 M[16];
#pragma unroll
      for (int i = 0; i < 16; i++) M[i] = make_double2(m[2 * i], m[2 * i + 1]);
      const uint32_t z0 = swz(1u << op.t0), z1 = swz(1u << op.t1);
      const uint32_t ngroups = tile_amps >> 2;
      for (uint32_t w = tid; w < ngroups; w += kThreads) {
        const uint32_t s0 = swz(map_w(op, w));
        const uint32_t s1 = s0 ^ z0, s2 = s0 ^ z1, s3 = s0 ^ z0 ^ z1;
        const double2 a0 = tile[s0], a1 = tile[s1], a2 = tile[s2], a3 = tile[s3];
        double2 r[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          r[i] = cmul(M[4 * i], a0);
          cfma(r[i], M[4 * i + 1], a1);
          cfma(r[i], M[4 * i + 2], a2);
          cfma(r[i], M[4 * i + 3], a3);
        }
        tile[s0] = r[0];
        tile[s1] = r[1];
        tile[s2] = r[2];
        tile[s3] = r[3];
      }
    } else if (op.kind == QSV_GATE_DENSE1) {
      const double2 m00 = make_double2(m[0], m[1]), m01 = make_double2(m[2], m[3]);
      const double2 m10 = make_double2(m[4], m[5]), m11 = make_double2(m[6], m[7]);
      const uint32_t z0 = swz(1u << op.t0);
      const uint32_t ngroups = tile_amps >> 1;
      for (uint32_t w = tid; w < ngroups; w += kThreads) {
        const uint32_t s0 = swz(map_w(op, w));
        const uint32_t s1 = s0 ^ z0;
        const double2 a0 = tile[s0], a1 = tile[s1];
        double2 r0 = cmul(m00, a0), r1 = cmul(m10, a0);
        cfma(r0, m01, a1);
        cfma(r1, m11, a1);
        tile[s0] = r0;
        tile[s1] = r1;
      }
    } else if (op.kind == QSV_GATE_DIAG1) {
      const double2 d0 = make_double2(m[0], m[1]), d1 = make_double2(m[2], m[3]);
      if (active) {
        for (uint32_t it = 0; it < iters; it++) {
          const uint32_t li = tid + it * kThreads;
          const uint32_t s = swz(li);
          tile[s] = cmul(((li >> op.t0) & 1u) ? d1 : d0, tile[s]);
        }
      }
    } else if (op.kind == QSV_GATE_DIAG2) {
      double2 dd[4];
#pragma unroll
      for (int i = 0; i < 4; i++) dd[i] = make_double2(m[2 * i], m[2 * i + 1]);
      if (active) {
        for (uint32_t it = 0; it < iters; it++) {
          const uint32_t li = tid + it * kThreads;
          const uint32_t s = swz(li);
          const uint32_t c = ((li >> op.t0) & 1u) | (((li >> op.t1) & 1u) << 1);
          const double2 f = c == 0 ? dd[0] : (c == 1 ? dd[1] : (c == 2 ? dd[2] : dd[3]));
          tile[s] = cmul(f, tile[s]);
        }
      }
    } else if (op.kind == QSV_GATE_CX) {
      // control t0 = 1: swap the target-0 and target-1 amplitudes
      const uint32_t zc = swz(1u << op.t0), zt = swz(1u << op.t1);
      const uint32_t ngroups = tile_amps >> 2;
      for (uint32_t w = tid; w < ngroups; w += kThreads) {
        const uint32_t s1 = swz(map_w(op, w)) ^ zc;
        const uint32_t s3 = s1 ^ zt;
        const double2 a = tile[s1], c = tile[s3];
        tile[s1] = c;
        tile[s3] = a;
      }
    }
  }
  __syncthreads();

  // ---- store: one HBM write of the tile ----
  if (active) {
#pragma unroll 4
    for (uint32_t it = 0; it < iters; it++) {
      const uint32_t li = tid + it * kThreads;
      g[hi_off[it]] = tile[swz(li)];
    }
  }
}

// ---- dense k-qubit gate, 3 <= k <= 12: out-of-place inside the tile (two smem buffers) ----
struct DenseKDesc {
  int32_t n, b, k, pad;
  uint8_t tb[16];   // tile bits, ascending physical positions
  uint8_t gt[16];   // tile-local position of gate qubit j (j = matrix index bit)
};

__global__ void __launch_bounds__(kThreads)
dense_k_kernel(double2* __restrict__ sv, const double2* __restrict__ mat,
               const __grid_constant__ DenseKDesc d) {
  extern __shared__ double2 smem[];
  const int b = d.b, k = d.k;
  const uint32_t tile_amps = 1u << b;
  const uint32_t K = 1u << k;
  double2* tin = smem;
  double2* tout = smem + tile_amps;
  uint32_t* coff = reinterpret_cast<uint32_t*>(smem + 2 * tile_amps);  // K entries
  const uint32_t tid = threadIdx.x;

  u64 base = blockIdx.x;
  for (int j = 0; j < b; j++) {
    const int p = d.tb[j];
    base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
  }
  uint32_t gmask = 0;
  for (int j = 0; j < k; j++) gmask |= 1u << d.gt[j];
  for (uint32_t c = tid; c < K; c += kThreads) {
    uint32_t off = 0;
    for (int j = 0; j < k; j++) off |= ((c >> j) & 1u) << d.gt[j];
    coff[c] = off;
  }
  for (uint32_t li = tid; li < tile_amps; li += kThreads) {
    u64 off = 0;
    for (int j = 0; j < b; j++) off |= (u64)((li >> j) & 1u) << d.tb[j];
    tin[li] = sv[base + off];
  }
  __syncthreads();
  for (uint32_t li = tid; li < tile_amps; li += kThreads) {
    uint32_t r = 0;
    for (int j = 0; j < k; j++) r |= ((li >> d.gt[j]) & 1u) << j;
    const uint32_t zb = li & ~gmask;
    const double2* row = mat + (size_t)r * K;
    double2 acc = make_double2(0.0, 0.0);
    for (uint32_t c = 0; c < K; c++) cfma(acc, __ldg(row + c), tin[zb | coff[c]]);
    tout[li] = acc;
  }
  __syncthreads();
  for (uint32_t li = tid; li < tile_amps; li += kThreads) {
    u64 off = 0;
    for (int j = 0; j < b; j++) off |= (u64)((li >> j) & 1u) << d.tb[j];
    sv[base + off] = tout[li];
  }
}

// ---- elementwise helpers ----
__global__ void __launch_bounds__(kThreads)
fill_zero_kernel(double2* __restrict__ sv, u64 n_amps) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride)
    sv[i] = make_double2(0.0, 0.0);
}

__global__ void set_diag_kernel(double2* __restrict__ sv, u64 count, u64 stride_amps, double2 v) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) sv[i * stride_amps] = v;
}

__global__ void __launch_bounds__(kThreads)
scale_kernel(double2* __restrict__ sv, u64 n_amps, double2 f) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride)
    sv[i] = cmul(f, sv[i]);
}

__global__ void __launch_bounds__(kThreads)
chop_kernel(double2* __restrict__ sv, u64 n_amps, double thr) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const double2 a = sv[i];
    // numpy abs(complex) is hypot(re, im); qasm_simulator.py:333
    if (hypot(a.x, a.y) < thr) sv[i] = make_double2(0.0, 0.0);
  }
}

static int grid_for(u64 n_amps) {
  u64 blocks = (n_amps + kThreads - 1) / kThreads;
  const u64 cap = (u64)kNumSMs * 16;
  return (int)std::max<u64>(1, std::min(blocks, cap));
}

// Pick three free tile bits with pairwise different residues mod 3 (see swz()), smallest first.
static void pick_lane_bits(int b, uint32_t gate_mask, PassOp* op) {
  int freeb[16], nf = 0;
  for (int j = 0; j < b; j++)
    if (!((gate_mask >> j) & 1u)) freeb[nf++] = j;
  int chosen[3], nc = 0;
  bool used_res[3] = {false, false, false};
  for (int i = 0; i < nf && nc < 3; i++) {
    if (!used_res[freeb[i] % 3]) {
      used_res[freeb[i] % 3] = true;
      chosen[nc++] = freeb[i];
    }
  }
  // not enough distinct residues (tiny tiles): fill with the lowest remaining free bits
  for (int i = 0; i < nf && nc < 3; i++) {
    bool dup = false;
    for (int c = 0; c < nc; c++) dup |= (chosen[c] == freeb[i]);
    if (!dup) chosen[nc++] = freeb[i];
  }
  op->nl = (uint8_t)nc;
  op->l0 = nc > 0 ? chosen[0] : 0;
  op->l1 = nc > 1 ? chosen[1] : 0;
  op->l2 = nc > 2 ? chosen[2] : 0;
  uint32_t ins_mask = gate_mask;
  for (int c = 0; c < nc; c++) ins_mask |= 1u << chosen[c];
  op->nins = 0;
  for (int j = 0; j < b; j++)
    if ((ins_mask >> j) & 1u) op->ins[op->nins++] = (uint8_t)j;
}

static bool g_attr_set[64] = {false};

// Opt in to > 48 KiB of dynamic shared memory, once per device.
static int ensure_attrs() {
  int dev = 0;
  QSV_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(QSV_ERANGE, "device ordinal out of range");
  if (!g_attr_set[dev]) {
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double2) << QSV_MAX_TILE_QUBITS)));
    QSV_CUDA(cudaFuncSetAttribute(dense_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * (sizeof(double2) << 12) + 4096 * sizeof(uint32_t))));
    g_attr_set[dev] = true;
  }
  return 0;
}

static int launch_pass(double* sv, int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                       int n_gates, cudaStream_t stream) {
  if (!sv || n < 1 || n > 40) return fail(QSV_EINVAL, "qsv_apply_pass: bad state / n_qubits");
  if (n_tile < 1 || n_tile > QSV_MAX_TILE_QUBITS || n_tile > n)
    return fail(QSV_ERANGE, "qsv_apply_pass: n_tile must be in [1, min(n_qubits, 13)]");
  if (n_gates < 0 || n_gates > kMaxOps) return fail(QSV_ERANGE, "qsv_apply_pass: too many gates");
  if (n - n_tile > 31) return fail(QSV_ERANGE, "qsv_apply_pass: grid too large");

  static thread_local PassDesc d;
  memset(&d, 0, sizeof(d));
  d.n = n;
  d.b = n_tile;
  int sorted[16];
  for (int j = 0; j < n_tile; j++) {
    if (tile_qubits[j] < 0 || tile_qubits[j] >= n) return fail(QSV_EINVAL, "qsv_apply_pass: tile qubit out of range");
    sorted[j] = tile_qubits[j];
  }
  std::sort(sorted, sorted + n_tile);
  for (int j = 0; j + 1 < n_tile; j++)
    if (sorted[j] == sorted[j + 1]) return fail(QSV_EINVAL, "qsv_apply_pass: duplicate tile qubit");
  int local_of[64];
  for (int q = 0; q < 64; q++) local_of[q] = -1;
  for (int j = 0; j < n_tile; j++) {
    d.tb[j] = (uint8_t)sorted[j];
    local_of[sorted[j]] = j;
  }
  int moff = 0;
  for (int i = 0; i < n_gates; i++) {
    const qsv_gate_t& gsrc = gates[i];
    PassOp& op = d.ops[i];
    int nq, ndbl;
    switch (gsrc.kind) {
      case QSV_GATE_DENSE1: nq = 1; ndbl = 8; break;
      case QSV_GATE_DENSE2: nq = 2; ndbl = 32; break;
      case QSV_GATE_DIAG1: nq = 1; ndbl = 4; break;
      case QSV_GATE_DIAG2: nq = 2; ndbl = 8; break;
      case QSV_GATE_CX: nq = 2; ndbl = 0; break;
      default: return fail(QSV_EINVAL, "qsv_apply_pass: unknown gate kind");
    }
    if (gsrc.q0 < 0 || gsrc.q0 >= n || local_of[gsrc.q0] < 0)
      return fail(QSV_EINVAL, "qsv_apply_pass: gate qubit q0 is not a tile qubit");
    if (nq == 2 && (gsrc.q1 < 0 || gsrc.q1 >= n || local_of[gsrc.q1] < 0 || gsrc.q1 == gsrc.q0))
      return fail(QSV_EINVAL, "qsv_apply_pass: gate qubit q1 is not a tile qubit / equals q0");
    if (moff + ndbl > kMaxMat) return fail(QSV_ERANGE, "qsv_apply_pass: matrix payload too large");
    op.kind = (uint8_t)gsrc.kind;
    op.t0 = (uint8_t)local_of[gsrc.q0];
    op.t1 = (uint8_t)(nq == 2 ? local_of[gsrc.q1] : 0);
    op.moff = (uint16_t)moff;
    memcpy(d.mats + moff, gsrc.m, sizeof(double) * ndbl);
    moff += ndbl;
    uint32_t gmask = 1u << op.t0;
    if (nq == 2) gmask |= 1u << op.t1;
    pick_lane_bits(n_tile, gmask, &op);
  }
  d.nops = n_gates;

  if (int rc = ensure_attrs()) return rc;
  const size_t smem = sizeof(double2) << n_tile;
  const unsigned grid = 1u << (n - n_tile);
  tile_pass_kernel<<<grid, kThreads, smem, stream>>>(reinterpret_cast<double2*>(sv), d);
  QSV_LAUNCHED();
  return 0;
}

// Default tile for a stand-alone gate: its qubits plus the lowest other qubits.
static int default_tile(int n, const int32_t* qubits, int k, int want, int32_t* out) {
  int nt = 0;
  bool in[64] = {false};
  for (int j = 0; j < k; j++) {
    if (qubits[j] < 0 || qubits[j] >= n || in[qubits[j]]) return -1;
    in[qubits[j]] = true;
    out[nt++] = qubits[j];
  }
  for (int q = 0; q < n && nt < want; q++)
    if (!in[q]) out[nt++] = q;
  return nt;
}

}  // namespace qsv

using namespace qsv;

extern "C" {

int qsv_version(void) { return QSV_VERSION; }
const char* qsv_last_error(void) { return g_err; }
uint64_t qsv_launch_count(void) { return g_launches; }

int qsv_init_basis0(double* sv, int n, double pr, double pi, void* stream) {
  if (!sv || n < 0 || n > 40) return fail(QSV_EINVAL, "qsv_init_basis0: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const u64 n_amps = 1ull << n;
  fill_zero_kernel<<<grid_for(n_amps), kThreads, 0, s>>>(reinterpret_cast<double2*>(sv), n_amps);
  QSV_LAUNCHED();
  set_diag_kernel<<<1, 32, 0, s>>>(reinterpret_cast<double2*>(sv), 1, 1, make_double2(pr, pi));
  QSV_LAUNCHED();
  return 0;
}

int qsv_init_identity(double* sv, int n, double pr, double pi, void* stream) {
  if (!sv || n < 0 || n > 20) return fail(QSV_EINVAL, "qsv_init_identity: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const u64 dim = 1ull << n;
  const u64 n_amps = dim * dim;
  fill_zero_kernel<<<grid_for(n_amps), kThreads, 0, s>>>(reinterpret_cast<double2*>(sv), n_amps);
  QSV_LAUNCHED();
  set_diag_kernel<<<(unsigned)((dim + 255) / 256), 256, 0, s>>>(reinterpret_cast<double2*>(sv), dim, dim + 1,
                                                                make_double2(pr, pi));
  QSV_LAUNCHED();
  return 0;
}

int qsv_scale(double* sv, int n, double re, double im, void* stream) {
  if (!sv || n < 0 || n > 40) return fail(QSV_EINVAL, "qsv_scale: bad arguments");
  const u64 n_amps = 1ull << n;
  scale_kernel<<<grid_for(n_amps), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<double2*>(sv), n_amps,
                                                                       make_double2(re, im));
  QSV_LAUNCHED();
  return 0;
}

int qsv_upload(double* sv, int n, const double* host_amps, size_t n_amps, double pr, double pi, void* stream) {
  if (!sv || !host_amps || n < 0 || n > 40) return fail(QSV_EINVAL, "qsv_upload: bad arguments");
  if (n_amps != ((size_t)1 << n)) return fail(QSV_EINVAL, "qsv_upload: n_amps != 2^n_qubits");
  QSV_CUDA(cudaMemcpyAsync(sv, host_amps, n_amps * 16, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  if (pr != 1.0 || pi != 0.0) return qsv_scale(sv, n, pr, pi, stream);
  return 0;
}

int qsv_chop(double* sv, int n, double thr, void* stream) {
  if (!sv || n < 0 || n > 40) return fail(QSV_EINVAL, "qsv_chop: bad arguments");
  const u64 n_amps = 1ull << n;
  chop_kernel<<<grid_for(n_amps), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<double2*>(sv), n_amps, thr);
  QSV_LAUNCHED();
  return 0;
}

int qsv_download(const double* sv, int n, double* host_amps, void* stream) {
  if (!sv || !host_amps || n < 0 || n > 40) return fail(QSV_EINVAL, "qsv_download: bad arguments");
  QSV_CUDA(cudaMemcpyAsync(host_amps, sv, ((size_t)16) << n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  QSV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int qsv_apply_pass(double* sv, int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                   int n_gates, void* stream) {
  if (!tile_qubits || (n_gates > 0 && !gates)) return fail(QSV_EINVAL, "qsv_apply_pass: null pointer");
  return launch_pass(sv, n, tile_qubits, n_tile, gates, n_gates, (cudaStream_t)stream);
}

int qsv_apply_matrix(double* sv, int n, const int32_t* qubits, int k, const double* host_matrix,
                     double* dev_mat_scratch, void* stream) {
  if (!sv || !qubits || !host_matrix || k < 1 || k > n) return fail(QSV_EINVAL, "qsv_apply_matrix: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  int32_t tile[16];
  if (k <= 2) {
    const int want = std::min(n, 12);
    const int nt = default_tile(n, qubits, k, want, tile);
    if (nt < 0) return fail(QSV_EINVAL, "qsv_apply_matrix: bad qubit list");
    qsv_gate_t g;
    memset(&g, 0, sizeof(g));
    g.kind = k == 1 ? QSV_GATE_DENSE1 : QSV_GATE_DENSE2;
    g.q0 = qubits[0];
    g.q1 = k == 2 ? qubits[1] : 0;
    memcpy(g.m, host_matrix, sizeof(double) * (k == 1 ? 8 : 32));
    return launch_pass(sv, n, tile, nt, &g, 1, s);
  }
  if (k > QSV_MAX_DENSE_K) return fail(QSV_ERANGE, "qsv_apply_matrix: k > 12 is not supported");
  if (!dev_mat_scratch) return fail(QSV_EINVAL, "qsv_apply_matrix: k >= 3 needs dev_mat_scratch");
  const int want = std::min(n, std::min(12, std::max(k + 3, 10)));
  const int nt = default_tile(n, qubits, k, std::max(want, k), tile);
  if (nt < 0) return fail(QSV_EINVAL, "qsv_apply_matrix: bad qubit list");
  DenseKDesc d;
  memset(&d, 0, sizeof(d));
  d.n = n;
  d.b = nt;
  d.k = k;
  int sorted[16];
  for (int j = 0; j < nt; j++) sorted[j] = tile[j];
  std::sort(sorted, sorted + nt);
  for (int j = 0; j < nt; j++) d.tb[j] = (uint8_t)sorted[j];
  for (int j = 0; j < k; j++)
    for (int t = 0; t < nt; t++)
      if (sorted[t] == qubits[j]) d.gt[j] = (uint8_t)t;
  const size_t mat_bytes = sizeof(double) * 2 * ((size_t)1 << (2 * k));
  QSV_CUDA(cudaMemcpyAsync(dev_mat_scratch, host_matrix, mat_bytes, cudaMemcpyHostToDevice, s));
  if (int rc = ensure_attrs()) return rc;
  const size_t smem = 2 * (sizeof(double2) << nt) + (sizeof(uint32_t) << k);
  dense_k_kernel<<<1u << (n - nt), kThreads, smem, s>>>(reinterpret_cast<double2*>(sv),
                                                        reinterpret_cast<const double2*>(dev_mat_scratch), d);
  QSV_LAUNCHED();
  return 0;
}

int qsv_collapse(double* sv, int n, int qubit, int outcome, double prob, int is_reset, void* stream) {
  if (!sv || qubit < 0 || qubit >= n || (outcome != 0 && outcome != 1) || !(prob > 0.0))
    return fail(QSV_EINVAL, "qsv_collapse: bad arguments");
  const double s = 1.0 / sqrt(prob);
  int32_t tile[16];
  const int32_t q = qubit;
  const int nt = default_tile(n, &q, 1, std::min(n, 12), tile);
  qsv_gate_t g;
  memset(&g, 0, sizeof(g));
  g.q0 = qubit;
  if (is_reset && outcome == 1) {
    g.kind = QSV_GATE_DENSE1;  // [[0, s], [0, 0]]  (qasm_simulator.py:272)
    g.m[2] = s;
  } else {
    g.kind = QSV_GATE_DIAG1;   // diag(s,0) or diag(0,s)  (qasm_simulator.py:249-251, :269)
    g.m[outcome == 0 ? 0 : 2] = s;
  }
  return launch_pass(sv, n, tile, nt, &g, 1, (cudaStream_t)stream);
}

}  // extern "C"
