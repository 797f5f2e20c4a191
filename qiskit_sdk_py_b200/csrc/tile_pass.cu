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
constexpr int kPT = 512;       // threads of the persistent pass kernel (16 warps, one CTA per SM)
constexpr int kTidBits = 9;

// One gate of a pass, lowered to XOR masks.  Every index map in the kernel is linear over GF(2):
//   slot(w) = XOR_j bit_j(w) * col[j]   is the (bank-swizzled) shared-memory slot of the first
// amplitude of group w; the other members of the group are slot ^ z0, slot ^ z1, slot ^ z0 ^ z1.
// The group index is split as  w = lane part (lb bits) | warp id (4 bits) | iteration (rest).
struct __align__(8) PassOp {
  uint8_t kind, t0, t1, nbits;  // t0/t1: tile-local gate bits; nbits = b - k (bits of the group index)
  uint16_t moff;                // offset (doubles) of the matrix in PassDesc::mats
  uint16_t z0, z1;              // swz(1 << t0), swz(1 << t1)
  uint8_t lb, pad;              // lane bits: 3 for the DMMA path (8 groups per warp step), else 5
  uint16_t col[13];             // col[j] = swz(1 << position of group-index bit j)
};

struct PassDesc {
  int32_t n, b, nops, pad;
  uint32_t n_tiles;
  uint32_t pad2[3];
  uint8_t tb[16];   // ascending physical bit position of each tile-local bit
  uint8_t vec[16];  // bank vector of each tile-local bit (pass-specific XOR swizzle), vec[j] = 1<<j for j<3
  PassOp ops[kMaxOps];
  double mats[kMaxMat];
};

// Pass-specific XOR swizzle of a tile-local amplitude index (16-byte slots; bank group = slot & 7).
// Tile bit j >= 3 toggles the bank bits vec[j]; three index bits whose vectors are linearly independent
// over GF(2) spread over all 8 bank groups, so the accesses of a quarter warp (LDS/STS.128) or of a
// half warp reading 8-byte halves (LDS.64) are conflict-free.  The host picks vec[] per pass so that
// the two gate bits of every dense gate get different vectors.
__host__ __device__ __forceinline__ uint32_t swz(const uint8_t* vec, int b, uint32_t li) {
  uint32_t f = 0;
  for (int j = 3; j < b; j++)
    if ((li >> j) & 1u) f ^= vec[j];
  return li ^ f;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor path (SASS: DMMA.8x8x4).
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// physical base address (in amplitudes) of tile t: spread t over the non-tile bits
__device__ __forceinline__ u64 tile_base(const PassDesc& d, u64 t) {
  for (int j = 0; j < d.b; j++) {
    const int p = d.tb[j];
    t = ((t >> p) << (p + 1)) | (t & ((1ull << p) - 1ull));
  }
  return t;
}

// Persistent, double-buffered pass kernel: CTA c processes tiles c, c+grid, c+2*grid, ...  While the
// gates of tile i run out of buffer (i & 1), the cp.async loads of tile i+1 stream into the other
// buffer, so HBM reads overlap the FP64 work; the write-back is plain stores, which retire
// asynchronously behind the next tile's compute.
//
// Dense 2-qubit gates run on the FP64 tensor path: a 4x4 complex gate is the 8x8 real matrix
// [[Re,-Im],[Im,Re]] acting on (re,im) of the 4 amplitudes of a group; one warp step takes 8 groups
// as the A operand (8 groups x 4 real components, two k-slices) against the gate as the B operand,
// and each lane receives one complex output amplitude (re,im) -- a single 16-byte store.
__global__ void __launch_bounds__(kPT, 1)
tile_pass_kernel(double2* __restrict__ sv, const __grid_constant__ PassDesc d) {
  extern __shared__ double2 smem[];
  __shared__ u64 hi_off[8];
  __shared__ uint16_t swz_it[8];
  __shared__ uint16_t tab[kMaxOps][56];  // per op: [0,32) lane part, [32,48) warp part, [48,56) iteration part
  __shared__ double frag[kMaxOps][64];   // per dense op: B fragments of the two k-slices, frag[s*32 + lane]

  const uint32_t tid = threadIdx.x;
  const int b = d.b;
  const uint32_t tile_amps = 1u << b;

  // ---- per-kernel setup (amortised over all tiles of this CTA) ----
  const int lo_bits = b < kTidBits ? b : kTidBits;
  u64 lo_off = 0;
  for (int j = 0; j < lo_bits; j++) lo_off |= (u64)((tid >> j) & 1u) << d.tb[j];
  if (tid < 8) {
    u64 off = 0;
    for (int j = kTidBits; j < b; j++) off |= (u64)((tid >> (j - kTidBits)) & 1u) << d.tb[j];
    hi_off[tid] = off;
    swz_it[tid] = (uint16_t)swz(d.vec, b, tid << kTidBits);
  }
  for (int idx = tid; idx < d.nops * 56; idx += kPT) {
    const int o = idx / 56, e = idx % 56;
    const int lb = d.ops[o].lb;
    uint32_t v = 0;
    if (e < 32) {
      for (int j = 0; j < 5; j++)
        if ((e >> j) & 1) v ^= d.ops[o].col[j];
    } else if (e < 48) {
      for (int j = 0; j < 4; j++)
        if (((e - 32) >> j) & 1) v ^= d.ops[o].col[lb + j];
    } else {
      for (int j = 0; j < 3; j++)
        if (((e - 48) >> j) & 1) v ^= d.ops[o].col[lb + 4 + j];
    }
    tab[o][e] = (uint16_t)v;
  }
  for (int idx = tid; idx < d.nops * 64; idx += kPT) {
    const int o = idx >> 6, e = idx & 63;
    if (d.ops[o].kind == QSV_GATE_DENSE2) {
      // B fragment of k-slice s: lane l holds Mreal[n = l >> 2][k = 4 s + (l & 3)]
      const int s = e >> 5, l = e & 31;
      const int nrow = l >> 2, kcol = 4 * s + (l & 3);
      const int r = nrow >> 1, c = kcol >> 1;
      const double* m = d.mats + d.ops[o].moff + 2 * (4 * r + c);
      double v;
      if ((nrow & 1) == 0) v = (kcol & 1) == 0 ? m[0] : -m[1];
      else v = (kcol & 1) == 0 ? m[1] : m[0];
      frag[o][e] = v;
    }
  }
  const uint32_t s_tid = swz(d.vec, b, tid);
  const uint32_t iters = tile_amps > kPT ? tile_amps / kPT : 1;
  const bool active = tid < tile_amps;
  const uint32_t lane = tid & 31u, wid = tid >> 5;
  __syncthreads();

  const u64 n_tiles = d.n_tiles;
  u64 t = blockIdx.x;
  u64 base_cur = 0, base_next = 0;
  if (t < n_tiles) {
    base_cur = tile_base(d, t);
    if (active) {
      const double2* g = sv + base_cur + lo_off;
#pragma unroll 4
      for (uint32_t it = 0; it < iters; it++) cp_async16(smem + (s_tid ^ swz_it[it]), g + hi_off[it]);
    }
    cp_async_commit();
  }
  uint32_t cur = 0;
  for (; t < n_tiles; t += gridDim.x, cur ^= 1u, base_cur = base_next) {
    double2* tile = smem + (cur ? tile_amps : 0u);
    const u64 t_next = t + gridDim.x;
    if (t_next < n_tiles) {
      base_next = tile_base(d, t_next);
      if (active) {
        double2* dst = smem + (cur ? 0u : tile_amps);
        const double2* g = sv + base_next + lo_off;
#pragma unroll 4
        for (uint32_t it = 0; it < iters; it++) cp_async16(dst + (s_tid ^ swz_it[it]), g + hi_off[it]);
      }
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    bool elementwise_prev = true;  // cp.async wrote element li = tid + it*512, same owner as DIAG ops
    for (int o = 0; o < d.nops; o++) {
      const PassOp& op = d.ops[o];
      const int kind = op.kind;
      const bool elementwise = (kind == QSV_GATE_DIAG1 || kind == QSV_GATE_DIAG2);
      if (o > 0 && !(elementwise && elementwise_prev)) __syncthreads();
      elementwise_prev = elementwise;
      const double* m = d.mats + op.moff;
      const uint32_t ngroups = 1u << op.nbits;

      if (kind == QSV_GATE_DENSE2) {
        // ---- FP64 tensor path: 8 groups per warp step ----
        const uint32_t g = lane >> 2, c = lane & 3u;
        const double b0 = frag[o][lane], b1 = frag[o][32 + lane];
        const uint32_t z0 = op.z0, z1 = op.z1;
        const uint32_t ld_x = (c & 2u) ? z0 : 0u;                                   // amplitude c>>1 of the slice
        const uint32_t st_x = ((c & 1u) ? z0 : 0u) ^ ((c & 2u) ? z1 : 0u);          // amplitude c of the group
        const uint32_t s_lw = (uint32_t)tab[o][g] ^ (uint32_t)tab[o][32 + wid];
        const uint32_t w_lw = g | (wid << 3);
        const char* tb8 = reinterpret_cast<const char*>(tile) + (c & 1u) * 8u;
        const uint32_t n_it = ngroups > 128u ? (ngroups >> 7) : 1u;
#pragma unroll 4
        for (uint32_t it = 0; it < n_it; it++) {
          const uint32_t s = s_lw ^ (uint32_t)tab[o][48 + it];
          const bool valid = (w_lw | (it << 7)) < ngroups;
          double x0 = 0.0, x1 = 0.0;
          if (valid) {
            const uint32_t sa = s ^ ld_x;
            x0 = *reinterpret_cast<const double*>(tb8 + (size_t)sa * 16u);
            x1 = *reinterpret_cast<const double*>(tb8 + (size_t)(sa ^ z1) * 16u);
          }
          double d0 = 0.0, d1 = 0.0;
          dmma884(d0, d1, x0, b0);
          dmma884(d0, d1, x1, b1);
          if (valid) tile[s ^ st_x] = make_double2(d0, d1);
        }
      } else if (kind == QSV_GATE_DENSE1) {
        // scalar path, only used when the tile has a single qubit (no partner bit for the tensor path)
        const double2 m00 = make_double2(m[0], m[1]), m01 = make_double2(m[2], m[3]);
        const double2 m10 = make_double2(m[4], m[5]), m11 = make_double2(m[6], m[7]);
        const uint32_t z0 = op.z0;
        const uint32_t s_w = (uint32_t)tab[o][lane] ^ (uint32_t)tab[o][32 + wid];
        uint32_t it = 0;
        for (uint32_t w = tid; w < ngroups; w += kPT, it++) {
          const uint32_t s0 = s_w ^ (uint32_t)tab[o][48 + it];
          const uint32_t s1 = s0 ^ z0;
          const double2 a0 = tile[s0], a1 = tile[s1];
          double2 r0 = cmul(m00, a0), r1 = cmul(m10, a0);
          cfma(r0, m01, a1);
          cfma(r1, m11, a1);
          tile[s0] = r0;
          tile[s1] = r1;
        }
      } else if (kind == QSV_GATE_DIAG1) {
        const double2 d0 = make_double2(m[0], m[1]), d1 = make_double2(m[2], m[3]);
        if (active) {
          for (uint32_t it = 0; it < iters; it++) {
            const uint32_t li = tid + it * kPT;
            const uint32_t s = s_tid ^ swz_it[it];
            tile[s] = cmul(((li >> op.t0) & 1u) ? d1 : d0, tile[s]);
          }
        }
      } else if (kind == QSV_GATE_DIAG2) {
        double2 dd[4];
#pragma unroll
        for (int i = 0; i < 4; i++) dd[i] = make_double2(m[2 * i], m[2 * i + 1]);
        if (active) {
          for (uint32_t it = 0; it < iters; it++) {
            const uint32_t li = tid + it * kPT;
            const uint32_t s = s_tid ^ swz_it[it];
            const uint32_t c = ((li >> op.t0) & 1u) | (((li >> op.t1) & 1u) << 1);
            const double2 f = c == 0 ? dd[0] : (c == 1 ? dd[1] : (c == 2 ? dd[2] : dd[3]));
            tile[s] = cmul(f, tile[s]);
          }
        }
      } else if (kind == QSV_GATE_CX) {
        // control t0 = 1: swap the target-0 and target-1 amplitudes
        const uint32_t zc = op.z0, zt = op.z1;
        const uint32_t s_w = (uint32_t)tab[o][lane] ^ (uint32_t)tab[o][32 + wid] ^ zc;
        uint32_t it = 0;
        for (uint32_t w = tid; w < ngroups; w += kPT, it++) {
          const uint32_t s1 = s_w ^ (uint32_t)tab[o][48 + it];
          const uint32_t s3 = s1 ^ zt;
          const double2 a = tile[s1], c = tile[s3];
          tile[s1] = c;
          tile[s3] = a;
        }
      }
    }
    __syncthreads();

    // ---- write-back: one HBM write of the tile ----
    if (active) {
      double2* g = sv + base_cur + lo_off;
#pragma unroll 4
      for (uint32_t it = 0; it < iters; it++) g[hi_off[it]] = tile[s_tid ^ swz_it[it]];
    }
    __syncthreads();  // the next iteration prefetches into this buffer
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

static inline int bank_vec(const PassDesc& d, int bit) { return d.vec[bit] & 7; }

static inline bool independent3(int a, int b, int c) {
  // three vectors of GF(2)^3 are independent iff none is zero, all differ, and c != a ^ b
  return a && b && c && a != b && a != c && b != c && c != (a ^ b);
}

// Choose the pass-specific bank vectors: bits 0..2 are the identity; every higher tile bit gets a
// non-zero vector such that, where possible, the two bits of each dense gate get different vectors.
static void choose_bank_vectors(PassDesc& d, int n_ops, const uint8_t* t0s, const uint8_t* t1s,
                                const uint8_t* dense) {
  const int b = d.b;
  for (int j = 0; j < 16; j++) d.vec[j] = (uint8_t)(1u << (j % 3));
  auto conflicts = [&]() {
    int c = 0;
    for (int i = 0; i < n_ops; i++)
      if (dense[i] && d.vec[t0s[i]] == d.vec[t1s[i]]) c++;
    return c;
  };
  int best = conflicts();
  for (int round = 0; round < 4 && best > 0; round++) {
    for (int i = 0; i < n_ops && best > 0; i++) {
      if (!dense[i] || d.vec[t0s[i]] != d.vec[t1s[i]]) continue;
      const int cand[2] = {t1s[i], t0s[i]};
      for (int ci = 0; ci < 2 && d.vec[t0s[i]] == d.vec[t1s[i]]; ci++) {
        const int bit = cand[ci];
        if (bit < 3 || bit >= b) continue;
        const uint8_t old = d.vec[bit];
        uint8_t best_v = old;
        for (uint8_t v = 1; v < 8; v++) {
          d.vec[bit] = v;
          const int c = conflicts();
          if (c < best) {
            best = c;
            best_v = v;
          }
        }
        d.vec[bit] = best_v;
      }
    }
  }
}

// Order the free (non-gate) tile bits of an op for its group index and fill op->col / nbits / lb.
// The first group-index bits are the ones that vary across the lanes served by one shared-memory
// wavefront, so they are chosen with linearly independent bank vectors:
//   DMMA dense gate : {f0, t0, t1} independent (STS.128 of a quarter warp: 2 groups x 4 amplitudes)
//                     {f0, f1, t0} independent (LDS.64 of a half warp: 4 groups x 2 amplitudes)
//   other gates     : {f0, f1, f2} independent (LDS/STS.128 of a quarter warp: 8 groups)
static void build_group_map(const PassDesc& d, uint32_t gate_mask, bool dmma, PassOp* op) {
  const int b = d.b;
  int freeb[16], nf = 0;
  for (int j = 0; j < b; j++)
    if (!((gate_mask >> j) & 1u)) freeb[nf++] = j;
  int lead[3] = {-1, -1, -1};
  if (dmma) {
    const int v0 = bank_vec(d, op->t0), v1 = bank_vec(d, op->t1);
    for (int i = 0; i < nf && lead[0] < 0; i++)
      if (independent3(bank_vec(d, freeb[i]), v0, v1)) lead[0] = freeb[i];
    if (lead[0] < 0) {  // t0/t1 share a vector (or tiny tile): at least avoid t0's vector
      for (int i = 0; i < nf && lead[0] < 0; i++)
        if (bank_vec(d, freeb[i]) != v0) lead[0] = freeb[i];
    }
    if (lead[0] >= 0) {
      for (int i = 0; i < nf && lead[1] < 0; i++)
        if (freeb[i] != lead[0] && independent3(bank_vec(d, lead[0]), bank_vec(d, freeb[i]), v0)) lead[1] = freeb[i];
    }
  } else {
    for (int i0 = 0; i0 < nf && lead[2] < 0; i0++)
      for (int i1 = i0 + 1; i1 < nf && lead[2] < 0; i1++)
        for (int i2 = i1 + 1; i2 < nf && lead[2] < 0; i2++)
          if (independent3(bank_vec(d, freeb[i0]), bank_vec(d, freeb[i1]), bank_vec(d, freeb[i2]))) {
            lead[0] = freeb[i0];
            lead[1] = freeb[i1];
            lead[2] = freeb[i2];
          }
  }
  int order[16], no = 0;
  bool taken[16] = {false};
  for (int k = 0; k < 3; k++) {
    if (lead[k] < 0) continue;
    bool dup = false;
    for (int q = 0; q < no; q++) dup |= (order[q] == lead[k]);
    if (dup) continue;
    order[no++] = lead[k];
    taken[lead[k]] = true;
  }
  for (int i = 0; i < nf; i++)
    if (!taken[freeb[i]]) order[no++] = freeb[i];
  op->nbits = (uint8_t)nf;
  op->lb = dmma ? 3 : 5;
  for (int j = 0; j < 13; j++) op->col[j] = j < nf ? (uint16_t)swz(d.vec, b, 1u << order[j]) : 0;
}

static bool g_attr_set[64] = {false};
static int g_sm_count[64] = {0};

static int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
  if (!g_sm_count[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMs;
    g_sm_count[dev] = v;
  }
  return g_sm_count[dev];
}

// Opt in to > 48 KiB of dynamic shared memory, once per device.
static int ensure_attrs() {
  int dev = 0;
  QSV_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(QSV_ERANGE, "device ordinal out of range");
  if (!g_attr_set[dev]) {
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * (sizeof(double2) << QSV_MAX_TILE_QUBITS))));
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
    return fail(QSV_ERANGE, "qsv_apply_pass: n_tile must be in [1, min(n_qubits, 12)]");
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
  // first pass over the gates: kinds, tile-local bits, matrices (1-qubit dense gates are widened to
  // I (x) M on a partner bit so that they run on the same FP64 tensor path as 2-qubit gates)
  int moff = 0;
  uint8_t t0s[kMaxOps], t1s[kMaxOps], dense[kMaxOps];
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
    op.kind = (uint8_t)gsrc.kind;
    op.t0 = (uint8_t)local_of[gsrc.q0];
    op.t1 = (uint8_t)(nq == 2 ? local_of[gsrc.q1] : 0);
    op.moff = (uint16_t)moff;
    if (gsrc.kind == QSV_GATE_DENSE1 && n_tile >= 2) {
      // partner = the next tile bit (cyclically); matrix index bit 1 = partner: kron(I2, M)
      op.kind = QSV_GATE_DENSE2;
      op.t1 = (uint8_t)((op.t0 + 1) % n_tile);
      ndbl = 32;
      if (moff + ndbl > kMaxMat) return fail(QSV_ERANGE, "qsv_apply_pass: matrix payload too large");
      double* m4 = d.mats + moff;
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          const bool same = (r >> 1) == (c >> 1);
          m4[2 * (4 * r + c)] = same ? gsrc.m[2 * (2 * (r & 1) + (c & 1))] : 0.0;
          m4[2 * (4 * r + c) + 1] = same ? gsrc.m[2 * (2 * (r & 1) + (c & 1)) + 1] : 0.0;
        }
    } else {
      if (moff + ndbl > kMaxMat) return fail(QSV_ERANGE, "qsv_apply_pass: matrix payload too large");
      memcpy(d.mats + moff, gsrc.m, sizeof(double) * ndbl);
    }
    moff += ndbl;
    t0s[i] = op.t0;
    t1s[i] = op.t1;
    dense[i] = op.kind == QSV_GATE_DENSE2;
  }
  choose_bank_vectors(d, n_gates, t0s, t1s, dense);
  for (int i = 0; i < n_gates; i++) {
    PassOp& op = d.ops[i];
    const bool two = !(op.kind == QSV_GATE_DENSE1 || op.kind == QSV_GATE_DIAG1);
    uint32_t gmask = 1u << op.t0;
    if (two) gmask |= 1u << op.t1;
    op.z0 = (uint16_t)swz(d.vec, n_tile, 1u << op.t0);
    op.z1 = (uint16_t)swz(d.vec, n_tile, 1u << op.t1);
    build_group_map(d, gmask, op.kind == QSV_GATE_DENSE2, &op);
  }
  if (n_gates == 0) choose_bank_vectors(d, 0, t0s, t1s, dense);
  d.nops = n_gates;

  if (int rc = ensure_attrs()) return rc;
  const size_t smem = 2 * (sizeof(double2) << n_tile);
  d.n_tiles = 1u << (n - n_tile);
  const unsigned grid = std::min<unsigned>(d.n_tiles, (unsigned)sm_count());
  tile_pass_kernel<<<grid, kPT, smem, stream>>>(reinterpret_cast<double2*>(sv), d);
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
