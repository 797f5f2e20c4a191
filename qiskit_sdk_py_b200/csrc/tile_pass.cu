// Cache-blocked gate application (K1-K5 of SURVEY 2.5) for sm_100a.
//
// One launch = one pass over the state: a CTA stages the 2^b amplitudes of a tile (the amplitudes
// that differ only in the b "tile qubits") in shared memory, applies a run of gates there and writes
// the tile back.  HBM traffic is exactly one read + one write of the vector per pass, however many
// gates the pass fuses.  Replaces runs of QasmSimulatorPy._add_unitary calls
// (/root/reference/qiskit/providers/basicaer/qasm_simulator.py:145-161), each of which is one
// out-of-place np.einsum over the whole vector.
//
// Kernel structure (tile_pass_kernel):
//   * persistent: 2-4 CTAs per SM (one warp per 2^8 amplitudes of the tile), each walking a contiguous
//     range of tiles with one tile buffer; while one CTA waits for HBM (cp.async load, write-back) the
//     others run their gates, so memory phases overlap FP64 / shared-memory phases across CTAs;
//   * the gates of a pass are grouped into STAGES.  In a stage every warp owns a sub-cube of 2^8
//     amplitudes of the tile (8 "inner" tile bits; the warp id supplies the others) and applies all
//     the stage's gates -- whose qubits are inner bits -- to it with only __syncwarp() between gates.
//     Warps therefore drift apart and the shared-memory traffic of one warp overlaps the FP64 tensor
//     work of another; block-wide barriers are needed only between stages;
//   * dense gates run on the FP64 tensor path (mma.sync m8n8k4 f64, SASS DMMA.8x8x4): a 4x4 complex
//     gate is the 8x8 real matrix [[Re,-Im],[Im,Re]]; one warp step multiplies 8 groups x 8 real
//     components (A operand, two k-slices) by the gate (B operand) and every lane receives one complex
//     output amplitude, i.e. one 16-byte shared-memory store.  (Measured on B200: DMMA 37 TF/s, DFMA
//     34 TF/s, no gain from mixing them -- tools/fp64_peak.cu -- but DMMA needs 8x fewer instructions
//     and 2 registers of gate matrix per lane instead of 64.)
//   * all index maps are linear over GF(2) (XOR masks), including a pass-specific bank swizzle.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace qsv {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

constexpr int kMaxOps = QSV_MAX_PASS_GATES;
constexpr int kMaxMat = 1536;  // doubles of matrix payload per pass (48 dense 2-qubit gates)
constexpr int kMaxPT = 512;    // threads of the pass kernel: 2^(b-3), at most 512 (one warp per 2^8 amplitudes)
constexpr int kWarps = kMaxPT / 32;
constexpr int kInner = 8;      // inner bits of a stage: 2^8 amplitudes per warp
constexpr int kMaxStages = kMaxOps;

struct __align__(4) PassOp {
  uint8_t kind, t0, t1, ngl;  // t0/t1: tile-local gate bits; ngl = log2(#groups in a warp's sub-cube)
  uint16_t moff;              // offset (doubles) of the matrix in PassDesc::mats
  uint16_t z0, z1;            // swz(1 << t0), swz(1 << t1)
  uint16_t col[6];            // swizzled masks of the free inner bits: col[0..2] vary inside a warp step
};

struct __align__(4) PassStage {
  uint8_t first_op, n_ops, n_inner, n_outer;
  uint8_t ebit[kInner];   // inner tile bits in element order (bit j of the in-warp element index)
  uint8_t obit[4];        // outer tile bits (bit j of the warp id)
  uint16_t ecol[kInner];  // swz(1 << ebit[j])
  uint16_t ocol[4];       // swz(1 << obit[j])
};

struct PassDesc {
  int32_t n, b, nops, nstages;
  uint32_t n_tiles;
  uint32_t pad2[3];
  uint8_t tb[16];   // ascending physical bit position of each tile-local bit
  uint8_t vec[16];  // bank vector of each tile-local bit (pass-specific XOR swizzle), vec[j] = 1<<j for j<3
  uint8_t src_of[16];  // write-back permutation: tile-local position j receives the content of bit src_of[j]
  PassStage stages[kMaxStages];
  PassOp ops[kMaxOps];
  double mats[kMaxMat];
};

// Pass-specific XOR swizzle of a tile-local amplitude index (16-byte slots; bank group = slot & 7).
// Tile bit j >= 3 toggles the bank bits vec[j]; three index bits whose vectors are linearly independent
// over GF(2) spread over all 8 bank groups, so the accesses of a quarter warp (LDS/STS.128) or of a
// half warp reading 8-byte halves (LDS.64) are conflict-free.  The host picks vec[] per pass.
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

// Per-op record staged in shared memory (one broadcast 16-byte load per op per warp).
struct __align__(16) OpRec {
  uint16_t z0, z1, c3, c4, c5;
  uint8_t kind, ngl, t0, t1;
  uint16_t moff;
};

// Persistent kernel, two 512-thread CTAs per SM, one 64 KiB tile buffer each: while one CTA waits for
// HBM (cp.async load of its next tile, write-back of the last) the other one runs its gates, so the
// memory phases of one overlap the FP64 / shared-memory phases of the other.  A CTA owns a contiguous
// range of tiles.
__global__ void __launch_bounds__(kMaxPT, 2)
tile_pass_kernel(double2* __restrict__ sv, const __grid_constant__ PassDesc d) {
  // dynamic shared memory: the tile, then the per-pass tables (sized by the pass, so that short passes
  // leave room for more resident CTAs)
  extern __shared__ double2 tile[];
  __shared__ u64 hi_off[8];
  __shared__ uint16_t swz_it[8];
  __shared__ uint16_t swz_it_st[8];
  const int nops_all = d.nops, nst_all = d.nstages;
  double (*frag)[64] = reinterpret_cast<double (*)[64]>(tile + (1u << d.b));   // [nops][64] B fragments
  OpRec* oprec = reinterpret_cast<OpRec*>(frag + nops_all);                     // [nops]
  uint32_t (*opl)[32] = reinterpret_cast<uint32_t (*)[32]>(oprec + nops_all);   // [nops][32] lane slots
  uint32_t (*st_warp)[kWarps] = reinterpret_cast<uint32_t (*)[kWarps]>(opl + nops_all);  // [nstages][16]
  uint32_t (*st_lane)[32] = reinterpret_cast<uint32_t (*)[32]>(st_warp + nst_all);       // [nstages][32]
  uint32_t (*st_it)[8] = reinterpret_cast<uint32_t (*)[8]>(st_lane + nst_all);           // [nstages][8]

  const uint32_t tid = threadIdx.x;
  const int b = d.b;
  const uint32_t tile_amps = 1u << b;
  const uint32_t dbg = d.pad2[0];  // QSV_DEBUG_FLAGS (profiling experiments): 1 = no write-back, 2 = no loads, 4 = no gates
  const int nops = d.nops, nstages = (dbg & 4u) ? 0 : d.nstages;

  // ---- per-kernel setup (amortised over all tiles of this CTA): every lane/warp dependent index of
  // the pass is tile-invariant, so it is digested once into shared-memory tables ----
  const uint32_t kPT = blockDim.x;
  const int kTidBits = 31 - __clz(kPT);
  const int lo_bits = b < kTidBits ? b : kTidBits;
  u64 lo_off = 0;
  for (int j = 0; j < lo_bits; j++) lo_off |= (u64)((tid >> j) & 1u) << d.tb[j];
  u64 tile_mask = 0;  // physical bit mask of the tile qubits
  for (int j = 0; j < b; j++) tile_mask |= 1ull << d.tb[j];
  if (tid < 8) {
    u64 off = 0;
    for (int j = kTidBits; j < b; j++) off |= (u64)((tid >> (j - kTidBits)) & 1u) << d.tb[j];
    hi_off[tid] = off;
    swz_it[tid] = (uint16_t)swz(d.vec, b, tid << kTidBits);
    uint32_t sst = 0;  // write-back source slot of loop index bits: position j reads the content of bit src_of[j]
    for (int j = kTidBits; j < b; j++)
      if ((tid >> (j - kTidBits)) & 1u) sst ^= swz(d.vec, b, 1u << d.src_of[j]);
    swz_it_st[tid] = (uint16_t)sst;
  }
  for (int idx = tid; idx < nops * 64; idx += kPT) {
    const int o = idx >> 6, e = idx & 63;
    const PassOp& op = d.ops[o];
    if (op.kind == QSV_GATE_DENSE2) {
      // B fragment of k-slice s: lane l holds Mreal[n = l >> 2][k = 4 s + (l & 3)]
      const int s = e >> 5, l = e & 31;
      const int nrow = l >> 2, kcol = 4 * s + (l & 3);
      const int r = nrow >> 1, c = kcol >> 1;
      const double* m = d.mats + op.moff + 2 * (4 * r + c);
      double v;
      if ((nrow & 1) == 0) v = (kcol & 1) == 0 ? m[0] : -m[1];
      else v = (kcol & 1) == 0 ? m[1] : m[0];
      frag[o][e] = v;
    }
    if (e < 32) {
      const uint32_t g = e >> 2, c = e & 3u;
      uint32_t s_g = 0;
      if (g & 1u) s_g ^= op.col[0];
      if (g & 2u) s_g ^= op.col[1];
      if (g & 4u) s_g ^= op.col[2];
      const uint32_t ld_x = (c & 2u) ? op.z0 : 0u;                                 // amplitude c>>1 of the k-slice
      const uint32_t st_x = ((c & 1u) ? op.z0 : 0u) ^ ((c & 2u) ? op.z1 : 0u);     // amplitude c of the group
      opl[o][e] = (s_g ^ ld_x) | ((s_g ^ st_x) << 16);
    }
    if (e == 32) {
      OpRec r;
      r.z0 = op.z0; r.z1 = op.z1; r.c3 = op.col[3]; r.c4 = op.col[4]; r.c5 = op.col[5];
      r.kind = op.kind; r.ngl = op.ngl; r.t0 = op.t0; r.t1 = op.t1;
      r.moff = op.moff;
      oprec[o] = r;
    }
  }
  for (int idx = tid; idx < nstages * 64; idx += kPT) {
    const int sidx = idx >> 6, e = idx & 63;
    const PassStage& st = d.stages[sidx];
    uint32_t li = 0, sl = 0;
    if (e < kWarps) {  // warp part
      for (int j = 0; j < st.n_outer; j++)
        if ((e >> j) & 1) { li |= 1u << st.obit[j]; sl ^= st.ocol[j]; }
      st_warp[sidx][e] = li | (sl << 16);
    } else if (e >= 24 && e < 32) {  // element iteration part (element bits 5, 6, 7)
      for (int j = 0; j < 3; j++)
        if ((((e - 24) >> j) & 1) && 5 + j < st.n_inner) { li |= 1u << st.ebit[5 + j]; sl ^= st.ecol[5 + j]; }
      st_it[sidx][e - 24] = li | (sl << 16);
    } else if (e >= 32) {  // lane part (element bits 0..4)
      for (int j = 0; j < 5 && j < st.n_inner; j++)
        if (((e - 32) >> j) & 1) { li |= 1u << st.ebit[j]; sl ^= st.ecol[j]; }
      st_lane[sidx][e - 32] = li | (sl << 16);
    }
  }
  const uint32_t s_tid = swz(d.vec, b, tid);
  uint32_t s_tid_st = 0;
  for (int j = 0; j < lo_bits; j++)
    if ((tid >> j) & 1u) s_tid_st ^= swz(d.vec, b, 1u << d.src_of[j]);
  const bool permuted = d.pad2[1] != 0;
  const uint32_t iters = tile_amps > kPT ? tile_amps / kPT : 1;
  const bool active = tid < tile_amps;
  const bool ld_on = active && !(dbg & 2u), st_on = active && !(dbg & 1u);
  const uint32_t lane = tid & 31u, wid = tid >> 5;
  const uint32_t g = lane >> 2;
  const char* tb8 = reinterpret_cast<const char*>(tile) + (lane & 1u) * 8u;
  __syncthreads();

  // contiguous range of tiles for this CTA; the tile base walks the non-tile bits: next = ((base | M) + 1) & ~M
  const u64 n_tiles = d.n_tiles;
  const u64 per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const u64 t_begin = (u64)blockIdx.x * per;
  const u64 t_end = t_begin + per < n_tiles ? t_begin + per : n_tiles;
  u64 base = t_begin < n_tiles ? tile_base(d, t_begin) : 0;
  for (u64 t = t_begin; t < t_end; t++, base = ((base | tile_mask) + 1) & ~tile_mask) {
    // ---- load: one HBM read of the tile ----
    if (ld_on) {
      const double2* gp = sv + base + lo_off;
#pragma unroll 8
      for (uint32_t it = 0; it < iters; it++) cp_async16(tile + (s_tid ^ swz_it[it]), gp + hi_off[it]);
    }
    cp_async_commit();
    cp_async_wait<0>();

    int o = 0;
    for (int sidx = 0; sidx < nstages; sidx++) {
      __syncthreads();  // stage boundary: warps re-partition the tile (the first one also publishes the loads)
      const uint32_t n_inner = d.stages[sidx].n_inner, n_outer = d.stages[sidx].n_outer;
      const int o_end = o + d.stages[sidx].n_ops;
      const bool warp_on = wid < (1u << n_outer);  // small tiles: fewer sub-cubes than warps
      const uint32_t wv = st_warp[sidx][wid];
      const uint32_t s_w = wv >> 16;
      bool first = true;
      for (; o < o_end; o++) {
        if (!warp_on) continue;
        const OpRec rec = oprec[o];
        if (!first) __syncwarp();
        first = false;

        if (rec.kind == QSV_GATE_DENSE2) {
          // ---- FP64 tensor path: 8 groups per warp step, up to 8 steps per sub-cube ----
          const double b0 = frag[o][lane], b1 = frag[o][32 + lane];
          const uint32_t v = opl[o][lane];
          const uint32_t a_ld = (s_w ^ (v & 0xffffu)) << 4;  // byte offsets of 16-byte slots
          const uint32_t a_st = (s_w ^ (v >> 16)) << 4;
          const uint32_t z1b = (uint32_t)rec.z1 << 4;
          const uint32_t c3 = (uint32_t)rec.c3 << 4, c4 = (uint32_t)rec.c4 << 4, c5 = (uint32_t)rec.c5 << 4;
          if (rec.ngl == 6u) {
#pragma unroll
            for (int hb = 0; hb < 2; hb++) {
              const uint32_t xh = hb ? c5 : 0u;
              const uint32_t xo[4] = {xh, xh ^ c3, xh ^ c4, xh ^ c3 ^ c4};
              double x0[4], x1[4];
#pragma unroll
              for (int it = 0; it < 4; it++) {
                const uint32_t a = a_ld ^ xo[it];
                x0[it] = *reinterpret_cast<const double*>(tb8 + a);
                x1[it] = *reinterpret_cast<const double*>(tb8 + (a ^ z1b));
              }
#pragma unroll
              for (int it = 0; it < 4; it++) {
                double d0 = 0.0, d1 = 0.0;
                dmma884(d0, d1, x0[it], b0);
                dmma884(d0, d1, x1[it], b1);
                *reinterpret_cast<double2*>(reinterpret_cast<char*>(tile) + (a_st ^ xo[it])) = make_double2(d0, d1);
              }
            }
          } else {
            const uint32_t ng = 1u << rec.ngl;
            const uint32_t n_it = ng > 8u ? (ng >> 3) : 1u;
            for (uint32_t it = 0; it < n_it; it++) {
              uint32_t xo = 0;
              if (it & 1u) xo ^= c3;
              if (it & 2u) xo ^= c4;
              if (it & 4u) xo ^= c5;
              const bool valid = (g | (it << 3)) < ng;
              double x0 = 0.0, x1 = 0.0;
              if (valid) {
                const uint32_t a = a_ld ^ xo;
                x0 = *reinterpret_cast<const double*>(tb8 + a);
                x1 = *reinterpret_cast<const double*>(tb8 + (a ^ z1b));
              }
              double d0 = 0.0, d1 = 0.0;
              dmma884(d0, d1, x0, b0);
              dmma884(d0, d1, x1, b1);
              if (valid) *reinterpret_cast<double2*>(reinterpret_cast<char*>(tile) + (a_st ^ xo)) = make_double2(d0, d1);
            }
          }
        } else {
          // ---- element-wise kinds: each lane owns elements e = lane | it << 5 of the sub-cube ----
          const uint32_t n_el = 1u << n_inner;
          const uint32_t lv = wv ^ st_lane[sidx][lane];  // li (low 16) and slot (high 16): disjoint bits, XOR == OR
          const double* m = d.mats + rec.moff;
          for (uint32_t it = 0; it < 8u; it++) {
            if ((lane | (it << 5)) >= n_el) break;
            const uint32_t ev = lv ^ st_it[sidx][it];
            const uint32_t li = ev & 0xffffu, s = ev >> 16;
            if (rec.kind == QSV_GATE_DIAG1) {
              const double* f = m + 2 * ((li >> rec.t0) & 1u);
              tile[s] = cmul(make_double2(f[0], f[1]), tile[s]);
            } else if (rec.kind == QSV_GATE_DIAG2) {
              const double* f = m + 2 * (((li >> rec.t0) & 1u) | (((li >> rec.t1) & 1u) << 1));
              tile[s] = cmul(make_double2(f[0], f[1]), tile[s]);
            } else if (rec.kind == QSV_GATE_CX) {
              // control t0 set, target t1 clear: swap with the target-set partner (same sub-cube)
              if (((li >> rec.t0) & 1u) && !((li >> rec.t1) & 1u)) {
                const uint32_t sp = s ^ rec.z1;
                const double2 a = tile[s], p = tile[sp];
                tile[s] = p;
                tile[sp] = a;
              }
            } else if (rec.kind == QSV_GATE_DENSE1) {
              // scalar 2x2 (only used when the tile has a single qubit): element with t0 clear owns the pair
              if (!((li >> rec.t0) & 1u)) {
                const uint32_t sp = s ^ rec.z0;
                const double2 a0 = tile[s], a1 = tile[sp];
                double2 r0 = cmul(make_double2(m[0], m[1]), a0), r1 = cmul(make_double2(m[4], m[5]), a0);
                cfma(r0, make_double2(m[2], m[3]), a1);
                cfma(r1, make_double2(m[6], m[7]), a1);
                tile[s] = r0;
                tile[sp] = r1;
              }
            }
          }
        }
      }
    }
    __syncthreads();

    // ---- write-back: one HBM write of the tile, optionally with the tile qubits permuted (the content
    // of tile bit src_of[j] lands on position j: a free qubit remap).  Without a permutation every thread
    // reads exactly the slots its own cp.async of the next tile will overwrite, so no barrier is needed ----
    if (st_on) {
      double2* gp = sv + base + lo_off;
#pragma unroll 8
      for (uint32_t it = 0; it < iters; it++) gp[hi_off[it]] = tile[s_tid_st ^ swz_it_st[it]];
    }
    if (permuted) __syncthreads();
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

// ------------------------------------------------------------------------------------------------
// Host-side lowering of a pass: stages, bank vectors, XOR masks.
// ------------------------------------------------------------------------------------------------
struct HostOp {
  int kind, t0, t1;
  uint32_t touch;  // every tile bit the gate reads (ordering)
  uint32_t need;   // tile bits that must be inner bits of the gate's stage
  int moff;
};

static inline bool independent3(int a, int b, int c) {
  // three vectors of GF(2)^3 are independent iff none is zero, all differ, and c != a ^ b
  return a && b && c && a != b && a != c && b != c && c != (a ^ b);
}

static inline int popc(uint32_t x) { return __builtin_popcount(x); }

// Greedy staging, same commutation rule as the planner: a gate may be hoisted into the current stage
// only over gates that touch none of its bits.
static int build_stages(const HostOp* ops, int n_ops, int b, int* order, PassStage* stages, uint32_t* inner_masks) {
  const int n_inner = std::min(b, kInner);
  bool done[kMaxOps] = {false};
  int n_done = 0, n_stages = 0, pos = 0;
  while (n_done < n_ops) {
    uint32_t inner = 0, blocked = 0;
    const int first = pos;
    for (int i = 0; i < n_ops; i++) {
      if (done[i]) continue;
      const HostOp& op = ops[i];
      if (!(op.touch & blocked) && popc(inner | op.need) <= n_inner) {
        inner |= op.need;
        done[i] = true;
        n_done++;
        order[pos++] = i;
      } else {
        blocked |= op.touch;
      }
    }
    // pad the inner set with the lowest unused tile bits
    for (int j = 0; j < b && popc(inner) < n_inner; j++)
      if (!((inner >> j) & 1u)) inner |= 1u << j;
    PassStage& st = stages[n_stages];
    memset(&st, 0, sizeof(st));
    st.first_op = (uint8_t)first;
    st.n_ops = (uint8_t)(pos - first);
    st.n_inner = (uint8_t)n_inner;
    st.n_outer = (uint8_t)(b - n_inner);
    inner_masks[n_stages] = inner;
    n_stages++;
  }
  return n_stages;
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
                                  (int)((sizeof(double2) << QSV_MAX_TILE_QUBITS) + 48 * 1024)));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QSV_CUDA(cudaFuncSetAttribute(dense_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * (sizeof(double2) << 12) + 4096 * sizeof(uint32_t))));
    g_attr_set[dev] = true;
  }
  return 0;
}

static int launch_pass(double* sv, int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                       int n_gates, cudaStream_t stream, const int32_t* dest_qubits = nullptr) {
  if (!sv || n < 1 || n > 40) return fail(QSV_EINVAL, "qsv_apply_pass: bad state / n_qubits");
  if (n_tile < 1 || n_tile > QSV_MAX_TILE_QUBITS || n_tile > n)
    return fail(QSV_ERANGE, "qsv_apply_pass: n_tile must be in [1, min(n_qubits, 12)]");
  if (n_gates < 0 || n_gates > kMaxOps) return fail(QSV_ERANGE, "qsv_apply_pass: too many gates");
  if (n - n_tile > 31) return fail(QSV_ERANGE, "qsv_apply_pass: grid too large");

  static thread_local PassDesc d;
  memset(&d, 0, sizeof(d));
  const int b = n_tile;
  d.n = n;
  d.b = b;
  int sorted[16];
  for (int j = 0; j < b; j++) {
    if (tile_qubits[j] < 0 || tile_qubits[j] >= n) return fail(QSV_EINVAL, "qsv_apply_pass: tile qubit out of range");
    sorted[j] = tile_qubits[j];
  }
  std::sort(sorted, sorted + b);
  for (int j = 0; j + 1 < b; j++)
    if (sorted[j] == sorted[j + 1]) return fail(QSV_EINVAL, "qsv_apply_pass: duplicate tile qubit");
  int local_of[64];
  for (int q = 0; q < 64; q++) local_of[q] = -1;
  for (int j = 0; j < b; j++) {
    d.tb[j] = (uint8_t)sorted[j];
    local_of[sorted[j]] = j;
  }
  // write-back permutation: content of tile_qubits[i] goes to dest_qubits[i]
  for (int j = 0; j < 16; j++) d.src_of[j] = (uint8_t)j;
  if (dest_qubits) {
    bool seen[16] = {false};
    for (int i = 0; i < b; i++) {
      const int dq = dest_qubits[i];
      if (dq < 0 || dq >= n || local_of[dq] < 0 || seen[local_of[dq]])
        return fail(QSV_EINVAL, "qsv_apply_pass_remap: dest_qubits must be a permutation of tile_qubits");
      seen[local_of[dq]] = true;
      d.src_of[local_of[dq]] = (uint8_t)local_of[tile_qubits[i]];
    }
    for (int j = 0; j < b; j++)
      if (d.src_of[j] != j) d.pad2[1] = 1;
  }

  // ---- gates -> host ops (tile-local bits, matrices) ----
  HostOp hops[kMaxOps];
  double mats[kMaxMat];
  int moff = 0;
  for (int i = 0; i < n_gates; i++) {
    const qsv_gate_t& gsrc = gates[i];
    HostOp& op = hops[i];
    int nq, ndbl;
    switch (gsrc.kind) {
      case QSV_GATE_DENSE1: nq = 1; ndbl = b >= 2 ? 32 : 8; break;
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
    op.kind = gsrc.kind;
    op.t0 = local_of[gsrc.q0];
    op.t1 = nq == 2 ? local_of[gsrc.q1] : -1;
    op.moff = moff;
    op.touch = 1u << op.t0;
    if (nq == 2) op.touch |= 1u << op.t1;
    switch (gsrc.kind) {
      case QSV_GATE_DENSE2: op.need = op.touch; break;
      case QSV_GATE_DENSE1: op.need = 1u << op.t0; break;
      case QSV_GATE_CX: op.need = 1u << op.t1; break;  // target; the control may be any tile bit
      default: op.need = 0; break;                      // diagonal gates need no particular partition
    }
    if (gsrc.kind == QSV_GATE_DENSE1 && b >= 2) {
      memset(mats + moff, 0, sizeof(double) * 32);  // widened once the partner bit is known
      memcpy(mats + moff, gsrc.m, sizeof(double) * 8);
    } else {
      memcpy(mats + moff, gsrc.m, sizeof(double) * ndbl);
    }
    moff += ndbl;
  }

  // ---- stages ----
  int order[kMaxOps];
  uint32_t inner_masks[kMaxStages];
  const int n_stages = build_stages(hops, n_gates, b, order, d.stages, inner_masks);
  d.nstages = n_stages;
  d.nops = n_gates;

  // ops in stage order; 1-qubit dense gates are widened to I (x) M on a partner inner bit so that they
  // run on the same FP64 tensor path as 2-qubit gates
  int stage_of[kMaxOps];
  for (int s = 0; s < n_stages; s++)
    for (int k = d.stages[s].first_op; k < d.stages[s].first_op + d.stages[s].n_ops; k++) stage_of[k] = s;
  HostOp sops[kMaxOps];
  for (int k = 0; k < n_gates; k++) {
    sops[k] = hops[order[k]];
    HostOp& op = sops[k];
    if (op.kind == QSV_GATE_DENSE1 && b >= 2) {
      const uint32_t inner = inner_masks[stage_of[k]];
      int partner = -1;
      for (int j = 0; j < b && partner < 0; j++)
        if (((inner >> j) & 1u) && j != op.t0) partner = j;
      double m2[8];
      memcpy(m2, mats + op.moff, sizeof(m2));
      double* m4 = mats + op.moff;
      for (int r = 0; r < 4; r++)
        for (int cc = 0; cc < 4; cc++) {
          const bool same = (r >> 1) == (cc >> 1);
          m4[2 * (4 * r + cc)] = same ? m2[2 * (2 * (r & 1) + (cc & 1))] : 0.0;
          m4[2 * (4 * r + cc) + 1] = same ? m2[2 * (2 * (r & 1) + (cc & 1)) + 1] : 0.0;
        }
      op.kind = QSV_GATE_DENSE2;
      op.t1 = partner;
    }
  }
  memcpy(d.mats, mats, sizeof(double) * moff);

  // ---- bank vectors: the two bits of every dense gate should differ ----
  for (int j = 0; j < 16; j++) d.vec[j] = (uint8_t)(1u << (j % 3));
  {
    auto conflicts = [&]() {
      int cnt = 0;
      for (int k = 0; k < n_gates; k++)
        if (sops[k].kind == QSV_GATE_DENSE2 && d.vec[sops[k].t0] == d.vec[sops[k].t1]) cnt++;
      // write-back reads: the 8 lanes of a quarter warp read the bits that land on positions 0..2
      if (b >= 3 && !independent3(d.vec[d.src_of[0]] & 7, d.vec[d.src_of[1]] & 7, d.vec[d.src_of[2]] & 7)) cnt += 2;
      return cnt;
    };
    int best = conflicts();
    // first make the write-back lane bits independent
    for (int j = 0; j < 3 && j < b && best > 0; j++) {
      const int bit = d.src_of[j];
      if (bit < 3) continue;
      uint8_t best_v = d.vec[bit];
      for (uint8_t v = 1; v < 8; v++) {
        d.vec[bit] = v;
        const int cnt = conflicts();
        if (cnt < best) {
          best = cnt;
          best_v = v;
        }
      }
      d.vec[bit] = best_v;
    }
    for (int round = 0; round < 4 && best > 0; round++) {
      for (int k = 0; k < n_gates && best > 0; k++) {
        if (sops[k].kind != QSV_GATE_DENSE2 || d.vec[sops[k].t0] != d.vec[sops[k].t1]) continue;
        const int cand[2] = {sops[k].t1, sops[k].t0};
        for (int ci = 0; ci < 2 && d.vec[sops[k].t0] == d.vec[sops[k].t1]; ci++) {
          const int bit = cand[ci];
          if (bit < 3) continue;
          uint8_t best_v = d.vec[bit];
          for (uint8_t v = 1; v < 8; v++) {
            d.vec[bit] = v;
            const int cnt = conflicts();
            if (cnt < best) {
              best = cnt;
              best_v = v;
            }
          }
          d.vec[bit] = best_v;
        }
      }
    }
  }
  auto bvec = [&](int bit) { return (int)(d.vec[bit] & 7); };
  auto sw1 = [&](int bit) { return (uint16_t)swz(d.vec, b, 1u << bit); };

  // ---- per-stage element order and per-op group order ----
  for (int s = 0; s < n_stages; s++) {
    PassStage& st = d.stages[s];
    const uint32_t inner = inner_masks[s];
    int ib[16], ni = 0, ob[16], no = 0;
    for (int j = 0; j < b; j++) {
      if ((inner >> j) & 1u) ib[ni++] = j;
      else ob[no++] = j;
    }
    // element order: three inner bits with independent bank vectors first (quarter-warp accesses)
    int lead[3] = {-1, -1, -1};
    for (int i0 = 0; i0 < ni && lead[2] < 0; i0++)
      for (int i1 = i0 + 1; i1 < ni && lead[2] < 0; i1++)
        for (int i2 = i1 + 1; i2 < ni && lead[2] < 0; i2++)
          if (independent3(bvec(ib[i0]), bvec(ib[i1]), bvec(ib[i2]))) {
            lead[0] = ib[i0];
            lead[1] = ib[i1];
            lead[2] = ib[i2];
          }
    int eo[16], ne = 0;
    for (int k = 0; k < 3; k++)
      if (lead[k] >= 0) eo[ne++] = lead[k];
    for (int i = 0; i < ni; i++) {
      bool dup = false;
      for (int q = 0; q < ne; q++) dup |= (eo[q] == ib[i]);
      if (!dup) eo[ne++] = ib[i];
    }
    for (int j = 0; j < kInner; j++) {
      st.ebit[j] = (uint8_t)(j < ni ? eo[j] : 0);
      st.ecol[j] = j < ni ? sw1(eo[j]) : 0;
    }
    for (int j = 0; j < 4; j++) {
      st.obit[j] = (uint8_t)(j < no ? ob[j] : 0);
      st.ocol[j] = j < no ? sw1(ob[j]) : 0;
    }
    for (int k = st.first_op; k < st.first_op + st.n_ops; k++) {
      const HostOp& h = sops[k];
      PassOp& op = d.ops[k];
      op.kind = (uint8_t)h.kind;
      op.t0 = (uint8_t)h.t0;
      op.t1 = (uint8_t)(h.t1 < 0 ? 0 : h.t1);
      op.moff = (uint16_t)h.moff;
      op.z0 = sw1(h.t0);
      op.z1 = h.t1 < 0 ? 0 : sw1(h.t1);
      if (h.kind != QSV_GATE_DENSE2) continue;
      // free inner bits of the gate: f0 with {f0,t0,t1} independent (STS.128: 2 groups x 4 amplitudes per
      // quarter warp), f1 with {f0,f1,t0} independent (LDS.64: 4 groups x 2 amplitudes per half warp)
      int fb[16], nf = 0;
      for (int i = 0; i < ni; i++)
        if (ib[i] != h.t0 && ib[i] != h.t1) fb[nf++] = ib[i];
      const int v0 = bvec(h.t0), v1 = bvec(h.t1);
      int f0 = -1, f1 = -1;
      for (int i = 0; i < nf && f0 < 0; i++)
        if (independent3(bvec(fb[i]), v0, v1)) f0 = fb[i];
      for (int i = 0; i < nf && f0 < 0; i++)
        if (bvec(fb[i]) != v0) f0 = fb[i];
      if (f0 >= 0)
        for (int i = 0; i < nf && f1 < 0; i++)
          if (fb[i] != f0 && independent3(bvec(f0), bvec(fb[i]), v0)) f1 = fb[i];
      int fo[16], nfo = 0;
      if (f0 >= 0) fo[nfo++] = f0;
      if (f1 >= 0) fo[nfo++] = f1;
      for (int i = 0; i < nf; i++)
        if (fb[i] != f0 && fb[i] != f1) fo[nfo++] = fb[i];
      op.ngl = (uint8_t)nf;
      for (int j = 0; j < 6; j++) op.col[j] = j < nf ? sw1(fo[j]) : 0;
    }
  }

  if (int rc = ensure_attrs()) return rc;
  const size_t smem = (sizeof(double2) << b) + (size_t)n_gates * (64 * sizeof(double) + sizeof(OpRec) + 32 * 4) +
                      (size_t)n_stages * ((kWarps + 32 + 8) * 4);
  d.n_tiles = 1u << (n - b);
  if (const char* e = getenv("QSV_DEBUG_FLAGS")) d.pad2[0] = (uint32_t)atoi(e);
  // one warp per 2^8 amplitudes: 512 threads for b = 12, 256 for b = 11, ...; at least one warp
  const int threads = std::max(32, std::min(kMaxPT, 1 << std::max(0, b - 3)));
  int per_sm = 1;
  QSV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_pass_kernel, threads, smem));
  if (per_sm < 1) return fail(QSV_ERANGE, "qsv_apply_pass: pass does not fit in shared memory");
  const unsigned grid = std::min<unsigned>(d.n_tiles, (unsigned)per_sm * (unsigned)sm_count());
  tile_pass_kernel<<<grid, threads, smem, stream>>>(reinterpret_cast<double2*>(sv), d);
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

int qsv_apply_pass_remap(double* sv, int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                         int n_gates, const int32_t* dest_qubits, void* stream) {
  if (!tile_qubits || (n_gates > 0 && !gates)) return fail(QSV_EINVAL, "qsv_apply_pass_remap: null pointer");
  return launch_pass(sv, n, tile_qubits, n_tile, gates, n_gates, (cudaStream_t)stream, dest_qubits);
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
