// Cache-blocked gate application (K1-K5 of SURVEY 2.5) for sm_100a.
//
// One launch = one pass over the state: a CTA stages the 2^b amplitudes of a tile (the amplitudes
// that differ only in the b "tile qubits") in shared memory, applies a run of gates there and writes
// the tile back.  HBM traffic is exactly one read + one write of the vector per pass, however many
// gates the pass fuses.  Replaces runs of QasmSimulatorPy._add_unitary calls
// (/root/reference/qiskit/providers/basicaer/qasm_simulator.py:145-161), each of which is one
// out-of-place np.einsum over the whole vector.
//
// Kernel structure (tile_pass_kernel, one body, two launch shapes):
//   * warp-group shape (tile_pass_kernel<768,1,3>; 11-qubit tiles, >= 2048 tiles, unpermuted write-back): one
//     persistent CTA per SM holds THREE independent groups of 8 warps (named barriers), each with TWO tile
//     buffers, all sharing the pass's read-only tables (that is what lets six 35 KiB buffers fit in 227 KiB).
//     A group requests tile t+1 into its other buffer as soon as tile t has landed, and the write-back of
//     tile t-1 drains while tile t is computed; a buffer is recycled without a barrier because the lane that
//     wrote a piece back is the lane that loads the next tile's same piece (it waits for its own bulk
//     async-group).  One elected lane per warp issues the warp's TMA pieces from uniform registers (offsets
//     in the kernel parameters), so there is no per-lane waterfall loop around the bulk-copy instruction;
//   * single-buffer shape (every other pass): 2-4 persistent CTAs per SM (one warp per 2^8 amplitudes of the
//     tile), each walking a contiguous range of tiles with one tile buffer; while one CTA waits for HBM (TMA
//     bulk load, write-back) the others run their gates;
//   * the gates of a pass are grouped into STAGES.  In a stage every warp owns a sub-cube of 2^8
//     amplitudes of the tile (8 "inner" tile bits; the warp id supplies the others) and applies all
//     the stage's gates -- whose qubits are inner bits -- to it with only __syncwarp() between gates.
//     Warps therefore drift apart and the shared-memory traffic of one warp overlaps the FP64 tensor
//     work of another; group-wide barriers are needed only between stages;
//   * dense gates run on the FP64 tensor path (mma.sync m8n8k4 f64, SASS DMMA.8x8x4): a 4x4 complex
//     gate is the 8x8 real matrix [[Re,-Im],[Im,Re]]; one warp step multiplies 8 groups x 8 real
//     components (A operand, two k-slices) by the gate (B operand) and every lane receives one complex
//     output amplitude, i.e. one 16-byte shared-memory store.  (Measured on B200: DMMA 37 TF/s, DFMA
//     34 TF/s, no gain from mixing them -- tools/fp64_peak.cu -- but DMMA needs 8x fewer instructions
//     and 2 registers of gate matrix per lane instead of 64.)
//   * on the host, pairs of 2-qubit gates that share a qubit are fused into 3-qubit blocks (8x8 complex):
//     one shared-memory round trip instead of two, and the block is applied in the 3-multiplication form of
//     the complex product (P(x+y), (P+Q)y, (Q-P)x: three real 8x8 products = 6 DMMA + 2 DADD per 64 amplitudes
//     instead of 8 DMMA for the 16x16 real form);
//   * diagonal gates never touch the tensor path: runs of them are folded on the host into lookup tables
//     over the tile bits they touch (one look-up + one complex multiply per amplitude per run), and
//     diagonal gates on qubits OUTSIDE the tile ride along as the pass's "tail" (TailDesc): outside the
//     tile a qubit is a constant bit of the tile's base address;
//   * every index map is a sum of per-bit increments (padded shared-memory layout, see pad_of_bit).
//
// Debug flags (qsv_debug_set_flags; the debug library also reads QSV_DEBUG_FLAGS): 1 no write-back, 2 no loads,
// 4 no gates (these three only in the debug library), 8 no gate-pair fusion, 16 no warp-group variant, 32 no L2
// prefetch, 64 no diagonal-run tables, 512 warp-group variant of round 1 (the compute warps issue their own bulk
// copies; then also 128 = request the next tile after the first op instead of before it, 256 = every thread polls
// the load mbarrier instead of one polling warp + a group barrier).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace qsv {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

// Kernel-variant / profiling knobs (bit values: see the top of this file).  They are set explicitly through
// qsv_debug_set_flags(); the production library accepts only the bits that select an equivalent kernel variant
// (results unchanged) and never reads the environment, so a stray variable cannot change results.  The debug
// library (libqsv_dbg.so, -DQSV_DEBUG_BUILD, used by tools/) also takes QSV_DEBUG_FLAGS and the bits that switch
// loads / stores / gates off.  2048 = full group barrier at every stage boundary (no partial barriers).
[[maybe_unused]] constexpr uint32_t kResultChangingFlags = 1u | 2u | 4u;
static uint32_t g_debug_flags = 0;
static bool g_debug_flags_set = false;
static uint32_t debug_flags() {
  if (g_debug_flags_set) return g_debug_flags;
#ifdef QSV_DEBUG_BUILD
  static const uint32_t env_flags = [] {
    const char* e = getenv("QSV_DEBUG_FLAGS");
    return e ? (uint32_t)atoi(e) : 0u;
  }();
  return env_flags;
#else
  return 0u;
#endif
}

#ifdef QSV_PROFILE
// per (CTA, group, role) phase totals in clocks; roles 0..7 = the group's compute warps, 8 = its producer warp
constexpr int kProfSlots = 20;  // 0-3 phases; 4 stage-barrier waits; 5 + o: op o of the pass (o < 15)
constexpr int kProfRoles = 9;
__device__ unsigned long long g_prof[160 * 3 * kProfRoles * kProfSlots];
#define PROF_DECL unsigned long long prof_acc[kProfSlots] = {0}; long long prof_t0 = clock64();
#define PROF_MARK(slot) do { const long long t1_ = clock64(); prof_acc[slot] += (unsigned long long)(t1_ - prof_t0); prof_t0 = t1_; } while (0)
#define PROF_OP_BEGIN long long prof_t2 = clock64();
#define PROF_OP_MARK(slot) do { const long long t3_ = clock64(); prof_acc[(slot) < kProfSlots ? (slot) : kProfSlots - 1] += (unsigned long long)(t3_ - prof_t2); prof_t2 = t3_; } while (0)
#define PROF_FLUSH(role) do { if ((threadIdx.x & 31u) == 0 && blockIdx.x < 160) \
    for (int q_ = 0; q_ < kProfSlots; q_++) g_prof[((blockIdx.x * 3 + grp) * kProfRoles + (role)) * kProfSlots + q_] = prof_acc[q_]; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(slot) do { } while (0)
#define PROF_OP_BEGIN
#define PROF_OP_MARK(slot) do { } while (0)
#define PROF_FLUSH(role) do { } while (0)
#endif

constexpr int kMaxOps = QSV_MAX_PASS_GATES;
constexpr int kMaxIn = 1536;   // doubles of matrix payload per pass as passed in (48 dense 2-qubit gates)
constexpr int kMaxMat = 3072;  // ... after pair fusion (24 fused 3-qubit blocks of 8x8 complex)
constexpr int kGateDense3 = 6; // internal kind: two 2-qubit gates sharing a qubit, fused by launch_pass
constexpr int kMaxPT = 512;    // threads of the pass kernel: 2^(b-3), at most 512 (one warp per 2^8 amplitudes)
constexpr int kWarps = kMaxPT / 32;
constexpr int kInner = 8;      // inner bits of a stage: 2^8 amplitudes per warp
constexpr int kMaxStages = kMaxOps;

struct __align__(4) PassOp {
  uint8_t kind, t0, t1, ngl;  // t0/t1: tile-local gate bits; ngl = log2(#groups in a warp's sub-cube)
  uint16_t moff;              // offset (doubles) of the matrix in PassDesc::mats
  uint16_t foff;              // offset (units of 64 doubles) of the B fragments in shared memory
  uint16_t z0, z1, z2;        // slot increments of the gate bits: cj[t0], cj[t1] (, cj[t2] for fused blocks)
  uint16_t col[6];            // slot increments of the free inner bits: col[0..2] vary inside a warp step
};

struct __align__(4) PassStage {
  uint8_t first_op, n_ops, n_inner, n_outer;
  uint8_t ebit[kInner];   // inner tile bits in element order (bit j of the in-warp element index)
  uint8_t obit[4];        // outer tile bits (bit j of the warp id)
  uint8_t sync_mask;      // warp-id bits whose outer tile bit is the same as in the previous stage: at the stage
  uint8_t sync_pad;       // boundary only the warps that agree on them exchange data (partial barrier); 0 = all warps
  uint16_t ecol[kInner];  // cj[ebit[j]]
  uint16_t ocol[4];       // cj[obit[j]]
};

constexpr int kMaxRuns = 8;      // diagonal runs of a pass that get a lookup table (2^nbits complex of shared memory each)
constexpr int kRunBits = 8;      // ... over at most 8 tile bits
constexpr int kMaxRunEntries = 512;  // all tables of a pass together: 8 KiB of shared memory

// A run of consecutive diagonal gates of one stage, folded on the host into ONE table over the tile bits the
// run touches: element li is multiplied by tab[x(li)], x = the run's bits of li gathered.  x is a sum of a
// warp part, a lane part and an iteration part, like every other index of the kernel.
struct DiagRun {
  uint8_t first_op, n_ops, nbits, pad;
  uint16_t moff;                      // table (2^nbits complex) in PassDesc::mats
  uint16_t toff;                      // ... and its offset (complex entries) in the shared-memory table area
  uint16_t xw[kWarps], xl[32], xit[8];
};

constexpr int kMaxTail = 192;    // diagonal tail gates of a pass (gates with qubits outside the tile)

// Diagonal TAIL of a pass: diagonal gates with at least one qubit outside the tile, applied after the pass's
// other gates.  Outside the tile a qubit is constant per tile (a bit of the tile's base address), so per
// tile every tail gate collapses to a factor pair on ONE tile bit (or to a scalar), the factors of a tile
// bit multiply up, and the whole tail -- however many gates -- becomes one or two small tables over the tile
// bits, rebuilt per tile, i.e. one or two complex multiplies per amplitude and NO extra pass over HBM.  (A
// controlled-phase ladder of a QFT touches every qubit of the register: without this each rung needs its
// two qubits resident, i.e. a pass of its own.)
struct TailDesc {
  int32_t n;                          // entries (0 = no tail)
  uint8_t start[16];                  // entries [start[j], start[j+1]) multiply tile bit j; index b = scalars
  uint8_t pk[kMaxTail];               // physical bit selecting the factor row (0xff: none)
  uint8_t pk2[kMaxTail];              // scalars only: second physical bit (0xff: none)
  uint8_t na, nb;                     // tile bits with factors, split in two table groups of at most 6 bits
  uint8_t abit[6], bbit[6];
  uint16_t xw[2][kWarps], xl[2][32], xit[2][8];  // table index parts in the last stage's element order
};

struct PassDesc {
  int32_t n, b, nops, nstages;
  int32_t nfrag;        // B-fragment units (64 doubles) of all ops: 1 per 2-qubit gate, 4 per fused block
  int32_t nruns;        // tabulated diagonal runs
  int32_t run_entries;  // their tables together (complex entries)
  uint32_t n_tiles;
  int32_t n_spread;     // entries of tbs: the tile bits plus the FIXED bits of a sub-block pass (qsv_apply_pass_sub)
  uint64_t fixed_mask;  // physical bits that keep the value the caller folded into the state pointer
  uint32_t flags;       // QSV_DEBUG_FLAGS (profiling experiments)
  uint32_t pb_ld;       // log2(amplitudes per bulk-load piece): contiguous low tile bits, at most 4 (256 B)
  uint32_t pb_st;       // same for the write-back (smaller if the write-back permutes low bits)
  uint8_t tb[16];       // physical bit of tile-local position j (positions 0..3 = the four lowest tile qubits)
  uint8_t tbs[16];      // the tile's (and the fixed) physical bits in ascending order (tile base computation)
  uint16_t cj[16];      // shared-memory slot increment of tile-local bit j (see slot layout below)
  uint8_t src_of[16];   // write-back permutation: tile-local position j receives the content of bit src_of[j]
  // warp-group variant: offsets of bulk-copy piece p (its index bits = the tile-local bits pb_ld..b-1), in
  // BYTES from the tile base in HBM / from the start of the tile buffer.  They live in the kernel parameters
  // so that the issuing lane's address arithmetic stays in uniform registers (no per-lane waterfall loop
  // around the bulk-copy instruction).
  uint64_t pc_g[256];
  uint16_t pc_s[256];
  DiagRun runs[kMaxRuns];
  TailDesc tail;
  PassStage stages[kMaxStages];
  PassOp ops[kMaxOps];
  double mats[kMaxMat];
};

static_assert(sizeof(PassDesc) + 16 <= 32764, "PassDesc travels as a __grid_constant__ kernel parameter");

// Shared-memory slot layout (16-byte slots; bank group = slot & 7).  Amplitude li of the tile lives at
//     slot(li) = sum_j bit_j(li) * cj[j],   cj[j] = 2^j for j < 4,   cj[j] = 2^j + pad_j for j >= 4,
// i.e. the 256-byte pieces (16 amplitudes, low four bits) stay contiguous -- they are moved by single
// TMA bulk copies, and the TMA request rate is what limits the copy phases, so pieces are as large as the
// contiguous runs of the tile in HBM -- and padding slots are inserted between pieces.  pad_j is
// super-increasing (1,2,4,9,18,36,73,146,292: pieces never overlap) with pad_j mod 8 = 2^((j-1) mod 3), so
// bit j >= 4 shifts the bank group by 2^((j-1) mod 3) and bits 0,1,2 by 1,2,4: any three index bits of
// pairwise different class spread over all 8 bank groups, which makes the accesses of a quarter warp
// (LDS/STS.128) or of a half warp reading 8-byte halves (LDS.64) conflict-free.  Bit 3 has no class (it
// moves by a whole 128-byte row): the host never lets it vary inside a wavefront of the dense paths when it
// can avoid it, and orders the tile qubits so that the two bits of every dense gate fall in different
// classes.  Every index map of the kernel is a sum of per-bit increments.
__host__ __device__ __forceinline__ uint32_t pad_of_bit(int j) {
  switch (j) {
    case 4: return 1;
    case 5: return 2;
    case 6: return 4;
    case 7: return 9;
    case 8: return 18;
    case 9: return 36;
    case 10: return 73;
    case 11: return 146;
    case 12: return 292;
    default: return 0;
  }
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) and their mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Long waits (the producer warp waiting for a tile's gates): the hardware parks the thread for up to `hint_ns` or
// until the phase completes, instead of re-issuing the test every few hundred cycles on a sub-partition that the
// compute warps need.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared, completion counted on the mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// one lane of the (converged) warp, the same one every time
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor path (SASS: DMMA.8x8x4).
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// D(8x8) = A(8x4) * B(4x8) + C with C kept (the accumulator input is a different register pair)
__device__ __forceinline__ void dmma884_c(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// physical base address (in amplitudes) of tile t: spread t over the non-tile bits
__device__ __forceinline__ u64 tile_base(const PassDesc& d, u64 t) {
  for (int j = 0; j < d.n_spread; j++) {
    const int p = d.tbs[j];
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

// Persistent kernel, 2-4 CTAs per SM (one warp per 2^8 amplitudes of the tile), one tile buffer each:
// while one CTA waits for HBM (TMA bulk load of its next tile, bulk write-back of the last) the others
// run their gates, so the memory phases of one overlap the FP64 / shared-memory phases of the others.
// A CTA owns a contiguous range of tiles.  The tile moves between HBM and shared memory with TMA bulk
// copies of 256-byte pieces (one per two threads), which bypass the LSU/L1TEX pipe that the gates saturate.
//
// kG = 1: the CTA is one group of threads with ONE tile buffer (2-4 CTAs per SM).
// kG = 3: one CTA per SM holds three independent warp GROUPS of 256 threads (named barriers), each with TWO
// tile buffers, and all groups share the pass's read-only tables, which is what makes six 35 KiB buffers fit
// in 227 KiB.  A group loads tile t+1 into its other buffer while it applies the gates of tile t, and the
// write-back of tile t drains while tile t+1 is computed: neither the HBM load latency nor the store drain
// is on the group's critical path, so the FP64 / shared-memory pipes stay busy.  Every thread loads and
// stores the SAME shared-memory piece, so a buffer is recycled without any group-wide barrier: a thread waits
// for its own store to have been read out (bulk wait_group.read) and then issues its own load.
template <int kG, bool kProd>
__device__ __forceinline__ void tile_pass_body(double2* __restrict__ sv, const PassDesc& d,
                                               const double2* __restrict__ tail_ent) {
  // dynamic shared memory: the (padded) tile buffers, then the per-pass tables (sized by the pass, so that
  // short passes leave room for more resident CTAs)
  extern __shared__ double2 bufs[];
  constexpr bool kDB = kG > 1;
  constexpr int kBuf = kDB ? 2 : 1;
  __shared__ __align__(8) uint64_t ld_bar[kG][kBuf];
  __shared__ __align__(8) uint64_t done_bar[kProd ? kG : 1][kBuf];  // kProd: "gates of this buffer's tile are done"
  const int nops_all = d.nops, nst_all = d.nstages;
  const int b = d.b;
  const uint32_t tile_amps = 1u << b;
  uint32_t tile_slots = tile_amps;
  for (int j = 4; j < b; j++) tile_slots += pad_of_bit(j);
  const uint32_t ctid = threadIdx.x, cthreads = blockDim.x;  // CTA-wide (table set-up)
  // kProd: threads [0, 256 kG) are the compute warps of the groups, warp 8 kG + g is the TMA producer of group g
  const bool is_prod = kProd && ctid >= 256u * kG;
  const uint32_t grp_raw = is_prod ? ((ctid - 256u * kG) >> 5) : (ctid >> 8);
  const uint32_t grp = kDB ? __shfl_sync(0xffffffffu, grp_raw, 0) : 0u;  // warp group (warp-uniform value)
  const uint32_t tid = kDB ? (ctid & 255u) : ctid;           // thread within the group
  const uint32_t kPT = kDB ? 256u : cthreads;                // threads per group
  double2* tile = bufs + (size_t)(grp * kBuf) * tile_slots;  // the group's current tile buffer
  auto gsync = [&]() {
    if constexpr (kDB) asm volatile("bar.sync %0, 256;\n" ::"r"(grp + 1u) : "memory");
    else __syncthreads();
  };
  double* frag = reinterpret_cast<double*>(bufs + (size_t)(kG * kBuf) * tile_slots);  // [nfrag][64] B fragments
  OpRec* oprec = reinterpret_cast<OpRec*>(frag + (size_t)d.nfrag * 64);         // [nops]
  uint32_t (*opl)[32] = reinterpret_cast<uint32_t (*)[32]>(oprec + nops_all);   // [nops][32] lane slots
  uint32_t (*st_warp)[kWarps] = reinterpret_cast<uint32_t (*)[kWarps]>(opl + nops_all);  // [nstages][16]
  uint32_t (*st_lane)[32] = reinterpret_cast<uint32_t (*)[32]>(st_warp + nst_all);       // [nstages][32]
  uint32_t (*st_it)[8] = reinterpret_cast<uint32_t (*)[8]>(st_lane + nst_all);           // [nstages][8]
  // tabulated diagonal runs: [nruns][256] complex (16-byte aligned), then [nruns][56] index parts
  const uintptr_t run_base = (reinterpret_cast<uintptr_t>(st_it + nst_all) + 15u) & ~(uintptr_t)15u;
  double2* runtab = reinterpret_cast<double2*>(run_base);  // the runs' tables back to back (DiagRun::toff)
  uint16_t (*runix)[56] = reinterpret_cast<uint16_t (*)[56]>(runtab + d.run_entries);
  __shared__ double2 tail_f_all[kG][16][2];    // per tile: factor pair of every tile bit; [b] = scalar
  __shared__ double2 tail_tab_all[kG][2][64];  // per tile: the two tables of the diagonal tail
  double2 (*tail_f)[2] = tail_f_all[grp];
  double2 (*tail_tab)[64] = tail_tab_all[grp];
  __shared__ uint16_t tail_ix[2][56];
  __shared__ uint8_t tail_pk[2][kMaxTail];   // row-selecting physical bits of the tail entries

  const uint32_t dbg = d.flags;  // 1 = no write-back, 2 = no loads, 4 = no gates
  const int nops = d.nops, nstages = (dbg & 4u) ? 0 : d.nstages;

  // ---- per-kernel setup (amortised over all tiles of this CTA): every lane/warp dependent index of
  // the pass is tile-invariant, so it is digested once into shared-memory tables / registers ----
  u64 tile_mask = d.fixed_mask;  // physical bits the tile walk skips: the tile qubits and a sub-block's fixed bits
  for (int j = 0; j < b; j++) tile_mask |= 1ull << d.tb[j];
  if (ctid == 0) {
    // kDB: one arrival per issuing warp (kProd: the producer's single arrive.expect_tx)
    for (int q = 0; q < kG * kBuf; q++) mbar_init(&ld_bar[0][0] + q, kProd ? 1u : kDB ? kPT >> 5 : kPT);
    if constexpr (kProd)
      for (int q = 0; q < kG * kBuf; q++) mbar_init(&done_bar[0][0] + q, kPT >> 5);  // one arrival per compute warp
  }
  for (int idx = ctid; idx < nops * 64; idx += cthreads) {
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
      frag[op.foff * 64 + e] = v;
    }
    if (e < 32) {
      const uint32_t g = e >> 2, c = e & 3u;
      uint32_t s_g = 0;
      if (g & 1u) s_g += op.col[0];
      if (g & 2u) s_g += op.col[1];
      if (g & 4u) s_g += op.col[2];
      const uint32_t ld_x = (c & 2u) ? op.z0 : 0u;                                 // amplitude c>>1 of the k-slice
      const uint32_t st_x = ((c & 1u) ? op.z0 : 0u) + ((c & 2u) ? op.z1 : 0u);     // amplitude c of the group
      opl[o][e] = (s_g + ld_x) | ((s_g + st_x) << 16);
    }
    if (e == 32) {
      OpRec r;
      r.z0 = op.z0; r.z1 = op.z1; r.c3 = op.col[3]; r.c4 = op.col[4]; r.c5 = op.col[5];
      r.kind = op.kind; r.ngl = op.ngl; r.t0 = op.t0; r.t1 = op.t1;
      r.moff = op.moff;
      if (op.kind == kGateDense3) { r.c5 = op.z2; r.moff = op.foff; }  // a fused block has at most 5 free bits
      if (op.kind == QSV_GATE_DENSE2 || op.kind == QSV_GATE_DIAG1 || op.kind == QSV_GATE_DIAG2) r.moff = op.foff;
      oprec[o] = r;
    }
    // diagonal gates: the 2 or 4 factors live in shared memory (a lane-dependent index into the kernel
    // parameters would be a divergent constant load, serialised per distinct address)
    if ((op.kind == QSV_GATE_DIAG1 || op.kind == QSV_GATE_DIAG2) && e >= 48 && e < 56)
      frag[op.foff * 64 + (e - 48)] = (op.kind == QSV_GATE_DIAG2 || e - 48 < 4) ? d.mats[op.moff + (e - 48)] : 0.0;
  }
  // B fragments of fused 3-qubit blocks, 3-multiplication form of the complex product (M = P + iQ, amplitudes
  // x + iy):  k1 = P (x + y),  re = k1 - (P + Q) y,  im = k1 + (Q - P) x  -- three real 8x8 products (6 DMMA per
  // 64 amplitudes) instead of the 16x16 real matrix [[P,-Q],[Q,P]] (8 DMMA).  Fragment f = 2 p + s: product p
  // (0: P, 1: -(P+Q), 2: Q-P), k-slice s; lane l holds Mp[row(n = l >> 2)][4 s + (l & 3)], where column n of
  // the accumulator is output amplitude (n >> 1) + 4 (n & 1): a lane's two accumulator columns then differ in
  // the block's third bit and its quad position gives the two low bits, as in the A operand.
  for (int idx = ctid; idx < nops * 192; idx += cthreads) {
    const int o = idx / 192, e = idx - 192 * o;
    const PassOp& op = d.ops[o];
    if (op.kind != kGateDense3) continue;
    const int f = e >> 5, l = e & 31;
    const int p = f >> 1, sl = f & 1;
    const int nn = l >> 2;
    const int r = (nn >> 1) + 4 * (nn & 1), c = 4 * sl + (l & 3);
    const double* m = d.mats + op.moff + 2 * (8 * r + c);
    frag[op.foff * 64 + e] = p == 0 ? m[0] : p == 1 ? -(m[0] + m[1]) : m[1] - m[0];
  }
  for (int idx = ctid; idx < nstages * 64; idx += cthreads) {
    const int sidx = idx >> 6, e = idx & 63;
    const PassStage& st = d.stages[sidx];
    uint32_t li = 0, sl = 0;
    if (e < kWarps) {  // warp part
      for (int j = 0; j < st.n_outer; j++)
        if ((e >> j) & 1) { li |= 1u << st.obit[j]; sl += st.ocol[j]; }
      st_warp[sidx][e] = li | (sl << 16);
    } else if (e >= 24 && e < 32) {  // element iteration part (element bits 5, 6, 7)
      for (int j = 0; j < 3; j++)
        if ((((e - 24) >> j) & 1) && 5 + j < st.n_inner) { li |= 1u << st.ebit[5 + j]; sl += st.ecol[5 + j]; }
      st_it[sidx][e - 24] = li | (sl << 16);
    } else if (e >= 32) {  // lane part (element bits 0..4)
      for (int j = 0; j < 5 && j < st.n_inner; j++)
        if (((e - 32) >> j) & 1) { li |= 1u << st.ebit[j]; sl += st.ecol[j]; }
      st_lane[sidx][e - 32] = li | (sl << 16);
    }
  }
  for (int idx = ctid; idx < d.nruns * (1 << kRunBits); idx += cthreads) {
    const int r = idx >> kRunBits, x = idx & ((1 << kRunBits) - 1);
    const DiagRun& run = d.runs[r];
    if (x < (1 << run.nbits)) runtab[run.toff + x] = make_double2(d.mats[run.moff + 2 * x], d.mats[run.moff + 2 * x + 1]);
    if (x < 56) runix[r][x] = x < kWarps ? run.xw[x] : x < kWarps + 32 ? run.xl[x - kWarps] : run.xit[x - kWarps - 32];
  }
  for (int idx = ctid; d.tail.n && idx < 112; idx += cthreads) {
    const int g2 = idx / 56, x = idx % 56;
    tail_ix[g2][x] = x < kWarps ? d.tail.xw[g2][x] : x < kWarps + 32 ? d.tail.xl[g2][x - kWarps] : d.tail.xit[g2][x - kWarps - 32];
  }
  for (int idx = ctid; idx < d.tail.n; idx += cthreads) {
    tail_pk[0][idx] = d.tail.pk[idx];
    tail_pk[1][idx] = d.tail.pk2[idx];
  }
  // bulk-copy pieces of this thread: piece p covers tile-local bits [0, pb); its index bits are the
  // tile-local bits pb.. ; thread t owns pieces t, t + kPT, ...
  const uint32_t pb_ld = d.pb_ld, pb_st = d.pb_st;
  const uint32_t np_ld = tile_amps >> pb_ld, np_st = tile_amps >> pb_st;
  const uint32_t ld_bytes = 16u << pb_ld, st_bytes = 16u << pb_st;
  // With fewer pieces than threads (256-byte pieces: every second thread) the issuing threads are spread
  // evenly over the warps, because a warp issues its bulk copies one lane at a time.
  const uint32_t sh_ld = np_ld < kPT ? (uint32_t)__ffs((int)(kPT / np_ld)) - 1u : 0u;
  const uint32_t sh_st = np_st < kPT ? (uint32_t)__ffs((int)(kPT / np_st)) - 1u : 0u;
  const uint32_t pc_ld0 = tid >> sh_ld, pc_st0 = tid >> sh_st;
  const bool own_ld = (tid & ((1u << sh_ld) - 1u)) == 0u && pc_ld0 < np_ld;
  const bool own_st = (tid & ((1u << sh_st) - 1u)) == 0u && pc_st0 < np_st;
  uint32_t my_ld_bytes = own_ld ? ld_bytes : 0u;
  for (uint32_t pc = tid + kPT; pc < np_ld; pc += kPT) my_ld_bytes += ld_bytes;
  // the common case (at most one piece per thread) keeps its two offsets in registers
  u64 g_ld0 = 0, g_st0 = 0;
  uint32_t s_ld0 = 0, s_st0 = 0;
  for (int j = 0; j + (int)pb_ld < b; j++)
    if ((pc_ld0 >> j) & 1u) { g_ld0 |= 1ull << d.tb[pb_ld + j]; s_ld0 += d.cj[pb_ld + j]; }
  for (int j = 0; j + (int)pb_st < b; j++)
    if ((pc_st0 >> j) & 1u) { g_st0 |= 1ull << d.tb[pb_st + j]; s_st0 += d.cj[d.src_of[pb_st + j]]; }
  const bool ld_on = !(dbg & 2u), st_on = !(dbg & 1u);
  // L2 prefetch of the next tile: every thread covers its share of one load piece
  const uint32_t pf_bytes = (ld_bytes >> sh_ld) ? (ld_bytes >> sh_ld) : ld_bytes;
  const uint32_t pf_off = (tid & ((1u << sh_ld) - 1u)) * (ld_bytes >> sh_ld);
  const bool pf_on = !(dbg & 32u) && pc_ld0 < np_ld && (ld_bytes >> sh_ld) != 0u;
  const uint32_t lane = tid & 31u, wid = tid >> 5;
  const uint32_t wid_u = kDB ? __shfl_sync(0xffffffffu, wid, 0) : wid;  // warp-uniform for the compiler
  const uint32_t ppw = np_ld / (kPT >> 5);                             // kDB: bulk-copy pieces per warp
  const uint32_t g = lane >> 2;
  const char* tb8 = reinterpret_cast<const char*>(tile) + (lane & 1u) * 8u;  // (re-pointed per tile when kDB)
  fence_proxy_async();  // make the mbarrier initialisation visible to the async proxy
  __syncthreads();

  // contiguous range of tiles for this group; the tile base walks the non-tile bits: next = ((base | M) + 1) & ~M
  const u64 n_tiles = d.n_tiles;
  const u64 n_groups = (u64)gridDim.x * kG;
  const u64 per = (n_tiles + n_groups - 1) / n_groups;
  const u64 t_begin = ((u64)blockIdx.x * kG + grp) * per;
  const u64 t_end = t_begin + per < n_tiles ? t_begin + per : n_tiles;
  u64 base = t_begin < n_tiles ? tile_base(d, t_begin) : 0;
  uint32_t parity = 0;
  if constexpr (kProd) {
    // ---- TMA producer warp of group `grp`: the compute warps of the group never issue a bulk copy, never poll
    // for a store to drain and need no group-wide barrier per tile.  Per tile t (buffer t & 1): wait until the
    // group's 8 compute warps have arrived on done_bar (gates finished, generic-proxy writes fenced), write the
    // tile back, wait until the bulk stores have READ the buffer, then request tile t + 2 into it. ----
    if (is_prod) {
      if (t_begin >= t_end) return;
      const uint32_t n_pc = np_ld;  // pieces per tile (np_ld == np_st in this shape)
      u64 base_st = base;           // tile being written back
      u64 base_ld = base;           // tile being requested
      auto issue_load = [&](uint32_t buf) {
        if (elect_one()) {
          if (ld_on) {
            mbar_arrive_expect_tx(&ld_bar[grp][buf], n_pc * ld_bytes);
            const char* gsrc = reinterpret_cast<const char*>(sv + base_ld);
            char* sdst = reinterpret_cast<char*>(bufs + (size_t)(grp * kBuf + buf) * tile_slots);
            for (uint32_t p = 0; p < n_pc; p++)
              bulk_load(sdst + d.pc_s[p], gsrc + d.pc_g[p], ld_bytes, &ld_bar[grp][buf]);
          } else {
            mbar_arrive(&ld_bar[grp][buf]);  // profiling without loads: keep the hand-shake
          }
        }
      };
      // prologue: tiles t_begin and t_begin + 1
      issue_load(0);
      base_ld = ((base_ld | tile_mask) + 1) & ~tile_mask;
      if (t_begin + 1 < t_end) {
        issue_load(1);
        base_ld = ((base_ld | tile_mask) + 1) & ~tile_mask;
      }
      uint32_t i = 0;
      PROF_DECL
      for (u64 t = t_begin; t < t_end; t++, i++) {
        const uint32_t buf = i & 1u;
        if (dbg & 1024u) mbar_wait(&done_bar[grp][buf], (i >> 1) & 1u);
        else mbar_wait_parked(&done_bar[grp][buf], (i >> 1) & 1u, 20000u);
        PROF_MARK(0);
        if (elect_one()) {
          if (st_on) {
            char* gdst = reinterpret_cast<char*>(sv + base_st);
            const char* ssrc = reinterpret_cast<const char*>(bufs + (size_t)(grp * kBuf + buf) * tile_slots);
            for (uint32_t p = 0; p < n_pc; p++) bulk_store(gdst + d.pc_g[p], ssrc + d.pc_s[p], st_bytes);
            bulk_commit();
          }
        }
        __syncwarp();
        PROF_MARK(1);
        if (elect_one())
          if (t + 2 < t_end) bulk_wait_read();  // the buffer has been read out: it may be overwritten
        __syncwarp();
        PROF_MARK(2);
        base_st = ((base_st | tile_mask) + 1) & ~tile_mask;
        if (t + 2 < t_end) {
          issue_load(buf);
          base_ld = ((base_ld | tile_mask) + 1) & ~tile_mask;
        }
        __syncwarp();
        PROF_MARK(3);
      }
      if (elect_one()) bulk_wait_read();
      PROF_FLUSH(8);
      return;
    }
  }
  if constexpr (kDB && !kProd) {  // prologue: the first tile goes into buffer 0
    if (ld_on && t_begin < t_end) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&ld_bar[grp][0], ppw * ld_bytes);
        const double2* gsrc = sv + base;
        for (uint32_t i = 0; i < ppw; i++) {
          const uint32_t p = wid_u * ppw + i;
          bulk_load(reinterpret_cast<char*>(tile) + d.pc_s[p], reinterpret_cast<const char*>(gsrc) + d.pc_g[p], ld_bytes,
                    &ld_bar[grp][0]);
        }
      }
    }
  }
  uint32_t t_idx = 0;
  PROF_DECL
  for (u64 t = t_begin; t < t_end; t++, t_idx++, base = ((base | tile_mask) + 1) & ~tile_mask) {
    bool next_issued = true;
    // kDB: load of the group's NEXT tile into its other buffer, issued by every thread for its own piece once
    // its write-back piece of the previous tile has left that buffer (no barrier needed, see above)
    auto issue_next = [&]() {
      if constexpr (kDB && !kProd) {
        if (elect_one()) {  // the lane that also issued the write-back of these pieces
          bulk_wait_read();
          const u64 nb = ((base | tile_mask) + 1) & ~tile_mask;
          const uint32_t nxt = (t_idx & 1u) ^ 1u;
          mbar_arrive_expect_tx(&ld_bar[grp][nxt], ppw * ld_bytes);
          const double2* gsrc = sv + nb;
          double2* sdst = bufs + (size_t)(grp * kBuf + nxt) * tile_slots;
          for (uint32_t i = 0; i < ppw; i++) {
            const uint32_t p = wid_u * ppw + i;
            bulk_load(reinterpret_cast<char*>(sdst) + d.pc_s[p], reinterpret_cast<const char*>(gsrc) + d.pc_g[p],
                      ld_bytes, &ld_bar[grp][nxt]);
          }
        }
        next_issued = true;
      }
    };
    if constexpr (kProd) {
      // ---- the group's producer warp requested this tile two tiles ago; every thread acquires the TMA writes ----
      const uint32_t cur = t_idx & 1u;
      tile = bufs + (size_t)(grp * kBuf + cur) * tile_slots;
      tb8 = reinterpret_cast<const char*>(tile) + (lane & 1u) * 8u;
      PROF_MARK(2);
      mbar_wait(&ld_bar[grp][cur], (t_idx >> 1) & 1u);
      PROF_MARK(0);
    }
    if constexpr (kDB && !kProd) {
      // ---- this tile was requested one tile ago ----
      const uint32_t cur = t_idx & 1u;
      tile = bufs + (size_t)(grp * kBuf + cur) * tile_slots;
      tb8 = reinterpret_cast<const char*>(tile) + (lane & 1u) * 8u;
      // one warp polls the mbarrier, the others park on the group barrier (debug flag 256: every thread polls
      // the mbarrier itself and there is no barrier here -- measured the same, profiles/r1_qv30_variants.txt)
      if (ld_on && (dbg & 256u)) mbar_wait(&ld_bar[grp][cur], (t_idx >> 1) & 1u);
      if (ld_on && !(dbg & 256u)) {
        if (wid == 0) mbar_wait(&ld_bar[grp][cur], (t_idx >> 1) & 1u);
        gsync();
      }
      next_issued = !(ld_on && t + 1 < t_end);
      // request the next tile right away: the write-back of the previous tile (issued one barrier ago) drains
      // quickly, and the load then has the whole of this tile's gates to arrive (flag 128: after the first op)
      if ((nstages == 0 || !(dbg & 128u)) && !next_issued) issue_next();
    }
    // ---- load: one HBM read of the tile (TMA bulk copies, completion on the mbarrier) ----
    if (!kDB && t != t_begin) {
      bulk_wait_read();   // my write-back pieces of the previous tile have left shared memory
      __syncthreads();    // ... and so have everybody else's
    }
    if (!kDB && ld_on) {
      uint64_t* const bar = &ld_bar[0][0];
      mbar_arrive_expect_tx(bar, my_ld_bytes);
      if (own_ld) bulk_load(tile + s_ld0, sv + base + g_ld0, ld_bytes, bar);
      for (uint32_t pc = tid + kPT; pc < np_ld; pc += kPT) {  // small pieces only (tile without the low qubits)
        u64 go = 0;
        uint32_t so = 0;
        for (int j = 0; j + (int)pb_ld < b; j++)
          if ((pc >> j) & 1u) { go |= 1ull << d.tb[pb_ld + j]; so += d.cj[pb_ld + j]; }
        bulk_load(tile + so, sv + base + go, ld_bytes, bar);
      }
      // one warp polls the mbarrier (acquiring the TMA writes); everybody else parks on the hardware
      // barrier instead of burning issue slots that the other CTAs of this SM need for their gates
      if (wid == 0) mbar_wait(bar, parity);
      parity ^= 1u;
      __syncthreads();
      // pull the next tile of this CTA into L2 while the gates run, so that its bulk load sees L2 latency
      // instead of HBM latency (the CTA has a single tile buffer, the load cannot start earlier)
      if (pf_on && t + 1 < t_end) {
        const u64 nb = ((base | tile_mask) + 1) & ~tile_mask;
        const char* pa = reinterpret_cast<const char*>(sv + nb + g_ld0) + pf_off;
        for (uint32_t o2 = 0; o2 < pf_bytes; o2 += 128u)
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(pa + o2));
      }
    }

    int o = 0;
    PROF_OP_BEGIN
    for (int sidx = 0; sidx < nstages; sidx++) {
      if (sidx > 0) {  // stage boundary: warps re-partition the tile
        const uint32_t sm = kDB ? d.stages[sidx].sync_mask : 0u;  // (one CTA per SM: its 16 named barriers are free)
        if (sm == 0u) {
          gsync();
        } else if constexpr (kDB) {
          // only the warps that agree on the outer bits both stages share exchange data: a barrier among those
          // 2 or 4 warps (named barriers 4 .. 15: four per group) instead of the whole group
          uint32_t sub = 0, k2 = 0;
          for (uint32_t j = 0; j < 3u; j++)
            if ((sm >> j) & 1u) sub |= ((wid_u >> j) & 1u) << k2++;
          asm volatile("bar.sync %0, %1;\n" ::"r"(4u + grp * 4u + sub), "r"(256u >> k2) : "memory");
        }
      }
      PROF_OP_MARK(4);
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
          const double* fr = frag + (uint32_t)rec.moff * 64u;
          const double b0 = fr[lane], b1 = fr[32 + lane];
          const uint32_t v = opl[o][lane];
          const uint32_t a_ld = (s_w + (v & 0xffffu)) << 4;  // byte offsets of 16-byte slots
          const uint32_t a_st = (s_w + (v >> 16)) << 4;
          const uint32_t z1b = (uint32_t)rec.z1 << 4;
          const uint32_t c3 = (uint32_t)rec.c3 << 4, c4 = (uint32_t)rec.c4 << 4, c5 = (uint32_t)rec.c5 << 4;
          if (rec.ngl == 6u) {
            // software pipeline over the sub-cube's eight DMMA steps, operands requested two steps ahead
            const uint32_t xo[8] = {0u, c3, c4, c3 + c4, c5, c5 + c3, c5 + c4, c5 + c3 + c4};
            double xa[3], xb[3];
#pragma unroll
            for (int j = 0; j < 2; j++) {
              xa[j] = *reinterpret_cast<const double*>(tb8 + (a_ld + xo[j]));
              xb[j] = *reinterpret_cast<const double*>(tb8 + (a_ld + xo[j] + z1b));
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
              if (j + 2 < 8) {
                xa[(j + 2) % 3] = *reinterpret_cast<const double*>(tb8 + (a_ld + xo[j + 2]));
                xb[(j + 2) % 3] = *reinterpret_cast<const double*>(tb8 + (a_ld + xo[j + 2] + z1b));
              }
              double d0 = 0.0, d1 = 0.0;
              dmma884(d0, d1, xa[j % 3], b0);
              dmma884(d0, d1, xb[j % 3], b1);
              *reinterpret_cast<double2*>(reinterpret_cast<char*>(tile) + (a_st + xo[j])) = make_double2(d0, d1);
            }
          } else {
            const uint32_t ng = 1u << rec.ngl;
            const uint32_t n_it = ng > 8u ? (ng >> 3) : 1u;
            for (uint32_t it = 0; it < n_it; it++) {
              uint32_t xo = 0;
              if (it & 1u) xo += c3;
              if (it & 2u) xo += c4;
              if (it & 4u) xo += c5;
              const bool valid = (g | (it << 3)) < ng;
              double x0 = 0.0, x1 = 0.0;
              if (valid) {
                const uint32_t a = a_ld + xo;
                x0 = *reinterpret_cast<const double*>(tb8 + a);
                x1 = *reinterpret_cast<const double*>(tb8 + (a + z1b));
              }
              double d0 = 0.0, d1 = 0.0;
              dmma884(d0, d1, x0, b0);
              dmma884(d0, d1, x1, b1);
              if (valid) *reinterpret_cast<double2*>(reinterpret_cast<char*>(tile) + (a_st + xo)) = make_double2(d0, d1);
            }
          }
        } else if (rec.kind == kGateDense3) {
          // ---- fused pair of 2-qubit gates sharing a qubit: one 8x8 complex block per 8 amplitudes, i.e. one
          // shared-memory round trip for two gates, in the 3-multiplication form (see the fragment set-up):
          // per 64 amplitudes 2 LDS.128 + 2 DADD + 6 DMMA + 2 STS.128 per lane.  Lane (g, c) reads amplitude c
          // (bits t0, t1) of the two halves (bit t2) of group g and writes the same two slots. ----
          const double* fr = frag + (uint32_t)rec.moff * 64u;
          double bq[6];
#pragma unroll
          for (int q = 0; q < 6; q++) bq[q] = fr[q * 32 + lane];
          const uint32_t v = opl[o][lane];
          char* const ta = reinterpret_cast<char*>(tile) + ((s_w + (v >> 16)) << 4);
          const uint32_t z2b = (uint32_t)rec.c5 << 4;
          const uint32_t c3 = (uint32_t)rec.c3 << 4, c4 = (uint32_t)rec.c4 << 4;
          if (rec.ngl == 5u) {
            // software pipeline over the sub-cube's four DMMA steps: the operands of step j + 1 are requested
            // before the six DMMA of step j, so a warp waits for shared memory only once per op and keeps the FP64
            // pipe and the shared-memory pipe busy at the same time
            const uint32_t xo[4] = {0u, c3, c4, c3 + c4};
            double2 u0n = *reinterpret_cast<const double2*>(ta + xo[0]);
            double2 u1n = *reinterpret_cast<const double2*>(ta + (xo[0] + z2b));
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const double2 u0 = u0n, u1 = u1n;
              if (j + 1 < 4) {
                u0n = *reinterpret_cast<const double2*>(ta + xo[j + 1]);
                u1n = *reinterpret_cast<const double2*>(ta + (xo[j + 1] + z2b));
              }
              double k0 = 0.0, k1 = 0.0, r0, r1, i0, i1;
              dmma884(k0, k1, u0.x + u0.y, bq[0]);
              dmma884(k0, k1, u1.x + u1.y, bq[1]);
              dmma884_c(r0, r1, u0.y, bq[2], k0, k1);
              dmma884_c(i0, i1, u0.x, bq[4], k0, k1);
              dmma884(r0, r1, u1.y, bq[3]);
              dmma884(i0, i1, u1.x, bq[5]);
              *reinterpret_cast<double2*>(ta + xo[j]) = make_double2(r0, i0);
              *reinterpret_cast<double2*>(ta + (xo[j] + z2b)) = make_double2(r1, i1);
            }
          } else {
            const uint32_t ng = 1u << rec.ngl;
            const uint32_t n_it = ng > 8u ? (ng >> 3) : 1u;
            for (uint32_t it = 0; it < n_it; it++) {
              uint32_t xo = 0;
              if (it & 1u) xo += c3;
              if (it & 2u) xo += c4;
              const bool valid = (g | (it << 3)) < ng;
              double2 u0 = make_double2(0.0, 0.0), u1 = make_double2(0.0, 0.0);
              if (valid) {
                u0 = *reinterpret_cast<const double2*>(ta + xo);
                u1 = *reinterpret_cast<const double2*>(ta + (xo + z2b));
              }
              double k0 = 0.0, k1 = 0.0, r0, r1, i0, i1;
              dmma884(k0, k1, u0.x + u0.y, bq[0]);
              dmma884(k0, k1, u1.x + u1.y, bq[1]);
              dmma884_c(r0, r1, u0.y, bq[2], k0, k1);
              dmma884_c(i0, i1, u0.x, bq[4], k0, k1);
              dmma884(r0, r1, u1.y, bq[3]);
              dmma884(i0, i1, u1.x, bq[5]);
              if (valid) {
                *reinterpret_cast<double2*>(ta + xo) = make_double2(r0, i0);
                *reinterpret_cast<double2*>(ta + (xo + z2b)) = make_double2(r1, i1);
              }
            }
          }
        } else {
          // ---- element-wise kinds: each lane owns elements e = lane | it << 5 of the sub-cube ----
          const uint32_t n_el = 1u << n_inner;
          const uint32_t lv = wv + st_lane[sidx][lane];  // li (low 16 bits) and slot (high 16): no carries between them
          const bool is_diag = rec.kind == QSV_GATE_DIAG1 || rec.kind == QSV_GATE_DIAG2;
          if (is_diag && rec.ngl) {
            // tabulated run: one table lookup and one complex multiply per element for the whole run
            const int r = rec.ngl - 1;
            const double2* tab = runtab + d.runs[r].toff;
            const uint32_t x0 = (uint32_t)runix[r][wid] + (uint32_t)runix[r][kWarps + lane];
            for (uint32_t it = 0; it < 8u; it++) {
              if ((lane | (it << 5)) >= n_el) break;
              const uint32_t sl = (lv + st_it[sidx][it]) >> 16;
              tile[sl] = cmul(tab[x0 + runix[r][kWarps + 32 + it]], tile[sl]);
            }
            o += d.runs[r].n_ops - 1;
          } else if (is_diag) {
            // a run of consecutive diagonal gates of the stage is applied in one read-modify-write
            int o_last = o;
            while (o_last + 1 < o_end && !oprec[o_last + 1].ngl &&
                   (oprec[o_last + 1].kind == QSV_GATE_DIAG1 || oprec[o_last + 1].kind == QSV_GATE_DIAG2))
              o_last++;
            for (uint32_t it = 0; it < 8u; it++) {
              if ((lane | (it << 5)) >= n_el) break;
              const uint32_t ev = lv + st_it[sidx][it];
              const uint32_t li = ev & 0xffffu, sl = ev >> 16;
              double2 a = tile[sl];
              for (int q = o; q <= o_last; q++) {
                const OpRec rq = oprec[q];
                const double2* fac = reinterpret_cast<const double2*>(frag + (uint32_t)rq.moff * 64u);
                uint32_t x = (li >> rq.t0) & 1u;
                if (rq.kind == QSV_GATE_DIAG2) x |= ((li >> rq.t1) & 1u) << 1;
                a = cmul(fac[x], a);
              }
              tile[sl] = a;
            }
            o = o_last;
          } else {
            const double* m = d.mats + rec.moff;
            for (uint32_t it = 0; it < 8u; it++) {
              if ((lane | (it << 5)) >= n_el) break;
              const uint32_t ev = lv + st_it[sidx][it];
              const uint32_t li = ev & 0xffffu, sl = ev >> 16;
              if (rec.kind == QSV_GATE_CX) {
                // control t0 set, target t1 clear: swap with the target-set partner (same sub-cube)
                if (((li >> rec.t0) & 1u) && !((li >> rec.t1) & 1u)) {
                  const uint32_t sp = sl + rec.z1;
                  const double2 a = tile[sl], pp = tile[sp];
                  tile[sl] = pp;
                  tile[sp] = a;
                }
              } else if (rec.kind == QSV_GATE_DENSE1) {
                // scalar 2x2 (only used when the tile has a single qubit): element with t0 clear owns the pair
                if (!((li >> rec.t0) & 1u)) {
                  const uint32_t sp = sl + rec.z0;
                  const double2 a0 = tile[sl], a1 = tile[sp];
                  double2 r0 = cmul(make_double2(m[0], m[1]), a0), r1 = cmul(make_double2(m[4], m[5]), a0);
                  cfma(r0, make_double2(m[2], m[3]), a1);
                  cfma(r1, make_double2(m[6], m[7]), a1);
                  tile[sl] = r0;
                  tile[sp] = r1;
                }
              }
            }
          }
        }
        // kDB with debug flag 128: request the next tile only now, after the first op of this tile
        if constexpr (kDB)
          if (!next_issued) issue_next();
        PROF_OP_MARK(5 + o);
      }
    }
    if constexpr (kDB)
      if (!next_issued) issue_next();

    // ---- diagonal tail (see TailDesc): per-tile factors -> two small tables -> one sweep over the tile ----
    if (d.tail.n && !(dbg & 4u)) {
      // (1) factor pair of every tile bit and the scalar: one HALF warp per range, its 16 lanes over the
      // range's entries, then a shuffle product inside the half warp
      {
        const int half = (int)(lane >> 4), hl = (int)(lane & 15u), n_hw = 2 * (int)(kPT >> 5);
        for (int jj0 = 0; jj0 <= b; jj0 += n_hw) {
          const int jj = jj0 + 2 * (int)wid + half;
          const bool live = jj <= b;
          double2 f0 = make_double2(1.0, 0.0), f1 = make_double2(1.0, 0.0);
          const int e_begin = live ? d.tail.start[jj] : 0, e_end = live ? d.tail.start[jj + 1] : 0;
          for (int e0 = e_begin;; e0 += 16) {
            const int e = e0 + hl;
            const bool act = e < e_end;
            if (!__any_sync(0xffffffffu, act)) break;
            if (act) {
              const uint32_t k1 = tail_pk[0][e], k2 = tail_pk[1][e];
              const uint32_t r1 = k1 == 0xffu ? 0u : (uint32_t)((base >> k1) & 1ull);
              const uint32_t r2 = k2 == 0xffu ? 0u : (uint32_t)((base >> k2) & 1ull);
              const double2* ent = tail_ent + (size_t)e * 4;
              if (jj < b) {  // entry = [row][bit of the tile qubit]
                f0 = cmul(__ldg(ent + 2 * r1), f0);
                f1 = cmul(__ldg(ent + 2 * r1 + 1), f1);
              } else {       // scalar entry = [row][second row]
                f0 = cmul(__ldg(ent + 2 * r1 + r2), f0);
              }
            }
          }
#pragma unroll
          for (int off = 8; off > 0; off >>= 1) {
            double2 g0, g1;
            g0.x = __shfl_xor_sync(0xffffffffu, f0.x, off); g0.y = __shfl_xor_sync(0xffffffffu, f0.y, off);
            g1.x = __shfl_xor_sync(0xffffffffu, f1.x, off); g1.y = __shfl_xor_sync(0xffffffffu, f1.y, off);
            f0 = cmul(g0, f0);
            f1 = cmul(g1, f1);
          }
          if (live && hl == 0) { tail_f[jj][0] = f0; tail_f[jj][1] = f1; }
        }
      }
      gsync();
      // (2) tables over the two groups of tile bits (the scalar goes into the first)
      for (int idx = tid; idx < 128; idx += kPT) {
        const int g2 = idx >> 6, x = idx & 63;
        const int nbits = g2 ? d.tail.nb : d.tail.na;
        if (x < (1 << nbits)) {
          double2 v = g2 ? make_double2(1.0, 0.0) : tail_f[b][0];
          for (int i = 0; i < nbits; i++) v = cmul(tail_f[g2 ? d.tail.bbit[i] : d.tail.abit[i]][(x >> i) & 1], v);
          tail_tab[g2][x] = v;
        }
      }
      gsync();
      // (3) sweep, in the element order of the last stage
      const int sl_idx = nstages - 1;
      if (wid < (1u << d.stages[sl_idx].n_outer)) {
        const uint32_t n_el = 1u << d.stages[sl_idx].n_inner;
        const uint32_t lv = st_warp[sl_idx][wid] + st_lane[sl_idx][lane];
        const uint32_t xa0 = (uint32_t)tail_ix[0][wid] + (uint32_t)tail_ix[0][kWarps + lane];
        const uint32_t xb0 = (uint32_t)tail_ix[1][wid] + (uint32_t)tail_ix[1][kWarps + lane];
        const bool two = d.tail.nb != 0;
        for (uint32_t it = 0; it < 8u; it++) {
          if ((lane | (it << 5)) >= n_el) break;
          const uint32_t sl = (lv + st_it[sl_idx][it]) >> 16;
          double2 a = cmul(tail_tab[0][xa0 + tail_ix[0][kWarps + 32 + it]], tile[sl]);
          if (two) a = cmul(tail_tab[1][xb0 + tail_ix[1][kWarps + 32 + it]], a);
          tile[sl] = a;
        }
      }
    }

    // ---- write-back: one HBM write of the tile (TMA bulk copies), optionally with the tile qubits
    // permuted: position j receives the content of tile bit src_of[j] (a free qubit remap) ----
    fence_proxy_async();  // the gates' generic-proxy stores must be visible to the async proxy
    if constexpr (kProd) {
      // hand the buffer to the producer warp: one arrival per compute warp (release), no group-wide barrier --
      // the next tile lives in the other buffer
      __syncwarp();
      if (lane == 0) mbar_arrive(&done_bar[grp][t_idx & 1u]);
      PROF_MARK(1);
      continue;
    }
    gsync();
    if (kDB && st_on) {
      if (elect_one()) {  // one lane per warp, its warp's pieces, addresses in uniform registers
        double2* gdst = sv + base;
        for (uint32_t i = 0; i < ppw; i++) {
          const uint32_t p = wid_u * ppw + i;
          bulk_store(reinterpret_cast<char*>(gdst) + d.pc_g[p], reinterpret_cast<const char*>(tile) + d.pc_s[p], st_bytes);
        }
        bulk_commit();
      }
    }
    if (!kDB && st_on) {
      if (own_st) bulk_store(sv + base + g_st0, tile + s_st0, st_bytes);
      for (uint32_t pc = tid + kPT; pc < np_st; pc += kPT) {
        u64 go = 0;
        uint32_t so = 0;
        for (int j = 0; j + (int)pb_st < b; j++)
          if ((pc >> j) & 1u) { go |= 1ull << d.tb[pb_st + j]; so += d.cj[d.src_of[pb_st + j]]; }
        bulk_store(sv + base + go, tile + so, st_bytes);
      }
      bulk_commit();
    }
  }
  bulk_wait_read();
  if constexpr (kProd) PROF_FLUSH(wid);
}

// One body, three launch shapes: several resident CTAs per SM (4 x 256 threads for the default 11-qubit
// tile) keep more tiles in flight, which is what hides the HBM round trip of each CTA.
template <int kBlock, int kMinBlocks, int kG = 1, bool kProd = false>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
tile_pass_kernel(double2* __restrict__ sv, const __grid_constant__ PassDesc d, const double2* __restrict__ tail_ent) {
  tile_pass_body<kG, kProd>(sv, d, tail_ent);
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
  const u64 cap = (u64)sm_count() * 16;
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
  int t2;          // third bit of a fused block
  int foff;        // B-fragment offset (units of 64 doubles)
};

// ---- pair fusion (host): 2-qubit dense gates that share a qubit and are adjacent on it become one
// 3-qubit block.  The block costs the same FP64 work as its two gates (8 complex MACs per amplitude) but
// only one shared-memory round trip, which is what bounds the single-gate path.  Gates that act inside the
// bits of the current block are multiplied in for free. ----
static inline int popc_u(uint32_t x) { return __builtin_popcount(x); }

// `g4` (4x4 complex, row-major, index bit 0 = position pa, bit 1 = position pb) embedded into `nb` index bits
static void embed_gate(const double* g4, int nb, int pa, int pb, double* out) {
  const int dim = 1 << nb;
  const int keep = (dim - 1) & ~((1 << pa) | (1 << pb));
  for (int r = 0; r < dim; r++)
    for (int c = 0; c < dim; c++) {
      double re = 0.0, im = 0.0;
      if ((r & keep) == (c & keep)) {
        const int rg = ((r >> pa) & 1) | (((r >> pb) & 1) << 1);
        const int cg = ((c >> pa) & 1) | (((c >> pb) & 1) << 1);
        re = g4[2 * (4 * rg + cg)];
        im = g4[2 * (4 * rg + cg) + 1];
      }
      out[2 * (dim * r + c)] = re;
      out[2 * (dim * r + c) + 1] = im;
    }
}

// any 1- or 2-qubit op as a dense complex matrix (2x2 or 4x4, row-major) in its own qubit order
static void op_as_dense(int kind, const double* m, double* out) {
  switch (kind) {
    case QSV_GATE_DENSE1: memcpy(out, m, sizeof(double) * 8); break;
    case QSV_GATE_DENSE2: memcpy(out, m, sizeof(double) * 32); break;
    case QSV_GATE_DIAG1:
      memset(out, 0, sizeof(double) * 8);
      for (int i = 0; i < 2; i++) { out[2 * (2 * i + i)] = m[2 * i]; out[2 * (2 * i + i) + 1] = m[2 * i + 1]; }
      break;
    case QSV_GATE_DIAG2:
      memset(out, 0, sizeof(double) * 32);
      for (int i = 0; i < 4; i++) { out[2 * (4 * i + i)] = m[2 * i]; out[2 * (4 * i + i) + 1] = m[2 * i + 1]; }
      break;
    default: {  // QSV_GATE_CX: q0 = control = index bit 0, q1 = target (basicaertools.py:70-72)
      memset(out, 0, sizeof(double) * 32);
      static const int to[4] = {0, 3, 2, 1};
      for (int c = 0; c < 4; c++) out[2 * (4 * to[c] + c)] = 1.0;
    }
  }
}

// `g2` (2x2 complex) on index bit `pa` embedded into `nb` index bits
static void embed_gate1(const double* g2, int nb, int pa, double* out) {
  const int dim = 1 << nb;
  const int keep = (dim - 1) & ~(1 << pa);
  for (int r = 0; r < dim; r++)
    for (int c = 0; c < dim; c++) {
      double re = 0.0, im = 0.0;
      if ((r & keep) == (c & keep)) {
        const int rg = (r >> pa) & 1, cg = (c >> pa) & 1;
        re = g2[2 * (2 * rg + cg)];
        im = g2[2 * (2 * rg + cg) + 1];
      }
      out[2 * (dim * r + c)] = re;
      out[2 * (dim * r + c) + 1] = im;
    }
}

// out = a * b (complex dim x dim, row-major); out must not alias a or b
static void cmat_mul(const double* a, const double* b, int dim, double* out) {
  for (int r = 0; r < dim; r++)
    for (int c = 0; c < dim; c++) {
      double re = 0.0, im = 0.0;
      for (int k = 0; k < dim; k++) {
        const double ar = a[2 * (dim * r + k)], ai = a[2 * (dim * r + k) + 1];
        const double br = b[2 * (dim * k + c)], bi = b[2 * (dim * k + c) + 1];
        re += ar * br - ai * bi;
        im += ar * bi + ai * br;
      }
      out[2 * (dim * r + c)] = re;
      out[2 * (dim * r + c) + 1] = im;
    }
}

// bank class of a tile-local bit: it shifts the bank group by 2^(j mod 3) (see the slot layout)
static inline int bank_class(int bit) { return bit < 3 ? bit : bit == 3 ? -1 : (bit - 1) % 3; }
static inline bool differ(int a, int b) { return a >= 0 && b >= 0 && a != b; }
static inline bool independent3(int a, int b, int c) {
  // three bits spread over all 8 bank groups iff their classes are pairwise different (bit 3 has none)
  return differ(a, b) && differ(a, c) && differ(b, c);
}

// swap the two index bits of a 4x4 complex matrix (roles of a 2-qubit gate's qubits exchanged)
static void swap_gate_bits(double* m) {
  double t[32];
  memcpy(t, m, sizeof(t));
  static const int sw[4] = {0, 2, 1, 3};
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      m[2 * (4 * r + c)] = t[2 * (4 * sw[r] + sw[c])];
      m[2 * (4 * r + c) + 1] = t[2 * (4 * sw[r] + sw[c]) + 1];
    }
}

static inline int popc(uint32_t x) { return __builtin_popcount(x); }

// Greedy staging, same commutation rule as the planner: a gate may be hoisted into the current stage
// only over gates that touch none of its bits.
static int build_stages(const HostOp* ops, int n_ops, int b, int* order, PassStage* stages, uint32_t* inner_masks) {
  const int n_inner = std::min(b, kInner);
  bool done[kMaxOps] = {false};
  int n_done = 0, n_stages = 0, pos = 0;
  uint32_t all_need = 0;
  for (int i = 0; i < n_ops; i++) all_need |= ops[i].need;
  if (popc(all_need) <= n_inner) all_need = 0;  // a single stage: plain padding
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
    // pad the inner set with the lowest unused tile bits -- in a pass of several stages first with the bits some op
    // of the pass needs, so that the bits NO op needs stay outer bits in every stage (the warps that agree on them
    // never exchange data, see PassStage::sync_mask)
    for (int j = 0; j < b && popc(inner) < n_inner; j++)
      if (!((inner >> j) & 1u) && ((all_need >> j) & 1u)) inner |= 1u << j;
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

// Three double-buffered warp groups per SM (see tile_pass_body) are used when the pass has the default
// 11-qubit tile, enough tiles to fill the machine, an unpermuted write-back with one piece per thread at most
// (a thread recycles the shared-memory piece it owns), and tables that fit beside six tile buffers.
constexpr int kGroupTileQubits = 11;
constexpr u64 kGroupMinTiles = 2048;
constexpr size_t kGroupStaticSmem = 2 * 3 * 2 * 8 + 3 * 16 * 2 * 16 + 3 * 2 * 64 * 16 + 2 * 56 * 2 + 2 * kMaxTail;
constexpr int kGroupThreads = 3 * 256 + 3 * 32;  // three groups of 8 compute warps + one TMA producer warp each
constexpr size_t kMaxSmemPerBlock = 232448;  // 227 KiB (sm_100 opt-in limit, static + dynamic)

static bool g_attr_set[64] = {false};
// Opt in to > 48 KiB of dynamic shared memory, once per device.
static int ensure_attrs() {
  int dev = 0;
  QSV_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(QSV_ERANGE, "device ordinal out of range");
  if (!g_attr_set[dev]) {
    const int max_dyn = (int)((sizeof(double2) << QSV_MAX_TILE_QUBITS) + 64 * 1024);
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<512, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<256, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<128, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cudaFuncAttributes fa;
    QSV_CUDA(cudaFuncGetAttributes(&fa, tile_pass_kernel<768, 1, 3>));
    if (fa.sharedSizeBytes > kGroupStaticSmem + 64) return fail(QSV_ERANGE, "tile_pass_kernel<768,1,3>: static shared memory grew");
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<768, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kMaxSmemPerBlock - fa.sharedSizeBytes)));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<768, 1, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QSV_CUDA(cudaFuncGetAttributes(&fa, tile_pass_kernel<kGroupThreads, 1, 3, true>));
    if (fa.sharedSizeBytes > kGroupStaticSmem + 64)
      return fail(QSV_ERANGE, "tile_pass_kernel<864,1,3,true>: static shared memory grew");
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<kGroupThreads, 1, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kMaxSmemPerBlock - fa.sharedSizeBytes)));
    QSV_CUDA(cudaFuncSetAttribute(tile_pass_kernel<kGroupThreads, 1, 3, true>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QSV_CUDA(cudaFuncSetAttribute(dense_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * (sizeof(double2) << 12) + 4096 * sizeof(uint32_t))));
    g_attr_set[dev] = true;
  }
  return 0;
}

struct PassPlan {
  size_t smem;
  int threads, n_in, n_fused3;
  uint8_t op_flags[kMaxOps];  // dense ops: bit 0 = 128-bit accesses conflict-free, bit 1 = 64-bit loads conflict-free
  bool groups;  // three double-buffered warp groups per CTA (tile_pass_kernel<768, 1, 3>)
};


// Host-only lowering of a pass (no CUDA call): validates the arguments, fuses gates, builds the stages and
// every index table of the kernel into `d`.
static int lower_pass(PassDesc& d, PassPlan& plan, int n, const int32_t* tile_qubits, int n_tile,
                      const qsv_gate_t* gates, int n_gates, const int32_t* dest_qubits,
                      const qsv_gate_t* tail = nullptr, int n_tail = 0, double* tail_ent = nullptr,
                      const double* block8 = nullptr, const int32_t* block_qubits = nullptr,
                      const int32_t* fixed_qubits = nullptr, int n_fixed = 0) {
  if (n_tail < 0 || n_tail > kMaxTail) return fail(QSV_ERANGE, "qsv_apply_pass_tail: too many tail gates");
  if (n_tail > 0 && (!tail || !tail_ent)) return fail(QSV_EINVAL, "qsv_apply_pass_tail: null pointer");
  if (n < 1 || n > 40) return fail(QSV_EINVAL, "qsv_apply_pass: bad state / n_qubits");
  if (n_tile < 1 || n_tile > QSV_MAX_TILE_QUBITS || n_tile > n)
    return fail(QSV_ERANGE, "qsv_apply_pass: n_tile must be in [1, min(n_qubits, 12)]");
  if (n_gates < 0 || n_gates > kMaxOps) return fail(QSV_ERANGE, "qsv_apply_pass: too many gates");
  if (n - n_tile > 31) return fail(QSV_ERANGE, "qsv_apply_pass: grid too large");

  memset(&d, 0, sizeof(d));
  memset(plan.op_flags, 0, sizeof(plan.op_flags));
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
  for (int j = 0; j < b; j++) d.tbs[j] = (uint8_t)sorted[j];
  d.n_spread = b;
  if (n_fixed < 0 || n_fixed > 4 || (n_fixed > 0 && !fixed_qubits) || b + n_fixed > n)
    return fail(QSV_EINVAL, "qsv_apply_pass_sub: bad fixed qubits");
  for (int i = 0; i < n_fixed; i++) {
    const int q = fixed_qubits[i];
    bool clash = q < 0 || q >= n || ((d.fixed_mask >> q) & 1ull);
    for (int j = 0; j < b; j++) clash |= sorted[j] == q;
    if (clash) return fail(QSV_EINVAL, "qsv_apply_pass_sub: fixed qubits must be distinct and outside the tile");
    d.fixed_mask |= 1ull << q;
    d.tbs[d.n_spread++] = (uint8_t)q;
  }
  std::sort(d.tbs, d.tbs + d.n_spread);

  // ---- tile-local positions: 0..3 = the four lowest tile qubits (they form the contiguous 256-byte
  // pieces when they are physical qubits 0..3); the others are placed so that the two qubits of every
  // dense gate get different bank classes ----
  int pos_phys[16];
  {
    const int n_fixed = std::min(b, 4);
    for (int j = 0; j < n_fixed; j++) pos_phys[j] = sorted[j];
    uint64_t adj[64] = {0};
    int deg[64] = {0};
    for (int i = 0; i < n_gates; i++) {
      if (gates[i].kind != QSV_GATE_DENSE2 && gates[i].kind != QSV_GATE_CX) continue;
      const int a = gates[i].q0, c = gates[i].q1;
      if (a < 0 || a >= n || c < 0 || c >= n || a == c) continue;  // reported below
      if (!((adj[a] >> c) & 1ull)) { adj[a] |= 1ull << c; adj[c] |= 1ull << a; deg[a]++; deg[c]++; }
    }
    // Exhaustive search over the bank classes of the free positions (7 positions -> 210 distinct assignments; 8 ->
    // 560): minimise the number of dense gates whose two qubits share a class (2-way bank conflicts on every access
    // of that gate), then the number of qubit pairs of one class that are both neighbours of a third qubit (such a
    // gate finds no conflict-free third bit).  The greedy colouring this replaces left ~8 % of the dense ops of a
    // quantum-volume circuit with conflicts -- on the busiest pipe of the kernel (L1TEX ~80 %).
    const int m = b - n_fixed;
    if (m > 0) {
      int cap[3] = {0, 0, 0};
      for (int ps = n_fixed; ps < b; ps++) cap[bank_class(ps)]++;
      int fq[16];  // free qubits, highest degree first (prunes the search early)
      for (int i = 0; i < m; i++) fq[i] = sorted[n_fixed + i];
      std::sort(fq, fq + m, [&](int x, int y) { return deg[x] != deg[y] ? deg[x] > deg[y] : x < y; });
      int cur[16], best[16], best_cost = 1 << 30;
      auto cost_of = [&](int upto) {  // conflicts among the fixed qubits and fq[0..upto)
        int c = 0;
        for (int i = 0; i < upto; i++) {
          const int q = fq[i];
          for (int j = 0; j < n_fixed; j++)
            if (((adj[q] >> sorted[j]) & 1ull) && bank_class(j) == cur[i]) c += 4;
          for (int j = 0; j < i; j++)
            if (((adj[q] >> fq[j]) & 1ull) && cur[j] == cur[i]) c += 4;
        }
        return c;
      };
      // iterative depth-first enumeration
      int depth = 0;
      cur[0] = -1;
      int used_cap[3] = {0, 0, 0};
      while (depth >= 0) {
        if (cur[depth] >= 0) used_cap[cur[depth]]--;
        int c = cur[depth] + 1;
        while (c < 3 && used_cap[c] >= cap[c]) c++;
        if (c >= 3) { depth--; continue; }
        cur[depth] = c;
        used_cap[c]++;
        const int partial = cost_of(depth + 1);
        if (partial >= best_cost) continue;  // prune: costs only grow
        if (depth + 1 == m) {
          best_cost = partial;
          memcpy(best, cur, sizeof(int) * m);
          if (best_cost == 0) break;
          continue;
        }
        depth++;
        cur[depth] = -1;
      }
      // positions of a class are handed out in ascending order
      bool pos_used[16] = {false};
      for (int i = 0; i < m; i++) {
        for (int ps = n_fixed; ps < b; ps++)
          if (!pos_used[ps] && bank_class(ps) == best[i]) {
            pos_used[ps] = true;
            pos_phys[ps] = fq[i];
            break;
          }
      }
    }
  }
  int local_of[64];
  for (int q = 0; q < 64; q++) local_of[q] = -1;
  for (int j = 0; j < b; j++) {
    d.tb[j] = (uint8_t)pos_phys[j];
    local_of[pos_phys[j]] = j;
    d.cj[j] = (uint16_t)((1u << j) + pad_of_bit(j));
  }
  // bulk-copy piece size: the leading tile positions that are physical qubits 0..3 (contiguous in HBM)
  int pb = 0;
  while (pb < 4 && pb < b && pos_phys[pb] == pb) pb++;
  d.pb_ld = (uint32_t)pb;
  d.pb_st = (uint32_t)pb;
  if (b - pb <= 8)
    for (int p = 0; p < (1 << (b - pb)); p++) {
      uint64_t go = 0;
      uint32_t so = 0;
      for (int j = 0; j + pb < b; j++)
        if ((p >> j) & 1) { go |= 1ull << pos_phys[pb + j]; so += (1u << (pb + j)) + pad_of_bit(pb + j); }
      d.pc_g[p] = go * sizeof(double2);
      d.pc_s[p] = (uint16_t)(so * sizeof(double2));
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
    // a write-back piece must be contiguous in shared memory too: only unpermuted leading positions
    int ps = 0;
    while (ps < pb && d.src_of[ps] == ps) ps++;
    d.pb_st = (uint32_t)ps;
  }

  // ---- gates -> host ops (tile-local bits, matrices) ----
  HostOp hops_in[kMaxOps];
  double mats_in[kMaxIn];
  int moff = 0;
  for (int i = 0; i < n_gates; i++) {
    const qsv_gate_t& gsrc = gates[i];
    HostOp& op = hops_in[i];
    op.t2 = -1;
    op.foff = 0;
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
    if (gsrc.kind == QSV_GATE_CX && b >= 2) ndbl = 32;  // runs as a dense permutation on the tensor path (see below)
    if (moff + ndbl > kMaxIn) return fail(QSV_ERANGE, "qsv_apply_pass: matrix payload too large");
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
    if (gsrc.kind == QSV_GATE_CX && b >= 2) {
      // A lone CX as an element-wise conditional swap diverges (half of the lanes idle) and cost more than a dense
      // gate; as the 0/1 permutation matrix of basicaertools.py:70-72 it takes the DMMA path -- every product is
      // exact, so the result is the same bits -- and it fuses into 3-qubit blocks like any other gate.
      op_as_dense(QSV_GATE_CX, nullptr, mats_in + moff);
      op.kind = QSV_GATE_DENSE2;
      op.need = op.touch;
    } else if (gsrc.kind == QSV_GATE_DENSE1 && b >= 2) {
      memset(mats_in + moff, 0, sizeof(double) * 32);  // widened once the partner bit is known
      memcpy(mats_in + moff, gsrc.m, sizeof(double) * 8);
    } else {
      memcpy(mats_in + moff, gsrc.m, sizeof(double) * ndbl);
    }
    moff += ndbl;
  }
  d.flags = debug_flags();

  // ---- pair fusion: hops_in (program order) -> hops (fused, still a valid order) ----
  HostOp hops[kMaxOps];
  double mats[kMaxMat];
  const int n_in = n_gates;
  {
    const bool fuse_on = !(d.flags & 8u) && b >= 3;
    bool used[kMaxOps] = {false};
    int n_out = 0, mo = 0;
    // a finished 3-qubit block (`cur`: 8x8 complex, index bit k <-> tile bit ub[k]) -> kernel op + matrix in role order
    auto emit_block = [&](const double* cur, const int* ub, uint32_t cur_touch, HostOp& out) {
      // roles of the three bits: (t0, t1) should differ in bank class (see the slot layout)
      int role[3] = {0, 1, 2};
      bool found = false;
      for (int a = 0; a < 3 && !found; a++)
        for (int c = 0; c < 3 && !found; c++)
          if (a != c && differ(bank_class(ub[a]), bank_class(ub[c]))) {
            role[0] = a; role[1] = c; role[2] = 3 - a - c;
            found = true;
          }
      double* m8 = mats + mo;
      for (int r = 0; r < 8; r++)
        for (int c = 0; c < 8; c++) {
          int ro = 0, co = 0;  // index in the ub order
          for (int k = 0; k < 3; k++) {
            ro |= ((r >> k) & 1) << role[k];
            co |= ((c >> k) & 1) << role[k];
          }
          m8[2 * (8 * r + c)] = cur[2 * (8 * ro + co)];
          m8[2 * (8 * r + c) + 1] = cur[2 * (8 * ro + co) + 1];
        }
      out.kind = kGateDense3;
      out.t0 = ub[role[0]];
      out.t1 = ub[role[1]];
      out.t2 = ub[role[2]];
      out.touch = cur_touch;
      out.need = cur_touch;
      out.moff = mo;
      out.foff = 0;
      mo += 128;
    };
    if (block8) {  // a 3-qubit `unitary` given as such (qsv_apply_matrix, k = 3): it is the pass's first op
      int ub[3];
      uint32_t touch = 0;
      for (int k = 0; k < 3; k++) {
        if (block_qubits[k] < 0 || block_qubits[k] >= n || local_of[block_qubits[k]] < 0 ||
            ((touch >> local_of[block_qubits[k]]) & 1u))
          return fail(QSV_EINVAL, "qsv_apply_matrix: bad qubit list");
        ub[k] = local_of[block_qubits[k]];
        touch |= 1u << ub[k];
      }
      if (b < 3) return fail(QSV_ERANGE, "qsv_apply_matrix: a 3-qubit gate needs a tile of 3 qubits");
      emit_block(block8, ub, touch, hops[n_out++]);
    }
    for (int i = 0; i < n_in; i++) {
      if (used[i]) continue;
      used[i] = true;
      const HostOp& src = hops_in[i];
      HostOp& out = hops[n_out++];
      out = src;
      out.moff = mo;
      int ndbl_i;
      switch (src.kind) {
        case QSV_GATE_DENSE1: ndbl_i = b >= 2 ? 32 : 8; break;
        case QSV_GATE_DENSE2: ndbl_i = 32; break;
        case QSV_GATE_DIAG1: ndbl_i = 4; break;
        case QSV_GATE_DIAG2: ndbl_i = 8; break;
        default: ndbl_i = 0; break;
      }
      if (src.kind != QSV_GATE_DENSE2 || !fuse_on) {
        memcpy(mats + mo, mats_in + src.moff, sizeof(double) * ndbl_i);
        mo += ndbl_i;
        continue;
      }
      // current block: index bit k of `cur` <-> tile bit ub[k]
      int ub[3] = {src.t0, src.t1, -1}, nb = 2;
      uint32_t cur_touch = src.touch;
      double cur[128], emb[128], tmp[128];
      memcpy(cur, mats_in + src.moff, sizeof(double) * 32);
      uint32_t touched = 0;
      for (int j = i + 1; j < n_in; j++) {
        if (used[j]) continue;
        const HostOp& oj = hops_in[j];
        // any later gate that acts inside the block's bits is multiplied in (free: the block's cost does not
        // depend on its matrix); a 2-qubit gate that shares a bit may grow a 2-bit block to three bits
        const bool one_q = oj.kind == QSV_GATE_DENSE1 || oj.kind == QSV_GATE_DIAG1;
        const bool inside = (oj.touch & ~cur_touch) == 0;
        if (!(oj.touch & touched) && (oj.touch & cur_touch) && (inside || (!one_q && nb == 2)) &&
            popc_u(oj.touch | cur_touch) <= 3) {
          double dense[32];
          op_as_dense(oj.kind, mats_in + oj.moff, dense);
          if (one_q) {
            int pa = 0;
            for (int k = 0; k < nb; k++)
              if (ub[k] == oj.t0) pa = k;
            embed_gate1(dense, nb, pa, emb);
            cmat_mul(emb, cur, 1 << nb, tmp);
            memcpy(cur, tmp, sizeof(double) * 2 * (1 << nb) * (1 << nb));
            used[j] = true;
            continue;
          }
          if (popc_u(oj.touch | cur_touch) == 3 && nb == 2) {  // grow to three bits
            for (int q = 0; q < 32; q++) tmp[q] = cur[q];
            embed_gate(tmp, 3, 0, 1, cur);
            ub[2] = (oj.touch & ~cur_touch) == (1u << oj.t0) ? oj.t0 : oj.t1;
            nb = 3;
            cur_touch |= oj.touch;
          }
          int pa = -1, pb2 = -1;
          for (int k = 0; k < nb; k++) {
            if (ub[k] == oj.t0) pa = k;
            if (ub[k] == oj.t1) pb2 = k;
          }
          embed_gate(dense, nb, pa, pb2, emb);
          cmat_mul(emb, cur, 1 << nb, tmp);
          memcpy(cur, tmp, sizeof(double) * 2 * (1 << nb) * (1 << nb));
          used[j] = true;
          continue;
        }
        touched |= oj.touch;
        if ((touched & cur_touch) == cur_touch) break;  // every bit of the block is fenced off
      }
      if (nb == 2) {
        memcpy(mats + mo, cur, sizeof(double) * 32);
        mo += 32;
        continue;
      }
      emit_block(cur, ub, cur_touch, out);
    }
    n_gates = n_out;
    moff = mo;
  }

  // ---- stages ----
  int order[kMaxOps];
  uint32_t inner_masks[kMaxStages];
  int n_stages = build_stages(hops, n_gates, b, order, d.stages, inner_masks);
  if (n_stages == 0 && n_tail > 0) {  // the tail sweep borrows the element order of the last stage
    PassStage& st = d.stages[0];
    memset(&st, 0, sizeof(st));
    st.n_inner = (uint8_t)std::min(b, kInner);
    st.n_outer = (uint8_t)(b - st.n_inner);
    inner_masks[0] = (1u << st.n_inner) - 1u;
    n_stages = 1;
  }
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
        if (((inner >> j) & 1u) && j != op.t0 && differ(bank_class(j), bank_class(op.t0))) partner = j;
      for (int j = 0; j < b && partner < 0; j++)
        if (((inner >> j) & 1u) && j != op.t0 && bank_class(j) >= 0) partner = j;
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
    // the lane-varying gate bit (t0) should have a bank class: bit 3 takes the other role
    if (op.kind == QSV_GATE_DENSE2 && bank_class(op.t0) < 0 && bank_class(op.t1) >= 0) {
      swap_gate_bits(mats + op.moff);
      std::swap(op.t0, op.t1);
    }
  }
  memcpy(d.mats, mats, sizeof(double) * moff);
  int nfrag = 0;
  for (int k = 0; k < n_gates; k++) {
    sops[k].foff = nfrag;
    if (sops[k].kind == QSV_GATE_DENSE2 || sops[k].kind == QSV_GATE_DIAG1 || sops[k].kind == QSV_GATE_DIAG2) nfrag += 1;
    if (sops[k].kind == kGateDense3) nfrag += 4;
  }
  d.nfrag = nfrag;

  auto bvec = [&](int bit) { return bank_class(bit); };
  auto sw1 = [&](int bit) { return d.cj[bit]; };

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
    // outer bits shared with the previous stage keep their warp-id position
    st.sync_mask = 0;
    if (s > 0 && no == d.stages[s - 1].n_outer && !(d.flags & 2048u)) {
      const PassStage& prev = d.stages[s - 1];
      int placed[4] = {-1, -1, -1, -1};
      bool used_ob[16] = {false};
      for (int j = 0; j < no; j++)
        for (int q = 0; q < no; q++)
          if (!used_ob[q] && ob[q] == prev.obit[j]) { placed[j] = ob[q]; used_ob[q] = true; st.sync_mask |= (uint8_t)(1u << j); break; }
      int q = 0;
      for (int j = 0; j < no; j++) {
        if (placed[j] >= 0) continue;
        while (used_ob[q]) q++;
        placed[j] = ob[q];
        used_ob[q] = true;
      }
      for (int j = 0; j < no; j++) ob[j] = placed[j];
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
      op.foff = (uint16_t)h.foff;
      op.z0 = sw1(h.t0);
      op.z1 = h.t1 < 0 ? 0 : sw1(h.t1);
      op.z2 = h.t2 < 0 ? 0 : sw1(h.t2);
      if (h.kind != QSV_GATE_DENSE2 && h.kind != kGateDense3) continue;
      // free inner bits of the gate: f0 with {f0,t0,t1} independent (STS.128: 2 groups x 4 amplitudes per
      // quarter warp), f1 with {f0,f1,t0} independent (LDS.64: 4 groups x 2 amplitudes per half warp)
      int fb[16], nf = 0;
      for (int i = 0; i < ni; i++)
        if (ib[i] != h.t0 && ib[i] != h.t1 && ib[i] != h.t2) fb[nf++] = ib[i];
      const int v0 = bvec(h.t0), v1 = bvec(h.t1);
      int f0 = -1, f1 = -1;
      for (int i = 0; i < nf && f0 < 0; i++)
        if (independent3(bvec(fb[i]), v0, v1)) f0 = fb[i];
      for (int i = 0; i < nf && f0 < 0; i++)
        if (differ(bvec(fb[i]), v0)) f0 = fb[i];
      for (int i = 0; i < nf && f0 < 0; i++)
        if (bvec(fb[i]) >= 0) f0 = fb[i];
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
      plan.op_flags[k] = (uint8_t)((f0 >= 0 && independent3(bvec(f0), v0, v1) ? 1 : 0) |
                                   (f0 >= 0 && f1 >= 0 && independent3(bvec(f0), bvec(f1), v0) ? 2 : 0));
    }

    // ---- runs of consecutive diagonal gates -> lookup tables (QFT-like circuits are mostly such runs) ----
    const bool tab_on = !(d.flags & 64u);
    int k = st.first_op;
    const int k_end = st.first_op + st.n_ops;
    while (tab_on && k < k_end) {
      auto is_diag = [&](int idx) { return sops[idx].kind == QSV_GATE_DIAG1 || sops[idx].kind == QSV_GATE_DIAG2; };
      if (!is_diag(k)) { k++; continue; }
      uint32_t bits = 0;
      int k2 = k;
      while (k2 < k_end && is_diag(k2) && popc(bits | sops[k2].touch) <= kRunBits) bits |= sops[k2++].touch;
      const int n_run = k2 - k, nb = popc(bits);
      // a table costs about as much as two gates of the plain path
      if (n_run >= 2 && d.nruns < kMaxRuns && moff + (2 << nb) <= kMaxMat && d.run_entries + (1 << nb) <= kMaxRunEntries) {
        DiagRun& run = d.runs[d.nruns];
        int ub[kRunBits], nu = 0;
        for (int j = 0; j < b; j++)
          if ((bits >> j) & 1u) ub[nu++] = j;
        auto gather = [&](uint32_t li) {
          uint32_t x = 0;
          for (int j = 0; j < nu; j++) x |= ((li >> ub[j]) & 1u) << j;
          return (uint16_t)x;
        };
        run.first_op = (uint8_t)k;
        run.n_ops = (uint8_t)n_run;
        run.nbits = (uint8_t)nb;
        run.moff = (uint16_t)moff;
        run.toff = (uint16_t)d.run_entries;
        d.run_entries += 1 << nb;
        for (uint32_t x = 0; x < (1u << nb); x++) {
          uint32_t li = 0;
          for (int j = 0; j < nu; j++) li |= ((x >> j) & 1u) << ub[j];
          double re = 1.0, im = 0.0;
          for (int q = k; q < k2; q++) {  // same order as the gates would be applied
            const HostOp& h = sops[q];
            uint32_t fx = (li >> h.t0) & 1u;
            if (h.kind == QSV_GATE_DIAG2) fx |= ((li >> h.t1) & 1u) << 1;
            const double fr = d.mats[h.moff + 2 * fx], fi = d.mats[h.moff + 2 * fx + 1];
            const double nr = fr * re - fi * im, ni = fr * im + fi * re;
            re = nr;
            im = ni;
          }
          d.mats[moff + 2 * x] = re;
          d.mats[moff + 2 * x + 1] = im;
        }
        moff += 2 << nb;
        for (int e = 0; e < kWarps; e++) {
          uint32_t li = 0;
          for (int j = 0; j < st.n_outer; j++)
            if ((e >> j) & 1) li |= 1u << st.obit[j];
          run.xw[e] = gather(li);
        }
        for (int e = 0; e < 32; e++) {
          uint32_t li = 0;
          for (int j = 0; j < 5 && j < st.n_inner; j++)
            if ((e >> j) & 1) li |= 1u << st.ebit[j];
          run.xl[e] = gather(li);
        }
        for (int e = 0; e < 8; e++) {
          uint32_t li = 0;
          for (int j = 0; j < 3; j++)
            if (((e >> j) & 1) && 5 + j < st.n_inner) li |= 1u << st.ebit[5 + j];
          run.xit[e] = gather(li);
        }
        d.ops[k].ngl = (uint8_t)(d.nruns + 1);
        d.nruns++;
      }
      k = k2;
    }
  }

  // ---- diagonal tail: entries sorted by the tile bit they multiply (index b = scalars) ----
  if (n_tail > 0) {
    TailDesc& td = d.tail;
    int range_of[kMaxTail];
    uint8_t pk1[kMaxTail], pk2[kMaxTail];
    double ent[kMaxTail][8];
    int count[16] = {0};
    for (int i = 0; i < n_tail; i++) {
      const qsv_gate_t& g = tail[i];
      if (g.kind != QSV_GATE_DIAG1 && g.kind != QSV_GATE_DIAG2)
        return fail(QSV_EINVAL, "qsv_apply_pass_tail: tail gates must be diagonal (QSV_GATE_DIAG1 / QSV_GATE_DIAG2)");
      const bool two = g.kind == QSV_GATE_DIAG2;
      if (g.q0 < 0 || g.q0 >= n || (two && (g.q1 < 0 || g.q1 >= n || g.q1 == g.q0)))
        return fail(QSV_EINVAL, "qsv_apply_pass_tail: tail gate qubit out of range");
      if (((d.fixed_mask >> g.q0) & 1ull) || (two && ((d.fixed_mask >> g.q1) & 1ull)))
        return fail(QSV_EINVAL, "qsv_apply_pass_sub: a tail gate may not act on a fixed qubit");
      const int l0 = local_of[g.q0], l1 = two ? local_of[g.q1] : -1;
      if (two && l0 >= 0 && l1 >= 0)
        return fail(QSV_EINVAL, "qsv_apply_pass_tail: a 2-qubit tail gate needs a qubit outside the tile");
      // factor of the gate for (bit of q0, bit of q1): m[b0 + 2 b1] (complex)
      auto fac = [&](int b0, int b1, double* out) {
        const int idx = b0 + 2 * b1;
        out[0] = g.m[2 * idx];
        out[1] = g.m[2 * idx + 1];
      };
      pk1[i] = 0xff;
      pk2[i] = 0xff;
      if (!two) {
        if (l0 >= 0) {  // in the tile: [row any][bit]
          range_of[i] = l0;
          for (int r = 0; r < 2; r++)
            for (int bj = 0; bj < 2; bj++) fac(bj, 0, ent[i] + 2 * (2 * r + bj));
        } else {        // scalar: [row = bit of q0][*]
          range_of[i] = b;
          pk1[i] = (uint8_t)g.q0;
          for (int r = 0; r < 2; r++)
            for (int r2 = 0; r2 < 2; r2++) fac(r, 0, ent[i] + 2 * (2 * r + r2));
        }
      } else if (l0 >= 0) {  // q0 in the tile, q1 selects the row
        range_of[i] = l0;
        pk1[i] = (uint8_t)g.q1;
        for (int r = 0; r < 2; r++)
          for (int bj = 0; bj < 2; bj++) fac(bj, r, ent[i] + 2 * (2 * r + bj));
      } else if (l1 >= 0) {  // q1 in the tile, q0 selects the row
        range_of[i] = l1;
        pk1[i] = (uint8_t)g.q0;
        for (int r = 0; r < 2; r++)
          for (int bj = 0; bj < 2; bj++) fac(r, bj, ent[i] + 2 * (2 * r + bj));
      } else {               // both outside: scalar [row = q0][row2 = q1]
        range_of[i] = b;
        pk1[i] = (uint8_t)g.q0;
        pk2[i] = (uint8_t)g.q1;
        for (int r = 0; r < 2; r++)
          for (int r2 = 0; r2 < 2; r2++) fac(r, r2, ent[i] + 2 * (2 * r + r2));
      }
      count[range_of[i]]++;
    }
    int at[17];
    at[0] = 0;
    for (int j = 0; j <= b; j++) at[j + 1] = at[j] + count[j];
    for (int j = 0; j <= b + 1; j++) td.start[j] = (uint8_t)at[j];
    int fill[16];
    for (int j = 0; j <= b; j++) fill[j] = at[j];
    for (int i = 0; i < n_tail; i++) {  // stable: gates of one range keep their order
      const int e = fill[range_of[i]]++;
      td.pk[e] = pk1[i];
      td.pk2[e] = pk2[i];
      memcpy(tail_ent + 8 * e, ent[i], sizeof(double) * 8);
    }
    td.n = n_tail;
    // table groups: the tile bits that have factors, at most 6 per table
    int nj = 0, jb[16];
    for (int j = 0; j < b; j++)
      if (count[j]) jb[nj++] = j;
    td.na = (uint8_t)std::min(nj, 6);
    td.nb = (uint8_t)(nj - td.na);
    for (int i = 0; i < td.na; i++) td.abit[i] = (uint8_t)jb[i];
    for (int i = 0; i < td.nb; i++) td.bbit[i] = (uint8_t)jb[6 + i];
    const PassStage& st = d.stages[n_stages - 1];
    for (int g2 = 0; g2 < 2; g2++) {
      const uint8_t* gb = g2 ? td.bbit : td.abit;
      const int ng = g2 ? td.nb : td.na;
      auto gather = [&](uint32_t li) {
        uint32_t x = 0;
        for (int i = 0; i < ng; i++) x |= ((li >> gb[i]) & 1u) << i;
        return (uint16_t)x;
      };
      for (int e = 0; e < kWarps; e++) {
        uint32_t li = 0;
        for (int j = 0; j < st.n_outer; j++)
          if ((e >> j) & 1) li |= 1u << st.obit[j];
        td.xw[g2][e] = gather(li);
      }
      for (int e = 0; e < 32; e++) {
        uint32_t li = 0;
        for (int j = 0; j < 5 && j < st.n_inner; j++)
          if ((e >> j) & 1) li |= 1u << st.ebit[j];
        td.xl[g2][e] = gather(li);
      }
      for (int e = 0; e < 8; e++) {
        uint32_t li = 0;
        for (int j = 0; j < 3; j++)
          if (((e >> j) & 1) && 5 + j < st.n_inner) li |= 1u << st.ebit[5 + j];
        td.xit[g2][e] = gather(li);
      }
    }
  }

  size_t tile_slots = (size_t)1 << b;
  for (int j = 4; j < b; j++) tile_slots += pad_of_bit(j);
  const size_t smem = tile_slots * sizeof(double2) + (size_t)nfrag * 64 * sizeof(double) +
                      (size_t)n_gates * (sizeof(OpRec) + 32 * 4) + (size_t)n_stages * ((kWarps + 32 + 8) * 4) + 16 +
                      (size_t)d.run_entries * sizeof(double2) + (size_t)d.nruns * 56 * sizeof(uint16_t);
  d.n_tiles = 1u << (n - b - n_fixed);
  // one warp per 2^8 amplitudes: 512 threads for b = 12, 256 for b = 11, ...; at least one warp
  plan.smem = smem;
  plan.threads = std::max(32, std::min(kMaxPT, 1 << std::max(0, b - 3)));
  plan.groups = false;
  if (b == kGroupTileQubits && d.n_tiles >= kGroupMinTiles && !(d.flags & 16u) && d.pb_ld == d.pb_st &&
      ((size_t)1 << (b - d.pb_ld)) <= 256) {
    bool ident = true;
    for (int j = 0; j < b; j++) ident &= d.src_of[j] == j;
    const size_t smem3 = smem + 5 * tile_slots * sizeof(double2);
    if (ident && smem3 + kGroupStaticSmem + 64 <= kMaxSmemPerBlock) {
      plan.groups = true;
      plan.smem = smem3;
      plan.threads = kGroupThreads;
    }
  }
  plan.n_in = n_in;
  plan.n_fused3 = 0;
  for (int k = 0; k < n_gates; k++) plan.n_fused3 += d.ops[k].kind == kGateDense3;
  return 0;
}

static int launch_pass(double* sv, int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                       int n_gates, cudaStream_t stream, const int32_t* dest_qubits = nullptr,
                       const qsv_gate_t* tail = nullptr, int n_tail = 0, void* dev_scratch = nullptr,
                       const double* block8 = nullptr, const int32_t* block_qubits = nullptr,
                       const int32_t* fixed_qubits = nullptr, int n_fixed = 0, uint32_t fixed_pattern = 0,
                       int max_ctas = 0) {
  if (!sv) return fail(QSV_EINVAL, "qsv_apply_pass: bad state / n_qubits");
  if (n_tail > 0 && !dev_scratch) return fail(QSV_EINVAL, "qsv_apply_pass_tail: null device scratch");
  static thread_local PassDesc d;
  static thread_local double tail_ent[kMaxTail * 8];
  PassPlan plan;
  if (int rc = lower_pass(d, plan, n, tile_qubits, n_tile, gates, n_gates, dest_qubits, tail, n_tail, tail_ent, block8,
                          block_qubits, fixed_qubits, n_fixed))
    return rc;
  // a sub-block pass walks only the tiles whose fixed bits read `fixed_pattern`: their value goes into the pointer
  for (int i = 0; i < n_fixed; i++)
    if ((fixed_pattern >> i) & 1u) sv += 2 * ((size_t)1 << fixed_qubits[i]);
  const unsigned cta_cap = max_ctas > 0 ? (unsigned)max_ctas : 0xffffffffu;
  if (int rc = ensure_attrs()) return rc;
  if (n_tail > 0)  // small pageable copy: staged by the driver before the call returns, ordered on the stream
    QSV_CUDA(cudaMemcpyAsync(dev_scratch, tail_ent, sizeof(double) * 8 * (size_t)n_tail, cudaMemcpyHostToDevice, stream));
  const double2* d_tail = reinterpret_cast<const double2*>(dev_scratch);
  const size_t smem = plan.smem;
  const int threads = plan.threads;
  if (plan.groups) {
    const unsigned grid3 = std::min(std::min<unsigned>((d.n_tiles + 2) / 3, (unsigned)sm_count()), cta_cap);
    if (d.flags & 512u)  // A/B: compute warps issue their own bulk copies (round-1 shape)
      tile_pass_kernel<768, 1, 3><<<grid3, 768, smem, stream>>>(reinterpret_cast<double2*>(sv), d, d_tail);
    else
      tile_pass_kernel<kGroupThreads, 1, 3, true><<<grid3, kGroupThreads, smem, stream>>>(reinterpret_cast<double2*>(sv), d,
                                                                                         d_tail);
    QSV_LAUNCHED();
    return 0;
  }
  int per_sm = 1;
  if (threads >= 512) {
    QSV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_pass_kernel<512, 2>, threads, smem));
  } else if (threads >= 256) {
    QSV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_pass_kernel<256, 4>, threads, smem));
  } else {
    QSV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_pass_kernel<128, 8>, threads, smem));
  }
  if (per_sm < 1) return fail(QSV_ERANGE, "qsv_apply_pass: pass does not fit in shared memory");
  const unsigned grid = std::min(std::min<unsigned>(d.n_tiles, (unsigned)per_sm * (unsigned)sm_count()),
                                 cta_cap == 0xffffffffu ? cta_cap : cta_cap * (unsigned)per_sm);
  if (threads >= 512)
    tile_pass_kernel<512, 2><<<grid, threads, smem, stream>>>(reinterpret_cast<double2*>(sv), d, d_tail);
  else if (threads >= 256)
    tile_pass_kernel<256, 4><<<grid, threads, smem, stream>>>(reinterpret_cast<double2*>(sv), d, d_tail);
  else
    tile_pass_kernel<128, 8><<<grid, threads, smem, stream>>>(reinterpret_cast<double2*>(sv), d, d_tail);
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

int qsv_apply_pass_tail(double* sv, int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                        int n_gates, const int32_t* dest_qubits, const qsv_gate_t* tail, int n_tail,
                        void* dev_scratch, void* stream) {
  if (!tile_qubits || (n_gates > 0 && !gates)) return fail(QSV_EINVAL, "qsv_apply_pass_tail: null pointer");
  return launch_pass(sv, n, tile_qubits, n_tile, gates, n_gates, (cudaStream_t)stream, dest_qubits, tail, n_tail,
                     dev_scratch);
}

int qsv_apply_pass_sub(double* sv, int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates, int n_gates,
                       const qsv_gate_t* tail, int n_tail, void* dev_scratch, const int32_t* fixed_qubits, int n_fixed,
                       uint32_t fixed_pattern, int max_sms, void* stream) {
  if (!tile_qubits || (n_gates > 0 && !gates)) return fail(QSV_EINVAL, "qsv_apply_pass_sub: null pointer");
  if (n_fixed > 0 && fixed_pattern >= (1u << n_fixed)) return fail(QSV_EINVAL, "qsv_apply_pass_sub: bad pattern");
  return launch_pass(sv, n, tile_qubits, n_tile, gates, n_gates, (cudaStream_t)stream, nullptr, tail, n_tail, dev_scratch,
                     nullptr, nullptr, fixed_qubits, n_fixed, fixed_pattern, max_sms);
}

size_t qsv_tail_scratch_bytes(void) { return sizeof(double) * 8 * kMaxTail; }

int qsv_alloc(void** dev_out, size_t bytes) {
  if (!dev_out || bytes == 0) return fail(QSV_EINVAL, "qsv_alloc: bad arguments");
  QSV_CUDA(cudaMalloc(dev_out, bytes));
  return 0;
}

int qsv_free(void* dev_ptr) {
  if (!dev_ptr) return 0;
  QSV_CUDA(cudaFree(dev_ptr));
  return 0;
}

int qsv_copy_d2d(void* dev_dst, const void* dev_src, size_t bytes, void* stream) {
  if (!dev_dst || !dev_src) return fail(QSV_EINVAL, "qsv_copy_d2d: null pointer");
  QSV_CUDA(cudaMemcpyAsync(dev_dst, dev_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int qsv_debug_set_flags(uint32_t flags) {
#ifndef QSV_DEBUG_BUILD
  if (flags & kResultChangingFlags)
    return fail(QSV_EINVAL, "qsv_debug_set_flags: bits 1, 2, 4 (no stores / loads / gates) need the debug library");
#endif
  g_debug_flags = flags;
  g_debug_flags_set = true;
  g_p2p_unroll = ((flags >> 12) & 3u) == 1u ? 8 : ((flags >> 12) & 3u) == 2u ? 2 : ((flags >> 12) & 3u) == 3u ? 1 : 4;
  g_p2p_chunked = (flags >> 14) & 1u;
#ifdef QSV_DEBUG_BUILD
  g_p2p_mode = (flags >> 15) & 3u;
#endif
  return 0;
}

int qsv_debug_profile(uint64_t* host_out, size_t n_words) {
#ifdef QSV_PROFILE
  if (!host_out || n_words > sizeof(g_prof) / 8) return fail(QSV_EINVAL, "qsv_debug_profile: bad arguments");
  QSV_CUDA(cudaDeviceSynchronize());
  QSV_CUDA(cudaMemcpyFromSymbol(host_out, g_prof, n_words * 8));
  return 0;
#else
  (void)host_out;
  (void)n_words;
  return fail(QSV_ERANGE, "qsv_debug_profile: this library was built without -DQSV_PROFILE");
#endif
}

int qsv_pass_info(int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates, int n_gates,
                  const int32_t* dest_qubits, int32_t* info) {
  if (!tile_qubits || !info || (n_gates > 0 && !gates)) return fail(QSV_EINVAL, "qsv_pass_info: null pointer");
  static thread_local PassDesc d;
  PassPlan plan;
  if (int rc = lower_pass(d, plan, n, tile_qubits, n_tile, gates, n_gates, dest_qubits)) return rc;
  info[0] = d.nops;
  info[1] = d.nstages;
  info[2] = plan.n_fused3;
  info[3] = (int32_t)plan.smem;
  info[4] = plan.threads;
  info[5] = (int32_t)d.pb_ld;
  info[6] = (int32_t)d.pb_st;
  info[7] = plan.n_in;
  return 0;
}

int qsv_pass_ops(int n, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates, int n_gates,
                 const int32_t* dest_qubits, int32_t* ops_out, int max_ops) {
  if (!tile_qubits || !ops_out || (n_gates > 0 && !gates)) return fail(QSV_EINVAL, "qsv_pass_ops: null pointer");
  static thread_local PassDesc d;
  PassPlan plan;
  if (int rc = lower_pass(d, plan, n, tile_qubits, n_tile, gates, n_gates, dest_qubits)) return rc;
  int o = 0;
  for (int sidx = 0; sidx < d.nstages; sidx++)
    for (int k = 0; k < d.stages[sidx].n_ops && o < max_ops; k++, o++) {
      const PassOp& op = d.ops[o];
      ops_out[4 * o] = op.kind;
      ops_out[4 * o + 1] = sidx;
      ops_out[4 * o + 2] = d.tb[op.t0] | (d.tb[op.t1] << 8);
      ops_out[4 * o + 3] = plan.op_flags[o];
    }
  return o;
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
  if (k == 3) {
    // one fused 3-qubit block of the pass kernel (8x8 complex in the 3-multiplication DMMA form): the path that
    // pairs of 2-qubit gates take, here with the caller's matrix
    const int nt = default_tile(n, qubits, k, std::min(n, 11), tile);
    if (nt < 0) return fail(QSV_EINVAL, "qsv_apply_matrix: bad qubit list");
    return launch_pass(sv, n, tile, nt, nullptr, 0, s, nullptr, nullptr, 0, nullptr, host_matrix, qubits);
  }
  if (!dev_mat_scratch) return fail(QSV_EINVAL, "qsv_apply_matrix: k >= 4 needs dev_mat_scratch");
  if (k <= 7 && n >= 11)  // complex GEMM over the tile's groups on the FP64 tensor path (dense_k.cu)
    return launch_dense_k_dmma(sv, n, qubits, k, host_matrix, dev_mat_scratch, s);
  // states smaller than one tile, and k >= 8 (the 2^k x 2^k matrix no longer fits beside the tile): scalar kernel
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
