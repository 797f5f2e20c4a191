/*
 * qsv.h -- C-ABI of libqsv, the B200 (sm_100a) statevector engine.
 *
 * This is the drop-in boundary for the one data-parallel hot path of Qiskit Terra 0.18's BasicAer
 * simulators.  The reference has NO native interface on this path: every entry point below replaces
 * a numpy call made from Python in
 *     qiskit/providers/basicaer/qasm_simulator.py      (QasmSimulatorPy / StatevectorSimulatorPy)
 *     qiskit/providers/basicaer/unitary_simulator.py   (UnitarySimulatorPy)
 * and is what a ctypes binding inside those files would call.  Each prototype
 * cites the reference lines it stands in for.
 *
 * Conventions
 *   - A state is 2^n complex128 amplitudes in device memory, interleaved (re, im) float64, i.e.
 *     numpy complex128 layout.  Qubit q is bit q of the flat amplitude index (the C-order flattening
 *     of the reference's rank-n tensor whose axis n-1-q is qubit q; basicaertools.py:132-180).
 *   - Matrices are row-major, interleaved (re, im) float64; for a k-qubit gate on qubits[0..k-1],
 *     qubits[0] is the LEAST significant bit of the row/column index (qasm_simulator.py:145-161).
 *   - Plain pointers and sizes only.  `sv` / `probs` / `ws` are DEVICE pointers owned by the caller;
 *     `host_*` are HOST pointers owned by the caller.  `stream` is a cudaStream_t passed as void*
 *     (NULL = default stream).  The library keeps no per-state handle and allocates no device memory behind the
 *     caller's back; qsv_alloc / qsv_free are there for callers that have no allocator of their own.
 *   - Every function returns 0 on success, a negative QSV_E* code for argument errors, or a positive
 *     cudaError_t.  qsv_last_error() returns a thread-local message for the last failure.
 *   - One host thread per state at a time (the reference backends are not re-entrant either;
 *     qasm_simulator.py:118-131 keeps per-run state on the instance).
 */
#ifndef QSV_H
#define QSV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSV_VERSION 100

#define QSV_EINVAL (-1)   /* bad argument */
#define QSV_ERANGE (-2)   /* size outside what the kernels support */
#define QSV_ENORM (-3)    /* probabilities do not sum to 1 (numpy choice() ValueError) */

#define QSV_MAX_TILE_QUBITS 12 /* two buffers of 2^12 amplitudes * 16 B = 128 KiB of shared memory */
#define QSV_MAX_PASS_GATES 48
#define QSV_MAX_DENSE_K 12

/* Gate kinds inside a fused pass. */
enum {
  QSV_GATE_DENSE1 = 1, /* 2x2 dense on q0: m = 4 complex, row-major                    */
  QSV_GATE_DENSE2 = 2, /* 4x4 dense on (q0,q1), q0 = LSB of the index: m = 16 complex  */
  QSV_GATE_DIAG1 = 3,  /* diag(m[0], m[1]) on q0 (u1, rz, measure collapse)            */
  QSV_GATE_DIAG2 = 4,  /* diag(m[0..3]) on (q0,q1)                                     */
  QSV_GATE_CX = 5      /* q0 = control, q1 = target (basicaertools.py:70-72)           */
};

typedef struct qsv_gate {
  int32_t kind;
  int32_t q0;
  int32_t q1;
  int32_t reserved;
  double m[32]; /* interleaved re,im */
} qsv_gate_t;

int qsv_version(void);
const char* qsv_last_error(void);

/* psi = e^{i phase} |0...0>.  Replaces _initialize_statevector (qasm_simulator.py:319-328) fused with
 * the global-phase multiply (qasm_simulator.py:522).  Writes 16*2^n bytes. */
int qsv_init_basis0(double* sv, int n_qubits, double phase_re, double phase_im, void* stream);

/* U = e^{i phase} * identity as a 2n-"qubit" vector, flat index row*2^n + col.  Replaces
 * _initialize_unitary (unitary_simulator.py:190-199). */
int qsv_init_identity(double* sv, int n_qubits, double phase_re, double phase_im, void* stream);

/* psi <- host amplitudes (initial_statevector option, qasm_simulator.py:299-311,326) times
 * (phase_re + i phase_im).  n_amps must be 2^n. */
int qsv_upload(double* sv, int n_qubits, const double* host_amps, size_t n_amps, double phase_re,
               double phase_im, void* stream);

/* psi *= (re + i im), elementwise. */
int qsv_scale(double* sv, int n_qubits, double re, double im, void* stream);

/* One cache-blocked pass: every CTA stages the 2^n_tile amplitudes spanned by `tile_qubits` in shared
 * memory (one HBM read), applies `gates` in order there, and writes the tile back (one HBM write).
 * All gate qubits must be in tile_qubits; n_tile <= min(n_qubits, QSV_MAX_TILE_QUBITS).  For full
 * memory coalescing tile_qubits should contain qubits 0..3 (256-byte runs = one TMA bulk copy each).  Replaces a RUN of
 * consecutive _add_unitary calls (qasm_simulator.py:145-161, loop :526-559), each of which is one
 * np.einsum pass over the whole vector in the reference. */
int qsv_apply_pass(double* sv, int n_qubits, const int32_t* tile_qubits, int n_tile,
                   const qsv_gate_t* gates, int n_gates, void* stream);

/* qsv_apply_pass followed, for free, by a permutation of the tile qubits: on write-back the content of
 * qubit tile_qubits[i] lands on qubit dest_qubits[i] (dest_qubits is a permutation of tile_qubits; NULL =
 * identity).  The caller tracks the resulting logical -> physical qubit map.  Used by the planner to
 * keep the qubits the next gates need in the always-resident low bits, and to undo the map before a
 * read-out.  (New: the reference never moves qubits; its einsum addresses them in place.) */
int qsv_apply_pass_remap(double* sv, int n_qubits, const int32_t* tile_qubits, int n_tile,
                         const qsv_gate_t* gates, int n_gates, const int32_t* dest_qubits, void* stream);

/* qsv_apply_pass_remap plus a DIAGONAL TAIL: after `gates`, and inside the same pass (no extra HBM traffic),
 * the `n_tail` diagonal gates of `tail` (QSV_GATE_DIAG1 / QSV_GATE_DIAG2) are applied.  Their qubits need
 * NOT be tile qubits: outside the tile a qubit is constant per tile, so per tile the whole tail collapses to
 * one factor pair per tile qubit and the kernel applies it with one or two table look-ups per amplitude,
 * however many gates the tail has.  A 2-qubit tail gate must have at least one qubit outside the tile (gates
 * inside the tile belong in `gates`); n_tail <= QSV_MAX_TAIL_GATES.  `dev_scratch` = device buffer of
 * qsv_tail_scratch_bytes() bytes (only read by this pass; reusable by the next call on the same stream).
 * Replaces the runs of _add_unitary calls (qasm_simulator.py:145-161) that a controlled-phase ladder makes in
 * the reference -- e.g. the 435 cp gates of a 30-qubit QFT, each five einsum passes there.  (The tail gates
 * are applied after ALL of `gates`: the caller guarantees that this order is equivalent to program order.) */
#define QSV_MAX_TAIL_GATES 192
int qsv_apply_pass_tail(double* sv, int n_qubits, const int32_t* tile_qubits, int n_tile,
                        const qsv_gate_t* gates, int n_gates, const int32_t* dest_qubits,
                        const qsv_gate_t* tail, int n_tail, void* dev_scratch, void* stream);
size_t qsv_tail_scratch_bytes(void);

/* qsv_apply_pass_tail restricted to a SUB-BLOCK of the state: only the amplitudes whose qubits fixed_qubits[i] read
 * bit i of fixed_pattern are touched (the fixed qubits are neither tile qubits nor acted on by any gate), on at most
 * max_sms SMs (0 = all).  The multi-GPU engine applies the passes that commute with a qubit-swap exchange sub-block by
 * sub-block while the other sub-blocks are in flight over NVLink, the exchange kernel running on the SMs left free
 * (dist.ShardedState).  Same gates / tail / scratch conventions as qsv_apply_pass_tail; no write-back permutation. */
int qsv_apply_pass_sub(double* sv, int n_qubits, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                       int n_gates, const qsv_gate_t* tail, int n_tail, void* dev_scratch,
                       const int32_t* fixed_qubits, int n_fixed, uint32_t fixed_pattern, int max_sms, void* stream);

/* Host-only introspection of the lowering qsv_apply_pass[_remap] would do for these arguments (no GPU is
 * touched, nothing is launched): info[0] = kernel ops after gate fusion (2-qubit gates that share a qubit
 * are merged into 3-qubit blocks), info[1] = stages (block-wide barriers + 1), info[2] = fused 3-qubit
 * blocks, info[3] = dynamic shared memory (bytes), info[4] = threads per CTA, info[5] / info[6] = log2 of
 * the amplitudes per TMA bulk-copy piece of the load / the write-back, info[7] = gates passed in.
 * Same argument validation and error codes as qsv_apply_pass_remap.  (New: planning aid and CPU-side test
 * hook; no reference counterpart.) */
int qsv_pass_info(int n_qubits, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates,
                  int n_gates, const int32_t* dest_qubits, int32_t* info);

/* Host-only, like qsv_pass_info: the kernel ops of the lowered pass in execution order, four ints per op --
 * kind (QSV_GATE_*; 6 = fused 3-qubit block), stage, physical qubits q0 | q1 << 8 of the op's two lane bits,
 * flags (dense ops: bit 0 = its 128-bit shared-memory accesses are bank-conflict free, bit 1 = its 64-bit loads
 * are).  Returns the number of ops written (<= max_ops) or a negative error.  Planning / profiling aid. */
int qsv_pass_ops(int n_qubits, const int32_t* tile_qubits, int n_tile, const qsv_gate_t* gates, int n_gates,
                 const int32_t* dest_qubits, int32_t* ops_out, int max_ops);

/* psi <- (G on qubits) psi for a dense k-qubit matrix, k >= 1.  Replaces one _add_unitary call
 * (qasm_simulator.py:145-161; `unitary` instruction :543-546).  k <= 2 goes through qsv_apply_pass; k = 3 is one
 * fused 3-qubit block of the pass kernel; 4 <= k <= 7 (states of at least 11 qubits) is a complex GEMM over the
 * tile's groups on the FP64 tensor path (DMMA, 3-multiplication form); larger k and smaller states use a scalar
 * staged kernel.  k >= 4 needs `dev_mat_scratch`, a device buffer of at least 4*4^k doubles (the matrix, re-arranged
 * into tensor-core fragments, is copied there). */
int qsv_apply_matrix(double* sv, int n_qubits, const int32_t* qubits, int k, const double* host_matrix,
                     double* dev_mat_scratch, void* stream);

/* (p0, p1) of one qubit: p_b = sum over amplitudes with bit `qubit` == b of |psi|^2.  Replaces the
 * np.sum(np.abs(psi)**2, axis=...) in _get_measure_outcome (qasm_simulator.py:173-176).  `ws` needs
 * qsv_reduce_workspace_bytes() bytes.  Synchronises the stream; result in host_out[2]. */
size_t qsv_reduce_workspace_bytes(void);
int qsv_prob_qubit(const double* sv, int n_qubits, int qubit, void* ws, double* host_out, void* stream);

/* Projective collapse + renormalisation: the diag / reset update matrices that _add_qasm_measure
 * (qasm_simulator.py:247-253) and _add_qasm_reset (:265-273) pass to _add_unitary.
 * is_reset = 0: amplitudes with bit != outcome are zeroed, the rest scaled by 1/sqrt(prob).
 * is_reset = 1: as above, and for outcome 1 the |1> half is moved to |0>. */
int qsv_collapse(double* sv, int n_qubits, int qubit, int outcome, double prob, int is_reset,
                 void* stream);

/* Shot-batched mid-circuit operations.  The reference simulates a circuit with mid-circuit
 * measure / reset / conditionals once per shot (qasm_simulator.py:512-531); the host instead keeps 2^n_shot_qubits
 * shots side by side in ONE vector of n_total = n + n_shot_qubits qubits (shot number = the top index bits): gates act
 * on all shots through the ordinary passes, qsv_marginal over (qubit, shot qubits) gives every shot's outcome
 * probabilities in one launch, and these two entry points do the per-shot parts:
 *   qsv_collapse_batch  _add_qasm_measure / _add_qasm_reset (:247-253, :265-273) with a different outcome and
 *                       1/sqrt(p) per shot (host arrays of 2^n_shot_qubits entries)
 *   qsv_apply_masked    a 1- or 2-qubit gate on the shots whose mask byte is non-zero (conditional gates, :527-540)
 * dev_ws: qsv_batch_workspace_bytes(n_shot_qubits) bytes of device memory.  Both synchronise the stream. */
size_t qsv_batch_workspace_bytes(int n_shot_qubits);
int qsv_collapse_batch(double* sv, int n_total, int n_shot_qubits, int qubit, const unsigned char* host_outcomes,
                       const double* host_scales, int is_reset, void* dev_ws, void* stream);
int qsv_apply_masked(double* sv, int n_total, int n_shot_qubits, const qsv_gate_t* gate, const unsigned char* host_mask,
                     void* dev_ws, void* stream);

/* Marginal probabilities of m measured qubits: probs[s] = sum |psi_i|^2 over all i whose bit
 * measured_sorted[pos] equals bit pos of s.  Replaces np.sum(np.abs(psi)**2, axis=...) in
 * _add_sample_measure (qasm_simulator.py:202-209).  dev_probs holds 2^m doubles. */
size_t qsv_marginal_workspace_bytes(int n_qubits, int m);
int qsv_marginal(const double* sv, int n_qubits, const int32_t* measured_sorted, int m,
                 double* dev_probs, void* ws, void* stream);

/* Draw `shots` samples: the RandomState.choice(range(2^m), shots, p) of _add_sample_measure
 * (qasm_simulator.py:213), i.e. cdf = cumsum(p); cdf /= cdf[-1]; idx = searchsorted(cdf, u, 'right')
 * for the caller-supplied uniforms u (RandomState.random_sample(shots)).
 *   src            : from_amplitudes != 0 -> 2^log2_bins complex amplitudes, p_i = |psi_i|^2
 *                    (all qubits measured); else 2^log2_bins doubles (a marginal from qsv_marginal).
 *   exact_order    : flags.  Bit 0 set -> the CDF is numpy's cumsum bit for bit (one float64 running sum, left to
 *                    right, evaluated in parallel; p_i = np.abs(psi_i)**2 with numpy's rounding), so samples
 *                    equal the reference's; clear -> blocked parallel scan.  Bit 1 set ->
 *                    skip the sum-to-1 check (a shard of a multi-GPU state sums to less than 1).
 *   host_total     : receives cdf[-1] before normalisation (must be ~1).
 * Returns QSV_ENORM if |total - 1| > sqrt(eps) as numpy's choice() does.  Synchronises the stream. */
size_t qsv_sample_workspace_bytes(int log2_bins, int shots);
int qsv_sample(const double* src, int from_amplitudes, int log2_bins, const double* host_uniforms,
               int shots, int64_t* host_out_idx, int exact_order, void* ws, double* host_total,
               void* stream);

/* The phases of qsv_sample's exact-order path, callable one by one so that the CDF can run across the SHARDS of
 * a sharded state in the reference's bin order (numpy's cumsum is one running sum over all 2^n bins; shard r
 * continues where shard r-1 stopped).  All use the workspace of qsv_sample_workspace_bytes(log2_bins, shots),
 * which carries the state from one phase to the next; src / from_amplitudes / log2_bins as in qsv_sample.
 *   qsv_cdf_prepare   blocked sums of the 1024-bin chunks and their prefix; host_approx_total = the array's sum
 *                     (rounded differently from the running sum: only used to steer the next phase).
 *   qsv_cdf_classify  per chunk: can the running sum be advanced by an integer number of ulps?  approx_offset =
 *                     approximate running sum before this array (sum of the previous shards' approx totals).
 *   qsv_cdf_chain     the running sum itself, bit for bit numpy's, starting from exact_start (0 for the first
 *                     shard, else the previous shard's host_exact_end); serial in the shards, milliseconds each.
 *   qsv_cdf_search    searchsorted(cdf / total, u, 'right') restricted to this array: bin index, or -1 if the
 *                     shot falls before the array (exact_start / total > u), -2 if after it.
 * Together they replace RandomState.choice in _add_sample_measure (qasm_simulator.py:213) for n > 33 qubits.
 * Each synchronises the stream. */
int qsv_cdf_prepare(const double* src, int from_amplitudes, int log2_bins, void* ws, double* host_approx_total,
                    void* stream);
int qsv_cdf_classify(const double* src, int from_amplitudes, int log2_bins, double approx_offset, void* ws,
                     void* stream);
int qsv_cdf_chain(const double* src, int from_amplitudes, int log2_bins, double exact_start, void* ws,
                  double* host_exact_end, void* stream);
int qsv_cdf_search(const double* src, int from_amplitudes, int log2_bins, double exact_start, double total,
                   const double* host_uniforms, int shots, int64_t* host_out_idx, void* ws, void* stream);

/* <psi| P |psi> for the Pauli string with Z (or Y) on the bits of z_mask and X (or Y) on the bits of x_mask,
 * times (phase_re + i phase_im) for the with-X case.  Replaces the serial Cython loops
 * expval_pauli_no_x / expval_pauli_with_x (qiskit/quantum_info/states/cython/exp_value.pyx:37-94, called
 * from Statevector.expectation_value, states/statevector.py:377-408) -- SURVEY 8(f) rank 4.  `ws` as for
 * qsv_prob_qubit.  Synchronises the stream. */
int qsv_expval_pauli(const double* sv, int n_qubits, uint64_t z_mask, uint64_t x_mask, double phase_re,
                     double phase_im, void* ws, double* host_out, void* stream);

/* psi[abs(psi) < threshold] = 0 (qasm_simulator.py:330-334, unitary_simulator.py:201-207). */
int qsv_chop(double* sv, int n_qubits, double threshold, void* stream);

/* Device -> host copy of 2^n amplitudes (the ndarray returned in data["statevector"]). */
int qsv_download(const double* sv, int n_qubits, double* host_amps, void* stream);

/* ---- multi-GPU support (sv sharded by the high-order "global" qubits) -------------- */

/* Pack / unpack the half of the local shard whose local bit `local_qubit` equals `bit_value` into a
 * contiguous buffer of 2^(n_local-1) amplitudes, chunk [chunk_begin, chunk_begin+chunk_amps).  Used
 * around the NCCL exchange that swaps a global qubit with a local one. */
int qsv_pack_half(const double* sv, int n_local, int local_qubit, int bit_value, double* dev_buf,
                  size_t chunk_begin, size_t chunk_amps, void* stream);
int qsv_unpack_half(double* sv, int n_local, int local_qubit, int bit_value, const double* dev_buf,
                    size_t chunk_begin, size_t chunk_amps, void* stream);

/* General form for swapping k global qubits with k local ones in a single all-to-all: the sub-block of
 * 2^(n_local-k) amplitudes whose local bits local_qubits[i] equal bit i of `pattern`. */
int qsv_pack_bits(const double* sv, int n_local, const int32_t* local_qubits, int k, uint32_t pattern,
                  double* dev_buf, size_t chunk_begin, size_t chunk_amps, void* stream);
int qsv_unpack_bits(double* sv, int n_local, const int32_t* local_qubits, int k, uint32_t pattern,
                    const double* dev_buf, size_t chunk_begin, size_t chunk_amps, void* stream);

/* Qubit swap over NVLink peer memory, in place and without staging: the k global qubits whose
 * values on this rank read `my_pattern` trade places with the local qubits local_qubits[0..k).  peer_sv[lam] is the
 * shard of the rank whose k global bits read lam, mapped into this process with qsv_ipc_import (entry my_pattern is
 * ignored).  ONE kernel per rank: for every peer it swaps, element by element, half of the pair's sub-block (the
 * partner does the other half) -- remote load / remote store over NVLink, so every moved amplitude crosses the
 * link once per direction and touches HBM once per side.  The caller must bracket the call with a stream-ordered
 * barrier among the participating ranks.  max_ctas > 0 limits the grid (to leave SMs to a concurrent pass). */
int qsv_swap_bits_p2p(double* sv, int n_local, const int32_t* local_qubits, int k, uint32_t my_pattern,
                      double* const* peer_sv, int max_ctas, void* stream);

/* CUDA IPC plumbing for the above (one process per GPU): export = 64-byte handle of the allocation that holds
 * dev_ptr + dev_ptr's offset inside it; import maps a peer's allocation and returns the peer pointer; release
 * unmaps it (before the owner frees the memory). */
int qsv_ipc_export(const void* dev_ptr, void* handle_out, uint64_t* offset_out);
int qsv_ipc_import(const void* handle, uint64_t offset, void** peer_ptr_out);
int qsv_ipc_release(void* peer_ptr, uint64_t offset);

/* dev_dst (2^log2_dst doubles) = 0 everywhere except dev_dst[base | sum_i bit_i(s) << dest_bit[i]] = dev_src[s]
 * for s in [0, 2^m): a shard's marginal over its local measured qubits lands in the bins of the full marginal that
 * its rank bits (base) select; the ranks' arrays are then all-reduced (qasm_simulator.py:202-209 across shards). */
int qsv_scatter_bits(const double* dev_src, int m, const int32_t* dest_bit, uint64_t base, double* dev_dst,
                     int log2_dst, void* stream);

/* SWAP gates on `n_pairs` disjoint qubit pairs (a_i = pairs[2 i], b_i = pairs[2 i + 1]) as ONE in-place pass over the
 * state: amplitude i trades places with amplitude pi(i) = i with the index bits of every pair exchanged.  The reference
 * applies a swap as three cx gates = three einsum passes over the vector (qasm_simulator.py:145-161, 545-552; the final
 * swaps of a QFT, the routing swaps of a transpiled circuit); the host takes swaps out of the gate stream by relabelling
 * the later gates and settles the accumulated qubit permutation here.  Every amplitude is read and written once (32 * 2^n
 * bytes); pairs of qubits >= 4 keep 256-byte runs on both sides.  n_pairs <= QSV_MAX_SWAP_PAIRS. */
#define QSV_MAX_SWAP_PAIRS 20
int qsv_swap_qubit_pairs(double* sv, int n_qubits, const int32_t* pairs, int n_pairs, void* stream);

/* Sum of |psi|^2 over the local shard (for distributed normalisation / sampling). */
int qsv_norm2(const double* sv, int n_qubits, void* ws, double* host_out, void* stream);

/* Device memory for callers without an allocator of their own (an FFI consumer per INTEGRATION.md; the Python host
 * uses them for state snapshots): cudaMalloc / cudaFree / cudaMemcpyAsync device-to-device on `stream`.  The
 * snapshot / restore of a shot's starting state (qasm_simulator.py:512-519 re-initialises per shot) is one
 * qsv_copy_d2d. */
int qsv_alloc(void** dev_out, size_t bytes);
int qsv_free(void* dev_ptr);
int qsv_copy_d2d(void* dev_dst, const void* dev_src, size_t bytes, void* stream);

/* Number of kernels launched by this library in the calling process (bench.py's gpu_launches). */
uint64_t qsv_launch_count(void);

/* Select an equivalent variant of the pass kernel (tests and tools/; no reference counterpart): bit 8 = no
 * gate-pair fusion, 16 = single-buffer shape instead of the warp-group shape, 32 = no L2 prefetch, 64 = no
 * diagonal-run tables, 512 = warp groups without producer warps.  None of these changes results.  Bits 1, 2, 4
 * (no stores / no loads / no gates: timing experiments) are accepted by the debug library only. */
int qsv_debug_set_flags(uint32_t flags);

/* Profiling aid (tools/ only; no reference counterpart): per-phase clock totals of the last warp-group pass,
 * [CTA][group][role 0..8][20] uint64 -- roles 0..7 = the group's compute warps (0 wait for the tile, 1 gates, 2 loop
 * overhead, 4 stage barriers, 5 + o = kernel op o), role 8 = the group's TMA producer warp (0 wait for the gates,
 * 1 issue stores, 2 wait for the stores to read the buffer, 3 issue loads).  Only the debug library (libqsv_dbg.so, -DQSV_PROFILE) records
 * them; the production library returns QSV_ERANGE. */
int qsv_debug_profile(uint64_t* host_out, size_t n_words);

#ifdef __cplusplus
}
#endif
#endif /* QSV_H */
