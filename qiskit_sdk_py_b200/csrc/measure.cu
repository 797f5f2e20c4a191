// Probabilities, marginals and shot sampling (K7, K9, K10, K13 of SURVEY 2.5) for sm_100a.
//
// Replaces, in /root/reference/qiskit/providers/basicaer/qasm_simulator.py:
//   _get_measure_outcome  :163-182   np.sum(np.abs(psi)**2, axis=all but one)
//   _add_sample_measure   :184-225   marginal + RandomState.choice == cumsum / normalise / searchsorted
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace qsv {

constexpr int kReduceBlocks = 1184;           // fixed grid (8 per SM of a B200) -> deterministic two-stage reduction
constexpr int kChunkLog2 = 10;               // sampler chunk: 1024 bins
constexpr int kChunk = 1 << kChunkLog2;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of two values, result valid in thread 0; fixed order -> deterministic
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[kThreads / 32], sb[kThreads / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    sa[wid] = a;
    sb[wid] = b;
  }
  __syncthreads();
  if (wid == 0) {
    a = lane < kThreads / 32 ? sa[lane] : 0.0;
    b = lane < kThreads / 32 ? sb[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
  __syncthreads();
}

// stage 1: partial[2*blk + bit] = sum over this block's grid-stride slice of |psi|^2 with bit q == bit
__global__ void __launch_bounds__(kThreads)
prob_qubit_kernel(const double2* __restrict__ sv, u64 n_amps, int q, double* __restrict__ partial) {
  double p0 = 0.0, p1 = 0.0;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
    const double2 a = sv[i];
    const double p = fma(a.x, a.x, a.y * a.y);
    if ((i >> q) & 1ull) p1 += p; else p0 += p;
  }
  block_sum2(p0, p1);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = p0;
    partial[2 * blockIdx.x + 1] = p1;
  }
}

// stage 2: one block adds the partials in a fixed order
__global__ void __launch_bounds__(kThreads)
reduce_pairs_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  double p0 = 0.0, p1 = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += kThreads) {
    p0 += partial[2 * i];
    p1 += partial[2 * i + 1];
  }
  block_sum2(p0, p1);
  if (threadIdx.x == 0) {
    out[0] = p0;
    out[1] = p1;
  }
}

// ---- Pauli expectation value <psi| P |psi> for a Pauli string given by its z / x bit masks.  Restates
// /root/reference/qiskit/quantum_info/states/cython/exp_value.pyx: expval_pauli_no_x (:37-49) when
// x_mask == 0, expval_pauli_with_x (:52-94) otherwise; the serial loop becomes a grid-stride reduction
// with the same two-stage deterministic summation as prob_qubit. ----
__global__ void __launch_bounds__(kThreads)
expval_pauli_kernel(const double2* __restrict__ sv, u64 n_amps, u64 z_mask, u64 x_mask, int x_max, double2 phase,
                    double* __restrict__ partial) {
  double acc = 0.0, dummy = 0.0;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  if (x_mask == 0) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
      const double2 a = sv[i];
      const double p = a.x * a.x + a.y * a.y;
      acc += (__popcll(i & z_mask) & 1) ? -p : p;
    }
  } else {
    const u64 mask_u = ~((2ull << x_max) - 1ull), mask_l = (1ull << x_max) - 1ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < (n_amps >> 1); i += stride) {
      const u64 i0 = ((i << 1) & mask_u) | (i & mask_l);
      const u64 i1 = i0 ^ x_mask;
      const double2 d0 = sv[i0], d1 = sv[i1];
      // Re(phase * d1 * conj(d0)) and Re(phase * d0 * conj(d1))
      const double re = d1.x * d0.x + d1.y * d0.y, im = d1.y * d0.x - d1.x * d0.y;
      const double v0 = phase.x * re - phase.y * im;
      const double v1 = phase.x * re + phase.y * im;
      acc += (__popcll(i0 & z_mask) & 1) ? -v0 : v0;
      acc += (__popcll(i1 & z_mask) & 1) ? -v1 : v1;
    }
  }
  block_sum2(acc, dummy);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = acc;
    partial[2 * blockIdx.x + 1] = 0.0;
  }
}

// ---- shot-batched mid-circuit operations: the top `b` qubits of the vector number independent shots ----------
// The reference re-runs a circuit with mid-circuit measure / reset / conditionals once per shot
// (qasm_simulator.py:512-531, :561-614).  Here 2^b shots are simulated side by side as one (n + b)-qubit vector:
// gates on the n circuit qubits act on every shot at once through the ordinary pass kernel; the per-shot parts are
// these two kernels.
__global__ void __launch_bounds__(kThreads)
collapse_batch_kernel(double2* __restrict__ sv, u64 n_pairs, int q, int shot_shift, const unsigned char* __restrict__ outcome,
                      const double* __restrict__ scale, int is_reset) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 low_mask = (1ull << q) - 1ull;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n_pairs; j += stride) {
    const u64 i0 = ((j >> q) << (q + 1)) | (j & low_mask);  // bit q clear
    const u64 i1 = i0 | (1ull << q);
    const u64 shot = i0 >> shot_shift;
    const int out = outcome[shot];
    const double f = scale[shot];
    const double2 a0 = sv[i0], a1 = sv[i1];
    double2 r0, r1 = make_double2(0.0, 0.0);
    if (out == 0) {
      r0 = make_double2(a0.x * f, a0.y * f);                      // diag(1/sqrt(p), 0)   (:249, :269)
    } else if (is_reset) {
      r0 = make_double2(a1.x * f, a1.y * f);                      // [[0, 1/sqrt(p)], [0, 0]]   (:272)
    } else {
      r0 = make_double2(0.0, 0.0);
      r1 = make_double2(a1.x * f, a1.y * f);                      // diag(0, 1/sqrt(p))   (:251)
    }
    sv[i0] = r0;
    sv[i1] = r1;
  }
}

// a 1- or 2-qubit gate (dense 2x2 / 4x4 complex, row-major) on the shots whose mask byte is set (conditional gates)
struct MaskedGate {
  int32_t k, q0, q1, shot_shift;
  double m[32];
};

__global__ void __launch_bounds__(kThreads)
apply_masked_kernel(double2* __restrict__ sv, u64 n_groups, const __grid_constant__ MaskedGate g,
                    const unsigned char* __restrict__ mask) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const int lo = g.k == 2 ? (g.q0 < g.q1 ? g.q0 : g.q1) : g.q0, hi = g.k == 2 ? (g.q0 < g.q1 ? g.q1 : g.q0) : -1;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n_groups; j += stride) {
    u64 base = ((j >> lo) << (lo + 1)) | (j & ((1ull << lo) - 1ull));
    if (hi >= 0) base = ((base >> hi) << (hi + 1)) | (base & ((1ull << hi) - 1ull));
    if (!mask[base >> g.shot_shift]) continue;
    const int dim = 1 << g.k;
    double2 a[4], r[4];
    for (int c = 0; c < dim; c++) {
      u64 idx = base;
      if (c & 1) idx |= 1ull << g.q0;
      if (c & 2) idx |= 1ull << g.q1;
      a[c] = sv[idx];
    }
    for (int rr = 0; rr < dim; rr++) {
      double2 acc = make_double2(0.0, 0.0);
      for (int c = 0; c < dim; c++) cfma(acc, make_double2(g.m[2 * (dim * rr + c)], g.m[2 * (dim * rr + c) + 1]), a[c]);
      r[rr] = acc;
    }
    for (int c = 0; c < dim; c++) {
      u64 idx = base;
      if (c & 1) idx |= 1ull << g.q0;
      if (c & 2) idx |= 1ull << g.q1;
      sv[idx] = r[c];
    }
  }
}

// ---- marginal over m measured qubits: grid = (parts, bins) ----
struct MargDesc {
  int32_t n, m, log2_parts, pad;
  uint8_t mq[40];   // measured qubits ascending
  uint8_t uq[40];   // unmeasured qubits ascending
};

__global__ void __launch_bounds__(kThreads)
marginal_kernel(const double2* __restrict__ sv, const __grid_constant__ MargDesc d,
                double* __restrict__ partial) {
  const u64 bin = blockIdx.y + (u64)blockIdx.z * gridDim.y;
  const u64 part = blockIdx.x;
  const int nu = d.n - d.m;
  u64 base = 0;
  for (int j = 0; j < d.m; j++) base |= ((bin >> j) & 1ull) << d.mq[j];
  const u64 per_part = (1ull << nu) >> d.log2_parts;
  double acc = 0.0, dummy = 0.0;
  for (u64 u = part * per_part + threadIdx.x; u < (part + 1) * per_part; u += kThreads) {
    u64 off = 0;
    for (int j = 0; j < nu; j++) off |= ((u >> j) & 1ull) << d.uq[j];
    const double2 a = sv[base | off];
    acc += fma(a.x, a.x, a.y * a.y);
  }
  block_sum2(acc, dummy);
  if (threadIdx.x == 0) partial[(bin << d.log2_parts) + part] = acc;
}

// Few unmeasured qubits (n - m <= 10): one thread per bin adds its 2^(n-m) amplitudes in a fixed order.
__global__ void __launch_bounds__(kThreads)
marginal_small_kernel(const double2* __restrict__ sv, const __grid_constant__ MargDesc d, u64 bins,
                      double* __restrict__ probs) {
  const u64 bin = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= bins) return;
  const int nu = d.n - d.m;
  u64 base = 0;
  for (int j = 0; j < d.m; j++) base |= ((bin >> j) & 1ull) << d.mq[j];
  double acc = 0.0;
  for (u64 u = 0; u < (1ull << nu); u++) {
    u64 off = 0;
    for (int j = 0; j < nu; j++) off |= ((u >> j) & 1ull) << d.uq[j];
    const double2 a = sv[base | off];
    acc += fma(a.x, a.x, a.y * a.y);
  }
  probs[bin] = acc;
}

__global__ void __launch_bounds__(kThreads)
marginal_finish_kernel(const double* __restrict__ partial, u64 bins, int log2_parts, double* __restrict__ probs) {
  const u64 bin = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= bins) return;
  double acc = 0.0;
  const u64 parts = 1ull << log2_parts;
  for (u64 p = 0; p < parts; p++) acc += partial[(bin << log2_parts) + p];
  probs[bin] = acc;
}

// ---- sampler ----
// |z|^2 exactly as the reference computes it: np.abs(psi) ** 2 (qasm_simulator.py:207).  numpy's complex128
// absolute (SIMD loop of loops_unary_complex.dispatch.c; AVX2/AVX-512 builds fuse the multiply-add) is
//     a = max * sqrt(fma(r, r, 1)),  r = min / max,  max/min of (|re|, |im|),  0 if max == 0
// -- not hypot() and not sqrt(re^2 + im^2); each differs from it in the last bit for about a third of random
// inputs -- and ** 2 is a * a.  Division, fma, sqrt and product are correctly rounded on both sides, so the
// probabilities agree bit for bit (tests/test_gpu_kernels.py::test_exact_cdf_from_amplitudes_is_numpy_abs2).
__device__ __forceinline__ double numpy_abs2(double2 a) {
  const double ax = fabs(a.x), ay = fabs(a.y);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  if (mx == 0.0) return 0.0;
  const double r = __ddiv_rn(mn, mx);
  const double h = __dmul_rn(__dsqrt_rn(__fma_rn(r, r, 1.0)), mx);
  return __dmul_rn(h, h);
}

// kExact: the sampler's exact-order path (bitwise the reference's probabilities); otherwise re^2 + im^2 with one
// rounding less, for sums that are compared with a tolerance or only steer the exact path.
template <bool kAmps, bool kExact = true>
__device__ __forceinline__ double load_p(const void* src, u64 i) {
  if (kAmps) {
    const double2 a = reinterpret_cast<const double2*>(src)[i];
    return kExact ? numpy_abs2(a) : fma(a.x, a.x, a.y * a.y);
  }
  return reinterpret_cast<const double*>(src)[i];
}

// Exact (numpy cumsum) order: ONE running float64 sum over all bins, left to right.  Warp 0 lane 0 is
// the adder; the other warps stage the next chunk's probabilities in shared memory meanwhile.  Records
// the running sum at the end of every chunk.
template <bool kAmps>
__global__ void __launch_bounds__(1024)
cdf_sequential_kernel(const void* __restrict__ src, u64 bins, double start, double* __restrict__ chunk_end) {
  __shared__ double buf[2][kChunk];
  const u64 nchunks = (bins + kChunk - 1) / kChunk;
  const int tid = threadIdx.x;
  double run = start;
  // prologue: stage chunk 0
  for (int i = tid; i < kChunk; i += 1024) buf[0][i] = (u64)i < bins ? load_p<kAmps>(src, i) : 0.0;
  __syncthreads();
  for (u64 c = 0; c < nchunks; c++) {
    const int cur = c & 1;
    if (tid >= 32) {
      const u64 nb = (c + 1) * kChunk;
      if (c + 1 < nchunks)
        for (int i = tid - 32; i < kChunk; i += 1024 - 32) buf[cur ^ 1][i] = nb + i < bins ? load_p<kAmps>(src, nb + i) : 0.0;
    } else if (tid == 0) {
      const u64 left = bins - c * kChunk;
      const int cnt = left < (u64)kChunk ? (int)left : kChunk;
      const double* p = buf[cur];
#pragma unroll 8
      for (int i = 0; i < cnt; i++) run += p[i];
      chunk_end[c] = run;
    }
    __syncthreads();
  }
}

// ---- exact (numpy cumsum) order, in parallel ---------------------------------------------------------
// numpy's cumsum is one running float64 sum, s <- fl(s + p_i).  While s stays inside one binade
// [2^k, 2^(k+1)) its ulp u = 2^(k-52) is fixed and every partial sum is a multiple of u, so
//     fl(s + p) = s + u * rint(p / u)      (unless p/u is exactly half-way: ties depend on s)
// i.e. the sequential sum is an INTEGER sum of q_i = rint(p_i / u), which is associative and can be
// reduced in parallel without changing a single bit.  Per 1024-bin chunk:
//   ZERO  all p_i == 0                                   -> running sum unchanged
//   INT   the (approximate) running sum provably stays in one binade over the chunk and no p_i / u is a
//         tie                                             -> add the chunk's integer increment
//   SLOW  anything else (binade crossings: ~60 chunks)    -> sequential float64 adds, as numpy does
// chunk_classify_kernel computes mode / binade / integer increment of every chunk in parallel;
// chunk_chain_kernel then walks the chunks once (a few dozen cycles per chunk).
constexpr int kModeZero = 0, kModeInt = 1, kModeSlow = 2;

template <bool kAmps>
__global__ void __launch_bounds__(kThreads)
chunk_classify_kernel(const void* __restrict__ src, u64 bins, const double* __restrict__ aprefix, double offset,
                      int* __restrict__ meta, long long* __restrict__ inc) {
  __shared__ long long s_sum[kThreads / 32];
  __shared__ int s_flag[kThreads / 32];
  const u64 c = blockIdx.x;
  // `offset`: (approximate) running sum before this array -- the shards before this one of a sharded state
  const double a_prev = offset + (c > 0 ? aprefix[c - 1] : 0.0), a_cur = offset + aprefix[c];
  const double lo = a_prev * (1.0 - 9.313225746154785e-10), hi = a_cur * (1.0 + 9.313225746154785e-10);  // 2^-30 margins
  int k = 0;
  bool int_ok = lo > 0.0;
  if (int_ok) {
    k = ilogb(lo);
    int_ok = hi < ldexp(1.0, k + 1) && k > -900;
  }
  long long acc = 0;
  int flag = 0;  // bit 0: some p != 0, bit 1: a tie
  for (int i = threadIdx.x; i < kChunk; i += kThreads) {
    const u64 idx = c * kChunk + i;
    if (idx >= bins) break;
    const double pv = load_p<kAmps>(src, idx);
    if (pv != 0.0) flag |= 1;
    if (int_ok) {
      const double x = ldexp(pv, 52 - k);
      if (x - floor(x) == 0.5) flag |= 2;
      acc += (long long)rint(x);
    }
  }
  // block reduction of (acc, flag)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    flag |= __shfl_xor_sync(0xffffffffu, flag, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s_sum[wid] = acc;
    s_flag[wid] = flag;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    int fl = 0;
    for (int w = 0; w < kThreads / 32; w++) {
      tot += s_sum[w];
      fl |= s_flag[w];
    }
    int mode = kModeSlow;
    if (!(fl & 1)) mode = kModeZero;
    else if (int_ok && !(fl & 2)) mode = kModeInt;
    meta[c] = mode | ((k + 2048) << 2);
    inc[c] = tot;
  }
}

// Groups of 32 chunks ("super-chunks") whose chunks are all INT in the same binade (or ZERO) are summarised
// in parallel -- total increment and per-chunk inclusive prefix -- so that the sequential chain below takes
// one integer add per 32768 bins; only mixed super-chunks (binade crossings) are walked chunk by chunk.
__global__ void __launch_bounds__(kThreads)
super_classify_kernel(u64 nchunks, const int* __restrict__ meta, const long long* __restrict__ inc,
                      int* __restrict__ smeta, long long* __restrict__ sinc, long long* __restrict__ pref) {
  const u64 sidx = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const u64 nsuper = (nchunks + 31) >> 5;
  if (sidx >= nsuper) return;
  const u64 c = sidx * 32 + lane;
  const int m = c < nchunks ? meta[c] : kModeZero;
  long long q = (c < nchunks && (m & 3) == kModeInt) ? inc[c] : 0;
  const int mode = m & 3, k = (m >> 2) - 2048;
  const unsigned int_mask = __ballot_sync(0xffffffffu, mode == kModeInt);
  const unsigned slow_mask = __ballot_sync(0xffffffffu, mode == kModeSlow);
  int k_first = 0;
  bool uniform = slow_mask == 0;
  if (int_mask) {
    k_first = __shfl_sync(0xffffffffu, k, __ffs(int_mask) - 1);
    uniform = uniform && __all_sync(0xffffffffu, mode != kModeInt || k == k_first);
  }
  // inclusive warp scan of the increments
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, q, o);
    if (lane >= o) q += t;
  }
  if (c < nchunks) pref[c] = q;
  if (lane == 31) {
    int sm = kModeSlow;  // mixed
    if (uniform) sm = int_mask ? kModeInt : kModeZero;
    smeta[sidx] = sm | ((k_first + 2048) << 2);
    sinc[sidx] = q;
  }
}

template <bool kAmps>
__global__ void __launch_bounds__(32)
chunk_chain_kernel(const void* __restrict__ src, u64 bins, u64 nchunks, const int* __restrict__ meta,
                   const long long* __restrict__ inc, const int* __restrict__ smeta,
                   const long long* __restrict__ sinc, double start, double* __restrict__ chunk_end,
                   long long* __restrict__ sstart, int* __restrict__ err) {
  const int lane = threadIdx.x;
  const u64 nsuper = (nchunks + 31) >> 5;
  double run = start;  // identical in every lane; `start` = the exact running sum where this array begins
  // inside a run of INT (super-)chunks of one binade the running sum is carried as an integer (units of
  // the ulp): the dependent chain is then a single integer add per step
  long long s_int = 0;
  int cur_k = 1 << 30;
  double dn = 0.0;
  auto enter_int = [&](int k) {
    if (k != cur_k) {
      // exact scaling by powers of two (k > -900, so both exponents are in range)
      const double up = __longlong_as_double((long long)(1023 + 52 - k) << 52);
      dn = __longlong_as_double((long long)(1023 - 52 + k) << 52);
      const double scaled = run * up;
      s_int = (long long)scaled;
      if (!(s_int >= (1ll << 52) && (double)s_int == scaled)) *err = 1;  // guarantee violated
      cur_k = k;
    }
  };
  for (u64 s0 = 0; s0 < nsuper; s0 += 32) {
    const u64 smine = s0 + lane;
    const int my_smeta = smine < nsuper ? smeta[smine] : 0;
    const long long my_sinc = smine < nsuper ? sinc[smine] : 0;
    long long my_start = 0;
    const int scnt = nsuper - s0 < 32 ? (int)(nsuper - s0) : 32;
    for (int sj = 0; sj < scnt; sj++) {
      const int sm = __shfl_sync(0xffffffffu, my_smeta, sj);
      const long long sq = __shfl_sync(0xffffffffu, my_sinc, sj);
      const int smode = sm & 3, sk = (sm >> 2) - 2048;
      if (smode == kModeInt) {
        enter_int(sk);
        if (lane == sj) my_start = s_int;  // running sum (in ulps of binade sk) before this super-chunk
        s_int += sq;
        if (s_int >= (1ll << 53)) *err = 1;
        run = (double)s_int * dn;
      } else if (smode == kModeZero) {
        if (lane == sj) my_start = __double_as_longlong(run);  // every chunk end of it equals `run`
      } else {
        // mixed super-chunk: walk its chunks
        const u64 c0 = (s0 + sj) * 32;
        const u64 mine = c0 + lane;
        const int my_meta = mine < nchunks ? meta[mine] : 0;
        const long long my_inc = mine < nchunks ? inc[mine] : 0;
        double my_end = 0.0;
        const int cnt = nchunks - c0 < 32 ? (int)(nchunks - c0) : 32;
        for (int j = 0; j < cnt; j++) {
          const int m = __shfl_sync(0xffffffffu, my_meta, j);
          const long long q = __shfl_sync(0xffffffffu, my_inc, j);
          const int mode = m & 3, k = (m >> 2) - 2048;
          if (mode == kModeInt) {
            enter_int(k);
            s_int += q;
            if (s_int >= (1ll << 53)) *err = 1;
            run = (double)s_int * dn;
          } else if (mode == kModeSlow) {
            const u64 begin = (c0 + j) * kChunk;
            double vals[kChunk / 32];
#pragma unroll
            for (int t = 0; t < kChunk / 32; t++) {
              const u64 idx = begin + (u64)t * 32 + lane;
              vals[t] = idx < bins ? load_p<kAmps>(src, idx) : 0.0;
            }
#pragma unroll
            for (int t = 0; t < kChunk / 32; t++)
              for (int l = 0; l < 32; l++) run += __shfl_sync(0xffffffffu, vals[t], l);
            cur_k = 1 << 30;  // leave integer mode: the next INT chunk re-derives it from `run`
          }
          if (lane == j) my_end = run;
        }
        if (mine < nchunks) chunk_end[mine] = my_end;
      }
    }
    if (smine < nsuper) sstart[smine] = my_start;
  }
}

// chunk ends of the uniform super-chunks, in parallel
__global__ void __launch_bounds__(kThreads)
chunk_fill_kernel(u64 nchunks, const int* __restrict__ smeta, const long long* __restrict__ sstart,
                  const long long* __restrict__ pref, double* __restrict__ chunk_end) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const int sm = smeta[c >> 5];
  const int smode = sm & 3, k = (sm >> 2) - 2048;
  if (smode == kModeInt) {
    const double dn = __longlong_as_double((long long)(1023 - 52 + k) << 52);
    chunk_end[c] = (double)(sstart[c >> 5] + pref[c]) * dn;
  } else if (smode == kModeZero) {
    chunk_end[c] = __longlong_as_double(sstart[c >> 5]);
  }
}

// Blocked order: chunk sums (sequential inside a chunk, one thread per chunk slice is too slow, so a
// fixed-shape tree per chunk), then an inclusive scan over the chunk sums.
template <bool kAmps>
__global__ void __launch_bounds__(kThreads)
chunk_sum_kernel(const void* __restrict__ src, u64 bins, double* __restrict__ chunk_sum) {
  const u64 c = blockIdx.x;
  double acc = 0.0, dummy = 0.0;
  for (int i = threadIdx.x; i < kChunk; i += kThreads) {
    const u64 idx = c * kChunk + i;
    if (idx < bins) acc += load_p<kAmps, false>(src, idx);
  }
  block_sum2(acc, dummy);
  if (threadIdx.x == 0) chunk_sum[c] = acc;
}

// In-place inclusive scan of `n` doubles, hierarchical with one thread per 1024-element segment doing a
// sequential scan (n <= 2^23 chunk sums at 33 qubits -> 2^13 segments -> one more level).
__global__ void seg_scan_kernel(double* __restrict__ data, u64 n, double* __restrict__ seg_total) {
  const u64 seg = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 begin = seg * 1024;
  if (begin >= n) return;
  const u64 end = begin + 1024 < n ? begin + 1024 : n;
  double run = 0.0;
  for (u64 i = begin; i < end; i++) {
    run += data[i];
    data[i] = run;
  }
  seg_total[seg] = run;
}

__global__ void serial_scan_kernel(double* __restrict__ data, u64 n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double run = 0.0;
    for (u64 i = 0; i < n; i++) {
      run += data[i];
      data[i] = run;
    }
  }
}

__global__ void seg_add_kernel(double* __restrict__ data, u64 n, const double* __restrict__ seg_incl) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 seg = i >> 10;
  if (seg > 0) data[i] += seg_incl[seg - 1];
}

// One thread per shot: first bin with cdf[bin] / total > u  (searchsorted side='right').  `start` / `total`: the
// running sum where this array begins and the grand total (0 and chunk_end[last] for a whole state; a shard of a
// sharded state begins where the previous shard ended).  Shots that fall before this array get -1, after it -2.
template <bool kAmps, bool kExact>
__global__ void __launch_bounds__(kThreads)
sample_kernel(const void* __restrict__ src, u64 bins, const double* __restrict__ chunk_end, u64 nchunks, double start,
              double total, const double* __restrict__ uniforms, int shots, long long* __restrict__ out) {
  const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= shots) return;
  const double u = uniforms[sidx];
  if (start / total > u) {
    out[sidx] = -1;
    return;
  }
  // first chunk whose end value exceeds u
  u64 lo = 0, hi = nchunks;  // answer in [lo, hi]
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if (chunk_end[mid] / total > u) hi = mid; else lo = mid + 1;
  }
  if (lo >= nchunks) {  // a later shard (for a whole state: u >= 1, which random_sample() never returns)
    out[sidx] = -2;
    return;
  }
  double run = lo > 0 ? chunk_end[lo - 1] : start;
  const u64 begin = lo * kChunk;
  const u64 end = begin + kChunk < bins ? begin + kChunk : bins;
  // exact order: the walk repeats the adds that produced chunk_end[lo], so it always finds the bin.  Blocked
  // order: chunk_end comes from a tree sum, the walk may end a rounding error short -- then the last bin of the
  // chunk with p > 0 is the answer (never a bin the reference cannot emit)
  u64 ans = end - 1, last_nonzero = end - 1;
  bool found = false;
  for (u64 i = begin; i < end; i++) {
    const double pv = load_p<kAmps, kExact>(src, i);
    if (pv > 0.0) last_nonzero = i;
    run += pv;
    if (run / total > u) {
      ans = i;
      found = true;
      break;
    }
  }
  out[sidx] = (long long)(found ? ans : last_nonzero);
}

struct SampleWs {
  double* chunk_end;   // nchunks
  double* aprefix;     // nchunks (exact mode: approximate inclusive prefix of the chunk sums)
  long long* inc;      // nchunks (exact mode: integer increment of each chunk)
  int* meta;           // nchunks + 1 (exact mode: mode | binade; last entry = error flag)
  long long* pref;     // nchunks (inclusive prefix of inc inside each super-chunk of 32 chunks)
  int* smeta;          // nsuper
  long long* sinc;     // nsuper
  long long* sstart;   // nsuper
  double* seg1;        // ceil(nchunks/1024)
  double* uniforms;    // shots
  long long* out;      // shots
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t sample_ws_layout(int log2_bins, int shots, void* ws, SampleWs* out) {
  const u64 bins = 1ull << log2_bins;
  const u64 nchunks = (bins + kChunk - 1) / kChunk;
  const u64 nseg = (nchunks + 1023) / 1024;
  size_t off = 0;
  char* base = reinterpret_cast<char*>(ws);
  if (out) out->chunk_end = reinterpret_cast<double*>(base + off);
  off += align_up(nchunks * sizeof(double));
  if (out) out->aprefix = reinterpret_cast<double*>(base + off);
  off += align_up(nchunks * sizeof(double));
  if (out) out->inc = reinterpret_cast<long long*>(base + off);
  off += align_up(nchunks * sizeof(long long));
  if (out) out->meta = reinterpret_cast<int*>(base + off);
  off += align_up((nchunks + 1) * sizeof(int));
  const u64 nsuper = (nchunks + 31) / 32;
  if (out) out->pref = reinterpret_cast<long long*>(base + off);
  off += align_up(nchunks * sizeof(long long));
  if (out) out->smeta = reinterpret_cast<int*>(base + off);
  off += align_up(nsuper * sizeof(int));
  if (out) out->sinc = reinterpret_cast<long long*>(base + off);
  off += align_up(nsuper * sizeof(long long));
  if (out) out->sstart = reinterpret_cast<long long*>(base + off);
  off += align_up(nsuper * sizeof(long long));
  if (out) out->seg1 = reinterpret_cast<double*>(base + off);
  off += align_up(nseg * sizeof(double));
  if (out) out->uniforms = reinterpret_cast<double*>(base + off);
  off += align_up((size_t)shots * sizeof(double));
  if (out) out->out = reinterpret_cast<long long*>(base + off);
  off += align_up((size_t)shots * sizeof(long long));
  return off;
}

// ---- the phases of the exact-order sampler (each usable on a shard of a sharded state) ----
// (1) chunk sums + inclusive scan -> ws.aprefix (approximate prefix, steers the classification); total to the host
template <bool kAmps>
static int cdf_prepare(const void* src, int log2_bins, void* ws, double* host_total, cudaStream_t s) {
  const u64 bins = 1ull << log2_bins;
  const u64 nchunks = (bins + kChunk - 1) / kChunk;
  SampleWs w;
  sample_ws_layout(log2_bins, 0, ws, &w);
  const u64 nseg = (nchunks + 1023) / 1024;
  chunk_sum_kernel<kAmps><<<(unsigned)nchunks, kThreads, 0, s>>>(src, bins, w.aprefix);
  QSV_LAUNCHED();
  seg_scan_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, s>>>(w.aprefix, nchunks, w.seg1);
  QSV_LAUNCHED();
  if (nseg > 1) {
    serial_scan_kernel<<<1, 32, 0, s>>>(w.seg1, nseg);
    QSV_LAUNCHED();
    seg_add_kernel<<<(unsigned)((nchunks + 255) / 256), 256, 0, s>>>(w.aprefix, nchunks, w.seg1);
    QSV_LAUNCHED();
  }
  if (host_total) {
    QSV_CUDA(cudaMemcpyAsync(host_total, w.aprefix + (nchunks - 1), sizeof(double), cudaMemcpyDeviceToHost, s));
    QSV_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
}

// (2) mode / binade / integer increment of every chunk, given the approximate running sum before the array
template <bool kAmps>
static int cdf_classify(const void* src, int log2_bins, double approx_offset, void* ws, cudaStream_t s) {
  const u64 bins = 1ull << log2_bins;
  const u64 nchunks = (bins + kChunk - 1) / kChunk;
  SampleWs w;
  sample_ws_layout(log2_bins, 0, ws, &w);
  QSV_CUDA(cudaMemsetAsync(w.meta + nchunks, 0, sizeof(int), s));
  chunk_classify_kernel<kAmps><<<(unsigned)nchunks, kThreads, 0, s>>>(src, bins, w.aprefix, approx_offset, w.meta, w.inc);
  QSV_LAUNCHED();
  const u64 nsuper = (nchunks + 31) / 32;
  super_classify_kernel<<<(unsigned)((nsuper * 32 + kThreads - 1) / kThreads), kThreads, 0, s>>>(
      nchunks, w.meta, w.inc, w.smeta, w.sinc, w.pref);
  QSV_LAUNCHED();
  return 0;
}

// (3) the sequential chain from the exact running sum `start`; chunk ends -> ws.chunk_end, last one to the host
template <bool kAmps>
static int cdf_chain(const void* src, int log2_bins, double start, void* ws, double* host_end, cudaStream_t s) {
  const u64 bins = 1ull << log2_bins;
  const u64 nchunks = (bins + kChunk - 1) / kChunk;
  SampleWs w;
  sample_ws_layout(log2_bins, 0, ws, &w);
  chunk_chain_kernel<kAmps><<<1, 32, 0, s>>>(src, bins, nchunks, w.meta, w.inc, w.smeta, w.sinc, start, w.chunk_end,
                                             w.sstart, w.meta + nchunks);
  QSV_LAUNCHED();
  chunk_fill_kernel<<<(unsigned)((nchunks + kThreads - 1) / kThreads), kThreads, 0, s>>>(nchunks, w.smeta, w.sstart,
                                                                                         w.pref, w.chunk_end);
  QSV_LAUNCHED();
  int err = 0;
  QSV_CUDA(cudaMemcpyAsync(&err, w.meta + nchunks, sizeof(int), cudaMemcpyDeviceToHost, s));
  QSV_CUDA(cudaStreamSynchronize(s));
  if (err) {  // cannot happen by construction; keep the plain sequential kernel as the safety net
    cdf_sequential_kernel<kAmps><<<1, 1024, 0, s>>>(src, bins, start, w.chunk_end);
    QSV_LAUNCHED();
  }
  if (host_end) {
    QSV_CUDA(cudaMemcpyAsync(host_end, w.chunk_end + (nchunks - 1), sizeof(double), cudaMemcpyDeviceToHost, s));
    QSV_CUDA(cudaStreamSynchronize(s));
  }
  return 0;
}

// (4) searchsorted(cdf / total, u, 'right') restricted to this array: -1 = before it, -2 = after it
template <bool kAmps, bool kExact>
static int cdf_search(const void* src, int log2_bins, double start, double total, const double* host_uniforms, int shots,
                      int64_t* host_out, void* ws, cudaStream_t s) {
  const u64 bins = 1ull << log2_bins;
  const u64 nchunks = (bins + kChunk - 1) / kChunk;
  SampleWs w;
  sample_ws_layout(log2_bins, shots, ws, &w);
  QSV_CUDA(cudaMemcpyAsync(w.uniforms, host_uniforms, (size_t)shots * sizeof(double), cudaMemcpyHostToDevice, s));
  sample_kernel<kAmps, kExact><<<(shots + kThreads - 1) / kThreads, kThreads, 0, s>>>(src, bins, w.chunk_end, nchunks, start,
                                                                                      total, w.uniforms, shots, w.out);
  QSV_LAUNCHED();
  QSV_CUDA(cudaMemcpyAsync(host_out, w.out, (size_t)shots * sizeof(long long), cudaMemcpyDeviceToHost, s));
  QSV_CUDA(cudaStreamSynchronize(s));
  return 0;
}

template <bool kAmps>
static int sample_impl(const void* src, int log2_bins, const double* host_uniforms, int shots,
                       int64_t* host_out, int flags, void* ws, double* host_total, cudaStream_t s) {
  const bool exact = (flags & 1) != 0, check_norm = (flags & 2) == 0;
  const u64 bins = 1ull << log2_bins;
  const u64 nchunks = (bins + kChunk - 1) / kChunk;
  SampleWs w;
  sample_ws_layout(log2_bins, shots, ws, &w);
  double total = 0.0;
  if (exact) {
    if (int rc = cdf_prepare<kAmps>(src, log2_bins, ws, nullptr, s)) return rc;
    if (int rc = cdf_classify<kAmps>(src, log2_bins, 0.0, ws, s)) return rc;
    if (int rc = cdf_chain<kAmps>(src, log2_bins, 0.0, ws, &total, s)) return rc;
  } else {
    chunk_sum_kernel<kAmps><<<(unsigned)nchunks, kThreads, 0, s>>>(src, bins, w.chunk_end);
    QSV_LAUNCHED();
    const u64 nseg = (nchunks + 1023) / 1024;
    seg_scan_kernel<<<(unsigned)((nseg + 127) / 128), 128, 0, s>>>(w.chunk_end, nchunks, w.seg1);
    QSV_LAUNCHED();
    if (nseg > 1) {
      serial_scan_kernel<<<1, 32, 0, s>>>(w.seg1, nseg);
      QSV_LAUNCHED();
      seg_add_kernel<<<(unsigned)((nchunks + 255) / 256), 256, 0, s>>>(w.chunk_end, nchunks, w.seg1);
      QSV_LAUNCHED();
    }
    QSV_CUDA(cudaMemcpyAsync(&total, w.chunk_end + (nchunks - 1), sizeof(double), cudaMemcpyDeviceToHost, s));
    QSV_CUDA(cudaStreamSynchronize(s));
  }
  if (host_total) *host_total = total;
  // numpy's choice(): "probabilities do not sum to 1" when |sum - 1| > sqrt(eps)
  if (check_norm && !(fabs(total - 1.0) <= 1.4901161193847656e-08))
    return fail(QSV_ENORM, "probabilities do not sum to 1");
  if (!(total > 0.0)) return fail(QSV_ENORM, "probabilities sum to zero");
  const int rc = exact ? cdf_search<kAmps, true>(src, log2_bins, 0.0, total, host_uniforms, shots, host_out, ws, s)
                       : cdf_search<kAmps, false>(src, log2_bins, 0.0, total, host_uniforms, shots, host_out, ws, s);
  if (rc) return rc;
  for (int i = 0; i < shots; i++)
    if (host_out[i] < 0) host_out[i] = (int64_t)bins;  // u >= 1: searchsorted returns len(cdf)
  return 0;
}

}  // namespace qsv

using namespace qsv;

extern "C" {

size_t qsv_reduce_workspace_bytes(void) { return sizeof(double) * 2 * (kReduceBlocks + 1); }

int qsv_prob_qubit(const double* sv, int n, int qubit, void* ws, double* host_out, void* stream) {
  if (!sv || !ws || !host_out || n < 1 || n > 40 || qubit < 0 || qubit >= n)
    return fail(QSV_EINVAL, "qsv_prob_qubit: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const u64 n_amps = 1ull << n;
  double* partial = reinterpret_cast<double*>(ws);
  const int blocks = (int)std::max<u64>(1, std::min<u64>((n_amps + kThreads - 1) / kThreads, kReduceBlocks));
  prob_qubit_kernel<<<blocks, kThreads, 0, s>>>(reinterpret_cast<const double2*>(sv), n_amps, qubit, partial);
  QSV_LAUNCHED();
  reduce_pairs_kernel<<<1, kThreads, 0, s>>>(partial, blocks, partial + 2 * kReduceBlocks);
  QSV_LAUNCHED();
  QSV_CUDA(cudaMemcpyAsync(host_out, partial + 2 * kReduceBlocks, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  QSV_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int qsv_norm2(const double* sv, int n, void* ws, double* host_out, void* stream) {
  if (!sv || !ws || !host_out || n < 0 || n > 40) return fail(QSV_EINVAL, "qsv_norm2: bad arguments");
  double p[2];
  if (n == 0) {
    double a[2];
    QSV_CUDA(cudaMemcpyAsync(a, sv, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    QSV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *host_out = a[0] * a[0] + a[1] * a[1];
    return 0;
  }
  const int rc = qsv_prob_qubit(sv, n, 0, ws, p, stream);
  if (rc) return rc;
  *host_out = p[0] + p[1];
  return 0;
}

int qsv_expval_pauli(const double* sv, int n, uint64_t z_mask, uint64_t x_mask, double phase_re, double phase_im,
                     void* ws, double* host_out, void* stream) {
  if (!sv || !ws || !host_out || n < 1 || n > 40) return fail(QSV_EINVAL, "qsv_expval_pauli: bad arguments");
  const u64 n_amps = 1ull << n;
  if ((z_mask | x_mask) >= n_amps) return fail(QSV_EINVAL, "qsv_expval_pauli: mask has bits beyond n_qubits");
  cudaStream_t s = (cudaStream_t)stream;
  int x_max = 0;
  for (int j = 0; j < n; j++)
    if ((x_mask >> j) & 1ull) x_max = j;
  double* partial = reinterpret_cast<double*>(ws);
  const u64 work = x_mask ? (n_amps >> 1) : n_amps;
  const int blocks = (int)std::max<u64>(1, std::min<u64>((work + kThreads - 1) / kThreads, kReduceBlocks));
  expval_pauli_kernel<<<blocks, kThreads, 0, s>>>(reinterpret_cast<const double2*>(sv), n_amps, z_mask, x_mask, x_max,
                                                  make_double2(phase_re, phase_im), partial);
  QSV_LAUNCHED();
  reduce_pairs_kernel<<<1, kThreads, 0, s>>>(partial, blocks, partial + 2 * kReduceBlocks);
  QSV_LAUNCHED();
  double out[2];
  QSV_CUDA(cudaMemcpyAsync(out, partial + 2 * kReduceBlocks, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  QSV_CUDA(cudaStreamSynchronize(s));
  *host_out = out[0];
  return 0;
}

int qsv_collapse_batch(double* sv, int n_total, int n_shot_qubits, int qubit, const unsigned char* host_outcomes,
                       const double* host_scales, int is_reset, void* dev_ws, void* stream) {
  const int n = n_total - n_shot_qubits;
  if (!sv || !host_outcomes || !host_scales || !dev_ws || n_shot_qubits < 0 || n < 1 || n_total > 40 || qubit < 0 ||
      qubit >= n)
    return fail(QSV_EINVAL, "qsv_collapse_batch: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t shots = (size_t)1 << n_shot_qubits;
  double* d_scale = reinterpret_cast<double*>(dev_ws);
  unsigned char* d_out = reinterpret_cast<unsigned char*>(d_scale + shots);
  QSV_CUDA(cudaMemcpyAsync(d_scale, host_scales, shots * sizeof(double), cudaMemcpyHostToDevice, s));
  QSV_CUDA(cudaMemcpyAsync(d_out, host_outcomes, shots, cudaMemcpyHostToDevice, s));
  const u64 n_pairs = 1ull << (n_total - 1);
  const int blocks = (int)std::max<u64>(1, std::min<u64>((n_pairs + kThreads - 1) / kThreads, (u64)sm_count() * 16));
  collapse_batch_kernel<<<blocks, kThreads, 0, s>>>(reinterpret_cast<double2*>(sv), n_pairs, qubit, n, d_out, d_scale,
                                                    is_reset);
  QSV_LAUNCHED();
  QSV_CUDA(cudaStreamSynchronize(s));  // the host arrays may be pageable temporaries
  return 0;
}

size_t qsv_batch_workspace_bytes(int n_shot_qubits) {
  if (n_shot_qubits < 0 || n_shot_qubits > 30) return 0;
  return ((size_t)1 << n_shot_qubits) * (sizeof(double) + 1) + 256;
}

int qsv_apply_masked(double* sv, int n_total, int n_shot_qubits, const qsv_gate_t* gate, const unsigned char* host_mask,
                     void* dev_ws, void* stream) {
  const int n = n_total - n_shot_qubits;
  if (!sv || !gate || !host_mask || !dev_ws || n_shot_qubits < 0 || n < 1 || n_total > 40)
    return fail(QSV_EINVAL, "qsv_apply_masked: bad arguments");
  MaskedGate g;
  memset(&g, 0, sizeof(g));
  g.shot_shift = n;
  g.q0 = gate->q0;
  g.q1 = gate->q1;
  const bool two = gate->kind == QSV_GATE_DENSE2 || gate->kind == QSV_GATE_DIAG2 || gate->kind == QSV_GATE_CX;
  g.k = two ? 2 : 1;
  if (g.q0 < 0 || g.q0 >= n || (two && (g.q1 < 0 || g.q1 >= n || g.q1 == g.q0)))
    return fail(QSV_EINVAL, "qsv_apply_masked: gate qubit out of range");
  switch (gate->kind) {  // every kind as a dense matrix in its own qubit order (q0 = index bit 0)
    case QSV_GATE_DENSE1: memcpy(g.m, gate->m, sizeof(double) * 8); break;
    case QSV_GATE_DENSE2: memcpy(g.m, gate->m, sizeof(double) * 32); break;
    case QSV_GATE_DIAG1:
      for (int i = 0; i < 2; i++) { g.m[2 * (2 * i + i)] = gate->m[2 * i]; g.m[2 * (2 * i + i) + 1] = gate->m[2 * i + 1]; }
      break;
    case QSV_GATE_DIAG2:
      for (int i = 0; i < 4; i++) { g.m[2 * (4 * i + i)] = gate->m[2 * i]; g.m[2 * (4 * i + i) + 1] = gate->m[2 * i + 1]; }
      break;
    case QSV_GATE_CX: {
      static const int to[4] = {0, 3, 2, 1};  // control = index bit 0 (basicaertools.py:70-72)
      for (int c = 0; c < 4; c++) g.m[2 * (4 * to[c] + c)] = 1.0;
      break;
    }
    default: return fail(QSV_EINVAL, "qsv_apply_masked: unknown gate kind");
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t shots = (size_t)1 << n_shot_qubits;
  unsigned char* d_mask = reinterpret_cast<unsigned char*>(dev_ws);
  QSV_CUDA(cudaMemcpyAsync(d_mask, host_mask, shots, cudaMemcpyHostToDevice, s));
  const u64 n_groups = 1ull << (n_total - g.k);
  const int blocks = (int)std::max<u64>(1, std::min<u64>((n_groups + kThreads - 1) / kThreads, (u64)sm_count() * 16));
  apply_masked_kernel<<<blocks, kThreads, 0, s>>>(reinterpret_cast<double2*>(sv), n_groups, g, d_mask);
  QSV_LAUNCHED();
  QSV_CUDA(cudaStreamSynchronize(s));
  return 0;
}

static int marginal_log2_parts(int n, int m) {
  // enough CTAs to fill the machine: bins * parts >= ~4 CTAs per SM, parts <= 2^(n-m) / 256
  int lp = 0;
  while (((1ull << m) << lp) < 592ull && (n - m - lp) > 10) lp++;  // fixed (not per device): it sizes the workspace
  return lp;
}

size_t qsv_marginal_workspace_bytes(int n, int m) {
  if (m < 0 || m > n) return 0;
  return sizeof(double) * (((size_t)1 << m) << marginal_log2_parts(n, m));
}

int qsv_marginal(const double* sv, int n, const int32_t* measured_sorted, int m, double* dev_probs, void* ws,
                 void* stream) {
  if (!sv || !measured_sorted || !dev_probs || !ws || n < 1 || n > 40 || m < 0 || m > n)
    return fail(QSV_EINVAL, "qsv_marginal: bad arguments");
  MargDesc d;
  memset(&d, 0, sizeof(d));
  d.n = n;
  d.m = m;
  bool meas[64] = {false};
  for (int j = 0; j < m; j++) {
    const int q = measured_sorted[j];
    if (q < 0 || q >= n || meas[q] || (j > 0 && q <= measured_sorted[j - 1]))
      return fail(QSV_EINVAL, "qsv_marginal: measured qubits must be ascending and unique");
    meas[q] = true;
    d.mq[j] = (uint8_t)q;
  }
  int nu = 0;
  for (int q = 0; q < n; q++)
    if (!meas[q]) d.uq[nu++] = (uint8_t)q;
  d.log2_parts = marginal_log2_parts(n, m);
  const u64 bins = 1ull << m;
  if (n - m <= 10 && bins >= 4096) {
    marginal_small_kernel<<<(unsigned)((bins + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(sv), d, bins, dev_probs);
    QSV_LAUNCHED();
    return 0;
  }
  if (bins > (1ull << 31)) return fail(QSV_ERANGE, "qsv_marginal: too many bins");
  const unsigned gy = (unsigned)std::min<u64>(bins, 32768), gz = (unsigned)(bins / gy);
  dim3 grid(1u << d.log2_parts, gy, gz);
  cudaStream_t s = (cudaStream_t)stream;
  marginal_kernel<<<grid, kThreads, 0, s>>>(reinterpret_cast<const double2*>(sv), d, reinterpret_cast<double*>(ws));
  QSV_LAUNCHED();
  marginal_finish_kernel<<<(unsigned)((bins + 255) / 256), 256, 0, s>>>(reinterpret_cast<double*>(ws), bins,
                                                                        d.log2_parts, dev_probs);
  QSV_LAUNCHED();
  return 0;
}

size_t qsv_sample_workspace_bytes(int log2_bins, int shots) {
  if (log2_bins < 0 || log2_bins > 40 || shots < 0) return 0;
  return sample_ws_layout(log2_bins, shots, nullptr, nullptr);
}

int qsv_cdf_prepare(const double* src, int from_amplitudes, int log2_bins, void* ws, double* host_approx_total,
                    void* stream) {
  if (!src || !ws || !host_approx_total || log2_bins < 0 || log2_bins > 40)
    return fail(QSV_EINVAL, "qsv_cdf_prepare: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  return from_amplitudes ? cdf_prepare<true>(src, log2_bins, ws, host_approx_total, s)
                         : cdf_prepare<false>(src, log2_bins, ws, host_approx_total, s);
}

int qsv_cdf_classify(const double* src, int from_amplitudes, int log2_bins, double approx_offset, void* ws,
                     void* stream) {
  if (!src || !ws || log2_bins < 0 || log2_bins > 40 || !(approx_offset >= 0.0))
    return fail(QSV_EINVAL, "qsv_cdf_classify: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  return from_amplitudes ? cdf_classify<true>(src, log2_bins, approx_offset, ws, s)
                         : cdf_classify<false>(src, log2_bins, approx_offset, ws, s);
}

int qsv_cdf_chain(const double* src, int from_amplitudes, int log2_bins, double exact_start, void* ws,
                  double* host_exact_end, void* stream) {
  if (!src || !ws || !host_exact_end || log2_bins < 0 || log2_bins > 40 || !(exact_start >= 0.0))
    return fail(QSV_EINVAL, "qsv_cdf_chain: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  return from_amplitudes ? cdf_chain<true>(src, log2_bins, exact_start, ws, host_exact_end, s)
                         : cdf_chain<false>(src, log2_bins, exact_start, ws, host_exact_end, s);
}

int qsv_cdf_search(const double* src, int from_amplitudes, int log2_bins, double exact_start, double total,
                   const double* host_uniforms, int shots, int64_t* host_out_idx, void* ws, void* stream) {
  if (!src || !ws || !host_uniforms || !host_out_idx || log2_bins < 0 || log2_bins > 40 || shots < 1 ||
      !(exact_start >= 0.0) || !(total > 0.0))
    return fail(QSV_EINVAL, "qsv_cdf_search: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  return from_amplitudes
             ? cdf_search<true, true>(src, log2_bins, exact_start, total, host_uniforms, shots, host_out_idx, ws, s)
             : cdf_search<false, true>(src, log2_bins, exact_start, total, host_uniforms, shots, host_out_idx, ws, s);
}

int qsv_sample(const double* src, int from_amplitudes, int log2_bins, const double* host_uniforms, int shots,
               int64_t* host_out_idx, int exact_order, void* ws, double* host_total, void* stream) {
  if (!src || !host_uniforms || !host_out_idx || !ws || log2_bins < 0 || log2_bins > 40 || shots < 1)
    return fail(QSV_EINVAL, "qsv_sample: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (from_amplitudes)
    return sample_impl<true>(src, log2_bins, host_uniforms, shots, host_out_idx, exact_order, ws, host_total, s);
  return sample_impl<false>(src, log2_bins, host_uniforms, shots, host_out_idx, exact_order, ws, host_total, s);
}

}  // extern "C"
