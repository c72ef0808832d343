// Radix-2^k number-theoretic transforms over Fr for sm_100a.
//
// A transform of size n = 2^logn runs as 1-3 passes (the quotient's 7 transforms as 15 instead of 21: see the fused
// kernels below); each pass stages a tile of up to 2^11
// elements in shared memory and performs up to 11 butterfly levels there (twiddles ω^k, k < n/2,
// are a table resident in HBM / L2).  Strided passes load `lo_tile` consecutive elements per row so
// global accesses stay in >= 256-byte runs.  DIF (Gentleman-Sande, natural -> bit-reversed) and DIT
// (Cooley-Tukey, bit-reversed -> natural) are paired so no bit-reversal permutation is ever done.
// Element-wise work is fused into the first / last pass: coset scaling by g^i (two-level power
// table), 1/n, and the quotient's pointwise (a*b - c) * 1/(g^n - 1).
//
// Replaces gnark-crypto `fft.Domain.FFT / FFTInverse` and gnark's `computeH`
// (reached from /root/reference/prover/prover_cpu.go:37).
#pragma once
#include "field.cuh"

namespace b200 {

enum NttScaleMode { SCALE_NONE = 0, SCALE_CONST = 1, SCALE_POW_NATURAL = 2, SCALE_POW_BITREV = 3 };

// factor(position i) = lo[e & (2^lo_bits - 1)] * hi[e >> lo_bits],  e = i or bitrev(i)
struct NttScale {
  int mode;
  int lo_bits;
  const void* lo;
  const void* hi;
  const void* cst;
};

struct NttPass {
  int logn;
  int s_log;         // log2 of the element stride between butterfly rows in this pass
  int logr;          // butterfly levels done in this pass
  int lo_tile_log;   // log2 of consecutive elements loaded per row
};

constexpr int kNttThreads = 256;

// The shared-memory tile is stored chunk-major: 16-byte chunk k of element i lives at tile[k * E + i].  With the
// natural array-of-elements layout consecutive threads would read 16-byte pieces 32 (or 48) bytes apart, a 2-way
// bank conflict on every 128-bit access; chunk-major rows are contiguous and conflict-free.
template <class El>
__device__ __forceinline__ void tile_load(El& x, const uint4* tile, uint32_t E, uint32_t i) {
  uint4* d = reinterpret_cast<uint4*>(&x);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(El) / 16); k++) d[k] = tile[k * E + i];
}
template <class El>
__device__ __forceinline__ void tile_store(uint4* tile, uint32_t E, uint32_t i, const El& x) {
  const uint4* s = reinterpret_cast<const uint4*>(&x);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(El) / 16); k++) tile[k * E + i] = s[k];
}

template <class Fr>
__device__ __forceinline__ void ntt_apply_scale(typename Fr::El& x, const NttScale& sc, uint32_t i, int logn) {
  using El = typename Fr::El;
  if (sc.mode == SCALE_NONE) return;
  if (sc.mode == SCALE_CONST) {
    El c;
    load16(c, reinterpret_cast<const El*>(sc.cst));
    Fr::mul(x, x, c);
    return;
  }
  uint32_t e = sc.mode == SCALE_POW_BITREV ? (__brev(i) >> (32 - logn)) : i;
  El l, h;
  load16(l, reinterpret_cast<const El*>(sc.lo) + (e & ((1u << sc.lo_bits) - 1u)));
  load16(h, reinterpret_cast<const El*>(sc.hi) + (e >> sc.lo_bits));
  Fr::mul(l, l, h);
  Fr::mul(x, x, l);
}

// ps.logr butterfly levels over a tile of 2^(logr + lo_tile_log) elements held chunk-major in shared memory; ends with
// a barrier.  lo0: position of the tile's first column inside its row group (twiddle exponents depend on it).
template <class Fr, bool DIT>
__device__ __forceinline__ void ntt_tile_levels(uint4* sm, const NttPass& ps, const typename Fr::El* __restrict__ tw,
                                                uint32_t lo0) {
  using El = typename Fr::El;
  const uint32_t E = 1u << (ps.logr + ps.lo_tile_log);
  const uint32_t lo_mask = (1u << ps.lo_tile_log) - 1u;
  const uint32_t half = E >> 1;
  for (int j = 0; j < ps.logr; j++) {
    const int log_dm = DIT ? j : (ps.logr - 1 - j);
    const int shift = ps.logn - 1 - ps.s_log - log_dm;   // twiddle exponent scale: d * 2^shift = n/2
    const uint32_t dm_mask = (1u << log_dm) - 1u;
    for (uint32_t q = threadIdx.x; q < half; q += blockDim.x) {
      uint32_t lo_l = q & lo_mask;
      uint32_t u = q >> ps.lo_tile_log;
      uint32_t mid_low = u & dm_mask;
      uint32_t mid = ((u >> log_dm) << (log_dm + 1)) | mid_low;
      uint32_t i0 = (mid << ps.lo_tile_log) | lo_l;
      uint32_t i1 = i0 + (1u << (log_dm + ps.lo_tile_log));
      uint32_t e = ((mid_low << ps.s_log) + lo0 + lo_l) << shift;
      El x, y, w;
      tile_load(x, sm, E, i0);
      tile_load(y, sm, E, i1);
      if (DIT) {
        if (e) {
          load16(w, tw + e);
          Fr::mul(y, y, w);
        }
        El t;
        Fr::add(t, x, y);
        Fr::sub(y, x, y);
        tile_store(sm, E, i0, t);
        tile_store(sm, E, i1, y);
      } else {
        El t;
        Fr::add(t, x, y);
        Fr::sub(y, x, y);
        if (e) {
          load16(w, tw + e);
          Fr::mul(y, y, w);
        }
        tile_store(sm, E, i0, t);
        tile_store(sm, E, i1, y);
      }
    }
    __syncthreads();
  }
}

// One pass.  If in_b != nullptr the load computes (data[i] * in_b[i] - in_c[i]) * den  (quotient).
template <class Fr, bool DIT>
__global__ void __launch_bounds__(kNttThreads)
k_ntt_pass(typename Fr::El* __restrict__ data, const typename Fr::El* __restrict__ tw, NttPass ps, NttScale pre,
           NttScale post, const typename Fr::El* __restrict__ in_b, const typename Fr::El* __restrict__ in_c,
           const typename Fr::El* __restrict__ den) {
  using El = typename Fr::El;
  extern __shared__ uint4 sm[];

  const int elog = ps.logr + ps.lo_tile_log;
  const uint32_t E = 1u << elog;
  const uint32_t lo_mask = (1u << ps.lo_tile_log) - 1u;
  const uint32_t tiles_per_hi = 1u << (ps.s_log - ps.lo_tile_log);
  const uint32_t tile = blockIdx.x;
  const uint32_t hi = tile / tiles_per_hi;
  const uint32_t lo0 = (tile % tiles_per_hi) << ps.lo_tile_log;
  const uint64_t base = ((uint64_t)hi << (ps.s_log + ps.logr)) + lo0;

  // ---- load (+ fused element-wise work)
  for (uint32_t l = threadIdx.x; l < E; l += kNttThreads) {
    uint32_t mid = l >> ps.lo_tile_log, lo_l = l & lo_mask;
    uint64_t gi = base + ((uint64_t)mid << ps.s_log) + lo_l;
    El x;
    load16_rw(x, data + gi);
    if (in_b) {
      El b, c, d;
      load16(b, in_b + gi);
      load16(c, in_c + gi);
      load16(d, den);
      Fr::mul(x, x, b);
      Fr::sub(x, x, c);
      Fr::mul(x, x, d);
    }
    ntt_apply_scale<Fr>(x, pre, (uint32_t)gi, ps.logn);
    tile_store(sm, E, l, x);
  }
  __syncthreads();

  // ---- butterflies
  ntt_tile_levels<Fr, DIT>(sm, ps, tw, lo0);

  // ---- store (+ fused scaling)
  for (uint32_t l = threadIdx.x; l < E; l += kNttThreads) {
    uint32_t mid = l >> ps.lo_tile_log, lo_l = l & lo_mask;
    uint64_t gi = base + ((uint64_t)mid << ps.s_log) + lo_l;
    El x;
    tile_load(x, sm, E, l);
    ntt_apply_scale<Fr>(x, post, (uint32_t)gi, ps.logn);
    store16(data + gi, x);
  }
}

// Fused middle of "interpolate, then evaluate on the coset" (the quotient does it for a, b and c): the LAST pass of the
// unscaled inverse DIF transform and the FIRST pass of the coset DIT transform both work on the same contiguous tile
// of 2^lc elements, so the tile makes one round trip through shared memory instead of two through HBM:
// lc DIF levels (omega^-1 twiddles), the coset / 1/n scaling (bit-reversed exponent), lc DIT levels (omega twiddles).
template <class Fr>
__global__ void __launch_bounds__(kNttThreads)
k_ntt_fused_mid(typename Fr::El* __restrict__ data, const typename Fr::El* __restrict__ tw_inv,
                const typename Fr::El* __restrict__ tw_fwd, NttPass ps, NttScale scale) {
  using El = typename Fr::El;
  extern __shared__ uint4 sm[];
  const uint32_t E = 1u << ps.logr;                   // contiguous pass: s_log = 0, lo_tile_log = 0
  const uint64_t base = (uint64_t)blockIdx.x << ps.logr;
  for (uint32_t l = threadIdx.x; l < E; l += blockDim.x) {
    El x;
    load16_rw(x, data + base + l);
    tile_store(sm, E, l, x);
  }
  __syncthreads();
  ntt_tile_levels<Fr, false>(sm, ps, tw_inv, 0);
  for (uint32_t l = threadIdx.x; l < E; l += blockDim.x) {
    El x;
    tile_load(x, sm, E, l);
    ntt_apply_scale<Fr>(x, scale, (uint32_t)(base + l), ps.logn);
    tile_store(sm, E, l, x);
  }
  __syncthreads();
  ntt_tile_levels<Fr, true>(sm, ps, tw_fwd, 0);
  for (uint32_t l = threadIdx.x; l < E; l += blockDim.x) {
    El x;
    tile_load(x, sm, E, l);
    store16(data + base + l, x);
  }
}

// Fused end of the quotient: the LAST (strided) pass of the three coset DIT transforms, the pointwise
// (a * b - c) / (g^n - 1) and the FIRST (strided) pass of the inverse coset DIF transform all address the same tile
// geometry.  Three tiles live in shared memory (192 KB for 32-byte elements); only `a` is written back.
constexpr int kNttQuotThreads = 512;
template <class Fr>
__global__ void __launch_bounds__(kNttQuotThreads)
k_ntt_fused_quot(typename Fr::El* __restrict__ a, const typename Fr::El* __restrict__ b,
                 const typename Fr::El* __restrict__ c, const typename Fr::El* __restrict__ tw_fwd,
                 const typename Fr::El* __restrict__ tw_inv, NttPass ps, const typename Fr::El* __restrict__ den) {
  using El = typename Fr::El;
  extern __shared__ uint4 sm[];
  const int elog = ps.logr + ps.lo_tile_log;
  const uint32_t E = 1u << elog;
  const uint32_t chunks = sizeof(El) / 16;
  uint4* ta = sm;
  uint4* tb = sm + (size_t)chunks * E;
  uint4* tc = tb + (size_t)chunks * E;
  const uint32_t lo_mask = (1u << ps.lo_tile_log) - 1u;
  const uint32_t tiles_per_hi = 1u << (ps.s_log - ps.lo_tile_log);
  const uint32_t hi = blockIdx.x / tiles_per_hi;
  const uint32_t lo0 = (blockIdx.x % tiles_per_hi) << ps.lo_tile_log;
  const uint64_t base = ((uint64_t)hi << (ps.s_log + ps.logr)) + lo0;
  for (uint32_t l = threadIdx.x; l < E; l += blockDim.x) {
    const uint64_t gi = base + ((uint64_t)(l >> ps.lo_tile_log) << ps.s_log) + (l & lo_mask);
    El x;
    load16_rw(x, a + gi);
    tile_store(ta, E, l, x);
    load16(x, b + gi);
    tile_store(tb, E, l, x);
    load16(x, c + gi);
    tile_store(tc, E, l, x);
  }
  __syncthreads();
  ntt_tile_levels<Fr, true>(ta, ps, tw_fwd, lo0);
  ntt_tile_levels<Fr, true>(tb, ps, tw_fwd, lo0);
  ntt_tile_levels<Fr, true>(tc, ps, tw_fwd, lo0);
  El d;
  load16(d, den);
  for (uint32_t l = threadIdx.x; l < E; l += blockDim.x) {
    El x, y, z;
    tile_load(x, ta, E, l);
    tile_load(y, tb, E, l);
    tile_load(z, tc, E, l);
    Fr::mul(x, x, y);
    Fr::sub(x, x, z);
    Fr::mul(x, x, d);
    tile_store(ta, E, l, x);
  }
  __syncthreads();
  ntt_tile_levels<Fr, false>(ta, ps, tw_inv, lo0);
  for (uint32_t l = threadIdx.x; l < E; l += blockDim.x) {
    const uint64_t gi = base + ((uint64_t)(l >> ps.lo_tile_log) << ps.s_log) + (l & lo_mask);
    El x;
    tile_load(x, ta, E, l);
    store16(a + gi, x);
  }
}

// out[k] = scale * base^(k * step)  for k < count   (square-and-multiply per thread)
template <class Fr>
__global__ void k_pow_table(typename Fr::El* __restrict__ out, const typename Fr::El* __restrict__ base_p,
                            const typename Fr::El* __restrict__ scale_p, uint64_t count, uint64_t step) {
  using El = typename Fr::El;
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  El base = *base_p, acc;
  if (scale_p) acc = *scale_p;
  else Fr::set_one(acc);
  uint64_t e = k * step;
  while (e) {
    if (e & 1) Fr::mul(acc, acc, base);
    e >>= 1;
    if (e) Fr::sqr(base, base);
  }
  out[k] = acc;
}

// Domain constants from (omega, g): consts[0]=omega^-1, [1]=g^-1, [2]=1/n, [3]=1/(g^n - 1), [4]=omega, [5]=g
template <class Fr>
__global__ void k_domain_consts(typename Fr::El* __restrict__ consts, const typename Fr::El* __restrict__ omega,
                                const typename Fr::El* __restrict__ g, int logn) {
  using El = typename Fr::El;
  if (threadIdx.x || blockIdx.x) return;
  El w = *omega, gg = *g, t, one;
  Fr::set_one(one);
  Fr::inv(consts[0], w);
  Fr::inv(consts[1], gg);
  // n as a field element: 2^logn in Montgomery form
  El nn = one;
  for (int i = 0; i < logn; i++) Fr::dbl(nn, nn);
  Fr::inv(consts[2], nn);
  t = gg;
  for (int i = 0; i < logn; i++) Fr::sqr(t, t);
  Fr::sub(t, t, one);
  Fr::inv(consts[3], t);
  consts[4] = w;
  consts[5] = gg;
}

// x[i] = x[i] * c  (element-wise; used for challenge-scaled commitment scalars)
template <class Fr>
__global__ void k_scale_vec(typename Fr::El* __restrict__ x, const typename Fr::El* __restrict__ c, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename Fr::El v, k;
  load16_rw(v, x + i);
  load16(k, c);
  Fr::mul(v, v, k);
  store16(x + i, v);
}

}  // namespace b200
