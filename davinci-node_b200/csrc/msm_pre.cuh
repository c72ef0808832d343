// Affine pre-reduction of the sorted MSM entries (batched additions with a shared inversion).
//
// A bucket accumulation in XYZZ coordinates costs 10 field multiplications per point.  Adding two AFFINE points costs
// 1 inversion + 3 multiplications, and Montgomery's trick shares one inversion over a whole block's batch at 3
// multiplications per element: 6 M per addition.  The sorted entry list is therefore halved `levels` times before
// the bucket accumulation sees it:   out[j] = in[2j] + in[2j+1]   over the whole list - bucket segments start at
// multiples of 2^levels (padded with sentinels that read as the point at infinity), so a pair never straddles two
// buckets and the kernel needs no per-bucket logic.  After L levels a bucket of k points holds ceil(k / 2^L) partial
// sums (stored as affine points, contiguous per bucket) and (1 - 2^-L) of its additions were done at 6-7 M instead of
// 10 M.
//
// Replaces nothing in the reference by itself: it is part of the bucket method behind `G1Affine.MultiExp`
// (gnark-crypto, go.mod:16; /root/reference/prover/prover_cpu.go:37).  gnark-crypto's CPU MSM uses the same idea
// (batch-affine buckets) for large instances.
#pragma once
#include "ec.cuh"

namespace b200 {

constexpr int kPreThreads = 128;
constexpr uint32_t kPreSentinel = 0xffffffffu;

// operands of pair j
template <class F, bool FROM_TABLE>
__device__ __forceinline__ void pre_load_pair(Affine<F>& p, Affine<F>& q, const Affine<F>* __restrict__ pts,
                                              const uint32_t* __restrict__ idx, uint64_t j) {
  if (FROM_TABLE) {
    const uint32_t e0 = idx[2 * j], e1 = idx[2 * j + 1];
    if (e0 == kPreSentinel) {
      F::set_zero(p.x);
      F::set_zero(p.y);
    } else {
      load16(p, pts + (e0 & 0x7fffffffu));
      if (e0 >> 31) EC<F>::neg(p);
    }
    if (e1 == kPreSentinel) {
      F::set_zero(q.x);
      F::set_zero(q.y);
    } else {
      load16(q, pts + (e1 & 0x7fffffffu));
      if (e1 >> 31) EC<F>::neg(q);
    }
  } else {
    load16(p, pts + 2 * j);
    load16(q, pts + 2 * j + 1);
  }
}

// kind of the addition p + q and its denominator d:  0 = generic (d = x2 - x1), 1 = doubling (d = 2 y1),
// 2 = no inversion needed (an operand is infinity, or p = -q): d = 1
template <class F>
__device__ __forceinline__ int pre_classify(typename F::El& d, const Affine<F>& p, const Affine<F>& q) {
  if (EC<F>::is_inf(p) || EC<F>::is_inf(q)) {
    F::set_one(d);
    return 2;
  }
  F::sub(d, q.x, p.x);
  if (!F::is_zero(d)) return 0;
  if (F::eq(p.y, q.y) && !F::is_zero(p.y)) {
    F::dbl(d, p.y);
    return 1;
  }
  F::set_one(d);
  return 2;
}

// One level  out[j] = in[2j] + in[2j+1]  (j < npairs = *count_slots >> (shift + 1), a device-side value) runs as
//   k_msm_pre_fwd      per thread: running product of its `per_thread` denominators (prefix products to scratch),
//                      thread total to totals[]                                     1 M per addition
//   k_msm_pre_inv_a/b/c  totals[t] <- 1 / totals[t] for ALL threads of the grid with ONE field inversion
//                      (Montgomery's trick in two levels; a few hundred thousand elements, microseconds of work)
//   k_msm_pre_bwd      per thread: peel its denominators off and finish every addition          5 M per addition
// so no thread ever waits at a barrier for an inversion.  Each thread owns `per_thread` pairs, strided by the block
// size so neighbouring threads touch neighbouring entries.
constexpr uint32_t kPreInvChunk = 128;

// x-coordinates decide the common case (both finite, x1 != x2): the forward pass then needs no y at all
template <class F, bool FROM_TABLE>
__device__ __forceinline__ void pre_denominator(typename F::El& d, const Affine<F>* __restrict__ pts,
                                                const uint32_t* __restrict__ idx, uint64_t j) {
  using El = typename F::El;
  El x1, x2;
  bool slow = false;
  if (FROM_TABLE) {
    const uint32_t e0 = idx[2 * j], e1 = idx[2 * j + 1];
    if (e0 == kPreSentinel || e1 == kPreSentinel) {
      slow = true;
    } else {
      load16(x1, &pts[e0 & 0x7fffffffu].x);
      load16(x2, &pts[e1 & 0x7fffffffu].x);
    }
  } else {
    load16(x1, &pts[2 * j].x);
    load16(x2, &pts[2 * j + 1].x);
  }
  if (!slow) {
    F::sub(d, x2, x1);
    slow = F::is_zero(d) || F::is_zero(x1) || F::is_zero(x2);
  }
  if (slow) {                      // infinity operand, doubling or cancellation: classify on the whole points
    Affine<F> p, q;
    pre_load_pair<F, FROM_TABLE>(p, q, pts, idx, j);
    pre_classify<F>(d, p, q);
  }
}

template <class F, bool FROM_TABLE>
__global__ void __launch_bounds__(kPreThreads, 4)
k_msm_pre_fwd(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ count_slots,
              uint32_t shift, uint32_t per_thread, typename F::El* __restrict__ prefix, typename F::El* __restrict__ totals) {
  using El = typename F::El;
  const uint64_t npairs = (uint64_t)(*count_slots >> shift) >> 1;
  const uint64_t block_base = (uint64_t)blockIdx.x * kPreThreads * per_thread;
  const uint64_t tglobal = (uint64_t)blockIdx.x * kPreThreads + threadIdx.x;
  El run;
  F::set_one(run);
  if (block_base < npairs) {
    El* my_prefix = prefix + tglobal;                       // element k at my_prefix[k * stride]
    const uint64_t stride = (uint64_t)gridDim.x * kPreThreads;
    for (uint32_t k = 0; k < per_thread; k++) {
      const uint64_t j = block_base + (uint64_t)k * kPreThreads + threadIdx.x;
      if (j >= npairs) break;
      El d;
      pre_denominator<F, FROM_TABLE>(d, pts, idx, j);
      store16(my_prefix + (uint64_t)k * stride, run);
      F::mul(run, run, d);
    }
  }
  store16(totals + tglobal, run);
}

// ---- totals[i] <- 1 / totals[i] for i < m  (m a multiple of kPreInvChunk; every element non-zero)
// a: thread u: exclusive prefix products of its chunk -> tpre, chunk product -> cprod[u]
template <class F>
__global__ void __launch_bounds__(64) k_msm_pre_inv_a(const typename F::El* __restrict__ totals, uint32_t nchunks,
                                                      typename F::El* __restrict__ tpre, typename F::El* __restrict__ cprod) {
  using El = typename F::El;
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nchunks) return;
  El run;
  F::set_one(run);
  for (uint32_t i = 0; i < kPreInvChunk; i++) {
    const uint64_t e = (uint64_t)u * kPreInvChunk + i;
    El t;
    load16_rw(t, totals + e);
    store16(tpre + e, run);
    F::mul(run, run, t);
  }
  store16(cprod + u, run);
}

// b: ONE block: cprod[u] <- 1 / cprod[u] for u < nchunks, with a single inversion
template <class F>
__global__ void __launch_bounds__(kPreThreads) k_msm_pre_inv_b(typename F::El* __restrict__ cprod, uint32_t nchunks,
                                                                typename F::El* __restrict__ scratch) {
  using El = typename F::El;
  __shared__ El sm_pre[kPreThreads];
  __shared__ El sm_suf[kPreThreads];
  __shared__ El sm_inv;
  const uint32_t per = (nchunks + kPreThreads - 1) / kPreThreads;
  const uint32_t lo = threadIdx.x * per, hi = lo + per < nchunks ? lo + per : nchunks;
  El run;
  F::set_one(run);
  for (uint32_t u = lo; u < hi; u++) {
    El t;
    load16_rw(t, cprod + u);
    store16(scratch + u, run);                            // exclusive prefix inside this thread's range
    F::mul(run, run, t);
  }
  sm_pre[threadIdx.x] = run;
  sm_suf[threadIdx.x] = run;
  __syncthreads();
  for (int s = 1; s < kPreThreads; s <<= 1) {
    El a, b;
    const bool up = (int)threadIdx.x >= s, dn = (int)threadIdx.x + s < kPreThreads;
    if (up) F::mul(a, sm_pre[threadIdx.x], sm_pre[threadIdx.x - s]);
    if (dn) F::mul(b, sm_suf[threadIdx.x], sm_suf[threadIdx.x + s]);
    __syncthreads();
    if (up) sm_pre[threadIdx.x] = a;
    if (dn) sm_suf[threadIdx.x] = b;
    __syncthreads();
  }
  if (threadIdx.x == 0) F::inv(sm_inv, sm_pre[kPreThreads - 1]);
  __syncthreads();
  El rinv = sm_inv;
  if (threadIdx.x > 0) F::mul(rinv, rinv, sm_pre[threadIdx.x - 1]);
  if (threadIdx.x + 1 < kPreThreads) F::mul(rinv, rinv, sm_suf[threadIdx.x + 1]);
  for (uint32_t u = hi; u > lo; u--) {
    El t, pre, o;
    load16_rw(t, cprod + (u - 1));
    load16_rw(pre, scratch + (u - 1));
    F::mul(o, rinv, pre);
    F::mul(rinv, rinv, t);
    store16(cprod + (u - 1), o);
  }
}

// c: thread u: back-substitution inside its chunk
template <class F>
__global__ void __launch_bounds__(64) k_msm_pre_inv_c(typename F::El* __restrict__ totals, uint32_t nchunks,
                                                      const typename F::El* __restrict__ tpre,
                                                      const typename F::El* __restrict__ cinv) {
  using El = typename F::El;
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nchunks) return;
  El rinv;
  load16_rw(rinv, cinv + u);
  for (int i = (int)kPreInvChunk - 1; i >= 0; i--) {
    const uint64_t e = (uint64_t)u * kPreInvChunk + i;
    El t, pre, o;
    load16_rw(t, totals + e);
    load16_rw(pre, tpre + e);
    F::mul(o, rinv, pre);
    F::mul(rinv, rinv, t);
    store16(totals + e, o);
  }
}

template <class F, bool FROM_TABLE>
__global__ void __launch_bounds__(kPreThreads, 3)
k_msm_pre_bwd(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ count_slots,
              uint32_t shift, uint32_t per_thread, Affine<F>* __restrict__ out, const typename F::El* __restrict__ prefix,
              const typename F::El* __restrict__ totals) {
  using El = typename F::El;
  const uint64_t npairs = (uint64_t)(*count_slots >> shift) >> 1;
  const uint64_t block_base = (uint64_t)blockIdx.x * kPreThreads * per_thread;
  if (block_base >= npairs) return;
  const uint64_t tglobal = (uint64_t)blockIdx.x * kPreThreads + threadIdx.x;
  const El* my_prefix = prefix + tglobal;
  const uint64_t stride = (uint64_t)gridDim.x * kPreThreads;
  // this thread's pairs are k = 0 .. kmax-1
  int kmax = 0;
  if (block_base + threadIdx.x < npairs) {
    const uint64_t span = npairs - block_base - threadIdx.x;            // >= 1
    const uint64_t cnt = (span + kPreThreads - 1) / kPreThreads;
    kmax = (int)(cnt < per_thread ? cnt : per_thread);
  }
  if (kmax == 0) return;
  El rinv;
  load16_rw(rinv, totals + tglobal);                                     // 1 / (product of this thread's denominators)
  // software pipeline: the operands of the next (lower) pair are in flight while this addition is finished
  Affine<F> np, nq;
  pre_load_pair<F, FROM_TABLE>(np, nq, pts, idx, block_base + (uint64_t)(kmax - 1) * kPreThreads + threadIdx.x);
  for (int k = kmax - 1; k >= 0; k--) {
    const uint64_t j = block_base + (uint64_t)k * kPreThreads + threadIdx.x;
    Affine<F> p = np, q = nq, r;
    if (k > 0) pre_load_pair<F, FROM_TABLE>(np, nq, pts, idx, j - kPreThreads);
    El d, pre, dinv;
    const int kind = pre_classify<F>(d, p, q);
    load16_rw(pre, my_prefix + (uint64_t)k * stride);
    F::mul(dinv, rinv, pre);
    F::mul(rinv, rinv, d);
    if (kind == 2) {
      if (EC<F>::is_inf(p)) r = q;
      else if (EC<F>::is_inf(q)) r = p;
      else {                                               // p = -q
        F::set_zero(r.x);
        F::set_zero(r.y);
      }
    } else {
      El num, lam, t;
      if (kind == 0) {
        F::sub(num, q.y, p.y);
      } else {                                             // 3 x^2
        F::sqr(t, p.x);
        F::dbl(num, t);
        F::add(num, num, t);
      }
      F::mul(lam, num, dinv);
      F::sqr(t, lam);
      F::sub(t, t, p.x);
      F::sub(r.x, t, q.x);
      F::sub(t, p.x, r.x);
      F::mul(t, lam, t);
      F::sub(r.y, t, p.y);
    }
    store16(out + j, r);
  }
}

// bucket b of the reduced list: [off[b] >> levels, + ceil(cnt / 2^levels)) ; totals2 = padded total >> levels
static __global__ void __launch_bounds__(256)
k_msm_pre_offsets(const uint32_t* __restrict__ off, const uint32_t* __restrict__ end, uint64_t total_b, uint32_t levels,
                  uint32_t* __restrict__ off2, uint32_t* __restrict__ end2, const uint32_t* __restrict__ totals,
                  uint32_t* __restrict__ totals2, uint32_t narrays) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < narrays) totals2[b] = totals[b] >> levels;
  if (b >= total_b) return;
  const uint32_t o = off[b], cnt = end[b] - o;
  off2[b] = o >> levels;
  end2[b] = (o >> levels) + ((cnt + (1u << levels) - 1u) >> levels);
}

}  // namespace b200
