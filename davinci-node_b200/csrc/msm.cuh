// Signed-digit Pippenger multi-scalar multiplication for sm_100a.
//
//   digits/histogram -> per-window exclusive scan -> counting-sort scatter (point index | sign)
//   -> bucket accumulation (one thread per bucket, 128-bit gathered affine loads, XYZZ mixed adds;
//      oversized buckets are split into fixed-size tasks and merged by a block reduction)
//   -> bucket reduction (running sums over groups of buckets, [base]*sum fix-up per group)
//   -> per-window block reduction -> Horner combine of the windows.
//
// Everything runs on one stream with no host synchronisation; scalars are consumed in
// gnark-crypto's Montgomery form (converted in-register), points in gnark's affine layout.
//
// Replaces `G1Affine.MultiExp` / `G2Affine.MultiExp` (gnark-crypto, go.mod:16) as called for the
// Ar / Bs1 / Bs / Krs / Krs2 and Pedersen commitments of groth16.Prove
// (/root/reference/prover/prover_cpu.go:37).
#pragma once
#include "ec.cuh"
#include "msm_plan.h"

namespace b200 {

// ------------------------------------------------------------------------------------ digits
// canonical scalar -> signed digit of window w (c <= 30).  `s` has NS limbs.
template <int NS>
__device__ __forceinline__ uint32_t window_bits(const uint32_t* s, int bit, int c) {
  int limb = bit >> 5, off = bit & 31;
  if (limb >= NS) return 0;
  uint64_t two = s[limb];
  if (limb + 1 < NS) two |= (uint64_t)s[limb + 1] << 32;
  return (uint32_t)(two >> off) & ((1u << c) - 1u);
}

// Walk all windows of one scalar, calling f(w, bucket_index (0-based), negative)
template <class Fr, class Fn>
__device__ __forceinline__ void for_each_digit(const typename Fr::El& mont, const MsmPlan& pl, Fn f) {
  typename Fr::El s;
  Fr::from_mont(s, mont);
  if (Fr::is_zero(s)) return;
  uint32_t carry = 0;
  const uint32_t half = 1u << (pl.c - 1);
  for (int w = 0; w < pl.nwin; w++) {
    uint32_t d = window_bits<Fr::N>(s.v, w * pl.c, pl.c) + carry;
    carry = 0;
    bool neg = false;
    if (d > half) {
      d = (1u << pl.c) - d;
      neg = true;
      carry = 1;
    }
    if (d) f(w, d - 1, neg);
  }
}

template <class Fr>
__global__ void k_msm_hist(const typename Fr::El* __restrict__ scalars, MsmPlan pl, uint32_t* __restrict__ hist,
                           const uint32_t* __restrict__ index_map) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pl.n) return;
  if (index_map && index_map[i] == 0xffffffffu) return;
  typename Fr::El s;
  load16(s, scalars + i);
  const bool table = pl.bwin == 1;
  for_each_digit<Fr>(s, pl, [&](int w, uint32_t b, bool) { atomicAdd(&hist[(table ? 0 : (uint64_t)w * pl.nb) + b], 1u); });
}

// one block per bucket window: exclusive scan of hist -> off (start offsets) and cur (running cursors);
// also totals[w] = number of (point, digit) entries of the window.  Each thread owns kScanItems
// consecutive buckets per sweep, so 2^19 buckets take 32 sweeps instead of 512.
constexpr int kScanItems = 16;
static __global__ void __launch_bounds__(1024)
k_msm_scan(const uint32_t* __restrict__ hist, MsmPlan pl, uint32_t* __restrict__ off, uint32_t* __restrict__ cur,
           uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t* h = hist + (uint64_t)blockIdx.x * pl.nb;
  uint32_t* o = off + (uint64_t)blockIdx.x * pl.nb;
  uint32_t* c = cur + (uint64_t)blockIdx.x * pl.nb;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t per_sweep = blockDim.x * kScanItems;
  for (uint32_t base = 0; base < pl.nb; base += per_sweep) {
    uint32_t first = base + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      v[k] = first + k < pl.nb ? h[first + k] : 0;
      sum += v[k];
    }
    uint32_t x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t ws = lane < nw ? warp_sums[lane] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, ws, d);
        if (lane >= d) ws += y;
      }
      warp_sums[lane] = ws;   // inclusive
    }
    __syncthreads();
    uint32_t prefix = carry_s + (wid ? warp_sums[wid - 1] : 0) + (x - sum);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      if (first + k < pl.nb) {
        o[first + k] = prefix;
        c[first + k] = prefix;
      }
      prefix += v[k];
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = prefix;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

template <class Fr>
__global__ void k_msm_scatter(const typename Fr::El* __restrict__ scalars, MsmPlan pl, uint32_t* __restrict__ cur,
                              uint32_t* __restrict__ sorted, const uint32_t* __restrict__ index_map) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pl.n) return;
  uint32_t pidx = (uint32_t)i;
  if (index_map) {
    pidx = index_map[i];
    if (pidx == 0xffffffffu) return;
  }
  typename Fr::El s;
  load16(s, scalars + i);
  const bool table = pl.bwin == 1;
  for_each_digit<Fr>(s, pl, [&](int w, uint32_t b, bool neg) {
    uint32_t pos = atomicAdd(&cur[(table ? 0 : (uint64_t)w * pl.nb) + b], 1u);
    // table mode: digit window w reads the precomputed multiple 2^(c w) P, stored at w * npts + i
    uint32_t entry = table ? (uint32_t)((uint64_t)w * pl.npts + pidx) : pidx;
    sorted[(table ? 0 : (uint64_t)w * pl.stride) + pos] = entry | (neg ? 0x80000000u : 0u);
  });
}

// ------------------------------------------------------------------------------------ accumulate
struct OvfTask {
  uint32_t bucket;   // global bucket id  w * nb + b
  uint32_t start;    // offset inside the window's sorted segment
  uint32_t len;
  uint32_t pad;
};
struct OvfBucket {
  uint32_t bucket;
  uint32_t first;    // first task index
  uint32_t ntasks;
  uint32_t pad;
};
struct OvfCounters {
  uint32_t ntasks;
  uint32_t nbuckets;
};

template <class F>
__device__ __forceinline__ void accumulate_run(XYZZ<F>& acc, const Affine<F>* __restrict__ points,
                                               const uint32_t* __restrict__ idx, uint32_t len) {
  using E = EC<F>;
  E::set_inf(acc);
  if (len == 0) return;
  // software pipeline: the next point is in flight while the current mixed add runs
  uint32_t e = idx[0];
  Affine<F> nxt;
  load16(nxt, points + (e & 0x7fffffffu));
  for (uint32_t k = 0; k < len; k++) {
    Affine<F> cur = nxt;
    bool neg = e >> 31;
    if (k + 1 < len) {
      e = idx[k + 1];
      load16(nxt, points + (e & 0x7fffffffu));
    }
    if (neg) E::neg(cur);
    E::madd(acc, cur);
  }
}

// Lock-step variant: every warp of the block runs the same trip count and meets at a barrier each
// iteration, so the (64 KB, larger than the instruction cache) loop body is streamed once per SM and
// iteration instead of once per warp.  Buckets of a block have (almost) equal sizes thanks to the
// size-sorted schedule, so the padding iterations are few.
template <class F>
__device__ __forceinline__ void accumulate_run_lockstep(XYZZ<F>& acc, const Affine<F>* __restrict__ points,
                                                        const uint32_t* __restrict__ idx, uint32_t len) {
  using E = EC<F>;
  __shared__ uint32_t s_trips;
  E::set_inf(acc);
  if (threadIdx.x == 0) s_trips = 0;
  __syncthreads();
  if (len) atomicMax(&s_trips, len);
  __syncthreads();
  const uint32_t trips = s_trips;
  uint32_t e = 0;
  Affine<F> nxt;
  if (len) {
    e = idx[0];
    load16(nxt, points + (e & 0x7fffffffu));
  }
  for (uint32_t k = 0; k < trips; k++) {
    __syncthreads();
    if (k < len) {
      Affine<F> cur = nxt;
      bool neg = e >> 31;
      if (k + 1 < len) {
        e = idx[k + 1];
        load16(nxt, points + (e & 0x7fffffffu));
      }
      if (neg) E::neg(cur);
      E::madd(acc, cur);
    }
  }
}

// ---- schedule: counting sort of the buckets by (clamped) size, largest first, so that the 32
// buckets of a warp have equal trip counts (no divergence) and the heavy buckets start first (no tail)
constexpr int kSizeBins = 1024;
constexpr int kSizeThreads = 1024;
constexpr int kSizePerThread = 8;

static __device__ __forceinline__ uint32_t size_bin(uint32_t cnt) { return cnt < kSizeBins ? cnt : kSizeBins - 1; }

static __global__ void __launch_bounds__(kSizeThreads)
k_msm_size_hist(const uint32_t* __restrict__ off, const uint32_t* __restrict__ end, uint64_t total_b,
                uint32_t* __restrict__ bins) {
  __shared__ uint32_t sh[kSizeBins];
  for (int b = threadIdx.x; b < kSizeBins; b += kSizeThreads) sh[b] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kSizeThreads * kSizePerThread;
  for (int k = 0; k < kSizePerThread; k++) {
    uint64_t gb = base + (uint64_t)k * kSizeThreads + threadIdx.x;
    if (gb < total_b) atomicAdd(&sh[size_bin(end[gb] - off[gb])], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSizeBins; b += kSizeThreads)
    if (sh[b]) atomicAdd(&bins[b], sh[b]);
}

// cursor[b] = number of buckets in strictly larger bins (descending order); single block
static __global__ void __launch_bounds__(kSizeBins) k_msm_size_scan(const uint32_t* __restrict__ bins,
                                                                     uint32_t* __restrict__ cursor) {
  __shared__ uint32_t sh[kSizeBins];
  int t = threadIdx.x;
  sh[t] = bins[kSizeBins - 1 - t];   // reversed: index 0 = largest bin
  __syncthreads();
  for (int d = 1; d < kSizeBins; d <<= 1) {
    uint32_t v = t >= d ? sh[t - d] : 0;
    __syncthreads();
    sh[t] += v;
    __syncthreads();
  }
  cursor[kSizeBins - 1 - t] = sh[t] - bins[kSizeBins - 1 - t];   // exclusive
}

static __global__ void __launch_bounds__(kSizeThreads)
k_msm_size_scatter(const uint32_t* __restrict__ off, const uint32_t* __restrict__ end, uint64_t total_b,
                   uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
  __shared__ uint32_t cnt[kSizeBins];
  __shared__ uint32_t basep[kSizeBins];
  for (int b = threadIdx.x; b < kSizeBins; b += kSizeThreads) cnt[b] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kSizeThreads * kSizePerThread;
  uint32_t mybin[kSizePerThread], mypos[kSizePerThread];
  for (int k = 0; k < kSizePerThread; k++) {
    uint64_t gb = base + (uint64_t)k * kSizeThreads + threadIdx.x;
    mybin[k] = 0xffffffffu;
    if (gb < total_b) {
      mybin[k] = size_bin(end[gb] - off[gb]);
      mypos[k] = atomicAdd(&cnt[mybin[k]], 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSizeBins; b += kSizeThreads)
    basep[b] = cnt[b] ? atomicAdd(&cursor[b], cnt[b]) : 0;
  __syncthreads();
  for (int k = 0; k < kSizePerThread; k++) {
    uint64_t gb = base + (uint64_t)k * kSizeThreads + threadIdx.x;
    if (mybin[k] != 0xffffffffu) perm[basep[mybin[k]] + mypos[k]] = (uint32_t)gb;
  }
}

constexpr uint32_t kOvfTask = kOvfTaskPoints;

template <class F>
#ifndef B200_ACC_NO_LOCKSTEP
#define B200_ACC_LOCKSTEP 1   // measured: -5% (G1) / -8% (G2) accumulate time at 128 threads, 3 blocks per SM
#endif
#ifndef B200_ACC_MIN_BLOCKS
#define B200_ACC_MIN_BLOCKS 3
#endif
#ifndef B200_ACC_THREADS
#define B200_ACC_THREADS 128
#endif
__global__ void __launch_bounds__(B200_ACC_THREADS, B200_ACC_MIN_BLOCKS)
k_msm_accumulate(const Affine<F>* __restrict__ points, const uint32_t* __restrict__ sorted,
                 const uint32_t* __restrict__ off, const uint32_t* __restrict__ end,
                 const uint32_t* __restrict__ perm, const uint32_t* __restrict__ totals, MsmPlan pl,
                 XYZZ<F>* __restrict__ buckets, OvfTask* __restrict__ tasks, OvfBucket* __restrict__ obuckets,
                 OvfCounters* __restrict__ ctr) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < (uint64_t)pl.bwin * pl.nb;
#ifndef B200_ACC_LOCKSTEP
  if (!valid) return;
#endif
  const uint32_t gb = valid ? perm[t] : 0;
  uint32_t w = gb / pl.nb;
  uint32_t start = valid ? off[gb] : 0, cnt = valid ? end[gb] - start : 0;
  uint32_t mine = cnt;
  // the cap follows the ACTUAL mean bucket load of this window (sparse witness vectors fill far fewer
  // digits than n * nwin): a single thread's chain of additions is pure latency (~8 us each)
  uint32_t task = 8u * (totals[w] / pl.nb + 1u);
  task = task < pl.task_min ? pl.task_min : (task > pl.task ? pl.task : task);
  if (cnt > task) {
    mine = task;
    uint32_t extra = (cnt - task + pl.ovf_task - 1) / pl.ovf_task;
    uint32_t first = atomicAdd(&ctr->ntasks, extra);
    if (first + extra <= pl.max_ovf) {
      uint32_t ob = atomicAdd(&ctr->nbuckets, 1u);
      obuckets[ob] = OvfBucket{gb, first, extra, 0};
      uint32_t s = start + task, left = cnt - task;
      for (uint32_t k = 0; k < extra; k++) {
        uint32_t l = left < pl.ovf_task ? left : pl.ovf_task;
        tasks[first + k] = OvfTask{gb, s, l, 0};
        s += l;
        left -= l;
      }
    } else {
      mine = cnt;   // cannot happen (capacity is n*nwin/kOvfTask + 1); stay correct anyway
    }
  }
  XYZZ<F> acc;
#ifdef B200_ACC_LOCKSTEP
  accumulate_run_lockstep<F>(acc, points, sorted + (uint64_t)w * pl.stride + start, mine);
  if (valid) store16(buckets + gb, acc);
#else
  accumulate_run<F>(acc, points, sorted + (uint64_t)w * pl.stride + start, mine);
  store16(buckets + gb, acc);
#endif
}

template <class F>
__global__ void __launch_bounds__(128)
k_msm_ovf_accumulate(const Affine<F>* __restrict__ points, const uint32_t* __restrict__ sorted, MsmPlan pl,
                     const OvfTask* __restrict__ tasks, const OvfCounters* __restrict__ ctr,
                     XYZZ<F>* __restrict__ partial) {
  uint32_t nt = ctr->ntasks < pl.max_ovf ? ctr->ntasks : pl.max_ovf;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    OvfTask tk = tasks[t];
    uint32_t w = tk.bucket / pl.nb;
    XYZZ<F> acc;
    accumulate_run<F>(acc, points, sorted + (uint64_t)w * pl.stride + tk.start, tk.len);
    store16(partial + t, acc);
  }
}

// Block-wide sum of XYZZ points through shared memory; result valid in thread 0.
template <class F, int THREADS>
__device__ __forceinline__ void block_sum(XYZZ<F>& v, XYZZ<F>* sm) {
  using E = EC<F>;
  store16(sm + threadIdx.x, v);
  __syncthreads();
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      XYZZ<F> o;
      load16_rw(o, sm + threadIdx.x + s);
      E::add(v, o);
      store16(sm + threadIdx.x, v);
    }
    __syncthreads();
  }
}

constexpr int kReduceThreads = 64;

// one block per oversized bucket: bucket += sum of its task partials
template <class F>
__global__ void __launch_bounds__(kReduceThreads)
k_msm_ovf_merge(const OvfBucket* __restrict__ obuckets, const OvfCounters* __restrict__ ctr,
                const XYZZ<F>* __restrict__ partial, XYZZ<F>* __restrict__ buckets) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  using E = EC<F>;
  uint32_t nbk = ctr->nbuckets;
  for (uint32_t ob = blockIdx.x; ob < nbk; ob += gridDim.x) {
    OvfBucket b = obuckets[ob];
    XYZZ<F> acc;
    E::set_inf(acc);
    for (uint32_t t = threadIdx.x; t < b.ntasks; t += kReduceThreads) {
      XYZZ<F> p;
      load16_rw(p, partial + b.first + t);
      E::add(acc, p);
    }
    block_sum<F, kReduceThreads>(acc, sm);
    if (threadIdx.x == 0) {
      XYZZ<F> cur;
      load16_rw(cur, buckets + b.bucket);
      E::add(cur, acc);
      store16(buckets + b.bucket, cur);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ bucket reduction
// thread (w, g): sum_{k in group g} (k+1) * B[w][k]  ->  groups[w * ngroups + g]
template <class F>
__global__ void __launch_bounds__(64)
k_msm_bucket_reduce(const XYZZ<F>* __restrict__ buckets, MsmPlan pl, XYZZ<F>* __restrict__ groups) {
  using E = EC<F>;
  uint32_t ngroups = pl.nb / pl.group;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint32_t)pl.bwin * ngroups) return;
  uint32_t w = t / ngroups, g = t % ngroups;
  const XYZZ<F>* B = buckets + (uint64_t)w * pl.nb + (uint64_t)g * pl.group;
  XYZZ<F> run, acc;
  E::set_inf(run);
  E::set_inf(acc);
  for (int k = (int)pl.group - 1; k >= 0; k--) {
    XYZZ<F> b;
    load16_rw(b, B + k);
    E::add(run, b);
    E::add(acc, run);
  }
  // bucket k of this group has weight g*group + k + 1 ; acc holds sum (k+1) B_k, run holds sum B_k
  uint32_t base = g * pl.group;
  if (base) {
    XYZZ<F> m;
    E::mul_u32(m, run, base);
    E::add(acc, m);
  }
  store16(groups + t, acc);
}

// block b: out[b] = sum of in[b * per_slice .. (b + 1) * per_slice)   (window sums, in one or two levels)
template <class F>
__global__ void __launch_bounds__(kReduceThreads)
k_msm_slice_sum(const XYZZ<F>* __restrict__ in, uint32_t per_slice, XYZZ<F>* __restrict__ out) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  using E = EC<F>;
  const XYZZ<F>* G = in + (uint64_t)blockIdx.x * per_slice;
  XYZZ<F> acc;
  E::set_inf(acc);
  for (uint32_t g = threadIdx.x; g < per_slice; g += kReduceThreads) {
    XYZZ<F> p;
    load16_rw(p, G + g);
    E::add(acc, p);
  }
  block_sum<F, kReduceThreads>(acc, sm);
  if (threadIdx.x == 0) store16(out + blockIdx.x, acc);
}

// result = sum_w 2^(c w) windows[w]   (single thread; latency hidden by the other MSM streams)
template <class F>
__global__ void k_msm_horner(const XYZZ<F>* __restrict__ windows, MsmPlan pl, XYZZ<F>* __restrict__ out) {
  using E = EC<F>;
  if (threadIdx.x || blockIdx.x) return;
  XYZZ<F> acc;
  E::set_inf(acc);
  for (int w = pl.bwin - 1; w >= 0; w--) {
    for (int i = 0; i < pl.c; i++) E::dbl(acc);
    XYZZ<F> ws;
    load16_rw(ws, windows + w);
    E::add(acc, ws);
  }
  store16(out, acc);
}

}  // namespace b200
