// Signed-digit Pippenger multi-scalar multiplication for sm_100a.
//
//   stage 1 (sort)    digits/histogram -> multi-block exclusive scan -> counting-sort scatter of (table index | sign)
//                     entries.  Table mode: one pass feeds up to 4 base sets that share the scalar vector (the A / B / K
//                     keys of a Groth16 proof), or - batched mode - many scalar vectors over one base set.
//   stage 2 (reduce)  [optional affine pre-reduction, msm_pre.cuh] -> size-sorted bucket schedule
//                     -> bucket accumulation (one thread per bucket, 128-bit gathered affine loads, XYZZ mixed adds,
//                        lock-step iterations; oversized buckets are cut into 32-point tasks)
//                     -> tail on a high-priority stream: overflow partial sums (small buckets: one thread; big ones:
//                        two-level block tree), bucket reduction (running sums over groups of buckets, [base]*sum
//                        fix-up per group), per-array block reductions, Horner (windowed mode only).
//
// No host synchronisation anywhere; scalars are consumed in gnark-crypto's Montgomery form (converted in-register),
// points in gnark's affine layout.
//
// Replaces `G1Affine.MultiExp` / `G2Affine.MultiExp` (gnark-crypto, go.mod:16) as called for the
// Ar / Bs1 / Bs / Krs / Krs2 and Pedersen commitments of groth16.Prove
// (/root/reference/prover/prover_cpu.go:37).
#pragma once
#include "ec.cuh"
#include "msm_plan.h"
#include "msm_pre.cuh"

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

// Signed-digit recoding of one canonical scalar, window by window (carry kept between calls).
// next() returns the 0-based bucket index of window w in `b` (digit magnitude - 1) and its sign;
// false when the digit is zero.
template <int NS>
struct DigitWalker {
  uint32_t carry = 0;
  __device__ __forceinline__ bool next(const uint32_t* s, int w, int c, uint32_t& b, bool& neg) {
    uint32_t d = window_bits<NS>(s, w * c, c) + carry;
    carry = 0;
    neg = false;
    if (d > (1u << (c - 1))) {
      d = (1u << c) - d;
      neg = true;
      carry = 1;
    }
    b = d - 1;
    return d != 0;
  }
};

// The base sets one sorting pass feeds: scalar i multiplies point map[j][i] of set j (0xffffffff = not in
// the set; a null map is the identity).  Windowed mode uses set 0 only.
struct MsmSets {
  const uint32_t* map[kMaxSets];
  uint64_t npts[kMaxSets];
};

constexpr uint32_t kSkip = 0xffffffffu;

// Shared front end of k_msm_hist / k_msm_scatter: loads scalar i, resolves its point index in every set and
// converts it out of Montgomery form.  Returns false when the scalar contributes nothing.
template <class Fr>
__device__ __forceinline__ bool msm_load_scalar(const typename Fr::El* __restrict__ scalars, const MsmPlan& pl,
                                                const MsmSets& sets, uint64_t i, typename Fr::El& s,
                                                uint32_t (&pidx)[kMaxSets]) {
  const int nsets = pl.table ? pl.bwin / pl.tstride : 1;
  bool any = false;
#pragma unroll
  for (int j = 0; j < kMaxSets; j++) {
    pidx[j] = kSkip;
    if (j < nsets && i < pl.n) {
      pidx[j] = sets.map[j] ? sets.map[j][i] : (uint32_t)i;
      any |= pidx[j] != kSkip;
    }
  }
  if (!any) return false;
  typename Fr::El mont;
  load16(mont, scalars + i);
  Fr::from_mont(s, mont);
  return !Fr::is_zero(s);
}

// Window 0 is where solved witnesses collide (a fifth of the wires hold the value 1, many more hold small
// integers): lanes of a warp that hit the same window-0 bucket are counted with ONE atomic (match_any +
// popc) instead of serialising on one L2 address.  All 32 lanes must call these.
template <class Fr>
__global__ void __launch_bounds__(256)
k_msm_hist(const typename Fr::El* __restrict__ scalars, MsmPlan pl, uint32_t* __restrict__ hist, MsmSets sets) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nsets = pl.table ? pl.bwin / pl.tstride : 1;
  const int ts = pl.tstride;
  const unsigned lane = threadIdx.x & 31u;
  typename Fr::El s;
  uint32_t pidx[kMaxSets];
  bool live = msm_load_scalar<Fr>(scalars, pl, sets, i, s, pidx);
  DigitWalker<Fr::N> dw;
  uint32_t b = 0;
  bool neg = false;
  bool has0 = live && dw.next(s.v, 0, pl.c, b, neg);
  const unsigned same = __match_any_sync(0xffffffffu, has0 ? b : (0x80000000u | lane));
#pragma unroll
  for (int j = 0; j < kMaxSets; j++) {
    if (j >= nsets) break;
    const bool in = has0 && pidx[j] != kSkip;
    const unsigned grp = same & __ballot_sync(0xffffffffu, in);
    if (in && lane == (unsigned)(__ffs(grp) - 1)) atomicAdd(&hist[(uint64_t)(j * ts) * pl.nb + b], (uint32_t)__popc(grp));
  }
  if (!live) return;
  int wr = 0;   // w % tstride
  for (int w = 1; w < pl.nwin; w++) {
    if (++wr == ts) wr = 0;
    if (!dw.next(s.v, w, pl.c, b, neg)) continue;
    if (pl.table) {
#pragma unroll
      for (int j = 0; j < kMaxSets; j++)
        if (j < nsets && pidx[j] != kSkip) atomicAdd(&hist[(uint64_t)(j * ts + wr) * pl.nb + b], 1u);
    } else {
      atomicAdd(&hist[(uint64_t)w * pl.nb + b], 1u);
    }
  }
}

// Exclusive scan of every bucket array's histogram -> off (segment starts) and cur (running cursors), plus
// totals[a] = entries of bucket array a.  Two kernels over chunks of kScanChunk buckets: chunk sums, then
// the in-chunk scan seeded with the sum of the earlier chunks (<= 2^10 chunks per array).
constexpr int kScanItems = 16;
constexpr int kScanThreads = 1024;
constexpr uint32_t kScanChunk = kScanItems * kScanThreads;

static __global__ void __launch_bounds__(kScanThreads)
k_msm_scan_sums(const uint32_t* __restrict__ hist, MsmPlan pl, uint32_t* __restrict__ chunk_sums) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t* h = hist + (uint64_t)blockIdx.y * pl.nb;
  const uint32_t first = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  uint32_t sum = 0;
  const uint32_t pad = (1u << pl.pre) - 1u;   // segments are padded to multiples of 2^pre (msm_pre.cuh)
#pragma unroll
  for (int k = 0; k < kScanItems; k++) sum += first + k < pl.nb ? ((h[first + k] + pad) & ~pad) : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v = warp_sums[threadIdx.x];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (threadIdx.x == 0) chunk_sums[blockIdx.y * gridDim.x + blockIdx.x] = v;
  }
}

static __global__ void __launch_bounds__(kScanThreads)
k_msm_scan(const uint32_t* __restrict__ hist, MsmPlan pl, const uint32_t* __restrict__ chunk_sums,
           uint32_t* __restrict__ off, uint32_t* __restrict__ cur, uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t* h = hist + (uint64_t)blockIdx.y * pl.nb;
  uint32_t* o = off + (uint64_t)blockIdx.y * pl.nb;
  uint32_t* c = cur + (uint64_t)blockIdx.y * pl.nb;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // sum of the chunks before this one (and, in the last block, the array total)
  if (wid == 0) {
    const uint32_t* cs = chunk_sums + blockIdx.y * gridDim.x;
    uint32_t before = 0, all = 0;
    for (uint32_t k = lane; k < gridDim.x; k += 32) {
      uint32_t v = cs[k];
      all += v;
      if (k < blockIdx.x) before += v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, d);
      all += __shfl_xor_sync(0xffffffffu, all, d);
    }
    if (lane == 0) {
      carry_s = before;
      if (blockIdx.x == 0) totals[blockIdx.y] = all;
    }
  }
  __syncthreads();
  const uint32_t first = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
  const uint32_t pad = (1u << pl.pre) - 1u;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    v[k] = first + k < pl.nb ? ((h[first + k] + pad) & ~pad) : 0;
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
    uint32_t ws = warp_sums[lane];
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
}

template <class Fr>
__global__ void __launch_bounds__(256)
k_msm_scatter(const typename Fr::El* __restrict__ scalars, MsmPlan pl, uint32_t* __restrict__ cur,
              uint32_t* __restrict__ sorted, MsmSets sets) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nsets = pl.table ? pl.bwin / pl.tstride : 1;
  const int ts = pl.tstride;
  const unsigned lane = threadIdx.x & 31u;
  typename Fr::El s;
  uint32_t pidx[kMaxSets];
  bool live = msm_load_scalar<Fr>(scalars, pl, sets, i, s, pidx);
  DigitWalker<Fr::N> dw;
  uint32_t b = 0;
  bool neg = false;
  bool has0 = live && dw.next(s.v, 0, pl.c, b, neg);
  const unsigned same = __match_any_sync(0xffffffffu, has0 ? b : (0x80000000u | lane));
#pragma unroll
  for (int j = 0; j < kMaxSets; j++) {
    if (j >= nsets) break;
    const bool in = has0 && pidx[j] != kSkip;
    const unsigned grp = same & __ballot_sync(0xffffffffu, in);
    if (in) {
      const int leader = __ffs(grp) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(&cur[(uint64_t)(j * ts) * pl.nb + b], (uint32_t)__popc(grp));
      base = __shfl_sync(grp, base, leader);
      const uint32_t pos = base + __popc(grp & ((1u << lane) - 1u));
      // table mode: window 0 reads table 0 (the points themselves) at index pidx
      sorted[(uint64_t)(j * ts) * pl.stride + pos] = pidx[j] | (neg ? 0x80000000u : 0u);
    }
  }
  if (!live) return;
  int wr = 0, wq = 0;   // w % tstride, w / tstride
  for (int w = 1; w < pl.nwin; w++) {
    if (++wr == ts) {
      wr = 0;
      wq++;
    }
    if (!dw.next(s.v, w, pl.c, b, neg)) continue;
    const uint32_t sign = neg ? 0x80000000u : 0u;
    if (pl.table) {
#pragma unroll
      for (int j = 0; j < kMaxSets; j++) {
        if (j < nsets && pidx[j] != kSkip) {
          uint32_t pos = atomicAdd(&cur[(uint64_t)(j * ts + wr) * pl.nb + b], 1u);
          // digit window w = wq s + wr reads the precomputed multiple 2^(c s wq) P, stored at wq * npts + index
          sorted[(uint64_t)(j * ts + wr) * pl.stride + pos] = (uint32_t)((uint64_t)wq * sets.npts[j] + pidx[j]) | sign;
        }
      }
    } else {
      uint32_t pos = atomicAdd(&cur[(uint64_t)w * pl.nb + b], 1u);
      sorted[(uint64_t)w * pl.stride + pos] = pidx[0] | sign;
    }
  }
}

// Batched mode (pl.batch_n > 0): `bwin` scalar vectors over one shared base set, e.g. the 128 quotient polynomials
// of the EIP-7594 cell proofs against the monomial SRS.  Vector v feeds bucket array v; plain atomics (the scalars
// of this mode are quotient coefficients, uniformly distributed).
template <class Fr>
__global__ void __launch_bounds__(256)
k_msm_hist_batch(const typename Fr::El* __restrict__ scalars, MsmPlan pl, uint32_t* __restrict__ hist, MsmSets sets) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pl.n) return;
  const uint32_t v = (uint32_t)(i / pl.batch_n), li = (uint32_t)(i % pl.batch_n);
  if ((sets.map[0] ? sets.map[0][li] : li) == kSkip) return;
  typename Fr::El s, mont;
  load16(mont, scalars + i);
  Fr::from_mont(s, mont);
  if (Fr::is_zero(s)) return;
  DigitWalker<Fr::N> dw;
  uint32_t b;
  bool neg;
  for (int w = 0; w < pl.nwin; w++)
    if (dw.next(s.v, w, pl.c, b, neg)) atomicAdd(&hist[(uint64_t)v * pl.nb + b], 1u);
}

template <class Fr>
__global__ void __launch_bounds__(256)
k_msm_scatter_batch(const typename Fr::El* __restrict__ scalars, MsmPlan pl, uint32_t* __restrict__ cur,
                    uint32_t* __restrict__ sorted, MsmSets sets) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pl.n) return;
  const uint32_t v = (uint32_t)(i / pl.batch_n), li = (uint32_t)(i % pl.batch_n);
  const uint32_t pidx = sets.map[0] ? sets.map[0][li] : li;
  if (pidx == kSkip) return;
  typename Fr::El s, mont;
  load16(mont, scalars + i);
  Fr::from_mont(s, mont);
  if (Fr::is_zero(s)) return;
  DigitWalker<Fr::N> dw;
  uint32_t b;
  bool neg;
  for (int w = 0; w < pl.nwin; w++) {
    if (!dw.next(s.v, w, pl.c, b, neg)) continue;
    uint32_t pos = atomicAdd(&cur[(uint64_t)v * pl.nb + b], 1u);
    sorted[(uint64_t)v * pl.stride + pos] = (uint32_t)((uint64_t)w * sets.npts[0] + pidx) | (neg ? 0x80000000u : 0u);
  }
}

// ------------------------------------------------------------------------------------ accumulate
struct OvfTask {
  uint32_t bucket;   // global bucket id  a * nb + b   (a = bucket array)
  uint32_t start;    // offset inside the array's sorted segment
  uint32_t len;
  uint32_t pad;
};
// an oversized bucket: its points beyond the accumulate thread's cap are cut into `ntasks` tasks
// (tasks / partial sums [first, first + ntasks)); buckets with many tasks are merged in two tree levels
// (level-1 block sums at mid[first2 ...])
struct OvfBucket {
  uint32_t bucket;
  uint32_t first;
  uint32_t ntasks;
  uint32_t start;    // first overflow entry inside the array's sorted segment
  uint32_t len;      // overflow entries
  uint32_t first2;
  uint32_t pad0, pad1;
};
struct OvfCounters {
  uint32_t ntasks;
  uint32_t nbuckets;
  uint32_t nmid;
  uint32_t pad;
};

// device pointers of the point arrays a launch reads: one per base set in table mode, [0] otherwise
struct MsmPts {
  const void* p[kMaxSets];
};

// DIRECT: the run is `len` affine points stored contiguously at `points` (output of the affine pre-reduction);
// otherwise `idx` holds (table index | sign) entries into `points`.
template <class F, bool DIRECT>
__device__ __forceinline__ void accumulate_fetch(Affine<F>& dst, uint32_t& e, const Affine<F>* __restrict__ points,
                                                 const uint32_t* __restrict__ idx, uint32_t k) {
  if (DIRECT) {
    e = 0;
    load16(dst, points + k);
  } else {
    e = idx[k];
    load16(dst, points + (e & 0x7fffffffu));
  }
}

template <class F, bool DIRECT>
__device__ __forceinline__ void accumulate_run(XYZZ<F>& acc, const Affine<F>* __restrict__ points,
                                               const uint32_t* __restrict__ idx, uint32_t len) {
  using E = EC<F>;
  E::set_inf(acc);
  if (len == 0) return;
  // software pipeline: the next point is in flight while the current mixed add runs
  uint32_t e;
  Affine<F> nxt;
  accumulate_fetch<F, DIRECT>(nxt, e, points, idx, 0);
  for (uint32_t k = 0; k < len; k++) {
    Affine<F> cur = nxt;
    bool neg = e >> 31;
    if (k + 1 < len) accumulate_fetch<F, DIRECT>(nxt, e, points, idx, k + 1);
    if (neg) E::neg(cur);
    E::madd(acc, cur);
  }
}

// Lock-step variant: every warp of the block runs the same trip count and meets at a barrier each
// iteration, so the (64 KB, larger than the instruction cache) loop body is streamed once per SM and
// iteration instead of once per warp.  Buckets of a block have (almost) equal sizes thanks to the
// size-sorted schedule, so the padding iterations are few.
template <class F>
struct AccCfg;
template <class F, bool DIRECT>
__device__ __forceinline__ void accumulate_run_lockstep(XYZZ<F>& acc, const Affine<F>* __restrict__ points,
                                                        const uint32_t* __restrict__ idx, uint32_t len) {
  using E = EC<F>;
  constexpr bool kPrefetch = AccCfg<F>::kPrefetch;
  __shared__ uint32_t s_trips;
  E::set_inf(acc);
  if (threadIdx.x == 0) s_trips = 0;
  __syncthreads();
  if (len) atomicMax(&s_trips, len);
  __syncthreads();
  const uint32_t trips = s_trips;
  uint32_t e = 0;
  Affine<F> nxt;
  if (kPrefetch && len) accumulate_fetch<F, DIRECT>(nxt, e, points, idx, 0);
  for (uint32_t k = 0; k < trips; k++) {
    __syncthreads();
    if (k < len) {
      if (!kPrefetch) accumulate_fetch<F, DIRECT>(nxt, e, points, idx, k);
      Affine<F> cur = nxt;
      bool neg = e >> 31;
      if (kPrefetch && k + 1 < len) accumulate_fetch<F, DIRECT>(nxt, e, points, idx, k + 1);
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

#ifndef B200_ACC_NO_LOCKSTEP
#define B200_ACC_LOCKSTEP 1   // measured: -5% (G1) / -8% (G2) accumulate time at 128 threads, 3 blocks per SM
#endif
#ifndef B200_ACC_MIN_BLOCKS
#define B200_ACC_MIN_BLOCKS 3
#endif
#ifndef B200_ACC_MIN_BLOCKS_BIG
#define B200_ACC_MIN_BLOCKS_BIG 3   // Fp2 over a 12-limb prime field (G2 of BLS12-377 / BLS12-381): measured better than 2
#endif
#ifndef B200_ACC_MIN_BLOCKS_WIDE
#define B200_ACC_MIN_BLOCKS_WIDE 2  // 24-limb prime field (BW6-761): one 1176-MAC product needs ~100 registers of its own;
                                    // 255 registers at 2 blocks per SM: 47.8 -> 38.5 ms for a 2^20 MSM
#endif
#ifndef B200_ACC_MIN_BLOCKS_SMALL
#define B200_ACC_MIN_BLOCKS_SMALL 4   // 32-byte coordinate fields (BN254 Fp): 126 registers, measured -5% vs 3 blocks
#endif
#ifndef B200_ACC_THREADS
#define B200_ACC_THREADS 128
#endif
template <class F>
struct AccCfg {
  static constexpr int kMinBlocks = F::BASE_N >= 24                 ? B200_ACC_MIN_BLOCKS_WIDE
                                    : sizeof(typename F::El) >= 96  ? B200_ACC_MIN_BLOCKS_BIG
                                    : sizeof(typename F::El) <= 32  ? B200_ACC_MIN_BLOCKS_SMALL
                                                                    : B200_ACC_MIN_BLOCKS;
  // one-point prefetch: not for Fp2 points (48 more registers in a kernel that already spills; measured +1.5% without)
  static constexpr bool kPrefetch = !(sizeof(typename F::El) >= 96 && F::BASE_N < 24);
};
template <class F, bool DIRECT>
__global__ void __launch_bounds__(B200_ACC_THREADS, AccCfg<F>::kMinBlocks)
k_msm_accumulate(MsmPts pts, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ off,
                 const uint32_t* __restrict__ end, const uint32_t* __restrict__ perm,
                 const uint32_t* __restrict__ totals, MsmPlan pl, XYZZ<F>* __restrict__ buckets,
                 OvfBucket* __restrict__ obuckets, OvfCounters* __restrict__ ctr) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < (uint64_t)pl.bwin * pl.nb;
#ifndef B200_ACC_LOCKSTEP
  if (!valid) return;
#endif
  const uint32_t gb = valid ? perm[t] : 0;
  uint32_t w = gb / pl.nb;
  uint32_t start = valid ? off[gb] : 0, cnt = valid ? end[gb] - start : 0;
  uint32_t mine = cnt;
  // the cap follows the ACTUAL mean bucket load of this array (sparse witness vectors fill far fewer
  // digits than n * nwin): a single thread's chain of additions is pure latency (~10 us each)
  uint32_t task = 8u * (totals[w] / pl.nb + 1u);
  task = task < pl.task_min ? pl.task_min : (task > pl.task ? pl.task : task);
  if (cnt > task) {
    uint32_t extra = (cnt - task + pl.ovf_task - 1) / pl.ovf_task;
    uint32_t first = atomicAdd(&ctr->ntasks, extra);
    if (first + extra <= pl.max_ovf) {
      mine = task;
      uint32_t ob = atomicAdd(&ctr->nbuckets, 1u);
      obuckets[ob] = OvfBucket{gb, first, extra, start + task, cnt - task, 0, 0, 0};
    }
    // else: cannot happen (capacity covers every entry); stay correct anyway by taking the whole bucket
  }
  // DIRECT: pts.p[0] is the pre-reduced affine list (single bucket array), `start` indexes it
  const Affine<F>* points = reinterpret_cast<const Affine<F>*>(pts.p[(pl.table && !pl.batch_n) ? w / pl.tstride : 0]) + (DIRECT ? start : 0);
  const uint32_t* idx = DIRECT ? nullptr : sorted + (uint64_t)w * pl.stride + start;
  XYZZ<F> acc;
#ifdef B200_ACC_LOCKSTEP
  accumulate_run_lockstep<F, DIRECT>(acc, points, idx, mine);
  if (valid) store16(buckets + gb, acc);
#else
  accumulate_run<F, DIRECT>(acc, points, idx, mine);
  store16(buckets + gb, acc);
#endif
}

// ---- oversized buckets (skewed scalar distributions, e.g. the many 1-valued witness wires)
// one warp per oversized bucket writes its task list
static __global__ void __launch_bounds__(256)
k_msm_ovf_expand(const OvfBucket* __restrict__ obuckets, const OvfCounters* __restrict__ ctr, MsmPlan pl,
                 OvfTask* __restrict__ tasks) {
  const uint32_t nbk = ctr->nbuckets;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t ob = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ob < nbk; ob += warps) {
    OvfBucket b = obuckets[ob];
    for (uint32_t k = lane; k < b.ntasks; k += 32) {
      uint32_t s = k * pl.ovf_task;
      uint32_t l = b.len - s < pl.ovf_task ? b.len - s : pl.ovf_task;
      tasks[b.first + k] = OvfTask{b.bucket, b.start + s, l, 0};
    }
  }
}

template <class F, bool DIRECT>
__global__ void __launch_bounds__(128)
k_msm_ovf_accumulate(MsmPts pts, const uint32_t* __restrict__ sorted, MsmPlan pl,
                     const OvfTask* __restrict__ tasks, const OvfCounters* __restrict__ ctr,
                     XYZZ<F>* __restrict__ partial) {
  uint32_t nt = ctr->ntasks < pl.max_ovf ? ctr->ntasks : pl.max_ovf;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    OvfTask tk = tasks[t];
    uint32_t w = tk.bucket / pl.nb;
    const Affine<F>* points = reinterpret_cast<const Affine<F>*>(pts.p[(pl.table && !pl.batch_n) ? w / pl.tstride : 0]) + (DIRECT ? tk.start : 0);
    XYZZ<F> acc;
    accumulate_run<F, DIRECT>(acc, points, DIRECT ? nullptr : sorted + (uint64_t)w * pl.stride + tk.start, tk.len);
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
// buckets with <= 32 partial sums: one thread adds them (0.4 ms).  Table mode makes the buckets below
// (r >> c (nwin-1)) systematically ~9x heavier than the mean - the top digit window only reaches that far - so a
// dense MSM has thousands of buckets with 5-10 partial sums each; they must not take the block-tree path.
constexpr uint32_t kOvfSmall = 32;
constexpr int kOvfMergeThreads = 128;
constexpr uint32_t kOvfPre = 4;                         // partial sums a thread adds before the block tree
constexpr uint32_t kOvfChunk = kOvfMergeThreads * kOvfPre;   // partial sums one level-1 block merges

// thread per oversized bucket: few partial sums are added directly; many get a level-1 output range
template <class F>
__global__ void __launch_bounds__(64)
k_msm_ovf_merge_small(OvfBucket* __restrict__ obuckets, OvfCounters* __restrict__ ctr,
                      const XYZZ<F>* __restrict__ partial, XYZZ<F>* __restrict__ buckets) {
  using E = EC<F>;
  const uint32_t nbk = ctr->nbuckets;
  for (uint32_t ob = blockIdx.x * blockDim.x + threadIdx.x; ob < nbk; ob += gridDim.x * blockDim.x) {
    OvfBucket b = obuckets[ob];
    if (b.ntasks > kOvfSmall) {
      obuckets[ob].first2 = atomicAdd(&ctr->nmid, (b.ntasks + kOvfChunk - 1) / kOvfChunk);
      continue;
    }
    XYZZ<F> acc;
    load16_rw(acc, buckets + b.bucket);
    for (uint32_t t = 0; t < b.ntasks; t++) {
      XYZZ<F> p;
      load16_rw(p, partial + b.first + t);
      E::add(acc, p);
    }
    store16(buckets + b.bucket, acc);
  }
}

// level 1: block (x, y) sums chunk x (+ k gridDim.x) of the partial sums of big bucket y (+ k gridDim.y)
template <class F>
__global__ void __launch_bounds__(kOvfMergeThreads)
k_msm_ovf_merge_l1(const OvfBucket* __restrict__ obuckets, const OvfCounters* __restrict__ ctr,
                   const XYZZ<F>* __restrict__ partial, XYZZ<F>* __restrict__ mid) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  using E = EC<F>;
  const uint32_t nbk = ctr->nbuckets;
  for (uint32_t ob = blockIdx.y; ob < nbk; ob += gridDim.y) {
    OvfBucket b = obuckets[ob];
    if (b.ntasks <= kOvfSmall) continue;
    const uint32_t nchunks = (b.ntasks + kOvfChunk - 1) / kOvfChunk;
    for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      XYZZ<F> acc;
      E::set_inf(acc);
      const uint32_t base = ch * kOvfChunk + threadIdx.x * kOvfPre;
      for (uint32_t k = 0; k < kOvfPre; k++) {
        if (base + k < b.ntasks) {
          XYZZ<F> p;
          load16_rw(p, partial + b.first + base + k);
          E::add(acc, p);
        }
      }
      block_sum<F, kOvfMergeThreads>(acc, sm);
      if (threadIdx.x == 0) store16(mid + b.first2 + ch, acc);
      __syncthreads();
    }
  }
}

// level 2: one block per big bucket adds its level-1 sums into the bucket
template <class F>
__global__ void __launch_bounds__(kOvfMergeThreads)
k_msm_ovf_merge_l2(const OvfBucket* __restrict__ obuckets, const OvfCounters* __restrict__ ctr,
                   const XYZZ<F>* __restrict__ mid, XYZZ<F>* __restrict__ buckets) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  using E = EC<F>;
  const uint32_t nbk = ctr->nbuckets;
  for (uint32_t ob = blockIdx.x; ob < nbk; ob += gridDim.x) {
    OvfBucket b = obuckets[ob];
    if (b.ntasks <= kOvfSmall) continue;
    const uint32_t nchunks = (b.ntasks + kOvfChunk - 1) / kOvfChunk;
    XYZZ<F> acc;
    E::set_inf(acc);
    for (uint32_t ch = threadIdx.x; ch < nchunks; ch += kOvfMergeThreads) {
      XYZZ<F> p;
      load16_rw(p, mid + b.first2 + ch);
      E::add(acc, p);
    }
    block_sum<F, kOvfMergeThreads>(acc, sm);
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
// sum_k (k + 1) B_k with two full additions per bucket and a log-depth tail.
//
// Level 0 cuts the nb buckets of an array into groups of G0; thread (w, g) walks its group from the top with the
// classic running sum and emits  acc_g = sum_j (j + 1) B[g G0 + j]  and  run_g = sum_j B[g G0 + j].  Then
//     sum_k (k + 1) B_k = sum_g acc_g + G0 * sum_g g run_g .
// The weighted sum over the (<= 2^14) run values is taken WITHOUT another serial pass: run_g is stored one slot to
// the left (slot g - 1), a Hillis-Steele suffix scan (log2(ng) rounds of one addition per element, k_msm_suffix_round)
// turns slot i into sum_{g > i} run_g, and the plain sum of all slots is sum_g g run_g.  Both plain sums run through the
// block-tree reduction (k_msm_range_sum).  Dependent-addition depth: 2 G0 (level 0) + log2(ng) + ~15, instead of the
// 2 G0 + 32-step double-and-add per group of round 1 - and the multiplier work drops by a third.
template <class F>
__global__ void __launch_bounds__(64)
k_msm_wsum_level0(const XYZZ<F>* __restrict__ buckets, uint32_t nb, uint32_t G, uint32_t narrays, uint32_t ngroups,
                  XYZZ<F>* __restrict__ acc_out, XYZZ<F>* __restrict__ run_shifted) {
  using E = EC<F>;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= narrays * ngroups) return;
  uint32_t w = t / ngroups, g = t % ngroups;
  const XYZZ<F>* X = buckets + (uint64_t)w * nb + (uint64_t)g * G;
  XYZZ<F> run, acc;
  E::set_inf(run);
  E::set_inf(acc);
  for (int j = (int)G - 1; j >= 0; j--) {
    XYZZ<F> b;
    load16_rw(b, X + j);
    E::add(run, b);
    E::add(acc, run);
  }
  store16(acc_out + t, acc);
  if (g) store16(run_shifted + t - 1, run);
  if (g == ngroups - 1) {
    E::set_inf(run);
    store16(run_shifted + t, run);   // the last slot of every array: nothing lies to its right
  }
}

// one round of the suffix scan: out[i] = in[i] + in[i + d] (inside each array of n elements)
template <class F>
__global__ void __launch_bounds__(128)
k_msm_suffix_round(const XYZZ<F>* __restrict__ in, XYZZ<F>* __restrict__ out, uint32_t n, uint32_t d, uint32_t total) {
  using E = EC<F>;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  uint32_t i = t % n;
  XYZZ<F> a;
  load16_rw(a, in + t);
  if (i + d < n) {
    XYZZ<F> b;
    load16_rw(b, in + t + d);
    E::add(a, b);
  }
  store16(out + t, a);
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

// block (slice, a): out[a * gridDim.x + slice] = sum of per_slice consecutive elements of array a, where arrays
// 0 .. split-1 live at base0 + a * n and arrays split .. at base1 + (a - split) * n (elements beyond n are skipped)
template <class F>
__global__ void __launch_bounds__(kReduceThreads)
k_msm_range_sum(const XYZZ<F>* __restrict__ base0, const XYZZ<F>* __restrict__ base1, uint32_t split, uint32_t n,
                uint32_t per_slice, XYZZ<F>* __restrict__ out) {
  extern __shared__ uint4 smem_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(smem_raw);
  using E = EC<F>;
  const uint32_t a = blockIdx.y;
  const XYZZ<F>* A = a < split ? base0 + (uint64_t)a * n : base1 + (uint64_t)(a - split) * n;
  const uint32_t lo = blockIdx.x * per_slice;
  const uint32_t hi = lo + per_slice < n ? lo + per_slice : n;
  XYZZ<F> acc;
  E::set_inf(acc);
  for (uint32_t g = lo + threadIdx.x; g < hi; g += kReduceThreads) {
    XYZZ<F> p;
    load16_rw(p, A + g);
    E::add(acc, p);
  }
  block_sum<F, kReduceThreads>(acc, sm);
  if (threadIdx.x == 0) store16(out + (uint64_t)a * gridDim.x + blockIdx.x, acc);
}

// thread w: windows[w] = sums[w] + 2^logG0 * sums[narr + w]   (sum acc_g + G0 * sum g run_g)
template <class F>
__global__ void k_msm_wsum_combine(const XYZZ<F>* __restrict__ sums, uint32_t narr, uint32_t logG0,
                                   XYZZ<F>* __restrict__ windows) {
  using E = EC<F>;
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= narr) return;
  XYZZ<F> T, s0;
  load16_rw(T, sums + narr + w);
  for (uint32_t k = 0; k < logG0; k++) E::dbl(T);
  load16_rw(s0, sums + w);
  E::add(T, s0);
  store16(windows + w, T);
}

// windowed mode: out[0] = sum_w 2^(c w) windows[w]  (single thread; latency hidden by the other MSM streams)
// table mode: one result per base set; with table stride 1 it is windows[j] (the tables carry the 2^(c w) factors)
template <class F>
__global__ void k_msm_horner(const XYZZ<F>* __restrict__ windows, MsmPlan pl, XYZZ<F>* __restrict__ out) {
  using E = EC<F>;
  if (blockIdx.x) return;
  if (pl.table) {
    // one result per base set (or per scalar vector in batched mode): sum_r 2^(c r) S_r over its tstride bucket sets
    const int ts = pl.tstride;
    for (int j = threadIdx.x; j < pl.bwin / ts; j += blockDim.x) {
      XYZZ<F> acc;
      load16_rw(acc, windows + j * ts + ts - 1);
      for (int r = ts - 2; r >= 0; r--) {
        for (int i = 0; i < pl.c; i++) E::dbl(acc);
        XYZZ<F> ws;
        load16_rw(ws, windows + j * ts + r);
        E::add(acc, ws);
      }
      store16(out + j, acc);
    }
    return;
  }
  if (threadIdx.x) return;
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
