// Per-curve backend: instantiates the templated kernels for one curve configuration and exposes
// them through the CurveBackend interface.  Included by curve_<name>.cu only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "backend.h"
#include "msm.cuh"
#include "ntt.cuh"
#include "kzg.cuh"
#include "serde.cuh"
#include "pairing.cuh"
#include <memory>
#include <map>
#include <mutex>

namespace b200 {

// ------------------------------------------------------------------------------------ debug kernels
template <class F>
__global__ void k_dbg_field(int op, const typename F::El* a, const typename F::El* b, typename F::El* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename F::El x = a[i], y, r;
  if (b) y = b[i];
  switch (op) {
    case OP_ADD: F::add(r, x, y); break;
    case OP_SUB: F::sub(r, x, y); break;
    case OP_MUL: F::mul(r, x, y); break;
    case OP_SQR: F::sqr(r, x); break;
    case OP_INV: F::inv(r, x); break;
    case OP_NEG: F::neg(r, x); break;
    case OP_SQRT:
      if (!F::sqrt(r, x)) F::set_zero(r);
      break;
    case OP_INV_BIN: F::inv_bin(r, x); break;
    default: r = x; break;
  }
  out[i] = r;
}

template <class F>
__global__ void k_dbg_field_mont(int op, const typename F::El* a, typename F::El* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename F::El x = a[i], r;
  if (op == OP_FROM_MONT) F::from_mont(r, x);
  else F::to_mont(r, x);
  out[i] = r;
}

// one kernel per operation (a single kernel switching over the operations shares one stack frame between unrelated
// cases; the scalar multiplication of BN254 - where Fr and Fp elements are the same C++ type - came out wrong that way)
template <class F, class Fr, int OP>
__global__ void k_dbg_ec(const void* a, const void* b, void* out, uint64_t n) {
  using E = EC<F>;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> p = reinterpret_cast<const XYZZ<F>*>(a)[i];
  if constexpr (OP == EC_MADD) {
    Affine<F> q = reinterpret_cast<const Affine<F>*>(b)[i];
    E::madd(p, q);
    reinterpret_cast<XYZZ<F>*>(out)[i] = p;
  } else if constexpr (OP == EC_ADD) {
    XYZZ<F> q = reinterpret_cast<const XYZZ<F>*>(b)[i];
    E::add(p, q);
    reinterpret_cast<XYZZ<F>*>(out)[i] = p;
  } else if constexpr (OP == EC_DBL) {
    E::dbl(p);
    reinterpret_cast<XYZZ<F>*>(out)[i] = p;
  } else if constexpr (OP == EC_TO_AFFINE) {
    Affine<F> q;
    E::to_affine(q, p);
    reinterpret_cast<Affine<F>*>(out)[i] = q;
  } else {
    typename Fr::El s = reinterpret_cast<const typename Fr::El*>(b)[i], sc;
    Fr::from_mont(sc, s);
    uint32_t k[Fr::N];
#pragma unroll
    for (int j = 0; j < Fr::N; j++) k[j] = sc.v[j];
    XYZZ<F> r;
    E::template mul_scalar<Fr::N>(r, p, k);
    reinterpret_cast<XYZZ<F>*>(out)[i] = r;
  }
}

// out = affine(sum of `count` XYZZ partial sums): the combine step of a range-split MSM whose per-GPU
// partials were all-gathered over NVLink (count = number of GPUs, a few hundred bytes each)
template <class F>
__global__ void k_sum_partials(const XYZZ<F>* in, uint32_t count, Affine<F>* out) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ<F> acc;
  EC<F>::set_inf(acc);
  for (uint32_t i = 0; i < count; i++) {
    XYZZ<F> p = in[i];
    EC<F>::add(acc, p);
  }
  Affine<F> o;
  EC<F>::template to_affine<true>(o, acc);
  *out = o;
}

template <class F>
__global__ void k_to_affine(const XYZZ<F>* in, Affine<F>* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> p = in[i];
  Affine<F> q;
  EC<F>::to_affine(q, p);
  out[i] = q;
}

// out[i] = affine([k_i] base): one thread per scalar (double-and-add, leading zero bits skipped)
template <class F, class Fr>
__global__ void __launch_bounds__(64) k_fixed_base(const Affine<F>* base, const typename Fr::El* scalars, uint64_t n,
                                                    Affine<F>* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename Fr::El k;
  load16(k, scalars + i);
  Fr::from_mont(k, k);
  XYZZ<F> b, r;
  Affine<F> ba = *base;
  EC<F>::from_affine(b, ba);
  EC<F>::template mul_scalar<Fr::N>(r, b, k.v);
  Affine<F> o;
  EC<F>::to_affine(o, r);
  store16(out + i, o);
}

// dependent multiply chain: measures the sustained Montgomery-multiply (IMAD.WIDE) issue rate
template <class F>
__global__ void __launch_bounds__(256) k_calib_mul(typename F::El* io, uint64_t n, int iters) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename F::El x = io[i], y = io[(i + 1) % n];
#pragma unroll 1
  for (int k = 0; k < iters; k += 2) {
    F::mul(x, x, y);
    F::mul(y, y, x);
  }
  F::add(x, x, y);
  io[i] = x;
}

template <class Cfg>
__global__ void k_curve_consts(typename Cfg::G1F::El* b1, typename Cfg::G2F::El* b2) {
  if (threadIdx.x || blockIdx.x) return;
  typename Cfg::G1F::El x;
  typename Cfg::G2F::El y;
  Cfg::curve_b(x, y);
  *b1 = x;
  *b2 = y;
}

// ------------------------------------------------------------------------------------ proof assembly
// out4 = {r, s, 1, -r*s} in Montgomery form
template <class Fr>
__global__ void k_prep_rs(const typename Fr::El* r, const typename Fr::El* s, typename Fr::El* out4) {
  typename Fr::El t;
  out4[0] = *r;
  out4[1] = *s;
  Fr::set_one(out4[2]);
  Fr::mul(t, *r, *s);
  Fr::neg(out4[3], t);
}

// tmp[0] = [s] Ar ; tmp[1] = [r] Bs1      (one thread each)
template <class F, class Fr>
__global__ void k_assemble_mul(const XYZZ<F>* ar, const XYZZ<F>* bs1, const typename Fr::El* rs, XYZZ<F>* tmp) {
  typename Fr::El k;
  XYZZ<F> p, r;
  if (blockIdx.x == 0) {
    Fr::from_mont(k, rs[1]);
    p = *ar;
  } else {
    Fr::from_mont(k, rs[0]);
    p = *bs1;
  }
  EC<F>::template mul_scalar_w4<Fr::N>(r, p, k.v);
  tmp[blockIdx.x] = r;
}

// k += tmp[0] + tmp[1]: a range-split slice folds s*Ar_g + r*Bs1_g into its K partial sum before the gather
template <class F>
__global__ void k_assemble_fold(XYZZ<F>* k, const XYZZ<F>* tmp) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ<F> acc = *k;
  EC<F>::add(acc, tmp[0]);
  EC<F>::add(acc, tmp[1]);
  *k = acc;
}

// block 0: Krs = K + Z + s*Ar + r*Bs1 -> affine ; 1: Ar -> affine ; 2: Bs (G2) -> affine ; 3: Pok -> affine
template <class F1, class F2>
__global__ void k_assemble_out(const XYZZ<F1>* ar, const XYZZ<F2>* bs2, const XYZZ<F1>* k, const XYZZ<F1>* z,
                               const XYZZ<F1>* pok, const XYZZ<F1>* tmp, Affine<F1>* out_ar, Affine<F2>* out_bs,
                               Affine<F1>* out_krs, Affine<F1>* out_pok) {
  if (blockIdx.x == 0) {
    XYZZ<F1> acc = *k;
    EC<F1>::add(acc, *z);
    EC<F1>::add(acc, tmp[0]);
    EC<F1>::add(acc, tmp[1]);
    Affine<F1> o;
    EC<F1>::template to_affine<true>(o, acc);
    *out_krs = o;
  } else if (blockIdx.x == 1) {
    Affine<F1> o;
    EC<F1>::template to_affine<true>(o, *ar);
    *out_ar = o;
  } else if (blockIdx.x == 2) {
    Affine<F2> o;
    EC<F2>::template to_affine<true>(o, *bs2);
    *out_bs = o;
  } else if (pok && out_pok) {
    Affine<F1> o;
    EC<F1>::template to_affine<true>(o, *pok);
    *out_pok = o;
  }
}

// sums[j] = sum over parts of partials[part][j]; slots 0..4 are G1 XYZZ, slot 5 is G2 XYZZ (range-split prove)
template <class F1, class F2>
__global__ void k_sum_sets(const uint8_t* __restrict__ partials, uint32_t nparts, uint8_t* __restrict__ sums) {
  const size_t x1 = sizeof(XYZZ<F1>), x2 = sizeof(XYZZ<F2>), set = 5 * x1 + x2;
  if (threadIdx.x) return;
  const int j = blockIdx.x;
  if (j < 5) {
    XYZZ<F1> acc;
    EC<F1>::set_inf(acc);
    for (uint32_t p = 0; p < nparts; p++) {
      XYZZ<F1> v = *reinterpret_cast<const XYZZ<F1>*>(partials + p * set + j * x1);
      EC<F1>::add(acc, v);
    }
    *reinterpret_cast<XYZZ<F1>*>(sums + j * x1) = acc;
  } else {
    XYZZ<F2> acc;
    EC<F2>::set_inf(acc);
    for (uint32_t p = 0; p < nparts; p++) {
      XYZZ<F2> v = *reinterpret_cast<const XYZZ<F2>*>(partials + p * set + 5 * x1);
      EC<F2>::add(acc, v);
    }
    *reinterpret_cast<XYZZ<F2>*>(sums + 5 * x1) = acc;
  }
}

// ------------------------------------------------------------------------------------ MSM driver
// T_j[i] = 2^(c j) P_i for j < nwin (affine, table j at offset j * npts): one thread per point.
// The XYZZ doubling chain runs through all windows; the nwin - 1 affine normalisations of a point share ONE field
// inversion (Montgomery's trick over d_j = ZZ_j * ZZZ_j), which is 3x less multiplier work than an inversion per
// window.  `scratch` holds {ZZ_j, ZZZ_j, prefix product} per (window, point) while the kernel runs; X_j, Y_j wait
// in their table slot.
template <class F>
__global__ void __launch_bounds__(64) k_build_tables(const Affine<F>* __restrict__ base, uint64_t npts, int c, int nwin,
                                                      Affine<F>* __restrict__ tables,
                                                      typename F::El* __restrict__ scratch) {
  using El = typename F::El;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  Affine<F> p;
  load16(p, base + i);
  store16(tables + i, p);
  XYZZ<F> x;
  EC<F>::from_affine(x, p);
  El run;
  F::set_one(run);
  for (int j = 1; j < nwin; j++) {
    for (int k = 0; k < c; k++) EC<F>::dbl(x);
    El* sc = scratch + ((uint64_t)(j - 1) * npts + i) * 3;
    Affine<F> xy;
    xy.x = x.x;
    xy.y = x.y;
    store16(tables + (uint64_t)j * npts + i, xy);
    El d;
    F::mul(d, x.zz, x.zzz);
    if (EC<F>::is_inf(x)) F::set_one(d);   // infinity (or a point of 2-power order): keeps the product invertible
    store16(sc + 0, x.zz);
    store16(sc + 1, x.zzz);
    store16(sc + 2, run);                  // product of d_1 .. d_{j-1}
    F::mul(run, run, d);
  }
  if (nwin < 2) return;
  El rinv;
  F::inv(rinv, run);
  for (int j = nwin - 1; j >= 1; j--) {
    El* sc = scratch + ((uint64_t)(j - 1) * npts + i) * 3;
    El zz, zzz, pre, dinv, d;
    load16_rw(zz, sc + 0);
    load16_rw(zzz, sc + 1);
    load16_rw(pre, sc + 2);
    Affine<F> xy, o;
    load16_rw(xy, tables + (uint64_t)j * npts + i);
    {
      // (no branch around the products: for G2 they are out-of-line calls, see the control-flow rule in ec.cuh)
      F::mul(dinv, rinv, pre);             // 1 / (ZZ_j ZZZ_j)
      El izz, izzz;
      F::mul(izz, dinv, zzz);
      F::mul(izzz, dinv, zz);
      F::mul(o.x, xy.x, izz);
      F::mul(o.y, xy.y, izzz);
      F::mul(d, zz, zzz);
      if (F::is_zero(zz)) {
        F::set_zero(o.x);
        F::set_zero(o.y);
        F::set_one(d);
      }
    }
    F::mul(rinv, rinv, d);
    store16(tables + (uint64_t)j * npts + i, o);
  }
}

// ---- stage 1: digits -> per-bucket-array counting sort of (table index | sign) entries
template <class Fr>
void msm_sort_launch(const void* d_scalars, const MsmPlan& pl, const MsmSets& sets, MsmWorkspace& ws, cudaStream_t s,
                     MsmSorted& out) {
  const uint64_t total_b = (uint64_t)pl.bwin * pl.nb;
  uint32_t* hist = (uint32_t*)ws.hist.get(total_b * 4);
  uint32_t* off = (uint32_t*)ws.off.get(total_b * 4);
  uint32_t* cur = (uint32_t*)ws.cur.get(total_b * 4);
  uint32_t* sorted = (uint32_t*)ws.sorted.get((uint64_t)pl.bwin * pl.stride * 4);
  const unsigned nchunks = (unsigned)((pl.nb + kScanChunk - 1) / kScanChunk);
  if (nchunks > 1024) throw std::runtime_error("msm: too many buckets per array");
  uint32_t* chunk_sums = (uint32_t*)ws.chunk_sums.get(((uint64_t)pl.bwin * nchunks + pl.bwin) * 4);
  uint32_t* totals = chunk_sums + (uint64_t)pl.bwin * nchunks;
  const int tok_sort = prof_begin(PROF_MSM_SORT, s);
  B200_CUDA(cudaMemsetAsync(hist, 0, total_b * 4, s));
  // pre-reduction: the padding slots of every bucket segment read as the point at infinity
  if (pl.pre) B200_CUDA(cudaMemsetAsync(sorted, 0xff, (uint64_t)pl.bwin * pl.stride * 4, s));
  const auto* sc = reinterpret_cast<const typename Fr::El*>(d_scalars);
  const unsigned sblocks = (unsigned)((pl.n + 255) / 256);
  if (pl.batch_n) k_msm_hist_batch<Fr><<<sblocks, 256, 0, s>>>(sc, pl, hist, sets);
  else k_msm_hist<Fr><<<sblocks, 256, 0, s>>>(sc, pl, hist, sets);
  k_msm_scan_sums<<<dim3(nchunks, pl.bwin), kScanThreads, 0, s>>>(hist, pl, chunk_sums);
  k_msm_scan<<<dim3(nchunks, pl.bwin), kScanThreads, 0, s>>>(hist, pl, chunk_sums, off, cur, totals);
  if (pl.batch_n) k_msm_scatter_batch<Fr><<<sblocks, 256, 0, s>>>(sc, pl, cur, sorted, sets);
  else k_msm_scatter<Fr><<<sblocks, 256, 0, s>>>(sc, pl, cur, sorted, sets);
  prof_end(tok_sort, s);
  prof_count_launches(4);
  B200_CUDA(cudaGetLastError());
  out.pl = pl;
  out.off = off;
  out.end = cur;
  out.sorted = sorted;
  out.totals = totals;
}

// ---- stage 2: bucket accumulation + reduction of `pl.bwin` bucket arrays -> d_out (one XYZZ point in
// windowed mode, one per base set in table mode)
template <class F, int GROUP>
void msm_reduce_launch(const MsmSorted& so, const MsmPts& pts, void* d_out, MsmWorkspace& ws, cudaStream_t s,
                       bool join) {
  using Pt = XYZZ<F>;
  const MsmPlan& pl = so.pl;
  const uint64_t total_b = (uint64_t)pl.bwin * pl.nb;
  const uint32_t ngroups = pl.nb / pl.group;
  Pt* buckets = (Pt*)ws.buckets.get(total_b * sizeof(Pt));
  OvfTask* tasks = (OvfTask*)ws.tasks.get((uint64_t)pl.max_ovf * sizeof(OvfTask));
  OvfBucket* obuckets = (OvfBucket*)ws.obuckets.get((uint64_t)pl.max_ovf * sizeof(OvfBucket));
  Pt* partial = (Pt*)ws.partial.get((uint64_t)pl.max_ovf * sizeof(Pt));
  const uint32_t max_mid = pl.max_ovf / kOvfChunk + pl.max_ovf / kOvfSmall + 2;
  Pt* mid = (Pt*)ws.mid.get((uint64_t)max_mid * sizeof(Pt));
  // window sums run in one or two slice-sum levels
  const uint32_t kSlices = 64;
  const bool two_level = ngroups >= 4 * kSlices;
  // bucket reduction workspace (msm.cuh): [acc | ping | pong] of bwin * ngroups points each, then the block sums
  const uint64_t arr_pts = (uint64_t)pl.bwin * ngroups;
  Pt* groups = (Pt*)ws.groups.get((3 * arr_pts + (uint64_t)2 * pl.bwin * (kSlices + 1)) * sizeof(Pt));
  Pt* ping = groups + arr_pts;
  Pt* pong = ping + arr_pts;
  Pt* mids = pong + arr_pts;                                  // 2 bwin arrays x kSlices slice sums
  Pt* sums = mids + (uint64_t)2 * pl.bwin * kSlices;          // 2 bwin totals: [sum acc | sum g run]
  Pt* windows = (Pt*)ws.windows.get((uint64_t)pl.bwin * sizeof(Pt));
  OvfCounters* ctr = (OvfCounters*)ws.ctr.get(sizeof(OvfCounters));
  uint32_t* perm = (uint32_t*)ws.perm.get(total_b * 4);
  uint32_t* bins = (uint32_t*)ws.bins.get(2 * kSizeBins * 4);   // [bins | cursor]

  const int tok_sched = prof_begin(PROF_MSM_SCHED, s);
  B200_CUDA(cudaMemsetAsync(ctr, 0, sizeof(OvfCounters), s));
  B200_CUDA(cudaMemsetAsync(bins, 0, 2 * kSizeBins * 4, s));
  // ---- affine pre-reduction (single bucket array): halve the sorted list pl.pre times, then accumulate the
  //      reduced affine list directly
  const uint32_t* off = so.off;
  const uint32_t* end = so.end;
  const uint32_t* totals = so.totals;
  MsmPts acc_pts = pts;
  const bool direct = pl.pre > 0;
  if (direct) {
    if (pl.bwin != 1 || !pl.table) throw std::runtime_error("msm: pre-reduction needs a single table-mode base set");
    using El = typename F::El;
    const uint64_t slots = pl.stride;                       // upper bound of the padded entry count
    Affine<F>* bufA = (Affine<F>*)ws.pre_a.get((slots / 2 + 1) * sizeof(Affine<F>));
    Affine<F>* bufB = pl.pre > 1 ? (Affine<F>*)ws.pre_b.get((slots / 4 + 1) * sizeof(Affine<F>)) : nullptr;
    const Affine<F>* in = nullptr;
    Affine<F>* out = bufA;
    const int tok_pre = prof_begin(PROF_MSM_PRE, s);
    for (int l = 0; l < pl.pre; l++) {
      const uint64_t npairs = (slots >> l) / 2 + 1;
      uint32_t per_thread = (uint32_t)std::min<uint64_t>(144, std::max<uint64_t>(8, npairs / (kPreThreads * 148ull * 8)));
      const unsigned blocks = (unsigned)((npairs + (uint64_t)kPreThreads * per_thread - 1) / ((uint64_t)kPreThreads * per_thread));
      El* prefix = (El*)ws.pre_prefix.get((uint64_t)blocks * kPreThreads * per_thread * sizeof(El));
      // thread totals / their exclusive prefixes / chunk products / scratch of the grid-wide inversion
      const uint32_t nthreads = blocks * kPreThreads, nchunks = nthreads / kPreInvChunk;
      El* totals_t = (El*)ws.pre_tot.get(((uint64_t)2 * nthreads + 2 * nchunks + 4) * sizeof(El));
      El* tpre = totals_t + nthreads;
      El* cprod = tpre + nthreads;
      El* cscr = cprod + nchunks;
      const Affine<F>* src = l == 0 ? (const Affine<F>*)pts.p[0] : in;
      if (l == 0)
        k_msm_pre_fwd<F, true><<<blocks, kPreThreads, 0, s>>>(src, so.sorted, so.totals, 0, per_thread, prefix, totals_t);
      else
        k_msm_pre_fwd<F, false><<<blocks, kPreThreads, 0, s>>>(src, nullptr, so.totals, (uint32_t)l, per_thread, prefix, totals_t);
      k_msm_pre_inv_a<F><<<(nchunks + 63) / 64, 64, 0, s>>>(totals_t, nchunks, tpre, cprod);
      k_msm_pre_inv_b<F><<<1, kPreThreads, 0, s>>>(cprod, nchunks, cscr);
      k_msm_pre_inv_c<F><<<(nchunks + 63) / 64, 64, 0, s>>>(totals_t, nchunks, tpre, cprod);
      if (l == 0)
        k_msm_pre_bwd<F, true><<<blocks, kPreThreads, 0, s>>>(src, so.sorted, so.totals, 0, per_thread, out, prefix, totals_t);
      else
        k_msm_pre_bwd<F, false><<<blocks, kPreThreads, 0, s>>>(src, nullptr, so.totals, (uint32_t)l, per_thread, out, prefix, totals_t);
      in = out;
      out = (out == bufA) ? bufB : bufA;
    }
    uint32_t* o2 = (uint32_t*)ws.off2.get((2 * total_b + 4) * 4);
    uint32_t* e2 = o2 + total_b;
    uint32_t* t2 = e2 + total_b;
    k_msm_pre_offsets<<<(unsigned)((total_b + 255) / 256), 256, 0, s>>>(so.off, so.end, total_b, (uint32_t)pl.pre, o2, e2,
                                                                         so.totals, t2, 1);
    prof_end(tok_pre, s);
    prof_count_launches(5 * pl.pre + 1);
    off = o2;
    end = e2;
    totals = t2;
    acc_pts.p[0] = in;
  }
  // size-sorted bucket schedule
  const unsigned szblocks = (unsigned)((total_b + kSizeThreads * kSizePerThread - 1) / (kSizeThreads * kSizePerThread));
  k_msm_size_hist<<<szblocks, kSizeThreads, 0, s>>>(off, end, total_b, bins);
  k_msm_size_scan<<<1, kSizeBins, 0, s>>>(bins, bins + kSizeBins);
  k_msm_size_scatter<<<szblocks, kSizeThreads, 0, s>>>(off, end, total_b, bins + kSizeBins, perm);
  prof_end(tok_sched, s);
  const int tok_acc = prof_begin(GROUP == 2 ? PROF_MSM_ACC_G2 : PROF_MSM_ACC_G1, s);
  const unsigned acc_blocks = (unsigned)((total_b + B200_ACC_THREADS - 1) / B200_ACC_THREADS);
  if (direct)
    k_msm_accumulate<F, true><<<acc_blocks, B200_ACC_THREADS, 0, s>>>(acc_pts, nullptr, off, end, perm, totals, pl, buckets,
                                                                       obuckets, ctr);
  else
    k_msm_accumulate<F, false><<<acc_blocks, B200_ACC_THREADS, 0, s>>>(acc_pts, so.sorted, off, end, perm, totals, pl,
                                                                        buckets, obuckets, ctr);
  prof_end(tok_acc, s);
  // ---- latency-bound tail on the high-priority stream
  ws.hop_to_tail(s);
  cudaStream_t t = ws.tail;
  // oversized buckets: task list, partial sums, then a small-bucket merge and a two-level tree for big ones
  const int tok_ovf = prof_begin(PROF_MSM_OVF, t);
  k_msm_ovf_expand<<<148, 256, 0, t>>>(obuckets, ctr, pl, tasks);
  unsigned ovf_blocks = (unsigned)std::min<uint64_t>((pl.max_ovf + 127) / 128, 148 * 8);
  if (direct) k_msm_ovf_accumulate<F, true><<<ovf_blocks, 128, 0, t>>>(acc_pts, nullptr, pl, tasks, ctr, partial);
  else k_msm_ovf_accumulate<F, false><<<ovf_blocks, 128, 0, t>>>(acc_pts, so.sorted, pl, tasks, ctr, partial);
  k_msm_ovf_merge_small<F><<<148, 64, 0, t>>>(obuckets, ctr, partial, buckets);
  const size_t merge_smem = kOvfMergeThreads * sizeof(Pt);
  k_msm_ovf_merge_l1<F><<<dim3(16, 256), kOvfMergeThreads, merge_smem, t>>>(obuckets, ctr, partial, mid);
  k_msm_ovf_merge_l2<F><<<256, kOvfMergeThreads, merge_smem, t>>>(obuckets, ctr, mid, buckets);
  prof_end(tok_ovf, t);
  size_t red_smem = kReduceThreads * sizeof(Pt);
  const int tok_br = prof_begin(PROF_MSM_BUCKET_REDUCE, t);
  k_msm_wsum_level0<F><<<(unsigned)((arr_pts + 63) / 64), 64, 0, t>>>(buckets, pl.nb, pl.group, (uint32_t)pl.bwin, ngroups,
                                                                       groups, ping);
  int rounds = 0;
  for (uint32_t d = 1; d < ngroups; d <<= 1) {
    k_msm_suffix_round<F><<<(unsigned)((arr_pts + 127) / 128), 128, 0, t>>>(ping, pong, ngroups, d, (uint32_t)arr_pts);
    std::swap(ping, pong);
    rounds++;
  }
  prof_end(tok_br, t);
  const int tok_sums = prof_begin(PROF_MSM_SUMS, t);
  const uint32_t narr2 = 2u * pl.bwin;
  if (two_level) {
    k_msm_range_sum<F><<<dim3(kSlices, narr2), kReduceThreads, red_smem, t>>>(groups, ping, (uint32_t)pl.bwin, ngroups,
                                                                             (ngroups + kSlices - 1) / kSlices, mids);
    k_msm_range_sum<F><<<dim3(1, narr2), kReduceThreads, red_smem, t>>>(mids, mids, narr2, kSlices, kSlices, sums);
  } else {
    k_msm_range_sum<F><<<dim3(1, narr2), kReduceThreads, red_smem, t>>>(groups, ping, (uint32_t)pl.bwin, ngroups, ngroups, sums);
  }
  int log_g0 = 0;
  while ((1u << log_g0) < pl.group) log_g0++;
  k_msm_wsum_combine<F><<<(pl.bwin + 31) / 32, 32, 0, t>>>(sums, (uint32_t)pl.bwin, (uint32_t)log_g0, windows);
  k_msm_horner<F><<<1, 128, 0, t>>>(windows, pl, (Pt*)d_out);
  prof_end(tok_sums, t);
  // join: the caller's stream waits for the tail; otherwise the result is ready when ws.e_back fires
  // (ws.wait_tail(other_stream)) and the caller's stream may run ahead with independent bulk work
  ws.hop_back(s, join);
  prof_count_launches((two_level ? 17 : 16) + rounds);
  B200_CUDA(cudaGetLastError());
}

template <class F>
void msm_set_smem_attrs() {
  const int bytes = (int)(kOvfMergeThreads * sizeof(XYZZ<F>));
  if (bytes > 48 * 1024) {
    B200_CUDA(cudaFuncSetAttribute(k_msm_ovf_merge_l1<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    B200_CUDA(cudaFuncSetAttribute(k_msm_ovf_merge_l2<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
}

// ------------------------------------------------------------------------------------ backend
// ------------------------------------------------------------------------------------ pairing check (pairing.cuh)
template <class Cfg>
using PairOf = PairingT<typename Cfg::Fp, typename Cfg::Tower, typename Cfg::Fr::Params>;

// One pairing per block, one thread: the Miller value of (P_i, Q_i).  flags[i] = 1 when r * P_i != infinity.
template <class Cfg>
__global__ void k_pairing_miller(const typename Cfg::Fp::El* g1, const typename Cfg::Fp::El* g2, uint32_t n,
                                 typename PairOf<Cfg>::Ext* f_out, uint32_t* flags) {
  using PT = PairOf<Cfg>;
  const uint32_t i = blockIdx.x;
  if (threadIdx.x != 0 || i >= n) return;
  typename PT::Ext f;
  typename Cfg::Fp::El P[2], Q[2 * PT::NQ];      // operands in local memory: miller() takes plain pointers
  for (int j = 0; j < 2; j++) P[j] = g1[2 * (size_t)i + j];
  for (int j = 0; j < 2 * PT::NQ; j++) Q[j] = g2[2 * PT::NQ * (size_t)i + j];
  const bool in_subgroup = PT::miller(f, P, Q);
  f_out[i] = f;
  flags[i] = in_subgroup ? 0u : 1u;
}
// Batch variants: thread t handles pair t / check t, so the lanes of a warp run INDEPENDENT pairings through one
// instruction stream (the Miller loop follows the bits of r, the exponentiation the bits of (p^k - 1) / r: no
// data-dependent control flow once the inversions are Fermat exponentiations).  This is the throughput form of the
// verifier: checks per second scale with the number of resident threads, not with the latency of one thread.
template <class Cfg>
using PairOfBatch = PairingT<typename Cfg::Fp, typename Cfg::Tower, typename Cfg::Fr::Params, true>;

template <class Cfg>
__global__ void k_pairing_miller_batch(const typename Cfg::Fp::El* g1, const typename Cfg::Fp::El* g2, uint32_t n,
                                       typename PairOfBatch<Cfg>::Ext* f_out, uint32_t* flags) {
  using PT = PairOfBatch<Cfg>;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename PT::Ext f;
  typename Cfg::Fp::El P[2], Q[2 * PT::NQ];
  for (int j = 0; j < 2; j++) P[j] = g1[2 * (size_t)i + j];
  for (int j = 0; j < 2 * PT::NQ; j++) Q[j] = g2[2 * PT::NQ * (size_t)i + j];
  const bool in_subgroup = PT::miller(f, P, Q);
  f_out[i] = f;
  flags[i] = in_subgroup ? 0u : 1u;
}
template <class Cfg>
__global__ void k_pairing_finish_batch(const typename PairOfBatch<Cfg>::Ext* f, const uint32_t* flags, uint32_t per,
                                       uint32_t n_checks, int32_t* results) {
  using PT = PairOfBatch<Cfg>;
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_checks) return;
  typename PT::Ext acc, g;
  PT::ext_one(acc);
  uint32_t bad = 0;
  for (uint32_t j = 0; j < per; j++) {
    g = f[(size_t)c * per + j];
    bad |= flags[(size_t)c * per + j];
    PT::ext_mul(acc, acc, g);
  }
  PT::final_exp(g, acc);
  results[c] = bad ? -1 : (PT::ext_is_one(g) ? 1 : 0);
}
// One thread: product of the Miller values, final exponentiation, comparison with one.
template <class Cfg>
__global__ void k_pairing_finish(const typename PairOf<Cfg>::Ext* f, uint32_t n, typename PairOf<Cfg>::Ext* gt_out,
                                 uint32_t* is_one) {
  using PT = PairOf<Cfg>;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  typename PT::Ext acc, g;
  PT::ext_one(acc);
  for (uint32_t i = 0; i < n; i++) {
    g = f[i];
    PT::ext_mul(acc, acc, g);
  }
  PT::final_exp(g, acc);
  *gt_out = g;
  *is_one = PT::ext_is_one(g) ? 1u : 0u;
}

template <class Cfg>
struct CurveImpl : CurveBackend {
  using Fp = typename Cfg::Fp;
  using Fr = typename Cfg::Fr;
  using G1F = typename Cfg::G1F;
  using G2F = typename Cfg::G2F;

  int id() const override { return Cfg::ID; }
  const char* name() const override { return Cfg::name(); }
  size_t fr_bytes() const override { return sizeof(typename Fr::El); }
  size_t fp_bytes() const override { return sizeof(typename Fp::El); }
  size_t affine_bytes(int g) const override { return g == 1 ? sizeof(Affine<G1F>) : sizeof(Affine<G2F>); }
  size_t xyzz_bytes(int g) const override { return g == 1 ? sizeof(XYZZ<G1F>) : sizeof(XYZZ<G2F>); }
  int fr_bits() const override { return Fr::BITS; }

  void msm(int group, const void* d_points, const void* d_scalars, uint64_t n, void* d_out, MsmWorkspace& ws,
           cudaStream_t s, int c_override, MsmStats* stats, const uint32_t* d_index_map,
           const MsmBases* bases, bool join, int pre) override {
    if (bases && bases->group != group) throw std::runtime_error("msm: base tables belong to the other group");
    if (n >= (1ull << 31)) throw std::runtime_error("msm: n must be < 2^31");
    MsmPlan pl = bases ? make_msm_plan_table(n, Fr::BITS, bases->c, bases->npts, 1, msm_pre_levels(n, bases->c, pre),
                                             bases->tstride)
                       : make_msm_plan(n, Fr::BITS, c_override);
    if (stats) *stats = MsmStats{pl.c, pl.nwin, pl.nb, pl.task, pl.group};
    if (n == 0) {
      B200_CUDA(cudaMemsetAsync(d_out, 0, xyzz_bytes(group), s));
      if (!join) {   // keep the contract: the result is ready when the tail event fires
        ws.hop_to_tail(s);
        ws.hop_back(s, false);
      }
      return;
    }
    MsmSets sets{};
    sets.map[0] = d_index_map;
    sets.npts[0] = bases ? bases->npts : 0;
    MsmPts pts{};
    pts.p[0] = bases ? bases->tables.p : d_points;
    MsmSorted so;
    const int tok_total = prof_begin(group == 2 ? PROF_MSM_TOTAL_G2 : PROF_MSM_TOTAL_G1, s);
    msm_sort_launch<Fr>(d_scalars, pl, sets, ws, s, so);
    reduce_dispatch(group, so, pts, d_out, ws, s, join);
    prof_end(tok_total, s);
  }

  // affine pre-reduction levels for a table-mode MSM of n scalars: `want` < 0 = automatic (dense scalars assumed:
  // one level per factor of two of the mean bucket load above 8, at most 3), B200_MSM_PRE overrides
  static int msm_pre_levels(uint64_t n, int c, int want) {
    const char* e = std::getenv("B200_MSM_PRE");   // read per call: the tests switch it
    const int env = e ? std::atoi(e) : -1;
    if (env >= 0) return std::min(env, 4);
    if (want >= 0) return std::min(want, 4);
    return 0;
  }

  void reduce_dispatch(int group, const MsmSorted& so, const MsmPts& pts, void* d_out, MsmWorkspace& ws,
                       cudaStream_t s, bool join = true) {
    if (group == 1) {
      msm_set_smem_attrs<G1F>();
      msm_reduce_launch<G1F, 1>(so, pts, d_out, ws, s, join);
    } else {
      msm_set_smem_attrs<G2F>();
      msm_reduce_launch<G2F, 2>(so, pts, d_out, ws, s, join);
    }
  }

  void msm_batch(int group, const MsmBases& bases, const void* d_scalars, uint64_t n_per, uint32_t batch,
                 const uint32_t* d_index_map, void* d_out, MsmWorkspace& ws, cudaStream_t s) override {
    if (bases.group != group) throw std::runtime_error("msm_batch: base tables belong to the other group");
    if (n_per == 0 || batch == 0 || n_per * batch >= (1ull << 31) || batch > 65535)
      throw std::runtime_error("msm_batch: bad sizes");
    MsmPlan pl = make_msm_plan_table(n_per, Fr::BITS, bases.c, bases.npts, (int)batch);
    pl.n = n_per * batch;
    pl.batch_n = (uint32_t)n_per;
    MsmSets sets{};
    sets.map[0] = d_index_map;
    sets.npts[0] = bases.npts;
    MsmPts pts{};
    pts.p[0] = bases.tables.p;
    MsmSorted so;
    msm_sort_launch<Fr>(d_scalars, pl, sets, ws, s, so);
    reduce_dispatch(group, so, pts, d_out, ws, s, true);
  }

  void msm_sort(const void* d_scalars, uint64_t n, const MsmBases* const* bases, const uint32_t* const* maps,
                int nsets, MsmWorkspace& ws, cudaStream_t s, MsmSorted& out) override {
    if (nsets < 1 || nsets > kMaxSets) throw std::runtime_error("msm_sort: 1..4 base sets");
    if (n == 0 || n >= (1ull << 31)) throw std::runtime_error("msm_sort: n must be in [1, 2^31)");
    MsmSets sets{};
    for (int j = 0; j < nsets; j++) {
      if (bases[j]->c != bases[0]->c || bases[j]->nwin != bases[0]->nwin || bases[j]->tstride != bases[0]->tstride)
        throw std::runtime_error("msm_sort: base sets must share the window width and table stride");
      sets.map[j] = maps[j];
      sets.npts[j] = bases[j]->npts;
    }
    MsmPlan pl = make_msm_plan_table(n, Fr::BITS, bases[0]->c, bases[0]->npts, nsets, 0, bases[0]->tstride);
    msm_sort_launch<Fr>(d_scalars, pl, sets, ws, s, out);
  }

  void msm_reduce(int group, const MsmSorted& so, int first_set, int count, const MsmBases* const* bases,
                  void* d_out, MsmWorkspace& ws, cudaStream_t s, bool join) override {
    const int ts = so.pl.tstride;
    if (!so.pl.table || first_set < 0 || count < 1 || (first_set + count) * ts > so.pl.bwin)
      throw std::runtime_error("msm_reduce: bad set range");
    MsmSorted sub = so;
    sub.pl = make_msm_plan_table(so.pl.n, Fr::BITS, so.pl.c, so.pl.npts, count, 0, ts);
    sub.off += (uint64_t)first_set * ts * so.pl.nb;
    sub.end += (uint64_t)first_set * ts * so.pl.nb;
    sub.sorted += (uint64_t)first_set * ts * so.pl.stride;
    sub.totals += first_set * ts;
    MsmPts pts{};
    for (int j = 0; j < count; j++) {
      if (bases[j]->group != group) throw std::runtime_error("msm_reduce: base tables belong to the other group");
      pts.p[j] = bases[j]->tables.p;
    }
    reduce_dispatch(group, sub, pts, d_out, ws, s, join);
  }

  int table_window(uint64_t npts, int tstride) const override { return msm_table_window(npts, Fr::BITS, tstride); }

  void build_tables(MsmBases& b, int group, const void* d_points, uint64_t npts, int window_bits,
                    cudaStream_t s, int tstride) override {
    b.group = group;
    b.npts = npts;
    b.c = window_bits > 0 ? window_bits : msm_table_window(npts, Fr::BITS, tstride);
    b.nwin = msm_nwin(Fr::BITS, b.c);
    b.tstride = std::max(1, std::min(tstride, b.nwin));
    b.ntab = msm_ntables(b.nwin, b.tstride);
    if ((uint64_t)b.ntab * npts >= (1ull << 31))
      throw std::runtime_error("base set too large for its table depth (raise the table stride)");
    const size_t pb = affine_bytes(group);
    void* t = b.tables.get(std::max<uint64_t>(npts, 1) * b.ntab * pb);
    if (!npts) return;
    unsigned blocks = (unsigned)((npts + 63) / 64);
    // scratch for the shared inversion: 3 coordinates per (table, point); freed (stream-ordered) after the kernel
    DevBuf scratch;
    void* sc = scratch.get(std::max<uint64_t>(npts * (uint64_t)(b.ntab - 1), 1) * 3 * (pb / 2));
    const int dbls = b.c * b.tstride;   // doublings between consecutive tables
    if (group == 1)
      k_build_tables<G1F><<<blocks, 64, 0, s>>>((const Affine<G1F>*)d_points, npts, dbls, b.ntab, (Affine<G1F>*)t,
                                                (typename G1F::El*)sc);
    else
      k_build_tables<G2F><<<blocks, 64, 0, s>>>((const Affine<G2F>*)d_points, npts, dbls, b.ntab, (Affine<G2F>*)t,
                                                (typename G2F::El*)sc);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(s));   // the scratch buffer is released on return
  }

  // ---------------------------------------------------------------- NTT
  using FrEl = typename Fr::El;
  static constexpr int kElogMax = sizeof(FrEl) <= 32 ? 11 : 10;   // tile <= 64 KB (32-B fr) / 48 KB (48-B fr)
  static constexpr int kStridedMax = 8;

  static std::vector<NttPass> plan_passes(int logn, bool dit) {
    std::vector<NttPass> v;
    int lc = std::min(kElogMax, logn);
    v.push_back(NttPass{logn, 0, lc, 0});
    int rest = logn - lc;
    int k = (rest + kStridedMax - 1) / kStridedMax;
    int s = lc;
    for (int i = 0; i < k; i++) {
      int r = (rest + (k - i) - 1) / (k - i);
      rest -= r;
      v.push_back(NttPass{logn, s, r, std::min(s, kElogMax - r)});
      s += r;
    }
    if (!dit) std::reverse(v.begin(), v.end());
    return v;
  }

  // passes [first, last) of the transform's plan; `pre` applies to pass 0 and `post` to the final pass of the WHOLE plan
  // (so a caller running a sub-range only gets them when the range touches that end)
  template <bool DIT>
  static void run_passes(int logn, FrEl* data, const FrEl* tw, NttScale pre, NttScale post, const FrEl* in_b,
                         const FrEl* in_c, const FrEl* den, cudaStream_t s, int first_pass = 0, int last_pass = -1) {
    // per-device attribute; cheap enough to set on every call (device pools live in one process)
    B200_CUDA(cudaFuncSetAttribute(k_ntt_pass<Fr, DIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(sizeof(FrEl) << kElogMax)));
    auto passes = plan_passes(logn, DIT);
    if (last_pass < 0) last_pass = (int)passes.size();
    const NttScale none{SCALE_NONE, 0, nullptr, nullptr, nullptr};
    for (int i = first_pass; i < last_pass; i++) {
      const NttPass& ps = passes[i];
      int elog = ps.logr + ps.lo_tile_log;
      unsigned blocks = 1u << (logn - elog);
      size_t smem = sizeof(FrEl) << elog;
      bool first = i == 0, last = i + 1 == (int)passes.size();
      const int tok = prof_begin(PROF_NTT_PASS, s);
      k_ntt_pass<Fr, DIT><<<blocks, kNttThreads, smem, s>>>(data, tw, ps, first ? pre : none, last ? post : none,
                                                           first ? in_b : nullptr, first ? in_c : nullptr, den);
      prof_end(tok, s);
    }
    prof_count_launches(last_pass - first_pass);
    B200_CUDA(cudaGetLastError());
  }

  // "interpolate, then evaluate on the coset" for one vector: inverse DIF without its last (contiguous) pass, the fused
  // middle (k_ntt_fused_mid), coset DIT without its first pass and - when `skip_last_dit` - without its last one
  // (which the fused quotient kernel performs).  npasses = passes of one plain transform.
  static void coset_evals_fused(NttDomain& d, FrEl* v, bool skip_last_dit, cudaStream_t s) {
    const NttScale none{SCALE_NONE, 0, nullptr, nullptr, nullptr};
    const FrEl* twf = (const FrEl*)d.tw_fwd.p;
    const FrEl* twi = (const FrEl*)d.tw_inv.p;
    const NttScale cs{SCALE_POW_BITREV, d.lo_bits, d.g_lo.p, d.g_hi_scaled.p, nullptr};
    auto dit = plan_passes(d.logn, true);
    const int np = (int)dit.size();
    run_passes<false>(d.logn, v, twi, none, none, nullptr, nullptr, nullptr, s, 0, np - 1);
    const NttPass& mid = dit[0];                       // contiguous pass: s_log = 0, lo_tile_log = 0
    const size_t smem = sizeof(FrEl) << mid.logr;
    B200_CUDA(cudaFuncSetAttribute(k_ntt_fused_mid<Fr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tok = prof_begin(PROF_NTT_PASS, s);
    k_ntt_fused_mid<Fr><<<1u << (d.logn - mid.logr), kNttThreads, smem, s>>>(v, twi, twf, mid, cs);
    prof_end(tok, s);
    prof_count_launches(1);
    run_passes<true>(d.logn, v, twf, none, none, nullptr, nullptr, nullptr, s, 1, skip_last_dit ? np - 1 : np);
  }

  void domain_init(NttDomain& d, int logn, const void* d_omega, const void* d_g, cudaStream_t s) override {
    if (logn < 1 || logn > 30) throw std::runtime_error("domain size out of range");
    d.logn = logn;
    d.lo_bits = (logn + 1) / 2;
    const uint64_t n = 1ull << logn, half = n >> 1;
    const uint64_t nlo = 1ull << d.lo_bits, nhi = 1ull << (logn - d.lo_bits);
    FrEl* consts = (FrEl*)d.consts.get(6 * sizeof(FrEl));
    k_domain_consts<Fr><<<1, 1, 0, s>>>(consts, (const FrEl*)d_omega, (const FrEl*)d_g, logn);
    auto table = [&](DevBuf& buf, const FrEl* base, const FrEl* scale, uint64_t count, uint64_t step) {
      FrEl* out = (FrEl*)buf.get(count * sizeof(FrEl));
      k_pow_table<Fr><<<(unsigned)((count + 127) / 128), 128, 0, s>>>(out, base, scale, count, step);
    };
    table(d.tw_fwd, consts + 4, nullptr, half, 1);
    table(d.tw_inv, consts + 0, nullptr, half, 1);
    table(d.g_lo, consts + 5, nullptr, nlo, 1);
    table(d.g_hi, consts + 5, nullptr, nhi, nlo);
    table(d.g_hi_scaled, consts + 5, consts + 2, nhi, nlo);
    table(d.gi_lo, consts + 1, nullptr, nlo, 1);
    table(d.gi_hi_scaled, consts + 1, consts + 2, nhi, nlo);
    B200_CUDA(cudaGetLastError());
  }

  void ntt(NttDomain& d, void* d_data, bool inverse, bool dit, bool coset, cudaStream_t s) override {
    const FrEl* consts = (const FrEl*)d.consts.p;
    NttScale pre{SCALE_NONE, 0, nullptr, nullptr, nullptr}, post = pre;
    if (!inverse) {
      if (coset) pre = NttScale{dit ? SCALE_POW_BITREV : SCALE_POW_NATURAL, d.lo_bits, d.g_lo.p, d.g_hi.p, nullptr};
    } else {
      if (coset) post = NttScale{dit ? SCALE_POW_NATURAL : SCALE_POW_BITREV, d.lo_bits, d.gi_lo.p, d.gi_hi_scaled.p, nullptr};
      else post = NttScale{SCALE_CONST, 0, nullptr, nullptr, consts + 2};
    }
    const FrEl* tw = (const FrEl*)(inverse ? d.tw_inv.p : d.tw_fwd.p);
    if (dit) run_passes<true>(d.logn, (FrEl*)d_data, tw, pre, post, nullptr, nullptr, nullptr, s);
    else run_passes<false>(d.logn, (FrEl*)d_data, tw, pre, post, nullptr, nullptr, nullptr, s);
  }

  void compute_h(NttDomain& d, void* d_a, void* d_b, void* d_c, cudaStream_t s) override {
    const FrEl* consts = (const FrEl*)d.consts.p;
    const NttScale none{SCALE_NONE, 0, nullptr, nullptr, nullptr};
    const FrEl* twf = (const FrEl*)d.tw_fwd.p;
    const FrEl* twi = (const FrEl*)d.tw_inv.p;
    FrEl* v[3] = {(FrEl*)d_a, (FrEl*)d_b, (FrEl*)d_c};
    // (a*b - c) / (g^n - 1) on the coset, then back: g^-i / n on store of the last inverse pass
    NttScale ci{SCALE_POW_BITREV, d.lo_bits, d.gi_lo.p, d.gi_hi_scaled.p, nullptr};
    auto dit = plan_passes(d.logn, true);
    const int np = (int)dit.size();
    static const bool fuse = [] {
      const char* e = std::getenv("B200_NTT_FUSE");   // 0 restores the 21-pass schedule (A/B measurements)
      return !(e && std::atoi(e) == 0);
    }();
    if (np < 2 || !fuse) {
      // one pass per transform (n <= 2^11) or fusion disabled: the plain 7-transform schedule
      for (int k = 0; k < 3; k++) run_passes<false>(d.logn, v[k], twi, none, none, nullptr, nullptr, nullptr, s);
      NttScale cs{SCALE_POW_BITREV, d.lo_bits, d.g_lo.p, d.g_hi_scaled.p, nullptr};
      for (int k = 0; k < 3; k++) run_passes<true>(d.logn, v[k], twf, cs, none, nullptr, nullptr, nullptr, s);
      run_passes<false>(d.logn, v[0], twi, none, ci, v[1], v[2], consts + 3, s);
      return;
    }
    // 15 passes instead of 21 at three passes per transform: per vector  (np - 1) + 1 fused + (np - 2),  then ONE fused
    // pass (last DIT level group of a, b, c + pointwise + first DIF level group), then the remaining np - 1 DIF passes
    for (int k = 0; k < 3; k++) coset_evals_fused(d, v[k], true, s);
    const NttPass& ps = dit[np - 1];                    // == first pass of the DIF plan
    const int elog = ps.logr + ps.lo_tile_log;
    const size_t smem = 3 * (sizeof(FrEl) << elog);
    B200_CUDA(cudaFuncSetAttribute(k_ntt_fused_quot<Fr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tok = prof_begin(PROF_NTT_PASS, s);
    k_ntt_fused_quot<Fr><<<1u << (d.logn - elog), kNttQuotThreads, smem, s>>>(v[0], v[1], v[2], twf, twi, ps, consts + 3);
    prof_end(tok, s);
    prof_count_launches(1);
    run_passes<false>(d.logn, v[0], twi, none, ci, nullptr, nullptr, nullptr, s, 1, np);
  }

  void coset_evals(NttDomain& d, void* d_v, cudaStream_t s) override {
    if (plan_passes(d.logn, true).size() >= 2) {
      coset_evals_fused(d, (FrEl*)d_v, false, s);
      return;
    }
    const NttScale none{SCALE_NONE, 0, nullptr, nullptr, nullptr};
    run_passes<false>(d.logn, (FrEl*)d_v, (const FrEl*)d.tw_inv.p, none, none, nullptr, nullptr, nullptr, s);
    NttScale cs{SCALE_POW_BITREV, d.lo_bits, d.g_lo.p, d.g_hi_scaled.p, nullptr};
    run_passes<true>(d.logn, (FrEl*)d_v, (const FrEl*)d.tw_fwd.p, cs, none, nullptr, nullptr, nullptr, s);
  }

  void compute_h_tail(NttDomain& d, void* d_a, void* d_b, void* d_c, cudaStream_t s) override {
    const FrEl* consts = (const FrEl*)d.consts.p;
    const NttScale none{SCALE_NONE, 0, nullptr, nullptr, nullptr};
    NttScale ci{SCALE_POW_BITREV, d.lo_bits, d.gi_lo.p, d.gi_hi_scaled.p, nullptr};
    run_passes<false>(d.logn, (FrEl*)d_a, (const FrEl*)d.tw_inv.p, none, ci, (const FrEl*)d_b, (const FrEl*)d_c, consts + 3, s);
  }

  // ---------------------------------------------------------------- point (de)compression, all curves
  // curve coefficients b (G1) and b' (G2), computed once per device
  std::mutex consts_mu;
  std::map<int, std::unique_ptr<DevBuf>> curve_consts;
  const uint8_t* curve_b_dev(cudaStream_t s) {
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(consts_mu);
    auto& slot = curve_consts[dev];
    if (!slot) {
      slot.reset(new DevBuf());
      uint8_t* p = (uint8_t*)slot->get(sizeof(typename G1F::El) + sizeof(typename G2F::El));
      k_curve_consts<Cfg><<<1, 1, 0, s>>>((typename G1F::El*)p, (typename G2F::El*)(p + sizeof(typename G1F::El)));
      B200_CUDA(cudaGetLastError());
      B200_CUDA(cudaStreamSynchronize(s));
    }
    return (const uint8_t*)slot->p;
  }
  size_t compressed_bytes(int group) const override {
    return group == 1 ? (size_t)CoordSerde<G1F>::BYTES : (size_t)CoordSerde<G2F>::BYTES;
  }
  void points_decompress(int group, const void* d_bytes, void* d_affine, uint64_t n, uint32_t* d_err,
                         cudaStream_t s) override {
    if (!n) return;
    const uint8_t* cb = curve_b_dev(s);
    const unsigned blocks = (unsigned)((n + 63) / 64);
    if (group == 1)
      k_points_decompress<G1F><<<blocks, 64, 0, s>>>((const uint8_t*)d_bytes, (Affine<G1F>*)d_affine, n, Cfg::FLAG_BITS,
                                                     (const typename G1F::El*)cb, d_err);
    else
      k_points_decompress<G2F><<<blocks, 64, 0, s>>>((const uint8_t*)d_bytes, (Affine<G2F>*)d_affine, n, Cfg::FLAG_BITS,
                                                     (const typename G2F::El*)(cb + sizeof(typename G1F::El)), d_err);
    prof_count_launches(1);
    B200_CUDA(cudaGetLastError());
  }
  size_t gt_bytes() const override { return sizeof(typename PairOf<Cfg>::Ext); }
  void pairing_check(const void* d_g1, const void* d_g2, uint32_t n, void* d_f, void* d_gt, uint32_t* d_flags,
                     cudaStream_t s) override {
    using Ext = typename PairOf<Cfg>::Ext;
    if (n) {
      k_pairing_miller<Cfg><<<n, 1, 0, s>>>((const typename Fp::El*)d_g1, (const typename Fp::El*)d_g2, n, (Ext*)d_f, d_flags);
      B200_CUDA(cudaGetLastError());
    }
    k_pairing_finish<Cfg><<<1, 1, 0, s>>>((const Ext*)d_f, n, (Ext*)d_gt, d_flags + n);
    prof_count_launches(n ? 2 : 1);
    B200_CUDA(cudaGetLastError());
  }

  void pairing_check_batch(const void* d_g1, const void* d_g2, uint32_t per, uint32_t n_checks, void* d_f, uint32_t* d_flags,
                           int32_t* d_results, cudaStream_t s) override {
    using Ext = typename PairOfBatch<Cfg>::Ext;
    const uint32_t n = per * n_checks;
    if (!n_checks) return;
    if (n) {
      k_pairing_miller_batch<Cfg><<<(n + 31) / 32, 32, 0, s>>>((const typename Fp::El*)d_g1, (const typename Fp::El*)d_g2, n,
                                                               (Ext*)d_f, d_flags);
      B200_CUDA(cudaGetLastError());
    }
    k_pairing_finish_batch<Cfg><<<(n_checks + 31) / 32, 32, 0, s>>>((const Ext*)d_f, d_flags, per, n_checks, d_results);
    prof_count_launches(n ? 2 : 1);
    B200_CUDA(cudaGetLastError());
  }

  void points_compress(int group, const void* d_affine, void* d_bytes, uint64_t n, cudaStream_t s) override {
    if (!n) return;
    const unsigned blocks = (unsigned)((n + 63) / 64);
    if (group == 1)
      k_points_compress<G1F><<<blocks, 64, 0, s>>>((const Affine<G1F>*)d_affine, (uint8_t*)d_bytes, n, Cfg::FLAG_BITS);
    else
      k_points_compress<G2F><<<blocks, 64, 0, s>>>((const Affine<G2F>*)d_affine, (uint8_t*)d_bytes, n, Cfg::FLAG_BITS);
    prof_count_launches(1);
    B200_CUDA(cudaGetLastError());
  }

  void g1_decompress(const void* d_bytes, void* d_affine, uint32_t n, uint32_t* d_err, cudaStream_t s) override {
    if constexpr (Cfg::ID == 3) {
      k_g1_decompress_bls<Fp><<<(n + 63) / 64, 64, 0, s>>>((const uint8_t*)d_bytes, (Affine<Fp>*)d_affine, n, 4, d_err);
      B200_CUDA(cudaGetLastError());
    } else {
      throw std::runtime_error("g1_decompress: only BLS12-381 (EIP-4844 SRS) is supported");
    }
  }
  void g1_compress(const void* d_affine, void* d_bytes, uint32_t n, cudaStream_t s) override {
    if constexpr (Cfg::ID == 3) {
      k_g1_compress_bls<Fp><<<(n + 63) / 64, 64, 0, s>>>((const Affine<Fp>*)d_affine, (uint8_t*)d_bytes, n);
      B200_CUDA(cudaGetLastError());
    } else {
      throw std::runtime_error("g1_compress: only BLS12-381 is supported");
    }
  }
  void blob_to_scalars(const void* d_blob, void* d_scalars, uint32_t n, uint32_t* d_err, cudaStream_t s) override {
    if constexpr (Cfg::ID == 3) {
      k_blob_to_scalars<Fr><<<(n + 127) / 128, 128, 0, s>>>((const uint8_t*)d_blob, (FrEl*)d_scalars, n, d_err);
      B200_CUDA(cudaGetLastError());
    } else {
      throw std::runtime_error("blob_to_scalars: only BLS12-381 is supported");
    }
  }

  void kzg_roots(void* d_roots, uint32_t n, cudaStream_t s) override {
    if constexpr (Cfg::ID == 3) {
      int logn = 0;
      while ((1u << logn) < n) logn++;
      k_kzg_roots<Fr><<<(n + 127) / 128, 128, 0, s>>>((FrEl*)d_roots, n, logn);
      B200_CUDA(cudaGetLastError());
    } else {
      throw std::runtime_error("kzg_roots: only BLS12-381 is supported");
    }
  }
  void kzg_open(const void* d_p, const void* d_roots, const void* d_z_be, uint32_t n, void* d_q, void* d_y_bytes,
                void* d_scratch, uint32_t* d_err, cudaStream_t s) override {
    if constexpr (Cfg::ID == 3) {
      int logn = 0;
      while ((1u << logn) < n) logn++;
      const uint32_t nblocks = (n + kKzgThreads - 1) / kKzgThreads;
      FrEl* inv = (FrEl*)d_scratch;
      FrEl* partial = inv + n;
      FrEl* partial2 = partial + nblocks;
      FrEl* z = partial2 + nblocks;
      FrEl* y = z + 1;
      uint32_t* hit = (uint32_t*)(y + 1);
      const FrEl* p = (const FrEl*)d_p;
      const FrEl* roots = (const FrEl*)d_roots;
      k_kzg_load_point<Fr><<<1, 1, 0, s>>>((const uint8_t*)d_z_be, z, hit, d_err);
      k_kzg_open_terms<Fr><<<nblocks, kKzgThreads, 0, s>>>(p, roots, z, inv, partial, hit, n);
      k_kzg_open_y<Fr><<<1, 32, 0, s>>>(partial, nblocks, p, z, hit, n, logn, y, (uint8_t*)d_y_bytes);
      k_kzg_open_quotient<Fr><<<nblocks, kKzgThreads, 0, s>>>(p, roots, inv, y, (FrEl*)d_q, partial2, n);
      k_kzg_open_fix<Fr><<<1, 32, 0, s>>>(partial2, nblocks, z, hit, (FrEl*)d_q);
      prof_count_launches(5);
      B200_CUDA(cudaGetLastError());
    } else {
      throw std::runtime_error("kzg_open: only BLS12-381 is supported");
    }
  }

  void kzg_cell_shifts(void* d_shifts, uint32_t ncells, uint32_t m, uint32_t n, cudaStream_t s) override {
    if constexpr (Cfg::ID == 3) {
      int log_ext = 0;
      while ((1u << log_ext) < 2 * n) log_ext++;
      k_kzg_cell_shifts<Fr><<<(ncells + 63) / 64, 64, 0, s>>>((FrEl*)d_shifts, ncells, m, log_ext);
      B200_CUDA(cudaGetLastError());
    } else {
      throw std::runtime_error("kzg_cell_shifts: only BLS12-381 is supported");
    }
  }
  void kzg_cell_quotients(const void* d_coeffs, const void* d_shifts, void* d_q, uint32_t n, uint32_t m, uint32_t ncells,
                          cudaStream_t s) override {
    if constexpr (Cfg::ID == 3) {
      k_kzg_cell_quotients<Fr><<<(ncells * m + 127) / 128, 128, 0, s>>>((const FrEl*)d_coeffs, (const FrEl*)d_shifts,
                                                                         (FrEl*)d_q, n, m, ncells);
      prof_count_launches(1);
      B200_CUDA(cudaGetLastError());
    } else {
      throw std::runtime_error("kzg_cell_quotients: only BLS12-381 is supported");
    }
  }

  void scale_vec(void* d_x, const void* d_c, uint64_t n, cudaStream_t s) override {
    if (!n) return;
    k_scale_vec<Fr><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((FrEl*)d_x, (const FrEl*)d_c, n);
    B200_CUDA(cudaGetLastError());
  }

  void prep_rs(const void* d_r, const void* d_s, void* d_out4, cudaStream_t s) override {
    prof_count_launches(1);
    k_prep_rs<Fr><<<1, 1, 0, s>>>((const FrEl*)d_r, (const FrEl*)d_s, (FrEl*)d_out4);
    B200_CUDA(cudaGetLastError());
  }

  void assemble(const AssembleArgs& a, cudaStream_t s, int phases) override {
    using P1 = XYZZ<G1F>;
    if (phases & 1) {
      prof_count_launches(1);
      k_assemble_mul<G1F, Fr><<<2, 1, 0, s>>>((const P1*)a.ar_msm, (const P1*)a.bs1_msm, (const FrEl*)a.rs, (P1*)a.tmp);
    }
    if (phases & 4) {
      prof_count_launches(1);
      k_assemble_fold<G1F><<<1, 1, 0, s>>>((P1*)a.k_msm, (const P1*)a.tmp);
    }
    if (phases & 2) {
      prof_count_launches(1);
      k_assemble_out<G1F, G2F><<<4, 1, 0, s>>>((const P1*)a.ar_msm, (const XYZZ<G2F>*)a.bs2_msm, (const P1*)a.k_msm,
                                                (const P1*)a.z_msm, (const P1*)a.pok_msm, (const P1*)a.tmp,
                                                (Affine<G1F>*)a.out_ar, (Affine<G2F>*)a.out_bs,
                                                (Affine<G1F>*)a.out_krs, (Affine<G1F>*)a.out_pok);
    }
    B200_CUDA(cudaGetLastError());
  }

  void sum_sets(const void* d_partials, uint32_t nparts, void* d_sums, cudaStream_t s) override {
    k_sum_sets<G1F, G2F><<<6, 32, 0, s>>>((const uint8_t*)d_partials, nparts, (uint8_t*)d_sums);
    prof_count_launches(1);
    B200_CUDA(cudaGetLastError());
  }

  void sum_partials(int group, const void* d_xyzz, uint32_t count, void* d_aff, cudaStream_t s) override {
    if (group == 1) k_sum_partials<G1F><<<1, 1, 0, s>>>((const XYZZ<G1F>*)d_xyzz, count, (Affine<G1F>*)d_aff);
    else k_sum_partials<G2F><<<1, 1, 0, s>>>((const XYZZ<G2F>*)d_xyzz, count, (Affine<G2F>*)d_aff);
    prof_count_launches(1);
    B200_CUDA(cudaGetLastError());
  }

  void to_affine(int group, const void* d_xyzz, void* d_aff, uint32_t count, cudaStream_t s) override {
    if (!count) return;
    if (group == 1) k_to_affine<G1F><<<(count + 31) / 32, 32, 0, s>>>((const XYZZ<G1F>*)d_xyzz, (Affine<G1F>*)d_aff, count);
    else k_to_affine<G2F><<<(count + 31) / 32, 32, 0, s>>>((const XYZZ<G2F>*)d_xyzz, (Affine<G2F>*)d_aff, count);
    B200_CUDA(cudaGetLastError());
  }

  template <class F>
  static void field_op(int op, const void* a, const void* b, void* out, uint64_t n, cudaStream_t s) {
    unsigned blocks = (unsigned)((n + 63) / 64);
    if (op == OP_FROM_MONT || op == OP_TO_MONT) {
      if constexpr (std::is_same<F, G2F>::value && !std::is_same<G2F, Fp>::value) {
        throw std::runtime_error("from/to_mont not defined for Fp2 debug op");
      } else {
        k_dbg_field_mont<F><<<blocks, 64, 0, s>>>(op, (const typename F::El*)a, (typename F::El*)out, n);
      }
    } else {
      k_dbg_field<F><<<blocks, 64, 0, s>>>(op, (const typename F::El*)a, (const typename F::El*)b,
                                           (typename F::El*)out, n);
    }
    B200_CUDA(cudaGetLastError());
  }

  void dbg_field_op(int field, int op, const void* a, const void* b, void* out, uint64_t n, cudaStream_t s) override {
    if (!n) return;
    if (field == FIELD_FP) field_op<Fp>(op, a, b, out, n, s);
    else if (field == FIELD_FR) field_op<Fr>(op, a, b, out, n, s);
    else if (field == FIELD_FP2) {
      if constexpr (std::is_same<G2F, Fp>::value) throw std::runtime_error("curve has no Fp2 (G2 is over Fp)");
      else field_op<G2F>(op, a, b, out, n, s);
    } else throw std::runtime_error("bad field selector");
  }

  template <class F>
  static void dbg_ec_launch(int op, const void* a, const void* b, void* out, uint64_t n, cudaStream_t s) {
    const unsigned blocks = (unsigned)((n + 31) / 32);
    switch (op) {
      case EC_MADD: k_dbg_ec<F, Fr, EC_MADD><<<blocks, 32, 0, s>>>(a, b, out, n); break;
      case EC_ADD: k_dbg_ec<F, Fr, EC_ADD><<<blocks, 32, 0, s>>>(a, b, out, n); break;
      case EC_DBL: k_dbg_ec<F, Fr, EC_DBL><<<blocks, 32, 0, s>>>(a, b, out, n); break;
      case EC_TO_AFFINE: k_dbg_ec<F, Fr, EC_TO_AFFINE><<<blocks, 32, 0, s>>>(a, b, out, n); break;
      case EC_MUL_SCALAR: k_dbg_ec<F, Fr, EC_MUL_SCALAR><<<blocks, 32, 0, s>>>(a, b, out, n); break;
      default: throw std::runtime_error("bad EC op");
    }
  }
  void dbg_ec_op(int group, int op, const void* a, const void* b, void* out, uint64_t n, cudaStream_t s) override {
    if (!n) return;
    if (group == 1) dbg_ec_launch<G1F>(op, a, b, out, n, s);
    else dbg_ec_launch<G2F>(op, a, b, out, n, s);
    B200_CUDA(cudaGetLastError());
  }

  void fixed_base(int group, const void* d_base, const void* d_scalars, uint64_t n, void* d_out,
                  cudaStream_t s) override {
    if (!n) return;
    unsigned blocks = (unsigned)((n + 63) / 64);
    if (group == 1)
      k_fixed_base<G1F, Fr><<<blocks, 64, 0, s>>>((const Affine<G1F>*)d_base, (const FrEl*)d_scalars, n, (Affine<G1F>*)d_out);
    else
      k_fixed_base<G2F, Fr><<<blocks, 64, 0, s>>>((const Affine<G2F>*)d_base, (const FrEl*)d_scalars, n, (Affine<G2F>*)d_out);
    B200_CUDA(cudaGetLastError());
  }

  void calib_mul(int field, void* d_inout, uint64_t nthreads, int iters, cudaStream_t s) override {
    unsigned blocks = (unsigned)((nthreads + 255) / 256);
    if (field == FIELD_FP) k_calib_mul<Fp><<<blocks, 256, 0, s>>>((typename Fp::El*)d_inout, nthreads, iters);
    else k_calib_mul<Fr><<<blocks, 256, 0, s>>>((typename Fr::El*)d_inout, nthreads, iters);
    B200_CUDA(cudaGetLastError());
  }
};

}  // namespace b200
