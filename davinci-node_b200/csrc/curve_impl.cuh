// Per-curve backend: instantiates the templated kernels for one curve configuration and exposes
// them through the CurveBackend interface.  Included by curve_<name>.cu only.
#pragma once
#include <algorithm>
#include <cmath>

#include "backend.h"
#include "msm.cuh"
#include "ntt.cuh"

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

template <class F, class Fr>
__global__ void k_dbg_ec(int op, const void* a, const void* b, void* out, uint64_t n) {
  using E = EC<F>;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> p = reinterpret_cast<const XYZZ<F>*>(a)[i];
  switch (op) {
    case EC_MADD: {
      Affine<F> q = reinterpret_cast<const Affine<F>*>(b)[i];
      E::madd(p, q);
      reinterpret_cast<XYZZ<F>*>(out)[i] = p;
      break;
    }
    case EC_ADD: {
      XYZZ<F> q = reinterpret_cast<const XYZZ<F>*>(b)[i];
      E::add(p, q);
      reinterpret_cast<XYZZ<F>*>(out)[i] = p;
      break;
    }
    case EC_DBL:
      E::dbl(p);
      reinterpret_cast<XYZZ<F>*>(out)[i] = p;
      break;
    case EC_TO_AFFINE: {
      Affine<F> q;
      E::to_affine(q, p);
      reinterpret_cast<Affine<F>*>(out)[i] = q;
      break;
    }
    case EC_MUL_SCALAR: {
      typename Fr::El s = reinterpret_cast<const typename Fr::El*>(b)[i], sc;
      Fr::from_mont(sc, s);
      XYZZ<F> r;
      E::template mul_scalar<Fr::N>(r, p, sc.v);
      reinterpret_cast<XYZZ<F>*>(out)[i] = r;
      break;
    }
  }
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

// ------------------------------------------------------------------------------------ MSM driver
template <class F, class Fr>
void msm_launch(const void* d_points, const void* d_scalars, uint64_t n, void* d_out, MsmWorkspace& ws,
                cudaStream_t s, int c_override, MsmStats* stats) {
  using Pt = XYZZ<F>;
  if (n >= (1ull << 31)) throw std::runtime_error("msm: n must be < 2^31");
  MsmPlan pl = make_msm_plan(n, Fr::BITS, c_override);
  if (stats) *stats = MsmStats{pl.c, pl.nwin, pl.nb, pl.task, pl.group};
  if (n == 0) {
    B200_CUDA(cudaMemsetAsync(d_out, 0, sizeof(Pt), s));
    return;
  }
  const uint64_t total_b = (uint64_t)pl.nwin * pl.nb;
  const uint32_t ngroups = pl.nb / pl.group;
  uint32_t* hist = (uint32_t*)ws.hist.get(total_b * 4);
  uint32_t* off = (uint32_t*)ws.off.get(total_b * 4);
  uint32_t* cur = (uint32_t*)ws.cur.get(total_b * 4);
  uint32_t* sorted = (uint32_t*)ws.sorted.get((uint64_t)pl.nwin * n * 4);
  Pt* buckets = (Pt*)ws.buckets.get(total_b * sizeof(Pt));
  OvfTask* tasks = (OvfTask*)ws.tasks.get((uint64_t)pl.max_ovf * sizeof(OvfTask));
  OvfBucket* obuckets = (OvfBucket*)ws.obuckets.get((uint64_t)pl.max_ovf * sizeof(OvfBucket));
  Pt* partial = (Pt*)ws.partial.get((uint64_t)pl.max_ovf * sizeof(Pt));
  Pt* groups = (Pt*)ws.groups.get((uint64_t)pl.nwin * ngroups * sizeof(Pt));
  Pt* windows = (Pt*)ws.windows.get((uint64_t)pl.nwin * sizeof(Pt));
  OvfCounters* ctr = (OvfCounters*)ws.ctr.get(sizeof(OvfCounters));

  B200_CUDA(cudaMemsetAsync(hist, 0, total_b * 4, s));
  B200_CUDA(cudaMemsetAsync(ctr, 0, sizeof(OvfCounters), s));
  const auto* sc = reinterpret_cast<const typename Fr::El*>(d_scalars);
  const auto* pts = reinterpret_cast<const Affine<F>*>(d_points);
  const unsigned sblocks = (unsigned)((n + 255) / 256);
  k_msm_hist<Fr><<<sblocks, 256, 0, s>>>(sc, pl, hist);
  k_msm_scan<<<pl.nwin, 1024, 0, s>>>(hist, pl, off, cur);
  k_msm_scatter<Fr><<<sblocks, 256, 0, s>>>(sc, pl, cur, sorted);
  k_msm_accumulate<F><<<(unsigned)((total_b + 127) / 128), 128, 0, s>>>(pts, sorted, off, cur, pl, buckets, tasks,
                                                                         obuckets, ctr);
  // oversized buckets (skewed scalar distributions, e.g. the many 1-valued witness wires)
  unsigned ovf_blocks = (unsigned)std::min<uint64_t>((pl.max_ovf + 127) / 128, 148 * 8);
  k_msm_ovf_accumulate<F><<<ovf_blocks, 128, 0, s>>>(pts, sorted, pl, tasks, ctr, partial);
  size_t red_smem = kReduceThreads * sizeof(Pt);
  k_msm_ovf_merge<F><<<(unsigned)std::min<uint64_t>(pl.max_ovf, 1024), kReduceThreads, red_smem, s>>>(obuckets, ctr,
                                                                                                      partial, buckets);
  k_msm_bucket_reduce<F><<<(pl.nwin * ngroups + 63) / 64, 64, 0, s>>>(buckets, pl, groups);
  k_msm_window_sum<F><<<pl.nwin, kReduceThreads, red_smem, s>>>(groups, pl, windows);
  k_msm_horner<F><<<1, 32, 0, s>>>(windows, pl, (Pt*)d_out);
  B200_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------ backend
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
           cudaStream_t s, int c_override, MsmStats* stats) override {
    if (group == 1) msm_launch<G1F, Fr>(d_points, d_scalars, n, d_out, ws, s, c_override, stats);
    else msm_launch<G2F, Fr>(d_points, d_scalars, n, d_out, ws, s, c_override, stats);
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

  void dbg_ec_op(int group, int op, const void* a, const void* b, void* out, uint64_t n, cudaStream_t s) override {
    if (!n) return;
    unsigned blocks = (unsigned)((n + 31) / 32);
    if (group == 1) k_dbg_ec<G1F, Fr><<<blocks, 32, 0, s>>>(op, a, b, out, n);
    else k_dbg_ec<G2F, Fr><<<blocks, 32, 0, s>>>(op, a, b, out, n);
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
