// EIP-4844 blob KZG commitment helpers (BLS12-381): SRS point decompression, blob -> scalars,
// commitment compression.  The commitment itself is the generic Pippenger MSM (msm.cuh) over the
// 4096 Lagrange-basis points with the bit-reversal permutation as its index map.
//
// Replaces go-ethereum `kzg4844.BlobToCommitment` as called by
// /root/reference/types/blobs.go:90-96 (state/blobs.go:99, crypto/blobs/blob.go:41).
#pragma once
#include "ec.cuh"

namespace b200 {

// big-endian bytes -> little-endian limbs (NB = 4*N bytes)
template <int N>
__device__ __forceinline__ void be_bytes_to_limbs(uint32_t* v, const uint8_t* b) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    const uint8_t* q = b + 4 * (N - 1 - k);
    v[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
}
template <int N>
__device__ __forceinline__ void limbs_to_be_bytes(uint8_t* b, const uint32_t* v) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    uint8_t* q = b + 4 * (N - 1 - k);
    q[0] = (uint8_t)(v[k] >> 24);
    q[1] = (uint8_t)(v[k] >> 16);
    q[2] = (uint8_t)(v[k] >> 8);
    q[3] = (uint8_t)v[k];
  }
}

// canonical a > (p-1)/2 ?
template <class Fp>
__device__ __forceinline__ bool lexicographically_largest(const typename Fp::El& canon) {
  using P = typename Fp::Params;
  // half = (p - 1) / 2 = p >> 1 (p odd)
  for (int i = Fp::N - 1; i >= 0; i--) {
    uint32_t h = (P::modulus(i) >> 1) | (i + 1 < Fp::N ? (P::modulus(i + 1) << 31) : 0u);
    if (canon.v[i] > h) return true;
    if (canon.v[i] < h) return false;
  }
  return false;
}

// ZCash/IETF compressed G1 (48 bytes: flags 0x80 compressed, 0x40 infinity, 0x20 y-largest) -> affine.
// err bit 0: bad encoding, bit 1: x not on curve.
template <class Fp>
__global__ void k_g1_decompress_bls(const uint8_t* __restrict__ in, Affine<Fp>* __restrict__ out, uint32_t n,
                                    int curve_b, uint32_t* err) {
  using El = typename Fp::El;
  using P = typename Fp::Params;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[4 * Fp::N];
  for (int k = 0; k < 4 * Fp::N; k++) buf[k] = in[(uint64_t)i * 4 * Fp::N + k];
  uint8_t flags = buf[0] & 0xe0;
  buf[0] &= 0x1f;
  Affine<Fp> r;
  Fp::set_zero(r.x);
  Fp::set_zero(r.y);
  if (!(flags & 0x80)) {
    atomicOr(err, 1u);
    out[i] = r;
    return;
  }
  if (flags & 0x40) {
    out[i] = r;
    return;
  }
  El x, y2, y, t;
  be_bytes_to_limbs<Fp::N>(x.v, buf);
  Fp::to_mont(x, x);
  Fp::sqr(y2, x);
  Fp::mul(y2, y2, x);
  El bb;
  Fp::set_one(t);
  Fp::set_zero(bb);
  for (int k = 0; k < curve_b; k++) Fp::add(bb, bb, t);
  Fp::add(y2, y2, bb);
  // y = y2^((p+1)/4)   (p = 3 mod 4)
  uint32_t e[Fp::N];
  uint32_t carry = 1;
  for (int k = 0; k < Fp::N; k++) {
    uint64_t s = (uint64_t)P::modulus(k) + carry;
    e[k] = (uint32_t)s;
    carry = (uint32_t)(s >> 32);
  }
  for (int k = 0; k < Fp::N; k++) e[k] = (e[k] >> 2) | (k + 1 < Fp::N ? (e[k + 1] << 30) : 0u);
  Fp::template pow<Fp::N>(y, y2, e);
  Fp::sqr(t, y);
  if (!Fp::eq(t, y2)) {
    atomicOr(err, 2u);
    out[i] = r;
    return;
  }
  El yc;
  Fp::from_mont(yc, y);
  bool largest = lexicographically_largest<Fp>(yc);
  if (largest != ((flags & 0x20) != 0)) Fp::neg(y, y);
  r.x = x;
  r.y = y;
  out[i] = r;
}

template <class Fp>
__global__ void k_g1_compress_bls(const Affine<Fp>* __restrict__ in, uint8_t* __restrict__ out, uint32_t n) {
  using El = typename Fp::El;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<Fp> p = in[i];
  uint8_t* o = out + (uint64_t)i * 4 * Fp::N;
  if (EC<Fp>::is_inf(p)) {
    for (int k = 0; k < 4 * Fp::N; k++) o[k] = 0;
    o[0] = 0xc0;
    return;
  }
  El xc, yc;
  Fp::from_mont(xc, p.x);
  Fp::from_mont(yc, p.y);
  limbs_to_be_bytes<Fp::N>(o, xc.v);
  o[0] |= 0x80;
  if (lexicographically_largest<Fp>(yc)) o[0] |= 0x20;
}

// blob: n big-endian 32-byte canonical scalars -> Montgomery Fr ; err bit 2 when a scalar >= r
template <class Fr>
__global__ void k_blob_to_scalars(const uint8_t* __restrict__ blob, typename Fr::El* __restrict__ out, uint32_t n,
                                  uint32_t* err) {
  using P = typename Fr::Params;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename Fr::El v;
  uint8_t buf[4 * Fr::N];
  for (int k = 0; k < 4 * Fr::N; k++) buf[k] = blob[(uint64_t)i * 4 * Fr::N + k];
  be_bytes_to_limbs<Fr::N>(v.v, buf);
  bool lt = false;   // v < r ?
  for (int k = Fr::N - 1; k >= 0; k--) {
    if (v.v[k] < P::modulus(k)) {
      lt = true;
      break;
    }
    if (v.v[k] > P::modulus(k)) break;
  }
  if (!lt) atomicOr(err, 4u);
  Fr::to_mont(v, v);
  out[i] = v;
}

// ------------------------------------------------------------------------------------ opening proofs
// EIP-4844 compute_kzg_proof_impl on the device: the blob is p's evaluations over the n-th roots of unity in
// bit-reversed order; proof = MSM(q, Lagrange SRS) with q_i = (p_i - y) / (w_i - z), y = p(z) by the
// barycentric formula.  Replaces go-ethereum `kzg4844.ComputeProof` / `ComputeBlobProof`
// (/root/reference/types/blobs.go:107-134).
constexpr int kKzgThreads = 128;

// roots[i] = w^brp(i), w = (2^32-th root of unity of BLS12-381 fr, crypto/blobs/barycentric.go:52)^(2^(32-logn))
template <class Fr>
__global__ void k_kzg_roots(typename Fr::El* __restrict__ roots, uint32_t n, int logn) {
  using El = typename Fr::El;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr uint32_t g32[8] = {0x439f0d2bu, 0x3829971fu, 0x8c2280b9u, 0xb6368350u,
                               0x22c813b4u, 0xd09b6819u, 0xdfe81f20u, 0x16a2a19eu};
  El w, acc;
#pragma unroll
  for (int k = 0; k < 8; k++) w.v[k] = g32[k];
  Fr::to_mont(w, w);
  for (int k = 0; k < 32 - logn; k++) Fr::sqr(w, w);
  uint32_t e = logn ? (__brev(i) >> (32 - logn)) : 0;
  Fr::set_one(acc);
  while (e) {
    if (e & 1) Fr::mul(acc, acc, w);
    e >>= 1;
    if (e) Fr::sqr(w, w);
  }
  roots[i] = acc;
}

// block-wide sum of Fr elements; result valid in thread 0
template <class Fr>
__device__ __forceinline__ void kzg_block_sum(typename Fr::El& v, typename Fr::El* sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = kKzgThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      Fr::add(v, v, sm[threadIdx.x + s]);
      sm[threadIdx.x] = v;
    }
    __syncthreads();
  }
}

// z: 32 big-endian bytes -> Montgomery; err bit 3 when z >= r
template <class Fr>
__global__ void k_kzg_load_point(const uint8_t* __restrict__ z_be, typename Fr::El* __restrict__ z_out, uint32_t* hit,
                                 uint32_t* err) {
  using P = typename Fr::Params;
  typename Fr::El v;
  uint8_t buf[32];
  for (int k = 0; k < 32; k++) buf[k] = z_be[k];
  be_bytes_to_limbs<Fr::N>(v.v, buf);
  bool lt = false;
  for (int k = Fr::N - 1; k >= 0; k--) {
    if (v.v[k] < P::modulus(k)) {
      lt = true;
      break;
    }
    if (v.v[k] > P::modulus(k)) break;
  }
  if (!lt) atomicOr(err, 8u);
  Fr::to_mont(v, v);
  *z_out = v;
  *hit = 0xffffffffu;
}

// inv[i] = 1/(z - w_i) (0 when z == w_i, whose index goes to *hit); partial[block] = sum p_i w_i inv_i
template <class Fr>
__global__ void __launch_bounds__(kKzgThreads)
k_kzg_open_terms(const typename Fr::El* __restrict__ p, const typename Fr::El* __restrict__ roots,
                 const typename Fr::El* __restrict__ z, typename Fr::El* __restrict__ inv,
                 typename Fr::El* __restrict__ partial, uint32_t* __restrict__ hit, uint32_t n) {
  using El = typename Fr::El;
  __shared__ El sm[kKzgThreads];
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  El term;
  Fr::set_zero(term);
  if (i < n) {
    El d, w = roots[i], iv;
    Fr::sub(d, *z, w);
    if (Fr::is_zero(d)) atomicMin(hit, i);
    Fr::inv(iv, d);
    inv[i] = iv;
    Fr::mul(term, p[i], w);
    Fr::mul(term, term, iv);
  }
  kzg_block_sum<Fr>(term, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = term;
}

// y = p(z): p[hit] inside the domain, else (z^n - 1)/n * sum of the partials.  Also as 32 big-endian bytes.
template <class Fr>
__global__ void k_kzg_open_y(const typename Fr::El* __restrict__ partial, uint32_t nblocks,
                             const typename Fr::El* __restrict__ p, const typename Fr::El* __restrict__ z,
                             const uint32_t* __restrict__ hit, uint32_t n, int logn, typename Fr::El* __restrict__ y_out,
                             uint8_t* __restrict__ y_bytes) {
  using El = typename Fr::El;
  if (threadIdx.x || blockIdx.x) return;
  El y;
  if (*hit != 0xffffffffu) {
    y = p[*hit];
  } else {
    El s, zn = *z, one, nn, ninv;
    Fr::set_zero(s);
    for (uint32_t k = 0; k < nblocks; k++) Fr::add(s, s, partial[k]);
    for (int k = 0; k < logn; k++) Fr::sqr(zn, zn);
    Fr::set_one(one);
    Fr::sub(zn, zn, one);
    nn = one;
    for (int k = 0; k < logn; k++) Fr::dbl(nn, nn);
    Fr::inv(ninv, nn);
    Fr::mul(y, s, zn);
    Fr::mul(y, y, ninv);
  }
  *y_out = y;
  El yc;
  Fr::from_mont(yc, y);
  uint8_t buf[32];
  limbs_to_be_bytes<Fr::N>(buf, yc.v);
  for (int k = 0; k < 32; k++) y_bytes[k] = buf[k];
}

// q_i = (p_i - y) / (w_i - z) = -(p_i - y) inv_i ; partial2[block] = sum (p_i - y) w_i inv_i
template <class Fr>
__global__ void __launch_bounds__(kKzgThreads)
k_kzg_open_quotient(const typename Fr::El* __restrict__ p, const typename Fr::El* __restrict__ roots,
                    const typename Fr::El* __restrict__ inv, const typename Fr::El* __restrict__ y,
                    typename Fr::El* __restrict__ q, typename Fr::El* __restrict__ partial2, uint32_t n) {
  using El = typename Fr::El;
  __shared__ El sm[kKzgThreads];
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  El term;
  Fr::set_zero(term);
  if (i < n) {
    El t, u;
    Fr::sub(t, p[i], *y);
    Fr::mul(u, t, inv[i]);
    Fr::mul(term, u, roots[i]);
    Fr::neg(u, u);
    q[i] = u;
  }
  kzg_block_sum<Fr>(term, sm);
  if (threadIdx.x == 0) partial2[blockIdx.x] = term;
}

// z inside the domain (z = w_hit): q_hit = (1/z) sum_{i != hit} (p_i - y) w_i / (z - w_i)
// (EIP-4844 compute_quotient_eval_within_domain; the i = hit term of the partial sums is zero because inv_hit = 0)
template <class Fr>
__global__ void k_kzg_open_fix(const typename Fr::El* __restrict__ partial2, uint32_t nblocks,
                               const typename Fr::El* __restrict__ z, const uint32_t* __restrict__ hit,
                               typename Fr::El* __restrict__ q) {
  using El = typename Fr::El;
  if (threadIdx.x || blockIdx.x) return;
  if (*hit == 0xffffffffu) return;
  El s, zi;
  Fr::set_zero(s);
  for (uint32_t k = 0; k < nblocks; k++) Fr::add(s, s, partial2[k]);
  Fr::inv(zi, *z);
  Fr::mul(s, s, zi);
  q[*hit] = s;
}

// ------------------------------------------------------------------------------------ cell proofs (EIP-7594)
// proof_k = MSM(q_k, monomial SRS), q_k = p(X) div (X^m - a_k), a_k = h_k^m, h_k = w_2n^brp(m k): replaces go-ethereum
// `kzg4844.ComputeCellProofs` (/root/reference/types/blobs.go:99-105).
// shifts[k] = a_k for k < ncells  (m = cell size, a power of two; n = blob size)
template <class Fr>
__global__ void k_kzg_cell_shifts(typename Fr::El* __restrict__ shifts, uint32_t ncells, uint32_t m, int log_ext) {
  using El = typename Fr::El;
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncells) return;
  constexpr uint32_t g32[8] = {0x439f0d2bu, 0x3829971fu, 0x8c2280b9u, 0xb6368350u,
                               0x22c813b4u, 0xd09b6819u, 0xdfe81f20u, 0x16a2a19eu};
  El w, acc;
#pragma unroll
  for (int i = 0; i < 8; i++) w.v[i] = g32[i];
  Fr::to_mont(w, w);
  for (int i = 0; i < 32 - log_ext; i++) Fr::sqr(w, w);      // primitive 2n-th root
  uint32_t e = __brev(m * k) >> (32 - log_ext);
  Fr::set_one(acc);
  while (e) {
    if (e & 1) Fr::mul(acc, acc, w);
    e >>= 1;
    if (e) Fr::sqr(w, w);
  }
  for (uint32_t t = 1; t < m; t <<= 1) Fr::sqr(acc, acc);    // h^m
  shifts[k] = acc;
}

// thread (k, r): the residue class r mod m of quotient k, from the top down: q[j] = c[j + m] + a_k q[j + m]
template <class Fr>
__global__ void k_kzg_cell_quotients(const typename Fr::El* __restrict__ coeffs, const typename Fr::El* __restrict__ shifts,
                                     typename Fr::El* __restrict__ q, uint32_t n, uint32_t m, uint32_t ncells) {
  using El = typename Fr::El;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncells * m) return;
  const uint32_t k = t / m, r = t % m;
  const El a = shifts[k];
  El* out = q + (uint64_t)k * n;
  El run, zero;
  Fr::set_zero(run);
  Fr::set_zero(zero);
  for (int64_t j = (int64_t)n - m + r; j >= (int64_t)(n - m); j -= m) out[j] = zero;   // degree >= n - m: zero
  for (int64_t j = (int64_t)n - 2 * m + r; j >= 0; j -= m) {
    El c = coeffs[j + m];
    Fr::mul(run, run, a);
    Fr::add(run, run, c);
    out[j] = run;
  }
}

}  // namespace b200
