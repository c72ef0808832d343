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

}  // namespace b200
