// gnark-crypto point (de)compression for every curve / group of the path: the on-disk format of the proving keys
// (`pk.WriteTo`, written by /root/reference/cmd/circuit-compile/main.go:507-512 and read back at
// /root/reference/circuits/artifacts.go:391-406) and of the EIP-4844 SRS (config/kzg_trusted_setup.txt).
//
//   big-endian X (G2 over Fp2: X.A1 || X.A0), flag bits in the top of byte 0:
//     BLS12-377 / BLS12-381 / BW6-761 (3 bits): 100 compressed, y smallest; 101 compressed, y largest; 110 infinity
//     BN254 (2 bits):                          10  compressed, y smallest; 11  compressed, y largest; 01  infinity
//   "largest" is lexicographic on the canonical value (Fp2: A1 first, A0 when A1 = 0).
// Loading a 10^7-point key is minutes of square roots on a CPU; here it is one kernel.
#pragma once
#include "ec.cuh"
#include "kzg.cuh"

namespace b200 {

template <class F>
struct CoordSerde;

// coordinates over Fp
template <class P>
struct CoordSerde<FpT<P>> {
  using F = FpT<P>;
  static constexpr int BYTES = 4 * P::N;
  static __device__ __forceinline__ void read(typename F::El& x, const uint8_t* b) {
    be_bytes_to_limbs<P::N>(x.v, b);
    F::to_mont(x, x);
  }
  static __device__ __forceinline__ void write(uint8_t* b, const typename F::El& x) {
    typename F::El c;
    F::from_mont(c, x);
    limbs_to_be_bytes<P::N>(b, c.v);
  }
  static __device__ __forceinline__ bool largest(const typename F::El& y) {
    typename F::El c;
    F::from_mont(c, y);
    return lexicographically_largest<F>(c);
  }
};

// coordinates over Fp2: A1 first
template <class P, int NRN>
struct CoordSerde<Fp2T<P, NRN>> {
  using F = Fp2T<P, NRN>;
  using B = FpT<P>;
  static constexpr int BYTES = 8 * P::N;
  static __device__ __forceinline__ void read(typename F::El& x, const uint8_t* b) {
    be_bytes_to_limbs<P::N>(x.c1.v, b);
    be_bytes_to_limbs<P::N>(x.c0.v, b + 4 * P::N);
    B::to_mont(x.c1, x.c1);
    B::to_mont(x.c0, x.c0);
  }
  static __device__ __forceinline__ void write(uint8_t* b, const typename F::El& x) {
    typename B::El c;
    B::from_mont(c, x.c1);
    limbs_to_be_bytes<P::N>(b, c.v);
    B::from_mont(c, x.c0);
    limbs_to_be_bytes<P::N>(b + 4 * P::N, c.v);
  }
  static __device__ __forceinline__ bool largest(const typename F::El& y) {
    typename B::El c;
    B::from_mont(c, y.c1);
    if (!B::is_zero(c)) return lexicographically_largest<B>(c);
    B::from_mont(c, y.c0);
    return lexicographically_largest<B>(c);
  }
};

struct PointFlags {
  uint8_t mask, smallest, largest, infinity;
};
__host__ __device__ inline PointFlags point_flags(int flag_bits) {
  return flag_bits == 2 ? PointFlags{0xc0, 0x80, 0xc0, 0x40} : PointFlags{0xe0, 0x80, 0xa0, 0xc0};
}

// err bit 0: not a compressed encoding, bit 1: x is not the abscissa of a curve point
template <class F>
__global__ void __launch_bounds__(64)
k_points_decompress(const uint8_t* __restrict__ in, Affine<F>* __restrict__ out, uint64_t n, int flag_bits,
                    const typename F::El* __restrict__ curve_b, uint32_t* err) {
  using S = CoordSerde<F>;
  using El = typename F::El;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PointFlags pf = point_flags(flag_bits);
  uint8_t buf[S::BYTES];
  for (int k = 0; k < S::BYTES; k++) buf[k] = in[i * S::BYTES + k];
  const uint8_t flags = buf[0] & pf.mask;
  buf[0] &= (uint8_t)~pf.mask;
  Affine<F> r;
  F::set_zero(r.x);
  F::set_zero(r.y);
  const bool compressed = flags == pf.smallest || flags == pf.largest;
  // the square root is an out-of-line call: entered by every thread of the converged group when any needs it
  if (__any_sync(__activemask(), compressed)) {
    El x, y2, y, ny;
    S::read(x, buf);
    F::sqr(y2, x);
    F::mul(y2, y2, x);
    F::add(y2, y2, *curve_b);
    const bool found = F::sqrt(y, y2);
    F::neg(ny, y);
    const bool flip = S::largest(y) != (flags == pf.largest);
    if (compressed && found) {
      r.x = x;
      r.y = flip ? ny : y;
    }
    if (compressed && !found) atomicOr(err, 2u);
  }
  if (!compressed && flags != pf.infinity) atomicOr(err, 1u);
  store16(out + i, r);
}

template <class F>
__global__ void __launch_bounds__(64)
k_points_compress(const Affine<F>* __restrict__ in, uint8_t* __restrict__ out, uint64_t n, int flag_bits) {
  using S = CoordSerde<F>;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PointFlags pf = point_flags(flag_bits);
  Affine<F> p;
  load16(p, in + i);
  uint8_t buf[S::BYTES];
  if (EC<F>::is_inf(p)) {
    for (int k = 0; k < S::BYTES; k++) buf[k] = 0;
    buf[0] = pf.infinity;
  } else {
    S::write(buf, p.x);
    buf[0] |= S::largest(p.y) ? pf.largest : pf.smallest;
  }
  for (int k = 0; k < S::BYTES; k++) out[i * S::BYTES + k] = buf[k];
}

}  // namespace b200
