// Field layer for the sm_100a Groth16 backend.
//
// The limb-level arithmetic is generated (tools/gen_field.py -> gen/field_*.cuh): one inline-PTX
// block per operation, IMAD.WIDE carry chains, modulus as immediates.  This header wraps a generated
// parameter struct P into the "field concept" the EC / MSM / NTT templates consume:
//
//   F::El                      element type (plain limbs, gnark-crypto Montgomery layout)
//   F::add/sub/mul/sqr/dbl/neg (El& r, const El& a[, const El& b])     r may alias a/b
//   F::is_zero / eq / set_zero / set_one
//   F::LIMBS                   uint32 words per element
//
// Replaces gnark-crypto `fp.Element` / `fr.Element` / `E2` arithmetic reached from
// /root/reference/prover/prover_cpu.go:37 (groth16.Prove).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include "gen/field_bn254_fp.cuh"
#include "gen/field_bn254_fr.cuh"
#include "gen/field_bls12_377_fp.cuh"
#include "gen/field_bls12_377_fr.cuh"
#include "gen/field_bls12_381_fp.cuh"
#include "gen/field_bls12_381_fr.cuh"
#include "gen/field_bw6_761_fp.cuh"

namespace b200 {

using bw6_761_fr = bls12_377_fp;   // BW6-761's scalar field is BLS12-377's base field (2-chain)

template <int N>
struct alignas(16) Limbs {
  uint32_t v[N];
};

// 128-bit vectorised global load / store of POD structs whose size is a multiple of 16 bytes
template <class T>
__device__ __forceinline__ void load16(T& dst, const T* src) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple expected");
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(&dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
}
template <class T>
__device__ __forceinline__ void load16_rw(T& dst, const T* src) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(&dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}
template <class T>
__device__ __forceinline__ void store16(T* dst, const T& src) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple expected");
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4* s = reinterpret_cast<const uint4*>(&src);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}


// ---------------------------------------------------------------------------------------- Fp
template <class P>
struct FpT {
  using Params = P;
  static constexpr int N = P::N;
  static constexpr int LIMBS = P::N;
  static constexpr int BASE_N = P::N;      // limbs of the prime field the arithmetic bottoms out in
  static constexpr int BITS = P::BITS;
  using El = Limbs<N>;

  static __device__ __forceinline__ void add(El& r, const El& a, const El& b) { P::add(r.v, a.v, b.v); }
  static __device__ __forceinline__ void sub(El& r, const El& a, const El& b) { P::sub(r.v, a.v, b.v); }
  static __device__ __forceinline__ void mul(El& r, const El& a, const El& b) { P::mul(r.v, a.v, b.v); }
  static __device__ __forceinline__ void sqr(El& r, const El& a) { P::sqr(r.v, a.v); }
  static __device__ __forceinline__ void dbl(El& r, const El& a) { P::add(r.v, a.v, a.v); }
  static __device__ __forceinline__ void from_mont(El& r, const El& a) { P::from_mont(r.v, a.v); }
  static __device__ __forceinline__ void to_mont(El& r, const El& a) {
    El r2;
#pragma unroll
    for (int i = 0; i < N; i++) r2.v[i] = P::r2(i);
    P::mul(r.v, a.v, r2.v);
  }
  static __device__ __forceinline__ void set_zero(El& r) {
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
  }
  static __device__ __forceinline__ void set_one(El& r) {
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::one(i);
  }
  static __device__ __forceinline__ bool is_zero(const El& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= a.v[i];
    return o == 0;
  }
  static __device__ __forceinline__ bool eq(const El& a, const El& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
  }
  static __device__ __forceinline__ void neg(El& r, const El& a) {
    El z;
    set_zero(z);
    P::sub(r.v, z.v, a.v);   // 0 - a mod p  (gives 0 for a == 0)
  }
  // r = a * k for a small non-negative compile-time-ish integer (used for curve constants 3, 5 ...)
  static __device__ __forceinline__ void mul_small(El& r, const El& a, int k) {
    El acc, base = a;
    set_zero(acc);
    while (k) {
      if (k & 1) add(acc, acc, base);
      k >>= 1;
      if (k) dbl(base, base);
    }
    r = acc;
  }
  // a^e for a canonical little-endian exponent of NE limbs (square-and-multiply, MSB first)
  template <int NE>
  static __device__ __noinline__ void pow(El& r, const El& a, const uint32_t* e) {
    El acc;
    set_one(acc);
    bool started = false;
    for (int i = NE * 32 - 1; i >= 0; i--) {
      if (started) sqr(acc, acc);
      if ((e[i >> 5] >> (i & 31)) & 1) {
        mul(acc, acc, a);
        started = true;
      }
    }
    r = acc;
  }
  // Tonelli-Shanks square root (any odd p; one exponentiation when p = 3 mod 4).  Returns false - and leaves r
  // unspecified - when a is not a square.  Used by point decompression.
  static __device__ __noinline__ bool sqrt(El& r, const El& a) {
    if (is_zero(a)) {
      set_zero(r);
      return true;
    }
    uint32_t e[N];
#pragma unroll
    for (int i = 0; i < N; i++) e[i] = P::sqrt_e(i);
    El w, x, b, z, one;
    set_one(one);
    pow<N>(w, a, e);          // a^((T-1)/2)
    mul(x, a, w);             // a^((T+1)/2)
    mul(b, x, w);             // a^T
#pragma unroll
    for (int i = 0; i < N; i++) z.v[i] = P::sqrt_z(i);
    int rr = P::TWO_ADICITY;
    while (!eq(b, one)) {
      int m = 0;
      El t = b;
      while (!eq(t, one) && m < rr) {
        sqr(t, t);
        m++;
      }
      if (m >= rr) return false;
      El g = z;
      for (int k = 0; k < rr - m - 1; k++) sqr(g, g);
      mul(x, x, g);
      sqr(z, g);
      mul(b, b, z);
      rr = m;
    }
    r = x;
    return true;
  }
  // a^-1 by the binary extended Euclidean algorithm on the canonical value (~2 log2 p iterations of N-limb shifts and
  // subtractions): ~10x less latency than the Fermat exponentiation for ONE thread (the proof's final affine
  // normalisation); data-dependent loops, so not for warps of independent elements.  0 -> 0.
  static __device__ __noinline__ void inv_bin(El& r, const El& a) {
    uint32_t u[N], v[N], x1[N], x2[N];
    El ac;
    from_mont(ac, a);
    bool zero = true;
    for (int i = 0; i < N; i++) {
      u[i] = ac.v[i];
      v[i] = P::modulus(i);
      x1[i] = i == 0;
      x2[i] = 0;
      zero = zero && u[i] == 0;
    }
    if (zero) {
      set_zero(r);
      return;
    }
    auto is_one = [](const uint32_t* w) {
      uint32_t o = w[0] ^ 1u;
      for (int i = 1; i < N; i++) o |= w[i];
      return o == 0;
    };
    auto shr1 = [](uint32_t* w) {
      for (int i = 0; i < N - 1; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
      w[N - 1] >>= 1;
    };
    auto add_p = [](uint32_t* w) {           // w += p  (w < p, 2p < 2^(32N) for every field here)
      uint64_t c = 0;
      for (int i = 0; i < N; i++) {
        c += (uint64_t)w[i] + P::modulus(i);
        w[i] = (uint32_t)c;
        c >>= 32;
      }
    };
    auto geq = [](const uint32_t* a_, const uint32_t* b_) {
      for (int i = N - 1; i >= 0; i--) {
        if (a_[i] != b_[i]) return a_[i] > b_[i];
      }
      return true;
    };
    auto sub = [](uint32_t* a_, const uint32_t* b_) {     // a -= b, returns the borrow
      uint64_t bw = 0;
      for (int i = 0; i < N; i++) {
        uint64_t t = (uint64_t)a_[i] - b_[i] - bw;
        a_[i] = (uint32_t)t;
        bw = (t >> 32) & 1u;
      }
      return (uint32_t)bw;
    };
    while (!is_one(u) && !is_one(v)) {
      while (!(u[0] & 1u)) {
        shr1(u);
        if (x1[0] & 1u) add_p(x1);
        shr1(x1);
      }
      while (!(v[0] & 1u)) {
        shr1(v);
        if (x2[0] & 1u) add_p(x2);
        shr1(x2);
      }
      if (geq(u, v)) {
        sub(u, v);
        if (sub(x1, x2)) add_p(x1);          // x1 = x1 - x2 mod p
      } else {
        sub(v, u);
        if (sub(x2, x1)) add_p(x2);
      }
    }
    El out;
    const uint32_t* res = is_one(u) ? x1 : x2;
    for (int i = 0; i < N; i++) out.v[i] = res[i];
    to_mont(r, out);
  }
  // a^-1 = a^(p-2); 0 -> 0
  static __device__ __noinline__ void inv(El& r, const El& a) {
    uint32_t e[N];
    uint32_t borrow = 2;   // e = p - 2 (low limb is 1 for the BLS12-377 fields and BLS12-381 fr)
#pragma unroll
    for (int i = 0; i < N; i++) {
      uint32_t m = P::modulus(i);
      e[i] = m - borrow;
      borrow = (m < borrow) ? 1u : 0u;
    }
    pow<N>(r, a, e);
  }
};

// ---------------------------------------------------------------------------------------- Fp2
// Fp2 = Fp[u] / (u^2 + NR_NEG)   i.e. u^2 = -NR_NEG   (1 for BN254 / BLS12-381, 5 for BLS12-377)
template <class P, int NR_NEG>
struct Fp2T {
  using Base = FpT<P>;
  using BEl = typename Base::El;
  static constexpr int LIMBS = 2 * P::N;
  static constexpr int BASE_N = P::N;
  struct alignas(16) El {
    BEl c0, c1;
  };

  static __device__ __forceinline__ void add(El& r, const El& a, const El& b) {
    Base::add(r.c0, a.c0, b.c0);
    Base::add(r.c1, a.c1, b.c1);
  }
  static __device__ __forceinline__ void sub(El& r, const El& a, const El& b) {
    Base::sub(r.c0, a.c0, b.c0);
    Base::sub(r.c1, a.c1, b.c1);
  }
  static __device__ __forceinline__ void dbl(El& r, const El& a) {
    Base::dbl(r.c0, a.c0);
    Base::dbl(r.c1, a.c1);
  }
  static __device__ __forceinline__ void neg(El& r, const El& a) {
    Base::neg(r.c0, a.c0);
    Base::neg(r.c1, a.c1);
  }
  static __device__ __forceinline__ void mul_nr_neg(BEl& r, const BEl& a) {   // r = NR_NEG * a
    if (NR_NEG == 1) {
      r = a;
    } else {   // 5a = 4a + a
      BEl t;
      Base::dbl(t, a);
      Base::dbl(t, t);
      Base::add(r, t, a);
    }
  }
  // Karatsuba: 3 base multiplications
  static __device__ __noinline__ void mul(El& r, const El& a, const El& b) {
    BEl t0, t1, sa, sb;
    Base::mul(t0, a.c0, b.c0);
    Base::mul(t1, a.c1, b.c1);
    Base::add(sa, a.c0, a.c1);
    Base::add(sb, b.c0, b.c1);
    Base::mul(sa, sa, sb);
    Base::sub(sa, sa, t0);
    Base::sub(r.c1, sa, t1);
    mul_nr_neg(t1, t1);
    Base::sub(r.c0, t0, t1);
  }
  static __device__ __noinline__ void sqr(El& r, const El& a) {
    // c1 = 2 a0 a1 ; c0 = a0^2 - NR_NEG a1^2 = (a0 + a1)(a0 - NR_NEG a1) + (NR_NEG - 1) a0 a1
    BEl m, s, d, t;
    Base::mul(m, a.c0, a.c1);
    Base::add(s, a.c0, a.c1);
    mul_nr_neg(t, a.c1);
    Base::sub(d, a.c0, t);
    Base::mul(s, s, d);
    if (NR_NEG == 1) {
      r.c0 = s;
    } else {
      Base::dbl(t, m);
      Base::dbl(t, t);   // 4 a0 a1
      Base::add(r.c0, s, t);
    }
    Base::dbl(r.c1, m);
  }
  static __device__ __forceinline__ void set_zero(El& r) {
    Base::set_zero(r.c0);
    Base::set_zero(r.c1);
  }
  static __device__ __forceinline__ void set_one(El& r) {
    Base::set_one(r.c0);
    Base::set_zero(r.c1);
  }
  static __device__ __forceinline__ bool is_zero(const El& a) { return Base::is_zero(a.c0) && Base::is_zero(a.c1); }
  static __device__ __forceinline__ bool eq(const El& a, const El& b) { return Base::eq(a.c0, b.c0) && Base::eq(a.c1, b.c1); }
  static __device__ __forceinline__ void mul_small(El& r, const El& a, int k) {
    Base::mul_small(r.c0, a.c0, k);
    Base::mul_small(r.c1, a.c1, k);
  }
  // square root in Fp2 through the norm (u^2 = -NR_NEG is a non-residue of Fp); false when a is not a square.
  // Every branch around the out-of-line base-field calls is a warp vote and each thread selects its own result
  // (control-flow rule in ec.cuh).
  static __device__ __forceinline__ void bsel(BEl& dst, const BEl& src, bool take) {
#pragma unroll
    for (int i = 0; i < P::N; i++) dst.v[i] = take ? src.v[i] : dst.v[i];
  }
  static __device__ __noinline__ bool sqrt(El& r, const El& a) {
    const unsigned grp = __activemask();
    const bool real_only = Base::is_zero(a.c1);
    bool ok = false;
    El out;
    set_zero(out);
    if (__any_sync(grp, !real_only)) {
      BEl half, two, t, n, s, x0, x0b, x1;
      Base::sqr(n, a.c0);
      Base::sqr(t, a.c1);
      mul_nr_neg(t, t);
      Base::add(n, n, t);                    // norm = a0^2 + NR_NEG a1^2
      bool have = Base::sqrt(s, n);
      Base::set_one(two);
      Base::dbl(two, two);
      Base::inv(half, two);
      Base::add(t, a.c0, s);
      Base::mul(t, t, half);                 // (a0 + s) / 2
      const bool first = Base::sqrt(x0, t);
      if (__any_sync(grp, have && !first)) {
        Base::sub(t, a.c0, s);
        Base::mul(t, t, half);               // (a0 - s) / 2
        const bool second = Base::sqrt(x0b, t);
        bsel(x0, x0b, !first);
        have = have && (first || second);
      }
      Base::dbl(t, x0);
      Base::inv(t, t);
      Base::mul(x1, a.c1, t);                // a1 / (2 x0)
      El cand, chk;
      cand.c0 = x0;
      cand.c1 = x1;
      sqr(chk, cand);
      const bool good = !real_only && have && eq(chk, a);
      bsel(out.c0, cand.c0, good);
      bsel(out.c1, cand.c1, good);
      ok = ok || good;
    }
    if (__any_sync(grp, real_only)) {
      BEl t, x, xi, zero;
      Base::set_zero(zero);
      const bool direct = Base::sqrt(x, a.c0);
      // sqrt(a0) = u sqrt(a0 / u^2) = u sqrt(-a0 / NR_NEG)
      Base::set_one(t);
      mul_nr_neg(t, t);
      Base::inv(t, t);
      Base::mul(t, t, a.c0);
      Base::neg(t, t);
      const bool imag = Base::sqrt(xi, t);
      bsel(out.c0, x, real_only && direct);
      bsel(out.c1, zero, real_only && direct);
      bsel(out.c0, zero, real_only && !direct && imag);
      bsel(out.c1, xi, real_only && !direct && imag);
      ok = ok || (real_only && (direct || imag));
    }
    r = out;
    return ok;
  }
  static __device__ __noinline__ void inv_bin(El& r, const El& a) {
    BEl n, t;
    Base::sqr(n, a.c0);
    Base::sqr(t, a.c1);
    mul_nr_neg(t, t);
    Base::add(n, n, t);
    Base::inv_bin(n, n);
    Base::mul(r.c0, a.c0, n);
    Base::mul(t, a.c1, n);
    Base::neg(r.c1, t);
  }
  static __device__ __noinline__ void inv(El& r, const El& a) {
    // 1/(a0 + a1 u) = (a0 - a1 u) / (a0^2 + NR_NEG a1^2)
    BEl n, t;
    Base::sqr(n, a.c0);
    Base::sqr(t, a.c1);
    mul_nr_neg(t, t);
    Base::add(n, n, t);
    Base::inv(n, n);
    Base::mul(r.c0, a.c0, n);
    Base::mul(t, a.c1, n);
    Base::neg(r.c1, t);
  }
};

}  // namespace b200
