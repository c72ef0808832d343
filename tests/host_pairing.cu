// CPU harness for davinci-node_b200/csrc/pairing.cuh: the SAME template code the GPU kernels instantiate, compiled for
// the host (B200_PQ = __host__) over a portable 32-bit-limb Montgomery field, so that tests/test_host_pairing.py can
// check every extension-field product, Miller step and the final exponentiation bit for bit against the big-integer
// pairing of the test oracle without a GPU.  Test infrastructure: nothing in the product links this file.
//
// I/O convention of the exported functions: canonical (non-Montgomery) little-endian 32-bit limbs.
#define B200_PAIRING_HOST 1
#define B200_PQ __host__
#include <stdint.h>
#include <string.h>

#include "field.cuh"      // generated parameter structs (modulus / one / r2 / M0 are host-readable)
#include "pairing.cuh"

namespace {

template <class P>
struct HostFp {
  static constexpr int N = P::N;
  struct El {
    uint32_t v[N];
  };
  static bool geq_p(const uint32_t* a) {
    for (int i = N - 1; i >= 0; i--) {
      uint32_t m = P::modulus(i);
      if (a[i] != m) return a[i] > m;
    }
    return true;
  }
  static void sub_p(uint32_t* a) {
    uint64_t bw = 0;
    for (int i = 0; i < N; i++) {
      uint64_t t = (uint64_t)a[i] - P::modulus(i) - bw;
      a[i] = (uint32_t)t;
      bw = (t >> 32) & 1u;
    }
  }
  static void add(El& r, const El& a, const El& b) {
    uint64_t c = 0;
    uint32_t t[N];
    for (int i = 0; i < N; i++) {
      c += (uint64_t)a.v[i] + b.v[i];
      t[i] = (uint32_t)c;
      c >>= 32;
    }
    if (c || geq_p(t)) sub_p(t);     // every modulus here leaves at least one spare bit, c stays 0
    memcpy(r.v, t, sizeof(t));
  }
  static void sub(El& r, const El& a, const El& b) {
    uint64_t bw = 0;
    uint32_t t[N];
    for (int i = 0; i < N; i++) {
      uint64_t d = (uint64_t)a.v[i] - b.v[i] - bw;
      t[i] = (uint32_t)d;
      bw = (d >> 32) & 1u;
    }
    if (bw) {
      uint64_t c = 0;
      for (int i = 0; i < N; i++) {
        c += (uint64_t)t[i] + P::modulus(i);
        t[i] = (uint32_t)c;
        c >>= 32;
      }
    }
    memcpy(r.v, t, sizeof(t));
  }
  // Montgomery product (CIOS, 32-bit limbs)
  static void mul(El& r, const El& a, const El& b) {
    uint32_t t[N + 2];
    memset(t, 0, sizeof(t));
    for (int i = 0; i < N; i++) {
      uint64_t c = 0;
      for (int j = 0; j < N; j++) {
        uint64_t s = (uint64_t)a.v[j] * b.v[i] + t[j] + c;
        t[j] = (uint32_t)s;
        c = s >> 32;
      }
      uint64_t s = (uint64_t)t[N] + c;
      t[N] = (uint32_t)s;
      t[N + 1] = (uint32_t)(s >> 32);
      uint32_t m = t[0] * P::M0;
      s = (uint64_t)m * P::modulus(0) + t[0];
      c = s >> 32;
      for (int j = 1; j < N; j++) {
        s = (uint64_t)m * P::modulus(j) + t[j] + c;
        t[j - 1] = (uint32_t)s;
        c = s >> 32;
      }
      s = (uint64_t)t[N] + c;
      t[N - 1] = (uint32_t)s;
      t[N] = t[N + 1] + (uint32_t)(s >> 32);
    }
    if (t[N] || geq_p(t)) sub_p(t);
    memcpy(r.v, t, N * 4);
  }
  static void set_zero(El& r) { memset(r.v, 0, sizeof(r.v)); }
  static void set_one(El& r) {
    for (int i = 0; i < N; i++) r.v[i] = P::one(i);
  }
  static bool is_zero(const El& a) {
    uint32_t o = 0;
    for (int i = 0; i < N; i++) o |= a.v[i];
    return o == 0;
  }
  static bool eq(const El& a, const El& b) { return memcmp(a.v, b.v, sizeof(a.v)) == 0; }
  static void neg(El& r, const El& a) {
    El z;
    set_zero(z);
    sub(r, z, a);
  }
  static void mul_small(El& r, const El& a, int k) {
    El acc, base = a;
    set_zero(acc);
    while (k) {
      if (k & 1) add(acc, acc, base);
      k >>= 1;
      if (k) add(base, base, base);
    }
    r = acc;
  }
  // a^(p-2); 0 -> 0 (the convention of FpT::inv_bin)
  static void inv_bin(El& r, const El& a) {
    uint32_t e[N];
    uint64_t bw = 2;
    for (int i = 0; i < N; i++) {                // e = p - 2
      uint64_t d = (uint64_t)P::modulus(i) - bw;
      e[i] = (uint32_t)d;
      bw = (d >> 32) & 1u;
    }
    El acc;
    set_one(acc);
    for (int i = N * 32 - 1; i >= 0; i--) {
      mul(acc, acc, acc);
      if ((e[i >> 5] >> (i & 31)) & 1u) mul(acc, acc, a);
    }
    r = acc;
  }
  static void inv(El& r, const El& a) { inv_bin(r, a); }
  static void to_mont(El& r, const uint32_t* canon) {
    El a, r2;
    for (int i = 0; i < N; i++) {
      a.v[i] = canon[i];
      r2.v[i] = P::r2(i);
    }
    mul(r, a, r2);
  }
  static void from_mont(uint32_t* canon, const El& a) {
    El one, t;
    set_zero(one);
    one.v[0] = 1;
    mul(t, a, one);
    memcpy(canon, t.v, sizeof(t.v));
  }
};

template <class FP, class T, class RP, bool FERMAT = false>
struct Harness {
  using F = HostFp<FP>;
  using PT = b200::PairingT<F, T, RP, FERMAT>;
  using El = typename F::El;
  static constexpr int N = FP::N;
  static void load(El* dst, const uint32_t* src, int count) {
    for (int i = 0; i < count; i++) F::to_mont(dst[i], src + (size_t)i * N);
  }
  static void store(uint32_t* dst, const typename PT::Ext& e) {
    for (int i = 0; i < PT::K; i++) F::from_mont(dst + (size_t)i * N, e.c[i]);
  }
  // out: K x N limbs; which: 0 = Miller value, 1 = reduced pairing
  static int pair(const uint32_t* g1, const uint32_t* g2, uint32_t* out, int which) {
    El P[2], Q[2 * PT::NQ];
    load(P, g1, 2);
    load(Q, g2, 2 * PT::NQ);
    typename PT::Ext f, g;
    bool in_subgroup = PT::miller(f, P, Q);
    if (which == 1) {
      PT::final_exp(g, f);
      store(out, g);
    } else {
      store(out, f);
    }
    return in_subgroup ? 1 : 0;
  }
  static int ext_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
    typename PT::Ext x, y, z;
    load(x.c, a, PT::K);
    load(y.c, b, PT::K);
    PT::ext_mul(z, x, y);
    store(out, z);
    return 0;
  }
  // prod_i t(P_i, Q_i) == 1 ; returns 1 / 0, or -1 when some P_i is outside the subgroup
  static int check(const uint32_t* g1, const uint32_t* g2, int n) {
    typename PT::Ext acc, f;
    PT::ext_one(acc);
    bool sub_ok = true;
    for (int i = 0; i < n; i++) {
      El P[2], Q[2 * PT::NQ];
      load(P, g1 + (size_t)i * 2 * N, 2);
      load(Q, g2 + (size_t)i * 2 * PT::NQ * N, 2 * PT::NQ);
      sub_ok = PT::miller(f, P, Q) && sub_ok;
      PT::ext_mul(acc, acc, f);
    }
    if (!sub_ok) return -1;
    PT::final_exp(f, acc);
    return PT::ext_is_one(f) ? 1 : 0;
  }
};

using H_bn254 = Harness<b200::bn254_fp, b200::pairing_bn254, b200::bn254_fr>;
using H_bls12_377 = Harness<b200::bls12_377_fp, b200::pairing_bls12_377, b200::bls12_377_fr>;
using H_bls12_381 = Harness<b200::bls12_381_fp, b200::pairing_bls12_381, b200::bls12_381_fr>;
using H_bw6_761 = Harness<b200::bw6_761_fp, b200::pairing_bw6_761, b200::bw6_761_fr>;
using HF_bn254 = Harness<b200::bn254_fp, b200::pairing_bn254, b200::bn254_fr, true>;   // the batch kernels' variant

}  // namespace

#define DISPATCH(curve, expr)              \
  switch (curve) {                         \
    case 1: { using HH = H_bn254; return expr; }      \
    case 2: { using HH = H_bls12_377; return expr; }  \
    case 3: { using HH = H_bls12_381; return expr; }  \
    case 4: { using HH = H_bw6_761; return expr; }    \
    case 101: { using HH = HF_bn254; return expr; }   \
    default: return -100;                  \
  }

extern "C" {
__attribute__((visibility("default"))) int hp_pair(int curve, const uint32_t* g1, const uint32_t* g2, uint32_t* out, int which) {
  DISPATCH(curve, HH::pair(g1, g2, out, which));
}
__attribute__((visibility("default"))) int hp_ext_mul(int curve, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  DISPATCH(curve, HH::ext_mul(a, b, out));
}
__attribute__((visibility("default"))) int hp_check(int curve, const uint32_t* g1, const uint32_t* g2, int n) {
  DISPATCH(curve, HH::check(g1, g2, n));
}
}
