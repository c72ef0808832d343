// CPU restatement of the Groth16 proving hot path (test infrastructure: ORACLE, not product).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  It restates, independently of the CUDA code (64-bit limbs + unsigned __int128 CIOS
// Montgomery, Jacobian-extended buckets, OpenMP across windows), the algorithms the reference reaches
// through gnark / gnark-crypto (go.mod:15-16) from /root/reference/prover/prover_cpu.go:37:
//   - fp/fr Montgomery arithmetic                     (gnark-crypto ecc/<curve>/fp, fr)
//   - G1/G2 MultiExp: signed-digit Pippenger           (gnark-crypto ecc/<curve>/multiexp.go)
//   - fft.Domain FFT / FFTInverse (DIF/DIT, coset)     (gnark-crypto ecc/<curve>/fr/fft)
//   - computeH and the Prove MSM schedule              (gnark backend/groth16/<curve>/prove.go)
// Parity: validated bit-for-bit against the big-int oracle (oracle/*.py) in tests/test_oracle_c.py;
// the reference itself cannot be built here (no Go toolchain), so this is cpu_baseline kind "port".
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

// ----------------------------------------------------------------------------- field parameters
struct FieldParams {
  int L;            // 64-bit limbs
  u64 p[12];        // modulus
  u64 inv;          // -p^-1 mod 2^64
  u64 one[12];      // R mod p
  u64 r2[12];       // R^2 mod p
  int bits;
};

static void hex_to_limbs(const char* hex, u64* out, int L) {
  memset(out, 0, 8 * L);
  int n = (int)strlen(hex);
  for (int i = 0; i < n; i++) {
    char ch = hex[n - 1 - i];
    u64 v = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
    out[i / 16] |= v << (4 * (i % 16));
  }
}

static int cmp_limbs(const u64* a, const u64* b, int L) {
  for (int i = L - 1; i >= 0; i--) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}
static u64 add_limbs(u64* r, const u64* a, const u64* b, int L) {
  u64 c = 0;
  for (int i = 0; i < L; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  return c;
}
static u64 sub_limbs(u64* r, const u64* a, const u64* b, int L) {
  u64 bw = 0;
  for (int i = 0; i < L; i++) {
    u128 t = (u128)a[i] - b[i] - bw;
    r[i] = (u64)t;
    bw = (u64)(t >> 64) & 1;
  }
  return bw;
}

static void init_field(FieldParams& f, const char* hex, int L) {
  f.L = L;
  hex_to_limbs(hex, f.p, L);
  u64 x = 1;   // Newton: x = p^-1 mod 2^64
  for (int i = 0; i < 6; i++) x *= 2 - f.p[0] * x;
  f.inv = (u64)(0 - x);
  // R mod p by doubling 1, 64L times ; R^2 by another 64L doublings
  u64 v[12];
  memset(v, 0, sizeof v);
  v[0] = 1;
  for (int k = 0; k < 2 * 64 * L; k++) {
    u64 c = add_limbs(v, v, v, L);
    if (c || cmp_limbs(v, f.p, L) >= 0) sub_limbs(v, v, f.p, L);
    if (k == 64 * L - 1) memcpy(f.one, v, 8 * L);
  }
  memcpy(f.r2, v, 8 * L);
  f.bits = 0;
  for (int i = L - 1; i >= 0 && !f.bits; i--)
    if (f.p[i]) f.bits = 64 * i + 64 - __builtin_clzll(f.p[i]);
}

// ----------------------------------------------------------------------------- Fp<L>
template <int L>
struct Fe {
  u64 v[L];
};

template <int L, int TAG>
struct Fp {
  static FieldParams P;
  typedef Fe<L> El;
  static void add(El& r, const El& a, const El& b) {
    u64 c = add_limbs(r.v, a.v, b.v, L);
    if (c || cmp_limbs(r.v, P.p, L) >= 0) sub_limbs(r.v, r.v, P.p, L);
  }
  static void sub(El& r, const El& a, const El& b) {
    if (sub_limbs(r.v, a.v, b.v, L)) add_limbs(r.v, r.v, P.p, L);
  }
  static void dbl(El& r, const El& a) { add(r, a, a); }
  static void neg(El& r, const El& a) {
    El z;
    memset(&z, 0, sizeof z);
    sub(r, z, a);
  }
  // CIOS Montgomery multiplication
  static void mul(El& r, const El& a, const El& b) {
    u64 t[L + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < L; i++) {
      u64 c = 0;
      for (int j = 0; j < L; j++) {
        u128 x = (u128)a.v[j] * b.v[i] + t[j] + c;
        t[j] = (u64)x;
        c = (u64)(x >> 64);
      }
      u128 x = (u128)t[L] + c;
      t[L] = (u64)x;
      t[L + 1] = (u64)(x >> 64);
      u64 m = t[0] * P.inv;
      x = (u128)m * P.p[0] + t[0];
      c = (u64)(x >> 64);
      for (int j = 1; j < L; j++) {
        x = (u128)m * P.p[j] + t[j] + c;
        t[j - 1] = (u64)x;
        c = (u64)(x >> 64);
      }
      x = (u128)t[L] + c;
      t[L - 1] = (u64)x;
      t[L] = t[L + 1] + (u64)(x >> 64);
    }
    if (t[L] || cmp_limbs(t, P.p, L) >= 0) sub_limbs(t, t, P.p, L);
    memcpy(r.v, t, 8 * L);
  }
  static void sqr(El& r, const El& a) { mul(r, a, a); }
  static void set_zero(El& r) { memset(&r, 0, sizeof r); }
  static void set_one(El& r) { memcpy(r.v, P.one, 8 * L); }
  static bool is_zero(const El& a) {
    u64 o = 0;
    for (int i = 0; i < L; i++) o |= a.v[i];
    return o == 0;
  }
  static bool eq(const El& a, const El& b) { return memcmp(a.v, b.v, 8 * L) == 0; }
  static void from_mont(El& r, const El& a) {
    El one;
    memset(&one, 0, sizeof one);
    one.v[0] = 1;
    mul(r, a, one);
  }
  static void to_mont(El& r, const El& a) {
    El r2;
    memcpy(r2.v, P.r2, 8 * L);
    mul(r, a, r2);
  }
  static void pow(El& r, const El& a, const u64* e, int ne) {
    El acc;
    set_one(acc);
    for (int i = ne * 64 - 1; i >= 0; i--) {
      sqr(acc, acc);
      if ((e[i >> 6] >> (i & 63)) & 1) mul(acc, acc, a);
    }
    r = acc;
  }
  static void inv(El& r, const El& a) {
    u64 e[L], two[L];
    memset(two, 0, sizeof two);
    two[0] = 2;
    sub_limbs(e, P.p, two, L);
    pow(r, a, e, L);
  }
};
template <int L, int TAG>
FieldParams Fp<L, TAG>::P;

// Fp2 = Fp[u]/(u^2 + NRN)
template <class B, int NRN>
struct Fp2 {
  typedef typename B::El BEl;
  struct El {
    BEl c0, c1;
  };
  static void add(El& r, const El& a, const El& b) {
    B::add(r.c0, a.c0, b.c0);
    B::add(r.c1, a.c1, b.c1);
  }
  static void sub(El& r, const El& a, const El& b) {
    B::sub(r.c0, a.c0, b.c0);
    B::sub(r.c1, a.c1, b.c1);
  }
  static void dbl(El& r, const El& a) { add(r, a, a); }
  static void neg(El& r, const El& a) {
    B::neg(r.c0, a.c0);
    B::neg(r.c1, a.c1);
  }
  static void mul_nrn(BEl& r, const BEl& a) {
    BEl acc = a;
    for (int i = 1; i < NRN; i++) B::add(acc, acc, a);
    r = acc;
  }
  static void mul(El& r, const El& a, const El& b) {   // Karatsuba, 3 base multiplications (as gnark-crypto's E2.Mul)
    BEl t0, t1, sa, sb;
    B::mul(t0, a.c0, b.c0);
    B::mul(t1, a.c1, b.c1);
    B::add(sa, a.c0, a.c1);
    B::add(sb, b.c0, b.c1);
    B::mul(sa, sa, sb);
    B::sub(sa, sa, t0);
    B::sub(sa, sa, t1);
    mul_nrn(t1, t1);
    B::sub(r.c0, t0, t1);
    r.c1 = sa;
  }
  static void sqr(El& r, const El& a) {   // 2 base multiplications
    BEl m, sum, d, t;
    B::mul(m, a.c0, a.c1);
    B::add(sum, a.c0, a.c1);
    mul_nrn(t, a.c1);
    B::sub(d, a.c0, t);
    B::mul(sum, sum, d);          // a0^2 - NRN a1^2 + (1 - NRN) a0 a1
    BEl corr = m;
    for (int i = 2; i < NRN; i++) B::add(corr, corr, m);   // (NRN - 1) a0 a1
    if (NRN > 1) B::add(sum, sum, corr);
    r.c0 = sum;
    B::dbl(r.c1, m);
  }
  static void set_zero(El& r) { memset(&r, 0, sizeof r); }
  static void set_one(El& r) {
    B::set_one(r.c0);
    B::set_zero(r.c1);
  }
  static bool is_zero(const El& a) { return B::is_zero(a.c0) && B::is_zero(a.c1); }
  static bool eq(const El& a, const El& b) { return B::eq(a.c0, b.c0) && B::eq(a.c1, b.c1); }
  static void inv(El& r, const El& a) {
    BEl n, t;
    B::sqr(n, a.c0);
    B::sqr(t, a.c1);
    mul_nrn(t, t);
    B::add(n, n, t);
    B::inv(n, n);
    B::mul(r.c0, a.c0, n);
    B::mul(t, a.c1, n);
    B::neg(r.c1, t);
  }
};

// ----------------------------------------------------------------------------- group law (Jacobian)
template <class F>
struct Curve {
  typedef typename F::El El;
  struct Aff {
    El x, y;
  };
  struct Jac {
    El x, y, z;
  };
  static bool aff_inf(const Aff& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
  static void set_inf(Jac& p) {
    F::set_one(p.x);
    F::set_one(p.y);
    F::set_zero(p.z);
  }
  static bool is_inf(const Jac& p) { return F::is_zero(p.z); }
  static void dbl(Jac& p) {   // dbl-2009-l (a = 0)
    if (is_inf(p)) return;
    El A, B, C, D, E, Fq, t;
    F::sqr(A, p.x);
    F::sqr(B, p.y);
    F::sqr(C, B);
    F::add(t, p.x, B);
    F::sqr(t, t);
    F::sub(t, t, A);
    F::sub(t, t, C);
    F::dbl(D, t);
    F::dbl(E, A);
    F::add(E, E, A);
    F::sqr(Fq, E);
    El z3;
    F::mul(z3, p.y, p.z);
    F::dbl(z3, z3);
    F::sub(p.x, Fq, D);
    F::sub(p.x, p.x, D);
    F::sub(t, D, p.x);
    F::mul(t, E, t);
    F::dbl(C, C);
    F::dbl(C, C);
    F::dbl(C, C);
    F::sub(p.y, t, C);
    p.z = z3;
  }
  static void add(Jac& p, const Jac& q) {   // add-2007-bl
    if (is_inf(q)) return;
    if (is_inf(p)) {
      p = q;
      return;
    }
    El z1z1, z2z2, u1, u2, s1, s2, h, r, hh, hhh, v, t;
    F::sqr(z1z1, p.z);
    F::sqr(z2z2, q.z);
    F::mul(u1, p.x, z2z2);
    F::mul(u2, q.x, z1z1);
    F::mul(s1, p.y, q.z);
    F::mul(s1, s1, z2z2);
    F::mul(s2, q.y, p.z);
    F::mul(s2, s2, z1z1);
    if (F::eq(u1, u2)) {
      if (F::eq(s1, s2)) dbl(p);
      else set_inf(p);
      return;
    }
    F::sub(h, u2, u1);
    F::sub(r, s2, s1);
    F::sqr(hh, h);
    F::mul(hhh, h, hh);
    F::mul(v, u1, hh);
    F::sqr(t, r);
    F::sub(t, t, hhh);
    F::sub(t, t, v);
    El x3;
    F::sub(x3, t, v);
    F::sub(t, v, x3);
    F::mul(t, r, t);
    F::mul(s1, s1, hhh);
    F::sub(p.y, t, s1);
    p.x = x3;
    F::mul(p.z, p.z, q.z);
    F::mul(p.z, p.z, h);
  }
  static void madd(Jac& p, const Aff& q, bool negate) {
    if (aff_inf(q)) return;
    Jac j;
    j.x = q.x;
    j.y = q.y;
    if (negate) F::neg(j.y, j.y);
    F::set_one(j.z);
    add(p, j);
  }
  static void to_affine(Aff& r, const Jac& p) {
    if (is_inf(p)) {
      F::set_zero(r.x);
      F::set_zero(r.y);
      return;
    }
    El zi, zi2, zi3;
    F::inv(zi, p.z);
    F::sqr(zi2, zi);
    F::mul(zi3, zi2, zi);
    F::mul(r.x, p.x, zi2);
    F::mul(r.y, p.y, zi3);
  }
  // extended-Jacobian bucket (x = X/ZZ, y = Y/ZZZ): gnark-crypto's g1JacExtended, mixed addition 8M + 2S
  struct Ext {
    El x, y, zz, zzz;
  };
  static void ext_set_inf(Ext& p) {
    F::set_zero(p.x);
    F::set_zero(p.y);
    F::set_zero(p.zz);
    F::set_zero(p.zzz);
  }
  static bool ext_is_inf(const Ext& p) { return F::is_zero(p.zz); }
  static void ext_dbl_affine(Ext& r, const El& qx, const El& qy) {   // mdbl-2008-s-1
    El u, v, w, s2, m, t;
    F::dbl(u, qy);
    F::sqr(v, u);
    F::mul(w, u, v);
    F::mul(s2, qx, v);
    F::sqr(m, qx);
    F::dbl(t, m);
    F::add(m, m, t);
    F::sqr(r.x, m);
    F::sub(r.x, r.x, s2);
    F::sub(r.x, r.x, s2);
    F::sub(t, s2, r.x);
    F::mul(t, m, t);
    F::mul(u, w, qy);
    F::sub(r.y, t, u);
    r.zz = v;
    r.zzz = w;
  }
  static void ext_madd(Ext& p, const Aff& q, bool negate) {   // madd-2008-s
    if (aff_inf(q)) return;
    El qy = q.y;
    if (negate) F::neg(qy, qy);
    if (ext_is_inf(p)) {
      p.x = q.x;
      p.y = qy;
      F::set_one(p.zz);
      F::set_one(p.zzz);
      return;
    }
    El u2, s2, pp, ppp, qq, t;
    F::mul(u2, q.x, p.zz);
    F::mul(s2, qy, p.zzz);
    F::sub(u2, u2, p.x);
    F::sub(s2, s2, p.y);
    if (F::is_zero(u2)) {
      if (F::is_zero(s2)) ext_dbl_affine(p, q.x, qy);
      else ext_set_inf(p);
      return;
    }
    F::sqr(pp, u2);
    F::mul(ppp, u2, pp);
    F::mul(qq, p.x, pp);
    F::sqr(t, s2);
    F::sub(t, t, ppp);
    F::sub(t, t, qq);
    F::sub(p.x, t, qq);
    F::sub(qq, qq, p.x);
    F::mul(qq, s2, qq);
    F::mul(t, p.y, ppp);
    F::sub(p.y, qq, t);
    F::mul(p.zz, p.zz, pp);
    F::mul(p.zzz, p.zzz, ppp);
  }
  // bucket -> Jacobian for the (Jacobian) running sums
  static void ext_to_jac(Jac& r, const Ext& p) {
    if (ext_is_inf(p)) {
      set_inf(r);
      return;
    }
    // choose Z = ZZ (then Z^2 = ZZ^2, Z^3 = ZZ^3 = ZZZ^2):  X_j = x Z^2 = X ZZ ;  Y_j = y Z^3 = Y ZZZ
    F::mul(r.x, p.x, p.zz);
    F::mul(r.y, p.y, p.zzz);
    r.z = p.zz;
  }
  static void mul_scalar(Jac& r, const Jac& p, const u64* k, int nk) {
    Jac acc;
    set_inf(acc);
    for (int i = nk * 64 - 1; i >= 0; i--) {
      dbl(acc);
      if ((k[i >> 6] >> (i & 63)) & 1) add(acc, p);
    }
    r = acc;
  }
};

// ----------------------------------------------------------------------------- Pippenger MSM
// scalars: canonical (non-Montgomery) little-endian, LS limbs each.  index_map semantics as the
// product's (scalar i multiplies points[map[i]], 0xffffffff skips) so the prove schedule can be
// restated with the same proving-key arrays.
// frac_k / frac_n: only digit windows [nwin k / n, nwin (k + 1) / n) are processed - a bounded SAMPLE of the MSM for
// the CPU baseline that runs exactly the code of its share of the full MSM (Pippenger's windows are independent).
template <class F>
static void msm(typename Curve<F>::Jac& out, const typename Curve<F>::Aff* pts, const u64* scal, int LS, size_t n,
                int scalar_bits, const uint32_t* map, int threads, int frac_k = 0, int frac_n = 1) {
  typedef Curve<F> C;
  typedef typename C::Jac Jac;
  typedef typename C::Ext Ext;
  const size_t lo_s = 0, hi_s = n;
  int c = 4;
  {
    double best = 1e300;
    for (int cc = 2; cc <= 16; cc++) {
      int nw = (scalar_bits + 1 + cc - 1) / cc;
      double cost = (double)nw * (n + 2.0 * (1 << (cc - 1)));
      if (cost < best) best = cost, c = cc;
    }
  }
  const int nwin = (scalar_bits + 1 + c - 1) / c;
  const int nb = 1 << (c - 1);
  // signed digits (carry across windows)
  std::vector<int32_t> dig((size_t)nwin * n);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (long i = (long)lo_s; i < (long)hi_s; i++) {
    const u64* s = scal + (size_t)i * LS;
    int carry = 0;
    bool skip = map && map[i] == 0xffffffffu;
    for (int w = 0; w < nwin; w++) {
      int bit = w * c, limb = bit >> 6, off = bit & 63;
      u64 v = 0;
      if (limb < LS) {
        v = s[limb] >> off;
        if (off + c > 64 && limb + 1 < LS) v |= s[limb + 1] << (64 - off);
      }
      int d = (int)(v & ((1u << c) - 1)) + carry;
      carry = 0;
      if (d > nb) {
        d -= (1 << c);
        carry = 1;
      }
      dig[(size_t)w * n + i] = skip ? 0 : d;
    }
  }
  std::vector<Jac> wins(nwin);
  const int w_lo = nwin * frac_k / frac_n, w_hi = nwin * (frac_k + 1) / frac_n, w_cnt = std::max(w_hi - w_lo, 1);
  // windows in parallel; within a window the point range is split so all threads stay busy
  int split = std::max(1, threads / w_cnt + (threads % w_cnt ? 1 : 0));
  std::vector<Jac> parts((size_t)nwin * split);
  for (auto& pj : parts) C::set_inf(pj);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int job = w_lo * split; job < w_hi * split; job++) {
    int w = job / split, part = job % split;
    size_t lo = lo_s + (hi_s - lo_s) * part / split, hi = lo_s + (hi_s - lo_s) * (part + 1) / split;
    std::vector<Ext> buckets(nb);
    for (int b = 0; b < nb; b++) C::ext_set_inf(buckets[b]);
    const int32_t* dw = dig.data() + (size_t)w * n;
    for (size_t i = lo; i < hi; i++) {
      int d = dw[i];
      if (!d) continue;
      size_t pi = map ? map[i] : i;
      if (d > 0) C::ext_madd(buckets[d - 1], pts[pi], false);
      else C::ext_madd(buckets[-d - 1], pts[pi], true);
    }
    Jac run, acc, bj;
    C::set_inf(run);
    C::set_inf(acc);
    for (int b = nb - 1; b >= 0; b--) {
      if (!C::ext_is_inf(buckets[b])) {
        C::ext_to_jac(bj, buckets[b]);
        C::add(run, bj);
      }
      C::add(acc, run);
    }
    parts[job] = acc;
  }
  for (int w = 0; w < nwin; w++) {
    C::set_inf(wins[w]);
    for (int s2 = 0; s2 < split; s2++) C::add(wins[w], parts[(size_t)w * split + s2]);
  }
  Jac total;
  C::set_inf(total);
  for (int w = nwin - 1; w >= 0; w--) {
    for (int i = 0; i < c; i++) C::dbl(total);
    C::add(total, wins[w]);
  }
  out = total;
}

// ----------------------------------------------------------------------------- NTT
static inline uint32_t brev(uint32_t i, int logn) {
  uint32_t r = 0;
  for (int b = 0; b < logn; b++) r |= ((i >> b) & 1u) << (logn - 1 - b);
  return r;
}

template <class Fr>
struct Ntt {
  typedef typename Fr::El El;
  // in-place, natural in -> bit-reversed out (DIF) with root w (Montgomery)
  static void dif(El* a, int logn, const std::vector<El>& tw, int threads) {
    size_t n = (size_t)1 << logn;
    for (int s = 0; s < logn; s++) {
      size_t d = n >> (s + 1);
#pragma omp parallel for num_threads(threads) schedule(static)
      for (long q = 0; q < (long)(n / 2); q++) {
        size_t blk = q / d, j = q % d;
        size_t i0 = blk * 2 * d + j, i1 = i0 + d;
        El t;
        Fr::add(t, a[i0], a[i1]);
        Fr::sub(a[i1], a[i0], a[i1]);
        Fr::mul(a[i1], a[i1], tw[j << s]);
        a[i0] = t;
      }
    }
  }
  // bit-reversed in -> natural out (DIT)
  static void dit(El* a, int logn, const std::vector<El>& tw, int threads) {
    size_t n = (size_t)1 << logn;
    for (int s = logn - 1; s >= 0; s--) {
      size_t d = n >> (s + 1);
#pragma omp parallel for num_threads(threads) schedule(static)
      for (long q = 0; q < (long)(n / 2); q++) {
        size_t blk = q / d, j = q % d;
        size_t i0 = blk * 2 * d + j, i1 = i0 + d;
        El t;
        Fr::mul(t, a[i1], tw[j << s]);
        Fr::sub(a[i1], a[i0], t);
        Fr::add(a[i0], a[i0], t);
      }
    }
  }
  static std::vector<El> powers(const El& base, size_t count, int threads = 1) {
    std::vector<El> t(count);
    const size_t chunk = (count + threads - 1) / std::max(threads, 1);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long c = 0; c < (long)threads; c++) {
      size_t lo = (size_t)c * chunk, hi = std::min(count, lo + chunk);
      if (lo >= hi) continue;
      El x;
      Fr::set_one(x);
      El b = base;
      for (size_t e = lo; e; e >>= 1) {   // x = base^lo
        if (e & 1) Fr::mul(x, x, b);
        Fr::sqr(b, b);
      }
      for (size_t i = lo; i < hi; i++) {
        t[i] = x;
        Fr::mul(x, x, base);
      }
    }
    return t;
  }
};

template <class Fr>
struct Domain {
  typedef typename Fr::El El;
  int logn;
  size_t n;
  El omega, omega_inv, g, g_inv, n_inv, den;
  std::vector<El> tw, twi, gp, gip;   // omega^k, omega^-k (k<n/2); g^k, g^-k (k<n)
  void pointwise(El* a, El* b, El* c, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long i = 0; i < (long)n; i++) {
      Fr::mul(a[i], a[i], b[i]);
      Fr::sub(a[i], a[i], c[i]);
      Fr::mul(a[i], a[i], den);
    }
  }
  void init(int logn_, const El& w, const El& gg, int threads = 1) {
    logn = logn_;
    n = (size_t)1 << logn;
    omega = w;
    g = gg;
    Fr::inv(omega_inv, w);
    Fr::inv(g_inv, gg);
    El nn;
    Fr::set_one(nn);
    for (int i = 0; i < logn; i++) Fr::dbl(nn, nn);
    Fr::inv(n_inv, nn);
    El t = gg, one;
    Fr::set_one(one);
    for (int i = 0; i < logn; i++) Fr::sqr(t, t);
    Fr::sub(t, t, one);
    Fr::inv(den, t);
    tw = Ntt<Fr>::powers(omega, std::max<size_t>(n / 2, 1), threads);
    twi = Ntt<Fr>::powers(omega_inv, std::max<size_t>(n / 2, 1), threads);
    gp = Ntt<Fr>::powers(g, n, threads);
    gip = Ntt<Fr>::powers(g_inv, n, threads);
  }
  // gnark fft.Domain semantics
  void fft(El* a, bool inverse, bool dit, bool coset, int threads) {
    if (!inverse && coset) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (long i = 0; i < (long)n; i++) Fr::mul(a[i], a[i], gp[dit ? brev((uint32_t)i, logn) : (size_t)i]);
    }
    if (dit) Ntt<Fr>::dit(a, logn, inverse ? twi : tw, threads);
    else Ntt<Fr>::dif(a, logn, inverse ? twi : tw, threads);
    if (inverse) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (long i = 0; i < (long)n; i++) {
        Fr::mul(a[i], a[i], n_inv);
        if (coset) Fr::mul(a[i], a[i], gip[dit ? (size_t)i : brev((uint32_t)i, logn)]);
      }
    }
  }
  void compute_h(El* a, El* b, El* c, int threads) {
    fft(a, true, false, false, threads);
    fft(b, true, false, false, threads);
    fft(c, true, false, false, threads);
    fft(a, false, true, true, threads);
    fft(b, false, true, true, threads);
    fft(c, false, true, true, threads);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long i = 0; i < (long)n; i++) {
      Fr::mul(a[i], a[i], b[i]);
      Fr::sub(a[i], a[i], c[i]);
      Fr::mul(a[i], a[i], den);
    }
    fft(a, true, false, true, threads);
  }
};

// ----------------------------------------------------------------------------- curve instances
static const char* HEX_P[5] = {
    "", "30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47",
    "1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001",
    "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab",
    "122e824fb83ce0ad187c94004faff3eb926186a81d14688528275ef8087be41707ba638e584e91903cebaff25b423048689c8ed12f9fd9071dcd3dc73ebff2e98a116c25667a8f8160cf8aeeaf0a437e6913e6870000082f49d00000000008b"};
static const char* HEX_R[5] = {
    "", "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001",
    "12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001",
    "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001",
    "1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001"};

template <int ID, int LP, int LR, int NRN, bool G2_OVER_FP>
struct CurveCfg {
  typedef Fp<LP, ID * 2> FP;
  typedef Fp<LR, ID * 2 + 1> FR;
  typedef FP G1F;
  typedef typename std::conditional<G2_OVER_FP, FP, Fp2<FP, NRN>>::type G2F;
  static void init() {
    static bool done = false;
#pragma omp critical(oracle_init)
    if (!done) {
      init_field(FP::P, HEX_P[ID], LP);
      init_field(FR::P, HEX_R[ID], LR);
      done = true;
    }
  }
};
typedef CurveCfg<1, 4, 4, 1, false> Bn254;
typedef CurveCfg<2, 6, 4, 5, false> Bls377;
typedef CurveCfg<3, 6, 4, 1, false> Bls381;
typedef CurveCfg<4, 12, 6, 1, true> Bw6;

// ----------------------------------------------------------------------------- generic entry points
template <class Cfg, class F>
static void msm_entry(const void* points, const void* scalars_mont, size_t n, const uint32_t* map, void* out_aff,
                      int threads) {
  typedef typename Cfg::FR FR;
  typedef Curve<F> C;
  constexpr int LS = sizeof(typename FR::El) / 8;
  std::vector<u64> canon(n * LS);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (long i = 0; i < (long)n; i++) {
    typename FR::El e;
    FR::from_mont(e, ((const typename FR::El*)scalars_mont)[i]);
    memcpy(&canon[(size_t)i * LS], e.v, 8 * LS);
  }
  typename C::Jac r;
  msm<F>(r, (const typename C::Aff*)points, canon.data(), LS, n, FR::P.bits, map, threads);
  C::to_affine(*(typename C::Aff*)out_aff, r);
}

template <class Cfg>
static int dispatch_msm(int group, const void* points, const void* scalars, size_t n, const uint32_t* map, void* out,
                        int threads) {
  Cfg::init();
  if (group == 1) msm_entry<Cfg, typename Cfg::G1F>(points, scalars, n, map, out, threads);
  else msm_entry<Cfg, typename Cfg::G2F>(points, scalars, n, map, out, threads);
  return 0;
}

template <class Cfg>
static int dispatch_fft(void* data, int logn, const void* omega, const void* g, int inverse, int dit, int coset,
                        int threads) {
  Cfg::init();
  typedef typename Cfg::FR FR;
  Domain<FR> d;
  d.init(logn, *(const typename FR::El*)omega, *(const typename FR::El*)g, threads);
  d.fft((typename FR::El*)data, inverse, dit, coset, threads);
  return 0;
}

template <class Cfg>
static int dispatch_h(void* a, void* b, void* c, int logn, const void* omega, const void* g, int threads) {
  Cfg::init();
  typedef typename Cfg::FR FR;
  Domain<FR> d;
  d.init(logn, *(const typename FR::El*)omega, *(const typename FR::El*)g, threads);
  d.compute_h((typename FR::El*)a, (typename FR::El*)b, (typename FR::El*)c, threads);
  return 0;
}

// The Prove MSM schedule with the same extended-array / index-map convention as the product
// (A_ext = A || delta || alpha with scalars W || r || s || 1 || -rs, see prover.cu), so one set of
// proving-key buffers feeds both.  Outputs affine Ar, Bs, Krs (and Pok when a sigma basis is given).
// nparts == 10 turns the call into a bounded SAMPLE of the proof for the CPU baseline: component `part` of
//   0: quotient H (7 transforms) + PoK MSM + assembly   1: MSM A   2: MSM B1   3, 4: halves of the G2 MSM (by window
//   range)   5: MSM K   6..9: quarters of the quotient MSM Z (by window range)
// each running exactly the code the full proof runs for it; the outputs are then meaningless.  comp_seconds (7
// doubles, optional) receives the wall time of H, A, B1, B2, K, Z, PoK + assembly.
struct ProveArgs {
  int curve, logn;
  const void *omega, *g;
  const void *A_ext, *B1_ext, *B2_ext, *K_ext, *Z;
  const uint32_t *mapA, *mapB, *mapK;
  uint64_t m, nb_public, nZ;
  const void* W_ext;          // m + 4 scalars (Montgomery): wires, r, s, 1, -rs
  void *a, *b, *c;            // n scalars each (overwritten)
  void *out_ar, *out_bs, *out_krs;
  int threads;
  // --- appended in round 2 (zero = absent)
  const void* sigma;          // BasisExpSigma points of the commitment (n_commit)
  const void* cvals;          // committed wire values (Montgomery)
  uint64_t n_commit;
  void* out_pok;
  int part, nparts;
  double* comp_seconds;
};

template <class Cfg>
static int prove_impl(const ProveArgs& p) {
  Cfg::init();
  typedef typename Cfg::FR FR;
  typedef Curve<typename Cfg::G1F> C1;
  typedef Curve<typename Cfg::G2F> C2;
  constexpr int LS = sizeof(typename FR::El) / 8;
  const int T = p.threads;
  const bool sample = p.nparts > 1;
  if (sample && p.nparts != 10) return -2;
  auto on = [&](int first, int last) { return !sample || (p.part >= first && p.part <= last); };
  double ts[7] = {0, 0, 0, 0, 0, 0, 0};
  typename FR::El* va = (typename FR::El*)p.a;
  typename FR::El* vb = (typename FR::El*)p.b;
  typename FR::El* vc = (typename FR::El*)p.c;
  double t0 = omp_get_wtime();
  if (on(0, 0)) {
    // gnark keeps the domain's twiddle tables in the proving key (fft.NewDomain at setup): build them once per
    // (curve, size) and keep them across proofs
    static Domain<FR> d;
    static int d_logn = -1;
    if (d_logn != p.logn || !FR::eq(d.omega, *(const typename FR::El*)p.omega) || !FR::eq(d.g, *(const typename FR::El*)p.g)) {
      d.init(p.logn, *(const typename FR::El*)p.omega, *(const typename FR::El*)p.g, T);
      d_logn = p.logn;
    }
    d.compute_h(va, vb, vc, T);
  }
  ts[0] = omp_get_wtime() - t0;
  auto canon = [&](const void* mont, size_t n) {
    std::vector<u64> v(n * LS);
#pragma omp parallel for num_threads(T) schedule(static)
    for (long i = 0; i < (long)n; i++) {
      typename FR::El e;
      FR::from_mont(e, ((const typename FR::El*)mont)[i]);
      memcpy(&v[(size_t)i * LS], e.v, 8 * LS);
    }
    return v;
  };
  typename C1::Jac ar, bs1, k, z, pok;
  typename C2::Jac bs2;
  const size_t nw = p.m + 4, nk = p.m - p.nb_public + 4;
  std::vector<u64> w;
  if (on(0, 5)) w = canon(p.W_ext, nw);
  t0 = omp_get_wtime();
  if (on(1, 1)) msm<typename Cfg::G1F>(ar, (const typename C1::Aff*)p.A_ext, w.data(), LS, nw, FR::P.bits, p.mapA, T);
  ts[1] = omp_get_wtime() - t0;
  t0 = omp_get_wtime();
  if (on(2, 2)) msm<typename Cfg::G1F>(bs1, (const typename C1::Aff*)p.B1_ext, w.data(), LS, nw, FR::P.bits, p.mapB, T);
  ts[2] = omp_get_wtime() - t0;
  t0 = omp_get_wtime();
  if (on(3, 4))
    msm<typename Cfg::G2F>(bs2, (const typename C2::Aff*)p.B2_ext, w.data(), LS, nw, FR::P.bits, p.mapB, T,
                           sample ? p.part - 3 : 0, sample ? 2 : 1);
  ts[3] = omp_get_wtime() - t0;
  t0 = omp_get_wtime();
  if (on(5, 5))
    msm<typename Cfg::G1F>(k, (const typename C1::Aff*)p.K_ext, w.data() + p.nb_public * LS, LS, nk, FR::P.bits, p.mapK, T);
  ts[4] = omp_get_wtime() - t0;
  t0 = omp_get_wtime();
  if (on(6, 9)) {
    std::vector<u64> h = canon(p.a, p.nZ);
    msm<typename Cfg::G1F>(z, (const typename C1::Aff*)p.Z, h.data(), LS, p.nZ, FR::P.bits, nullptr, T,
                           sample ? p.part - 6 : 0, sample ? 4 : 1);
  }
  ts[5] = omp_get_wtime() - t0;
  t0 = omp_get_wtime();
  if (on(0, 0) && p.sigma && p.n_commit) {
    std::vector<u64> cv = canon(p.cvals, p.n_commit);
    msm<typename Cfg::G1F>(pok, (const typename C1::Aff*)p.sigma, cv.data(), LS, p.n_commit, FR::P.bits, nullptr, T);
    if (p.out_pok) C1::to_affine(*(typename C1::Aff*)p.out_pok, pok);
  }
  if (!sample) {
    typename C1::Jac t0j, t1j;
    C1::mul_scalar(t0j, ar, w.data() + (p.m + 1) * LS, LS);    // s * Ar
    C1::mul_scalar(t1j, bs1, w.data() + (p.m + 0) * LS, LS);   // r * Bs1
    C1::add(k, z);
    C1::add(k, t0j);
    C1::add(k, t1j);
    C1::to_affine(*(typename C1::Aff*)p.out_ar, ar);
    C2::to_affine(*(typename C2::Aff*)p.out_bs, bs2);
    C1::to_affine(*(typename C1::Aff*)p.out_krs, k);
  }
  ts[6] = omp_get_wtime() - t0;
  if (p.comp_seconds) memcpy(p.comp_seconds, ts, sizeof ts);
  return 0;
}

// out[i] = [k0 + i] G as affine points (Montgomery): valid, distinct subgroup points for the CPU baseline's
// synthetic key.  Blocks of B points Q_j + T_i (Q_j = [k0 + j B] G, T_i = [i] G) share one field inversion.
template <class Cfg, class F>
static void gen_points_impl(const void* gen_affine, uint64_t k0, size_t n, void* out_v, int threads) {
  typedef Curve<F> C;
  typedef typename C::Aff Aff;
  typedef typename C::Jac Jac;
  typedef typename F::El El;
  const size_t B = 256;
  Aff* out = (Aff*)out_v;
  const Aff g = *(const Aff*)gen_affine;
  Jac gj;
  gj.x = g.x;
  gj.y = g.y;
  F::set_one(gj.z);
  // T_i = [i] G, i < B (T_0 = infinity)
  std::vector<Aff> Tt(B);
  {
    Jac acc;
    C::set_inf(acc);
    for (size_t i = 0; i < B; i++) {
      C::to_affine(Tt[i], acc);
      C::add(acc, gj);
    }
  }
  const size_t nblocks = (n + B - 1) / B;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (long j = 0; j < (long)nblocks; j++) {
    u64 kk[1] = {k0 + (u64)j * B};
    Jac qj;
    C::mul_scalar(qj, gj, kk, 1);
    Aff q;
    C::to_affine(q, qj);
    const size_t cnt = std::min(B, n - (size_t)j * B);
    // affine additions q + T_i, i = 1 .. cnt-1, with Montgomery's batch inversion of (x_T - x_q)
    std::vector<El> dx(cnt), pre(cnt);
    El run;
    F::set_one(run);
    for (size_t i = 1; i < cnt; i++) {
      F::sub(dx[i], Tt[i].x, q.x);
      pre[i] = run;
      F::mul(run, run, dx[i]);
    }
    El inv;
    F::inv(inv, run);
    out[(size_t)j * B] = q;
    for (size_t i = cnt - 1; i >= 1; i--) {
      El dinv, lam, t, x3, y3;
      F::mul(dinv, inv, pre[i]);
      F::mul(inv, inv, dx[i]);
      F::sub(t, Tt[i].y, q.y);
      F::mul(lam, t, dinv);
      F::sqr(x3, lam);
      F::sub(x3, x3, q.x);
      F::sub(x3, x3, Tt[i].x);
      F::sub(t, q.x, x3);
      F::mul(y3, lam, t);
      F::sub(y3, y3, q.y);
      out[(size_t)j * B + i].x = x3;
      out[(size_t)j * B + i].y = y3;
    }
  }
}

template <class Cfg>
static int dispatch_gen(int group, const void* gen, uint64_t k0, size_t n, void* out, int threads) {
  Cfg::init();
  if (group == 1) gen_points_impl<Cfg, typename Cfg::G1F>(gen, k0, n, out, threads);
  else gen_points_impl<Cfg, typename Cfg::G2F>(gen, k0, n, out, threads);
  return 0;
}

template <class Cfg>
static int fr_mul_impl(const void* a, const void* b, void* out, uint64_t n, int to_mont, int threads) {
  Cfg::init();
  typedef typename Cfg::FR FR;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (long i = 0; i < (long)n; i++) {
    typename FR::El x = ((const typename FR::El*)a)[i];
    if (to_mont) FR::to_mont(((typename FR::El*)out)[i], x);
    else FR::mul(((typename FR::El*)out)[i], x, ((const typename FR::El*)b)[i]);
  }
  return 0;
}

#define DISPATCH(curve, CALL)            \
  switch (curve) {                       \
    case 1: return CALL(Bn254);          \
    case 2: return CALL(Bls377);         \
    case 3: return CALL(Bls381);         \
    case 4: return CALL(Bw6);            \
    default: return -1;                  \
  }

extern "C" {

int oc_num_threads(void) { return omp_get_max_threads(); }

int oc_msm(int curve, int group, const void* points, const void* scalars_mont, uint64_t n, const uint32_t* map,
           void* out_affine, int threads) {
  if (threads <= 0) threads = omp_get_max_threads();
#define CALL(C) dispatch_msm<C>(group, points, scalars_mont, (size_t)n, map, out_affine, threads)
  DISPATCH(curve, CALL)
#undef CALL
}

int oc_fft(int curve, void* data, int logn, const void* omega, const void* g, int inverse, int dit, int coset,
           int threads) {
  if (threads <= 0) threads = omp_get_max_threads();
#define CALL(C) dispatch_fft<C>(data, logn, omega, g, inverse, dit, coset, threads)
  DISPATCH(curve, CALL)
#undef CALL
}

int oc_compute_h(int curve, void* a, void* b, void* c, int logn, const void* omega, const void* g, int threads) {
  if (threads <= 0) threads = omp_get_max_threads();
#define CALL(C) dispatch_h<C>(a, b, c, logn, omega, g, threads)
  DISPATCH(curve, CALL)
#undef CALL
}

int oc_prove(const ProveArgs* p) {
  ProveArgs q = *p;
  if (q.threads <= 0) q.threads = omp_get_max_threads();
#define CALL(C) prove_impl<C>(q)
  DISPATCH(q.curve, CALL)
#undef CALL
}

// element-wise Montgomery product (used by the oracle self-test and the CPU baseline's c = a * b)
int oc_fr_mul(int curve, const void* a, const void* b, void* out, uint64_t n) {
  const int threads = omp_get_num_procs();
#define CALL(C) fr_mul_impl<C>(a, b, out, n, 0, threads)
  DISPATCH(curve, CALL)
#undef CALL
}

// canonical -> Montgomery, element-wise
int oc_fr_to_mont(int curve, const void* a, void* out, uint64_t n, int threads) {
  if (threads <= 0) threads = omp_get_num_procs();
#define CALL(C) fr_mul_impl<C>(a, nullptr, out, n, 1, threads)
  DISPATCH(curve, CALL)
#undef CALL
}

// out[i] = [k0 + i] gen  (affine, Montgomery; gen must not be a small-order point and k0 + n must stay far below r)
int oc_gen_points(int curve, int group, const void* gen_affine, uint64_t k0, uint64_t n, void* out, int threads) {
  if (threads <= 0) threads = omp_get_num_procs();
  if (group != 1 && group != 2) return -1;
#define CALL(C) dispatch_gen<C>(group, gen_affine, k0, (size_t)n, out, threads)
  DISPATCH(curve, CALL)
#undef CALL
}

int oc_num_procs(void) { return omp_get_num_procs(); }

}  // extern "C"
