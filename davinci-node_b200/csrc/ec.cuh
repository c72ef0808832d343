// Elliptic-curve group law for y^2 = x^3 + b (a = 0 on every curve of the path), templated on the
// coordinate field F (FpT<...> for G1 and BW6-761 G2, Fp2T<...> for the other G2 groups).
//
// Buckets and partial sums live in extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2): mixed addition costs 8M+2S, the cheapest inversion-free bucket update.
// Affine points use gnark-crypto's layout {X, Y}, infinity = (0, 0)  (SURVEY.md A.4).
//
// Replaces gnark-crypto's g1JacExtended / g2JacExtended bucket arithmetic used by
// `G1Affine.MultiExp` / `G2Affine.MultiExp` under groth16.Prove
// (/root/reference/prover/prover_cpu.go:37).
#pragma once
#include "field.cuh"

namespace b200 {

template <class F>
struct alignas(16) Affine {
  typename F::El x, y;
};

template <class F>
struct alignas(16) XYZZ {
  typename F::El x, y, zz, zzz;
};

template <class F>
struct EC {
  using El = typename F::El;
  using Aff = Affine<F>;
  using Pt = XYZZ<F>;

  static __device__ __forceinline__ void set_inf(Pt& p) {
    F::set_zero(p.x);
    F::set_zero(p.y);
    F::set_zero(p.zz);
    F::set_zero(p.zzz);
  }
  static __device__ __forceinline__ bool is_inf(const Pt& p) { return F::is_zero(p.zz); }
  static __device__ __forceinline__ bool is_inf(const Aff& p) { return F::is_zero(p.x) && F::is_zero(p.y); }
  static __device__ __forceinline__ void from_affine(Pt& r, const Aff& p) {
    if (is_inf(p)) {
      set_inf(r);
      return;
    }
    r.x = p.x;
    r.y = p.y;
    F::set_one(r.zz);
    F::set_one(r.zzz);
  }
  static __device__ __forceinline__ void neg(Aff& p) { F::neg(p.y, p.y); }
  static __device__ __forceinline__ void neg(Pt& p) { F::neg(p.y, p.y); }

  // r = 2 * (affine p)          (mdbl-2008-s-1)
  static __device__ __noinline__ void dbl_affine(Pt& r, const Aff& p) {
    El u, v, w, s, m, t;
    F::dbl(u, p.y);
    F::sqr(v, u);
    F::mul(w, u, v);
    F::mul(s, p.x, v);
    F::sqr(m, p.x);
    F::dbl(t, m);
    F::add(m, m, t);          // 3 x^2
    F::sqr(r.x, m);
    F::sub(r.x, r.x, s);
    F::sub(r.x, r.x, s);
    F::sub(t, s, r.x);
    F::mul(t, m, t);
    F::mul(u, w, p.y);
    F::sub(r.y, t, u);
    r.zz = v;
    r.zzz = w;
  }

  // p = 2 * p                   (dbl-2008-s-1)
  static __device__ __noinline__ void dbl(Pt& p) {
    // no early exit for the point at infinity: ZZ' = V ZZ = 0 keeps it at infinity (Fp2 products are calls, see the
    // control-flow rule below), X and Y are zeroed at the end for hygiene
    const bool inf = is_inf(p);
    El u, v, w, s, m, t;
    F::dbl(u, p.y);
    F::sqr(v, u);
    F::mul(w, u, v);
    F::mul(s, p.x, v);
    F::sqr(m, p.x);
    F::dbl(t, m);
    F::add(m, m, t);
    F::mul(u, w, p.y);        // W * Y1 (before Y is overwritten)
    F::sqr(p.x, m);
    F::sub(p.x, p.x, s);
    F::sub(p.x, p.x, s);
    F::sub(t, s, p.x);
    F::mul(t, m, t);
    F::sub(p.y, t, u);
    F::mul(p.zz, v, p.zz);
    F::mul(p.zzz, w, p.zzz);
    if (inf) set_inf(p);
  }

  // Control-flow rule for everything below: a branch that encloses an OUT-OF-LINE call (dbl, dbl_affine, the Fp2
  // products) must be warp-uniform.  ptxas keeps warp-uniform values - stack addresses of by-reference arguments, the
  // Montgomery constant - in per-WARP uniform registers, and a callee entered by a divergent subset of a warp
  // overwrites them under the feet of the sibling threads that are still on another path of the caller (found with
  // compute-sanitizer: add() -> dbl() loaded M0 into the uniform register holding add()'s frame pointer; rewriting the
  // source as "plain ifs" does not help, the optimiser re-threads the paths).  So: the case of every thread is
  // classified first, each piece of arithmetic runs when ANY thread of the converged group needs it (warp vote), and
  // every thread keeps its own result with selects.
  enum : int { kModeSkip = 0, kModeCopy = 1, kModeGeneric = 2, kModeDouble = 3, kModeCancel = 4 };

  static __device__ __forceinline__ void csel(El& dst, const El& src, bool take) {
    uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(&src);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(El) / 4); i++) d[i] = take ? q[i] : d[i];
  }
  static __device__ __forceinline__ void csel(Pt& dst, const Pt& src, bool take) {
    csel(dst.x, src.x, take);
    csel(dst.y, src.y, take);
    csel(dst.zz, src.zz, take);
    csel(dst.zzz, src.zzz, take);
  }

  // acc += (affine p)           (madd-2008-s), all special cases handled.  The hot path of the bucket accumulation:
  // plain early exits and inline products; the (rare, out-of-line) doubling is deferred to the end and entered by vote.
  static __device__ __forceinline__ void madd(Pt& acc, const Aff& p) {
    bool need_dbl = false;
    if (!is_inf(p)) {
      if (is_inf(acc)) {
        acc.x = p.x;
        acc.y = p.y;
        F::set_one(acc.zz);
        F::set_one(acc.zzz);
      } else {
        El u2, s2, pp, ppp, q, t;
        F::mul(u2, p.x, acc.zz);
        F::mul(s2, p.y, acc.zzz);
        F::sub(u2, u2, acc.x);     // P
        F::sub(s2, s2, acc.y);     // R
        if (F::is_zero(u2)) {
          need_dbl = F::is_zero(s2);
          if (!need_dbl) set_inf(acc);
        } else {
          F::sqr(pp, u2);
          F::mul(ppp, u2, pp);
          F::mul(q, acc.x, pp);
          F::sqr(t, s2);
          F::sub(t, t, ppp);
          F::sub(t, t, q);
          F::sub(acc.x, t, q);       // X3 = R^2 - PPP - 2Q
          F::sub(q, q, acc.x);
          F::mul(q, s2, q);          // R (Q - X3)
          F::mul(t, acc.y, ppp);
          F::sub(acc.y, q, t);
          F::mul(acc.zz, acc.zz, pp);
          F::mul(acc.zzz, acc.zzz, ppp);
        }
      }
    }
    if (__any_sync(__activemask(), need_dbl)) {     // every thread of the converged group makes the call
      Pt d;
      dbl_affine(d, p);
      csel(acc, d, need_dbl);
    }
  }

  // acc += b                    (add-2008-s), all special cases handled; every branch around arithmetic is a vote
  static __device__ __noinline__ void add(Pt& acc, const Pt& b) {
    const unsigned grp = __activemask();
    const bool b_inf = is_inf(b), a_inf = is_inf(acc);
    const bool live = !b_inf && !a_inf;
    int mode = b_inf ? kModeSkip : kModeCopy;
    El u1, u2, s1, s2;
    if (__any_sync(grp, live)) {
      F::mul(u1, acc.x, b.zz);
      F::mul(u2, b.x, acc.zz);
      F::mul(s1, acc.y, b.zzz);
      F::mul(s2, b.y, acc.zzz);
      F::sub(u2, u2, u1);        // P
      F::sub(s2, s2, s1);        // R
      if (live) mode = !F::is_zero(u2) ? kModeGeneric : (F::is_zero(s2) ? kModeDouble : kModeCancel);
    }
    if (__any_sync(grp, mode == kModeGeneric)) {
      El pp, ppp, q, t;
      Pt r3;
      F::sqr(pp, u2);
      F::mul(ppp, u2, pp);
      F::mul(q, u1, pp);
      F::sqr(t, s2);
      F::sub(t, t, ppp);
      F::sub(t, t, q);
      F::sub(r3.x, t, q);
      F::sub(q, q, r3.x);
      F::mul(q, s2, q);
      F::mul(t, s1, ppp);
      F::sub(r3.y, q, t);
      F::mul(r3.zz, acc.zz, b.zz);
      F::mul(r3.zz, r3.zz, pp);
      F::mul(r3.zzz, acc.zzz, b.zzz);
      F::mul(r3.zzz, r3.zzz, ppp);
      csel(acc, r3, mode == kModeGeneric);
    }
    csel(acc, b, mode == kModeCopy);
    if (mode == kModeCancel) set_inf(acc);
    if (__any_sync(grp, mode == kModeDouble)) {
      Pt d = acc;
      dbl(d);
      csel(acc, d, mode == kModeDouble);
    }
  }

  // acc += b with plain early exits: ONLY for callers whose warp-mates cannot be on a sibling path - the one-thread
  // proof-assembly kernels and the scalar-multiplication helpers below (control-flow rule above)
  static __device__ __noinline__ void add_st(Pt& acc, const Pt& b) {
    if (is_inf(b)) return;
    if (is_inf(acc)) {
      acc = b;
      return;
    }
    El u1, u2, s1, s2, pp, ppp, q, t;
    F::mul(u1, acc.x, b.zz);
    F::mul(u2, b.x, acc.zz);
    F::mul(s1, acc.y, b.zzz);
    F::mul(s2, b.y, acc.zzz);
    F::sub(u2, u2, u1);        // P
    F::sub(s2, s2, s1);        // R
    if (F::is_zero(u2)) {
      if (F::is_zero(s2)) {
        dbl(acc);
      } else {
        set_inf(acc);
      }
      return;
    }
    F::sqr(pp, u2);
    F::mul(ppp, u2, pp);
    F::mul(q, u1, pp);
    F::sqr(t, s2);
    F::sub(t, t, ppp);
    F::sub(t, t, q);
    F::sub(acc.x, t, q);
    F::sub(q, q, acc.x);
    F::mul(q, s2, q);
    F::mul(t, s1, ppp);
    F::sub(acc.y, q, t);
    F::mul(acc.zz, acc.zz, b.zz);
    F::mul(acc.zz, acc.zz, pp);
    F::mul(acc.zzz, acc.zzz, b.zzz);
    F::mul(acc.zzz, acc.zzz, ppp);
  }

  // r = [k] p for a small unsigned k (double-and-add, MSB first)
  static __device__ __noinline__ void mul_u32(Pt& r, const Pt& p, uint32_t k) {
    Pt acc;
    set_inf(acc);
    for (int i = 31; i >= 0; i--) {
      dbl(acc);
      if ((k >> i) & 1) add_st(acc, p);
    }
    r = acc;
  }

  // r = [s] p for a canonical (non-Montgomery) little-endian scalar of NS limbs
  template <int NS>
  static __device__ __noinline__ void mul_scalar(Pt& r, const Pt& p, const uint32_t* s) {
    Pt acc;
    set_inf(acc);
    bool started = false;
    for (int i = NS * 32 - 1; i >= 0; i--) {
      if (started) dbl(acc);
      if ((s[i >> 5] >> (i & 31)) & 1) {
        add_st(acc, p);
        started = true;
      }
    }
    r = acc;
  }

  // r = [s] p, fixed 4-bit windows: a table of 1..15 multiples (14 additions), then 4 doublings and at most one
  // addition per window - a quarter of the additions of double-and-add.  Used by the proof assembly (one thread
  // per multiplication; the table lives in local memory).
  template <int NS>
  static __device__ __noinline__ void mul_scalar_w4(Pt& r, const Pt& p, const uint32_t* s) {
    Pt tab[16];
    set_inf(tab[0]);
    tab[1] = p;
    for (int k = 2; k < 16; k++) {
      tab[k] = tab[k >> 1];
      if (k & 1) {
        tab[k] = tab[k - 1];
        add_st(tab[k], p);
      } else {
        dbl(tab[k]);
      }
    }
    Pt acc;
    set_inf(acc);
    bool started = false;
    for (int i = NS * 8 - 1; i >= 0; i--) {
      if (started) {
        dbl(acc);
        dbl(acc);
        dbl(acc);
        dbl(acc);
      }
      uint32_t d = (s[i >> 3] >> ((i & 7) * 4)) & 15u;
      if (d) {
        add_st(acc, tab[d]);
        started = true;
      }
    }
    r = acc;
  }

  // affine normalisation: x = X/ZZ, y = Y/ZZZ ; infinity -> (0, 0).  Inlined (the inversion inside
  // is the only out-of-line call) and written through a local with a single exit.
  // SINGLE_THREAD: the caller is a one-thread kernel (proof assembly) - use the low-latency binary-GCD inversion
  template <bool SINGLE_THREAD = false>
  static __device__ __forceinline__ void to_affine(Aff& r, const Pt& p) {
    Aff o;
    F::set_zero(o.x);
    F::set_zero(o.y);
    if (!is_inf(p)) {
      // 1/ZZ and 1/ZZZ from a single inversion of ZZ*ZZZ
      El t, ti, izz, izzz;
      F::mul(t, p.zz, p.zzz);
      if (SINGLE_THREAD) F::inv_bin(ti, t);
      else F::inv(ti, t);
      F::mul(izz, ti, p.zzz);
      F::mul(izzz, ti, p.zz);
      F::mul(o.x, p.x, izz);
      F::mul(o.y, p.y, izzz);
    }
    r = o;
  }
};

}  // namespace b200
