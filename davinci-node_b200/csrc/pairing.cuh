// Reduced Tate pairing and product-of-pairings check for the four curves of the path.
//
// What it replaces: the pairing check inside gnark's groth16.Verify, which the reference runs right after every proof
// (/root/reference/circuits/artifacts.go:595-613), i.e. gnark-crypto's <curve>.PairingCheck (third-party, go.mod:16), and
// the EIP-197 precompile call of /root/reference/config/statetransition_vkey.sol:720-746.  Those are all
// "prod_i e(P_i, Q_i) == 1" predicates, which every non-degenerate bilinear pairing on (G1, G2) decides identically, so
// this file implements the one with the least machinery:
//
//     t(P, Q) = f_{r,P}(psi(Q)) ^ ((p^k - 1) / r),      P in G1 over Fp,  Q in G2 on the sextic twist,
//
// with F_{p^k} = Fp[w] / (w^k + MH w^(k/2) + M0) held as k coefficients over Fp (schoolbook products, no tower, no
// Frobenius tables), psi the untwisting map, affine Miller steps (one binary-GCD inversion each) and a plain
// square-and-multiply final exponentiation.  One thread per pairing: this is the FUNCTIONAL verifier of SURVEY.md 8(f),
// not a throughput kernel (a Groth16 verification is four Miller loops and one exponentiation next to a proof's ~10^9
// field products).  gnark's optimal-ate values differ from these by a fixed exponent; the product checks agree.
//
// The code is written against the "field concept" of field.cuh (El, add, sub, mul, neg, mul_small, is_zero, eq,
// set_zero, set_one, inv_bin, inv) and is compiled twice: for the device inside curve_impl.cuh, and - with B200_PQ = __host__
// and a portable Montgomery field - by tests/host_pairing.cu, where every function below is checked bit for bit against
// the big-integer pairing of the test oracle on the CPU.
#pragma once
#include <stdint.h>

#include "gen/pairing_consts.cuh"

#ifndef B200_PQ
#define B200_PQ __device__
#endif

namespace b200 {

// F: prime field of the coordinates; T: extension shape (gen/pairing_consts.cuh); RP: parameter struct of the scalar
// field (its modulus r drives the Miller loop); FERMAT: invert by a^(p-2) (fixed instruction stream: the choice when the
// lanes of a warp run independent pairings) instead of the binary GCD (data-dependent loops: fastest for one thread)
template <class F, class T, class RP, bool FERMAT = false>
struct PairingT {
  using El = typename F::El;
  static constexpr int K = T::K;
  static constexpr int H = T::K / 2;
  static constexpr int NQ = T::U1 ? 2 : 1;      // Fp coordinates per twist coordinate
  struct Ext {
    El c[K];
  };

  static B200_PQ void inv(El& r, const El& a) {
    if (FERMAT) {
      F::inv(r, a);
    } else {
      F::inv_bin(r, a);
    }
  }
  static B200_PQ void ext_zero(Ext& r) {
    for (int i = 0; i < K; i++) F::set_zero(r.c[i]);
  }
  static B200_PQ void ext_one(Ext& r) {
    ext_zero(r);
    F::set_one(r.c[0]);
  }
  static B200_PQ bool ext_is_one(const Ext& a) {
    El one;
    F::set_one(one);
    bool ok = F::eq(a.c[0], one);
    for (int i = 1; i < K; i++) ok = ok && F::is_zero(a.c[i]);
    return ok;
  }
  // r = v * a for a small signed integer v
  static B200_PQ void scale_small(El& r, const El& a, int v) {
    El t;
    F::mul_small(t, a, v < 0 ? -v : v);
    if (v < 0) F::neg(t, t);
    r = t;
  }
  // r = a * b mod (w^K + MH w^H + M0)
  static B200_PQ __noinline__ void ext_mul(Ext& r, const Ext& a, const Ext& b) {
    El t[2 * K - 1];
    bool bz[K];
    for (int i = 0; i < 2 * K - 1; i++) F::set_zero(t[i]);
    for (int j = 0; j < K; j++) bz[j] = F::is_zero(b.c[j]);
    for (int i = 0; i < K; i++) {
      if (F::is_zero(a.c[i])) continue;
      for (int j = 0; j < K; j++) {
        if (bz[j]) continue;
        El m;
        F::mul(m, a.c[i], b.c[j]);
        F::add(t[i + j], t[i + j], m);
      }
    }
    // w^d = -(MH w^(d-H) + M0 w^(d-K)), highest degree first
    for (int d = 2 * K - 2; d >= K; d--) {
      if (F::is_zero(t[d])) continue;
      El m;
      if (T::MH != 0) {
        scale_small(m, t[d], T::MH);
        F::sub(t[d - H], t[d - H], m);
      }
      scale_small(m, t[d], T::M0);
      F::sub(t[d - K], t[d - K], m);
    }
    for (int i = 0; i < K; i++) r.c[i] = t[i];
  }
  // a ^ ((p^K - 1) / r), most significant bit first
  static B200_PQ __noinline__ void final_exp(Ext& r, const Ext& a) {
    Ext acc;
    ext_one(acc);
    for (int i = T::FE_BITS - 1; i >= 0; i--) {
      ext_mul(acc, acc, acc);
      if ((T::final_exp(i >> 5) >> (i & 31)) & 1u) ext_mul(acc, acc, a);
    }
    r = acc;
  }

  // twist coordinate (NQ Fp limbs groups; Fp2 as c0 + c1 u with u = U0 + U1 w^H) -> F_{p^K}
  static B200_PQ void embed(Ext& out, const El* c) {
    ext_zero(out);
    if (T::U1 == 0) {
      out.c[0] = c[0];
    } else {
      El t;
      scale_small(t, c[NQ - 1], T::U0);
      F::add(out.c[0], c[0], t);
      scale_small(out.c[H], c[NQ - 1], T::U1);
    }
  }
  // psi: D-type twist (x w^2, y w^3); M-type twist (x / w^2, y / w^3)
  static B200_PQ void untwist(Ext& xq, Ext& yq, const El* Q) {
    Ext ex, ey, w1, w2, w3;
    embed(ex, Q);
    embed(ey, Q + NQ);
    ext_zero(w1);
    if (T::DTWIST) {
      F::set_one(w1.c[1]);
    } else {
      // 1 / w = -(w^(K-1) + MH w^(H-1)) / M0
      El one, m0, im0;
      F::set_one(one);
      scale_small(m0, one, T::M0);
      inv(im0, m0);
      F::neg(w1.c[K - 1], im0);
      if (T::MH != 0) {
        scale_small(m0, im0, T::MH);
        F::neg(w1.c[H - 1], m0);
      }
    }
    ext_mul(w2, w1, w1);
    ext_mul(w3, w2, w1);
    ext_mul(xq, ex, w2);
    ext_mul(yq, ey, w3);
  }

  // l(Q) = (yq - y0) - lam (xq - x0)
  static B200_PQ void line(Ext& l, const Ext& xq, const Ext& yq, const El& lam, const El& x0, const El& y0) {
    for (int i = 0; i < K; i++) {
      if (F::is_zero(xq.c[i])) {
        l.c[i] = yq.c[i];
      } else {
        El m;
        F::mul(m, xq.c[i], lam);
        F::sub(l.c[i], yq.c[i], m);
      }
    }
    El m;
    F::mul(m, lam, x0);
    F::sub(m, m, y0);
    F::add(l.c[0], l.c[0], m);
  }
  // tangent step at T = (tx, ty), ty != 0: f *= l_{T,T}(Q), T = 2T
  static B200_PQ void step_tangent(Ext& f, El& tx, El& ty, const Ext& xq, const Ext& yq) {
    El lam, num, den, nx, t;
    F::mul(num, tx, tx);
    F::mul_small(num, num, 3);
    F::add(den, ty, ty);
    inv(den, den);
    F::mul(lam, num, den);
    Ext l;
    line(l, xq, yq, lam, tx, ty);
    ext_mul(f, f, l);
    F::mul(nx, lam, lam);
    F::sub(nx, nx, tx);
    F::sub(nx, nx, tx);
    F::sub(t, tx, nx);
    F::mul(t, lam, t);
    F::sub(ty, t, ty);
    tx = nx;
  }
  // chord step through T and P (tx != xp): f *= l_{T,P}(Q), T = T + P
  static B200_PQ void step_chord(Ext& f, El& tx, El& ty, const El& xp, const El& yp, const Ext& xq, const Ext& yq) {
    El lam, num, den, nx, t;
    F::sub(num, ty, yp);
    F::sub(den, tx, xp);
    inv(den, den);
    F::mul(lam, num, den);
    Ext l;
    line(l, xq, yq, lam, tx, ty);
    ext_mul(f, f, l);
    F::mul(nx, lam, lam);
    F::sub(nx, nx, tx);
    F::sub(nx, nx, xp);
    F::sub(t, tx, nx);
    F::mul(t, lam, t);
    F::sub(ty, t, ty);
    tx = nx;
  }
  // Miller function f_{r,P}(psi(Q)) (vertical lines dropped: they lie in a proper subfield and die in the final
  // exponentiation).  P = {x, y}, Q = {x (NQ limbs groups), y}; the all-zero encoding is the point at infinity (gnark's
  // affine convention) and gives f = 1.  Returns false when r P != infinity (P outside the order-r subgroup).
  static B200_PQ __noinline__ bool miller(Ext& f, const El* P, const El* Q) {
    ext_one(f);
    bool p_inf = F::is_zero(P[0]) && F::is_zero(P[1]);
    bool q_inf = true;
    for (int i = 0; i < 2 * NQ; i++) q_inf = q_inf && F::is_zero(Q[i]);
    if (p_inf || q_inf) return true;
    Ext xq, yq;
    untwist(xq, yq, Q);
    const El xp = P[0], yp = P[1];
    El tx = xp, ty = yp;
    bool inf = false;
    for (int i = RP::BITS - 2; i >= 0; i--) {
      ext_mul(f, f, f);
      if (!inf) {
        if (F::is_zero(ty)) {
          inf = true;                              // vertical tangent
        } else {
          step_tangent(f, tx, ty, xq, yq);
        }
      }
      if ((RP::modulus(i >> 5) >> (i & 31)) & 1u) {
        if (inf) {
          tx = xp;
          ty = yp;
          inf = false;
        } else if (F::eq(tx, xp)) {
          El s;
          F::add(s, ty, yp);
          if (F::is_zero(s)) {
            inf = true;                            // vertical line through T and -T
          } else {
            step_tangent(f, tx, ty, xq, yq);
          }
        } else {
          step_chord(f, tx, ty, xp, yp, xq, yq);
        }
      }
    }
    return inf;
  }
};

}  // namespace b200
