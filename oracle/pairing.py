"""Big-integer pairings for the four curves of the path.  TEST INFRASTRUCTURE ONLY (see oracle/params.py header).

What the reference does with pairings on this path:
  * every proof is verified right after proving (/root/reference/circuits/artifacts.go:595-613 -> groth16.Verify),
  * the on-chain verifier of the statetransition proofs is /root/reference/config/statetransition_vkey.sol:653-746
    (EIP-197 pairing-check precompile),
  * KZG openings are checked against [tau]_2 (/root/reference/crypto/blobs/kzg.go:26-45).
All three are *product-of-pairings equals one* checks, so any non-degenerate bilinear pairing on (G1, G2) decides them
identically.  This file therefore implements the simplest one, the reduced Tate pairing

    t(P, Q) = f_{r,P}(psi(Q)) ^ ((p^k - 1) / r),      P in G1 over Fp,  Q in G2 on the sextic twist,

with F_{p^k} = Fp[w] / (modulus) as plain polynomials and psi the untwisting isomorphism - not gnark's optimal-ate
Miller loops (third-party: gnark-crypto v0.19.3 ecc/<curve>/pairing.go, go.mod:16), whose *values* differ by a fixed
exponent but whose product checks agree.

Towers (gnark-crypto's):
  BN254      Fp2 = Fp[u]/(u^2+1),  w^6 = 9+u   -> w^12 - 18 w^6 + 82,  D-twist  b' = 3/(9+u)
  BLS12-381  Fp2 = Fp[u]/(u^2+1),  w^6 = 1+u   -> w^12 -  2 w^6 +  2,  M-twist  b' = 4(1+u)
  BLS12-377  Fp2 = Fp[u]/(u^2+5),  w^6 = u     -> w^12 + 5,            D-twist  b' = 1/u
  BW6-761    k = 6, G2 over Fp,    w^6 = -4    -> w^6 + 4,             M-twist  b' = 4   (b = -1)

Pinned by: bilinearity / non-degeneracy (tests/test_oracle_pairing.py), the EIP-4844 ceremony relation
e([tau]_1, G_2) = e(G_1, [tau]_2) between the reference's SRS file and crypto/blobs/kzg.go, and the subgroup / twist
membership of every G2 constant of config/statetransition_vkey.sol.
"""
from . import curve as C

# name -> (k, m_half, m_0, u_as_poly (coefficient of w^0, coefficient of w^(k/2)) or None, twist type)
#   modulus = w^k + m_half w^(k/2) + m_0 ;  u = u0 + u1 w^6
_TOWERS = {
    "bn254": (12, -18, 82, (-9, 1), "D"),
    "bls12_381": (12, -2, 2, (-1, 1), "M"),
    "bls12_377": (12, 0, 5, (0, 1), "D"),
    "bw6_761": (6, 0, 4, None, "M"),
}


class ExtField:
    """F_{p^k} = Fp[w] / (w^k + m_half w^(k/2) + m_0); elements are k-tuples of ints (little-endian in w)."""

    def __init__(self, p, k, m_half, m_0):
        self.p, self.k, self.mh, self.m0 = p, k, m_half % p, m_0 % p
        self.one = (1,) + (0,) * (k - 1)
        self.zero = (0,) * k

    def add(self, a, b):
        p = self.p
        return tuple((x + y) % p for x, y in zip(a, b))

    def sub(self, a, b):
        p = self.p
        return tuple((x - y) % p for x, y in zip(a, b))

    def scale(self, a, c):
        p = self.p
        return tuple(x * c % p for x in a)

    def mul(self, a, b):
        k, p, h = self.k, self.p, self.k // 2
        t = [0] * (2 * k - 1)
        for i, ai in enumerate(a):
            if ai:
                for j, bj in enumerate(b):
                    if bj:
                        t[i + j] += ai * bj
        mh, m0 = self.mh, self.m0
        for d in range(2 * k - 2, k - 1, -1):       # w^d = -(mh w^(d - k/2) + m0 w^(d - k))
            c = t[d] % p
            if c:
                if mh:
                    t[d - h] -= c * mh
                t[d - k] -= c * m0
        return tuple(x % p for x in t[:k])

    def sqr(self, a):
        return self.mul(a, a)

    def pow(self, a, e):
        acc = self.one
        for bit in bin(e)[2:]:
            acc = self.mul(acc, acc)
            if bit == "1":
                acc = self.mul(acc, a)
        return acc


class Pairing:
    def __init__(self, name):
        self.cx = cx = C.ctx(name)
        self.name = name
        k, mh, m0, self.u_poly, self.twist = _TOWERS[name]
        self.k = k
        self.F = ExtField(cx.p, k, mh, m0)
        self.r = cx.r
        self.final_exp = (cx.p ** k - 1) // cx.r
        assert (cx.p ** k - 1) % cx.r == 0
        p = cx.p
        # w^2, w^3 and their inverses as field elements (untwisting factors)
        F = self.F
        w = (0, 1) + (0,) * (k - 2)
        self.w2 = F.mul(w, w)
        self.w3 = F.mul(self.w2, w)
        # 1/w = -(w^(k-1) + mh w^(k/2-1)) / m0
        inv_m0 = pow(m0 % p, -1, p)
        winv = [0] * k
        winv[k - 1] = (-inv_m0) % p
        winv[k // 2 - 1] = (-mh * inv_m0) % p
        winv = tuple(winv)
        assert F.mul(winv, w) == F.one
        self.w2i = F.mul(winv, winv)
        self.w3i = F.mul(self.w2i, winv)

    # ---- embeddings
    def embed_base(self, c):
        """Coordinate of a twist point (Fp2 tuple, or int for BW6-761) -> F_{p^k}."""
        k, p = self.k, self.cx.p
        out = [0] * k
        if self.u_poly is None:
            out[0] = c % p
        else:
            c0, c1 = c
            u0, u1 = self.u_poly
            out[0] = (c0 + c1 * u0) % p
            out[k // 2] = (c1 * u1) % p
        return tuple(out)

    def untwist(self, Q):
        """psi: E'(F_{p^(k/6)}) -> E(F_{p^k}).  D-type: (x w^2, y w^3); M-type: (x / w^2, y / w^3)."""
        F = self.F
        x, y = self.embed_base(Q[0]), self.embed_base(Q[1])
        if self.twist == "D":
            return F.mul(x, self.w2), F.mul(y, self.w3)
        return F.mul(x, self.w2i), F.mul(y, self.w3i)

    # ---- Miller loop f_{r,P}(Q), P affine over Fp, Q = (xq, yq) in F_{p^k}; vertical lines dropped
    def miller(self, P, Q):
        F, p = self.F, self.cx.p
        if P is None or Q is None:
            return F.one
        xq, yq = self.untwist(Q)
        xp, yp = P
        f = F.one
        tx, ty = xp, yp

        def line(lam, x0, y0):
            # l(Q) = (yq - y0) - lam (xq - x0)
            v = list(F.sub(yq, F.scale(xq, lam)))
            v[0] = (v[0] - y0 + lam * x0) % p
            return tuple(v)

        inf = False
        for bit in bin(self.r)[3:]:
            f = F.sqr(f)
            if not inf:
                if ty == 0:
                    inf = True          # vertical tangent
                else:
                    lam = 3 * tx * tx * pow(2 * ty, -1, p) % p
                    f = F.mul(f, line(lam, tx, ty))
                    nx = (lam * lam - 2 * tx) % p
                    ty = (lam * (tx - nx) - ty) % p
                    tx = nx
            if bit == "1":
                if inf:
                    tx, ty, inf = xp, yp, False
                elif tx == xp:
                    if (ty + yp) % p == 0:
                        inf = True      # vertical line through T and -T
                    else:
                        lam = 3 * tx * tx * pow(2 * ty, -1, p) % p
                        f = F.mul(f, line(lam, tx, ty))
                        nx = (lam * lam - 2 * tx) % p
                        ty = (lam * (tx - nx) - ty) % p
                        tx = nx
                else:
                    lam = (ty - yp) * pow(tx - xp, -1, p) % p
                    f = F.mul(f, line(lam, tx, ty))
                    nx = (lam * lam - tx - xp) % p
                    ty = (lam * (tx - nx) - ty) % p
                    tx = nx
        assert inf, "P is not in the order-r subgroup"
        return f

    def pair(self, P, Q):
        """Reduced Tate pairing t(P, Q) as an F_{p^k} element (1 when either argument is the point at infinity)."""
        return self.F.pow(self.miller(P, Q), self.final_exp)

    def product_is_one(self, pairs):
        """prod_i t(P_i, Q_i) == 1 - the EIP-197 pairing-check / gnark PairingCheck predicate."""
        F = self.F
        acc = F.one
        for P, Q in pairs:
            acc = F.mul(acc, self.miller(P, Q))
        return F.pow(acc, self.final_exp) == F.one


_CACHE = {}


def get(name) -> Pairing:
    if name not in _CACHE:
        _CACHE[name] = Pairing(name)
    return _CACHE[name]
