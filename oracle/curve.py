"""Big-integer field / elliptic-curve arithmetic: the ground-truth oracle.

TEST INFRASTRUCTURE ONLY (see oracle/params.py header).  Restates, with plain
Python integers, the group law the reference reaches through gnark-crypto
(`ecc/<curve>` G1Affine/G2Affine/G1Jac, go.mod:16) from
/root/reference/prover/prover_cpu.go:37 (`groth16.Prove`).  Proof elements are
canonical affine points, so any correct group law yields gnark's bytes
 - this file therefore uses the textbook affine / Jacobian
formulas rather than mimicking gnark's extended-Jacobian bucket code.

Points: affine tuples (x, y) or None for infinity.  Field elements: int for Fp,
(c0, c1) tuples for Fp2 = Fp[u]/(u^2 - nonresidue).
"""
import json
import os
from math import isqrt

from . import params as P


# ----------------------------------------------------------------------------- field ops
class FpOps:
    deg = 1

    def __init__(self, p):
        self.p = p
        self.zero = 0
        self.one = 1

    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def neg(self, a): return (-a) % self.p
    def mul(self, a, b): return (a * b) % self.p
    def sqr(self, a): return (a * a) % self.p
    def inv(self, a): return pow(a, -1, self.p)
    def is_zero(self, a): return a % self.p == 0
    def muli(self, a, k): return (a * k) % self.p
    def from_int(self, k): return k % self.p
    def eq(self, a, b): return (a - b) % self.p == 0

    def sqrt(self, a):
        return sqrt_mod(a, self.p)


class Fp2Ops:
    deg = 2

    def __init__(self, p, nonresidue):
        self.p = p
        self.nr = nonresidue % p
        self.zero = (0, 0)
        self.one = (1, 0)

    def add(self, a, b): return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)
    def sub(self, a, b): return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)
    def neg(self, a): return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.nr * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a): return self.mul(a, a)

    def inv(self, a):
        p = self.p
        n = pow((a[0] * a[0] - self.nr * a[1] * a[1]) % p, -1, p)
        return ((a[0] * n) % p, (-a[1] * n) % p)

    def is_zero(self, a): return a[0] % self.p == 0 and a[1] % self.p == 0
    def muli(self, a, k): return ((a[0] * k) % self.p, (a[1] * k) % self.p)
    def from_int(self, k): return (k % self.p, 0)
    def eq(self, a, b): return (a[0] - b[0]) % self.p == 0 and (a[1] - b[1]) % self.p == 0

    def sqrt(self, a):
        """Square root in Fp2 via the norm trick; returns None if a is a non-residue."""
        p = self.p
        if self.is_zero(a):
            return (0, 0)
        a0, a1 = a
        if a1 % p == 0:
            s = sqrt_mod(a0, p)
            if s is not None:
                return (s, 0)
            # sqrt(a0) = u * sqrt(a0 / nr)
            s = sqrt_mod(a0 * pow(self.nr, -1, p) % p, p)
            return None if s is None else (0, s)
        norm = (a0 * a0 - self.nr * a1 * a1) % p
        n = sqrt_mod(norm, p)
        if n is None:
            return None
        inv2 = pow(2, -1, p)
        for nn in (n, p - n):
            x2 = (a0 + nn) * inv2 % p
            x = sqrt_mod(x2, p)
            if x is None or x == 0:
                continue
            y = a1 * pow(2 * x, -1, p) % p
            if self.eq(self.mul((x, y), (x, y)), a):
                return (x, y)
        return None


def sqrt_mod(a, p):
    """Tonelli-Shanks; returns a root or None."""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return r


# ----------------------------------------------------------------------------- group law
class Group:
    """Short-Weierstrass curve y^2 = x^3 + b over field ops F (a = 0 for every curve on the path)."""

    def __init__(self, F, b, name=""):
        self.F = F
        self.b = b
        self.name = name

    def on_curve(self, pt):
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.eq(F.sqr(y), F.add(F.mul(F.sqr(x), x), self.b))

    def neg(self, pt):
        return None if pt is None else (pt[0], self.F.neg(pt[1]))

    # --- affine (used for small cases and as the definition)
    def add(self, p1, p2):
        F = self.F
        if p1 is None:
            return p2
        if p2 is None:
            return p1
        x1, y1 = p1
        x2, y2 = p2
        if F.eq(x1, x2):
            if F.eq(y1, y2) and not F.is_zero(y1):
                lam = F.mul(F.muli(F.sqr(x1), 3), F.inv(F.muli(y1, 2)))
            else:
                return None
        else:
            lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    # --- Jacobian (X, Y, Z), Z == 0 -> infinity; used for speed
    def jac_from_affine(self, pt):
        F = self.F
        return (F.one, F.one, F.zero) if pt is None else (pt[0], pt[1], F.one)

    def jac_to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jac_double(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        D = F.muli(F.sub(F.sub(F.sqr(F.add(X, B)), A), C), 2)
        E = F.muli(A, 3)
        Fq = F.sqr(E)
        X3 = F.sub(Fq, F.muli(D, 2))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muli(C, 8))
        Z3 = F.muli(F.mul(Y, Z), 2)
        return (X3, Y3, Z3)

    def jac_add(self, J1, J2):
        F = self.F
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        if F.is_zero(Z1):
            return J2
        if F.is_zero(Z2):
            return J1
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if F.eq(U1, U2):
            if F.eq(S1, S2):
                return self.jac_double(J1)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        R = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(R), HHH), F.muli(V, 2))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def jac_mixed_add(self, J1, pt):
        if pt is None:
            return J1
        return self.jac_add(J1, (pt[0], pt[1], self.F.one))

    def mul(self, pt, k):
        """Scalar multiplication [k]pt (k any integer)."""
        if pt is None or k == 0:
            return None
        if k < 0:
            return self.mul(self.neg(pt), -k)
        F = self.F
        acc = (F.one, F.one, F.zero)
        base = (pt[0], pt[1], F.one)
        for bit in bin(k)[2:]:
            acc = self.jac_double(acc)
            if bit == "1":
                acc = self.jac_add(acc, base)
        return self.jac_to_affine(acc)

    def msm_naive(self, points, scalars):
        """Definition of the multi-exponentiation: sum_i [s_i] P_i, affine."""
        acc = self.jac_from_affine(None)
        for pt, s in zip(points, scalars):
            if s and pt is not None:
                acc = self.jac_add(acc, self.jac_from_affine(self.mul(pt, s)))
        return self.jac_to_affine(acc)

    def msm(self, points, scalars, c=None):
        """Bucket-method MSM (same value as msm_naive; faster for n >~ 64)."""
        n = len(points)
        if n < 32:
            return self.msm_naive(points, scalars)
        if c is None:
            c = max(2, min(16, n.bit_length() - 3))
        nbits = max(1, max((int(s).bit_length() for s in scalars), default=1))
        F = self.F
        inf = (F.one, F.one, F.zero)
        total = inf
        nwin = (nbits + c - 1) // c
        for w in reversed(range(nwin)):
            for _ in range(c):
                total = self.jac_double(total)
            buckets = [inf] * (1 << c)
            sh = w * c
            mask = (1 << c) - 1
            for pt, s in zip(points, scalars):
                d = (int(s) >> sh) & mask
                if d and pt is not None:
                    buckets[d] = self.jac_mixed_add(buckets[d], pt)
            run = inf
            acc = inf
            for d in range((1 << c) - 1, 0, -1):
                run = self.jac_add(run, buckets[d])
                acc = self.jac_add(acc, run)
            total = self.jac_add(total, acc)
        return self.jac_to_affine(total)

    def sum(self, points):
        acc = self.jac_from_affine(None)
        for pt in points:
            acc = self.jac_mixed_add(acc, pt)
        return self.jac_to_affine(acc)


# ----------------------------------------------------------------------------- curve objects
def _cm_trace(p):
    """(t, f) with 4p = t^2 + 3 f^2 (j = 0 curves), by Cornacchia."""
    s = sqrt_mod(p - 3, p)
    assert s is not None
    if s % 2 == 0:
        s = p - s
    a, b = 2 * p, s
    lim = isqrt(4 * p)
    while b > lim:
        a, b = b, a % b
    t = b
    rem = 4 * p - t * t
    assert rem % 3 == 0
    f = isqrt(rem // 3)
    assert f * f * 3 == rem
    return t, f


def _twist_orders(q, t, f):
    """All six possible group orders of j=0 curves over F_q given 4q = t^2 + 3f^2."""
    traces = {t, -t}
    for st in (1, -1):
        for sf in (1, -1):
            v = st * t + sf * 3 * f
            assert v % 2 == 0
            traces.add(v // 2)
    return [q + 1 - tr for tr in traces]


def _find_generator(G: Group, r, q, t, f, seed_x=1):
    """Deterministically derive a point of exact order r on G (cofactor clearing)."""
    F = G.F
    orders = [n for n in _twist_orders(q, t, f) if n % r == 0]
    x = seed_x
    while True:
        xe = F.from_int(x) if F.deg == 1 else (x, 1)
        rhs = F.add(F.mul(F.sqr(xe), xe), G.b)
        y = F.sqrt(rhs)
        x += 1
        if y is None:
            continue
        # canonical choice: lexicographically smaller y
        ny = F.neg(y)
        y = min(y, ny) if F.deg == 1 else min(y, ny, key=lambda v: (v[1], v[0]))
        pt = (xe, y)
        for n in orders:
            if G.mul(pt, n) is None:
                g = G.mul(pt, n // r)
                if g is not None and G.mul(g, r) is None:
                    return g
        # this x gave a point whose order is on another twist - cannot happen for a fixed
        # curve, but keep searching to stay robust
    raise AssertionError


_CACHE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "generators.json")


class CurveCtx:
    """Bundles G1 / G2 groups and generators for one curve."""

    def __init__(self, c: P.Curve, cache):
        self.c = c
        self.name = c.name
        self.p, self.r = c.p, c.r
        self.Fp = FpOps(c.p)
        self.Fr = FpOps(c.r)
        self.G1 = Group(self.Fp, c.b % c.p, c.name + ".G1")
        if c.g2_degree == 2:
            self.F2 = Fp2Ops(c.p, c.nonresidue)
            if c.name == "bn254":
                b2 = self.F2.mul((3, 0), self.F2.inv((9, 1)))          # D-twist 3/(9+u)
            elif c.name == "bls12_377":
                b2 = self.F2.inv((0, 1))                                # D-twist 1/u
            else:
                b2 = tuple(v % c.p for v in c.b2)
        else:
            self.F2 = self.Fp
            b2 = c.b2[0] % c.p
        self.G2 = Group(self.F2, b2, c.name + ".G2")

        ent = cache.get(c.name, {})
        if c.g1 is not None:
            self.g1 = c.g1
        elif "g1" in ent:
            self.g1 = tuple(int(v, 16) for v in ent["g1"])
        else:
            t, f = _cm_trace(c.p)
            self.g1 = _find_generator(self.G1, c.r, c.p, t, f)
            ent["g1"] = [hex(v) for v in self.g1]
        if "g2" in ent:
            g = ent["g2"]
            if c.g2_degree == 2:
                self.g2 = ((int(g[0], 16), int(g[1], 16)), (int(g[2], 16), int(g[3], 16)))
            else:
                self.g2 = (int(g[0], 16), int(g[1], 16))
        else:
            t, f = _cm_trace(c.p)
            if c.g2_degree == 2:
                q, t2, f2 = c.p * c.p, t * t - 2 * c.p, t * f
                self.g2 = _find_generator(self.G2, c.r, q, t2, f2)
                ent["g2"] = [hex(self.g2[0][0]), hex(self.g2[0][1]), hex(self.g2[1][0]), hex(self.g2[1][1])]
            else:
                self.g2 = _find_generator(self.G2, c.r, c.p, t, f)
                ent["g2"] = [hex(v) for v in self.g2]
        cache[c.name] = ent
        assert self.G1.on_curve(self.g1) and self.G2.on_curve(self.g2)

    def group(self, which):
        return self.G1 if which == 1 else self.G2

    def gen(self, which):
        return self.g1 if which == 1 else self.g2


_CTX = {}


def ctx(name) -> CurveCtx:
    """Curve context by name ('bn254', 'bls12_377', 'bls12_381', 'bw6_761') or id."""
    if isinstance(name, int):
        name = P.BY_ID[name].name
    if name not in _CTX:
        cache = {}
        if os.path.exists(_CACHE_PATH):
            with open(_CACHE_PATH) as fh:
                cache = json.load(fh)
        before = json.dumps(cache, sort_keys=True)
        _CTX[name] = CurveCtx(P.CURVES[name], cache)
        if json.dumps(cache, sort_keys=True) != before:
            with open(_CACHE_PATH, "w") as fh:
                json.dump(cache, fh, indent=1, sort_keys=True)
    return _CTX[name]
