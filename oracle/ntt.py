"""NTT / quotient-polynomial oracle (big-int).  TEST INFRASTRUCTURE ONLY.

Restates gnark-crypto `fft.Domain.FFT / FFTInverse` (DIF: natural in -> bit-reversed out, DIT:
bit-reversed in -> natural out, OnCoset: shift by FrMultiplicativeGen) and gnark's `computeH`
(third-party: gnark v0.14.1-0.20260126121332-407111efab55 backend/groth16/*/prove.go,
reached from /root/reference/prover/prover_cpu.go:37).  Parity unpinned by the reference's own tests
(no golden transform vectors exist there); pinned mathematically: `tests/test_oracle_ntt.py` checks
the transforms against the O(n^2) DFT definition and h against the polynomial identity
A(x)B(x) - C(x) = h(x)(x^n - 1) at random points.
"""
from . import params as P


def bitrev(i, logn):
    r = 0
    for _ in range(logn):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def bit_reverse_list(a):
    n = len(a)
    logn = n.bit_length() - 1
    return [a[bitrev(i, logn)] for i in range(n)]


class Domain:
    """gnark-crypto fft.Domain for a curve's scalar field."""

    def __init__(self, curve: P.Curve, n: int):
        assert n & (n - 1) == 0 and n >= 1
        self.q = curve.r
        self.n = n
        self.logn = n.bit_length() - 1
        assert self.logn <= curve.two_adicity
        self.omega = pow(curve.root_of_unity, 1 << (curve.two_adicity - self.logn), self.q)
        self.omega_inv = pow(self.omega, -1, self.q)
        self.g = curve.mult_gen
        self.g_inv = pow(self.g, -1, self.q)
        self.n_inv = pow(n, -1, self.q)


def dft_natural(a, w, q):
    """A[k] = sum_j a[j] w^(jk): iterative radix-2, natural in / natural out."""
    n = len(a)
    if n == 1:
        return list(a)
    logn = n.bit_length() - 1
    A = bit_reverse_list(a)
    size = 2
    while size <= n:
        wm = pow(w, n // size, q)
        half = size // 2
        for start in range(0, n, size):
            x = 1
            for j in range(half):
                u = A[start + j]
                v = A[start + j + half] * x % q
                A[start + j] = (u + v) % q
                A[start + j + half] = (u - v) % q
                x = x * wm % q
        size *= 2
    return A


def dft_definition(a, w, q):
    n = len(a)
    return [sum(a[j] * pow(w, j * k, q) for j in range(n)) % q for k in range(n)]


def fft(a, dom: Domain, inverse=False, dit=False, coset=False):
    """gnark semantics.  DIF: `a` natural, result bit-reversed.  DIT: `a` bit-reversed, result natural."""
    q, n = dom.q, dom.n
    nat = bit_reverse_list(a) if dit else list(a)
    if not inverse:
        if coset:
            x = 1
            for j in range(n):
                nat[j] = nat[j] * x % q
                x = x * dom.g % q
        out = dft_natural(nat, dom.omega, q)
    else:
        out = dft_natural(nat, dom.omega_inv, q)
        out = [v * dom.n_inv % q for v in out]
        if coset:
            x = 1
            for j in range(n):
                out[j] = out[j] * x % q
                x = x * dom.g_inv % q
    return out if dit else bit_reverse_list(out)


def compute_h(a, b, c, dom: Domain):
    """gnark computeH: coefficients of (A*B - C)/(X^n - 1), returned in BIT-REVERSED order."""
    q, n = dom.q, dom.n
    pad = lambda v: list(v) + [0] * (n - len(v))
    a, b, c = pad(a), pad(b), pad(c)
    a = fft(a, dom, inverse=True)
    b = fft(b, dom, inverse=True)
    c = fft(c, dom, inverse=True)
    a = fft(a, dom, dit=True, coset=True)
    b = fft(b, dom, dit=True, coset=True)
    c = fft(c, dom, dit=True, coset=True)
    den = pow((pow(dom.g, n, q) - 1) % q, -1, q)
    h = [(x * y - z) * den % q for x, y, z in zip(a, b, c)]
    return fft(h, dom, inverse=True, coset=True)
