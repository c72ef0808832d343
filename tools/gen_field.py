#!/usr/bin/env python3
"""Generator for the sm_100a Montgomery field kernels (32-bit limbs, IMAD.WIDE carry chains).

Every field operation is built as a straight-line program over a tiny IR whose
ops are exactly PTX integer instructions with explicit carry-flag semantics
(`mad.lo.cc.u32`, `madc.hi.cc.u32`, `addc.cc.u32`, ...).  The same op list is

  * emitted as ONE inline-asm block per field operation (the carry flag never
    lives across an asm boundary, moduli are immediates so they cost no
    registers), and
  * interpreted in Python (`run`) so the exact instruction sequence is checked
    against big-integer arithmetic on a machine without a GPU
    (tests/test_field_ir.py).

Layout contract: an element is N little-endian uint32 limbs holding x*R mod p,
R = 2^(32N) - bit-identical to gnark-crypto's `[N/2]uint64` Montgomery elements
(SURVEY.md A.4), so proving-key memory is consumed without conversion.

Usage:  python tools/gen_field.py <out_dir>
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MASK = 0xFFFFFFFF


# ----------------------------------------------------------------------------- IR
class Prog:
    """Straight-line program.  Operands are register names (str) or immediates (int)."""

    def __init__(self, name, inputs, n_out):
        self.name = name
        self.inputs = list(inputs)          # input register names
        self.outputs = ["o%d" % i for i in range(n_out)]
        self.ops = []
        self.temps = []
        self._n = 0

    def tmp(self):
        t = "t%d" % self._n
        self._n += 1
        self.temps.append(t)
        return t

    def op(self, opc, d, *srcs):
        self.ops.append((opc, d) + tuple(srcs))
        return d

    # --- interpreter (PTX semantics, CC.CF modelled explicitly)
    def run(self, in_vals, strict=False):
        """Interpret with PTX semantics.  strict=True additionally raises when an addition WITHOUT a carry-out
        (.cc) overflows, i.e. when the program would silently drop a carry."""
        reg = dict(zip(self.inputs, in_vals))
        cf = 0

        def lost(t, opc):
            if strict and (t >> 32):
                raise OverflowError("carry dropped by %s in %s" % (opc, self.name))

        def v(x):
            return x & MASK if isinstance(x, int) else reg[x]

        for ins in self.ops:
            opc, d, s = ins[0], ins[1], ins[2:]
            base = opc.split(".")[0]
            cc_out = opc.endswith(".cc")
            cin = base in ("addc", "subc", "madc")
            if base in ("add", "addc"):
                t = v(s[0]) + v(s[1]) + (cf if cin else 0)
                if cc_out:
                    cf = t >> 32
                else:
                    lost(t, opc)
                reg[d] = t & MASK
            elif base in ("sub", "subc"):
                t = v(s[0]) - v(s[1]) - (cf if cin else 0)
                if cc_out:
                    cf = 1 if t < 0 else 0
                reg[d] = t & MASK
            elif base in ("mad", "madc"):
                prod = v(s[0]) * v(s[1])
                part = (prod & MASK) if ".lo" in opc else (prod >> 32)
                t = part + v(s[2]) + (cf if cin else 0)
                if cc_out:
                    cf = t >> 32
                else:
                    lost(t, opc)
                reg[d] = t & MASK
            elif base == "mul":
                prod = v(s[0]) * v(s[1])
                reg[d] = (prod & MASK) if ".lo" in opc else (prod >> 32)
            elif base == "mov":
                reg[d] = v(s[0])
            elif base == "selp":       # d = s[2] != 0 ? s[0] : s[1]   (emitted as setp + selp)
                reg[d] = v(s[0]) if v(s[2]) != 0 else v(s[1])
            elif base == "or":
                reg[d] = v(s[0]) | v(s[1])
            else:
                raise ValueError(opc)
        return [reg[o] for o in self.outputs]

    # --- PTX emission
    def emit_asm(self, indent="    "):
        """Return (asm_text_lines, n_outputs, n_inputs); placeholders: outputs first."""
        ph = {}
        k = 0
        for o in self.outputs:
            ph[o] = "%%%d" % k
            k += 1
        for i in self.inputs:
            ph[i] = "%%%d" % k
            k += 1

        def f(x):
            if isinstance(x, int):
                return "0x%08x" % (x & MASK)
            return ph.get(x, x)

        lines = ["{"]
        # temps in chunks
        for i in range(0, len(self.temps), 16):
            lines.append(".reg .u32 " + ", ".join(self.temps[i:i + 16]) + ";")
        if any(o[0].startswith("selp") for o in self.ops):
            lines.append(".reg .pred pq;")
        for ins in self.ops:
            opc, d, s = ins[0], ins[1], ins[2:]
            if opc == "selp":
                lines.append("setp.ne.u32 pq, %s, 0;" % f(s[2]))
                lines.append("selp.u32 %s, %s, %s, pq;" % (f(d), f(s[0]), f(s[1])))
            elif opc == "or":
                lines.append("or.b32 %s, %s;" % (f(d), ", ".join(f(x) for x in s)))
            else:
                lines.append("%s.u32 %s, %s;" % (opc, f(d), ", ".join(f(x) for x in s)))
        lines.append("}")
        return lines


def limbs(x, n):
    return [(x >> (32 * i)) & MASK for i in range(n)]


def from_limbs(v):
    return sum(int(x) << (32 * i) for i, x in enumerate(v))


# ----------------------------------------------------------------------------- programs
def _cond_sub_p(P, T, p_l, outs):
    """outs = T - p if T >= p else T   (T < 2p < 2^(32N))."""
    n = len(T)
    U = [P.tmp() for _ in range(n)]
    P.op("sub.cc", U[0], T[0], p_l[0])
    for k in range(1, n):
        P.op("subc.cc", U[k], T[k], p_l[k])
    bw = P.tmp()
    P.op("subc", bw, 0, 0)                      # 0xffffffff when T < p
    for k in range(n):
        P.op("selp", outs[k], T[k], U[k], bw)


def prog_add(n, p):
    A = ["a%d" % i for i in range(n)]
    B = ["b%d" % i for i in range(n)]
    P = Prog("add", A + B, n)
    p_l = limbs(p, n)
    T = [P.tmp() for _ in range(n)]
    P.op("add.cc", T[0], A[0], B[0])
    for k in range(1, n - 1):
        P.op("addc.cc", T[k], A[k], B[k])
    P.op("addc", T[n - 1], A[n - 1], B[n - 1])
    _cond_sub_p(P, T, p_l, P.outputs)
    return P


def prog_sub(n, p):
    A = ["a%d" % i for i in range(n)]
    B = ["b%d" % i for i in range(n)]
    P = Prog("sub", A + B, n)
    p_l = limbs(p, n)
    T = [P.tmp() for _ in range(n)]
    P.op("sub.cc", T[0], A[0], B[0])
    for k in range(1, n):
        P.op("subc.cc", T[k], A[k], B[k])
    bw = P.tmp()
    P.op("subc", bw, 0, 0)                      # all-ones when a < b
    U = [P.tmp() for _ in range(n)]
    P.op("add.cc", U[0], T[0], p_l[0])
    for k in range(1, n - 1):
        P.op("addc.cc", U[k], T[k], p_l[k])
    P.op("addc", U[n - 1], T[n - 1], p_l[n - 1])
    for k in range(n):
        P.op("selp", P.outputs[k], U[k], T[k], bw)
    return P


def _reduce_row(P, X, Y, p_l, m0):
    """X (pos 0..n-1) / Y (pos 1..n): add m*p so that X[0] becomes 0."""
    n = len(X)
    m = P.tmp()
    unit = m0 == MASK and p_l[0] == 1 and "zr" in P.inputs
    if unit:
        # p = 1 mod 2^32 (BLS12-377 fp / fr, BLS12-381 fr): m = -X[0] and the first column is X[0] + m = 0 with
        # carry (X[0] != 0) - both are plain integer adds on the ALU pipe instead of an IMAD and an IMAD.HI on the
        # (binding) multiplier pipe.  The zero comes from the constant bank so ptxas cannot fold the negation
        # into the other columns' multiplies (which would break their IMAD.WIDE fusion).
        P.op("sub", m, "zr", X[0])
    else:
        # m0 == 2^32-1 would let ptxas rewrite m = -X[0] and fold the negation into the modulus
        # immediates of the lo half only, which blocks IMAD.WIDE fusion - keep it opaque instead.
        P.op("mul.lo", m, X[0], "m0r" if (m0 == MASK and "m0r" in P.inputs) else m0)
    # odd limbs of p into Y
    for j in range(0, n, 2):
        P.op("mad.lo.cc" if j == 0 else "madc.lo.cc", Y[j], m, p_l[j + 1], Y[j])
        P.op("madc.hi.cc" if j + 2 < n else "madc.hi", Y[j + 1], m, p_l[j + 1], Y[j + 1])
    # even limbs of p into X
    for j in range(0, n, 2):
        if unit and j == 0:
            P.op("add.cc", X[0], X[0], m)
            P.op("addc.cc", X[1], X[1], 0)
            continue
        P.op("mad.lo.cc" if j == 0 else "madc.lo.cc", X[j], m, p_l[j], X[j])
        P.op("madc.hi.cc", X[j + 1], m, p_l[j], X[j + 1])
    P.op("addc", Y[n - 1], Y[n - 1], 0)


def _opaque_inputs(p, n):
    """extra inputs read through the constant bank: the opaque zero (unit-low-limb moduli) or the opaque M0"""
    m0 = (-pow(p, -1, 1 << 32)) & MASK
    # opt-in (B200_GEN_UNIT=1): measured on B200 the binding resource is the IMAD.WIDE count, which this
    # variant does not change (the IMAD / IMAD.HI it removes issue beside the wide MACs), so it is off by default
    if m0 == MASK and (p & MASK) == 1 and os.environ.get("B200_GEN_UNIT"):
        return ["zr"]
    return ["m0r"] if m0 == MASK else []


def prog_mul(n, p, b_is_a=False):
    """Montgomery product a*b/R mod p, even/odd column accumulation (2n^2+n wide MACs)."""
    assert n % 2 == 0
    A = ["a%d" % i for i in range(n)]
    B = A if b_is_a else ["b%d" % i for i in range(n)]
    m0 = (-pow(p, -1, 1 << 32)) & MASK
    extra = _opaque_inputs(p, n)
    P = Prog("sqr" if b_is_a else "mul", (A if b_is_a else A + B) + extra, n)
    p_l = limbs(p, n)
    X = [P.tmp() for _ in range(n)]
    Y = [P.tmp() for _ in range(n)]
    for j in range(0, n, 2):
        P.op("mul.lo", X[j], A[j], B[0])
        P.op("mul.hi", X[j + 1], A[j], B[0])
    for j in range(0, n, 2):
        P.op("mul.lo", Y[j], A[j + 1], B[0])
        P.op("mul.hi", Y[j + 1], A[j + 1], B[0])
    _reduce_row(P, X, Y, p_l, m0)
    for i in range(1, n):
        Xp, X = X, Y                               # shift by one limb: old odd column becomes even
        Y = [P.tmp() for _ in range(n)]
        P.op("add.cc", X[0], X[0], Xp[1])
        for j in range(0, n, 2):
            lo_add = Xp[j + 2] if j + 2 < n else 0
            hi_add = Xp[j + 3] if j + 3 < n else 0
            P.op("madc.lo.cc", Y[j], A[j + 1], B[i], lo_add)
            P.op("madc.hi.cc" if j + 2 < n else "madc.hi", Y[j + 1], A[j + 1], B[i], hi_add)
        for j in range(0, n, 2):
            P.op("mad.lo.cc" if j == 0 else "madc.lo.cc", X[j], A[j], B[i], X[j])
            P.op("madc.hi.cc", X[j + 1], A[j], B[i], X[j + 1])
        P.op("addc", Y[n - 1], Y[n - 1], 0)
        _reduce_row(P, X, Y, p_l, m0)
    T = [P.tmp() for _ in range(n)]
    P.op("add.cc", T[0], X[1], Y[0])
    for k in range(1, n - 1):
        P.op("addc.cc", T[k], X[k + 1], Y[k])
    P.op("addc", T[n - 1], Y[n - 1], 0)
    _cond_sub_p(P, T, p_l, P.outputs)
    return P


def _mont_rows(P, X, Y, p_l, m0, n):
    """n Montgomery reduction rows over the single-width value held in X (+ Y = 0): returns the n-limb
    list T = (X + q p) / R  (< p + 1), not yet conditionally reduced."""
    _reduce_row(P, X, Y, p_l, m0)
    for i in range(1, n):
        Xp, X = X, Y
        Y = [P.tmp() for _ in range(n)]
        P.op("add.cc", X[0], X[0], Xp[1])
        for k in range(n):
            src = Xp[k + 2] if k + 2 < n else 0
            P.op("addc.cc" if k < n - 1 else "addc", Y[k], src, 0)
        _reduce_row(P, X, Y, p_l, m0)
    T = [P.tmp() for _ in range(n)]
    P.op("add.cc", T[0], X[1], Y[0])
    for k in range(1, n - 1):
        P.op("addc.cc", T[k], X[k + 1], Y[k])
    P.op("addc", T[n - 1], Y[n - 1], 0)
    return T


def prog_sqr(n, p):
    """Montgomery square a*a/R mod p with the off-diagonal products computed once:
    n(n-1)/2 + n wide MACs for a^2 (instead of n^2) plus the n^2 + n of the reduction rows; the doubling and
    the three 2n-limb carry chains run on the ALU pipe, which has slack next to the multiplier pipe."""
    assert n % 2 == 0
    A = ["a%d" % i for i in range(n)]
    m0 = (-pow(p, -1, 1 << 32)) & MASK
    P = Prog("sqr", A + _opaque_inputs(p, n), n)
    p_l = limbs(p, n)
    # two accumulators so that the (lo, hi) pairs of one carry chain never overlap: EV holds pairs that start at
    # even limb positions, OD pairs that start at odd positions
    EV = [P.tmp() for _ in range(2 * n)]
    OD = [P.tmp() for _ in range(2 * n)]
    for k in range(2 * n):
        P.op("mov", EV[k], 0)
        P.op("mov", OD[k], 0)
    for i in range(n - 1):
        for first_j in (i + 1, i + 2):
            js = list(range(first_j, n, 2))
            if not js:
                continue
            acc = OD if (i + first_j) % 2 else EV
            for idx, j in enumerate(js):
                pos = i + j
                P.op("mad.lo.cc" if idx == 0 else "madc.lo.cc", acc[pos], A[i], A[j], acc[pos])
                P.op("madc.hi.cc", acc[pos + 1], A[i], A[j], acc[pos + 1])
            top = i + js[-1] + 1
            # absorb the chain's carry (the limbs above `top` hold at most earlier carries)
            # (top <= 2n - 2: the product a_{n-2} a_{n-1} ends at limb 2n - 2)
            if top + 2 < 2 * n:
                P.op("addc.cc", acc[top + 1], acc[top + 1], 0)
                P.op("addc", acc[top + 2], acc[top + 2], 0)
            else:
                P.op("addc", acc[top + 1], acc[top + 1], 0)
    # S = EV + OD ; T = 2 S + diag
    S = [P.tmp() for _ in range(2 * n)]
    P.op("add.cc", S[0], EV[0], OD[0])
    for k in range(1, 2 * n - 1):
        P.op("addc.cc", S[k], EV[k], OD[k])
    P.op("addc", S[2 * n - 1], EV[2 * n - 1], OD[2 * n - 1])
    S2 = [P.tmp() for _ in range(2 * n)]
    P.op("add.cc", S2[0], S[0], S[0])
    for k in range(1, 2 * n - 1):
        P.op("addc.cc", S2[k], S[k], S[k])
    P.op("addc", S2[2 * n - 1], S[2 * n - 1], S[2 * n - 1])
    D = [P.tmp() for _ in range(2 * n)]
    for i in range(n):
        P.op("mul.lo", D[2 * i], A[i], A[i])
        P.op("mul.hi", D[2 * i + 1], A[i], A[i])
    T = [P.tmp() for _ in range(2 * n)]
    P.op("add.cc", T[0], S2[0], D[0])
    for k in range(1, 2 * n - 1):
        P.op("addc.cc", T[k], S2[k], D[k])
    P.op("addc", T[2 * n - 1], S2[2 * n - 1], D[2 * n - 1])
    # Montgomery-reduce the low half, then add the high half:  (T_lo + q p) / R + T_hi  <  2p
    X = [P.tmp() for _ in range(n)]
    Y = [P.tmp() for _ in range(n)]
    for k in range(n):
        P.op("mov", X[k], T[k])
        P.op("mov", Y[k], 0)
    Tr = _mont_rows(P, X, Y, p_l, m0, n)
    U = [P.tmp() for _ in range(n)]
    P.op("add.cc", U[0], Tr[0], T[n])
    for k in range(1, n - 1):
        P.op("addc.cc", U[k], Tr[k], T[n + k])
    P.op("addc", U[n - 1], Tr[n - 1], T[2 * n - 1])
    _cond_sub_p(P, U, p_l, P.outputs)
    return P


def prog_from_mont(n, p):
    """a / R mod p  (Montgomery reduction of a single-width value; n^2+n wide MACs)."""
    A = ["a%d" % i for i in range(n)]
    m0 = (-pow(p, -1, 1 << 32)) & MASK
    P = Prog("from_mont", A + _opaque_inputs(p, n), n)
    p_l = limbs(p, n)
    X = [P.tmp() for _ in range(n)]
    Y = [P.tmp() for _ in range(n)]
    for k in range(n):
        P.op("mov", X[k], A[k])
        P.op("mov", Y[k], 0)
    _reduce_row(P, X, Y, p_l, m0)
    for i in range(1, n):
        Xp, X = X, Y
        Y = [P.tmp() for _ in range(n)]
        P.op("add.cc", X[0], X[0], Xp[1])
        for k in range(n):
            src = Xp[k + 2] if k + 2 < n else 0
            P.op("addc.cc" if k < n - 1 else "addc", Y[k], src, 0)
        _reduce_row(P, X, Y, p_l, m0)
    T = [P.tmp() for _ in range(n)]
    P.op("add.cc", T[0], X[1], Y[0])
    for k in range(1, n - 1):
        P.op("addc.cc", T[k], X[k + 1], Y[k])
    P.op("addc", T[n - 1], Y[n - 1], 0)
    _cond_sub_p(P, T, p_l, P.outputs)
    return P


# ----------------------------------------------------------------------------- fields
def curve_moduli():
    """name -> (p, r) from the product's own table (davinci-node_b200/layout.py; loaded by path: the package itself
    needs the library this generator is a build step of)."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "davinci-node_b200", "layout.py")
    spec = importlib.util.spec_from_file_location("_b200_layout", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return {v[0]: (v[1], v[2]) for v in mod.CURVES.values()}


def field_table():
    M = curve_moduli()
    return [
        ("bn254_fp", M["bn254"][0], 8),
        ("bn254_fr", M["bn254"][1], 8),
        ("bls12_377_fp", M["bls12_377"][0], 12),
        ("bls12_377_fr", M["bls12_377"][1], 8),
        ("bls12_381_fp", M["bls12_381"][0], 12),
        ("bls12_381_fr", M["bls12_381"][1], 8),
        ("bw6_761_fp", M["bw6_761"][0], 24),
        # bw6_761_fr == bls12_377_fp (same modulus): aliased in field.cuh
    ]


def programs(n, p):
    return {
        "add": prog_add(n, p),
        "sub": prog_sub(n, p),
        "mul": prog_mul(n, p),
        # the dedicated squaring (prog_sqr) saves 22% of the wide MACs of a square but its three 2n-limb carry
        # chains lengthen the dependent instruction stream: measured -4% on the bucket-accumulation kernel, which
        # runs at ~82% of the IMAD.WIDE pipe and is otherwise latency-limited.  Opt-in with B200_GEN_SQR=1.
        "sqr": prog_sqr(n, p) if os.environ.get("B200_GEN_SQR") else prog_mul(n, p, b_is_a=True),
        "from_mont": prog_from_mont(n, p),
    }


def emit_fn(prog, fname, n, arity):
    """C++ wrapper: static __device__ void fname(uint32_t* r, const uint32_t* a[, const uint32_t* b])."""
    lines = prog.emit_asm()
    args = "uint32_t* __restrict__ r, const uint32_t* a" + (", const uint32_t* b" if arity == 2 else "")
    out = ["  static __device__ __forceinline__ void %s(%s) {" % (fname, args)]
    # outputs go to fresh locals so r may alias a / b
    out.append("    uint32_t " + ", ".join("o%d" % i for i in range(n)) + ";")
    out.append("    asm(")
    for ln in lines:
        out.append('      "%s\\n\\t"' % ln)
    outs = ", ".join('"=r"(o%d)' % i for i in range(n))
    ins = ", ".join('"r"(a[%d])' % i for i in range(n))
    if arity == 2:
        ins += ", " + ", ".join('"r"(b[%d])' % i for i in range(n))
    if "m0r" in prog.inputs:
        ins += ', "r"(k_m0_opaque[0])'
    if "zr" in prog.inputs:
        ins += ', "r"(k_m0_opaque[1])'
    out.append("      : %s" % outs)
    out.append("      : %s);" % ins)
    out.append("    " + " ".join("r[%d] = o%d;" % (i, i) for i in range(n)))
    out.append("  }")
    return "\n".join(out)


def emit_field(name, p, n):
    progs = programs(n, p)
    R = 1 << (32 * n)
    hdr = []
    hdr.append("// GENERATED by tools/gen_field.py - do not edit.  Field %s, %d x 32-bit limbs." % (name, n))
    hdr.append("#pragma once")
    hdr.append("#include <stdint.h>")
    hdr.append("namespace b200 {")
    hdr.append("#ifndef B200_M0_OPAQUE_DEFINED")
    hdr.append("#define B200_M0_OPAQUE_DEFINED")
    hdr.append("// read through the constant bank so ptxas cannot see that M0 == 2^32-1 (see gen_field.py)")
    hdr.append("static __device__ __constant__ uint32_t k_m0_opaque[2] = {0xffffffffu, 0u};")
    hdr.append("#endif")
    hdr.append("struct %s {" % name)
    hdr.append("  static constexpr int N = %d;" % n)
    hdr.append("  static constexpr int BITS = %d;" % p.bit_length())

    def arr(label, val):
        return "  static __device__ __host__ __forceinline__ uint32_t %s(int i) { constexpr uint32_t v[%d] = {%s}; return v[i]; }" % (
            label, n, ", ".join("0x%08xu" % x for x in limbs(val, n)))

    hdr.append(arr("modulus", p))
    hdr.append(arr("one", R % p))             # Montgomery form of 1
    hdr.append(arr("r2", (R * R) % p))        # to-Montgomery multiplier
    hdr.append(arr("r3", (R * R * R) % p))
    hdr.append("  static constexpr uint32_t M0 = 0x%08xu;" % ((-pow(p, -1, 1 << 32)) & MASK))
    # Tonelli-Shanks constants (point decompression): p - 1 = 2^S * T, sqrt_e = (T - 1) / 2, sqrt_z = nqr^T (Montgomery)
    S_, T_ = 0, p - 1
    while T_ % 2 == 0:
        T_ //= 2
        S_ += 1
    nqr = 2
    while pow(nqr, (p - 1) // 2, p) != p - 1:
        nqr += 1
    hdr.append("  static constexpr int TWO_ADICITY = %d;" % S_)
    hdr.append(arr("sqrt_e", (T_ - 1) // 2))
    hdr.append(arr("sqrt_z", pow(nqr, T_, p) * R % p))
    hdr.append(emit_fn(progs["add"], "add", n, 2))
    hdr.append(emit_fn(progs["sub"], "sub", n, 2))
    hdr.append(emit_fn(progs["mul"], "mul", n, 2))
    hdr.append(emit_fn(progs["sqr"], "sqr", n, 1))
    hdr.append(emit_fn(progs["from_mont"], "from_mont", n, 1))
    hdr.append("};")
    hdr.append("}  // namespace b200")
    return "\n".join(hdr) + "\n"


# ----------------------------------------------------------------------------- pairing constants
def pairing_table():
    """(curve, p, r, k, m_half, m_0, u0, u1, D-twist): F_{p^k} = Fp[w] / (w^k + m_half w^(k/2) + m_0) and the image
    u = u0 + u1 w^(k/2) of the Fp2 generator (u1 = 0: G2 lives over Fp).  Towers of gnark-crypto:
      BN254      u^2 = -1, w^6 = 9 + u   BLS12-381  u^2 = -1, w^6 = 1 + u
      BLS12-377  u^2 = -5, w^6 = u       BW6-761    k = 6, w^6 = -4"""
    M = curve_moduli()
    return [
        ("bn254",) + M["bn254"] + (12, -18, 82, -9, 1, True),
        ("bls12_377",) + M["bls12_377"] + (12, 0, 5, 0, 1, True),
        ("bls12_381",) + M["bls12_381"] + (12, -2, 2, -1, 1, False),
        ("bw6_761",) + M["bw6_761"] + (6, 0, 4, 0, 0, False),
    ]


def emit_pairing():
    out = ["// GENERATED by tools/gen_field.py - do not edit.  Extension-field shapes and final exponents (p^k - 1) / r of the",
           "// reduced Tate pairing (csrc/pairing.cuh).",
           "#pragma once", "#include <stdint.h>",
           "#ifdef B200_PAIRING_HOST", "#define B200_PAIRING_CONST static const", "#define B200_PAIRING_FN static inline",
           "#else", "#define B200_PAIRING_CONST static __device__ const",
           "#define B200_PAIRING_FN static __device__ __forceinline__", "#endif",
           "namespace b200 {"]
    for name, p, r, k, mh, m0, u0, u1, dtw in pairing_table():
        assert (p ** k - 1) % r == 0
        e = (p ** k - 1) // r
        nw = (e.bit_length() + 31) // 32
        out.append("B200_PAIRING_CONST uint32_t k_final_exp_%s[%d] = {%s};" % (
            name, nw, ", ".join("0x%08xu" % x for x in limbs(e, nw))))
        out.append("struct pairing_%s {" % name)
        out.append("  static constexpr int K = %d, MH = %d, M0 = %d, U0 = %d, U1 = %d;" % (k, mh, m0, u0, u1))
        out.append("  static constexpr bool DTWIST = %s;" % ("true" if dtw else "false"))
        out.append("  static constexpr int FE_BITS = %d;" % e.bit_length())
        out.append("  B200_PAIRING_FN uint32_t final_exp(int i) { return k_final_exp_%s[i]; }" % name)
        out.append("};")
    out.append("}  // namespace b200")
    return "\n".join(out) + "\n"


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "davinci-node_b200", "csrc", "gen")
    os.makedirs(out_dir, exist_ok=True)
    for name, p, n in field_table():
        path = os.path.join(out_dir, "field_%s.cuh" % name)
        txt = emit_field(name, p, n)
        old = open(path).read() if os.path.exists(path) else None
        if old != txt:
            with open(path, "w") as fh:
                fh.write(txt)
        print("generated", path, "(%d limbs)" % n)
    path = os.path.join(out_dir, "pairing_consts.cuh")
    txt = emit_pairing()
    if (open(path).read() if os.path.exists(path) else None) != txt:
        with open(path, "w") as fh:
            fh.write(txt)
    print("generated", path)


if __name__ == "__main__":
    main()
