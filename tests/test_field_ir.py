"""The generated Montgomery field programs (tools/gen_field.py), interpreted instruction by
instruction with PTX carry-flag semantics, must equal big-integer arithmetic.  This pins the exact
instruction sequences the CUDA kernels execute without needing a GPU."""
import random

import pytest

import gen_field as g

FIELDS = g.field_table()


@pytest.mark.parametrize("name,p,n", FIELDS, ids=[f[0] for f in FIELDS])
def test_field_programs_match_bigint(name, p, n):
    rnd = random.Random(0xF1E1D + n)
    pr = g.programs(n, p)
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    edge = [0, 1, 2, p - 1, p - 2, R % p, (R * R) % p, (p - 1) // 2, (1 << (p.bit_length() - 1))]
    cases = [(a, b) for a in edge for b in edge]
    cases += [(rnd.randrange(p), rnd.randrange(p)) for _ in range(150)]
    extra = [0xFFFFFFFF] if "m0r" in pr["mul"].inputs else ([0] if "zr" in pr["mul"].inputs else [])
    for a, b in cases:
        A, B = g.limbs(a, n), g.limbs(b, n)
        assert g.from_limbs(pr["add"].run(A + B)) == (a + b) % p
        assert g.from_limbs(pr["sub"].run(A + B)) == (a - b) % p
        assert g.from_limbs(pr["mul"].run(A + B + extra)) == (a * b * Ri) % p
        assert g.from_limbs(pr["sqr"].run(A + extra, strict=True)) == (a * a * Ri) % p
        assert g.from_limbs(pr["from_mont"].run(A + extra)) == (a * Ri) % p


@pytest.mark.parametrize("name,p,n", FIELDS, ids=[f[0] for f in FIELDS])
def test_mul_uses_2n2_plus_n_wide_macs(name, p, n):
    """P_mul(N) = 2N^2 + N 32x32->64 MACs; lo/hi halves fuse into IMAD.WIDE.  Moduli with a
    unit low limb (p = 1 mod 2^32) need no multiplier for m = -X[0] nor for the first column of every
    reduction row: 2N multiplier-pipe instructions fewer."""
    pr = g.programs(n, p)["mul"]
    halves = sum(1 for o in pr.ops if o[0].startswith(("mad", "mul")))
    if "zr" in pr.inputs:
        assert halves == 2 * (2 * n * n) + n - 3 * n
    else:
        assert halves == 2 * (2 * n * n) + n


@pytest.mark.parametrize("name,p,n", FIELDS, ids=[f[0] for f in FIELDS])
def test_sqr_dedicated_program(name, p, n):
    """The squaring program computes the off-diagonal products once (n(n-1)/2 + n wide MACs for a^2) and must
    never drop a carry, also for operands made of all-ones limbs (strict interpreter mode)."""
    pr = g.prog_sqr(n, p)
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    extra = [0xFFFFFFFF] if "m0r" in pr.inputs else ([0] if "zr" in pr.inputs else [])
    rnd = random.Random(0x5C0 + n)
    M = 0xFFFFFFFF
    cases = [0, 1, p - 1, (p - 1) // 2]
    top = p >> (32 * (n - 1))
    for _ in range(60):
        v = 0
        for i in range(n - 1):
            v |= (M if rnd.random() < 0.7 else rnd.randrange(1 << 32)) << (32 * i)
        v |= rnd.randrange(top) << (32 * (n - 1))
        cases.append(v % p)
    for a in cases:
        assert g.from_limbs(pr.run(g.limbs(a, n) + extra, strict=True)) == (a * a * Ri) % p
    halves = sum(1 for o in pr.ops if o[0].startswith(("mad", "mul")))
    unit = 3 * n if "zr" in pr.inputs else 0
    assert halves == 2 * (n * (n - 1) // 2 + n + n * n) + n - unit


@pytest.mark.parametrize("name,p,n", [f for f in FIELDS if f[1] % (1 << 32) == 1], ids=lambda f: str(f)[:12])
def test_unit_low_limb_variant(name, p, n, monkeypatch):
    """Opt-in variant for p = 1 mod 2^32 (B200_GEN_UNIT=1): m = -X[0] and the first reduction column on the ALU."""
    monkeypatch.setenv("B200_GEN_UNIT", "1")
    pr = g.prog_mul(n, p)
    assert "zr" in pr.inputs
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    rnd = random.Random(77 + n)
    for _ in range(100):
        a, b = rnd.randrange(p), rnd.randrange(p)
        assert g.from_limbs(pr.run(g.limbs(a, n) + g.limbs(b, n) + [0], strict=True)) == (a * b * Ri) % p
    halves = sum(1 for o in pr.ops if o[0].startswith(("mad", "mul")))
    assert halves == 2 * (2 * n * n) + n - 3 * n
