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
    extra = [0xFFFFFFFF] if "m0r" in pr["mul"].inputs else []
    for a, b in cases:
        A, B = g.limbs(a, n), g.limbs(b, n)
        assert g.from_limbs(pr["add"].run(A + B)) == (a + b) % p
        assert g.from_limbs(pr["sub"].run(A + B)) == (a - b) % p
        assert g.from_limbs(pr["mul"].run(A + B + extra)) == (a * b * Ri) % p
        assert g.from_limbs(pr["sqr"].run(A + extra)) == (a * a * Ri) % p
        assert g.from_limbs(pr["from_mont"].run(A + extra)) == (a * Ri) % p


@pytest.mark.parametrize("name,p,n", FIELDS, ids=[f[0] for f in FIELDS])
def test_mul_uses_2n2_plus_n_wide_macs(name, p, n):
    """P_mul(N) = 2N^2 + N 32x32->64 MACs; lo/hi halves fuse into IMAD.WIDE."""
    pr = g.programs(n, p)["mul"]
    halves = sum(1 for o in pr.ops if o[0].startswith(("mad", "mul")))
    assert halves == 2 * (2 * n * n) + n
