"""csrc/pairing.cuh - the template the GPU pairing kernels instantiate - compiled for the CPU (tests/host_pairing.cu)
and compared bit for bit with the big-integer pairing oracle: extension-field products, Miller values, reduced
pairings, product checks, the infinity / non-subgroup edge cases.  What is NOT covered here is the device field
arithmetic underneath (tests/test_gpu_field_ec.py) and the kernel plumbing (tests/test_gpu_zz_verify.py)."""
import random

import pytest

import host_pairing_util as HP
from oracle import curve as OC
from oracle import pairing

CURVES = ["bn254", "bls12_377", "bls12_381", "bw6_761"]


@pytest.mark.parametrize("name", CURVES)
def test_ext_mul_matches_oracle(name):
    pr = pairing.get(name)
    rnd = random.Random(7)
    p, k = pr.cx.p, pr.k
    for trial in range(4):
        a = tuple(rnd.randrange(p) for _ in range(k))
        b = tuple(rnd.randrange(p) if (trial < 2 or i in (0, 2, 3)) else 0 for i in range(k))   # dense and sparse
        assert HP.ext_mul(name, a, b) == pr.F.mul(a, b)
    assert HP.ext_mul(name, pr.F.one, a) == a
    top = tuple([0] * (k - 1) + [p - 1])
    assert HP.ext_mul(name, top, top) == pr.F.mul(top, top)


@pytest.mark.parametrize("name", CURVES)
def test_pairing_bit_exact_and_bilinear(name):
    pr = pairing.get(name)
    cx = pr.cx
    a, b = 0x1234567, 0xfedcba9
    P, Q = cx.G1.mul(cx.g1, a), cx.G2.mul(cx.g2, b)
    f_host, in_sub = HP.pair(name, P, Q, which=0)
    assert in_sub and f_host == pr.miller(P, Q)
    e_host, _ = HP.pair(name, P, Q)
    assert e_host == pr.pair(P, Q)
    e1, _ = HP.pair(name, cx.g1, cx.g2)
    assert e1 == pr.pair(cx.g1, cx.g2) and e_host == pr.F.pow(e1, a * b % cx.r) and e1 != pr.F.one


@pytest.mark.parametrize("name", CURVES)
def test_product_check_and_edge_cases(name):
    pr = pairing.get(name)
    cx = pr.cx
    a = 0x9abcdef123
    P, Q = cx.g1, cx.g2
    good = [(cx.G1.mul(P, a), Q), (cx.G1.neg(P), cx.G2.mul(Q, a))]
    bad = [(cx.G1.mul(P, a), Q), (cx.G1.neg(P), cx.G2.mul(Q, a + 1))]
    assert HP.check(name, good) == 1 and pr.product_is_one(good)
    assert HP.check(name, bad) == 0 and not pr.product_is_one(bad)
    # the point at infinity on either side contributes 1
    assert HP.check(name, good + [(None, Q), (P, None)]) == 1
    one, _ = HP.pair(name, None, Q)
    assert one == pr.F.one
    # a curve point outside the order-r subgroup is reported (the oracle asserts on it); BN254's G1 has cofactor 1
    if name != "bn254":
        rnd = random.Random(3)
        b1 = cx.G1.b
        while True:
            x = rnd.randrange(cx.p)
            y = OC.sqrt_mod((x * x * x + b1) % cx.p, cx.p)
            if y is not None and cx.G1.on_curve((x, y)) and cx.G1.mul((x, y), cx.r) is not None:
                break
        assert HP.check(name, [((x, y), Q)]) == -1


def test_fermat_inversion_variant_is_the_same_function():
    """PairingT<..., FERMAT = true> (what the batch kernels instantiate: a^(p-2) instead of the binary GCD) gives the
    same Miller values and decisions."""
    pr = pairing.get("bn254")
    cx = pr.cx
    P, Q = cx.G1.mul(cx.g1, 77), cx.G2.mul(cx.g2, 1234567)
    f, in_sub = HP.pair("bn254_fermat", P, Q, which=0)
    assert in_sub and f == pr.miller(P, Q)
    a = 991
    assert HP.check("bn254_fermat", [(cx.G1.mul(cx.g1, a), cx.g2), (cx.G1.neg(cx.g1), cx.G2.mul(cx.g2, a))]) == 1
    assert HP.check("bn254_fermat", [(cx.G1.mul(cx.g1, a), cx.g2), (cx.G1.neg(cx.g1), cx.G2.mul(cx.g2, a + 1))]) == 0
