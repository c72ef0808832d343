"""Curve / field parameter table for the Groth16 hot path.

TEST INFRASTRUCTURE ONLY (oracle).  Nothing under ``oracle/`` is imported by the
product path (``davinci-node_b200/``); only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it.

Sources (the arithmetic itself lives in un-vendored go.mod dependencies of the
reference: gnark-crypto v0.19.3-0.20260126145145-b5cf053fbc34, go.mod:16):
  * BN254 p, r:           /root/reference/config/statetransition_vkey.sol:41-42
  * BLS12-381 fr root:    /root/reference/crypto/blobs/barycentric.go:52
  * BLS12-381 G1/G2 gens: /root/reference/crypto/blobs/kzg.go:26-45 (compressed)
  * remaining constants:  SURVEY.md Appendix A.2 / B (published curve parameters,
    re-verified numerically in tests/test_oracle_params.py)
"""
from dataclasses import dataclass, field as _dc_field


@dataclass(frozen=True)
class Curve:
    name: str
    cid: int                  # enum value used by include/b200_groth16.h
    p: int                    # base field modulus
    r: int                    # scalar field modulus (subgroup order)
    b: int                    # G1: y^2 = x^3 + b
    fp_limbs64: int
    fr_limbs64: int
    two_adicity: int
    root_of_unity: int        # primitive 2^two_adicity-th root in fr
    mult_gen: int             # gnark FrMultiplicativeGen (coset shift)
    g2_degree: int            # 2 -> G2 over Fp2, 1 -> G2 over Fp (BW6-761)
    nonresidue: int           # Fp2 = Fp[u]/(u^2 - nonresidue)     (ignored when g2_degree == 1)
    b2: tuple                 # twist coefficient (c0, c1) or (c0,) for degree 1
    g1: tuple                 # subgroup generator (x, y)
    g2: tuple = None          # filled lazily by oracle.curve (derived by cofactor clearing when not published here)

    @property
    def fp_bits(self):
        return self.p.bit_length()

    @property
    def fr_bits(self):
        return self.r.bit_length()

    @property
    def fp_limbs32(self):
        return 2 * self.fp_limbs64

    @property
    def fr_limbs32(self):
        return 2 * self.fr_limbs64


BN254 = Curve(
    name="bn254", cid=1,
    p=0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
    r=0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
    b=3, fp_limbs64=4, fr_limbs64=4, two_adicity=28,
    root_of_unity=19103219067921713944291392827692070036145651957329286315305642004821462161904,
    mult_gen=5, g2_degree=2, nonresidue=-1,
    b2=None,  # 3/(9+u), computed in oracle.curve
    g1=(1, 2),
)

BLS12_377 = Curve(
    name="bls12_377", cid=2,
    p=0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
    r=0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
    b=1, fp_limbs64=6, fr_limbs64=4, two_adicity=47,
    root_of_unity=8065159656716812877374967518403273466521432693661810619979959746626482506078,
    mult_gen=22, g2_degree=2, nonresidue=-5,
    b2=None,  # D-twist: 1/u, computed in oracle.curve
    g1=(0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef,
        0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6),
)

BLS12_381 = Curve(
    name="bls12_381", cid=3,
    p=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    b=4, fp_limbs64=6, fr_limbs64=4, two_adicity=32,
    root_of_unity=10238227357739495823651030575849232062558860180284477541189508159991286009131,
    mult_gen=7, g2_degree=2, nonresidue=-1,
    b2=(4, 4),  # M-twist: 4(1+u)
    g1=(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
)

BW6_761 = Curve(
    name="bw6_761", cid=4,
    p=0x122e824fb83ce0ad187c94004faff3eb926186a81d14688528275ef8087be41707ba638e584e91903cebaff25b423048689c8ed12f9fd9071dcd3dc73ebff2e98a116c25667a8f8160cf8aeeaf0a437e6913e6870000082f49d00000000008b,
    r=BLS12_377.p,
    b=-1, fp_limbs64=12, fr_limbs64=6, two_adicity=46,
    root_of_unity=32863578547254505029601261939868325669770508939375122462904745766352256812585773382134936404344547323199885654433,
    mult_gen=15, g2_degree=1, nonresidue=0,
    b2=(4,),  # M-twist over Fp: y^2 = x^3 + 4
    g1=None,  # derived by cofactor clearing in oracle.curve
)

CURVES = {c.name: c for c in (BN254, BLS12_377, BLS12_381, BW6_761)}
BY_ID = {c.cid: c for c in CURVES.values()}


def mont_R(limbs64: int) -> int:
    """gnark-crypto Montgomery radix: R = 2^(64*limbs)."""
    return 1 << (64 * limbs64)
