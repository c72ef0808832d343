"""Curve constants the host side needs (subgroup generators, 2-adic roots, gnark coset generators).
Static data; cross-checked against the oracle in tests/test_host_cpu.py.  Sources: SURVEY.md A.2 / B;
G2 and BW6-761 generators are the deterministic cofactor-cleared points also derived by oracle/curve.py."""

CONSTS = {
    1: {
        "name": "bn254", "two_adicity": 28,
        "root_of_unity": 0x2a3c09f0a58a7e8500e0a7eb8ef62abc402d111e41112ed49bd61b6e725b19f0,
        "mult_gen": 5,
        "g1": (0x1, 0x2),
        "g2": ((0x27d409ede13256511fb71acc9b73965ec3ee0cf9768aa74bfdaa33a3d1af123c, 0xcc52155d015f5bfe14a977f613d2d1fc2dd71966abf025a3dea3444afb7eeed), (0x1759a4021cdfb053b8c69b6252938195aee9bf0671a3eae93a2d48014484113f, 0x2416348afc3f5bd4601cfad808b770bb394d994742255c15d7567e9e3ed37b68)),
    },
    2: {
        "name": "bls12_377", "two_adicity": 47,
        "root_of_unity": 0x11d4b7f60cb92cc160c69477d1a8a12f9b506ee363e3f04a476ef4a4ec2a895e,
        "mult_gen": 22,
        "g1": (0x8848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef, 0x1914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6),
        "g2": ((0x6f72205595a839df693176b247c2fa251f7e02a29061e50540dc9e1c2bf1957bf1bab2288c257c2cb36b58f2418bc9, 0x138c24b2b4e17888beed0a9802aac837cdea39890effe00072f754ecb0152dd6cb524f281298966dbaeca23d3e462b8), (0x16235fdea6c3faf2a83d3730f6ab2c033ef6c2739002946f7dc48e4688bca1af1c9b417d58220817e0dc644b5e7d916, 0x707ac6cc7d192827fc54eb83267f3bed8511bd3c74f63a1ea75eabb66476769c8786f2af2a75166f33142379b4963c)),
    },
    3: {
        "name": "bls12_381", "two_adicity": 32,
        "root_of_unity": 0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b,
        "mult_gen": 7,
        "g1": (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb, 0x8b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
        "g2": ((0x4d1cc4ad56b68cdb595adb46cad2cc82e3d0da9a75ef283b6bbd91df14533e1a45128ec26f8ab25072da969d7628b70, 0x13a471d5149813b306fe76921cff7bb8d5c03fdc24a613f3e7a7fb8deb8097699751485a0bd2ad391718aaa4419ce75b), (0xa3d002cac5c50eb9e97e8b62ca30ffc5bf5aaacec121cdb63e19a5e358c4804439edb98366c02fd2840c7b9004f8b99, 0x1834907430540701fa8aa597f79e63960ec77037a7d9a06606c4c58bd8019969edabb81b77fae18489a80d47bab79d25)),
    },
    4: {
        "name": "bw6_761", "two_adicity": 46,
        "root_of_unity": 0x36a92e05198a8030f152488aeffc9b40fbe05b4512a3d4b44d994a0ddff8c606df0a4306fe0bc37eca603cc563b9a1,
        "mult_gen": 15,
        "g1": (0xd82cbf66753123ed25942ffadbec116b901330673728468b1653febae12aa13a5d68dc240a36cfbe185365abc6cb0cc5042c14be9179f0c6c05fc952c93a806d5316c2b601db66bd557011eb2c7dd0c1891418e3ce0e512da946c2ca98c56f, 0xa62fd67fdd91e327a96c02bc80385547a171b11241a2653b54d7359cd7569806b159fd05975390f644cd4d4d121918f1f84be0e364c557f196bd4095e732d987ca22009ba7577b80aaa35b641488679ed9ef0d43b32e776ad507137f20a2dd),
        "g2": (0xb57e4c181f2d61f9f68074b8b339da2da5cb0f398dad1a696575790f81a64889e99e92b694535070923045a2bd226be5a65f563e88e9f685b5f9b1d81e5d0cd3dcf42709ae8d9248fa04fc72b6a0ffca5c80d003fcfa9292828ee95ecacbb5, 0xe38788b22985f8434ad682fa4186c1a22045e5f189caad93979c088409d9a236123604483af21173517a02e6b7788d54818eeb547af836e7ebbcb997d7f33dfdeebacf614a4d2e37ebd1481bf92fc0fc870e8edd2e2758f59922008b96f3f5),
    },
}


def domain_constants(curve_id, logn):
    """(omega, coset generator) of gnark's fft.Domain of size 2^logn (canonical ints)."""
    from .layout import CURVES
    c = CONSTS[curve_id]
    r = CURVES[curve_id][2]
    if logn > c["two_adicity"]:
        raise ValueError("domain larger than the field's 2-adicity")
    return pow(c["root_of_unity"], 1 << (c["two_adicity"] - logn), r), c["mult_gen"]
