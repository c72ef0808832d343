"""GPU parity at scale: prove against the structured synthetic key (every key point is [k]G with a
known k) and check every proof element bit-for-bit against its closed form in the exponent
(SURVEY.md section 7 step 0).  The quotient h comes from the C++ CPU restatement (oracle/c).  Both the
host-buffer entry point (b200_prove) and the device-resident one (b200_prove_dev) are exercised."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cport
from oracle import curve as OC

pytestmark = pytest.mark.gpu


def _ints(canon):
    """(n, limbs) uint64 canonical -> list of Python ints."""
    out = canon[:, 0].astype(object)
    for j in range(1, canon.shape[1]):
        out = out + (canon[:, j].astype(object) << (64 * j))
    return out


def _dot(vals, ks, q):
    return int(np.dot(vals, ks.astype(object)) % q)


def expected_exponents(wl, sol, r, s):
    L = wl.L
    q = L.r
    W = _ints(sol["W_canon"])
    # quotient on the CPU restatement
    n = wl.n
    pad = lambda c: np.concatenate([c, np.zeros((n - c.shape[0], c.shape[1]), dtype=np.uint64)])
    a, b = pad(sol["a_canon"]), pad(sol["b_canon"])
    from davinci_node_b200.curve_consts import domain_constants
    omega, g = domain_constants(L.id, wl.logn)
    am = L.enc_fr(_ints(a).tolist())
    bm = L.enc_fr(_ints(b).tolist())
    cm = L.enc_fr([(x * y) % q for x, y in zip(_ints(a).tolist(), _ints(b).tolist())])
    assert cport.lib().oc_compute_h(L.id, cport.p(am), cport.p(bm), cport.p(cm), wl.logn, cport.p(L.enc_fr([omega])),
                                    cport.p(L.enc_fr([g])), 0) == 0
    h = np.array(L.dec_fr(am), dtype=object)
    A = (wl.alpha + _dot(W[~wl.infA], wl.kA, q) + r * wl.delta) % q
    B = (wl.beta + _dot(W[~wl.infB], wl.kB, q) + s * wl.delta) % q
    keep = np.ones(wl.m, dtype=bool)
    keep[:wl.nb_public] = False
    keep[wl.krs_skip] = False
    K = (_dot(W[keep], wl.kK, q) - r * s * wl.delta) % q
    Z = _dot(h[:wl.n - 1], wl.kZ, q)
    com = _dot(W[wl.committed], wl.kBasis, q)
    return {"Ar": A, "Bs": B, "Krs": (K + Z + s * A + r * B) % q, "Commitment": com, "Pok": com * wl.sigma % q}


@pytest.mark.parametrize("cname,logn,mix", [("bls12_377", 14, "witness"), ("bls12_377", 17, "witness"),
                                            ("bls12_377", 15, "ones"), ("bn254", 13, "zeros"),
                                            ("bn254", 15, "uniform"), ("bw6_761", 12, "witness"),
                                            ("bls12_377", 22, "witness")])     # BASELINE.json's full size (~40 s)
def test_structured_key_closed_form(cname, logn, mix):
    import torch
    from davinci_node_b200 import capi, prover, synthetic
    capi.init()
    if logn >= 20 and os.environ.get("B200_SKIP_FULLSIZE"):
        pytest.skip("full-size case disabled by B200_SKIP_FULLSIZE")
    cx = OC.ctx(cname)
    wl = synthetic.SyntheticWorkload(cname, logn, seed=logn)
    h = wl.register()
    try:
        sol = wl.solution(seed=7, mix=mix)
        r, s = 0x1234567890ABCDEF1234567890ABCDEF % cx.r, 0xFEDCBA0987654321FEDCBA0987654321 % cx.r
        want = expected_exponents(wl, sol, r, s)
        G1, G2 = cx.G1, cx.G2
        for on_device in (False, True):
            pin, pout, out, keep = wl.prove_args(sol, r, s, on_device=on_device)
            fn = capi.lib.b200_prove_dev if on_device else capi.lib.b200_prove
            capi.check(fn(h, C.byref(pin), C.byref(pout), 0))
            torch.cuda.synchronize()
            got = wl.decode_proof(out)
            assert got["Ar"] == G1.mul(cx.g1, want["Ar"]), (cname, on_device)
            assert got["Bs"] == G2.mul(cx.g2, want["Bs"])
            assert got["Krs"] == G1.mul(cx.g1, want["Krs"])
            assert got["CommitmentPok"] == G1.mul(cx.g1, want["Pok"])
        # the commitment hint entry point
        L = wl.L
        vals = sol["W"].numpy()[int(wl.committed[0]) * L.fr_bytes:(int(wl.committed[0]) + wl.n_c) * L.fr_bytes]
        outc = np.zeros(L.affine_bytes(1), dtype=np.uint8)
        capi.check(capi.lib.b200_commit(h, 0, capi.Slice(vals.ctypes.data, wl.n_c), outc.ctypes.data, 0))
        assert L.dec_affine(outc, 1)[0] == G1.mul(cx.g1, want["Commitment"])
    finally:
        prover.release_proving_key(wl.pk)
