"""Drives oracle/c's prove schedule from oracle-level objects (used by CPU tests and as the checker
for larger GPU parity cases)."""
import numpy as np

from davinci_node_b200.layout import Layout
from oracle import cport
from oracle import ntt as N


def extended_key(L, cs, pk):
    """A_ext / B_ext / K_ext arrays and index maps, the same convention as csrc/prover.cu."""
    SKIP = 0xFFFFFFFF
    m = cs.nb_wires
    infA, infB = pk["InfinityA"], pk["InfinityB"]
    mapA = np.full(m + 4, SKIP, dtype=np.uint32)
    mapB = np.full(m + 4, SKIP, dtype=np.uint32)
    ia = ib = 0
    for i in range(m):
        if not infA[i]:
            mapA[i] = ia
            ia += 1
        if not infB[i]:
            mapB[i] = ib
            ib += 1
    mapA[m + 0], mapA[m + 2] = ia, ia + 1
    mapB[m + 1], mapB[m + 2] = ib, ib + 1
    skip = set()
    for cm in cs.commitments:
        skip.update(cm["private_committed"])
        skip.add(cm["commitment_index"])
    npriv = m - cs.nb_public
    mapK = np.full(npriv + 4, SKIP, dtype=np.uint32)
    ik = 0
    for j in range(npriv):
        if cs.nb_public + j not in skip:
            mapK[j] = ik
            ik += 1
    mapK[npriv + 3] = ik
    G1, G2 = pk["G1"], pk["G2"]
    return {
        "A": L.enc_affine(G1["A"] + [G1["Delta"], G1["Alpha"]], 1),
        "B1": L.enc_affine(G1["B"] + [G1["Delta"], G1["Beta"]], 1),
        "B2": L.enc_affine(G2["B"] + [G2["Delta"], G2["Beta"]], 2),
        "K": L.enc_affine(G1["K"] + [G1["Delta"]], 1),
        "Z": L.enc_affine(G1["Z"], 1),
        "mapA": mapA, "mapB": mapB, "mapK": mapK,
    }


def c_oracle_prove(cname, cs, pk, W, r, s, threads=0):
    from oracle import groth16 as OG
    L = Layout(cname)
    q = L.r
    ek = extended_key(L, cs, pk)
    n = pk["domain_size"]
    a, b, c = OG.constraint_values(cs, W, q)
    pad = lambda v: v + [0] * (n - len(v))
    ba, bb, bc = L.enc_fr(pad(a)), L.enc_fr(pad(b)), L.enc_fr(pad(c))
    wext = L.enc_fr(list(W) + [r, s, 1, (-r * s) % q])
    omega, g = L.enc_fr([pk["generator"]]), L.enc_fr([pk["coset_gen"]])
    out_ar = np.zeros(L.affine_bytes(1), dtype=np.uint8)
    out_krs = np.zeros(L.affine_bytes(1), dtype=np.uint8)
    out_bs = np.zeros(L.affine_bytes(2), dtype=np.uint8)
    args = cport.ProveArgs(
        curve=L.id, logn=n.bit_length() - 1, omega=cport.p(omega), g=cport.p(g),
        A_ext=cport.p(ek["A"]), B1_ext=cport.p(ek["B1"]), B2_ext=cport.p(ek["B2"]), K_ext=cport.p(ek["K"]),
        Z=cport.p(ek["Z"]), mapA=cport.p(ek["mapA"]), mapB=cport.p(ek["mapB"]), mapK=cport.p(ek["mapK"]),
        m=cs.nb_wires, nb_public=cs.nb_public, nZ=len(pk["G1"]["Z"]), W_ext=cport.p(wext),
        a=cport.p(ba), b=cport.p(bb), c=cport.p(bc), out_ar=cport.p(out_ar), out_bs=cport.p(out_bs),
        out_krs=cport.p(out_krs), threads=threads)
    assert cport.lib().oc_prove(args) == 0
    return {"Ar": L.dec_affine(out_ar, 1)[0], "Bs": L.dec_affine(out_bs, 2)[0], "Krs": L.dec_affine(out_krs, 1)[0]}
