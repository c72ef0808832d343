"""gnark-crypto point compression and gnark's proving-key stream, big-int restatement.  TEST INFRASTRUCTURE ONLY.

Independent of the product's reader / writer (davinci-node_b200/artifacts.py): points are compressed here from Python
integers with the flag rules of SURVEY.md A.4 (gnark-crypto ecc/<curve>/marshal.go, un-vendored go.mod dependency):
BN254 2 flag bits (10 smallest y, 11 largest, 01 infinity), the other curves 3 bits (100 / 101 / 110); G2 over Fp2 writes
X.A1 || X.A0 and orders y by (A1, A0).  The stream layout is gnark's `pk.WriteTo`
(/root/reference/cmd/circuit-compile/main.go:507-512 writes it, circuits/artifacts.go:391-406 reads it back).
Parity: the framing is as surveyed (from memory of gnark v0.14) - unpinned by reference bytes; the point encoding of
BLS12-381 G1 / G2 is pinned by the EIP-4844 SRS file and crypto/blobs/kzg.go.
"""
import struct


def _flags(curve_name):
    if curve_name == "bn254":
        return dict(smallest=0x80, largest=0xC0, infinity=0x40)
    return dict(smallest=0x80, largest=0xA0, infinity=0xC0)


def _largest(y, p):
    if isinstance(y, tuple):
        return (y[1] > (p - 1) // 2) if y[1] else (y[0] > (p - 1) // 2)
    return y > (p - 1) // 2


def compress_point(cx, group, pt) -> bytes:
    fl = _flags(cx.name)
    nb = (cx.p.bit_length() + 7) // 8
    nb += (-nb) % 8                       # whole 64-bit limbs: 32 / 48 / 96 bytes
    wide = group == 2 and cx.c.g2_degree == 2
    size = 2 * nb if wide else nb
    if pt is None:
        return bytes([fl["infinity"]]) + bytes(size - 1)
    x, y = pt
    raw = (x[1].to_bytes(nb, "big") + x[0].to_bytes(nb, "big")) if wide else x.to_bytes(nb, "big")
    first = raw[0] | (fl["largest"] if _largest(y, cx.p) else fl["smallest"])
    return bytes([first]) + raw[1:]


def write_proving_key(cx, pk) -> bytes:
    """pk: the oracle's key dict (oracle/groth16.py setup)."""
    r = cx.r
    nbr = (r.bit_length() + 7) // 8
    nbr += (-nbr) % 8
    out = bytearray()
    card = pk["domain_size"]
    gen, coset = pk["generator"], pk["coset_gen"]
    out += struct.pack(">Q", card)
    for v in (pow(card, -1, r), gen, pow(gen, -1, r), coset, pow(coset, -1, r)):
        out += v.to_bytes(nbr, "big")
    out += b"\x00"
    pts = lambda group, lst: struct.pack(">I", len(lst)) + b"".join(compress_point(cx, group, q) for q in lst)
    G1, G2 = pk["G1"], pk["G2"]
    for q in (G1["Alpha"], G1["Beta"], G1["Delta"]):
        out += compress_point(cx, 1, q)
    for lst in (G1["A"], G1["B"], G1["Z"], G1["K"]):
        out += pts(1, lst)
    out += compress_point(cx, 2, G2["Beta"]) + compress_point(cx, 2, G2["Delta"]) + pts(2, G2["B"])
    infa, infb = pk["InfinityA"], pk["InfinityB"]
    out += struct.pack(">QQQ", len(infa), sum(map(bool, infa)), sum(map(bool, infb)))
    out += struct.pack(">I", len(infa)) + bytes(1 if v else 0 for v in infa)
    out += struct.pack(">I", len(infb)) + bytes(1 if v else 0 for v in infb)
    out += struct.pack(">I", len(pk["CommitmentKeys"]))
    for key in pk["CommitmentKeys"]:
        out += pts(1, key["Basis"]) + pts(1, key["BasisExpSigma"])
    return bytes(out)
