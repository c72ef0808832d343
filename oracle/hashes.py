"""Hash functions of the BSB22 commitment challenge and the Solidity verifier.  TEST INFRASTRUCTURE ONLY.

  * keccak256 - legacy Keccak padding (Ethereum), used by `solidity.WithProverTargetSolidityVerifier`
    (/root/reference/circuits/statetransition/artifacts.go:18) and by the verifier contract
    (/root/reference/config/statetransition_vkey.sol:668-677).  Known answers: keccak256("") and keccak256("abc").
  * expand_message_xmd / fr_hash - RFC 9380 section 5.3.1 with SHA-256 and gnark-crypto's `fr.Hash(msg, dst, count)`
    (L = 16 + ceil(bits / 8) bytes per element, big-endian, reduced mod r): gnark's default commitment hash
    `hash_to_field.New([]byte("bsb22-commitment"))` and the proof-of-knowledge fold challenge (dst "G16-BSB22").
    Known answers: RFC 9380 appendix K.1 (expand_message_xmd, SHA-256).
gnark / gnark-crypto are un-vendored go.mod dependencies (go.mod:15-16); the call sites restated here are listed in
SURVEY.md A.1 steps 3 and 5.
"""
import hashlib

_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def _keccak_f(A):
    for rc in _RC:
        Cc = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [Cc[(x - 1) % 5] ^ _rol(Cc[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= rc
    return A


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        A = _keccak_f(A)
    out = b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


def expand_message_xmd(msg: bytes, dst: bytes, length: int) -> bytes:
    """RFC 9380 5.3.1 with H = SHA-256 (b = 32 bytes, block 64 bytes)."""
    b_len, block = 32, 64
    ell = -(-length // b_len)
    if ell > 255 or length > 65535 or len(dst) > 255:
        raise ValueError("expand_message_xmd: parameters out of range")
    dst_prime = dst + bytes([len(dst)])
    h = lambda x: hashlib.sha256(x).digest()
    b0 = h(bytes(block) + msg + length.to_bytes(2, "big") + b"\x00" + dst_prime)
    bi = h(b0 + b"\x01" + dst_prime)
    out = bytearray(bi)
    for i in range(2, ell + 1):
        bi = h(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime)
        out += bi
    return bytes(out[:length])


def fr_hash(msg: bytes, dst: bytes, count: int, modulus: int):
    """gnark-crypto fr.Hash: `count` field elements from expand_message_xmd, 16 + ceil(bits/8) bytes each."""
    L = 16 + (modulus.bit_length() + 7) // 8
    raw = expand_message_xmd(msg, dst, count * L)
    return [int.from_bytes(raw[i * L:(i + 1) * L], "big") % modulus for i in range(count)]


def fr_bytes_len(modulus: int) -> int:
    return (modulus.bit_length() + 7) // 8


def g1_raw_bytes(pt, coord_bytes: int) -> bytes:
    """G1Affine.Marshal(): uncompressed X || Y, big-endian (infinity: 0x40 flag then zeros)."""
    if pt is None:
        return bytes([0x40]) + bytes(2 * coord_bytes - 1)
    return pt[0].to_bytes(coord_bytes, "big") + pt[1].to_bytes(coord_bytes, "big")


def commitment_challenge(kind: str, commitment_pt, public_committed, r: int, p: int) -> int:
    """Value of a BSB22 commitment wire: hash of Commitment.Marshal() || every public committed
    value as a big-endian fr, reduced mod r.  kind: 'default' (hash_to_field "bsb22-commitment") or 'solidity' (keccak)."""
    nb = fr_bytes_len(r)
    data = g1_raw_bytes(commitment_pt, (p.bit_length() + 7) // 8) + b"".join(int(v).to_bytes(nb, "big") for v in public_committed)
    if kind == "solidity":
        return int.from_bytes(keccak256(data), "big") % r
    if kind == "default":
        return fr_hash(data, b"bsb22-commitment", 1, r)[0]
    raise ValueError("unknown commitment hash kind %r" % kind)


def fold_challenge(commitment_wire_values, r: int) -> int:
    """pedersen.BatchProve's challenge: fr.Hash(concat of the commitment wires' values, "G16-BSB22", 1)[0]."""
    nb = fr_bytes_len(r)
    return fr_hash(b"".join(int(v).to_bytes(nb, "big") for v in commitment_wire_values), b"G16-BSB22", 1, r)[0]
