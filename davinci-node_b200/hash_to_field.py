"""Host-side hashes of the BSB22 commitment challenge (the part of groth16.Prove that stays on the CPU in the reference
too: gnark's `opt.HashToFieldFn`; in the Go shim this is gnark's own code, untouched).

  default   gnark-crypto `hash_to_field.New([]byte("bsb22-commitment"))`: RFC 9380 expand_message_xmd (SHA-256),
            16 + ceil(bits/8) bytes per element, reduced mod r
  solidity  `solidity.WithProverTargetSolidityVerifier(backend.GROTH16)` -> legacy Keccak-256, reduced mod r
            (/root/reference/circuits/statetransition/artifacts.go:18; contract side:
            /root/reference/config/statetransition_vkey.sol:668-677)
Also the proof-of-knowledge fold challenge `fr.Hash(.., "G16-BSB22", 1)` used with two or more commitments.
"""
import hashlib

_MASK = (1 << 64) - 1
_ROUND_CONSTANTS = []
_ROTATIONS = [0] * 25
_PI = [0] * 25


def _init_tables():
    # round constants from the degree-8 LFSR, rotation offsets and the pi permutation from the (x, y) walk of the spec
    r = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            r = ((r << 1) ^ ((r >> 7) * 0x71)) & 0xFF
            if r & 2:
                rc ^= 1 << ((1 << j) - 1)
        _ROUND_CONSTANTS.append(rc)
    x, y = 1, 0
    for t in range(24):
        _ROTATIONS[x + 5 * y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5
    for x in range(5):
        for y in range(5):
            _PI[y + 5 * ((2 * x + 3 * y) % 5)] = x + 5 * y


_init_tables()


def _permute(st):
    for rc in _ROUND_CONSTANTS:
        col = [st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20] for x in range(5)]
        for x in range(5):
            c1 = col[(x + 1) % 5]
            d = col[(x + 4) % 5] ^ (((c1 << 1) | (c1 >> 63)) & _MASK)
            for y in range(0, 25, 5):
                st[x + y] ^= d
        moved = [0] * 25
        for dst in range(25):
            v, n = st[_PI[dst]], _ROTATIONS[_PI[dst]]
            moved[dst] = ((v << n) | (v >> (64 - n))) & _MASK if n else v
        for y in range(0, 25, 5):
            row = moved[y:y + 5]
            for x in range(5):
                st[x + y] = row[x] ^ (~row[(x + 1) % 5] & _MASK & row[(x + 2) % 5])
        st[0] ^= rc


def keccak256(data: bytes) -> bytes:
    """Legacy Keccak-256 (multi-rate padding 0x01 .. 0x80), as sha3.NewLegacyKeccak256 / Solidity keccak256."""
    rate = 136
    padded = bytearray(data) + b"\x01" + bytes((-len(data) - 1) % rate)
    padded[-1] |= 0x80
    st = [0] * 25
    for off in range(0, len(padded), rate):
        block = padded[off:off + rate]
        for i in range(rate // 8):
            st[i] ^= int.from_bytes(block[8 * i:8 * i + 8], "little")
        _permute(st)
    return b"".join(v.to_bytes(8, "little") for v in st[:4])


def expand_message_xmd(msg: bytes, dst: bytes, n_bytes: int) -> bytes:
    if len(dst) > 255 or n_bytes > 255 * 32:
        raise ValueError("expand_message_xmd: out of range")
    tag = dst + len(dst).to_bytes(1, "big")
    b0 = hashlib.sha256(bytes(64) + msg + n_bytes.to_bytes(2, "big") + b"\x00" + tag).digest()
    blocks = [hashlib.sha256(b0 + b"\x01" + tag).digest()]
    while 32 * len(blocks) < n_bytes:
        mixed = bytes(p ^ q for p, q in zip(b0, blocks[-1]))
        blocks.append(hashlib.sha256(mixed + (len(blocks) + 1).to_bytes(1, "big") + tag).digest())
    return b"".join(blocks)[:n_bytes]


def hash_to_fr(msg: bytes, dst: bytes, count: int, r: int):
    """gnark-crypto fr.Hash."""
    width = 16 + (r.bit_length() + 7) // 8
    stream = expand_message_xmd(msg, dst, count * width)
    return [int.from_bytes(stream[k * width:(k + 1) * width], "big") % r for k in range(count)]


KINDS = ("default", "solidity")


def commitment_challenge(kind, commitment_xy, public_committed, r, fp_bytes):
    """Value of a commitment wire: hash of Commitment.Marshal() (uncompressed X || Y, big-endian) followed by every
    public committed value as a big-endian fr element (constraint.SerializeCommitment), reduced mod r."""
    nb = (r.bit_length() + 7) // 8
    if commitment_xy is None:
        head = bytes([0x40]) + bytes(2 * fp_bytes - 1)
    else:
        head = commitment_xy[0].to_bytes(fp_bytes, "big") + commitment_xy[1].to_bytes(fp_bytes, "big")
    data = head + b"".join(int(v).to_bytes(nb, "big") for v in public_committed)
    if kind == "solidity":
        return int.from_bytes(keccak256(data), "big") % r
    if kind == "default":
        return hash_to_fr(data, b"bsb22-commitment", 1, r)[0]
    raise ValueError("unknown hash-to-field kind %r (have %s)" % (kind, ", ".join(KINDS)))


def fold_challenge(commitment_wire_values, r):
    """pedersen.BatchProve: fr.Hash(commitment wire values as big-endian fr, "G16-BSB22", 1)[0]."""
    nb = (r.bit_length() + 7) // 8
    return hash_to_fr(b"".join(int(v).to_bytes(nb, "big") for v in commitment_wire_values), b"G16-BSB22", 1, r)[0]
