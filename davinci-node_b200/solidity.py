"""On-chain proof format of the statetransition / results proofs: mirror of /root/reference/solidity/solidity.go
(`Groth16CommitmentProof.FromGnarkProof` :29-70, `ABIEncode` :85-116) over the BN254 proof the backend returns.

ABI layout `(uint256[8], uint256[2], uint256[2])`: twelve 32-byte big-endian words, all static - A.x, A.y, B.x1, B.x0,
B.y1, B.y0 (EIP-197: imaginary part first), C.x, C.y, commitment x, y, proof-of-knowledge x, y - exactly the calldata
`verifyProof` of /root/reference/config/statetransition_vkey.sol:653-658 takes."""
import json

from .gnark_types import Proof

BN254 = 1


class SolidityProofError(ValueError):
    pass


class Groth16CommitmentProof:
    def __init__(self):
        self.Ar = self.Bs = self.Krs = None
        self.Commitments = self.CommitmentPok = None

    def FromGnarkProof(self, proof: Proof):
        """solidity.go:29-70: expects a BN254 proof with (at least) one commitment."""
        if proof.curve_id != BN254:
            raise SolidityProofError("expected groth16_bn254.Proof, got curve id %r" % (proof.curve_id,))
        pts = proof.points()
        if not pts["Commitments"]:
            raise SolidityProofError("proof has no commitment")
        zero = (0, 0)
        ar, bs, krs = pts["Ar"] or zero, pts["Bs"] or ((0, 0), (0, 0)), pts["Krs"] or zero
        self.Ar = [ar[0], ar[1]]
        self.Bs = [[bs[0][1], bs[0][0]], [bs[1][1], bs[1][0]]]
        self.Krs = [krs[0], krs[1]]
        self.Commitments = list(pts["Commitments"][0] or zero)
        self.CommitmentPok = list(pts["CommitmentPok"] or zero)
        return self

    def words(self):
        return [self.Ar[0], self.Ar[1], self.Bs[0][0], self.Bs[0][1], self.Bs[1][0], self.Bs[1][1], self.Krs[0], self.Krs[1],
                self.Commitments[0], self.Commitments[1], self.CommitmentPok[0], self.CommitmentPok[1]]

    def ABIEncode(self) -> bytes:
        """solidity.go:85-116."""
        return b"".join(int(w).to_bytes(32, "big") for w in self.words())

    @staticmethod
    def ABIDecode(data: bytes):
        if len(data) != 12 * 32:
            raise SolidityProofError("expected 384 bytes")
        w = [int.from_bytes(data[32 * i:32 * (i + 1)], "big") for i in range(12)]
        p = Groth16CommitmentProof()
        p.Ar, p.Bs, p.Krs = w[0:2], [w[2:4], w[4:6]], w[6:8]
        p.Commitments, p.CommitmentPok = w[8:10], w[10:12]
        return p

    def String(self) -> str:
        """solidity.go:72-83 (JSON with the reference's field names; big integers as decimal numbers)."""
        return json.dumps({"proof": {"Ar": self.Ar, "Bs": self.Bs, "Krs": self.Krs}, "commitments": self.Commitments,
                           "commitment_pok": self.CommitmentPok})
