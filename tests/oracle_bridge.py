"""Converts oracle objects (Python ints / point tuples) into the product's gnark-layout containers."""
import numpy as np

from davinci_node_b200 import gnark_types as T
from davinci_node_b200.layout import Layout


def ccs_from_oracle(cs, curve_id):
    first_internal = min([cm["commitment_index"] for cm in cs.commitments] + [cs.O[0][0][0]])
    return T.ConstraintSystem(curve_id=curve_id, nb_wires=cs.nb_wires, nb_public=cs.nb_public,
                              nb_secret=first_internal - cs.nb_public, L=cs.L, R=cs.R, O=cs.O,
                              commitments=cs.commitments)


def pk_from_oracle(pk, curve_id):
    L = Layout(curve_id)
    a1 = lambda pts: L.enc_affine(pts, 1)
    a2 = lambda pts: L.enc_affine(pts, 2)
    return T.ProvingKey(
        curve_id=curve_id, domain_cardinality=pk["domain_size"],
        domain_generator=L.enc_fr([pk["generator"]]), domain_coset_gen=L.enc_fr([pk["coset_gen"]]),
        g1_alpha=a1([pk["G1"]["Alpha"]]), g1_beta=a1([pk["G1"]["Beta"]]), g1_delta=a1([pk["G1"]["Delta"]]),
        g1_A=a1(pk["G1"]["A"]), g1_B=a1(pk["G1"]["B"]), g1_Z=a1(pk["G1"]["Z"]), g1_K=a1(pk["G1"]["K"]),
        g2_beta=a2([pk["G2"]["Beta"]]), g2_delta=a2([pk["G2"]["Delta"]]), g2_B=a2(pk["G2"]["B"]),
        infinity_a=np.array(pk["InfinityA"], dtype=np.uint8), infinity_b=np.array(pk["InfinityB"], dtype=np.uint8),
        commitment_keys=[{"Basis": a1(k["Basis"]), "BasisExpSigma": a1(k["BasisExpSigma"])} for k in pk["CommitmentKeys"]],
    )


def vk_from_oracle(vk, curve_id):
    """oracle.groth16.verifying_key(...) -> davinci_node_b200.setup.VerifyingKey"""
    from davinci_node_b200.setup import VerifyingKey
    L = Layout(curve_id)
    a1 = lambda pts: L.enc_affine(pts, 1)
    a2 = lambda pts: L.enc_affine(pts, 2)
    return VerifyingKey(curve_id=curve_id, g1_alpha=a1([vk["G1"]["Alpha"]]), g1_K=a1(vk["G1"]["K"]),
                        g2_beta=a2([vk["G2"]["Beta"]]), g2_gamma=a2([vk["G2"]["Gamma"]]), g2_delta=a2([vk["G2"]["Delta"]]),
                        commitment_keys=[{"G": a2([k["G"]]), "GSigmaNeg": a2([k["GSigmaNeg"]])} for k in vk["CommitmentKeys"]],
                        public_and_commitment_committed=[list(w) for w in vk["public_committed"]])


def proof_from_oracle(proof, curve_id):
    """oracle.groth16.prove(...) point dict -> davinci_node_b200.gnark_types.Proof"""
    L = Layout(curve_id)
    p = T.Proof(curve_id)
    p.Ar, p.Krs, p.Bs = L.enc_affine([proof["Ar"]], 1), L.enc_affine([proof["Krs"]], 1), L.enc_affine([proof["Bs"]], 2)
    p.Commitments = [L.enc_affine([c], 1) for c in proof["Commitments"]]
    p.CommitmentPok = L.enc_affine([proof.get("CommitmentPok")], 1)
    return p
