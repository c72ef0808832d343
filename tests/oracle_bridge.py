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
