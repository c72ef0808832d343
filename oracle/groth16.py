"""Groth16 setup / prove / verify oracle (big-int).  TEST INFRASTRUCTURE ONLY.

Restates gnark's `groth16.Setup` and `groth16.Prove` (third-party module gnark
v0.14.1-0.20260126121332-407111efab55, go.mod:15, reached from
/root/reference/prover/prover_cpu.go:37,57 and prover/setup.go:25-27) with the BSB22 commitment
extension, over a SYNTHETIC trusted setup whose toxic waste is known, so every proof element also
has a closed form in the exponent:

    Ar  = (alpha + sum_i w_i A_i(tau) + r delta) G1
    Bs  = (beta  + sum_i w_i B_i(tau) + s delta) G2
    Krs = (sum_priv w_i K_i + h(tau) Z(tau)/delta - r s delta + s Ar + r Bs1) G1

Parity status: gnark's own proof BYTES are unpinned (the reference never pins any, SURVEY.md fact 4, and gnark cannot
be built here).  Everything else is pinned: `verify` is gnark's pairing verifier and `solidity_verify_proof` a
line-by-line port of the contract the reference deploys (config/statetransition_vkey.sol:653-746) over the pairing
oracle (oracle/pairing.py, itself pinned to reference-held material); `verify_exponent` re-checks the same equation in
the exponent.  The commitment challenges use gnark's real hash-to-field functions (oracle/hashes.py).
"""
import random
from dataclasses import dataclass, field

from . import curve as C
from . import ntt as N


# ----------------------------------------------------------------------------- synthetic R1CS
@dataclass
class R1CS:
    nb_wires: int
    nb_public: int                      # includes wire 0 (constant one)
    L: list                             # per constraint: list of (wire, coeff)
    R: list
    O: list
    commitments: list = field(default_factory=list)   # dicts: private_committed [wires], commitment_index wire

    @property
    def nb_constraints(self):
        return len(self.L)


def synthetic_circuit(nb_constraints, nb_public, q, seed, n_commit=0, n_private_committed=0, mix="witness",
                      n_public_committed=0):
    """'N multiplications + BSB22 commitments' circuit in the style of
    /root/reference/circuits/test/statetransition/statetransition_dummy.go:23-57, with a satisfying
    assignment.  Wire layout follows gnark: [one, public.., secret/internal..].

    mix = 'witness': 40% zero / 20% one / 25% < 2^64 / 15% full-width free wires;
    mix = 'uniform': all free wires uniform."""
    rnd = random.Random(seed)
    W = [1] + [rnd.randrange(q) for _ in range(nb_public - 1)]
    L, R, O = [], [], []

    def free_value():
        if mix == "uniform":
            return rnd.randrange(q)
        u = rnd.random()
        if u < 0.40:
            return 0
        if u < 0.60:
            return 1
        if u < 0.85:
            return rnd.randrange(1 << 64)
        return rnd.randrange(q)

    # free private wires (committed ones come first among them)
    n_free = max(4, nb_constraints // 4, n_commit * n_private_committed + 2)
    first_free = len(W)
    for _ in range(n_free):
        W.append(free_value())
    commitments = []
    nxt = first_free
    for ci in range(n_commit):
        priv = list(range(nxt, nxt + n_private_committed))
        nxt += n_private_committed
        # public committed wires only enter the challenge hash (gnark: PublicAndCommitmentCommitted)
        commitments.append({"private_committed": priv, "commitment_index": None,
                            "public_committed": list(range(1, 1 + min(n_public_committed, nb_public - 1)))})
    # commitment wires: value chosen by the (host-side) hash in the real prover; any value is a
    # valid assignment for the synthetic circuit because it only enters as a free factor.
    for cm in commitments:
        cm["commitment_index"] = len(W)
        W.append(rnd.randrange(q))

    def lin(max_terms):
        k = rnd.randint(1, max_terms)
        terms = {}
        for _ in range(k):
            w = rnd.randrange(len(W))
            terms[w] = (terms.get(w, 0) + rnd.choice([1, 1, 1, 2, 3, q - 1])) % q
        return [(w, cf) for w, cf in sorted(terms.items()) if cf]

    def ev(terms):
        return sum(W[w] * cf for w, cf in terms) % q

    for k in range(nb_constraints):
        l, r = lin(2), lin(2)
        if commitments and k < len(commitments):
            l = [(commitments[k]["commitment_index"], 1)]       # make the commitment wire matter
        if not l:
            l = [(0, 1)]
        if not r:
            r = [(0, 1)]
        o = len(W)
        W.append(ev(l) * ev(r) % q)
        L.append(l)
        R.append(r)
        O.append([(o, 1)])
    cs = R1CS(nb_wires=len(W), nb_public=nb_public, L=L, R=R, O=O, commitments=commitments)
    return cs, W


def constraint_values(cs: R1CS, W, q):
    ev = lambda terms: sum(W[w] * cf for w, cf in terms) % q
    a = [ev(t) for t in cs.L]
    b = [ev(t) for t in cs.R]
    c = [ev(t) for t in cs.O]
    assert all((x * y - z) % q == 0 for x, y, z in zip(a, b, c)), "assignment does not satisfy the R1CS"
    return a, b, c


# ----------------------------------------------------------------------------- setup
@dataclass
class Toxic:
    tau: int
    alpha: int
    beta: int
    gamma: int
    delta: int
    sigmas: list


def lagrange_at(dom: N.Domain, tau):
    """L_j(tau) for j < n over the domain <omega>."""
    q, n = dom.q, dom.n
    zn = (pow(tau, n, q) - 1) % q
    out = []
    wj = 1
    for _ in range(n):
        out.append(zn * wj % q * pow(n * (tau - wj) % q, -1, q) % q)
        wj = wj * dom.omega % q
    return out


def setup_exponents(cs: R1CS, cx: C.CurveCtx, tox: Toxic):
    """Discrete logs of every proving / verifying key element (gnark Setup)."""
    q = cx.r
    n = 1
    while n < cs.nb_constraints:
        n *= 2
    dom = N.Domain(cx.c, n)
    lag = lagrange_at(dom, tox.tau)
    m = cs.nb_wires
    A, B, Cc = [0] * m, [0] * m, [0] * m
    for k in range(cs.nb_constraints):
        for w, cf in cs.L[k]:
            A[w] = (A[w] + cf * lag[k]) % q
        for w, cf in cs.R[k]:
            B[w] = (B[w] + cf * lag[k]) % q
        for w, cf in cs.O[k]:
            Cc[w] = (Cc[w] + cf * lag[k]) % q
    dinv, ginv = pow(tox.delta, -1, q), pow(tox.gamma, -1, q)
    Kfull = [(tox.beta * A[i] + tox.alpha * B[i] + Cc[i]) % q for i in range(m)]
    committed = set()
    commit_wires = set()
    for cm in cs.commitments:
        committed.update(cm["private_committed"])
        commit_wires.add(cm["commitment_index"])
    pub = [i for i in range(m) if i < cs.nb_public or i in commit_wires]
    priv = [i for i in range(cs.nb_public, m) if i not in committed and i not in commit_wires]
    zn = (pow(tox.tau, n, q) - 1) % q
    Z_nat = [pow(tox.tau, j, q) * zn % q * dinv % q for j in range(n)]
    Z = N.bit_reverse_list(Z_nat)[: n - 1]
    return {
        "n": n, "dom": dom, "A": A, "B": B, "C": Cc,
        "pk_K": [Kfull[i] * dinv % q for i in priv], "priv_wires": priv,
        "vk_K": [Kfull[i] * ginv % q for i in pub], "pub_wires": pub,
        "basis": [[Kfull[i] * ginv % q for i in cm["private_committed"]] for cm in cs.commitments],
        "Z": Z,
    }


def setup(cs: R1CS, cx: C.CurveCtx, tox: Toxic):
    """Proving key as points (affine tuples), gnark field names."""
    ex = setup_exponents(cs, cx, tox)
    G1, G2, g1, g2 = cx.G1, cx.G2, cx.g1, cx.g2
    mul1 = lambda k: G1.mul(g1, k)
    mul2 = lambda k: G2.mul(g2, k)
    infA = [a == 0 for a in ex["A"]]
    infB = [b == 0 for b in ex["B"]]
    pk = {
        "domain_size": ex["n"], "generator": ex["dom"].omega, "coset_gen": ex["dom"].g,
        "G1": {
            "Alpha": mul1(tox.alpha), "Beta": mul1(tox.beta), "Delta": mul1(tox.delta),
            "A": [mul1(a) for a, inf in zip(ex["A"], infA) if not inf],
            "B": [mul1(b) for b, inf in zip(ex["B"], infB) if not inf],
            "K": [mul1(k) for k in ex["pk_K"]],
            "Z": [mul1(z) for z in ex["Z"]],
        },
        "G2": {
            "Beta": mul2(tox.beta), "Delta": mul2(tox.delta),
            "B": [mul2(b) for b, inf in zip(ex["B"], infB) if not inf],
        },
        "InfinityA": infA, "InfinityB": infB,
        "CommitmentKeys": [
            {"Basis": [mul1(k) for k in basis], "BasisExpSigma": [mul1(k * sg % cx.r) for k in basis]}
            for basis, sg in zip(ex["basis"], tox.sigmas)
        ],
    }
    return pk, ex


def verifying_key(cs: R1CS, cx: C.CurveCtx, tox: Toxic, ex=None):
    """gnark groth16.VerifyingKey as points: G1 {Alpha, K[]}, G2 {Beta, Gamma, Delta}, and the Pedersen verifying keys
    {G, GSigmaNeg = -sigma_i G} (gnark-crypto pedersen.VerifyingKey; the pairing check of
    config/statetransition_vkey.sol:678-699 pairs the commitment with GSigmaNeg and the proof of knowledge with G)."""
    ex = ex or setup_exponents(cs, cx, tox)
    G1, G2, g1, g2 = cx.G1, cx.G2, cx.g1, cx.g2
    return {
        "G1": {"Alpha": G1.mul(g1, tox.alpha), "K": [G1.mul(g1, k) for k in ex["vk_K"]]},
        "G2": {"Beta": G2.mul(g2, tox.beta), "Gamma": G2.mul(g2, tox.gamma), "Delta": G2.mul(g2, tox.delta)},
        "CommitmentKeys": [{"G": g2, "GSigmaNeg": G2.neg(G2.mul(g2, sg))} for sg in tox.sigmas[:len(cs.commitments)]],
        "nb_public": cs.nb_public,
        "public_committed": [cm.get("public_committed", []) for cm in cs.commitments],
    }


# ----------------------------------------------------------------------------- verify (pairings)
def verify(vk, proof, public_witness, cx: C.CurveCtx, hash_kind="default"):
    """gnark groth16.Verify (reached from /root/reference/circuits/artifacts.go:595-613 right after every proof):
    recomputes the commitment challenges from the proof's commitments and the public committed inputs, checks the
    Pedersen proof of knowledge (folded with the "G16-BSB22" challenge when there are several commitments) and
        e(Ar, Bs) = e(alpha, beta) * e(L, gamma) * e(Krs, delta),   L = sum_i pub_i K_i + sum_j chal_j K_(np+j) + sum_j C_j.
    public_witness: the public wire values WITHOUT the constant-one wire (gnark's public witness vector)."""
    from . import hashes as H
    from . import pairing
    pr = pairing.get(cx.name)
    G1 = cx.G1
    q = cx.r
    pub = [1] + [int(v) % q for v in public_witness]
    assert len(pub) == vk["nb_public"]
    comms = proof["Commitments"]
    if len(comms) != len(vk["CommitmentKeys"]):
        return False
    chals = [H.commitment_challenge(hash_kind, cm, [pub[w] for w in wires], q, cx.p)
             for cm, wires in zip(comms, vk["public_committed"])]
    if comms:
        fold = H.fold_challenge(chals, q) if len(comms) > 1 else 1
        pairs, f = [], 1
        for cm, key in zip(comms, vk["CommitmentKeys"]):
            pairs.append((G1.mul(cm, f), key["GSigmaNeg"]))
            f = f * fold % q
        pairs.append((proof["CommitmentPok"], vk["CommitmentKeys"][0]["G"]))
        if not pr.product_is_one(pairs):
            return False
    L = G1.msm_naive(vk["G1"]["K"], pub + chals)
    for cm in comms:
        L = G1.add(L, cm)
    return pr.product_is_one([(G1.neg(proof["Ar"]), proof["Bs"]), (vk["G1"]["Alpha"], vk["G2"]["Beta"]),
                              (L, vk["G2"]["Gamma"]), (proof["Krs"], vk["G2"]["Delta"])])


# ----------------------------------------------------------------------------- Solidity verifier (port)
def solidity_constants(vk, cx: C.CurveCtx):
    """The constant block gnark's Solidity exporter writes for a BN254 key with ONE commitment
    (/root/reference/config/statetransition_vkey.sol:60-115): G2 points negated, Fp2 coordinates as (_0 real, _1 imaginary)."""
    assert cx.name == "bn254" and len(vk["CommitmentKeys"]) == 1
    G2 = cx.G2
    c = {"ALPHA_X": vk["G1"]["Alpha"][0], "ALPHA_Y": vk["G1"]["Alpha"][1]}

    def put2(name, pt):
        (x0, x1), (y0, y1) = pt
        c[name + "_X_0"], c[name + "_X_1"], c[name + "_Y_0"], c[name + "_Y_1"] = x0, x1, y0, y1

    put2("BETA_NEG", G2.neg(vk["G2"]["Beta"]))
    put2("GAMMA_NEG", G2.neg(vk["G2"]["Gamma"]))
    put2("DELTA_NEG", G2.neg(vk["G2"]["Delta"]))
    put2("PEDERSEN_G", vk["CommitmentKeys"][0]["G"])
    put2("PEDERSEN_GSIGMANEG", vk["CommitmentKeys"][0]["GSigmaNeg"])
    K = vk["G1"]["K"]
    c["CONSTANT_X"], c["CONSTANT_Y"] = K[0] or (0, 0)
    for i, pt in enumerate(K[1:]):
        c["PUB_%d_X" % i], c["PUB_%d_Y" % i] = pt or (0, 0)        # (0, 0) is the precompiles' point at infinity
    return c


def solidity_verify_proof(c, proof8, commitments2, pok2, inputs, committed_input_indices, cx: C.CurveCtx):
    """Line-by-line port of `verifyProof` (/root/reference/config/statetransition_vkey.sol:653-746) over the constant
    block `c`: proof8 = A.x, A.y, B.x1, B.x0, B.y1, B.y0, C.x, C.y (EIP-197 order), commitments2 / pok2 = G1 points as
    (x, y), inputs = the public inputs, committed_input_indices = which inputs are hashed into the commitment
    challenge (the contract hard-codes input[2], `:663-666`).  Returns True where the contract would not revert."""
    from . import hashes as H
    from . import pairing
    P_, R_ = cx.p, cx.r
    pr = pairing.get("bn254")
    G1 = cx.G1
    if any(not (0 <= int(v) < R_) for v in inputs):
        return False                                   # PublicInputNotInField
    # HashToField (:660-677): keccak256(abi.encodePacked(commitments[0], commitments[1], publicAndCommitmentCommitted)) % R
    packed = b"".join(int(v).to_bytes(32, "big") for v in list(commitments2) + [inputs[i] for i in committed_input_indices])
    public_commitment = int.from_bytes(H.keccak256(packed), "big") % R_

    def g1(x, y):
        if x == 0 and y == 0:
            return None
        pt = (x % P_, y % P_)
        if not G1.on_curve(pt):
            raise ValueError("precompile failure: point not on curve")
        return pt

    def g2(x1, x0, y1, y0):                            # precompile order: imaginary part first
        pt = ((x0, x1), (y0, y1))
        if not cx.G2.on_curve(pt):
            raise ValueError("precompile failure: G2 point not on curve")
        return pt

    cg2 = lambda n: g2(c[n + "_X_1"], c[n + "_X_0"], c[n + "_Y_1"], c[n + "_Y_0"])
    try:
        # Pedersen (:678-699): e(commitment, GSigmaNeg) * e(pok, G) == 1
        if not pr.product_is_one([(g1(*commitments2), cg2("PEDERSEN_GSIGMANEG")), (g1(*pok2), cg2("PEDERSEN_G"))]):
            return False                               # CommitmentInvalid
        # publicInputMSM (:393-493): CONSTANT + commitment + sum input_i PUB_i + publicCommitments[0] PUB_last
        acc = G1.add(g1(c["CONSTANT_X"], c["CONSTANT_Y"]), g1(*commitments2))
        for i, v in enumerate(list(inputs) + [public_commitment]):
            pub_i = g1(c["PUB_%d_X" % i], c["PUB_%d_Y" % i])
            if pub_i is not None:
                acc = G1.add(acc, G1.mul(pub_i, int(v)))
        A = g1(proof8[0], proof8[1])
        B = g2(proof8[2], proof8[3], proof8[4], proof8[5])
        Cc = g1(proof8[6], proof8[7])
        # (:700-746): e(A, B) * e(C, -delta) * e(alpha, -beta) * e(L_pub, -gamma) == 1
        return pr.product_is_one([(A, B), (Cc, cg2("DELTA_NEG")), (g1(c["ALPHA_X"], c["ALPHA_Y"]), cg2("BETA_NEG")),
                                  (acc, cg2("GAMMA_NEG"))])
    except ValueError:
        return False


# ----------------------------------------------------------------------------- prove
def prove(cs: R1CS, pk, W, r, s, cx: C.CurveCtx, fold_challenge=None):
    """gnark groth16.Prove with pinned (r, s); returns affine points."""
    q = cx.r
    G1, G2 = cx.G1, cx.G2
    a, b, c = constraint_values(cs, W, q)
    dom = N.Domain(cx.c, pk["domain_size"])
    assert dom.omega == pk["generator"]
    h = N.compute_h(a, b, c, dom)
    wA = [w for w, inf in zip(W, pk["InfinityA"]) if not inf]
    wB = [w for w, inf in zip(W, pk["InfinityB"]) if not inf]
    delta1 = pk["G1"]["Delta"]
    ar = G1.add(G1.add(G1.msm(pk["G1"]["A"], wA), pk["G1"]["Alpha"]), G1.mul(delta1, r))
    bs1 = G1.add(G1.add(G1.msm(pk["G1"]["B"], wB), pk["G1"]["Beta"]), G1.mul(delta1, s))
    bs = G2.add(G2.add(G2.msm(pk["G2"]["B"], wB), pk["G2"]["Beta"]), G2.mul(pk["G2"]["Delta"], s))
    committed = set()
    for cm in cs.commitments:
        committed.update(cm["private_committed"])
        committed.add(cm["commitment_index"])
    wK = [W[i] for i in range(cs.nb_public, cs.nb_wires) if i not in committed]
    krs = G1.msm(pk["G1"]["K"], wK)
    krs = G1.add(krs, G1.msm(pk["G1"]["Z"], h[: len(pk["G1"]["Z"])]))
    krs = G1.add(krs, G1.mul(delta1, (-r * s) % q))
    krs = G1.add(krs, G1.mul(ar, s))
    krs = G1.add(krs, G1.mul(bs1, r))
    commitments, pok = [], None
    for i, (cm, key) in enumerate(zip(cs.commitments, pk["CommitmentKeys"])):
        vals = [W[w] for w in cm["private_committed"]]
        commitments.append(G1.msm(key["Basis"], vals))
        p_i = G1.msm(key["BasisExpSigma"], vals)
        if i:
            p_i = G1.mul(p_i, pow(fold_challenge, i, q))
        pok = G1.add(pok, p_i)
    return {"Ar": ar, "Bs": bs, "Krs": krs, "Commitments": commitments, "CommitmentPok": pok, "h": h}


# ----------------------------------------------------------------------------- checks in the exponent
def proof_exponents(cs: R1CS, ex, tox: Toxic, W, r, s, q):
    """Closed-form discrete logs of (Ar, Bs, Krs) - O(n) field work, no group operations."""
    a, b, c = constraint_values(cs, W, q)
    h = N.compute_h(a, b, c, ex["dom"])
    A = (tox.alpha + sum(w * x for w, x in zip(W, ex["A"])) + r * tox.delta) % q
    B = (tox.beta + sum(w * x for w, x in zip(W, ex["B"])) + s * tox.delta) % q
    kp = sum(W[i] * k for i, k in zip(ex["priv_wires"], ex["pk_K"])) % q
    hz = sum(x * z for x, z in zip(h, ex["Z"])) % q
    Cx = (kp + hz - r * s * tox.delta + s * A + r * B) % q
    return A, B, Cx


def verify_exponent(cs: R1CS, ex, tox: Toxic, W, A, B, Cx, q):
    """Groth16 verifier equation in the exponent: A*B = alpha*beta + gamma*L + delta*C."""
    Lx = sum(W[i] * k for i, k in zip(ex["pub_wires"], ex["vk_K"])) % q
    for cm, basis in zip(cs.commitments, ex["basis"]):
        Lx = (Lx + sum(W[w] * k for w, k in zip(cm["private_committed"], basis))) % q
    return (A * B - tox.alpha * tox.beta - tox.gamma * Lx - tox.delta * Cx) % q == 0
