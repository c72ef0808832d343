"""CPU: host-side logic that needs no GPU - C ABI exports vs the header, curve constants vs the
oracle, layouts, MSM planning, the R1CS solver mirror."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from davinci_node_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "b200_groth16.h")).read()
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(capi.lib, name), "missing export " + name
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)


def test_no_gpu_means_loud_failure():
    """Without a visible GPU the backend must refuse to work (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from davinci_node_b200 import capi
    with pytest.raises(capi.B200Error):
        capi.init()


def test_curve_constants_match_oracle():
    from davinci_node_b200 import curve_consts, layout
    from oracle import curve as OC
    from oracle import params as OP
    from oracle import ntt as N
    for cid, c in OP.BY_ID.items():
        cx = OC.ctx(c.name)
        k = curve_consts.CONSTS[cid]
        assert k["g1"] == cx.g1 and k["g2"] == cx.g2
        L = layout.Layout(cid)
        assert (L.p, L.r, L.fp_l, L.fr_l) == (c.p, c.r, c.fp_limbs64, c.fr_limbs64)
        dom = N.Domain(c, 1 << 10)
        assert curve_consts.domain_constants(cid, 10) == (dom.omega, dom.g)


def test_layout_round_trip():
    from davinci_node_b200 import layout
    from oracle import curve as OC
    for name in ("bn254", "bw6_761"):
        L = layout.Layout(name)
        cx = OC.ctx(name)
        pts1 = [cx.g1, None, cx.G1.mul(cx.g1, 5)]
        pts2 = [cx.g2, None, cx.G2.mul(cx.g2, 7)]
        assert L.dec_affine(L.enc_affine(pts1, 1), 1) == pts1
        assert L.dec_affine(L.enc_affine(pts2, 2), 2) == pts2
        assert len(L.enc_affine(pts2, 2)) == 3 * L.affine_bytes(2)
        assert L.dec_fr(L.enc_fr([0, 1, L.r - 1])) == [0, 1, L.r - 1]
        # gnark Montgomery form: 1 encodes as R mod r, little-endian
        one = int.from_bytes(bytes(L.enc_fr([1])), "little")
        assert one == (1 << (64 * L.fr_l)) % L.r


def test_msm_plan():
    from davinci_node_b200 import capi
    p = capi.msm_plan(2, 1 << 22)
    assert p["nwin"] * p["c"] >= 254 and p["nb"] == 1 << (p["c"] - 1)
    assert capi.msm_plan(4, 1 << 22)["nwin"] * capi.msm_plan(4, 1 << 22)["c"] >= 378
    assert capi.msm_plan(3, 4096, 8) == dict(c=8, nwin=32, nb=128, task=p["task"] if False else capi.msm_plan(3, 4096, 8)["task"], group=capi.msm_plan(3, 4096, 8)["group"])


def test_solver_mirror_and_witness_errors():
    from davinci_node_b200 import gnark_types as T
    from oracle import groth16 as OG
    from oracle import params as OP
    from oracle_bridge import ccs_from_oracle
    q = OP.BN254.r
    cs, W = OG.synthetic_circuit(20, 3, q, seed=2)
    ccs = ccs_from_oracle(cs, 1)
    w = T.Witness(1, W[1:cs.nb_public], W[cs.nb_public:cs.nb_public + ccs.nb_secret])
    sol = ccs.solve(w)
    assert sol.values == W
    a, b, c = OG.constraint_values(cs, W, q)
    from davinci_node_b200.layout import Layout
    assert Layout(1).dec_fr(sol.A) == a and Layout(1).dec_fr(sol.C) == c
    with pytest.raises(ValueError):
        ccs.solve(T.Witness(1, W[1:2], []))


def test_bench_reference_arm_contract_and_no_cpu_fallback():
    """bench.py --impl reference (the CPU restatement timed on the host cores) prints ONE JSON line with the
    contract's keys; the B200 arm refuses to run without a GPU instead of falling back to the CPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1")      # what torchrun sets: the arm must ignore it
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "10", "--warmup", "1",
                        "--logn", "12"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    cb = d["cpu_baseline"]
    assert d["impl"] == "reference" and cb["kind"] == "port" and d["steps"] == 10 and d["warmup"] == 1
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 0
    # the warm-up step proved the whole workload once; ten component samples add up to one proof
    assert cb["full_proof_seconds"] > 0 and abs(cb["proofs_equivalent_timed"] - 1.0) < 1e-9
    # the arm never maps the product library (the driver records the .so files each arm loads)
    probe = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--logn', '10'];"
             "runpy.run_path(%r, run_name='__main__');"
             "maps = open('/proc/self/maps').read();"
             "assert 'libb200groth16' not in maps and 'liboracle_c' in maps;"
             "assert not any(m.startswith('davinci_node_b200') for m in sys.modules)") % os.path.join(root, "bench.py")
    r = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1", "--logn", "10"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def test_solidity_abi_encoding():
    """solidity/solidity.go:29-116 mirror: word order (B's imaginary parts first), 384-byte static ABI layout."""
    import json
    from davinci_node_b200 import gnark_types as T, solidity
    from davinci_node_b200.layout import Layout
    from oracle import curve as OC
    cx = OC.ctx("bn254")
    L = Layout(1)
    g1 = lambda k: cx.G1.mul(cx.g1, k)
    proof = T.Proof(1)
    proof.Ar, proof.Krs, proof.CommitmentPok = L.enc_affine([g1(3)], 1), L.enc_affine([g1(5)], 1), L.enc_affine([g1(7)], 1)
    bs = cx.G2.mul(cx.g2, 11)
    proof.Bs = L.enc_affine([bs], 2)
    proof.Commitments = [L.enc_affine([g1(9)], 1)]
    sp = solidity.Groth16CommitmentProof().FromGnarkProof(proof)
    data = sp.ABIEncode()
    assert len(data) == 12 * 32
    w = [int.from_bytes(data[32 * i:32 * i + 32], "big") for i in range(12)]
    assert w[0:2] == list(g1(3)) and w[6:8] == list(g1(5)) and w[8:10] == list(g1(9)) and w[10:12] == list(g1(7))
    assert w[2:6] == [bs[0][1], bs[0][0], bs[1][1], bs[1][0]]
    assert solidity.Groth16CommitmentProof.ABIDecode(data).words() == w
    assert json.loads(sp.String())["proof"]["Bs"][0] == [bs[0][1], bs[0][0]]
    bad = T.Proof(2)
    with pytest.raises(solidity.SolidityProofError):
        solidity.Groth16CommitmentProof().FromGnarkProof(bad)


def test_product_and_build_never_import_the_oracle():
    """oracle/ is test infrastructure: nothing under davinci-node_b200/ (package, build recipe, CUDA sources) and not the
    field generator the build runs may import, include or execute it; bench.py may only through its cpu_baseline /
    reference legs (oracle.cport)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b|#\s*include\s+[\"<].*oracle)", re.M)
    offenders = []
    for base, _, files in os.walk(os.path.join(root, "davinci-node_b200")):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".go", ".tmpl")):
                txt = open(os.path.join(base, f), errors="replace").read()
                if pat.search(txt) or "oracle/" in txt and "dlopen" in txt:
                    offenders.append(os.path.join(base, f))
    gen = open(os.path.join(root, "tools", "gen_field.py")).read()
    if pat.search(gen):
        offenders.append("tools/gen_field.py")
    assert not offenders, offenders
    # bench.py: oracle imports only inside the CPU-baseline prover (class CpuProver), which the B200 arm's timed region
    # never touches
    bench = open(os.path.join(root, "bench.py")).read()
    lo, hi = bench.index("class CpuProver"), bench.index("def cpu_reference_run")
    for m in re.finditer(r"^\s*(from\s+oracle\b|import\s+oracle\b)", bench, re.M):
        assert lo < m.start() < hi, "oracle imported outside CpuProver: " + bench[m.start():m.start() + 60]
