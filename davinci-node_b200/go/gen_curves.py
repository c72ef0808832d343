#!/usr/bin/env python3
"""Generates the per-curve Go files of the B200 prover shim from prover/curve.go.tmpl (one body, four curves - the
curves the reference's callGPUProver dispatches, /root/reference/prover/prover_gpu.go:24-61).

  python davinci-node_b200/go/gen_curves.py          # rewrites prover/prover_b200_<curve>.go
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CURVES = [
    # ID (Go identifier suffix), display name, gnark package directory, C enum
    ("BN254", "BN254", "bn254", "B200_BN254"),
    ("BLS12377", "BLS12-377", "bls12-377", "B200_BLS12_377"),
    ("BLS12381", "BLS12-381", "bls12-381", "B200_BLS12_381"),
    ("BW6761", "BW6-761", "bw6-761", "B200_BW6_761"),
]


def render(tmpl, ident, name, pkg, enum):
    return tmpl.replace("{{ID}}", ident).replace("{{NAME}}", name).replace("{{PKG}}", pkg).replace("{{ENUM}}", enum)


def outputs():
    tmpl = open(os.path.join(HERE, "prover", "curve.go.tmpl")).read()
    return {os.path.join(HERE, "prover", "prover_b200_%s.go" % pkg.replace("-", "")): render(tmpl, ident, name, pkg, enum)
            for ident, name, pkg, enum in CURVES}


if __name__ == "__main__":
    for path, text in outputs().items():
        with open(path, "w") as fh:
            fh.write(text)
        print("generated", os.path.relpath(path, HERE))
