"""davinci-node_b200: B200-native (sm_100a) Groth16 proving backend behind davinci-node's
`prover/` interface.  The compute path is libb200groth16.so (hand-written CUDA behind a C ABI,
include/b200_groth16.h); this package is the host-side mirror of the reference's prover interface
plus ctypes bindings.  There is no CPU fallback: importing `capi` fails loudly if the library is
missing."""

__version__ = "0.1.0"

CURVE_IDS = {"bn254": 1, "bls12_377": 2, "bls12_381": 3, "bw6_761": 4}
