#!/usr/bin/env python3
"""NTT alone (for ncu captures and timing): python tools/ntt_probe.py [curve] [logn] [reps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from davinci_node_b200 import capi, layout, synthetic  # noqa: E402
from davinci_node_b200.curve_consts import domain_constants  # noqa: E402


def main():
    cname = sys.argv[1] if len(sys.argv) > 1 else "bls12_377"
    logn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    capi.init(1)
    L = layout.Layout(cname)
    n = 1 << logn
    rng = np.random.default_rng(1)
    buf = torch.from_numpy(synthetic.rand_canonical(rng, n, L.fr_l, L.r.bit_length()).view(np.uint8).reshape(-1)).cuda()
    omega, g = domain_constants(L.id, logn)
    gw, gc = L.enc_fr([omega]), L.enc_fr([g])
    dom = C.c_uint64(0)
    capi.check(capi.lib.b200_domain_create(L.id, n, gw.ctypes.data, gc.ctypes.data, C.byref(dom)))
    st = torch.cuda.current_stream().cuda_stream
    for inverse, dit, coset in ((0, 0, 0), (1, 1, 1)):
        capi.check(capi.lib.b200_ntt_dev(dom.value, buf.data_ptr(), inverse, dit, coset, st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            capi.check(capi.lib.b200_ntt_dev(dom.value, buf.data_ptr(), inverse, dit, coset, st))
        e1.record()
        torch.cuda.synchronize()
        print("ntt %s 2^%d inverse=%d dit=%d coset=%d: %.3f ms" % (cname, logn, inverse, dit, coset, e0.elapsed_time(e1) / reps))
    capi.check(capi.lib.b200_domain_release(dom.value))


if __name__ == "__main__":
    main()
