"""GPU: behaviour a drop-in must have beyond arithmetic - re-entrancy (calls arrive from several
goroutines: sequencer/ballot.go:136, aggregate.go:447, finalizer.go:389), error reporting without a
CPU fallback, and pageable host buffers."""
import ctypes as C
import threading

import numpy as np
import pytest

from oracle import curve as OC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wl():
    from davinci_node_b200 import capi, prover, synthetic
    capi.init()
    w = synthetic.SyntheticWorkload("bn254", 13, seed=5)
    w.register()
    yield w
    prover.release_proving_key(w.pk)


def test_concurrent_proofs_are_deterministic(wl):
    """8 threads prove 3 different witnesses concurrently (4 slots per GPU); with pinned (r, s) every
    result must equal the single-threaded proof of the same witness."""
    import torch
    from davinci_node_b200 import capi
    L = wl.L
    sols = [wl.solution(seed=100 + i, pinned=False) for i in range(3)]
    r, s = 1234567 % L.r, 7654321 % L.r
    ref = []
    for sol in sols:
        pin, pout, out, keep = wl.prove_args(sol, r, s, on_device=False)
        capi.check(capi.lib.b200_prove(wl.handle, C.byref(pin), C.byref(pout), 0))
        ref.append(out.numpy().copy())
    assert not np.array_equal(ref[0], ref[1])
    results, errors = {}, []

    def work(tid):
        try:
            torch.cuda.set_device(0)
            for rep in range(3):
                k = (tid + rep) % 3
                pin, pout, out, keep = wl.prove_args(sols[k], r, s, on_device=False)
                capi.check(capi.lib.b200_prove(wl.handle, C.byref(pin), C.byref(pout), -1))
                results[(tid, rep)] = (k, out.numpy().copy())
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 24
    for (tid, rep), (k, got) in results.items():
        assert np.array_equal(got, ref[k]), (tid, rep, k)


def test_errors_are_reported_not_swallowed(wl):
    from davinci_node_b200 import capi
    L = wl.L
    sol = wl.solution(seed=1, pinned=False)
    pin, pout, out, keep = wl.prove_args(sol, 5, 7, on_device=False)
    pin.wires.len = wl.m - 1                                    # wrong wire-vector length
    assert capi.lib.b200_prove(wl.handle, C.byref(pin), C.byref(pout), 0) != 0
    assert b"nb_wires" in capi.lib.b200_last_error()
    pin.wires.len = wl.m
    assert capi.lib.b200_prove(0xDEAD, C.byref(pin), C.byref(pout), 0) != 0   # unknown handle
    assert b"handle" in capi.lib.b200_last_error()
    assert capi.lib.b200_prove(wl.handle, C.byref(pin), C.byref(pout), 7) != 0  # key not resident on device 7
    out8 = np.zeros(64, dtype=np.uint8)
    assert capi.lib.b200_msm(99, 1, None, None, 0, out8.ctypes.data, 0) != 0    # unknown curve
    assert capi.lib.b200_msm(1, 3, None, None, 0, out8.ctypes.data, 0) != 0     # bad group
    # after the failures the key still works
    capi.check(capi.lib.b200_prove(wl.handle, C.byref(pin), C.byref(pout), 0))


def test_pageable_and_pinned_inputs_agree(wl):
    from davinci_node_b200 import capi
    L = wl.L
    r, s = 11 % L.r, 13 % L.r
    outs = []
    for pinned in (True, False):
        sol = wl.solution(seed=77, pinned=pinned)
        pin, pout, out, keep = wl.prove_args(sol, r, s, on_device=False)
        capi.check(capi.lib.b200_prove(wl.handle, C.byref(pin), C.byref(pout), 0))
        outs.append(out.numpy().copy())
    assert np.array_equal(outs[0], outs[1])


def test_host_register_roundtrip():
    """b200_host_register page-locks caller memory (what a Go shim does with its long-lived solver buffers); a proof
    from registered numpy buffers equals the proof from pageable ones."""
    import ctypes as C
    import numpy as np
    import torch
    from davinci_node_b200 import capi, prover, synthetic
    capi.init()
    wl = synthetic.SyntheticWorkload("bn254", 12, seed=5)
    h = wl.register()
    try:
        sol = wl.solution(seed=3, pinned=False)
        L = wl.L
        r, s = 123456789 % L.r, 987654321 % L.r
        outs = []
        for register in (False, True):
            bufs = {k: np.array(sol[k].numpy(), copy=True) for k in ("W", "a", "b", "c")}
            if register:
                for v in bufs.values():
                    capi.check(capi.lib.b200_host_register(v.ctypes.data, v.nbytes))
            fake = dict(sol)
            fake.update({k: torch.from_numpy(v) for k, v in bufs.items()})
            pin, pout, out, keep = wl.prove_args(fake, r, s, on_device=False)
            capi.check(capi.lib.b200_prove(h, C.byref(pin), C.byref(pout), 0))
            outs.append(bytes(out.numpy()))
            if register:
                for v in bufs.values():
                    capi.check(capi.lib.b200_host_unregister(v.ctypes.data))
        assert outs[0] == outs[1]
        assert capi.lib.b200_host_register(None, 16) != 0
    finally:
        prover.release_proving_key(wl.pk)
