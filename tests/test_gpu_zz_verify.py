"""GPU Groth16 verification (the reference verifies every proof right after proving,
/root/reference/circuits/artifacts.go:595-613): b200_pairing_check + the verifier mirror against the oracle.

STATUS: the pairing kernels were written after this round's GPU budget was spent.  Their template (csrc/pairing.cuh) is
verified bit for bit on the CPU (tests/test_host_pairing.py) and the host logic of the mirror in
tests/test_verifier_cpu.py, but this file had never run on hardware when it was committed - hence the non-strict xfail:
an XPASS in the round-end GPU run means the kernels work as verified on the CPU, an XFAIL that they do not yet.  The
work runs in its own process and is the last GPU test file, so a fault cannot disturb the rest of the suite."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="pairing kernels not yet run on hardware (GPU budget exhausted); CPU-verified template")
def test_gpu_pairing_and_groth16_verify():
    r = subprocess.run([sys.executable, os.path.join(HERE, "gpu_verify_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"]
    print("pairing / verify wall ms:", res["pairing_ms"], res["verify_ms"])


def _run_worker(flag, timeout):
    r = subprocess.run([sys.executable, os.path.join(HERE, "gpu_verify_worker.py"), flag], capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"]
    return res


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="batch pairing kernels not yet run on hardware (GPU budget exhausted); CPU-verified template")
def test_gpu_pairing_check_batch():
    """Generic inputs: every lane of a warp follows the same instruction stream."""
    res = _run_worker("--batch", 900)
    print("batched pairing checks per second:", res["checks_per_s"])


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="batch pairing kernels not yet run on hardware (GPU budget exhausted); CPU-verified template")
def test_gpu_pairing_check_batch_edge_cases():
    """Points at infinity and G1 points outside the subgroup inside a batch (lanes leave the Miller loop early)."""
    _run_worker("--batch-edge", 900)
