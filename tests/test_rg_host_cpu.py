"""CPU: the register-group arithmetic and index maps (csrc/tq_sv_rg.cuh) run as host code over whole swizzled tiles
against a gate-by-gate reference (tests/native/rg_check.cu): swizzle involution, bank-conflict freedom of the chosen
register bits, forward and adjoint application of random groups, gradient terms."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_register_group_host_check(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "rg_check")
    src = os.path.join(HERE, "native", "rg_check.cu")
    res = subprocess.run([nvcc, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w",
                          "-I", os.path.join(HERE, "..", "include"), src, "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0 and "rg_check: ok" in run.stdout, run.stdout[-2000:]
