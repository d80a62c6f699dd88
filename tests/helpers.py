"""Shared helpers for the parity tests."""
import numpy as np
import torch

import tedq_b200 as qb
from tedq_b200 import workloads as W

# north_star tolerances: complex64 results/gradients within relative 1e-5, complex128 within 1e-11.
# Real outputs (expectation values, probabilities, gradients) are O(1) quantities that pass through zero; the
# reference itself rounds its goldens to 5 decimals (test_pytorch_backend.py:406-409), so their bound is
# |a-b| <= tol * max(1, |b|).  Complex outputs (state vectors, amplitudes) have entries of magnitude 2^-n/2: their
# bound is RELATIVE to the largest reference entry, |a-b| <= tol * max|b|.
TOL = {"c64": 1e-5, "c128": 1e-11}


def rdtype(dtype):
    return torch.float32 if dtype == "c64" else torch.float64


def cdtype(dtype):
    return torch.complex64 if dtype == "c64" else torch.complex128


def golden_out(case):
    out = np.asarray(case["out"], dtype=np.float64)
    if case["spec"]["meas"][0][0] == "state":
        out = out[..., 0] + 1j * out[..., 1]
    return out


def build(case_or_spec, dtype="c64", flat=None):
    spec = case_or_spec.get("spec", case_or_spec)
    wrap = lambda v: torch.tensor(float(v), dtype=rdtype(dtype))
    return W.build_circuit(spec, qb, flat, tensor_fn=wrap)


def assert_close(a, b, tol, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    if np.iscomplexobj(b) and b.size:
        scale = float(np.abs(b).max())
        assert float(err.max()) <= tol * scale, \
            f"{what}: max |a-b| = {float(err.max()):.3e} exceeds {tol:g}*max|b| = {tol * scale:.3e}"
        return
    bound = tol * np.maximum(1.0, np.abs(b))
    worst = float(np.max(err - bound)) if err.size else 0.0
    assert worst <= 0, f"{what}: max |a-b| = {float(err.max()):.3e} exceeds {tol:g}*max(1,|b|)"


def all_kinds_spec():
    """Every parametrised gate kind of the reference on an entangled 4-qubit state (controls live)."""
    b = W._Builder("hess4", 4)
    for q in range(4):
        b.g("RY", [q], b.p())
    b.g("CNOT", [0, 1])
    b.g("CRX", [1, 2], b.p())
    b.g("Rot", [3], b.p(), b.p(), b.p())
    b.g("CRY", [3, 0], b.p())
    b.g("ControlledPhaseShift", [2, 3], b.p())
    b.g("Hadamard", [2])
    b.g("CRZ", [0, 2], b.p())
    b.g("PhaseShift", [1], b.p())
    b.g("RZ", [0], b.p())
    b.g("CNOT", [2, 3])
    for q in range(4):
        b.g("RX", [q], b.p())
    for q in range(3):
        b.expval(("PauliZ", [q]))
    return b.spec
