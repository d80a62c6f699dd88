"""CPU: pin the oracle (oracle/sv_ref.py) against fixtures produced by the reference itself, and the
host-side integer plans (a1) bit-exact against the reference's lists."""
import numpy as np
import pytest
import torch

from conftest import case_id, load_golden
from helpers import TOL, assert_close, build, cdtype, golden_out, rdtype
from oracle import sv_ref

CASES = load_golden("sv_cases.json")
SMALL = [c for c in CASES if c["spec"]["num_qubits"] <= 12]
LARGE = [c for c in CASES if c["spec"]["num_qubits"] > 12]


def test_reference_own_goldens():
    """test/test_pytorch_backend.py:392-472 (values typed into the reference's tests)."""
    exp = {"expval": [0.85154057, 0.0], "probs_all": [[0.9257702, 0.0], [0.07422972, 0.0]], "probs_1": [1.0, 0.0],
           "state": [[0.9620366 + 0.01599429j, 0.0], [0.05779156 - 0.26625147j, 0.0]]}
    got = [c for c in CASES if c["spec"]["name"] == "ref_golden_2q"]
    assert np.array_equal(np.round(np.float32(got[0]["out"][0]), 5), np.round(np.float32(exp["expval"]), 5))
    assert np.array_equal(np.round(np.float32(got[1]["out"][0][0]), 5), np.round(np.float32(exp["probs_all"]), 5))
    assert np.array_equal(np.round(np.float32(got[2]["out"][0][0]), 5), np.round(np.float32(exp["probs_1"]), 5))
    st = np.float32(got[3]["out"][0][0])
    assert np.array_equal(np.round(np.complex64(st[..., 0] + 1j * st[..., 1]), 5), np.round(np.complex64(exp["state"]), 5))
    # gradients of <Z0>: test_pytorch_backend.py:505-508
    circ = build(got[0])
    x = torch.tensor([0.54, 0.12], requires_grad=True)
    sv_ref.run_sv(circ, x)[0].backward()
    assert np.array_equal(np.round(x.grad.numpy(), 5), np.round(np.float32([-0.5104387, -0.10267819]), 5))


@pytest.mark.parametrize("case", SMALL, ids=case_id)
def test_oracle_matches_reference_fixture(case):
    dt = case["dtype"]
    circ = build(case, dt, case["flat"][0] if case["flat"] and case["flat"][0] else None)
    flat = torch.tensor(case["flat"], dtype=rdtype(dt)).reshape(len(case["flat"]), -1)
    ct = None
    if case["cotangent"]:
        ct = torch.tensor(case["cotangent"], dtype=rdtype(dt))
        if case["spec"]["meas"][0][0] == "state":
            ct = torch.view_as_complex(ct.contiguous())
    out, grad = sv_ref.run_batch(circ, flat, cdtype(dt), ct)
    # same algorithm, same torch kernels: agreement is far inside the parity tolerance
    assert_close(out.numpy(), golden_out(case), TOL[dt] * 0.1, "out")
    if grad is not None:
        assert_close(grad.numpy(), np.asarray(case["grad"]), TOL[dt] * 0.1, "grad")


@pytest.mark.parametrize("case", LARGE[:3], ids=case_id)
def test_oracle_matches_reference_fixture_large(case):
    test_oracle_matches_reference_fixture(case)


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_sv_plan_bit_exact(case):
    """_axeslist / _permutationlist equal the reference's (compiled_circuit.py:126-202), via the host mirror."""
    from tedq_b200.ir import build_ir

    ir = build_ir(build(case, case["dtype"]))
    assert [[list(a), list(b)] for a, b in ir.axeslist] == case["axeslist"]
    assert ir.permutationlist == case["permutationlist"]


def test_frontend_gate_matrices_match_reference():
    import tedq_b200 as qb

    for name, rec in load_golden("gate_matrices.json").items():
        nq = len(rec["re"]).bit_length() - 1
        op = getattr(qb, name)(*rec["params"], qubits=list(range(nq)), do_queue=False)
        ref = np.asarray(rec["re"]) + 1j * np.asarray(rec["im"])
        assert np.allclose(np.asarray(op.matrix, dtype=complex), ref, rtol=0, atol=1e-15), name


def test_oracle_matches_c4_depth20_fixture():
    """The oracle at BASELINE config 4 "depth 20" (17 819 gates, 16 qubits, complex128) against the reference's own
    output for the same parameters (forward only: keeps the CPU suite short; gradients are pinned on the GPU side)."""
    from tedq_b200 import workloads as W

    case = load_golden("c4d20_case.json")
    spec = W.mbl_2d(*case["spec"]["args"])
    circ = build(spec, "c128", case["flat"][0])
    with torch.no_grad():
        out = sv_ref.run_sv(circ, torch.tensor(case["flat"][0], dtype=torch.float64), torch.complex128)
    assert_close(out.numpy()[None], np.asarray(case["out"]), 1e-12, "c4d20 out")
