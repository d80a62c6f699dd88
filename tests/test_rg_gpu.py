"""GPU parity of the register-group sweeps (plan_opts["structure"] = 2, csrc/tq_sv_rg.cuh): same values and gradients as
the reference fixtures and the oracle, in the whole-state kernels and in forced HBM tiles; circuits that do not
qualify fall back to the default sweeps."""
import numpy as np
import pytest
import torch

import tedq_b200 as qb
from conftest import case_id, load_golden
from helpers import TOL, assert_close, build, cdtype, golden_out, rdtype
from oracle import sv_ref
from tedq_b200 import workloads as W
from test_engine_gpu import run_engine

pytestmark = pytest.mark.gpu
CASES = load_golden("sv_cases.json")
RG_CASES = [c for c in CASES if c["dtype"] == "c64" and c["spec"]["num_qubits"] >= 9]
TILES = {"whole": {}, "tiled": {"max_local_qubits_fwd": 9, "max_local_qubits_bwd": 9, "coalesce_bits": 3},
         "tiled10": {"max_local_qubits_fwd": 11, "max_local_qubits_bwd": 10, "coalesce_bits": 2}}


def qualifies(spec):
    return not ({g[0] for g in spec["gates"]} & {"SWAP", "CSWAP"})


@pytest.mark.parametrize("case", RG_CASES, ids=case_id)
@pytest.mark.parametrize("tiles", list(TILES), ids=list(TILES))
def test_register_group_sweeps_match_reference_fixture(case, tiles):
    n = case["spec"]["num_qubits"]
    opts = dict(TILES[tiles], structure=2)
    if tiles != "whole" and n <= opts["max_local_qubits_fwd"]:
        pytest.skip("state fits one tile")
    out, grad = run_engine(case, opts)
    assert_close(out, golden_out(case), TOL["c64"], "out")
    if grad is not None:
        assert_close(grad, np.asarray(case["grad"]), TOL["c64"], "grad")
    cc = build(case, "c64", case["flat"][0]).compilecircuit(backend="pytorch_b200", dtype=torch.complex64, plan_opts=opts)
    groups = cc.plan().num_register_groups(False), cc.plan().num_register_groups(True)
    if qualifies(case["spec"]):
        assert groups[0] > 0 and groups[1] > 0, "the plan did not take the register-group sweeps"
    else:
        assert groups == (0, 0)


POOL = ["Hadamard", "PauliX", "PauliY", "PauliZ", "S", "T", "SX", "RX", "RY", "RZ", "Rot", "PhaseShift",
        "CNOT", "CZ", "CY", "ControlledPhaseShift", "CRX", "CRY", "CRZ", "Toffoli"]
MEAS = [
    [["expval", [["PauliZ", [0]]]], ["expval", [["PauliX", [3]]]], ["expval", [["PauliZ", [1]], ["PauliZ", [7]]]]],
    [["probs", [2, 5]]],
    [["state"]],
    [["probs", None]],
]


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("tiles", ["whole", "tiled"])
def test_register_group_sweeps_match_oracle_on_random_circuits(seed, tiles):
    """Every gate kind a register group can hold (dense / diagonal one-target blocks with 0-2 controls, controlled-X
    swaps, multi-target diagonals), every measurement kind, 3 parameter sets."""
    n = 10 + seed % 3
    spec = W.random_circuit(n, 70 + 10 * seed, 300 + seed, gate_pool=POOL, meas=MEAS[seed % len(MEAS)])
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float32))
    opts = dict(TILES[tiles], structure=2)
    cc = circ.compilecircuit(backend="pytorch_b200", dtype=torch.complex64, plan_opts=opts)
    flat = torch.tensor(np.random.RandomState(seed).uniform(-np.pi, np.pi, (3, spec["n_params"])), dtype=torch.float32)
    x = flat.cuda().requires_grad_(True)
    y = cc.batched(x)
    yr = torch.view_as_real(y) if y.is_complex() else y
    ct = torch.tensor(np.random.RandomState(100 + seed).uniform(-1, 1, tuple(yr.shape[1:])), dtype=torch.float32)
    (yr * ct.cuda()).sum().backward()
    ref_y, ref_g = sv_ref.run_batch(circ, flat, torch.complex64, torch.view_as_complex(ct.contiguous()) if y.is_complex() else ct)
    assert cc.plan().num_register_groups(False) > 0 and cc.plan().num_register_groups(True) > 0
    assert_close(y.detach().cpu().numpy().reshape(3, -1), ref_y.numpy().reshape(3, -1), 1e-5, "out")
    assert_close(x.grad.cpu().numpy(), ref_g.numpy(), 4e-5, "grad")


def test_register_group_request_falls_back_when_the_circuit_does_not_qualify():
    spec = W.random_circuit(10, 40, 5, gate_pool=["RX", "SWAP", "CNOT", "RY"])
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float32))
    flat = torch.rand(2, spec["n_params"])
    outs = []
    for opts in (None, {"structure": 2}):
        cc = circ.compilecircuit(backend="pytorch_b200", dtype=torch.complex64, plan_opts=opts)
        outs.append(cc.batched(flat.cuda()).cpu().numpy())
        assert cc.plan().num_register_groups(False) == 0
    assert np.allclose(outs[0], outs[1], rtol=0, atol=1e-6)   # (shared-memory atomics: the last bit may differ run to run)


@pytest.mark.parametrize("case", RG_CASES, ids=case_id)
def test_automatic_choice_and_forced_default_sweeps(case):
    """plan_opts["structure"]: 0 (default) picks the register groups for the hardware-efficient ansatz circuits (the same
    arithmetic in half the passes over the tile) and the default sweeps for the many-body-localisation circuits (their pair blocks fold
    ~18 gates each); -1 forces the default sweeps, which still match the fixtures on every case."""
    name = case["spec"]["name"]
    cc = build(case, "c64", case["flat"][0]).compilecircuit(backend="pytorch_b200", dtype=torch.complex64)
    groups = cc.plan().num_register_groups(False)
    if name.startswith("hea"):
        assert groups > 0
    if name.startswith("mbl") or not qualifies(case["spec"]):
        assert groups == 0
    out, grad = run_engine(case, {"structure": -1})
    assert_close(out, golden_out(case), TOL["c64"], "out")
    if grad is not None:
        assert_close(grad, np.asarray(case["grad"]), TOL["c64"], "grad")
    cc = build(case, "c64", case["flat"][0]).compilecircuit(backend="pytorch_b200", dtype=torch.complex64,
                                                             plan_opts={"structure": -1})
    assert cc.plan().num_register_groups(False) == 0 and cc.plan().num_register_groups(True) == 0


def test_gradients_without_shared_memory_cells_match(tmp_path):
    """A sweep with thousands of trainable slots has no room for the gradient cells: the adjoint kernel then adds every
    warp's sums to the gradient in global memory.  TQ_RG_GRAD_DIRECT forces that path (read once per process, hence the
    subprocess); same gradients as the default sweeps, whole-state and tiled."""
    import os
    import subprocess
    import sys

    code = """
import sys, numpy as np, torch
sys.path.insert(0, %r)
import tedq_b200 as qb
from tedq_b200 import workloads as W
for n, opts in ((12, {}), (16, {})):
    spec = W.hea(n, 3)
    circ = W.build_circuit(spec, qb)
    flat = torch.tensor(np.random.RandomState(1).uniform(-3, 3, (2, spec["n_params"])), dtype=torch.float32)
    grads = []
    for structure in (-1, 2):
        cc = circ.compilecircuit(backend="pytorch_b200", plan_opts=dict(opts, structure=structure))
        x = flat.cuda().requires_grad_(True)
        y = cc.batched(x)
        (y * torch.arange(1, y.shape[1] + 1, device="cuda")).sum().backward()
        grads.append(x.grad.cpu().numpy())
        if structure == 2:
            assert cc.plan().num_register_groups(True) > 0
    err = float(np.abs(grads[0] - grads[1]).max())
    assert err < 2e-5, (n, err)
print("direct ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TQ_RG_GRAD_DIRECT="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "direct ok" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


@pytest.mark.parametrize("tiles", ["whole", "tiled"])
def test_register_group_sweeps_start_from_a_user_state(tiles):
    """InitStateVector (pytorch_backend.py:500-522) with the register-group sweeps: the first sweep copies the user's
    state into the swizzled tile; values and gradients against the oracle."""
    n = 11
    rng = np.random.RandomState(5)
    v = rng.randn(1 << n) + 1j * rng.randn(1 << n)
    v /= np.linalg.norm(v)

    def circuit_def(t):
        qb.InitStateVector(v)
        for q in range(n):
            qb.RY(t[q], qubits=[q])
        for q in range(n - 1):
            qb.CNOT(qubits=[q, q + 1])
        for q in range(n):
            qb.RZ(t[n + q], qubits=[q])
        return [qb.expval(qb.PauliZ(qubits=[q])) for q in (0, 5, 10)]

    flat = torch.tensor(rng.uniform(-2, 2, 2 * n), dtype=torch.float32)
    circ = qb.Circuit(circuit_def, n, flat)
    cc = circ.compilecircuit(backend="pytorch_b200", plan_opts=dict(TILES[tiles], structure=2))
    x = flat.cuda().requires_grad_(True)
    y = cc(x)
    ct = torch.tensor([0.5, -1.0, 2.0])
    (y * ct.cuda()).sum().backward()
    assert cc.plan().num_register_groups(False) > 0
    ref_y, ref_g = sv_ref.run_batch(circ, flat[None], torch.complex64, ct)
    assert_close(y.detach().cpu().numpy().reshape(-1), ref_y.numpy().reshape(-1), 1e-5, "out")
    assert_close(x.grad.cpu().numpy().reshape(-1), ref_g.numpy().reshape(-1), 4e-5, "grad")
