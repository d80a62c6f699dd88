"""CPU, build container only: the drop-in boundary against the REAL reference front end (skipped where
/root/reference does not exist, e.g. on the GPU box)."""
import os
import sys
from unittest.mock import MagicMock

import pytest
import torch

REF = os.environ.get("TEDQ_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tedq")), reason="reference tree not present")


@pytest.fixture(scope="module")
def qai():
    for m in ["jax", "jax.numpy", "jaxlib", "qiskit", "qiskit.circuit", "quafu", "matplotlib", "matplotlib.patches",
              "matplotlib.pyplot", "toolz", "panel", "IPython", "IPython.display", "ray"]:
        sys.modules.setdefault(m, MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import tedq

    import tedq_b200

    tedq_b200.register_backend()
    return tedq


def test_reference_circuit_compiles_to_b200_backend(qai):
    import tedq_b200

    def circuitDef(*params):
        qai.RX(params[0], qubits=[1])
        qai.Hadamard(qubits=[0])
        qai.CNOT(qubits=[0, 1])
        qai.Rot(params[1], params[2], params[3], qubits=[0], trainable_params=[0, 2])
        return qai.expval(qai.PauliZ(qubits=[0]))

    ps = [torch.tensor([0.1 * (i + 1)]) for i in range(4)]
    circuit = qai.Circuit(circuitDef, 2, *ps)
    ref = circuit.compilecircuit(backend="pytorch")
    cc = circuit.compilecircuit(backend="pytorch_b200")
    assert isinstance(cc, tedq_b200.B200Backend)
    # same accessors as the reference backend (compiled_circuit.py:549-647)
    assert cc.gates_names == ref.gates_names
    assert cc.qubits == ref.qubits
    assert cc.backend == "pytorch_b200" and cc.interface == ref.interface and cc.diff_method == ref.diff_method
    assert len(cc.operators) == len(ref.operators) and cc.measurements is circuit.measurements
    assert cc._axeslist == ref._axeslist and cc._permutationlist == ref._permutationlist
    # positional binding with trainable_params=[0, 2] (test_compiled_circuit.py:166-210): 3 flat slots
    assert cc._ir.n_params == 3
    assert cc._ir.gates[3].param_idx == (1, -1, 2) and abs(cc._ir.gates[3].param_const[1] - 0.3) < 1e-7
    # other backend strings still reach the reference's own dispatch
    with pytest.raises(ValueError, match="unknown backend input"):
        circuit.compilecircuit(backend="nope")


def test_frontend_traces_the_same_gate_table_as_the_reference(qai):
    import tedq_b200 as qb
    from tedq_b200 import workloads as W
    from tedq_b200.ir import build_ir

    for spec in (W.qnn4(), W.mbl_1d(6), W.random_circuit(4, 30, seed=4, meas=[["probs", [1, 3]]])):
        a = build_ir(W.build_circuit(spec, qai))
        b = build_ir(W.build_circuit(spec, qb))
        assert a.n_params == b.n_params and len(a.gates) == len(b.gates)
        for ga, gb in zip(a.gates, b.gates):
            assert (ga.name, ga.kind, ga.qubits, ga.param_idx) == (gb.name, gb.kind, gb.qubits, gb.param_idx)
            assert ga.param_const == pytest.approx(gb.param_const)
            if ga.matrix is not None:
                assert abs(ga.matrix - gb.matrix).max() < 1e-15
        assert a.axeslist == b.axeslist and a.permutationlist == b.permutationlist


def test_reference_backend_accepts_b200_tree_plugins(qai):
    """The reference's OWN PyTorchBackend, with the B200 contraction tree injected through use_jdopttn= /
    use_cotengra= (compiled_circuit.py:356-393): constructors are called with the reference's arguments and build
    one plan per measurement network; the contraction itself needs a GPU (no CPU fallback: loud error here)."""
    import tedq_b200

    def circuitDef(a, b):
        qai.RX(a, qubits=[0])
        qai.Hadamard(qubits=[1])
        qai.CNOT(qubits=[0, 1])
        qai.RY(b, qubits=[1])
        return [qai.expval(qai.PauliZ(qubits=[0])), qai.expval(qai.PauliZ(qubits=[1]))]

    a, b = torch.tensor([0.3]), torch.tensor([0.7])
    circuit = qai.Circuit(circuitDef, 2, a, b)
    for kw in ({"use_jdopttn": tedq_b200.B200OptTN}, {"use_cotengra": tedq_b200.ctg_compat}):
        cc = circuit.compilecircuit(backend="pytorch", requires_grad=False, tn_simplify=False,
                                    hyper_opt={"max_repeats": 4, "slicing_opts": {"target_num_slices": 2}}, **kw)
        trees = cc._optimize_order_trees
        assert len(trees) == 2 and all(isinstance(t, tedq_b200.B200OptTN) for t in trees)
        assert all(t.info.n_slices >= 2 and len(t.inputs) == 2 + 4 + 1 + 4 + 2 for t in trees)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            cc(a, b)
