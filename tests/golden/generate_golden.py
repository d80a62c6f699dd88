"""Generate tests/golden/*.json by running the UNMODIFIED reference (TeD-Q) in the build container.

    python tests/golden/generate_golden.py            # needs /root/reference, CPU only

The reference cannot travel to the GPU box, so its outputs are committed as fixtures; this script is
the committed recipe that made them.  Missing optional imports of the reference (jax, qiskit,
matplotlib, ...) are stubbed with MagicMock exactly as SURVEY.md 8c describes; none of them is on the
pytorch state-vector path.  complex128 fixtures use the documented monkey-patch of the hard-coded
``tcomplex`` (pytorch_backend.py:38).
"""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

for _m in ["jax", "jax.numpy", "jaxlib", "qiskit", "qiskit.circuit", "quafu", "matplotlib", "matplotlib.patches",
           "matplotlib.pyplot", "toolz", "panel", "IPython", "IPython.display", "ray"]:
    sys.modules.setdefault(_m, MagicMock())
sys.path.insert(0, os.environ.get("TEDQ_REFERENCE", "/root/reference"))

import tedq as qai  # noqa: E402
import tedq.backends.pytorch_backend as ref_backend  # noqa: E402
from tedq.tensor_network import gen_tensor_networks  # noqa: E402

from tedq_b200 import workloads as W  # noqa: E402  (specs only; no engine code runs here)


def to_list(t):
    t = t.detach().cpu()
    if t.is_complex():
        return torch.view_as_real(t).numpy().tolist()
    return t.numpy().tolist()


def run_case(spec, flat_batch, dtype="c64", seed=0, with_grad=True):
    """Reference forward (+ backward of sum(cotangent*out)) for every parameter set in flat_batch."""
    rdt = torch.float32 if dtype == "c64" else torch.float64
    ref_backend.tcomplex = torch.complex64 if dtype == "c64" else torch.complex128
    wrap = lambda v: torch.tensor(float(v), dtype=rdt)
    circuit = W.build_circuit(spec, qai, flat_batch[0] if len(flat_batch) else None, tensor_fn=wrap)
    cc = circuit.compilecircuit(backend="pytorch")
    rng = np.random.RandomState(seed + 1234)
    outs, grads, cots = [], [], []
    for row in flat_batch:
        x = torch.tensor(row, dtype=rdt, requires_grad=with_grad and len(row) > 0)
        y = cc(x) if len(row) else cc()
        outs.append(to_list(y))
        if with_grad and len(row):
            if y.is_complex():
                ct = torch.tensor(rng.uniform(-1, 1, size=tuple(y.shape) + (2,)), dtype=rdt)
                loss = torch.sum(torch.view_as_real(y) * ct)
            else:
                ct = torch.tensor(rng.uniform(-1, 1, size=tuple(y.shape)), dtype=rdt)
                loss = torch.sum(y * ct)
            loss.backward()
            grads.append(x.grad.numpy().tolist())
            cots.append(ct.numpy().tolist())
    ref_backend.tcomplex = torch.complex64
    case = {"spec": spec, "dtype": dtype, "flat": [list(map(float, r)) for r in flat_batch], "out": outs,
            "cotangent": cots, "grad": grads,
            "axeslist": [[list(a), list(b)] for a, b in cc._axeslist],
            "permutationlist": [list(p) for p in cc._permutationlist]}
    return case


def tn_maps(spec):
    """(input_indices, output_indices, size_dict keys) per measurement from the reference's gen_tensor_networks."""
    circuit = W.build_circuit(spec, qai)
    cc = circuit.compilecircuit(backend="pytorch")
    tns = gen_tensor_networks(cc._num_qubits, cc._operators, cc._appliedqubits, cc._measurements)
    return [{"inputs": [list(ix) for ix in tn.input_indices], "output": list(tn.output_indices),
             "size_keys": list(tn.size_dict.keys())} for tn in tns]


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    cases = []

    # 1. the reference's own golden circuit (test_pytorch_backend.py:386-584)
    for meas in ([["expval", [["PauliZ", [0]]]], ["expval", [["PauliX", [1]]]]], [["probs", None]], [["probs", [1]]],
                 [["state"]]):
        spec = {"name": "ref_golden_2q", "num_qubits": 2, "n_params": 2,
                "gates": [["RX", [0], ["p0"]], ["RY", [0], ["p1"]]], "meas": meas}
        cases.append(run_case(spec, [[0.54, 0.12]]))

    # 2. every gate, every measurement kind, both precisions
    meas_sets = [
        [["expval", [["PauliZ", [0]]]], ["expval", [["PauliX", [1]]]], ["expval", [["PauliY", [2]]]],
         ["expval", [["Hadamard", [0]]]]],
        [["expval", [["PauliZ", [0]], ["PauliZ", [2]]]], ["expval", [["PauliX", [0]], ["PauliY", [1]]]]],
        [["probs", None]], [["probs", [2, 0]]], [["probs", [1]], ["probs", [0]]], [["state"]],
    ]
    rng = np.random.RandomState(7)
    k = 0
    for n in (3, 4, 5, 6):
        for ms in meas_sets:
            for dtype in ("c64", "c128"):
                spec = W.random_circuit(n, 36, seed=100 + k, meas=ms)
                flat = rng.uniform(-np.pi, np.pi, size=(2, spec["n_params"]))
                cases.append(run_case(spec, flat.tolist(), dtype, seed=k))
                k += 1
    # single-qubit and no-parameter corner cases
    cases.append(run_case(W.random_circuit(1, 12, seed=5, meas=[["state"]]), rng.uniform(-3, 3, size=(1, W.random_circuit(1, 12, seed=5)["n_params"])).tolist()))
    spec0 = {"name": "noparam", "num_qubits": 3, "n_params": 0,
             "gates": [["Hadamard", [0], []], ["CNOT", [0, 1], []], ["Toffoli", [0, 1, 2], []], ["RX", [2], [0.3]]],
             "meas": [["probs", None]]}
    cases.append(run_case(spec0, [[]], with_grad=False))

    # 3. BASELINE configs (reduced batch): C1, C2, C3, C4
    spec = W.qnn4()
    cases.append(run_case(spec, rng.uniform(0, 1, size=(4, spec["n_params"])).tolist()))
    spec = W.mbl_1d(12)
    cases.append(run_case(spec, W.c2_inputs(256, 12, 0)[[0, 255]].tolist()))
    spec = W.mbl_1d(8)
    cases.append(run_case(spec, W.c2_inputs(4, 8, 1).tolist()))
    for n, depth in ((15, 2), (16, 3), (17, 2)):
        spec = W.hea(n, depth)
        cases.append(run_case(spec, rng.uniform(0, 1, size=(1, spec["n_params"])).tolist()))
    spec = W.hea(20, 10)
    cases.append(run_case(spec, rng.uniform(0, 1, size=(1, spec["n_params"])).tolist()))
    spec = W.mbl_2d(4, 1)
    cases.append(run_case(spec, rng.uniform(0, 1, size=(1, spec["n_params"])).tolist(), "c128"))
    spec = W.mbl_2d(3, 2)
    cases.append(run_case(spec, rng.uniform(0, 1, size=(1, spec["n_params"])).tolist(), "c128"))
    # tiled path with non-trivial measurements / every gate kind at n=15
    spec = W.random_circuit(15, 60, seed=77, meas=[["expval", [["PauliZ", [3]]]], ["expval", [["PauliX", [14]]]],
                                                  ["expval", [["PauliZ", [0]], ["PauliZ", [9]]]]])
    cases.append(run_case(spec, rng.uniform(-3, 3, size=(1, spec["n_params"])).tolist()))
    spec = W.random_circuit(15, 40, seed=78, meas=[["probs", [14, 2, 7]]])
    cases.append(run_case(spec, rng.uniform(-3, 3, size=(1, spec["n_params"])).tolist(), "c128"))

    with open(os.path.join(HERE, "sv_cases.json"), "w") as fh:
        json.dump(cases, fh)
    print("sv cases:", len(cases))

    # 4. gate matrices (test_pytorch_backend.py:138-360 compares the same set)
    mats = {}
    with qai.QInterpreter.circuits.storage_base.CircuitStorage():
        for name in W_GATES:
            npar = {"RX": 1, "RY": 1, "RZ": 1, "Rot": 3, "PhaseShift": 1, "ControlledPhaseShift": 1, "CRX": 1, "CRY": 1,
                    "CRZ": 1}.get(name, 0)
            nq = 3 if name in ("CSWAP", "Toffoli") else (2 if name in ("CNOT", "CZ", "CY", "SWAP", "ControlledPhaseShift",
                                                                        "CRX", "CRY", "CRZ") else 1)
            pars = [0.3, 0.4, 0.5][:npar] if npar == 3 else [0.5] * npar
            op = getattr(qai, name)(*pars, qubits=list(range(nq)), do_queue=False)
            m = np.asarray(op.matrix, dtype=complex)
            mats[name] = {"params": pars, "re": m.real.tolist(), "im": m.imag.tolist()}
    with open(os.path.join(HERE, "gate_matrices.json"), "w") as fh:
        json.dump(mats, fh)

    # 5. tensor-network index maps (gen_tensor_networks, tensor_network.py:850-1099)
    tn = []
    tn_specs = [
        {"name": "tn_a", "num_qubits": 2, "n_params": 2, "gates": [["RY", [0], ["p0"]], ["RZ", [1], ["p1"]]],
         "meas": [["expval", [["PauliZ", [0]]]], ["state"]]},
        {"name": "tn_b", "num_qubits": 2, "n_params": 1,
         "gates": [["Hadamard", [0], []], ["CNOT", [0, 1], []], ["RX", [1], ["p0"]]],
         "meas": [["expval", [["PauliZ", [1]]]]]},
        W.random_circuit(4, 20, seed=3, meas=[["expval", [["PauliZ", [0]], ["PauliX", [3]]]], ["probs", [2, 1]],
                                               ["probs", None], ["state"]]),
        W.qnn4(), W.mbl_1d(12), W.hea(6, 2), W.lattice_rcs(3, 3, 4, seed=1),
    ]
    for spec in tn_specs:
        tn.append({"spec": spec, "networks": tn_maps(spec)})
    with open(os.path.join(HERE, "tn_index_maps.json"), "w") as fh:
        json.dump(tn, fh, ensure_ascii=True)
    print("tn maps:", len(tn))

    # 6. post-measurement states: probs(qubits, after_state=True), pytorch_backend.py:474-493 (the reference's
    #    probs() helper drops the flag, measurement.py:206, so it is set on the returned object)
    def after_def(a, b, c):
        qai.RX(a, qubits=[0]); qai.RY(b, qubits=[1]); qai.CNOT(qubits=[0, 2]); qai.RZ(c, qubits=[2])
        qai.Hadamard(qubits=[1]); qai.CNOT(qubits=[1, 2])
        m = qai.measurement.probs(qubits=[0, 2])
        m.after_state = True
        return m
    pa = [torch.tensor(0.54), torch.tensor(0.12), torch.tensor(-0.8)]
    cc = qai.Circuit(after_def, 3, *pa).compilecircuit(backend="pytorch")
    y = cc(*pa)
    states = {k: [[float(x.real), float(x.imag)] for x in torch.stack(v).reshape(-1)]
              for k, v in cc.states_after_measurement.items()}
    with open(os.path.join(HERE, "after_state.json"), "w") as fh:
        json.dump({"params": [0.54, 0.12, -0.8], "probs": y.detach().numpy().tolist(), "states": states}, fh)


W_GATES = ["I", "Hadamard", "PauliX", "PauliY", "PauliZ", "S", "T", "SX", "CNOT", "CZ", "CY", "SWAP", "CSWAP", "Toffoli",
           "RX", "RY", "RZ", "Rot", "PhaseShift", "ControlledPhaseShift", "CRX", "CRY", "CRZ"]

if __name__ == "__main__":
    main()
