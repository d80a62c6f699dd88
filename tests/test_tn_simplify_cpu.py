"""CPU: the structure-aware network builder (tedq_b200/tn_simplify.py) against the oracle.

* every gate class the table calls diagonal / controlled really vanishes outside the kept entries, for the
  reference's gate tensors (oracle/sv_ref.gate_tensor restates pytorch_backend.py:579-1188) at random angles;
* the simplified network, contracted pairwise in numpy with the reduced operands, gives the same expval / probs /
  state as the reference-exact network of the same circuit."""
import numpy as np
import pytest
import torch

import tedq_b200 as qb
from oracle import sv_ref, tn_ref
from tedq_b200 import planner, tn_simplify, workloads as W

GATES = {"I": 1, "PauliZ": 1, "S": 1, "T": 1, "RZ": 1, "PhaseShift": 1, "CZ": 2, "ControlledPhaseShift": 2, "CRZ": 2,
         "CNOT": 2, "CY": 2, "CRX": 2, "CRY": 2, "Toffoli": 3, "CSWAP": 3}
N_PAR = {"RZ": 1, "PhaseShift": 1, "ControlledPhaseShift": 1, "CRZ": 1, "CRX": 1, "CRY": 1}


@pytest.mark.parametrize("name", sorted(GATES))
def test_structure_table_matches_reference_gate_tensors(name):
    k = GATES[name]
    spec = W._Builder("one", k)
    params = [0.37 + 0.9 * j for j in range(N_PAR.get(name, 0))]
    spec.g(name, list(range(k)), *params)
    spec.state()
    circ = W.build_circuit(spec.spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    op = circ.operators[0]
    st = tn_simplify.verified_structure(op.name, k, op.matrix)
    assert st is not None and st == tn_simplify.structure(name, k)
    full = np.asarray(op.matrix, dtype=np.complex128).reshape(-1)
    keep = np.zeros(full.shape, dtype=bool)
    keep[list(st[2])] = True
    assert np.all(full[~keep] == 0)
    assert len(st[2]) == 2 ** (st[0] + 2 * st[1])


def test_dense_gates_are_left_alone():
    for name, k in (("Hadamard", 1), ("RX", 1), ("RY", 1), ("Rot", 1), ("SX", 1), ("PauliX", 1), ("SWAP", 2)):
        assert tn_simplify.structure(name, k) is None


@pytest.mark.parametrize("seed", range(6))
def test_simplified_network_equals_reference_network(seed):
    meas = [["expval", [["PauliZ", [1]]]], ["probs", [0, 3]], ["state"], ["expval", [["PauliX", [2]]]]]
    spec = W.random_circuit(5, 45, seed=seed, meas=meas)
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    flat = torch.rand(spec["n_params"], dtype=torch.float64)
    nets = tn_simplify.networks_of_circuit(circ)
    ref = tn_ref.run_tn(circ, flat, torch.complex128)
    arrays_all = tn_ref.operands(circ, flat, torch.complex128)
    n_reduced = 0
    for net, arrs, r in zip(nets, arrays_all, ref):
        ops = []
        for a, red in zip(arrs, net.reductions):
            a = np.asarray(a)
            if red is not None:
                n_reduced += 1
                a = a.reshape(-1)[list(red)].reshape((2,) * int(np.log2(len(red))))
            ops.append(a)
        assert [len(t) for t in net.inputs] == [o.ndim for o in ops]
        info = planner.find_path(net.inputs, net.output, repeats=2, seed=seed)
        got = np.asarray(tn_ref.contract_path(ops, net.inputs, net.output, info.path))
        want = np.asarray(r)
        got = got if np.iscomplexobj(want) else np.real(got)
        assert np.abs(got - want).max() < 1e-12
    assert n_reduced > 0


def test_simplification_cuts_the_cost_of_the_lattice_circuit():
    spec = W.lattice_rcs(4, 5, 8, seed=2)
    circ = W.build_circuit(spec, qb)
    gq = [list(op.qubits) for op in circ.operators]
    from tedq_b200 import tn_index
    dense = tn_index.index_maps(20, gq, [("state", None)])[0]
    simp = tn_simplify.index_maps(20, gq, tn_simplify.gate_structures(circ), [("state", None)])[0]
    cap = lambda net: [list(t) for t in net.inputs] + [[ix] for ix in net.output]
    fd = planner.find_path(cap(dense), [], repeats=4).flops_log2
    fs = planner.find_path(cap(simp), [], repeats=4).flops_log2
    assert fs < fd - 2, (fs, fd)


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("simplify", [False, True], ids=["dense", "structured"])
def test_light_cone_pruned_networks_equal_the_full_ones(seed, simplify):
    """hyper_opt["light_cone"]: networks of expval / marginal measurements without the gates outside the causal
    cone contract (oracle, numpy) to the same numbers as the reference-exact networks; state() and probs() are
    left whole; the cone really removes gates when the observable sits on few qubits of a shallow circuit."""
    from tedq_b200 import tn_index

    meas = [["expval", [["PauliZ", [1]]]], ["probs", [0, 3]], ["state"], ["expval", [["PauliX", [2]], ["PauliZ", [4]]]],
            ["probs", None]]
    spec = W.random_circuit(6, 14, seed=seed, meas=meas)
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    flat = torch.rand(spec["n_params"], dtype=torch.float64)
    mod = tn_simplify if simplify else tn_index
    full = mod.networks_of_circuit(circ)
    pruned = mod.networks_of_circuit(circ, prune_light_cone=True)
    ref = tn_ref.run_tn(circ, flat, torch.complex128)
    arrays_all = tn_ref.operands(circ, flat, torch.complex128)
    fewer = 0
    for m, (net_f, net_p, arrs, r) in enumerate(zip(full, pruned, arrays_all, ref)):
        if meas[m][0] == "state" or (meas[m][0] == "probs" and meas[m][1] is None):
            assert net_p.inputs == net_f.inputs and net_p.operands == net_f.operands
            continue
        lookup = {}
        red_f = getattr(net_f, "reductions", None) or [None] * len(net_f.operands)
        for key, a, red in zip(net_f.operands, arrs, red_f):
            a = np.asarray(a)
            if red is not None:
                a = a.reshape(-1)[list(red)].reshape((2,) * int(np.log2(len(red))))
            lookup[key] = a
        ops = [lookup[key] for key in net_p.operands]
        assert [len(t) for t in net_p.inputs] == [o.ndim for o in ops]
        assert len(net_p.inputs) <= len(net_f.inputs)
        fewer += len(net_f.inputs) - len(net_p.inputs)
        gates_f = sorted(ref_ for kind, ref_ in net_f.operands if kind == tn_index.OPD_GATE)
        gates_p = sorted(ref_ for kind, ref_ in net_p.operands if kind == tn_index.OPD_GATE)
        adj_p = sorted(ref_ for kind, ref_ in net_p.operands if kind == tn_index.OPD_ADJ)
        assert gates_p == adj_p and set(gates_p) <= set(gates_f)
        info = planner.find_path(net_p.inputs, net_p.output, repeats=2, seed=seed)
        got = np.asarray(tn_ref.contract_path(ops, net_p.inputs, net_p.output, info.path))
        want = np.asarray(r)
        got = got if np.iscomplexobj(want) else np.real(got)
        assert np.abs(got - want).max() < 1e-12, (m, np.abs(got - want).max())
    assert fewer > 0


def test_light_cone_of_a_brick_circuit():
    from tedq_b200 import tn_index

    gq = [[0, 1], [2, 3], [4, 5], [1, 2], [3, 4], [0], [5]]
    # gate 4 = [3, 4] acts after [4, 5] and nothing later links it to qubit 5: it cancels against its adjoint
    assert tn_index.light_cone(gq, [5]) == [2, 6]
    assert tn_index.light_cone(gq, [0]) == [0, 5]
    assert tn_index.light_cone(gq, [0, 5]) == [0, 2, 5, 6]
    assert tn_index.light_cone(gq, [3]) == [1, 2, 4]
    assert tn_index.light_cone(gq, []) == []
