"""Structure-aware ("simplified") tensor networks of a circuit: ``tn_simplify=True``.

The reference builds every k-qubit gate as a dense rank-2k tensor with k fresh indices
(tedq/tensor_network/tensor_network.py:871-919) and leaves the reduction of the network to a simplifier
(``TensorNetwork.simplify``, tensor_network.py:90-129) that does not work (tensor_network.py:94).  This module
is the working counterpart for the two structures that dominate circuit networks:

* a DIAGONAL gate (RZ, PhaseShift, Z, S, T, I; CZ, CRZ, ControlledPhaseShift) does not change the wire index:
  it becomes a rank-k tensor ``d[w_1..w_k]`` on the CURRENT wire indices (a hyper-index shared by more than two
  tensors) instead of a rank-2k tensor with k new indices;
* a CONTROLLED gate (CNOT, CY, CRX, CRY, Toffoli, CSWAP) does not change its control wires: only the target
  wires get fresh indices, the tensor is ``B[c.., out_t.., in_t..]``.

The reduced tensors are sub-sets of the entries of the full gate tensors (``offsets`` below index the C-ordered
``[out..., in...]`` tensor), so the operands are gathered from the same device buffers the unsimplified path
uses; the adjoint half keeps the same structure (the adjoint of a diagonal / controlled gate is diagonal /
controlled on the same wires).  On the 40-qubit lattice circuit of BASELINE config 5 the contraction cost drops
from 2^44 to 2^31 flops.  The index maps of THIS module are ours; the reference-exact maps stay in tn_index.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .tn_index import OPD_ADJ, OPD_CAP, OPD_GATE, OPD_OBS, Network, cone_of_measurement, cone_qubits

DIAG1 = {"I", "PauliZ", "S", "T", "RZ", "PhaseShift"}
DIAG2 = {"CZ", "ControlledPhaseShift", "CRZ"}
CTRL1 = {"CNOT", "CY", "CRX", "CRY"}


def structure(name: str, n_qubits: int):
    """-> (n_control_or_diag_wires, n_target_wires, offsets) or None for a dense gate.
    Reduced tensor layout: [shared wires..., target outs..., target ins...] (C order)."""
    if name in DIAG1 and n_qubits == 1:
        return 1, 0, (0, 3)
    if name in DIAG2 and n_qubits == 2:
        return 2, 0, tuple(a * 10 + b * 5 for a in (0, 1) for b in (0, 1))
    if name in CTRL1 and n_qubits == 2:
        return 1, 1, tuple(c * 10 + o * 4 + i for c in (0, 1) for o in (0, 1) for i in (0, 1))
    if name == "Toffoli" and n_qubits == 3:
        return 2, 1, tuple(c0 * 36 + c1 * 18 + o * 8 + i for c0 in (0, 1) for c1 in (0, 1) for o in (0, 1)
                           for i in (0, 1))
    if name == "CSWAP" and n_qubits == 3:
        return 1, 2, tuple(c * 36 + o1 * 16 + o2 * 8 + i1 * 2 + i2 for c in (0, 1) for o1 in (0, 1)
                           for o2 in (0, 1) for i1 in (0, 1) for i2 in (0, 1))
    return None


def verified_structure(name: str, n_qubits: int, matrix) -> Optional[tuple]:
    """``structure`` only if the gate's trace-time matrix really vanishes outside the kept entries (a gate class
    whose convention differs from the table falls back to the dense tensor instead of giving wrong numbers)."""
    st = structure(name, n_qubits)
    if st is None or matrix is None:
        return st
    full = np.asarray(matrix, dtype=np.complex128).reshape(-1)
    mask = np.ones(full.shape, dtype=bool)
    mask[list(st[2])] = False
    return st if np.all(np.abs(full[mask]) < 1e-14) else None


def _thread(wire: List[int], cur: int, qubits: Sequence[int], st):
    """Indices of one tensor; dense: [new..., old...]; structured: [shared..., new_t..., old_t...]."""
    if st is None:
        k = len(qubits)
        new = [cur + 1 + j for j in range(k)]
        idx = new + [wire[q] for q in qubits]
        for q, i in zip(qubits, new):
            wire[q] = i
        return idx, cur + k
    n_sh, n_t, _ = st
    shared, targets = list(qubits[:n_sh]), list(qubits[n_sh:n_sh + n_t])
    new = [cur + 1 + j for j in range(n_t)]
    idx = [wire[q] for q in shared] + new + [wire[q] for q in targets]
    for q, i in zip(targets, new):
        wire[q] = i
    return idx, cur + n_t


def index_maps(num_qubits: int, gate_qubits, gate_structs, measurements, obs_structs=None,
               prune_light_cone: bool = False) -> List[Network]:
    """Same contract as tn_index.index_maps plus ``gate_structs[g]`` (``structure`` result or None) and, per
    expval measurement, ``obs_structs[m][j]``.  Network.reductions[t] = entry offsets of operand t or None."""
    n = num_qubits
    wire0 = list(range(n))
    cur0 = n - 1
    base_inputs = [[q] for q in range(n)]
    base_ops = [(OPD_CAP, q) for q in range(n)]
    base_red: List[Optional[tuple]] = [None] * n
    for gi, qs in enumerate(gate_qubits):
        idx, cur0 = _thread(wire0, cur0, list(qs), gate_structs[gi])
        base_inputs.append(idx)
        base_ops.append((OPD_GATE, gi))
        base_red.append(gate_structs[gi][2] if gate_structs[gi] else None)
    nets = []
    for mi, (kind, payload) in enumerate(measurements):
        wire, cur = list(wire0), cur0
        inputs, ops, red = [list(t) for t in base_inputs], list(base_ops), list(base_red)
        gates = list(range(len(gate_qubits)))
        cone = cone_of_measurement(gate_qubits, kind, payload) if prune_light_cone else None
        kept_q = list(range(n))
        if cone is not None and len(cone) < len(gates):      # tn_index.light_cone: the other gates cancel
            gates = cone
            kept_q = cone_qubits(gate_qubits, cone, kind, payload)
            wire, cur = list(range(n)), n - 1
            inputs, ops = [[q] for q in kept_q], [(OPD_CAP, q) for q in kept_q]
            red = [None] * len(kept_q)
            for gi in gates:
                idx, cur = _thread(wire, cur, list(gate_qubits[gi]), gate_structs[gi])
                inputs.append(idx)
                ops.append((OPD_GATE, gi))
                red.append(gate_structs[gi][2] if gate_structs[gi] else None)
        output: List[int] = []
        if kind == "state":
            net = Network(inputs, [wire[q] for q in range(n)], ops)
            net.reductions = red
            nets.append(net)
            continue
        if kind == "expval":
            for oi, qs in enumerate(payload):
                st = obs_structs[mi][oi] if obs_structs else None
                idx, cur = _thread(wire, cur, list(qs), st)
                inputs.append(idx)
                ops.append((OPD_OBS, oi))
                red.append(st[2] if st else None)
        elif kind == "probs":
            if payload is not None:
                output = [wire[q] for q in payload]
        else:
            raise ValueError(kind)
        for gi in reversed(gates):
            qs = list(gate_qubits[gi])
            if len(qs) > 3:
                raise ValueError("Error!! unknown operator with len of applied qubits larger than 3!")
            idx, cur = _thread(wire, cur, qs, gate_structs[gi])
            inputs.append(idx)
            ops.append((OPD_ADJ, gi))
            red.append(gate_structs[gi][2] if gate_structs[gi] else None)
        for q in kept_q:
            inputs.append([wire[q]])
            ops.append((OPD_CAP, q))
            red.append(None)
        net = Network(inputs, output, ops)
        net.reductions = red
        nets.append(net)
    return nets


def gate_structures(circuit):
    return [verified_structure(op.name, len(op.qubits), getattr(op, "matrix", None)) for op in circuit.operators]


def networks_of_circuit(circuit, prune_light_cone: bool = False) -> List[Network]:
    meas, obs_structs = [], []
    for ms in circuit.measurements:
        rt = getattr(ms.return_type, "value", ms.return_type)
        if rt == "expval":
            obs = ms.obs if isinstance(ms.obs, list) else [ms.obs]
            meas.append(("expval", [list(o.qubits) for o in obs]))
            obs_structs.append([verified_structure(o.name, len(o.qubits), getattr(o, "matrix", None)) for o in obs])
        elif rt == "probs":
            meas.append(("probs", None if ms.qubits is None else list(ms.qubits)))
            obs_structs.append(None)
        elif rt == "state":
            meas.append(("state", None))
            obs_structs.append(None)
        else:
            raise NotImplementedError(rt)
    return index_maps(circuit.num_qubits, [list(op.qubits) for op in circuit.operators], gate_structures(circuit),
                      meas, obs_structs, prune_light_cone)
