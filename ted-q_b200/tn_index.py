"""Tensor-network index maps of a circuit — bit-exact mirror of the reference's
``gen_tensor_networks`` (tedq/tensor_network/tensor_network.py:850-1099) and ``get_symbol`` (:1109-1127).

One network per measurement.  Wire q starts on index id q.  Walking the gates in circuit order, a
k-qubit gate takes k fresh ids ``cur+1 .. cur+k``; its tensor is indexed ``[new_1..new_k, old_1..old_k]``
(output legs first — the layout of ``matrix.reshape([2]*2k)``) and every touched wire moves to its new id.
For ``expval`` the observable tensor(s) follow, for ``expval``/``probs`` the whole gate list is walked again
in REVERSE order with fresh ids (operands there are the adjoint gates, pytorch_backend.py:524-546), then
one closing cap per wire in wire order.  ``state`` has no adjoint half and leaves the final wire ids open.

Index ids are integers; ``symbol(i)`` is the reference's character for id i, so
``[[symbol(i) for i in t] for t in net.inputs]`` equals the reference's ``input_indices``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

_BASE = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"

# operand kinds (what array sits at each position, pytorch_backend.py:311-336)
OPD_CAP, OPD_GATE, OPD_OBS, OPD_ADJ = 0, 1, 2, 3


def symbol(i: int) -> str:
    return _BASE[i] if i < 52 else chr(i + 140)


@dataclass
class Network:
    inputs: List[List[int]]                 # index ids per tensor, slow -> fast (C order)
    output: List[int]
    operands: List[Tuple[int, int]] = field(default_factory=list)  # (OPD_*, gate index | obs index | wire)
    # tn_simplify.py only: per operand, the entries of the full tensor it keeps (None = the whole tensor)
    reductions: List = field(default_factory=list)

    def symbols(self):
        return [[symbol(i) for i in t] for t in self.inputs], [symbol(i) for i in self.output]

    def size_keys(self):
        """Key order of the reference's size_dict (tensor_network.py:1080-1086)."""
        seen = []
        s = set()
        for t in self.inputs:
            for i in t:
                if i not in s:
                    s.add(i)
                    seen.append(symbol(i))
        for i in self.output:
            if i not in s:
                s.add(i)
                seen.append(symbol(i))
        return seen


def _thread(wire_id: List[int], cur: int, qubits: Sequence[int]):
    """Indices of one k-qubit tensor; returns (indices, new cur)."""
    k = len(qubits)
    new = [cur + 1 + j for j in range(k)]
    idx = new + [wire_id[q] for q in qubits]
    for q, i in zip(qubits, new):
        wire_id[q] = i
    return idx, cur + k


def light_cone(gate_qubits: Sequence[Sequence[int]], qubits: Sequence[int]) -> List[int]:
    """Gates (ids, ascending) inside the causal cone of ``qubits`` at the end of the circuit: walking the gate list
    backwards, a gate belongs to the cone when it touches a qubit the cone has reached.  Every other gate meets its
    own adjoint in <0|U^dagger O U|0> (or under the partial trace of a marginal) and cancels: U^dagger U = 1."""
    active = set(int(q) for q in qubits)
    cone = []
    for gi in range(len(gate_qubits) - 1, -1, -1):
        qs = set(int(q) for q in gate_qubits[gi])
        if active & qs:
            cone.append(gi)
            active |= qs
    return cone[::-1]


def cone_of_measurement(gate_qubits, kind, payload):
    """Gate ids a measurement's network needs under light-cone pruning, or None (no pruning: state(), probs())."""
    if kind == "expval":
        return light_cone(gate_qubits, [q for qs in payload for q in qs])
    if kind == "probs" and payload is not None:
        return light_cone(gate_qubits, payload)
    return None


def cone_qubits(gate_qubits, cone, kind, payload) -> List[int]:
    """Qubits a pruned network keeps (ascending): those the cone's gates touch plus the measured ones.  Every other
    wire is an opening |0> cap contracted with its own closing cap, <0|0> = 1: both caps are dropped."""
    qs = {int(q) for gi in cone for q in gate_qubits[gi]}
    if kind == "expval":
        qs |= {int(q) for ob in payload for q in ob}
    elif payload is not None:
        qs |= {int(q) for q in payload}
    return sorted(qs)


def index_maps(num_qubits: int, gate_qubits: Sequence[Sequence[int]], measurements,
               prune_light_cone: bool = False) -> List[Network]:
    """``measurements``: sequence of (kind, payload) with kind in {"expval", "probs", "state"};
    payload = list of observable qubit-lists (expval: one entry per observable of a list observable,
    or a single multi-qubit entry), kept-qubit list or None (probs), None (state).
    ``prune_light_cone`` (not in the reference, whose networks always hold every gate twice): expval / marginal
    networks keep only the gates inside the measurement's causal cone; operand refs stay the circuit's gate ids."""
    n = num_qubits
    wire0 = list(range(n))
    cur0 = n - 1
    base_inputs = [[q] for q in range(n)]
    base_ops = [(OPD_CAP, q) for q in range(n)]
    for gi, qs in enumerate(gate_qubits):
        idx, cur0 = _thread(wire0, cur0, list(qs))
        base_inputs.append(idx)
        base_ops.append((OPD_GATE, gi))

    nets = []
    for kind, payload in measurements:
        wire = list(wire0)
        cur = cur0
        inputs = [list(t) for t in base_inputs]
        ops = list(base_ops)
        gates = list(range(len(gate_qubits)))
        cone = cone_of_measurement(gate_qubits, kind, payload) if prune_light_cone else None
        kept_q = list(range(n))
        if cone is not None and len(cone) < len(gates):
            gates = cone
            kept_q = cone_qubits(gate_qubits, cone, kind, payload)
            wire, cur = list(range(n)), n - 1
            inputs, ops = [[q] for q in kept_q], [(OPD_CAP, q) for q in kept_q]
            for gi in gates:
                idx, cur = _thread(wire, cur, list(gate_qubits[gi]))
                inputs.append(idx)
                ops.append((OPD_GATE, gi))
        output: List[int] = []
        if kind == "state":
            output = [wire[q] for q in range(n)]
            nets.append(Network(inputs, output, ops))
            continue
        if kind == "expval":
            for oi, qs in enumerate(payload):
                idx, cur = _thread(wire, cur, list(qs))
                inputs.append(idx)
                ops.append((OPD_OBS, oi))
        elif kind == "probs":
            if payload is not None:
                output = [wire[q] for q in payload]
        else:
            raise ValueError(kind)
        for gi in reversed(gates):
            qs = list(gate_qubits[gi])
            if len(qs) > 3:
                raise ValueError("Error!! unknown operator with len of applied qubits larger than 3!")
            idx, cur = _thread(wire, cur, qs)
            inputs.append(idx)
            ops.append((OPD_ADJ, gi))
        for q in kept_q:
            inputs.append([wire[q]])
            ops.append((OPD_CAP, q))
        nets.append(Network(inputs, output, ops))
    return nets


def networks_of_circuit(circuit, prune_light_cone: bool = False) -> List[Network]:
    meas = []
    for ms in circuit.measurements:
        rt = getattr(ms.return_type, "value", ms.return_type)
        if rt == "expval":
            obs = ms.obs if isinstance(ms.obs, list) else [ms.obs]
            meas.append(("expval", [list(o.qubits) for o in obs]))
        elif rt == "probs":
            meas.append(("probs", None if ms.qubits is None else list(ms.qubits)))
        elif rt == "state":
            meas.append(("state", None))
        else:
            raise NotImplementedError(rt)
    return index_maps(circuit.num_qubits, [list(op.qubits) for op in circuit.operators], meas, prune_light_cone)
