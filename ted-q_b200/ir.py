"""Circuit -> static IR (host side of the drop-in boundary).

What the reference keeps in per-gate OrderedDicts and rebuilds on every call
(compiled_circuit.py:82-89, :418-440, :492-547) is resolved ONCE here:

* gate table: kind, qubits, and per parameter slot either an index into the flat
  parameter vector (trainable slot; positional binding of compiled_circuit.py:522-547)
  or the trace-time constant; gates without trainable slots carry their trace-time
  matrix (compiled_circuit.py:432-435, pytorch_backend.py:567-577);
* state-vector plan integers ``_axeslist`` / ``_permutationlist`` — bit-exact with
  compiled_circuit.py:126-202 (pinned by test_compiled_circuit.py:88-104);
* measurement table (pytorch_backend.py:393-498).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# must match enum tq_gate_kind in include/tedq_b200.h
GATE_KIND = {
    "FIXED": 0, "RX": 1, "RY": 2, "RZ": 3, "Rot": 4, "PhaseShift": 5,
    "ControlledPhaseShift": 6, "CRX": 7, "CRY": 8, "CRZ": 9,
}
MEAS_EXPVAL, MEAS_PROBS, MEAS_STATE = 0, 1, 2
MF_ZSTRING = 1


def _scalar(p) -> float:
    if hasattr(p, "detach"):
        p = p.detach().cpu().numpy()
    return float(np.asarray(p).reshape(-1)[0])


@dataclass
class GateRec:
    name: str
    kind: int
    qubits: Tuple[int, ...]
    param_idx: Tuple[int, ...] = ()        # flat index per parameter slot, -1 = constant
    param_const: Tuple[float, ...] = ()
    matrix: Optional[np.ndarray] = None    # FIXED only, complex128 (2^k, 2^k)


@dataclass
class MeasRec:
    kind: int
    flags: int = 0
    qubits: Tuple[int, ...] = ()
    matrix: Optional[np.ndarray] = None    # dense observable
    shape: Tuple[int, ...] = ()            # shape of this measurement's result (one parameter set)
    is_complex: bool = False
    after_state: bool = False


@dataclass
class CircuitIR:
    num_qubits: int
    gates: List[GateRec]
    meas: List[MeasRec]
    n_params: int
    init_state: Optional[np.ndarray] = None
    axeslist: list = field(default_factory=list)         # reversed gate order, as the reference stores it
    permutationlist: list = field(default_factory=list)


def sv_axes_perm(num_qubits: int, qubits: Sequence[int]):
    """(gate_pos, state_pos) and permutation of one gate (compiled_circuit.py:126-198).

    tensordot(gate, state, (gate_in_axes, qubits)) leaves the gate's output axes
    first, then the untouched state axes in ascending order; ``perm[q]`` is where
    qubit q's axis sits in that result.
    """
    k = len(qubits)
    gate_pos = [1] if k == 1 else list(range(k, 2 * k))
    where = {q: i for i, q in enumerate(qubits)}
    nxt = k
    perm = []
    for q in range(num_qubits):
        if q in where:
            perm.append(where[q])
        else:
            perm.append(nxt)
            nxt += 1
    return (gate_pos, list(qubits)), perm


def _kron_obs(obs_list, num_qubits):
    """Product of a list of 1-qubit observables, applied in list order (pytorch_backend.py:402-424):
    <psi| O_last ... O_first |psi>.  The reference contracts axis 1 of each observable with one qubit,
    so list entries are 1-qubit operators."""
    qs = []
    for ob in obs_list:
        if len(ob.qubits) != 1:
            raise ValueError("a list observable must be made of 1-qubit observables")
        if ob.qubits[0] not in qs:
            qs.append(int(ob.qubits[0]))
    k = len(qs)
    full = np.eye(2 ** k, dtype=complex)
    for ob in obs_list:
        pos = qs.index(int(ob.qubits[0]))
        m = np.asarray(ob.matrix, dtype=complex).reshape(2, 2)
        emb = np.kron(np.kron(np.eye(2 ** pos), m), np.eye(2 ** (k - pos - 1)))
        full = emb @ full
    return tuple(qs), full


def _is_z_string(obs_list) -> bool:
    seen = set()
    for ob in obs_list:
        if len(ob.qubits) != 1 or ob.qubits[0] in seen:
            return False
        seen.add(ob.qubits[0])
        m = np.asarray(ob.matrix, dtype=complex)
        if m.shape != (2, 2) or not np.array_equal(m, np.array([[1, 0], [0, -1]], dtype=complex)):
            return False
    return True


def build_ir(circuit) -> CircuitIR:
    n = int(circuit.num_qubits)
    gates: List[GateRec] = []
    axes, perms = [], []
    count = 0
    for op in circuit.operators:
        qubits = tuple(int(q) for q in op.qubits)
        a, p = sv_axes_perm(n, list(qubits))
        axes.append(a)
        perms.append(p)
        trainable = list(op.trainable_params)
        if trainable:
            if op.name not in GATE_KIND:
                raise ValueError(f"{op.name}: gate has trainable parameters but no device formula")
            nslots = len(op.parameters)
            idx = [-1] * nslots
            const = [0.0] * nslots
            for i, pos in enumerate(trainable):   # i-th trainable slot <- flat[count + i]
                idx[pos] = count + i
            for pos in range(nslots):
                if idx[pos] < 0:
                    const[pos] = _scalar(op.parameters[pos])
            count += len(trainable)
            gates.append(GateRec(op.name, GATE_KIND[op.name], qubits, tuple(idx), tuple(const)))
        else:
            k = len(qubits)
            m = np.asarray(op.matrix, dtype=np.complex128).reshape(2 ** k, 2 ** k)
            gates.append(GateRec(op.name, GATE_KIND["FIXED"], qubits, matrix=m))
    axes.reverse()
    perms.reverse()

    meas: List[MeasRec] = []
    for ms in circuit.measurements:
        rt = getattr(ms.return_type, "value", ms.return_type)
        if rt == "expval":
            obs = ms.obs if isinstance(ms.obs, list) else [ms.obs]
            if _is_z_string(obs):
                meas.append(MeasRec(MEAS_EXPVAL, MF_ZSTRING, tuple(int(o.qubits[0]) for o in obs), shape=()))
            else:
                if isinstance(ms.obs, list):
                    qs, mat = _kron_obs(obs, n)
                else:
                    qs = tuple(int(q) for q in ms.obs.qubits)
                    mat = np.asarray(ms.obs.matrix, dtype=np.complex128).reshape(2 ** len(qs), 2 ** len(qs))
                meas.append(MeasRec(MEAS_EXPVAL, 0, qs, matrix=mat, shape=()))
        elif rt == "probs":
            if ms.qubits is None:
                meas.append(MeasRec(MEAS_PROBS, 0, (), shape=(2,) * n, after_state=bool(ms.after_state)))
            else:
                qs = tuple(int(q) for q in ms.qubits)
                meas.append(MeasRec(MEAS_PROBS, 0, qs, shape=(2,) * len(set(qs)), after_state=bool(ms.after_state)))
        elif rt == "state":
            meas.append(MeasRec(MEAS_STATE, 0, (), shape=(2,) * n, is_complex=True))
        else:
            raise NotImplementedError(f"measurement type {rt!r} is not implemented (measurement.py:158-171)")

    init = None
    if circuit.init_state is not None and circuit.init_state:
        init = np.asarray(circuit.init_state.matrix, dtype=np.complex128).reshape(-1)
        if init.size != 2 ** n:
            raise ValueError("initial state has the wrong number of amplitudes")
    return CircuitIR(n, gates, meas, count, init, axes, perms)
