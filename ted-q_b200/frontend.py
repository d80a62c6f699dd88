"""Stand-alone circuit tracer with TeD-Q's user-facing names.

The B200 backend is a drop-in for ``qai.Circuit(...).compilecircuit(backend=...)``
of the reference (tedq/QInterpreter/circuits/circuit.py:42-98, :268-285) and consumes
the reference's own ``Circuit`` objects when ``tedq`` is importable (see
``register.py``).  The reference is pure Python and is NOT present on the GPU
box, so tests, ``smoke()`` and ``bench.py`` need a front end that travels with the
repo.  This module is that front end: the same call shapes

    def circuit_def(*params):
        RX(params[0], qubits=[0]); CNOT(qubits=[0, 1])
        return expval(PauliZ(qubits=[0]))
    cc = Circuit(circuit_def, 2, a, b).compilecircuit(backend="pytorch_b200")

producing objects with exactly the attributes the backend reads from a reference
circuit (compiled_circuit.py:71-93): ``num_qubits``, ``operators`` (``instance_id
name qubits parameters trainable_params matrix num_qubits``), ``measurements``
(``return_type obs qubits after_state``) and ``init_state`` (``matrix``).

It is deliberately small: one table of gate definitions instead of the
reference's 24 classes (tedq/QInterpreter/operators/qubit.py), no drawing, no
decompositions, no hardware back ends.  Gate matrices follow qubit.py
(RX :990-997, RY :1048-1055, RZ :1106-1112, Rot :1169-1190, PhaseShift :1253-1257,
ControlledPhaseShift :1327-1334, CRX :1429-1436, CRY :1531-1538, CRZ :1631-1642)
and are checked against matrices dumped from the reference in
tests/golden/gate_matrices.json.
"""
from __future__ import annotations

import cmath
import enum
import itertools
import math
from typing import Callable, List, Optional, Sequence

import numpy as np

__all__ = [
    "Circuit", "expval", "probs", "state", "var", "sample", "InitStateVector", "MeasurementReturnTypes",
    "Expectation", "Probability", "State", "Variance", "Sample", "GATE_NAMES", "HardwareEfficient", "Unitary",
]


class MeasurementReturnTypes(enum.Enum):
    """Same member names/values as measurement.py:240-255."""

    Sample = "sample"
    Variance = "var"
    Expectation = "expval"
    Probability = "probs"
    State = "state"


Expectation = MeasurementReturnTypes.Expectation
Probability = MeasurementReturnTypes.Probability
State = MeasurementReturnTypes.State
Variance = MeasurementReturnTypes.Variance
Sample = MeasurementReturnTypes.Sample

_ids = itertools.count()
_trace_stack: List[list] = []


def _scalar(p) -> float:
    """Trace-time numeric value of a parameter (python number, numpy or torch scalar)."""
    if hasattr(p, "detach"):
        p = p.detach().cpu().numpy()
    return float(np.asarray(p).reshape(-1)[0])


_S2 = 1.0 / math.sqrt(2.0)


def _m_rx(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]])


def _m_ry(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -s], [s, c]])


def _m_rz(t):
    p = cmath.exp(-0.5j * t)
    return np.array([[p, 0], [0, p.conjugate()]])


def _m_rot(a, b, w):
    c, s = math.cos(b / 2), math.sin(b / 2)
    return np.array([
        [cmath.exp(-0.5j * (a + w)) * c, -cmath.exp(0.5j * (a - w)) * s],
        [cmath.exp(-0.5j * (a - w)) * s, cmath.exp(0.5j * (a + w)) * c],
    ])


def _m_phase(p):
    return np.array([[1, 0], [0, cmath.exp(1j * p)]])


def _controlled(block):
    out = np.eye(4, dtype=complex)
    out[2:, 2:] = block
    return out


def _perm(n, mapping):
    m = np.zeros((n, n))
    for src, dst in enumerate(mapping):
        m[dst, src] = 1.0
    return m


# name -> (num_qubits, num_params, matrix function, is_observable)
_GATES = {
    "I": (1, 0, lambda: np.eye(2), True),
    "Hadamard": (1, 0, lambda: np.array([[_S2, _S2], [_S2, -_S2]]), True),
    "PauliX": (1, 0, lambda: np.array([[0, 1], [1, 0]]), True),
    "PauliY": (1, 0, lambda: np.array([[0, -1j], [1j, 0]]), True),
    "PauliZ": (1, 0, lambda: np.array([[1, 0], [0, -1]]), True),
    "S": (1, 0, lambda: np.array([[1, 0], [0, 1j]]), False),
    "T": (1, 0, lambda: np.array([[1, 0], [0, cmath.exp(1j * math.pi / 4)]]), False),
    "SX": (1, 0, lambda: 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]), False),
    "CNOT": (2, 0, lambda: _perm(4, [0, 1, 3, 2]), False),
    "CZ": (2, 0, lambda: np.diag([1, 1, 1, -1]), False),
    "CY": (2, 0, lambda: _controlled(np.array([[0, -1j], [1j, 0]])), False),
    "SWAP": (2, 0, lambda: _perm(4, [0, 2, 1, 3]), False),
    "CSWAP": (3, 0, lambda: _perm(8, [0, 1, 2, 3, 4, 6, 5, 7]), False),
    "Toffoli": (3, 0, lambda: _perm(8, [0, 1, 2, 3, 4, 5, 7, 6]), False),
    "RX": (1, 1, _m_rx, False),
    "RY": (1, 1, _m_ry, False),
    "RZ": (1, 1, _m_rz, False),
    "Rot": (1, 3, _m_rot, False),
    "PhaseShift": (1, 1, _m_phase, False),
    "ControlledPhaseShift": (2, 1, lambda p: np.diag([1, 1, 1, cmath.exp(1j * p)]), False),
    "CRX": (2, 1, lambda t: _controlled(_m_rx(t)), False),
    "CRY": (2, 1, lambda t: _controlled(_m_ry(t)), False),
    "CRZ": (2, 1, lambda t: _controlled(_m_rz(t)), False),
}
GATE_NAMES = tuple(_GATES)


class Operator:
    """One traced gate / observable (the attribute set of ops_abc.py:44-86)."""

    is_observable = False

    def __init__(self, name, params, qubits, do_queue=True, trainable_params=None, is_preparation=False,
                 matrix=None, is_observable=None):
        nq, npar, fn, obs = _GATES[name] if name in _GATES else (len(qubits), 0, None, bool(is_observable))
        if len(params) != npar:
            raise ValueError(f"{name}: # of parameters is not matched! expected {npar}parameters, but got {len(params)}.")
        if name in _GATES and len(qubits) != nq:
            raise ValueError(
                f"{name}: # of qubits this operator applied on is not matched! expected {nq} qubits, but got {len(qubits)}")
        self.name = name
        self.num_qubits = nq
        self.num_params = npar
        self.is_observable = obs
        self.instance_id = next(_ids)
        self.qubits = [int(q) for q in qubits]
        self.parameters = list(params)
        self.trainable_params = list(range(npar)) if trainable_params is None else list(trainable_params)
        self._is_preparation = is_preparation
        self.matrix = matrix if matrix is not None else fn(*[_scalar(p) for p in params])
        if do_queue:
            if not _trace_stack:
                raise ValueError("There's no global_deque for storing information!")
            _trace_stack[-1].append(self)

    def __repr__(self):
        return f"{self.name}(qubits={self.qubits})"


def _make_gate(name):
    def ctor(*params, qubits, do_queue=True, **kwargs):
        return Operator(name, params, qubits, do_queue=do_queue, trainable_params=kwargs.get("trainable_params"))

    ctor.__name__ = name
    ctor.__doc__ = f"{name} gate; see tedq/QInterpreter/operators/qubit.py."
    return ctor


for _n in _GATES:
    globals()[_n] = _make_gate(_n)
    __all__.append(_n)


def Unitary(matrix, qubits, do_queue=True, **kwargs):
    """User-defined gate / observable from its matrix (qubit.py:1696-1730; the reference's own constructor stops at
    an undefined name, :1719).  ``matrix``: (2^k, 2^k) or [2]*2k, first qubit = most significant bit.  As a gate
    k <= 3, as an observable (``expval`` / ``var`` / ``sample``) k <= 4 and the matrix must be Hermitian."""
    k = len(qubits)
    m = np.asarray(matrix, dtype=complex).reshape(2 ** k, 2 ** k)
    return Operator("Unitary", (), qubits, do_queue=do_queue, matrix=m, is_observable=True)


def InitStateVector(matrix, do_queue=True):
    """User-defined initial state (prepared_state.py: IintStateVector)."""
    return Operator("InitStateVector", (), [], do_queue=do_queue, is_preparation=True, matrix=np.asarray(matrix))


class QuantumMeasurement:
    """measurement.py:29-103: the observable(s) are popped off the trace and replaced by the measurement."""

    def __init__(self, return_type, obs=None, qubits=None, do_queue=True, after_state=False):
        if qubits is not None and obs is not None:
            raise ValueError("If an observable is provied, the qubit(s) can not be specified!")
        self.return_type = return_type
        self.obs = obs
        self.qubits = None if qubits is None else [int(q) for q in qubits]
        self.after_state = after_state
        if not do_queue:
            return
        if not _trace_stack:
            raise RuntimeError("No active circuit trace")
        ctx = _trace_stack[-1]
        if obs is not None:
            for ob in reversed(obs if isinstance(obs, list) else [obs]):
                if not ctx or ctx[-1].instance_id != ob.instance_id:
                    raise ValueError("Last content operator should be the same as 'obs'!")
                ctx.pop()
        ctx.append(self)


def expval(observable, do_queue=True):
    for ob in observable if isinstance(observable, list) else [observable]:
        if not getattr(ob, "is_observable", False):
            raise ValueError(f"{getattr(ob, 'name', ob)} is not a subclass of ObservableBase: cannot be used with expval")
    return QuantumMeasurement(Expectation, obs=observable, do_queue=do_queue)


def _check_observables(observable, what):
    for ob in observable if isinstance(observable, list) else [observable]:
        if not getattr(ob, "is_observable", False):
            raise ValueError(f"{getattr(ob, 'name', ob)} is not a subclass of ObservableBase: cannot be used with {what}")


def var(observable, do_queue=True):
    """Variance <O^2> - <O>^2 of an observable (or a product of single-qubit observables given as a list).  The
    reference declares it and raises NotImplementedError (measurement.py:158-163)."""
    _check_observables(observable, "var")
    return QuantumMeasurement(Variance, obs=observable, do_queue=do_queue)


def sample(observable, num_shots, do_queue=True):
    """``num_shots`` eigenvalue samples of an observable on at most 4 qubits, drawn on the device from the exact
    outcome distribution (reference: declared, NotImplementedError, measurement.py:166-171).  Not differentiable."""
    _check_observables(observable, "sample")
    ms = QuantumMeasurement(Sample, obs=observable, do_queue=do_queue)
    ms.num_shots = int(num_shots)
    return ms


def probs(qubits=None, do_queue=True, after_state=False):
    # the reference drops after_state here (measurement.py:206); kept for callers that set it on the object
    return QuantumMeasurement(Probability, qubits=qubits, do_queue=do_queue, after_state=after_state)


def state(do_queue=True):
    return QuantumMeasurement(State, do_queue=do_queue)


class Circuit:
    """Trace ``func(*params)`` once into gate and measurement lists (circuit.py:42-98)."""

    def __init__(self, func: Callable, num_qubits: int, *params, **kwargs):
        if not num_qubits:
            raise ValueError("Error in Circuit class, num_qubits cannot be None!")
        self._num_qubits = int(num_qubits)
        shapes = kwargs.get("parameter_shapes")
        if shapes:
            params = tuple((np.random.rand(*s) + 0.01) * np.e / 1.77 for s in shapes)
        _trace_stack.append([])
        try:
            func(*params)
        finally:
            ctx = _trace_stack.pop()
        self._init_state = ctx[0] if ctx and getattr(ctx[0], "_is_preparation", False) else None
        self._operators = [o for o in ctx if isinstance(o, Operator) and not o._is_preparation]
        self._measurements = [o for o in ctx if isinstance(o, QuantumMeasurement)]
        if not self._measurements:
            raise ValueError("No measurement! please specify a quantum measurement!")
        top = max([max(o.qubits) for o in self._operators if o.qubits] + [-1])
        if top + 1 > self._num_qubits:
            raise ValueError(
                f"Input number of qubits is not large enough! Maximum qubit number of operators is {top + 1}")

    num_qubits = property(lambda self: self._num_qubits)
    operators = property(lambda self: self._operators)
    measurements = property(lambda self: self._measurements)
    init_state = property(lambda self: self._init_state)

    def compilecircuit(self, backend=None, **kwargs):
        """String dispatch as circuit.py:268-285; this front end only knows the B200 backend."""
        from .backend import BACKEND_NAME, QUDIO_BACKEND_NAME, B200Backend, B200QUDIOBackend

        if backend == BACKEND_NAME:
            return B200Backend(backend, self, **kwargs)
        if backend == QUDIO_BACKEND_NAME:
            return B200QUDIOBackend(backend, self, **kwargs)
        raise ValueError(f"{backend}: unknown backend input")


def HardwareEfficient(n_wires: int, depth: int, params):
    """Same gate sequence as tedq/templates/layers.py:96-115 (RY,RZ on every wire, CNOT brick, final RY,RZ)."""
    RY, RZ, CNOT = globals()["RY"], globals()["RZ"], globals()["CNOT"]
    for layer in range(depth):
        for w in range(n_wires):
            RY(params[2 * layer][w], qubits=[w])
            RZ(params[2 * layer + 1][w], qubits=[w])
        for first in (2, 1):
            for w in range(first, n_wires, 2):
                CNOT(qubits=[w - 1, w])
    for w in range(n_wires):
        RY(params[2 * depth][w], qubits=[w])
        RZ(params[2 * depth + 1][w], qubits=[w])
