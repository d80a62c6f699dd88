"""Circuit specifications for the BASELINE.json configurations, in a front-end-neutral form.

A *spec* is plain data (JSON-serialisable):

    {"name", "num_qubits", "n_params",
     "gates": [[gate_name, [qubits], [param, ...]], ...],      param = "p<k>" (flat parameter k) | float
     "meas":  [["expval", [[obs_name, [qubits]], ...]] | ["probs", qubits|None] | ["state"], ...]}

``build_circuit(spec, qai, flat)`` traces it with any module that has TeD-Q's names: the real
``tedq`` (golden generation, in the build container only) or ``tedq_b200.frontend`` (everywhere).
Gate order matters: parameter binding is positional (compiled_circuit.py:522-547).

Generators restate the reference's example circuits:
  c1  4-qubit QNN                      SURVEY.md 8d (RX encoding, 2 x [RX,RY per wire + CNOT ladder], Z expvals)
  c2  1-D many-body localisation       examples/comparison/Many_body_Localization_1D_JT.py:49-137
  c3  hardware-efficient ansatz        tedq/templates/layers.py:96-115
  c4  2-D many-body localisation       examples/Many_body_Localization_2D.ipynb cell 5
  c5  lattice random circuit           SURVEY.md 8d option B (amplitude; reference gate classes only)
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

G_T = 2.185e6 * 1 * 200e-9  # g * h_bar * t_d  (1D_JT.py:24-28)


class _Builder:
    def __init__(self, name, n):
        self.spec = {"name": name, "num_qubits": n, "n_params": 0, "gates": [], "meas": []}

    def p(self):
        k = self.spec["n_params"]
        self.spec["n_params"] = k + 1
        return f"p{k}"

    def g(self, name, qubits, *params):
        self.spec["gates"].append([name, [int(q) for q in qubits], list(params)])

    def expval(self, *obs):
        self.spec["meas"].append(["expval", [[o, [int(q) for q in qs]] for o, qs in obs]])

    def probs(self, qubits=None):
        self.spec["meas"].append(["probs", None if qubits is None else [int(q) for q in qubits]])

    def state(self):
        self.spec["meas"].append(["state"])


def _h0(b, idx, jdx):
    """The fixed XX+YY coupling block (1D_JT.py:57-78): 18 gates."""
    for _ in range(1):
        b.g("Hadamard", [idx]); b.g("Hadamard", [jdx]); b.g("CNOT", [idx, jdx])
        b.g("RZ", [jdx], G_T)
        b.g("CNOT", [idx, jdx]); b.g("Hadamard", [idx]); b.g("Hadamard", [jdx])
        b.g("S", [idx]); b.g("S", [jdx])
        b.g("Hadamard", [idx]); b.g("Hadamard", [jdx]); b.g("CNOT", [idx, jdx])
        b.g("RZ", [jdx], G_T)
        b.g("CNOT", [idx, jdx]); b.g("Hadamard", [idx]); b.g("Hadamard", [jdx])
        b.g("PhaseShift", [idx], -math.pi / 2.0); b.g("PhaseShift", [jdx], -math.pi / 2.0)


def _hd(b, idx, jdx):
    b.g("RZ", [jdx], b.p())  # disorder field d[count], trainable slot (1D_JT.py:52-55)
    _h0(b, idx, jdx)


def qnn4(n=4, layers=2):
    """C1: RX(x_q) encoding; `layers` x [RX(w), RY(w) on every wire; CNOT ladder]; <Z_q> on every wire."""
    b = _Builder(f"qnn{n}", n)
    for q in range(n):
        b.g("RX", [q], b.p())
    for _ in range(layers):
        for q in range(n):
            b.g("RX", [q], b.p())
            b.g("RY", [q], b.p())
        for q in range(n - 1):
            b.g("CNOT", [q, q + 1])
    for q in range(n):
        b.expval(("PauliZ", [q]))
    return b.spec


def mbl_1d(n=12):
    """C2 (1D_JT.py:81-137).  Flat parameters: d[2(n-1)] then (n+1) x (RY, RX, RY) angles."""
    b = _Builder(f"mbl1d_{n}", n)
    for i in range(0, n, 2):
        b.g("PauliX", [i])
    for i in range(n):
        if i + 1 < n:
            _hd(b, i + 1, i)
        if i - 1 >= 0:
            _hd(b, i - 1, i)
    for i in range(n):
        b.g("RY", [i], b.p()); b.g("RX", [i], b.p()); b.g("RY", [i], b.p())
    for i in range(0, n - 7, 5):
        if i + 7 < n:
            b.g("CNOT", [i, i + 7])
    for i in range(n):
        if i + 1 < n:
            _h0(b, i + 1, i)
        if i - 1 >= 0:
            _h0(b, i - 1, i)
    b.g("RY", [n - 1], b.p()); b.g("RX", [n - 1], b.p()); b.g("RY", [n - 1], b.p())
    b.probs([n - 1])
    return b.spec


def mbl_2d(size=4, sweeps=1):
    """C4 (notebook cell 5).  ``sweeps`` > 1 repeats the Hd sweep and the H0 sweep that many times
    ("depth 20" = 10 + 10 Trotter sweeps); disorder slots are fresh per repetition."""
    n = size * size
    b = _Builder(f"mbl2d_{size}x{size}_s{sweeps}", n)
    ix = lambda i, j: size * i + j
    for i in range(size):
        for j in range(size):
            if (i + j) % 2 == 0:
                b.g("PauliX", [ix(i, j)])

    def neighbours(i, j):
        if i + 1 < size:
            yield ix(i + 1, j)
        if i - 1 >= 0:
            yield ix(i - 1, j)
        if j + 1 < size:
            yield ix(i, j + 1)
        if j - 1 >= 0:
            yield ix(i, j - 1)

    for _ in range(sweeps):
        for i in range(size):
            for j in range(size):
                for nb in neighbours(i, j):
                    _hd(b, nb, ix(i, j))
    for q in range(n):
        b.g("RY", [q], b.p()); b.g("RX", [q], b.p()); b.g("RY", [q], b.p())
    for _ in range(sweeps):
        for i in range(size):
            for j in range(size):
                for nb in neighbours(i, j):
                    _h0(b, nb, ix(i, j))
    b.g("RY", [n - 1], b.p()); b.g("RX", [n - 1], b.p()); b.g("RY", [n - 1], b.p())
    b.probs([n - 1])
    return b.spec


def hea(n=20, depth=10, measure="z_all"):
    """C3: templates/layers.py:96-115 (RY,RZ on every wire; CNOT brick (1,2),(3,4).. then (0,1),(2,3)..; final RY,RZ)."""
    b = _Builder(f"hea{n}_d{depth}", n)
    # the template indexes params[2*layer][wire] / params[2*layer+1][wire] of a (2*depth+2, n) array;
    # gates are issued wire-major (RY then RZ per wire), so flat binding is NOT the array's C order:
    # callers pass the flat vector in gate order (see hea_flat_from_matrix)
    for _ in range(depth):
        for w in range(n):
            b.g("RY", [w], b.p())
            b.g("RZ", [w], b.p())
        for first in (2, 1):
            for w in range(first, n, 2):
                b.g("CNOT", [w - 1, w])
    for w in range(n):
        b.g("RY", [w], b.p())
        b.g("RZ", [w], b.p())
    if measure == "z_all":
        for q in range(n):
            b.expval(("PauliZ", [q]))
    elif measure == "state":
        b.state()
    return b.spec


def lattice_rcs(rows=5, cols=8, cycles=12, seed=0, measure="state"):
    """C5 circuit family: every cycle = a random RX/RY/RZ(theta in [0, 2pi)) on every qubit followed by CNOTs
    on one of the 4 edge colourings of the rows x cols lattice.  Fixed angles (non-trainable gates)."""
    rng = np.random.RandomState(seed)
    n = rows * cols
    b = _Builder(f"rcs{rows}x{cols}_c{cycles}_s{seed}", n)
    ix = lambda r, c: r * cols + c
    colourings = [
        [(ix(r, c), ix(r, c + 1)) for r in range(rows) for c in range(0, cols - 1, 2)],
        [(ix(r, c), ix(r + 1, c)) for r in range(0, rows - 1, 2) for c in range(cols)],
        [(ix(r, c), ix(r, c + 1)) for r in range(rows) for c in range(1, cols - 1, 2)],
        [(ix(r, c), ix(r + 1, c)) for r in range(1, rows - 1, 2) for c in range(cols)],
    ]
    for cyc in range(cycles):
        for q in range(n):
            b.g(["RX", "RY", "RZ"][rng.randint(3)], [q], float(rng.uniform(0, 2 * math.pi)))
        for a, c in colourings[cyc % 4]:
            b.g("CNOT", [a, c])
    if measure == "state":
        b.state()
    return b.spec


def random_circuit(n, n_gates, seed, gate_pool=None, trainable_ratio=0.7, meas=None):
    """Differential-testing circuits over the whole gate set (SURVEY.md 4 "Implication for the build")."""
    rng = np.random.RandomState(seed)
    one = ["I", "Hadamard", "PauliX", "PauliY", "PauliZ", "S", "T", "SX", "RX", "RY", "RZ", "Rot", "PhaseShift"]
    two = ["CNOT", "CZ", "CY", "SWAP", "ControlledPhaseShift", "CRX", "CRY", "CRZ"]
    three = ["CSWAP", "Toffoli"]
    npar = {"RX": 1, "RY": 1, "RZ": 1, "Rot": 3, "PhaseShift": 1, "ControlledPhaseShift": 1, "CRX": 1, "CRY": 1, "CRZ": 1}
    pool = list(gate_pool) if gate_pool else one + (two if n >= 2 else []) + (three if n >= 3 else [])
    b = _Builder(f"rand{n}_{n_gates}_s{seed}", n)
    for _ in range(n_gates):
        name = pool[rng.randint(len(pool))]
        k = 1 if name in one else (2 if name in two else 3)
        qs = rng.choice(n, size=k, replace=False).tolist()
        k_par = npar.get(name, 0)
        if k_par and rng.rand() < trainable_ratio:
            params = [b.p() for _ in range(k_par)]
        else:
            params = [float(rng.uniform(-math.pi, math.pi)) for _ in range(k_par)]
        b.g(name, qs, *params)
    for m in (meas or [["expval", [["PauliZ", [0]]]]]):
        b.spec["meas"].append(m)
    return b.spec


# ---------------------------------------------------------------------------
def build_circuit(spec, qai, flat=None, tensor_fn=None):
    """Trace ``spec`` with front end ``qai``; ``flat`` = trace-time values of the flat parameters
    (defaults to 0.1*(k+1)).  ``tensor_fn`` wraps parameter values (the reference wants torch tensors)."""
    import torch

    P = spec["n_params"]
    if flat is None:
        flat = [0.1 * (k + 1) for k in range(P)]
    wrap = tensor_fn or (lambda v: torch.tensor(float(v)))
    measurement = getattr(qai, "measurement", qai)

    def circuit_def():
        for name, qubits, params in spec["gates"]:
            ctor = getattr(qai, name)
            n_train = sum(isinstance(p, str) for p in params)
            if not params:
                ctor(qubits=list(qubits))
            elif n_train == 0:
                ctor(*[wrap(p) for p in params], qubits=list(qubits), trainable_params=[])
            else:
                vals = [wrap(flat[int(p[1:])]) if isinstance(p, str) else wrap(p) for p in params]
                tp = [i for i, p in enumerate(params) if isinstance(p, str)]
                if len(tp) == len(params):
                    ctor(*vals, qubits=list(qubits))
                else:
                    ctor(*vals, qubits=list(qubits), trainable_params=tp)
        for m in spec["meas"]:
            if m[0] == "expval":
                obs = [getattr(qai, o)(qubits=list(qs)) for o, qs in m[1]]
                measurement.expval(obs[0] if len(obs) == 1 else obs)
            elif m[0] == "probs":
                measurement.probs(qubits=m[1])
            elif m[0] == "state":
                measurement.state()
            else:
                raise ValueError(m)

    return qai.Circuit(circuit_def, spec["num_qubits"])


def c2_inputs(batch=256, n=12, seed=0):
    """Synthetic C2 inputs (1D_JT.py:153-156, :184): [batch, 2(n-1) + 3(n+1)] float32 flat parameter sets.
    Half the sets are ergodic disorder, half localised; rotation angles cat(p, -p[:, :1]) with p ~ U[0,1)."""
    rng = np.random.RandomState(seed)
    N = 2 * (n - 1)
    h_erg, h_loc, t_d = 1e6, 40e6, 200e-9
    half = batch // 2
    d_erg = (rng.rand(half, N) * 2 - 1) * h_erg * t_d * math.pi
    d_loc = (rng.rand(batch - half, N) * 39 / 40.0 + 1 / 40.0) * rng.choice([-1.0, 1.0], size=(batch - half, N)) \
        * h_loc * t_d * math.pi
    d = np.concatenate([d_erg, d_loc], 0)
    p = rng.rand(batch, n + 1, 2)
    cir = np.concatenate([p, -p[:, :, :1]], axis=2).reshape(batch, -1)
    return np.concatenate([d, cir], axis=1).astype(np.float32)
