"""ORACLE — test infrastructure, NOT product code.

CPU restatement (torch, CPU tensors) of the reference's state-vector path, written from its
behaviour, every function citing the reference lines it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module; the product package never does.

Pinned (tests/test_oracle_golden.py) against
  * the reference's own golden vectors (test/test_pytorch_backend.py:386-584), and
  * outputs/gradients of the reference itself run in the build container, committed as fixtures
    under tests/golden/ by tests/golden/generate_golden.py.

The arithmetic deliberately uses the same torch ops as the reference (tensordot + permute,
autograd for gradients) so that timing it is a fair stand-in for "the reference's CPU pytorch
path" on a box where the reference package itself is absent.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch


def _scalar(p):
    if hasattr(p, "detach"):
        p = p.detach().cpu().numpy()
    return float(np.asarray(p).reshape(-1)[0])


# ---- gate tensors from parameters: pytorch_backend.py:866-1188 -------------------------------------
def _cat(entries, cdtype):
    # the reference concatenates one-element tensors then casts to tcomplex (:887-889)
    return torch.cat([e.reshape(1).to(cdtype) for e in entries], dim=0)


def gate_tensor(name: str, p: Sequence[torch.Tensor], cdtype) -> torch.Tensor:
    one = torch.ones(1)
    zero = torch.zeros(1)
    if name == "RX":      # :866-894
        c, js = torch.cos(p[0] / 2.0), 1j * torch.sin(-p[0] / 2.0)
        return _cat([c, js, js, c], cdtype).reshape(2, 2)
    if name == "RY":      # :897-922
        c, s = torch.cos(p[0] / 2.0), torch.sin(p[0] / 2.0)
        return _cat([c, -s, s, c], cdtype).reshape(2, 2)
    if name == "RZ":      # :925-954
        e = torch.exp(-0.5j * p[0])
        return _cat([e, zero, zero, e.conj()], cdtype).reshape(2, 2)
    if name == "Rot":     # :957-984
        c, s = torch.cos(p[1] / 2.0), torch.sin(p[1] / 2.0)
        return _cat([torch.exp(-0.5j * (p[0] + p[2])) * c, -torch.exp(0.5j * (p[0] - p[2])) * s,
                     torch.exp(-0.5j * (p[0] - p[2])) * s, torch.exp(0.5j * (p[0] + p[2])) * c], cdtype).reshape(2, 2)
    if name == "PhaseShift":  # :987-1013
        return _cat([one, zero, zero, torch.exp(1.0j * p[0])], cdtype).reshape(2, 2)
    if name == "ControlledPhaseShift":  # :1016-1054
        e = [zero] * 16
        e[0] = e[5] = e[10] = one
        e[15] = torch.exp(1.0j * p[0])
        return _cat(e, cdtype).reshape(2, 2, 2, 2)
    if name in ("CRX", "CRY", "CRZ"):  # :1057-1188
        e = [zero] * 16
        e[0] = e[5] = one
        if name == "CRX":
            c, js = torch.cos(p[0] / 2.0), 1.0j * torch.sin(-p[0] / 2.0)
            e[10], e[11], e[14], e[15] = c, js, js, c
        elif name == "CRY":
            c, s = torch.cos(p[0] / 2.0), torch.sin(p[0] / 2.0)
            e[10], e[11], e[14], e[15] = c, -s, s, c
        else:
            e[10], e[15] = torch.exp(-0.5j * p[0]), torch.exp(0.5j * p[0])
        return _cat(e, cdtype).reshape(2, 2, 2, 2)
    raise KeyError(name)


def fixed_tensor(matrix, k, cdtype):
    """pytorch_backend.py:567-577: trace-time numpy matrix -> tcomplex, reshaped [2]*2k."""
    return torch.from_numpy(np.asarray(matrix)).type(cdtype).reshape([2] * (2 * k))


def perm_for(qubits, n):
    """compiled_circuit.py:126-198: where each qubit axis sits after tensordot(gate, state)."""
    rest = [q for q in range(n) if q not in qubits]
    order = list(qubits) + rest
    return [order.index(i) for i in range(n)]


def run_sv(circuit, flat: torch.Tensor, cdtype=torch.complex64, return_state=False):
    """One parameter set through the circuit.  ``flat``: 1-D real tensor, positional binding
    (compiled_circuit.py:492-547: i-th trainable slot of each gate, gate order)."""
    n = circuit.num_qubits
    # initial state (pytorch_backend.py:500-522)
    if circuit.init_state is not None and circuit.init_state:
        state = torch.from_numpy(np.asarray(circuit.init_state.matrix)).type(cdtype).reshape([2] * n)
    else:
        state = torch.zeros([2] * n, dtype=cdtype)
        state.view(-1)[0] = 1.0
    count = 0
    for op in circuit.operators:
        k = len(op.qubits)
        tp = list(op.trainable_params)
        if tp:
            pars = [p if torch.is_tensor(p) else torch.tensor([float(_scalar(p))], dtype=flat.dtype)
                    for p in op.parameters]
            pars = [q.reshape(1).to(flat.dtype) if torch.is_tensor(q) else q for q in pars]
            for i, pos in enumerate(tp):
                pars[pos] = flat[count + i].reshape(1)
            count += len(tp)
            g = gate_tensor(op.name, pars, cdtype)
        else:
            g = fixed_tensor(op.matrix, k, cdtype)
        # the hot loop, pytorch_backend.py:365-379
        state = torch.tensordot(g, state, (list(range(k, 2 * k)), list(op.qubits)))
        state = state.permute(perm_for(list(op.qubits), n))
    if count != flat.numel():
        raise ValueError(f"Error!!!! number of parameters are not matched!! required {count} but {flat.numel()} are given")
    if return_state:
        return state
    return torch.stack(measure(circuit, state, cdtype), 0)


def measure(circuit, state, cdtype) -> List[torch.Tensor]:
    """pytorch_backend.py:393-498."""
    n = circuit.num_qubits
    res = []
    axes = list(range(n))
    for ms in circuit.measurements:
        rt = getattr(ms.return_type, "value", ms.return_type)
        if rt == "expval":
            if isinstance(ms.obs, list):   # :402-424
                cur = state
                for ob in ms.obs:
                    t = fixed_tensor(ob.matrix, 1, cdtype)
                    cur = torch.tensordot(t, cur, dims=([1], list(ob.qubits))).permute(perm_for(list(ob.qubits), n))
            else:                          # :428-459
                k = len(ms.obs.qubits)
                t = fixed_tensor(ms.obs.matrix, k, cdtype)
                cur = torch.tensordot(t, state, dims=(list(range(k, 2 * k)), list(ms.obs.qubits)))
                cur = cur.permute(perm_for(list(ms.obs.qubits), n))
            res.append(torch.squeeze(torch.tensordot(torch.conj(state), cur, dims=(axes, axes)).real))
        elif rt == "probs":                # :461-472
            pr = torch.abs(state) ** 2
            if ms.qubits is not None:
                drop = [a for a in axes if a not in ms.qubits]
                if drop:
                    pr = torch.sum(pr, dim=drop)
            res.append(pr)
        elif rt == "state":                # :495-496
            res.append(state)
        else:
            raise NotImplementedError(rt)
    return res


def run_batch(circuit, flat_batch: torch.Tensor, cdtype=torch.complex64, cotangent=None):
    """Python loop over parameter sets (the reference has no batch entry: SURVEY.md 8b).
    Returns (outputs [B, ...], grads [B, P] or None); grads are of sum(cotangent * output)."""
    outs, grads = [], []
    for b in range(flat_batch.shape[0]):
        x = flat_batch[b].clone().requires_grad_(cotangent is not None)
        y = run_sv(circuit, x, cdtype)
        outs.append(y.detach())
        if cotangent is not None:
            ct = cotangent[b] if cotangent.dim() == y.dim() + 1 else cotangent
            if y.is_complex():
                # torch convention: L = Re(sum(conj(ct) * y)) has dL/dy* = ct
                loss = torch.sum(torch.view_as_real(y) * torch.view_as_real(ct.to(y.dtype)))
            else:
                loss = torch.sum(y * ct.to(y.dtype))
            loss.backward()
            grads.append(x.grad.detach().clone())
    return torch.stack(outs, 0), (torch.stack(grads, 0) if grads else None)
