"""ORACLE — test infrastructure, NOT product code.

CPU restatement (numpy, complex128 or complex64) of the reference's tensor-network branch:
  * index maps            tedq/tensor_network/tensor_network.py:850-1099 (gen_tensor_networks)
  * operand assembly      tedq/backends/compiled_circuit.py:442-467, pytorch_backend.py:311-336
  * adjoint operands      pytorch_backend.py:524-546  (reshape(d, d).T.conj().reshape)
  * contraction           the reference hands the arrays to a third-party tree
                          (``tree.contract(arrays, backend='torch')``, pytorch_backend.py:339; cotengra /
                          jdtensorpath / opt_einsum are not vendored, not pinned, not installable here):
                          restated as the published pairwise-einsum algorithm over an ssa path.

PARITY UNPINNED for contraction values by the reference's own tests (there are none:
test_pytorch_backend.py:16 "#TODO: add tests for cotengra"); pinned instead by (i) index maps bit-exact
against fixtures produced by the reference's gen_tensor_networks (tests/golden/tn_index_maps.json) and
(ii) TN result == state-vector result of the pinned SV oracle on the same circuit.
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch

from . import sv_ref


def index_maps(circuit):
    """[(inputs, output)] per measurement; integer index ids (wire q starts on id q)."""
    n = circuit.num_qubits
    wire0, cur0 = list(range(n)), n - 1
    base = [[q] for q in range(n)]

    def thread(wire, cur, qs):
        new = [cur + 1 + j for j in range(len(qs))]
        idx = new + [wire[q] for q in qs]          # [out..., in...], tensor_network.py:871-919
        for q, i in zip(qs, new):
            wire[q] = i
        return idx, cur + len(qs)

    for op in circuit.operators:
        idx, cur0 = thread(wire0, cur0, list(op.qubits))
        base.append(idx)
    nets = []
    for ms in circuit.measurements:
        rt = getattr(ms.return_type, "value", ms.return_type)
        wire, cur, inputs, out = list(wire0), cur0, [list(t) for t in base], []
        if rt == "state":                           # :953-955
            nets.append((inputs, [wire[q] for q in range(n)]))
            continue
        if rt == "expval":                          # :986-1015
            for ob in (ms.obs if isinstance(ms.obs, list) else [ms.obs]):
                idx, cur = thread(wire, cur, list(ob.qubits))
                inputs.append(idx)
        elif ms.qubits is not None:                 # :1017-1022
            out = [wire[q] for q in ms.qubits]
        for op in reversed(circuit.operators):      # :1027-1072
            idx, cur = thread(wire, cur, list(op.qubits))
            inputs.append(idx)
        inputs += [[wire[q]] for q in range(n)]     # :1076-1078
        nets.append((inputs, out))
    return nets


def operands(circuit, flat, cdtype=torch.complex128):
    """Per measurement, the arrays in the reference's order (pytorch_backend.py:311-336)."""
    n = circuit.num_qubits
    gates, adj = [], []
    count = 0
    for op in circuit.operators:
        k = len(op.qubits)
        tp = list(op.trainable_params)
        if tp:
            pars = [p.reshape(1).to(flat.dtype) if torch.is_tensor(p) else torch.tensor([float(p)], dtype=flat.dtype)
                    for p in op.parameters]
            for i, pos in enumerate(tp):
                pars[pos] = flat[count + i].reshape(1)
            count += len(tp)
            g = sv_ref.gate_tensor(op.name, pars, cdtype)
        else:
            g = sv_ref.fixed_tensor(op.matrix, k, cdtype)
        g = g.detach().numpy()
        d = 2 ** k
        gates.append(g)
        adj.append(g.reshape(d, d).T.conj().reshape(g.shape))   # complex_conjugate, :524-546
    cap = np.array([1.0, 0.0], dtype=gates[0].dtype if gates else np.complex128)
    res = []
    for ms in circuit.measurements:
        rt = getattr(ms.return_type, "value", ms.return_type)
        arrays = [cap] * n + list(gates)
        if rt == "state":
            res.append(arrays)
            continue
        if rt == "expval":
            for ob in (ms.obs if isinstance(ms.obs, list) else [ms.obs]):
                kk = len(ob.qubits)
                arrays.append(np.asarray(ob.matrix).astype(cap.dtype).reshape([2] * (2 * kk)))
        arrays += list(reversed(adj))
        arrays += [cap] * n
        res.append(arrays)
    return res


def contract_path(arrays, inputs, output, path):
    """Pairwise contraction along an ssa path (what every tree.contract does): einsum on each pair keeping
    every index that still appears elsewhere or in the output."""
    live = {i: (np.asarray(a), list(ix)) for i, (a, ix) in enumerate(zip(arrays, inputs))}
    nxt = len(arrays)
    for a, b in path:
        A, ia = live.pop(a)
        B, ib = live.pop(b)
        others = set(output)
        for _, ix in live.values():
            others |= set(ix)
        keep = [i for i in dict.fromkeys(ia + ib) if i in others]
        sym = {ix: j for j, ix in enumerate(dict.fromkeys(ia + ib))}
        live[nxt] = (np.einsum(A, [sym[i] for i in ia], B, [sym[i] for i in ib], [sym[i] for i in keep]), keep)
        nxt += 1
    (T, ix), = live.values()
    return np.transpose(T, [ix.index(i) for i in output]) if output else T


def contract_slice_torch(arrays, inputs, output, path, sliced=(), slice_id=0, dtype=torch.complex64):
    """One slice of a sliced contraction the way the reference's planners run it on the CPU
    (``tree.contract(arrays, backend='torch')``, pytorch_backend.py:339: every pairwise step is a
    ``torch.tensordot`` = one cgemm on the host cores; PathOptimizer.rst:38-63: a sliced index is fixed to
    one value per slice and the slice results are summed).  ``sliced[j]`` takes bit j of ``slice_id``."""
    fixed = {ix: (slice_id >> j) & 1 for j, ix in enumerate(sliced)}
    live = {}
    for i, (a, ix) in enumerate(zip(arrays, inputs)):
        t = torch.as_tensor(np.asarray(a)).to(dtype)
        keep = []
        sel = []
        for ax in ix:
            if ax in fixed:
                sel.append(fixed[ax])
            else:
                sel.append(slice(None))
                keep.append(ax)
        live[i] = (t[tuple(sel)], keep)
    count = {}
    for _, ix in live.values():
        for ax in ix:
            count[ax] = count.get(ax, 0) + 1
    for ax in output:
        count[ax] = count.get(ax, 0) + 1
    nxt = len(arrays)
    for a, b in path:
        A, ia = live.pop(a)
        B, ib = live.pop(b)
        shared = [ax for ax in ia if ax in ib]
        contract = [ax for ax in shared if count[ax] == 2]
        if len(contract) == len(shared):
            T = torch.tensordot(A, B, dims=([ia.index(ax) for ax in contract], [ib.index(ax) for ax in contract]))
            keep = [ax for ax in ia if ax not in contract] + [ax for ax in ib if ax not in contract]
        else:   # an index shared with a third tensor survives: einsum keeps it
            keep = [ax for ax in dict.fromkeys(ia + ib) if ax not in contract]
            sym = {ax: j for j, ax in enumerate(dict.fromkeys(ia + ib))}
            T = torch.einsum(A, [sym[ax] for ax in ia], B, [sym[ax] for ax in ib], [sym[ax] for ax in keep])
            for ax in shared:
                if ax not in contract:
                    count[ax] -= 1
        for ax in contract:
            count[ax] = 0
        live[nxt] = (T, keep)
        nxt += 1
    (T, ix), = live.values()
    return T.permute([ix.index(ax) for ax in output]) if output else T


def greedy_path(inputs, output):
    """Smallest-result-first pairwise order; only used to contract oracle networks on the CPU."""
    live = {i: set(ix) for i, ix in enumerate(inputs)}
    nxt = len(inputs)
    path = []
    while len(live) > 1:
        best = None
        keys = list(live)
        for x in range(len(keys)):
            for y in range(x + 1, len(keys)):
                a, b = keys[x], keys[y]
                if not (live[a] & live[b]) and best is not None:
                    continue
                others = set(output)
                for k2, s in live.items():
                    if k2 not in (a, b):
                        others |= s
                size = len((live[a] | live[b]) & others)
                score = (0 if live[a] & live[b] else 1, size)
                if best is None or score < best[0]:
                    best = (score, a, b, (live[a] | live[b]) & others)
        _, a, b, res = best
        del live[a], live[b]
        live[nxt] = res
        path.append((a, b))
        nxt += 1
    return path


def run_tn(circuit, flat, cdtype=torch.complex128, path_fn=None):
    """All measurements of one parameter set through the TN branch -> list of numpy results
    (``.real`` for expval/probs as pytorch_backend.py:340,:348)."""
    out = []
    nets = index_maps(circuit)
    for ms, (inputs, output), arrays in zip(circuit.measurements, nets, operands(circuit, flat, cdtype)):
        path = (path_fn or greedy_path)(inputs, output)
        r = contract_path(arrays, inputs, output, path)
        rt = getattr(ms.return_type, "value", ms.return_type)
        out.append(r if rt == "state" else np.real(r))
    return out
