"""CPU: the tensor-network host logic.  C index maps / lowering (host-only entry points of the C ABI) against the
Python mirrors bit-exactly; planner invariants; oracle TN == oracle SV."""
import numpy as np
import pytest
import torch

import tedq_b200 as qb
from conftest import load_golden
from oracle import sv_ref, tn_ref
from tedq_b200 import capi, lowering, planner, tn_index
from tedq_b200 import workloads as W

MAPS = load_golden("tn_index_maps.json")


def _meas_args(circ, mi):
    ms = circ.measurements[mi]
    rt = getattr(ms.return_type, "value", ms.return_type)
    if rt == "expval":
        obs = ms.obs if isinstance(ms.obs, list) else [ms.obs]
        return dict(kind="expval", obs_qubits=[list(o.qubits) for o in obs])
    if rt == "probs":
        return dict(kind="probs", kept=None if ms.qubits is None else list(ms.qubits))
    return dict(kind="state")


@pytest.mark.parametrize("rec", MAPS, ids=lambda r: r["spec"]["name"])
def test_c_index_map_matches_reference_fixture(rec):
    circ = W.build_circuit(rec["spec"], qb)
    gq = [list(op.qubits) for op in circ.operators]
    for mi, ref in enumerate(rec["networks"]):
        ins, out = capi.tn_index_map(circ.num_qubits, gq, **_meas_args(circ, mi))
        assert [[capi.tn_symbol(i) for i in t] for t in ins] == ref["inputs"]
        assert [capi.tn_symbol(i) for i in out] == ref["output"]


@pytest.mark.parametrize("spec", [W.qnn4(), W.mbl_1d(6), W.hea(6, 2), W.lattice_rcs(3, 3, 4, seed=1, measure="state"),
                                  W.random_circuit(5, 30, seed=9, meas=[["probs", [3, 1]]])], ids=lambda s: s["name"])
@pytest.mark.parametrize("n_slices", [1, 8])
def test_lowering_c_matches_python_bit_exact(spec, n_slices):
    circ = W.build_circuit(spec, qb)
    for net in tn_index.networks_of_circuit(circ):
        info = planner.find_path(net.inputs, net.output, repeats=3, seed=1)
        if n_slices > 1:
            info = planner.slice_path(net.inputs, net.output, info, target_num_slices=n_slices)
            assert info.n_slices >= n_slices
        low = lowering.lower(net.inputs, net.output, info.path, info.sliced)
        c_steps, c_sl, c_fp = capi.tn_lower(net.inputs, net.output, info.path, info.sliced)
        assert c_steps == low.as_tuples()
        assert c_fp == low.final_perm
        assert c_sl == [(t, o, b) for t, l in enumerate(low.in_slice_bits) for (o, b) in l]


def test_planner_is_deterministic_and_valid():
    spec = W.lattice_rcs(3, 4, 6, seed=2, measure="state")
    circ = W.build_circuit(spec, qb)
    from tedq_b200.tn_backend import amplitude_network
    net = amplitude_network(tn_index.networks_of_circuit(circ)[0], [0] * 12)  # closed network: every index sliceable
    a = planner.find_path(net.inputs, net.output, repeats=6, seed=3)
    b = planner.find_path(net.inputs, net.output, repeats=6, seed=3)
    assert a.path == b.path and len(a.path) == len(net.inputs) - 1
    used = [x for p in a.path for x in p]
    assert sorted(used) == list(range(2 * len(net.inputs) - 2))
    s = planner.slice_path(net.inputs, net.output, a, target_size_log2=max(4, a.width - 2))
    assert s.width <= max(4, a.width - 2) and not (set(s.sliced) & set(net.output))


@pytest.mark.parametrize("spec", [
    W.random_circuit(3, 14, seed=21, meas=[["expval", [["PauliZ", [0]]]], ["expval", [["PauliX", [2]]]]]),
    W.random_circuit(4, 16, seed=22, meas=[["expval", [["PauliZ", [0]], ["PauliY", [3]]]]]),
    W.random_circuit(3, 12, seed=23, meas=[["probs", [2, 0]]]),
    W.random_circuit(3, 12, seed=24, meas=[["state"]]),
], ids=lambda s: s["name"])
def test_oracle_tn_equals_oracle_sv(spec):
    """TN result == SV result for any circuit small enough to run both."""
    wrap = lambda v: torch.tensor(float(v), dtype=torch.float64)
    circ = W.build_circuit(spec, qb, tensor_fn=wrap)
    flat = torch.tensor(np.random.RandomState(1).uniform(-3, 3, spec["n_params"]), dtype=torch.float64)
    sv = sv_ref.run_sv(circ, flat, torch.complex128).numpy()
    tn = tn_ref.run_tn(circ, flat, torch.complex128)
    for i, t in enumerate(tn):
        ref = sv[i]
        if spec["meas"][i][0] == "probs" and spec["meas"][i][1] is not None:
            # the TN branch orders the open legs as listed, the SV branch by ascending qubit (torch.sum over the rest)
            order = np.argsort(np.argsort(spec["meas"][i][1]))
            ref = np.transpose(ref, order)
        assert np.allclose(t, ref, atol=1e-12)


def test_slice_groups_partition_and_sum():
    """Host logic of hyper_opt["slice_batch"]: the grouped sliced indices leave the network and become the plan's
    batch dimension.  The groups partition the path's slices, no operand drops to rank 0, and contracting the
    re-laid-out operands set by set (oracle) reproduces the sum of the member slices."""
    import numpy as np
    import torch
    import tedq_b200 as qb
    from oracle import tn_ref
    from tedq_b200 import planner, tn_index
    from tedq_b200 import workloads as W
    from tedq_b200.tn_backend import TNExecutor, amplitude_network

    spec = W.lattice_rcs(3, 3, 5, seed=4, measure="state")
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    net = amplitude_network(tn_index.networks_of_circuit(circ)[0], [0] * 9)
    info = planner.slice_path(net.inputs, net.output, planner.find_path(net.inputs, net.output, repeats=2),
                              target_num_slices=16)
    arrays = [np.asarray(a) for a in tn_ref.operands(circ, torch.zeros(0, dtype=torch.float64))[0]] + \
        [np.array([1.0, 0.0])] * 9
    ex = object.__new__(TNExecutor)
    ex.ho = {"slice_batch": 2}
    ex.contract_parallel = False
    ex._amp = [net, info, None]
    grp = ex._slice_group(net, info, [False] * len(net.inputs))
    assert grp is not None and len(grp["indices"]) == 2
    assert sorted(grp["indices"] + grp["rest"]) == sorted(info.sliced)
    ex._amp_group = grp
    # default size: up to 2^5 slices (here all 16), fewer when the network is wide (memory) — never for
    # parameter-batched operands
    ex.ho = {}
    assert len(ex._slice_group(net, info, [False] * len(net.inputs))["indices"]) == 4
    wide = lambda w: planner.PathInfo(info.path, info.sliced, w, info.flops_log2, info.n_steps)
    assert len(ex._slice_group(net, wide(27), [False] * len(net.inputs))["indices"]) == 1
    assert ex._slice_group(net, wide(30), [False] * len(net.inputs)) is None
    assert ex._slice_group(net, info, [True] + [False] * (len(net.inputs) - 1)) is None
    ex.ho = {"slice_batch": 2}
    n_plan = 1 << len(grp["rest"])
    seen = sorted(s for i in range(n_plan) for s in ex.slice_members(i))
    assert seen == list(range(info.n_slices))
    gset = set(grp["indices"])
    inputs2 = [[ix for ix in t if ix not in gset] for t in net.inputs]
    assert all(len(t) >= 1 for t in inputs2)

    def one_slice(arrs, inputs, sliced, sid):
        sl_a, sl_i = [], []
        for a, ix in zip(arrs, inputs):
            sel = tuple(((sid >> sliced.index(i)) & 1) if i in sliced else slice(None) for i in ix)
            sl_a.append(a[sel])
            sl_i.append([i for i in ix if i not in sliced])
        return complex(tn_ref.contract_path(sl_a, sl_i, [], info.path))

    for i in (0, n_plan - 1):
        want = sum(one_slice(arrays, net.inputs, list(info.sliced), s) for s in ex.slice_members(i))
        got = 0.0
        for sset in range(4):      # the re-layout of TNExecutor._amplitude_operands, in numpy
            arrs = []
            for t, a in enumerate(arrays):
                sel = [slice(None)] * a.ndim
                for ax, gi in grp["axes"].get(t, []):
                    sel[ax] = (sset >> gi) & 1
                arrs.append(a[tuple(sel)])
            got += one_slice(arrs, inputs2, list(grp["rest"]), i)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
