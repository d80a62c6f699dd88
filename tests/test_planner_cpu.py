"""CPU: the planner's subtree reconfiguration keeps the path valid, never raises its cost, and the reconfigured
tree contracts to the same numbers (oracle pairwise contraction)."""
import numpy as np
import pytest
import torch

import tedq_b200 as qb
from oracle import tn_ref
from tedq_b200 import lowering, planner, tn_index, workloads as W


@pytest.mark.parametrize("seed", range(3))
def test_reconfigured_path_is_valid_cheaper_and_equal(seed):
    spec = W.lattice_rcs(3, 4, 6, seed=seed, measure="state")
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    inputs, output = tn_ref.index_maps(circ)[0]
    arrays = tn_ref.operands(circ, torch.zeros(0, dtype=torch.float64))[0]
    caps = [np.array([1.0, 0.0]), np.array([0.0, 1.0])]
    bits = [(seed + q) % 2 for q in range(12)]
    inputs = [list(t) for t in inputs] + [[ix] for ix in output]
    arrays = list(arrays) + [caps[b] for b in bits]
    base = planner.find_path(inputs, [], repeats=2, seed=seed)
    better = planner.find_path(inputs, [], repeats=2, seed=seed, reconf_sweeps=3, reconf_leaves=6)
    assert better.flops_log2 <= base.flops_log2 + 1e-9
    assert sorted(x for p in better.path for x in p) == sorted(x for p in base.path for x in p)   # every id used once
    lowering.lower(inputs, [], better.path, [])                                                  # a valid ssa tree
    a = complex(tn_ref.contract_path(arrays, inputs, [], base.path))
    b = complex(tn_ref.contract_path(arrays, inputs, [], better.path))
    assert abs(a - b) < 1e-12
    sliced = planner.slice_path(inputs, [], better, target_num_slices=4, reconf_sweeps=2, reconf_leaves=6)
    lowering.lower(inputs, [], sliced.path, sliced.sliced)
    total = sum(complex(tn_ref.contract_slice_torch(arrays, inputs, [], sliced.path, sliced.sliced, s, torch.complex128))
                for s in range(sliced.n_slices))
    assert abs(total - a) < 1e-12


def test_plan_cache_round_trip(tmp_path):
    """hyper_opt["plan_cache"]: the second search of the same network reads the stored plan; a corrupt or foreign
    file is ignored (the plan is re-costed against the network on load)."""
    inputs = [[0, 1], [1, 2, 3], [2, 4], [3, 4, 5], [0, 5]]
    calls = []

    def build():
        calls.append(1)
        return planner.slice_path(inputs, [], planner.find_path(inputs, [], repeats=2), target_num_slices=2)

    a = planner.cached_plan(str(tmp_path), inputs, [], build, max_repeats=2, target_num_slices=2)
    b = planner.cached_plan(str(tmp_path), inputs, [], build, max_repeats=2, target_num_slices=2)
    assert len(calls) == 1 and a.path == b.path and a.sliced == b.sliced and abs(a.flops_log2 - b.flops_log2) < 1e-12
    planner.cached_plan(str(tmp_path), inputs, [], build, max_repeats=3, target_num_slices=2)   # other key: new search
    assert len(calls) == 2
    for f in tmp_path.iterdir():
        f.write_text("{not json")
    c = planner.cached_plan(str(tmp_path), inputs, [], build, max_repeats=2, target_num_slices=2)
    assert len(calls) == 3 and c.path == a.path
    assert planner.cached_plan(None, inputs, [], build).path == a.path and len(calls) == 4       # no cache dir: search


def test_committed_bench_plans_are_hit(monkeypatch):
    """ted-q_b200/plans/ holds the pre-searched plans of BASELINE config 5 (scripts/make_bench_plans.py): with
    bench.py's options the planner's cache must answer without searching, and the stored plan must be the one the
    bench documents (64 slices, width 21, 8.3e11 flop per amplitude: the model-best of 8 searches under the calibrated
    step-time model)."""
    import bench

    def no_search(*a, **k):
        raise AssertionError("plan cache miss: re-run scripts/make_bench_plans.py")

    monkeypatch.setattr(planner, "find_path", no_search)
    monkeypatch.setattr(planner, "search_plan", no_search)
    dt, amps, n_slices, flops_per_slice = bench.c5_cpu_slices(0, slice_ids=[])
    assert n_slices == 64 and amps == []
    assert 7.5e11 < flops_per_slice * n_slices < 9.0e11


def test_native_subtree_dp_is_bit_identical_to_python_mirror():
    """tq_tn_subtree_order (C ABI, csrc/tq_planner.cu) against planner._subtree_dp_py on random subtrees: same cost
    bits and the same split for every subset, with and without the step-time model."""
    import random

    rng = random.Random(7)
    for trial in range(44):
        L = rng.randint(2, 8) if trial < 40 else rng.randint(9, 11)   # a few large subtrees (the C limit is 12)
        n_idx = rng.randint(1, 70)
        count = {ix: rng.randint(1, 4) for ix in range(n_idx)}
        leaf_inside, leaf_sets = [], []
        for _ in range(L):
            m = {ix: rng.randint(1, count[ix]) for ix in rng.sample(range(n_idx), rng.randint(1, min(n_idx, 24)))}
            leaf_inside.append(m)
            leaf_sets.append(frozenset(ix for ix, c in m.items() if c < count[ix] or rng.random() < 0.1))
        for ix in range(n_idx):      # a leaf set may not hold more carriers than the network has
            tot = sum(m.get(ix, 0) for m in leaf_inside)
            count[ix] = max(count[ix], tot)
        for model in (None, (2.0e14, 2.5e12, 1.2e-5)):
            b_py, s_py = planner._subtree_dp_py(leaf_sets, leaf_inside, count, model)
            b_c, s_c = planner._subtree_dp_native(leaf_sets, leaf_inside, count, model)
            assert b_py == b_c, (trial, b_py, b_c)
            assert all(int(s_c[S]) == s_py[S] for S in s_py), trial


def test_native_reconfigure_reproduces_python_path_and_committed_plan():
    """Whole reconfiguration: native and mirror give the same ssa path; and the search with bench.py's options still
    reproduces the committed 40-qubit plan (ted-q_b200/plans/), i.e. the compiled planner did not move the plan the
    GPU numbers were measured on."""
    spec = W.lattice_rcs(3, 4, 6, seed=1, measure="state")
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    inputs, output = tn_ref.index_maps(circ)[0]
    inputs = [list(t) for t in inputs] + [[ix] for ix in output]
    start = planner.find_path(inputs, [], repeats=2, seed=0).path
    # (a small network: the calibrated model's rates are scaled so that its step classes actually differ here)
    for model in (None, (2.0e14, 2.5e12, 1.2e-5), planner.CALIBRATED_TIME_MODEL, (2.0e10, 2.5e8, 1.2e-7, 2.5e9, 1.5e8)):
        a = planner.reconfigure(inputs, [], start, sweeps=2, time_model=model, native=True)
        b = planner.reconfigure(inputs, [], start, sweeps=2, time_model=model, native=False)
        assert a == b


def test_cached_plan_rejects_malformed_files(tmp_path):
    """A stored plan that does not fit the network (reused / dead ssa ids, a sliced open index, another planner
    version) is searched again and overwritten, not handed to the lowering."""
    import json
    import os

    inputs = [[0, 1], [1, 2], [2, 3], [3, 0, 4]]
    output = [4]
    calls = []

    def build():
        calls.append(1)
        return planner.find_path(inputs, output, repeats=2, seed=0)

    good = planner.cached_plan(str(tmp_path), inputs, output, build, k=1)
    (fname,) = os.listdir(tmp_path)
    path = os.path.join(tmp_path, fname)
    assert planner.valid_plan(inputs, output, good.path, good.sliced)
    bad_files = [
        {"path": [[0, 1], [0, 2], [4, 3]], "sliced": []},            # tensor 0 consumed twice
        {"path": [[0, 1], [2, 6], [4, 3]], "sliced": []},            # id 6 used before it exists
        {"path": [[0, 1], [2, 3]], "sliced": []},                    # too short
        {"path": [list(p) for p in good.path], "sliced": [4]},       # an open output index cannot be sliced
        {"path": [list(p) for p in good.path], "sliced": [9]},       # not an index of the network
        {"version": planner.PLANNER_VERSION + 1, "path": [list(p) for p in good.path], "sliced": []},
    ]
    for i, bad in enumerate(bad_files):
        with open(path, "w") as fh:
            json.dump(bad, fh)
        n = len(calls)
        info = planner.cached_plan(str(tmp_path), inputs, output, build, k=1)
        assert len(calls) == n + 1, i                 # searched again
        assert info.path == good.path
        with open(path) as fh:
            assert json.load(fh)["path"] == [list(p) for p in good.path]   # overwritten with a valid plan


@pytest.mark.parametrize("alpha,temp", [(1.0, 0.0), (0.5, 0.3), (0.0, 1.0), (1.0, 1.0)])
def test_native_greedy_pass_is_bit_identical_to_its_python_mirror(alpha, temp):
    """tq_tn_greedy_path (csrc/tq_planner.cu) against planner._greedy_once: the same ssa path and the same state of
    the random stream afterwards, on closed and open networks, with disconnected parts, and through find_path."""
    import random

    nets = []
    for seed, (rows, cols, cycles) in enumerate([(3, 4, 6), (4, 5, 8), (2, 3, 4)]):
        spec = W.lattice_rcs(rows, cols, cycles, seed=seed, measure="state")
        circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
        inputs, output = tn_ref.index_maps(circ)[0]
        nets.append(([list(t) for t in inputs], list(output)))                                   # open: a state network
        nets.append(([list(t) for t in inputs] + [[ix] for ix in output], []))                   # closed: an amplitude
    nets.append(([[0, 1], [1, 2], [7, 8], [8, 9, 7], [20]], [0, 2]))                             # disconnected parts
    nets.append(([[5]], [5]))                                                                    # one tensor
    for inputs, output in nets:
        r1, r2 = random.Random(11), random.Random(11)
        a = planner._greedy_once(inputs, output, r1, alpha, temp)
        b = planner._greedy_once_native(inputs, output, r2, alpha, temp)
        assert a == b
        assert r1.getstate() == r2.getstate()
    inputs, output = nets[3]
    p1 = planner.find_path(inputs, output, repeats=6, seed=3, native_greedy=False)
    p2 = planner.find_path(inputs, output, repeats=6, seed=3, native_greedy=True)
    assert p1.path == p2.path and p1.flops_log2 == p2.flops_log2


def test_search_plan_restarts_keep_the_model_best_search():
    """planner.search_plan: R restarts = R independent searches (seeds s .. s+R-1); the plan with the smallest modelled
    time is kept, and the executor's cache key carries the restart count only when it is used."""
    from tedq_b200 import tn_backend

    spec = W.lattice_rcs(3, 4, 6, seed=2, measure="state")
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    inputs, output = tn_ref.index_maps(circ)[0]
    inputs = [list(t) for t in inputs] + [[ix] for ix in output]
    model = (2.0e10, 2.5e8, 1.2e-7, 2.5e9, 1.5e8)      # (rates scaled down so that a 12-qubit network has real trade-offs)
    kw = dict(max_repeats=4, reconf_sweeps=2, reconf_leaves=6, time_model=model, target_num_slices=4)
    singles = [planner.search_plan(inputs, [], seed=s, **kw) for s in range(3)]
    times = [planner.path_time(inputs, [], p.path, p.sliced, model) * p.n_slices for p in singles]
    best = planner.search_plan(inputs, [], seed=0, restarts=3, **kw)
    assert best.path == singles[times.index(min(times))].path
    lowering.lower(inputs, [], best.path, best.sliced)
    ex = tn_backend.TNExecutor.__new__(tn_backend.TNExecutor)
    ex.ho = {"max_repeats": 4}
    assert "restarts" not in ex._plan_key()
    ex.ho = {"max_repeats": 4, "restarts": 8}
    assert ex._plan_key()["restarts"] == 8


def test_calibrated_step_time_model_follows_the_dispatch_classes():
    """planner.step_time_model with five values charges a step by the kernel class the engine would dispatch it to
    (csrc/tq_tn.cu: build_schedule)."""
    tc, bw, t0, fma, elem = planner.CALIBRATED_TIME_MODEL
    m5 = planner.CALIBRATED_TIME_MODEL
    flops = lambda nu: 8.0 * 2.0 ** nu
    # tensor cores: 128 x 16 free extents, k + m + n >= 20 (k 9, m 7, n 4)
    assert planner.step_time_model(16, 13, 11, 20, m5) == max(flops(20) / tc, 8.0 * (2 ** 16 + 2 ** 13 + 2 ** 11) / bw) + t0
    # the step of the mis-ranked plan: K = 2^15, 64 x 32 outputs -> one thread per output element
    assert planner.step_time_model(21, 20, 11, 26, m5) == flops(26) / elem + t0
    # 64 x 64 x 16 -> the FP32 GEMM
    assert planner.step_time_model(10, 10, 12, 16, m5) == flops(16) / fma + t0
    # <= 64 outputs with K >= 4096 -> split-K reduction, bandwidth bound
    assert planner.step_time_model(18, 17, 5, 20, m5) == max(flops(20) / tc, 8.0 * (2 ** 18 + 2 ** 17 + 2 ** 5) / bw) + t0
    # small steps ride in a fused run
    assert planner.step_time_model(8, 8, 10, 13, m5) == 1e-7
    # three values: every launched step at the tensor-core rate (the pre-calibration model)
    assert planner.step_time_model(21, 20, 11, 26, m5[:3]) == max(flops(26) / tc, 8.0 * (2 ** 21 + 2 ** 20 + 2 ** 11) / bw) + t0
