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
