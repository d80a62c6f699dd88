"""CPU, world_size 2, gloo: the N>1 host logic — slice-range sharding + one all-reduce reproduces the full sum,
batch sharding covers every row exactly once."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world_size, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        import tedq_b200 as qb
        from oracle import tn_ref
        from tedq_b200 import dist as tqd
        from tedq_b200 import planner, tn_index
        from tedq_b200 import workloads as W
        from tedq_b200.tn_backend import amplitude_network

        spec = W.lattice_rcs(2, 3, 4, seed=3, measure="state")
        circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
        net = amplitude_network(tn_index.networks_of_circuit(circ)[0], [0] * 6)
        info = planner.slice_path(net.inputs, net.output, planner.find_path(net.inputs, net.output, repeats=2),
                                  target_num_slices=8)
        arrays = tn_ref.operands(circ, torch.zeros(0, dtype=torch.float64))[0] + [np.array([1.0, 0.0])] * 6
        lo, hi = tqd.shard_range(info.n_slices, rank, world_size)
        part = 0.0 + 0.0j
        for s in range(lo, hi):   # fix the sliced indices, contract, sum (PathOptimizer.rst:38-63)
            sl_arrays, sl_inputs = [], []
            for a, ix in zip(arrays, net.inputs):
                a = np.asarray(a)
                sel = tuple(((s >> info.sliced.index(i)) & 1) if i in info.sliced else slice(None) for i in ix)
                sl_arrays.append(a[sel])
                sl_inputs.append([i for i in ix if i not in info.sliced])
            part += complex(tn_ref.contract_path(sl_arrays, sl_inputs, [], info.path))
        t = torch.tensor([part], dtype=torch.complex128)
        tqd.allreduce_sum_(t)
        full = complex(tn_ref.contract_path(arrays, net.inputs, [], info.path))
        rows = list(range(*tqd.shard_range(11, rank, world_size)))
        gathered = [None] * world_size
        dist.all_gather_object(gathered, rows)
        # measurements dealt round-robin, combined with one all-reduce (TN mode: one network per measurement)
        n_meas = 5
        mine = tqd.my_measurements(n_meas)
        truth = torch.arange(3 * n_meas * 2, dtype=torch.float64).reshape(3, n_meas, 2)
        combined = tqd.combine_measurements({i: truth[:, i] for i in mine}, n_meas, truth[:, 0])
        # data rows sharded over ranks with shared weights (the QUDIO front): every rank sees all values, the
        # autograd graph covers this rank's rows only; one all-reduce completes the shared-weight gradient
        class RowEngine:
            @staticmethod
            def batched(X, w, in_dims=None):
                return (X * w).sum(1, keepdim=True)

        X = torch.arange(15, dtype=torch.float64).reshape(5, 3)
        w = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64, requires_grad=True)
        local = tqd.sharded_batched(RowEngine, X, w, in_dims=(0, None), gather=False)
        rows_all = tqd.gather_rows(local, 5, X.device, keep_local_graph=True)
        rows_all.sum().backward()
        g = tqd.allreduce_sum_(w.grad.clone())
        rows_ok = bool(torch.equal(rows_all.detach(), (X * w.detach()).sum(1, keepdim=True))) and \
            bool(torch.allclose(g, X.sum(0)))
        ret[rank] = (complex(t[0]), full, gathered, info.n_slices, mine,
                     bool(torch.equal(combined, truth)) and rows_ok)
    finally:
        dist.destroy_process_group()


def test_slice_sharding_and_allreduce_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    for rank in range(2):
        total, full, gathered, n_slices, mine, combined_ok = ret[rank]
        assert n_slices >= 8
        assert abs(total - full) < 1e-12
        assert sorted(x for part in gathered for x in part) == list(range(11))
        assert mine == list(range(rank, 5, 2)) and combined_ok
