"""Config 5 (40-qubit amplitude, the bench's plan) timed for several slice-group sizes (hyper_opt["slice_batch"]).
Usage: python scripts/c5_groups.py [g ...]"""
import os
import sys
import time

sys.path.insert(0, ".")
import torch

import bench
import tedq_b200 as qb
from tedq_b200 import capi, workloads as W

gs = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 3]
spec = W.lattice_rcs(5, 8, 12, seed=0)
circ = W.build_circuit(spec, qb)
bits = [0] * 40
for g in gs:
    hyper = {"max_repeats": bench.C5_HYPER["max_repeats"], "reconf_sweeps": bench.C5_HYPER["reconf_sweeps"],
             "time_model": bench.C5_HYPER["time_model"], "slicing_opts": dict(bench.C5_HYPER["slicing_opts"]),
             "plan_cache": "/tmp/tq_plans", "slice_batch": g,
             "engine_opts": {capi.TN_OPT_TC_GATHER: int(os.environ.get("TQ_GATHER", "0"))}}
    t = time.time()
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=hyper)
    amp = cc.amplitude(bits)
    torch.cuda.synchronize()
    t_first = time.time() - t
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        cc.amplitude(bits)
    ev0.record()
    for _ in range(5):
        amp = cc.amplitude(bits)
    ev1.record()
    torch.cuda.synchronize()
    plan = cc._tn._amplitude_plan()[2]
    print("slice_batch %d: %d launch sequences of %d slices, %.2f ms per amplitude, amp %s, workspace %.2f GiB "
          "(first call incl. planning %.1f s)" % (g, plan.n_slices, len(cc._tn.slice_members(0)),
                                                  ev0.elapsed_time(ev1) / 5, complex(amp.cpu()),
                                                  plan.workspace_bytes(len(cc._tn.slice_members(0))) / 2 ** 30, t_first),
          flush=True)
    del cc
    torch.cuda.empty_cache()
