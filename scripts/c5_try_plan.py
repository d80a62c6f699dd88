"""Config 5 with alternative planner options (one GPU): python scripts/c5_try_plan.py REPEATS SWEEPS LEAVES T0 [G [SEED]]
T0 = the time model's seconds per launched step.  With --plan-only the plans are searched and cached, no GPU; with
--calibrated the five-value model of planner.CALIBRATED_TIME_MODEL (engine dispatch classes) at that T0."""
import os
import sys
import time

sys.path.insert(0, ".")
import torch

import bench
import tedq_b200 as qb
from tedq_b200 import workloads as W

args = [a for a in sys.argv[1:] if not a.startswith("--")]
reps, sweeps, leaves, t0 = int(args[0]), int(args[1]), int(args[2]), float(args[3])
g = int(args[4]) if len(args) > 4 else 3
seed = int(args[5]) if len(args) > 5 else 0
spec = W.lattice_rcs(5, 8, 12, seed=0)
circ = W.build_circuit(spec, qb)
from tedq_b200 import planner
tmodel = (2.0e14, 2.5e12, t0) + (tuple(planner.CALIBRATED_TIME_MODEL[3:]) if "--calibrated" in sys.argv else ())
hyper = {"max_repeats": reps, "reconf_sweeps": sweeps, "reconf_leaves": leaves, "time_model": tmodel,
         "slicing_opts": dict(bench.C5_HYPER["slicing_opts"]), "plan_cache": os.environ.get("TQ_PLAN_DIR", bench.PLAN_CACHE),
         "slice_batch": g, "seed": seed}
t = time.time()
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=hyper)
info = cc._tn._amplitude_plan()[1]
print("plan: %d slices, width %d, %.3e flop per amplitude (%.1f s)" % (info.n_slices, info.width,
                                                                     2.0 ** info.flops_log2 * info.n_slices, time.time() - t), flush=True)
if "--plan-only" in sys.argv:
    sys.exit(0)
bits = [0] * 40
amp = cc.amplitude(bits)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cc.amplitude(bits)
ev0.record()
for _ in range(5):
    amp = cc.amplitude(bits)
ev1.record()
torch.cuda.synchronize()
plan = cc._tn._amplitude_plan()[2]
kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
print("%.2f ms per amplitude, amp %s (reference plan: 2.2340735e-07+7.1186860e-08j), tensor-core steps %d, "
      "launch sequences %d" % (ev0.elapsed_time(ev1) / 5, complex(amp.cpu()), kinds.count(2), plan.n_slices), flush=True)
