"""Per-step profile (tq_tn_profile: CUDA events around every step) of a config-5 plan found with the given planner
options:  python scripts/c5_plan_profile.py REPEATS SWEEPS LEAVES T0 G OUT.json   (plan from the on-disk cache)."""
import json
import sys

sys.path.insert(0, ".")
import torch

import bench
import tedq_b200 as qb
from tedq_b200 import workloads as W

sys_argv = sys.argv
reps, sweeps, leaves, t0, g, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
spec = W.lattice_rcs(5, 8, 12, seed=0)
circ = W.build_circuit(spec, qb)
hyper = {"max_repeats": reps, "reconf_sweeps": sweeps, "reconf_leaves": leaves,
         "time_model": None if sweeps == 0 else (2.0e14, 2.5e12, t0) + ((2.5e13, 1.5e12) if "--calibrated" in sys.argv else ()),
         "slicing_opts": dict(bench.C5_HYPER["slicing_opts"]), "plan_cache": bench.PLAN_CACHE, "slice_batch": g}
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=hyper)
bits = torch.zeros((1, 40), dtype=torch.int64)
cc.amplitudes(bits)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(3):
    cc.amplitudes(bits)
ev1.record()
torch.cuda.synchronize()
rows, amps_per_seq, sets_per_seq, pin_ms = cc._tn.amplitudes_profile(torch.zeros((1, 0), device="cuda"), bits, 0)
info = cc._tn._amplitude_plan()[1]
json.dump({"ms_per_amplitude": ev0.elapsed_time(ev1) / 3, "sets_per_seq": sets_per_seq, "pin_ms": pin_ms, "n_slices": info.n_slices,
           "width": info.width, "flops_log2": info.flops_log2, "rows": rows}, open(out, "w"))
per = [r for r in rows if r["per_slice"]]
print("%.2f ms per amplitude; per-slice steps %d: %.3f ms per launch sequence of %d slices; once-per-call %.3f ms + pinned images %.3f ms"
      % (ev0.elapsed_time(ev1) / 3, len(per), sum(r["ms"] for r in per), sets_per_seq, sum(r["ms"] for r in rows if not r["per_slice"]), pin_ms))
