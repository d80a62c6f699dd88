"""Search the contraction plans bench.py uses for BASELINE config 5 once (CPU only, no GPU needed) and store them in
the planner's on-disk cache under ted-q_b200/plans/ (bench.PLAN_CACHE).  Re-run after changing the planner or
bench.C5_HYPER; bench.py falls back to searching when a file is missing."""
import sys
import time

sys.path.insert(0, ".")
import bench
import tedq_b200 as qb
from tedq_b200 import workloads as W

spec = W.lattice_rcs(5, 8, 12, seed=0)
circ = W.build_circuit(spec, qb)
for greedy in (False, True):
    hyper = bench.c5_hyper(greedy, False)
    t = time.time()
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=hyper)
    info = cc._tn._amplitude_plan()[1]
    print("c5%s:" % ("g" if greedy else ""), info.n_slices, "slices, width", info.width, "log2 flops/slice %.2f" % info.flops_log2,
          "%.1f s" % (time.time() - t))
t = time.time()
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=True,
                         hyper_opt={"max_repeats": 64, "plan_cache": bench.PLAN_CACHE})
info = cc._tn._amplitude_plan()[1]
print("c5s:", info.n_slices, "slices, width", info.width, "log2 flops/slice %.2f" % info.flops_log2, "%.1f s" % (time.time() - t))
t = time.time()
dt_info = bench.c5_cpu_slices(0, slice_ids=[])
print("cpu arm plan: %d slices (%.1f s, a cache hit shares the engine's file)" % (dt_info[2], time.time() - t))
