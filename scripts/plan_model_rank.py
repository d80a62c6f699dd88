"""Step-time model against measurement on the config-5 plans: estimated ms per amplitude of each plan under the old model
(every step above the fused-run size at the tensor-core rate) and the calibrated one (steps the engine cannot put on
tensor cores at the FP32 GEMM rate), next to the measured B200 times.  CPU only.
    python scripts/plan_model_rank.py"""
import json
import sys

sys.path.insert(0, ".")
import bench
import tedq_b200 as qb
from tedq_b200 import planner, tn_index, workloads as W

spec = W.lattice_rcs(5, 8, 12, seed=0)
circ = W.build_circuit(spec, qb)
net = tn_index.index_maps(circ.num_qubits, [list(op.qubits) for op in circ.operators], [("state", None)])[0]
inputs = [list(t) for t in net.inputs] + [[ix] for ix in net.output]
import os
CACHE = os.path.join(bench.ROOT, "profiles", "r02_plans")   # the plans that were measured (copies of the planner's cache files)
variants = {
    "old (first round-2 bench plan: 64 repeats, 6 sweeps, 3-value model)": dict(max_repeats=64, reconf_sweeps=6, reconf_leaves=8,
                                                                                time_model=(2.0e14, 2.5e12, 1.2e-5)),
    "c5g (plain greedy, 64 repeats)": dict(max_repeats=64, reconf_sweeps=0, reconf_leaves=8, time_model=None),
    "large search under the 3-value model (128 repeats, 10 sweeps, 9 leaves)": dict(max_repeats=128, reconf_sweeps=10, reconf_leaves=9,
                                                                                   time_model=(2.0e14, 2.5e12, 1.5e-6)),
    "seed0 of the same search under the calibrated model": dict(max_repeats=128, reconf_sweeps=10, reconf_leaves=9,
                                                                time_model=(2.0e14, 2.5e12, 1.5e-6, 2.5e13, 1.5e12)),
    "c5 (BENCH PLAN: model-best of 8 restarts of that search = seed 5)": dict(max_repeats=128, reconf_sweeps=10, reconf_leaves=9, restarts=8,
                                                                            time_model=(2.0e14, 2.5e12, 1.5e-6, 2.5e13, 1.5e12)),
    "small search under the calibrated model (64 repeats, 6 sweeps, 8 leaves)": dict(max_repeats=64, reconf_sweeps=6, reconf_leaves=8,
                                                                                    time_model=(2.0e14, 2.5e12, 1.5e-6, 2.5e13, 1.5e12)),
}
# ms per amplitude on one B200, 32 slices per launch sequence (profiles/r02_plan_profile_*.json)
measured = {"old": 18.6, "c5g": 41.3, "large": 26.9, "seed0": 9.8, "c5": 6.3, "small": 14.3}
for name, kw in variants.items():
    def search():
        raise SystemExit("plan not in the cache: run scripts/make_bench_plans.py / scripts/c5_try_plan.py ... --plan-only first")
    info = planner.cached_plan(bench.PLAN_CACHE if "restarts" in kw else CACHE, inputs, [], search, seed=0, minimize="flops", target_size=2 ** 27,
                               target_num_slices=64, **kw)
    row = []
    for label, model in (("3-value", (2.0e14, 2.5e12, 1.5e-6)), ("calibrated", (2.0e14, 2.5e12, 1.5e-6, 2.5e13, 1.5e12))):
        t = planner.path_time(inputs, [], info.path, info.sliced, model) * info.n_slices
        row.append("%s %.1f ms" % (label, t * 1e3))
    key = name.split(" ")[0]
    print("%-80s flops/amplitude %.2e  width %d | model: %s | measured %.1f ms" % (
        name, 2.0 ** info.flops_log2 * info.n_slices, info.width, ", ".join(row), measured[key]))
