"""Short C5 run for ncu: compile the bench's 40-qubit amplitude plan and contract N amplitudes (default 2).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_c5.py
"""
import sys
sys.path.insert(0, ".")
import torch
import bench
import tedq_b200 as qb
from tedq_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spec = W.lattice_rcs(5, 8, 12, seed=0)
cc = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                                              hyper_opt=dict(bench.c5_hyper(False, False), overlap_prepare=False))
bits = bench.c5_bitstrings(max(2, n))
for a in range(n):
    amp = cc.amplitude(bits[a].tolist())
torch.cuda.synchronize()
print("amplitude", complex(amp.cpu()))
plan = cc._tn._amplitude_plan()[2]
for s_ in range(plan.n_steps):
    if plan.step_kernel(s_) == 2:
        st = plan.step(s_)
        print("tc step", s_, "k%d m%d n%d b%d" % st[2:6], "flags", plan.step_flags(s_), "fuse_to", plan.step_fuse_to(s_),
              "mode", plan.step_fuse_mode(s_))
