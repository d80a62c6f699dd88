"""Config 5 with tn_simplify=True: CNOT controls and RZ gates stay on shared wire indices (tn_simplify.py)."""
import sys, time
sys.path.insert(0, ".")
import torch
import tedq_b200 as qb
from tedq_b200 import workloads as W
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
slices = int(sys.argv[2]) if len(sys.argv) > 2 else 1
spec = W.lattice_rcs(5, 8, 12, seed=0)
circ = W.build_circuit(spec, qb)
t = time.time()
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=True,
                         hyper_opt={"max_repeats": reps, "slicing_opts": {"target_num_slices": slices}})
bits = [0] * 40
amp = cc.amplitude(bits); torch.cuda.synchronize()
print("compile+first %.2fs" % (time.time() - t))
net, info, plan = cc._tn._amplitude_plan()[:3]
kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
print(info, "flops %.3e" % (plan.flops * plan.n_slices), {k: kinds.count(k) for k in sorted(set(kinds))})
for _ in range(3):
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10):
        amp = cc.amplitude(bits)
    torch.cuda.synchronize(); dt = (time.time() - t) / 10
    print("amp", complex(amp.cpu()), "time %.3f ms" % (dt * 1e3), "TFLOP/s %.2f" % (plan.flops * plan.n_slices / dt / 1e12))
rows = cc._tn.amplitude_profile(torch.zeros((1, 0), device="cuda"), bits, 0)
print("profiled total %.3f ms" % sum(r["ms"] for r in rows))
for r in sorted(rows, key=lambda r: -r["ms"])[:12]:
    print(r)
