import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
import tedq_b200 as qb
from tedq_b200 import workloads as W
spec = W.lattice_rcs(5, 8, 12, seed=0)
cc = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=bench.c5_hyper(False, False))
bits = bench.c5_bitstrings(4)
cc.amplitudes(bits[:1])
members = cc._tn.slice_members(0)
for which in range(3):
    dt, amps, ns, fl = bench.c5_cpu_slices(0, slice_ids=members, bits=bits[which])
    got = complex(cc.amplitude(bits[which].tolist(), slice_range=(0, 1)).cpu())
    print(which, bits[which].tolist(), "cpu", sum(amps), "gpu", got, "scale", max(abs(a) for a in amps))
# single-bit patterns
for q in (0, 7, 20, 39):
    b = [0] * 40; b[q] = 1
    dt, amps, ns, fl = bench.c5_cpu_slices(0, slice_ids=members, bits=b)
    got = complex(cc.amplitude(b, slice_range=(0, 1)).cpu())
    print("bit", q, "cpu", sum(amps), "gpu", got)
