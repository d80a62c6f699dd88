"""bench.py's parity metric for config-5 plans, with the complex128 GPU contraction as the referee: error of the first
32-slice group relative to the largest single slice of the group, per accumulation chunk.  [SEED [PLAN_DIR]]"""
import sys
sys.path.insert(0, ".")
import torch
import bench
import tedq_b200 as qb
from tedq_b200 import capi, workloads as W
spec = W.lattice_rcs(5, 8, 12, seed=0)
hyper = bench.c5_hyper(False, False)
if len(sys.argv) > 1:
    hyper["seed"] = int(sys.argv[1]); hyper["plan_cache"] = sys.argv[2] if len(sys.argv) > 2 else "scratch_plans"
circ64 = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
h1 = dict(hyper); h1["slice_batch"] = 0
cc = circ64.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=h1, dtype=torch.complex128)
allbits = bench.c5_bitstrings(4)
refs = [[complex(cc.amplitude(b.tolist(), slice_range=(i, i + 1)).cpu()) for i in range(64)] for b in allbits]
circ32 = W.build_circuit(spec, qb)
for chunk in (16, 32):
    h = dict(hyper); h["engine_opts"] = {capi.TN_OPT_TC_CHUNK: chunk}
    c32 = circ32.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=h)
    members = [c32._tn.slice_members(i) for i in range(2)]
    out = []
    for b, ref in zip(allbits, refs):
        for i, m in enumerate(members):
            got = complex(c32.amplitude(b.tolist(), slice_range=(i, i + 1)).cpu())
            want = sum(ref[s] for s in m)
            scale = max(abs(ref[s]) for s in m)
            out.append((abs(got - want) / scale, abs(want) / scale))
    print("chunk %d: bench metric per (bitstring, group): %s" % (chunk, ", ".join("%.2e (|sum|/max %.1f)" % e for e in out)), flush=True)
