"""Accuracy and time of the config-5 bench plan against its complex128 contraction on the same GPU, per accumulation
chunk of the tensor-core path (TQ_TN_OPT_TC_CHUNK: complex k accumulated inside the tensor core before the fp32
round-to-nearest drain)."""
import sys, time
sys.path.insert(0, ".")
import torch
import bench
import tedq_b200 as qb
from tedq_b200 import capi, workloads as W
spec = W.lattice_rcs(5, 8, 12, seed=0)
hyper = bench.c5_hyper(False, False)
if len(sys.argv) > 1:
    hyper["seed"] = int(sys.argv[1]); hyper["restarts"] = 1; hyper["plan_cache"] = sys.argv[2] if len(sys.argv) > 2 else bench.PLAN_CACHE
circ64 = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
cc = circ64.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=hyper, dtype=torch.complex128)
allbits = bench.c5_bitstrings(4)
refs = []
for b in allbits:
    groups = [complex(cc.amplitude(b.tolist(), slice_range=(i, i + 1)).cpu()) for i in range(2)]
    refs.append(groups)
circ32 = W.build_circuit(spec, qb)
for chunk in (8, 16, 32, 64):
    h = dict(hyper); h["engine_opts"] = {capi.TN_OPT_TC_CHUNK: chunk}
    c32 = circ32.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=h)
    errs = []
    for b, ref in zip(allbits, refs):
        for i in range(2):
            a = complex(c32.amplitude(b.tolist(), slice_range=(i, i + 1)).cpu())
            errs.append(abs(a - ref[i]) / max(abs(r) for r in ref))
    c32.amplitude(allbits[0].tolist()); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c32.amplitude(allbits[0].tolist())
    e1.record(); torch.cuda.synchronize()
    print("chunk %d: %.2f ms per amplitude, max rel err of a 32-slice group vs complex128 %.2e (mean %.2e)" % (
        chunk, e0.elapsed_time(e1) / 5, max(errs), sum(errs) / len(errs)), flush=True)
