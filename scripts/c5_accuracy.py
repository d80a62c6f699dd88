import sys, time
sys.path.insert(0, ".")
import torch
import tedq_b200 as qb
from tedq_b200 import capi, workloads as W
spec = W.lattice_rcs(5, 8, 12, seed=0)
wrap = lambda v: torch.tensor(float(v), dtype=torch.float64)
circ = W.build_circuit(spec, qb, tensor_fn=wrap)
ho = {"max_repeats": 16, "slicing_opts": {"target_size": 2 ** 27, "target_num_slices": 64}}
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=ho, dtype=torch.complex128)
bits = [0] * 40
t = time.time(); ref = complex(cc.amplitude(bits).cpu()); print("c128", ref, time.time() - t)
circ32 = W.build_circuit(spec, qb)
for tc in (0, 1):
    for chunk in ((32,) if tc == 0 else (16, 32, 64)):
        h = dict(ho); h["engine_opts"] = {capi.TN_OPT_TENSOR_CORE: tc, capi.TN_OPT_TC_CHUNK: chunk}
        c32 = circ32.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=h)
        a = complex(c32.amplitude(bits).cpu())
        print("tc", tc, "chunk", chunk, a, "rel err vs c128 %.3e" % (abs(a - ref) / abs(ref)))
        # per-slice relative errors
        if tc == 1 and chunk == 32:
            errs = []
            for s in range(0, 64, 8):
                r = complex(cc.amplitude(bits, slice_range=(s, s + 1)).cpu()); x = complex(c32.amplitude(bits, slice_range=(s, s + 1)).cpu())
                errs.append((abs(r), abs(x - r) / abs(r)))
            print("slices |ref|, rel err:", ["%.2e %.1e" % e for e in errs])
