"""C2 in tensor-network mode, forward only, a few calls (for ncu launch lists / timing)."""
import sys, time
sys.path.insert(0, ".")
import torch
import tedq_b200 as qb
from tedq_b200 import workloads as W
spec = W.mbl_1d(12)
circ = W.build_circuit(spec, qb)
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt={"max_repeats": 8})
x = torch.tensor(W.c2_inputs(256, 12, 0), device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = cc.batched(x)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        y = cc.batched(x)
    torch.cuda.synchronize()
    print("fwd wall ms per call", (time.perf_counter() - t) * 100)
