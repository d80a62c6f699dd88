"""C3 (20-qubit HEA depth 10), a few forward+backward steps on batch 32 (for ncu captures)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import tedq_b200 as qb
from tedq_b200 import workloads as W
spec = W.hea(20, 10)
circ = W.build_circuit(spec, qb)
cc = circ.compilecircuit(backend="pytorch_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.rand(B, spec["n_params"], device="cuda")
for _ in range(2):
    xx = x.clone().requires_grad_(True)
    cc.batched(xx).sum().backward()
torch.cuda.synchronize()
