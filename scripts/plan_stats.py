"""Print fusion / sweep statistics of the BASELINE circuits (needs a GPU: plans upload their tables)."""
import sys
sys.path.insert(0, ".")
import torch
import tedq_b200 as qb
from tedq_b200 import workloads as W
for spec, dt in ((W.qnn4(), torch.complex64), (W.mbl_1d(12), torch.complex64), (W.hea(20, 10), torch.complex64),
                 (W.mbl_2d(4, 1), torch.complex128)):
    rd = torch.float32 if dt == torch.complex64 else torch.float64
    cc = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd)).compilecircuit(
        backend="pytorch_b200", dtype=dt)
    p = cc.plan()
    print(spec["name"], "gates", len(spec["gates"]), "blocks", p.num_blocks(), "sweeps fwd/bwd", p.num_sweeps(False),
          p.num_sweeps(True), "hbm bytes fwd/bwd", p.hbm_bytes(False), p.hbm_bytes(True),
          "gates per fwd sweep", [p.sweep_num_gates(s, False) for s in range(p.num_sweeps(False))])
