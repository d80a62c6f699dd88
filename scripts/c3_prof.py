import sys
sys.path.insert(0, ".")
import numpy as np, torch
import tedq_b200 as qb
from tedq_b200 import workloads as W
structure = int(sys.argv[1]) if len(sys.argv) > 1 else 0    # tq_plan_opts.structure: -1 default sweeps, 0 automatic, 1 real blocks + diagonal layers, 2 register groups
spec = W.hea(20, 10)
circ = W.build_circuit(spec, qb)
B = 16
x = torch.tensor(np.random.RandomState(0).rand(B, spec["n_params"]), dtype=torch.float32, device="cuda")
cc = circ.compilecircuit(backend="pytorch_b200", plan_opts={"structure": structure})
plan = cc.plan(x.device)
out = torch.empty((B, plan.out_reals), device="cuda"); dy = torch.ones_like(out); grad = torch.empty((B, plan.n_params), device="cuda")
wsb = plan.workspace_bytes(B, True); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); st = torch.cuda.current_stream().cuda_stream
plan.forward(x.data_ptr(), B, out.data_ptr(), ws.data_ptr(), wsb, True, st)
plan.backward(x.data_ptr(), B, dy.data_ptr(), grad.data_ptr(), ws.data_ptr(), wsb, st)
torch.cuda.synchronize()
