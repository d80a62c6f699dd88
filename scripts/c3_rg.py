"""C3 (20-qubit HEA) forward / backward times: default sweeps against the register-group sweeps (structure = 2)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import tedq_b200 as qb
from tedq_b200 import workloads as W

spec = W.hea(20, 10)
circ = W.build_circuit(spec, qb)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.tensor(np.random.RandomState(0).rand(B, spec["n_params"]), dtype=torch.float32, device="cuda")
variants = [{"structure": -1}, {}, {"structure": 2, "coalesce_bits": 2}, {"structure": 2, "coalesce_bits": 1}]
if len(sys.argv) > 2:
    variants = [eval(v) for v in sys.argv[2:]]
ref = None
for opts in variants:
    cc = circ.compilecircuit(backend="pytorch_b200", plan_opts=opts or None)
    plan = cc.plan(x.device)
    out = torch.empty((B, plan.out_reals), device="cuda")
    dy = torch.ones_like(out)
    grad = torch.empty((B, plan.n_params), device="cuda")
    wsb = plan.workspace_bytes(B, True)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    def step():
        plan.forward(x.data_ptr(), B, out.data_ptr(), ws.data_ptr(), wsb, True, st)
        e1.record()
        plan.backward(x.data_ptr(), B, dy.data_ptr(), grad.data_ptr(), ws.data_ptr(), wsb, st)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    f = b = 0.0
    for _ in range(3):
        e0.record(); step(); e2.record(); torch.cuda.synchronize()
        f += e0.elapsed_time(e1) / 3; b += e1.elapsed_time(e2) / 3
    if ref is None:
        ref = (out.clone(), grad.clone())
    print(opts, "blocks", plan.num_blocks(), "ops", plan.op_stats(False)[0], plan.op_stats(True)[0], "groups",
          plan.num_register_groups(False), plan.num_register_groups(True), "sweeps", plan.num_sweeps(False), plan.num_sweeps(True))
    print("    fwd %.2f ms bwd %.2f ms -> %.0f evals/s" % (f, b, B / ((f + b) * 1e-3)),
          "max|dout| %.2e max|dgrad| %.2e" % (float((out - ref[0]).abs().max()), float((grad - ref[1]).abs().max())), flush=True)
