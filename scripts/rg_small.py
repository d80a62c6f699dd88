"""Register-group sweeps against the default sweeps on smaller HEA circuits (whole state resident in shared memory up to
13-14 qubits, tiled above): forward / adjoint ms per batch."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import tedq_b200 as qb
from tedq_b200 import workloads as W
for n, depth, B in ((10, 10, 1024), (12, 10, 1024), (13, 10, 512), (14, 10, 256), (16, 10, 256)):
    spec = W.hea(n, depth)
    circ = W.build_circuit(spec, qb)
    x = torch.tensor(np.random.RandomState(0).rand(B, spec["n_params"]), dtype=torch.float32, device="cuda")
    line = []
    for structure in (-1, 2):
        cc = circ.compilecircuit(backend="pytorch_b200", plan_opts={"structure": structure})
        plan = cc.plan(x.device)
        out = torch.empty((B, plan.out_reals), device="cuda"); dy = torch.ones_like(out)
        grad = torch.empty((B, plan.n_params), device="cuda")
        wsb = plan.workspace_bytes(B, True); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
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
        line.append("%s: fwd %.3f bwd %.3f ms -> %.0f evals/s" % ("groups" if structure == 2 else "default", f, b, B / ((f + b) * 1e-3)))
    print("hea%d depth %d, %d sets | %s" % (n, depth, B, " | ".join(line)), flush=True)
