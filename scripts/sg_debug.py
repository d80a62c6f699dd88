import os, sys
sys.path.insert(0, ".")
import torch
import tedq_b200 as qb
from tedq_b200 import capi, workloads as W
from oracle import sv_ref

spec = W.lattice_rcs(4, 5, 10, seed=2, measure="state")
circ = W.build_circuit(spec, qb)
ref = sv_ref.run_sv(circ, torch.zeros(0), torch.complex64, return_state=True).numpy().reshape(-1)
bits = [0, 1] * 10
want = ref[int("".join(str(b) for b in bits), 2)]
print("want", want, "NO_PIN", os.environ.get("TQ_TN_NO_PIN"))
def per_set(eo, g, slices):
    ho = {"max_repeats": 8, "slice_batch": g, "engine_opts": eo, "plan_cache": "/tmp/tqp",
          "slicing_opts": {"target_num_slices": slices}}
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=ho)
    tn = cc._tn
    plan, ptrs, strides, out, ws, ws_bytes, any_b, keep = tn._amplitude_operands(torch.zeros((1, 0), device="cuda"), bits)
    res = []
    for sl in range(plan.n_slices):
        out.zero_()
        plan.contract(ptrs, strides, out.shape[0], sl, sl + 1, out.data_ptr(), ws.data_ptr(), ws_bytes,
                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        res.append(out.reshape(-1).cpu().clone())
    return torch.stack(res), plan
for slices in (4, 16):
    a, _ = per_set({capi.TN_OPT_TENSOR_CORE: 0}, 2, slices)
    b, plan = per_set({capi.TN_OPT_TC_MIN_LOG2: 12}, 2, slices)
    print("slices", slices, "plan slices", plan.n_slices, "sum err tc_off %.2e tc %.2e" % (abs(a.sum() - want) / abs(want), abs(b.sum() - want) / abs(want)))
    scale = a.abs().max()
    print(" per (plan slice, set) |tc - fma| / max:", ((a - b).abs() / scale).numpy().round(6).tolist())
    print(" tc steps:", [(s, plan.step_flags(s), plan.step(s)[2:6]) for s in range(plan.n_steps) if plan.step_kernel(s) == 2])
