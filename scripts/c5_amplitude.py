"""BASELINE config 5 on one GPU: 40-qubit lattice random circuit, one amplitude by sliced contraction.
Prints the plan, the amplitude, whole-contraction timing and the per-step table of one slice
(tq_tn_profile) with each step's roofline.  Usage: python scripts/c5_amplitude.py [max_repeats] [tc 0|1]"""
import json
import sys
import time

sys.path.insert(0, ".")
import torch

import tedq_b200 as qb
from tedq_b200 import capi, workloads as W

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
use_tc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
HBM = peaks.get("hbm_gbs", 6650.0) * 1e9
TF32 = peaks.get("bf16_tflops", 1590.0) / 2 * 1e12      # derived: no TF32 figure in MEASURED_PEAKS.json
rows_, cols_, cyc = 5, 8, 12
spec = W.lattice_rcs(rows_, cols_, cyc, seed=0)
circ = W.build_circuit(spec, qb)
t = time.time()
cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                         hyper_opt={"max_repeats": reps, "reconf_sweeps": int(sys.argv[3]) if len(sys.argv) > 3 else 0, "slicing_opts": {"target_size": 2 ** 27, "target_num_slices": 64},
                                    "engine_opts": {capi.TN_OPT_TENSOR_CORE: use_tc}})
print("compile %.2fs" % (time.time() - t))
bits = [0] * 40
flat = torch.zeros((1, 0), device="cuda")
amp = cc.amplitude(bits)
torch.cuda.synchronize()
net, info, plan = cc._tn._amplitude_plan()[:3]
total_flops = 2.0 ** info.flops_log2 * info.n_slices
print(info, "flops/slice %.3e" % 2.0 ** info.flops_log2, "slices", info.n_slices, "in", plan.n_slices,
      "launch sequences, width", plan.width, "steps", plan.n_steps)
for _ in range(2):
    torch.cuda.synchronize()
    t = time.time()
    amp = cc.amplitude(bits)
    torch.cuda.synchronize()
    dt = time.time() - t
    print("amp", complex(amp.cpu()), "time %.4f s" % dt, "algorithmic TFLOP/s %.2f" % (total_flops / dt / 1e12))
a = cc.amplitude(bits, slice_range=(0, plan.n_slices // 2)) + cc.amplitude(bits, slice_range=(plan.n_slices // 2, plan.n_slices))
print("halves", complex(a.cpu()))
rows = cc._tn.amplitude_profile(torch.zeros((1, 0), device="cuda"), bits, 0)
tot = sum(r["ms"] for r in rows)
print("one slice (+ invariant part): %.3f ms over %d steps; by kernel:" % (tot, len(rows)),
      {k: round(sum(r["ms"] for r in rows if r["kernel"] == k), 3) for k in (0, 1, 2)},
      "counts", {k: sum(1 for r in rows if r["kernel"] == k) for k in (0, 1, 2)})
print("once-per-call packing of slice-invariant operand images: %.3f ms" % plan.last_pinned_pack_ms)
per = [r for r in rows if r["per_slice"]]
print("per-slice steps: %d, %.3f ms; invariant steps: %d, %.3f ms" % (
    len(per), sum(r["ms"] for r in per), len(rows) - len(per), sum(r["ms"] for r in rows if not r["per_slice"])))
print("step  k  m  n  b kern      ms  pack_ms  alg.TFLOP/s  bound   roofline_ms  frac")
for r in sorted(rows, key=lambda r: -r["ms"])[:25]:
    fl = 8.0 * 2.0 ** (r["k"] + r["m"] + r["n"] + r["b"])
    byts = 8.0 * (2.0 ** (r["k"] + r["m"] + r["b"]) + 2.0 ** (r["k"] + r["n"] + r["b"]) + 2.0 ** (r["m"] + r["n"] + r["b"]))
    t_fl, t_by = 3 * fl / TF32, byts / HBM     # 24 TF32 flops per complex MAC (4M x 3-term split)
    roof = max(t_fl, t_by) * 1e3
    print("%4d %2d %2d %2d %2d %4d %8.3f %8.3f %10.2f  %-6s %10.4f %6.3f" % (
        r["step"], r["k"], r["m"], r["n"], r["b"], r["kernel"], r["ms"], r["pack_ms"], fl / r["ms"] / 1e9,
        "tensor" if t_fl > t_by else "hbm", roof, roof / r["ms"] if r["ms"] > 0 else 0))
json.dump(rows, open("gpurun_out/c5_steps_tc%d.json" % use_tc, "w"))
