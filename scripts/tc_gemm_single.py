"""One contraction step C[m, n] = sum_k A[m, k] B[k, n] through the C ABI on the tensor-core path, timed with the
engine's own per-step events (tq_tn_profile).  Used for ncu captures of k_tc_gemm / k_tc_pack and for kernel
experiments.  Usage: python scripts/tc_gemm_single.py n_m n_n n_k [chunk] [reps]"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from tedq_b200 import capi

n_m, n_n, n_k = (int(v) for v in sys.argv[1:4])
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 32
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
c128 = len(sys.argv) > 6 and sys.argv[6] == "c128"
ids = list(range(n_m + n_n + n_k))
M, N, K = ids[:n_m], ids[n_m:n_m + n_n], ids[n_m + n_n:]
a_idx, b_idx, o_idx = M + K, K + N, M + N
plan = capi.TnPlan([a_idx, b_idx], o_idx, [(0, 1)], [], [False, False], capi.TQ_C128 if c128 else capi.TQ_C64)
plan.set_option(capi.TN_OPT_TC_MIN_LOG2, 0)
plan.set_option(capi.TN_OPT_TC_CHUNK, chunk)
import os
plan.set_option(capi.TN_OPT_TC_GATHER, int(os.environ.get("TQ_GATHER", "0")))
assert plan.step_kernel(0) == (1 if c128 else 2)
g = torch.Generator(device="cuda").manual_seed(0)
rd = torch.float64 if c128 else torch.float32
A = torch.randn(2 ** (n_m + n_k), 2, device="cuda", generator=g, dtype=rd)
B = torch.randn(2 ** (n_k + n_n), 2, device="cuda", generator=g, dtype=rd)
out = torch.zeros((1, 2 ** (n_m + n_n)), dtype=torch.complex128 if c128 else torch.complex64, device="cuda")
ws_bytes = plan.workspace_bytes(1)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
flops = 8.0 * 2.0 ** (n_m + n_n + n_k)
for r in range(reps):
    ms = plan.profile([A.data_ptr(), B.data_ptr()], [0, 0], 1, 0, out.data_ptr(), ws.data_ptr(), ws_bytes, stream)
    whole, pack = float(ms[0, 0]), float(ms[0, 1])
    if c128:
        print("c128 M=2^%d N=2^%d K=2^%d: %.4f ms -> %.2f FP64 TFLOP/s" % (n_m, n_n, n_k, whole, flops / whole / 1e9))
        continue
    print("M=2^%d N=2^%d K=2^%d chunk=%d: step %.4f ms (pack %.4f, gemm %.4f) -> %.1f algorithmic TFLOP/s, gemm alone "
          "%.1f TF32 TFLOP/s executed" % (n_m, n_n, n_k, chunk, whole, pack, whole - pack, flops / whole / 1e9,
                                          3 * flops / (whole - pack) / 1e9))
