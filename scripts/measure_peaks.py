#!/usr/bin/env python
"""Measure the tensor-core peaks bench.py divides by (MEASURED_PEAKS.json only has HBM and bf16):

  * TF32 dense: cuBLAS sgemm with TF32 tensor-core math (torch.matmul fp32, allow_tf32), 8192^3
  * FP64 dense: cuBLAS dgemm 8192^3 (DMMA)
  * bf16 dense and the HBM copy again, so that the ratios to the driver's numbers can be checked

each as burst (best of 10) and sustained (back to back for ~4 s), with the nvidia-smi clocks sampled during
the sustained part.  Writes gpurun_out/r02_peaks.json (copy it to profiles/).

    gpurun -- python scripts/measure_peaks.py
"""
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Clocks(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                self.rows.append([p.strip() for p in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        return {"samples": len(self.rows), "sm_mhz_median": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "power_w_max": max((float(r[2]) for r in self.rows if len(r) > 2), default=None),
                "sw_power_cap": any(len(r) > 3 and r[3].lower().startswith("active") for r in self.rows),
                "hw_slowdown": any(len(r) > 4 and r[4].lower().startswith("active") for r in self.rows)}


def gemm_peak(dtype, n, tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    flop = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    # sustained: back to back for ~4 s
    clk = Clocks()
    clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, int(4.0 / (flop / (best * 1e12))))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    clk.stop = True
    clk.join(timeout=2)
    sustained = flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    torch.backends.cuda.matmul.allow_tf32 = False
    return {"burst_tflops": best, "sustained_tflops": sustained, "n": n, "reps_sustained": reps, "clocks": clk.summary()}


def copy_peak():
    n = 1 << 30
    a = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    b = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    a.fill_(1.0)
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def main():
    assert torch.cuda.is_available()
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "how": "torch.matmul 8192^3 (2*N^3 flop): fp32 with allow_tf32 (cuBLAS TF32 tensor-core sgemm), fp64 (cuBLAS dgemm), "
                  "bf16; best of 10 (burst) and back to back for ~4 s (sustained); copy: b.copy_(a) over 1 Gi bf16, read+write bytes"}
    res["hbm_gbs"] = copy_peak()
    res["bf16"] = gemm_peak(torch.bfloat16, 8192)
    res["tf32"] = gemm_peak(torch.float32, 8192, tf32=True)
    res["fp32_simt"] = gemm_peak(torch.float32, 8192, tf32=False)
    res["fp64"] = gemm_peak(torch.float64, 8192)
    out = os.path.join(ROOT, "gpurun_out", "r02_peaks.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    sys.exit(main())
