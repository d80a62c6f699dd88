#!/usr/bin/env python
"""bench.py — circuit evaluations/s of the compiled-circuit hot path on N B200s (one rank per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4] [--impl b200|reference]

A *step* is one pass of the hot path over one batch of synthetic parameter sets: forward through the
whole circuit including measurements, and the backward (adjoint) pass for the gradient of sum(outputs)
w.r.t. every flat parameter.  An *evaluation* = one parameter set through one step.
Default workload = BASELINE.json configs[1]: 12-qubit 1-D many-body-localisation circuit, complex64,
batch 256 parameter sets per GPU (weak scaling: every rank runs its own 256 sets, no collective on the
data path; rank 0 reduces the timings).  Rank 0 prints ONE JSON line.

Timed region: per step a CUDA-event pair on the launching stream; between steps L2 is flushed by writing
a 256 MiB buffer (outside the event pair).  `value` has inputs resident in HBM; `e2e` goes through the
C-ABI host entry point (tq_execute_host) with HOST buffers: H2D of parameters and cotangent and D2H of
results and gradients are inside its timed region.  `--impl reference` times the CPU restatement of the
reference's own pytorch path (oracle/sv_ref.py: the reference is pure Python and cannot travel to the
box; same torch ops, same python loop) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def workload(name):
    """-> (spec, inputs [B, P] float np array, complex dtype string, description)"""
    from tedq_b200 import workloads as W

    rng = np.random.RandomState(0)
    if name == "c2":
        spec = W.mbl_1d(12)
        return spec, W.c2_inputs(256, 12, 0), "c64", "c2: 12-qubit MBL-1D (860 gates, 61 params), probs(q11), batch 256, fwd+bwd"
    if name == "c1":
        spec = W.qnn4()
        return spec, rng.uniform(0, 1, size=(64, spec["n_params"])).astype(np.float32), "c64", \
            "c1: 4-qubit QNN (26 gates, 20 params), 4 Z expvals, batch 64, fwd+bwd"
    if name == "c3":
        spec = W.hea(20, 10)
        b = int(os.environ.get("TQ_C3_BATCH", "128"))
        return spec, rng.uniform(0, 1, size=(b, spec["n_params"])).astype(np.float32), "c64", \
            f"c3: 20-qubit HEA depth 10 (630 gates, 440 params), 20 Z expvals, batch {b} per GPU, fwd+bwd"
    if name == "c4":
        spec = W.mbl_2d(4, 1)
        return spec, rng.uniform(0, 1, size=(64, spec["n_params"])).astype(np.float64), "c128", \
            "c4: 16-qubit MBL-2D 4x4 depth-1 (1835 gates, 99 params), probs(q15), complex128, batch 64, fwd+bwd"
    if name == "c4d20":
        spec = W.mbl_2d(4, 10)
        return spec, rng.uniform(0, 1, size=(64, spec["n_params"])).astype(np.float64), "c128", \
            (f"c4 at depth 20: 16-qubit MBL-2D 4x4, 10 Hd + 10 H0 Trotter sweeps ({len(spec['gates'])} gates, "
             f"{spec['n_params']} params), probs(q15), complex128, batch 64, fwd+bwd")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def time_cpu_reference(spec, flat, cdt, budget_s, max_sets):
    """Reference CPU path (oracle port): python loop over parameter sets, forward + backward each."""
    import tedq_b200 as qb
    from oracle import sv_ref
    from tedq_b200 import workloads as W

    rd = torch.float32 if cdt == "c64" else torch.float64
    cd = torch.complex64 if cdt == "c64" else torch.complex128
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    sv_ref.run_batch(circ, torch.tensor(flat[:1], dtype=rd), cd, torch.ones(()))  # warm-up (allocator, threads)
    outs = []
    n = 0
    t0 = time.perf_counter()
    while n < max_sets:
        xg = torch.tensor(flat[n], dtype=rd, requires_grad=True)
        y = sv_ref.run_sv(circ, xg, cd)
        (torch.view_as_real(y).sum() if y.is_complex() else y.sum()).backward()
        outs.append(y.detach())
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt, (torch.stack(outs) if outs else None)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c5":
        n_timed = int(os.environ.get("TQ_C5_CPU_SLICES", "4"))
        vals = []
        t_all = time.perf_counter()
        for _ in range(max(1, args.steps)):
            dt, amps, n_slices, fl_slice = c5_cpu_slices(n_timed)
            vals.append(1.0 / (dt * n_slices))
        total = time.perf_counter() - t_all
        value = float(np.mean(vals))
        print(json.dumps({
            "impl": "reference", "metric": "circuit_evals_per_sec", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c64", "data": "synthetic",
            "config": {"workload": "c5: 40-qubit 5x8 lattice random circuit, 12 cycles, amplitude <0|U|0>, "
                                   f"{n_slices} slices", "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n_timed} of {n_slices} slices per step with torch.tensordot complex64, "
                                       f"extrapolated x{n_slices}; host has {os.cpu_count()} logical cores"},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    spec, flat, cdt, desc = workload(args.workload)
    per_step_budget = float(os.environ.get("TQ_REF_STEP_SECONDS", "8"))
    rates, sets = [], 0
    for _ in range(args.warmup):
        time_cpu_reference(spec, flat, cdt, 1.0, 2)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, n, dt, _ = time_cpu_reference(spec, flat, cdt, per_step_budget, len(flat))
        rates.append(r)
        sets = n
    total = time.perf_counter() - t_all
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "circuit_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cdt, "data": "synthetic",
        "config": {"workload": desc, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sets} parameter sets per step, python loop fwd+bwd (the reference has no batch entry); "
                                   f"host has {os.cpu_count()} logical cores"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def measure_workload(name, steps, warmup, device, dist_on, world, do_e2e=True, do_cpu=True, kernel_timing=True):
    import tedq_b200 as qb
    from tedq_b200 import workloads as W

    spec, flat_np, cdt, desc = workload(name)
    rd = torch.float32 if cdt == "c64" else torch.float64
    cd = torch.complex64 if cdt == "c64" else torch.complex128
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    cc = circ.compilecircuit(backend="pytorch_b200", dtype=cd)
    plan = cc.plan()
    B, P = flat_np.shape
    flat = torch.tensor(flat_np, dtype=rd, device=device)
    out = torch.empty((B, plan.out_reals), dtype=rd, device=device)
    dy = torch.ones((B, plan.out_reals), dtype=rd, device=device)
    grad = torch.empty((B, P), dtype=rd, device=device)
    ws_bytes = plan.workspace_bytes(B, True)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream

    def step():
        plan.forward(flat.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws_bytes, True, stream)
        plan.backward(flat.data_ptr(), B, dy.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws_bytes, stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(device)
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    evf = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t_wall = time.perf_counter()
    for i in range(steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        plan.forward(flat.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws_bytes, True, stream)
        evf[i].record()
        plan.backward(flat.data_ptr(), B, dy.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        ev[i][1].record()
    torch.cuda.synchronize(device)
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - t_wall
    sampler.stop_flag = True
    sampler.join(timeout=2)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    fwd_ms = [a.elapsed_time(f) for (a, _), f in zip(ev, evf)]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if dist_on:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(t.item())
    value = B * steps * world / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel (the adjoint sweep): algorithmic HBM bytes / measured duration
    hbm_peak, peak_kind = load_peaks()
    bwd_ms = float(np.mean([s - f for s, f in zip(step_ms, fwd_ms)]))
    fwd_avg = float(np.mean(fwd_ms))
    bytes_b = plan.hbm_bytes(True) * B
    bytes_f = plan.hbm_bytes(False) * B
    n_sweeps_b = max(1, plan.num_sweeps(True))
    roof = {
        "bound": "hbm", "kernel": "k_sweep_bwd",
        "achieved": bytes_b / (bwd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": bytes_b / (bwd_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_kind": peak_kind,
        "algorithmic_bytes_per_step": bytes_b + bytes_f,
        "launch_ms": bwd_ms / n_sweeps_b, "sweeps_bwd": plan.num_sweeps(True), "sweeps_fwd": plan.num_sweeps(False),
        "fwd_ms": fwd_avg, "bwd_ms": bwd_ms,
        "fwd_achieved_gbs": bytes_f / (fwd_avg * 1e-3) / 1e9,
        "note": ("state resident in shared memory for the whole circuit: HBM sees parameters, outputs and gradients "
                 "only; the kernel is shared-memory/issue bound, see DESIGN.md" if plan.num_sweeps(True) == 1 and
                 spec["num_qubits"] <= 13 else "tiled sweeps: each sweep reads+writes psi and lambda once"),
    }
    res = {"value": value, "ms_per_step": total_ms / steps, "B": B, "P": P, "desc": desc, "roofline": roof,
           "clocks": sampler.summary(), "wall_s": wall,
           "gpu_launches": int((plan.launches(False) + plan.launches(True)) * steps), "dtype": cdt}

    # ---- e2e: HOST buffers through the C-ABI host entry point
    if do_e2e:
        hp = np.ascontiguousarray(flat_np, dtype=np.float32 if cdt == "c64" else np.float64)
        hdy = np.ones((B, plan.out_reals), dtype=hp.dtype)
        for _ in range(max(1, warmup)):
            plan.execute_host(hp, hdy)
        if dist_on:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            plan.execute_host(hp, hdy)
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if dist_on:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
        item = hp.dtype.itemsize
        res["e2e"] = {"value": B * steps * world / e2e_s, "unit": "evals/s",
                      "h2d_bytes_per_step": int(B * (P + plan.out_reals) * item),
                      "d2h_bytes_per_step": int(B * (P + plan.out_reals) * item),
                      "api": "tq_execute_host (C ABI, host buffers)"}

    # ---- CPU baseline: the oracle port on the host cores, bounded sample, rank 0 only, with a parity check
    if do_cpu:
        budget = float(os.environ.get("TQ_CPU_BASELINE_SECONDS", "12"))
        rate, n, dt, ref_out = time_cpu_reference(spec, flat_np, cdt, budget, B)
        got = out[:n].reshape(ref_out.shape if not ref_out.is_complex() else (n, -1)).cpu() if ref_out is not None else None
        err = None
        if ref_out is not None and not ref_out.is_complex():
            err = float((got.double() - ref_out.double()).abs().max())
        res["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{n} of {B} parameter sets, python loop fwd+bwd, {dt:.1f} s; "
                                         f"host has {os.cpu_count()} logical cores; max |gpu-cpu| on those sets = {err}"}
    return res


# Pre-searched plans of the BASELINE networks (the planner's on-disk cache, hyper_opt["plan_cache"]): searched once
# by scripts/make_bench_plans.py with exactly the options below and committed, so that a bench run spends its
# time on the device, not in the 12 s path search.  A missing or foreign file only means the search runs again.
PLAN_CACHE = os.path.join(ROOT, "ted-q_b200", "plans")
C5_HYPER = {"max_repeats": int(os.environ.get("TQ_C5_REPEATS", "64")),
            "reconf_sweeps": int(os.environ.get("TQ_C5_RECONF", "6")),   # subtree reconfiguration of the greedy tree
            # objective of the reconfiguration: estimated step time = max(flops / 200 TFLOP/s, bytes / 2.5 TB/s) +
            # 12 us per launched step (planner.step_time_model), instead of the bare flop count
            "time_model": None if os.environ.get("TQ_C5_TIME_MODEL", "1") == "0" else (2.0e14, 2.5e12, 1.2e-5),
            "slicing_opts": {"target_size": 2 ** 27, "target_num_slices": 64}}


def tf32_peak():
    """No TF32 figure in MEASURED_PEAKS.json: half the measured bf16 burst (tcgen05 kind::tf32 runs at half the
    kind::f16 rate), labelled derived."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["bf16_tflops"]) / 2.0, "derived: measured bf16 burst / 2"
    except Exception:
        return 1590.0 / 2.0, "derived from the fallback bf16 figure / 2"


def c5_cpu_slices(n_slices_timed, slice_ids=None):
    """Reference CPU arm of config 5: the same circuit, the same path and sliced indices, contracted slice by slice
    with torch.tensordot on the host cores (what tree.contract(arrays, backend='torch') does,
    pytorch_backend.py:339).  -> (seconds per slice, [slice amplitudes], n_slices, flops per slice)"""
    import tedq_b200 as qb
    from oracle import tn_ref
    from tedq_b200 import planner, workloads as W

    spec = W.lattice_rcs(5, 8, 12, seed=0, measure="state")
    circ = W.build_circuit(spec, qb)
    inputs, output = tn_ref.index_maps(circ)[0]
    arrays = tn_ref.operands(circ, torch.zeros(0), torch.complex64)[0]
    cap0 = np.array([1, 0], dtype=np.complex64)
    inputs = [list(t) for t in inputs] + [[ix] for ix in output]
    arrays = list(arrays) + [cap0] * 40

    def search():
        first = planner.find_path(inputs, [], repeats=C5_HYPER["max_repeats"], seed=0,
                                  reconf_sweeps=C5_HYPER["reconf_sweeps"], time_model=C5_HYPER["time_model"])
        return planner.slice_path(inputs, [], first, target_size_log2=27, target_num_slices=64,
                                  reconf_sweeps=min(3, C5_HYPER["reconf_sweeps"]), time_model=C5_HYPER["time_model"])

    # same key as TNExecutor._plan_key(): the engine's amplitude plan and this one share a cache file
    info = planner.cached_plan(PLAN_CACHE, inputs, [], search, max_repeats=C5_HYPER["max_repeats"], seed=0,
                               minimize="flops", reconf_sweeps=C5_HYPER["reconf_sweeps"], reconf_leaves=8,
                               time_model=C5_HYPER["time_model"], target_size=2 ** 27, target_num_slices=64)
    ids = list(slice_ids) if slice_ids is not None else list(range(n_slices_timed))
    tn_ref.contract_slice_torch(arrays, inputs, [], info.path, info.sliced, 0)   # warm-up (threads, allocator)
    amps = []
    t0 = time.perf_counter()
    for sid in ids:
        amps.append(complex(tn_ref.contract_slice_torch(arrays, inputs, [], info.path, info.sliced, sid)))
    dt = (time.perf_counter() - t0) / max(1, len(ids))
    return dt, amps, info.n_slices, 2.0 ** info.flops_log2


def measure_c5(steps, warmup, device, dist_on, world, do_cpu=False, greedy_plan=False):
    """BASELINE config 5: 40-qubit lattice random circuit (5x8, 12 cycles), single amplitude <0..0|U|0..0>, sliced
    contraction; slices are sharded over ranks and combined with one all-reduce (strong scaling)."""
    import tedq_b200 as qb
    from tedq_b200 import workloads as W

    spec = W.lattice_rcs(5, 8, 12, seed=0)
    circ = W.build_circuit(spec, qb)
    # greedy_plan: the plain 64-repeat greedy tree (no reconfiguration): 2.5x the flops in larger, squarer GEMMs —
    # the plan on which the tensor-core kernels are closest to their roofline
    hyper = {"max_repeats": C5_HYPER["max_repeats"], "reconf_sweeps": 0 if greedy_plan else C5_HYPER["reconf_sweeps"],
             "time_model": None if greedy_plan else C5_HYPER["time_model"],
             "slicing_opts": dict(C5_HYPER["slicing_opts"], contract_parallel=dist_on), "plan_cache": PLAN_CACHE}
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=hyper)
    bits = [0] * 40
    for _ in range(max(1, warmup)):
        amp = cc.amplitude(bits)
    torch.cuda.synchronize(device)
    plan = cc._tn._amplitude_plan()[2]
    if dist_on:
        torch.distributed.barrier()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        amp = cc.amplitude(bits)
    ev1.record()
    torch.cuda.synchronize(device)
    if dist_on:
        torch.distributed.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    if dist_on:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(t.item())
    info = cc._tn._amplitude_plan()[1]
    group = len(cc._tn.slice_members(0))          # slices that share one launch sequence (hyper_opt["slice_batch"])
    flops = 2.0 ** info.flops_log2 * info.n_slices
    peak, kind = tf32_peak()
    hbm_peak, _ = load_peaks()
    ach = flops * steps / (total_ms * 1e-3) / 1e12

    # per-step table of one slice (tq_tn_profile: CUDA events around every step)
    rows = cc._tn.amplitude_profile(torch.zeros((1, 0), device=device), bits, 0)
    per_slice = [r for r in rows if r["per_slice"]]
    slice_ms = sum(r["ms"] for r in per_slice)
    table = []
    tc_flops = tc_ms = tc_gemm_ms = 0.0
    for r in sorted(per_slice, key=lambda r: -r["ms"]):
        if r["ms"] < 0.01 * slice_ms:
            break
        if r["kernel"] == 4:   # a fused run of small steps is timed as one launch (on its first member)
            table.append({"kernel": "k_tn_fused (run of small steps)", "ms": round(r["ms"], 4)})
            continue
        sets = r["sets"] if r.get("per_set") else 1   # a grouped step runs `sets` slices' GEMMs in one launch
        fl = sets * 8.0 * 2.0 ** (r["k"] + r["m"] + r["n"] + r["b"])
        byts = sets * 8.0 * (2.0 ** (r["k"] + r["m"] + r["b"]) + 2.0 ** (r["k"] + r["n"] + r["b"]) + 2.0 ** (r["m"] + r["n"] + r["b"]))
        t_fl = 3.0 * fl / (peak * 1e12)          # 4M x 3-term split-TF32: 24 TF32 flops per 8 algorithmic
        t_by = byts / (hbm_peak * 1e9)
        roof_ms = max(t_fl, t_by) * 1e3
        table.append({"M": 2 ** r["m"], "N": 2 ** r["n"], "K": 2 ** r["k"], "batch": 2 ** r["b"] * sets,
                      "kernel": ["k_tn_step", "k_tn_gemm", "k_tc_pack+k_tc_gemm", "k_tn_dot", "k_tn_fused", "k_tn_apply"][r["kernel"]],
                      "ms": round(r["ms"], 4), "pack_ms": round(r["pack_ms"], 4),
                      "algorithmic_tflops": round(fl / (r["ms"] * 1e-3) / 1e12, 2),
                      "bound": "tensor" if t_fl > t_by else "hbm", "roofline_ms": round(roof_ms, 4),
                      "frac": round(roof_ms / r["ms"], 3)})
        if r["kernel"] == 2 and t_fl > t_by:
            tc_flops += fl
            tc_ms += r["ms"]
            tc_gemm_ms += r["ms"] - r["pack_ms"]
    res = {
        "value": steps / (total_ms * 1e-3), "unit": "evals/s", "ms_per_step": total_ms / steps, "dtype": "c64",
        "desc": f"c5: 40-qubit 5x8 lattice random circuit, 12 cycles, amplitude <0|U|0>, {info.n_slices} slices"
                + (f" in groups of {group} (one launch sequence per group)" if group > 1 else "") +
                f", width {plan.width}, {plan.n_steps} pairwise steps, {flops:.3e} flop per amplitude",
        "amplitude": [float(amp.real), float(amp.imag)],
        "roofline": {
            "bound": "tensor", "kernel": "k_tc_gemm", "unit": "TFLOP/s", "peak": peak, "peak_kind": kind,
            # whole contraction (all 760 steps x 64 slices, packing and launch gaps included), algorithmic 8MNK flops
            "achieved": ach, "frac": ach / peak, "traffic": None,
            # the tensor-core-bound dominant steps of one slice
            "dominant_steps_algorithmic_tflops": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
            "dominant_steps_tf32_executed_tflops": 3 * tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
            "dominant_steps_frac_of_complex_gemm_roofline": 3 * tc_flops / (tc_ms * 1e-3) / 1e12 / peak if tc_ms else None,
            "gemm_kernel_only_tf32_executed_tflops": 3 * tc_flops / (tc_gemm_ms * 1e-3) / 1e12 if tc_gemm_ms else None,
            "note": "achieved/frac count ALGORITHMIC flops (8*M*N*K per complex GEMM) over the whole amplitude against "
                    "the TF32 peak. A complex64 GEMM at fp32 accuracy on TF32 tensor cores executes 24*M*N*K TF32 flops "
                    "(4M real decomposition x 3-term error-compensated split), so the complex-GEMM tensor-core "
                    "roofline is peak/3: dominant_steps_frac_of_complex_gemm_roofline = executed TF32 flops / time "
                    "(operand packing included) / peak for the tensor-bound steps.",
        },
        "per_slice_ms_profiled": slice_ms / group, "slices_per_launch_sequence": group, "steps_per_slice": len(per_slice), "steps_once_per_call": len(rows) - len(per_slice),
        "step_table": table,
        "clocks": sampler.summary(),
        "gpu_launches": int(sum((3 if r["kernel"] == 2 else 2 if r["kernel"] == 3 else 1 if r["kernel"] < 4 else
                                 (1 if r["ms"] > 0 else 0)) for r in per_slice) * plan.n_slices * steps / max(1, world)),
        "scaling": "strong",
    }
    if do_cpu:
        n_cpu = int(os.environ.get("TQ_C5_CPU_SLICES", "4"))
        groups = max(1, (n_cpu + group - 1) // group)
        members = [cc._tn.slice_members(i) for i in range(groups)]
        dt, amps, n_slices, fl_slice = c5_cpu_slices(0, slice_ids=[sid for m in members for sid in m])
        got = [complex(cc.amplitude(bits, slice_range=(i, i + 1)).cpu()) for i in range(groups)]
        want = [sum(amps[i * group:(i + 1) * group]) for i in range(groups)]
        scale = max(abs(a) for a in amps) or 1.0     # a slice can be exactly zero (a sliced wire next to a |0> cap)
        err = max(abs(g - a) / scale for g, a in zip(got, want))
        res["cpu_baseline"] = {
            "value": 1.0 / (dt * n_slices), "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(amps)} of {n_slices} slices contracted with torch.tensordot complex64 on the host "
                      f"({dt:.2f} s per slice, {fl_slice / dt / 1e9:.0f} GFLOP/s), extrapolated x{n_slices}; host has "
                      f"{os.cpu_count()} logical cores; max |gpu-cpu| on those slice amplitudes relative to the largest = {err:.2e}"}
    return res


def measure_c5_simplified(steps, warmup, device):
    """Config 5 again with tn_simplify=True (the reference's default flag; its own simplifier does not work):
    CNOT controls and RZ gates stay on shared wire indices, the plan needs no slicing."""
    import tedq_b200 as qb
    from tedq_b200 import workloads as W

    spec = W.lattice_rcs(5, 8, 12, seed=0)
    cc = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=True,
                                                  hyper_opt={"max_repeats": 64, "plan_cache": PLAN_CACHE})
    bits = [0] * 40
    for _ in range(max(1, warmup)):
        amp = cc.amplitude(bits)
    torch.cuda.synchronize(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        amp = cc.amplitude(bits)
    ev1.record()
    torch.cuda.synchronize(device)
    ms = ev0.elapsed_time(ev1) / steps
    plan = cc._tn._amplitude_plan()[2]
    kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
    return {"value": 1e3 / ms, "unit": "evals/s", "ms_per_step": ms, "dtype": "c64",
            "amplitude": [float(amp.real), float(amp.imag)],
            "steps_by_kernel": {k: kinds.count(k) for k in sorted(set(kinds))},
            "workload": f"c5 with tn_simplify=True: same circuit and amplitude, diagonal / controlled gates on shared "
                        f"wire indices: {plan.n_slices} slice(s), width {plan.width}, "
                        f"{plan.flops * plan.n_slices:.3e} flop per amplitude (dense network: 1.9e13)"}


def measure_tn_mode(name, steps, warmup, device):
    """BASELINE configs 2 and 4 in the mode they name: tensor-network contraction (tn_mode=True) through the public
    API — values from the contraction plan (one network per measurement, batched gate operands), gradient from the
    adjoint sweeps (backend.B200Execute, B200Backend._vjp)."""
    import tedq_b200 as qb
    from tedq_b200 import workloads as W

    spec, flat_np, cdt, desc = workload(name)
    rd = torch.float32 if cdt == "c64" else torch.float64
    cd = torch.complex64 if cdt == "c64" else torch.complex128
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, dtype=cd,
                             hyper_opt={"max_repeats": 8})
    x = torch.tensor(flat_np, dtype=rd, device=device)
    nb = x.shape[0]

    def step():
        xx = x.clone().requires_grad_(True)
        y = cc.batched(xx)
        y.sum().backward()
        return y

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        y = step()
    ev1.record()
    torch.cuda.synchronize(device)
    ms = ev0.elapsed_time(ev1) / steps
    with torch.no_grad():
        cc.batched(x)
        torch.cuda.synchronize(device)
        ev0.record()
        for _ in range(steps):
            cc.batched(x)
        ev1.record()
        torch.cuda.synchronize(device)
    fwd_ms = ev0.elapsed_time(ev1) / steps
    plan = cc._tn._plan(0)
    kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
    # the same step with the gradient taken by reverse mode through the contraction tree (tq_tn_backward)
    cc_tree = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, dtype=cd,
                                  hyper_opt={"max_repeats": 8, "tn_backward": "tree"})

    def step_tree():
        xx = x.clone().requires_grad_(True)
        cc_tree.batched(xx).sum().backward()

    for _ in range(warmup):
        step_tree()
    torch.cuda.synchronize(device)
    ev0.record()
    for _ in range(steps):
        step_tree()
    ev1.record()
    torch.cuda.synchronize(device)
    tree_ms = ev0.elapsed_time(ev1) / steps
    return {"value": nb / (ms * 1e-3), "unit": "evals/s", "ms_per_step": ms, "fwd_only_ms": fwd_ms,
            "tree_backward_ms_per_step": tree_ms,
            "fwd_only_evals_per_s": nb / (fwd_ms * 1e-3),
            "steps_by_kernel": {k: kinds.count(k) for k in sorted(set(kinds))},
            "workload": f"{desc.split(',')[0]} in tensor-network mode, batch {nb}: fwd = contraction plan "
                        f"({plan.n_steps} pairwise steps, width {plan.width}, {plan.flops:.3e} flop per set, "
                        f"in-repo planner) + bwd = adjoint sweeps",
            "dtype": cdt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--extras", default=os.environ.get("TQ_BENCH_EXTRAS", "c2tn,c1,c3,c4,c4d20,c4tn,c5,c5g,c5s"),
                    help="other BASELINE configs measured briefly and reported inside the same JSON line")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist_on = world > 1
    if dist_on:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: ONE JSON line
        torch.distributed.init_process_group("nccl", device_id=device)

    if args.workload == "c5":
        r = measure_c5(max(1, min(args.steps, 5)), 1, device, dist_on, world, do_cpu=(rank == 0 and world == 1))
        if rank == 0:
            extra = {k: r[k] for k in ("cpu_baseline", "step_table", "per_slice_ms_profiled", "steps_per_slice",
                                       "steps_once_per_call") if k in r}
            print(json.dumps({**extra, 
                "metric": "circuit_evals_per_sec", "value": r["value"], "unit": r["unit"], "n_gpus": world,
                "steps": max(1, min(args.steps, 5)), "warmup": 1, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
                "config": {"workload": r["desc"], "parallelism": f"slices sharded over {world} GPU(s), one all-reduce",
                           "l2": "intermediates (2^23..2^27 complex) exceed L2 between steps"},
                "roofline": r["roofline"], "clocks": r["clocks"], "gpu_launches": r["gpu_launches"],
                "amplitude": r["amplitude"],
                "e2e": {"value": r["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
                        "note": "the circuit has no runtime inputs (fixed random angles); the timed call is the "
                                "public cc.amplitude(bits), only the 8-byte amplitude returns to the host"}}))
        if dist_on:
            torch.distributed.destroy_process_group()
        return
    main_res = measure_workload(args.workload, args.steps, args.warmup, device, dist_on, world,
                                do_e2e=True, do_cpu=(rank == 0 and world == 1))
    extras = {}
    if world == 1 and args.extras and args.extras != "none":
        for name in [e for e in args.extras.split(",") if e and e != args.workload]:
            if name == "c5":
                r = measure_c5(3, 1, device, False, 1, do_cpu=os.environ.get("TQ_EXTRAS_CPU", "1") == "1")
                extras[name] = {"value": r["value"], "unit": "evals/s", "ms_per_step": r["ms_per_step"],
                                "workload": r["desc"], "roofline": r["roofline"], "dtype": "c64",
                                "step_table": r["step_table"], "per_slice_ms_profiled": r["per_slice_ms_profiled"],
                                "steps_per_slice": r["steps_per_slice"], "amplitude": r["amplitude"]}
                if "cpu_baseline" in r:
                    extras[name]["cpu_baseline"] = r["cpu_baseline"]
                continue
            if name == "c5g":
                r = measure_c5(3, 1, device, False, 1, do_cpu=False, greedy_plan=True)
                extras[name] = {"value": r["value"], "unit": "evals/s", "ms_per_step": r["ms_per_step"],
                                "workload": r["desc"] + " (plain greedy tree, no subtree reconfiguration)",
                                "roofline": r["roofline"], "dtype": "c64", "step_table": r["step_table"][:8],
                                "per_slice_ms_profiled": r["per_slice_ms_profiled"], "amplitude": r["amplitude"]}
                continue
            if name == "c5s":
                extras[name] = measure_c5_simplified(20, 3, device)
                continue
            if name in ("c2tn", "c4tn"):
                extras[name] = measure_tn_mode(name[:2], 5, 3, device)
                continue
            r = measure_workload(name, max(2, min(5, args.steps)), 3, device, False, 1, do_e2e=False,
                                 do_cpu=os.environ.get("TQ_EXTRAS_CPU", "0") == "1")
            extras[name] = {"value": r["value"], "unit": "evals/s", "ms_per_step": r["ms_per_step"], "workload": r["desc"],
                            "roofline": r["roofline"], "dtype": r["dtype"]}
            if "cpu_baseline" in r:
                extras[name]["cpu_baseline"] = r["cpu_baseline"]

    if rank == 0:
        # TeD-Q tensor-network mode, CPU, n=12: 2000 evals / 370 s = 5.4 evals/s (draw_comparison.ipynb:67)
        published = 2000.0 / 370.0 if args.workload == "c2" else None
        line = {
            "metric": "circuit_evals_per_sec", "value": main_res["value"], "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (main_res["value"] / published) if published else None,
            "dtype": main_res["dtype"], "data": "synthetic",
            "config": {"workload": main_res["desc"], "sets_per_gpu": main_res["B"], "params_per_set": main_res["P"],
                       "l2": "flushed between timed iterations (256 MiB write)", "parallelism": f"dp{world} (sets sharded)",
                       "published_baseline": "TeD-Q TN-mode CPU n=12: 5.4 evals/s, hardware unstated (BASELINE.md A2)"},
            "roofline": main_res["roofline"], "clocks": main_res["clocks"], "e2e": main_res.get("e2e"),
            "gpu_launches": main_res["gpu_launches"],
        }
        if "cpu_baseline" in main_res:
            line["cpu_baseline"] = main_res["cpu_baseline"]
        if extras:
            line["other_configs"] = extras
        print(json.dumps(line))
    if dist_on:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
