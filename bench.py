#!/usr/bin/env python
"""bench.py — circuit evaluations/s of the compiled-circuit hot path on N B200s (one rank per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c1|c2|c3|c4|c4d20] [--impl b200|reference]

Headline (default) = BASELINE.json configs[4], the north-star workload and the only one with a collective on its data
path: single amplitudes <b|U|0> of a 40-qubit random circuit (5x8 lattice, 12 cycles) by SLICED tensor-network
contraction.  A *step* is one pass of the hot path over one batch of synthetic input: AMPS (96) random bitstrings b;
an *evaluation* is one amplitude.  STRONG scaling: the 64 slices of every amplitude are sharded over the ranks
(contiguous ranges of slice groups), the partial sums of the whole batch are combined with ONE NCCL all-reduce per
step, inside the timed region.  `value` = amplitudes/s, device-timed (CUDA events on the launching stream, max
over ranks); `e2e` = the same through the public API cc.amplitudes(host bitstrings) -> host result, wall clock,
host<->device copies inside.  `roofline` = algorithmic 8*M*N*K flops of the tensor-bound steps / their measured time
(operand packing included) against the complex-GEMM tensor-core roofline = measured TF32 peak / 3 (4M decomposition x
3-term error-compensated split = 24 TF32 flops per complex MAC).

`other_configs` (every N): the other BASELINE configs, each with value / roofline / e2e and, at N = 1, cpu_baseline:
C1 (4-qubit QNN), C2 (12-qubit MBL-1D: state-vector mode and the tensor-network mode BASELINE names), C3 (20-qubit
HEA, 1024 parameter sets in total sharded over the ranks), C4 (4x4 MBL-2D complex128: depth 1, "depth 20", and
tensor-network mode), C3 in tensor-network mode with its 20 measurement networks dealt to the ranks (one all-reduce),
C5 on the plain greedy plan and with tn_simplify=True.

Parity is ASSERTED before any time is printed: every config's GPU results are compared with the CPU oracle (the
restatement of the reference's pytorch path) on the sampled sets / slices; a mismatch aborts the run.
`--impl reference` times that CPU path on the host cores (all threads), same config dict, bounded sample per step.
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

# bitstrings per step of the headline: a multiple of the 16 amplitudes one launch sequence holds on 8 ranks, and enough
# that 20 steps stay above a second of device time at N = 8 (96 amplitudes: 0.54 s per step on one B200, 78 ms on eight)
AMPS = int(os.environ.get("TQ_C5_AMPS", "96"))


# ----------------------------------------------------------------------------------------------------------
# peaks
# ----------------------------------------------------------------------------------------------------------
def load_peaks():
    """HBM GB/s from MEASURED_PEAKS.json (driver-written); TF32 / FP64 / FP32 TFLOP/s from profiles/r02_peaks.json
    (scripts/measure_peaks.py: cuBLAS sgemm-TF32 / dgemm / sgemm 8192^3 on this pool's B200)."""
    p = {"hbm_gbs": 6650.0, "hbm_kind": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        p["hbm_gbs"], p["hbm_kind"] = float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        p["bf16"] = float(d["bf16_tflops"])
    except Exception:
        p["bf16"] = 1590.0
    try:
        with open(os.path.join(ROOT, "profiles", "r02_peaks.json")) as fh:
            d = json.load(fh)
        p["tf32"], p["tf32_sustained"] = float(d["tf32"]["burst_tflops"]), float(d["tf32"]["sustained_tflops"])
        p["fp64"], p["fp32"] = float(d["fp64"]["burst_tflops"]), float(d["fp32_simt"]["burst_tflops"])
        p["tc_kind"] = "measured: cuBLAS 8192^3 burst (profiles/r02_peaks.json)"
    except Exception:
        p["tf32"] = p["tf32_sustained"] = p["bf16"] / 2.0
        p["fp64"], p["fp32"] = 37.0, 74.0
        p["tc_kind"] = "derived: bf16 / 2 (profiles/r02_peaks.json missing)"
    return p


# ----------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------
C3_TOTAL_SETS = int(os.environ.get("TQ_C3_SETS", "1024"))


def workload(name, world=1, rank=None):
    """-> (spec, inputs [B, P] float np array of THIS rank's share, complex dtype string, description, scaling)"""
    from tedq_b200 import workloads as W

    rng = np.random.RandomState(0)
    if name == "c2":
        spec = W.mbl_1d(12)
        return spec, W.c2_inputs(256, 12, 0), "c64", \
            "c2: 12-qubit MBL-1D (860 gates, 61 params), probs(q11), batch 256 per GPU, fwd+bwd", "weak"
    if name == "c1":
        spec = W.qnn4()
        return spec, rng.uniform(0, 1, size=(64, spec["n_params"])).astype(np.float32), "c64", \
            "c1: 4-qubit QNN (26 gates, 20 params), 4 Z expvals, batch 64 per GPU, fwd+bwd", "weak"
    if name == "c3":
        spec = W.hea(20, 10)
        full = rng.uniform(0, 1, size=(C3_TOTAL_SETS, spec["n_params"])).astype(np.float32)
        rank = (int(os.environ.get("RANK", "0")) if world > 1 else 0) if rank is None else rank
        per = (C3_TOTAL_SETS + world - 1) // world
        return spec, full[rank * per:(rank + 1) * per], "c64", \
            (f"c3: 20-qubit HEA depth 10 (630 gates, 440 params), 20 Z expvals, {C3_TOTAL_SETS} parameter sets in "
             f"total sharded over the GPUs, fwd+bwd"), "strong"
    if name == "c4":
        spec = W.mbl_2d(4, 1)
        return spec, rng.uniform(0, 1, size=(64, spec["n_params"])).astype(np.float64), "c128", \
            "c4: 16-qubit MBL-2D 4x4 depth-1 (1835 gates, 99 params), probs(q15), complex128, batch 64 per GPU, fwd+bwd", \
            "weak"
    if name == "c4d20":
        spec = W.mbl_2d(4, 10)
        return spec, rng.uniform(0, 1, size=(64, spec["n_params"])).astype(np.float64), "c128", \
            (f"c4 at depth 20: 16-qubit MBL-2D 4x4, 10 Hd + 10 H0 Trotter sweeps ({len(spec['gates'])} gates, "
             f"{spec['n_params']} params), probs(q15), complex128, batch 64 per GPU, fwd+bwd"), "weak"
    raise SystemExit(f"unknown workload {name}")


C5_DESC = (f"c5: 40-qubit 5x8 lattice random circuit, 12 cycles, single amplitudes <b|U|0> by sliced tensor-network "
           f"contraction, 64 slices sharded over the GPUs, {AMPS} random bitstrings per step")


def config_of(name, desc):
    """The `config` dict, identical in both arms (the driver compares them)."""
    if name == "c5":
        return {"workload": desc, "amplitudes_per_step": AMPS, "slices": 64, "scaling": "strong",
                "l2": "GPU arm: the intermediates of a launch sequence (128 slice x amplitude sets of up to 2^21 "
                      "complex64 each, 2 GB per step of the plan) exceed the 126 MB L2 between steps, no flush needed; "
                      "CPU arm: n/a"}
    return {"workload": desc, "l2": "GPU arm: L2 flushed between timed iterations (256 MiB write); CPU arm: n/a"}


def c5_bitstrings(n_amps, seed=0):
    rng = np.random.RandomState(1000 + seed)
    bits = rng.randint(0, 2, size=(n_amps, 40)).astype(np.int64)
    bits[0] = 0           # the first amplitude is <0...0|U|0...0> (round-1 records quote it)
    return bits


# ----------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------
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
            time.sleep(0.1)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


class Dist:
    def __init__(self, device):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.on = self.world > 1
        self.device = device

    def barrier(self):
        torch.cuda.synchronize(self.device)
        if self.on:
            torch.distributed.barrier()
        torch.cuda.synchronize(self.device)

    def max(self, v):
        if not self.on:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        if not self.on:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device)
        torch.distributed.all_reduce(t)
        return float(t.item())


class ParityError(SystemExit):
    pass


def assert_parity(what, err, tol):
    if not (err <= tol):
        raise ParityError(f"bench.py: PARITY FAILURE in {what}: error {err:.3e} exceeds {tol:g}; no timing is reported")


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's pytorch path on the host cores
# ----------------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def time_cpu_reference(name, spec, flat, cdt, budget_s, max_sets):
    """Reference CPU path (oracle port): python loop over parameter sets, forward + backward each (the reference has
    no batch entry).  -> dict(rate, n, seconds, outputs, grads); cached per workload name."""
    if name in _CPU_CACHE:
        return _CPU_CACHE[name]
    import tedq_b200 as qb
    from oracle import sv_ref
    from tedq_b200 import workloads as W

    rd = torch.float32 if cdt == "c64" else torch.float64
    cd = torch.complex64 if cdt == "c64" else torch.complex128
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    if spec["num_qubits"] <= 16 and len(spec["gates"]) < 5000:
        sv_ref.run_batch(circ, torch.tensor(flat[:1], dtype=rd), cd, torch.ones(()))  # warm-up (allocator, threads)
    outs, grads = [], []
    n = 0
    t0 = time.perf_counter()
    while n < max_sets:
        xg = torch.tensor(flat[n], dtype=rd, requires_grad=True)
        y = sv_ref.run_sv(circ, xg, cd)
        (torch.view_as_real(y).sum() if y.is_complex() else y.sum()).backward()
        outs.append(y.detach())
        grads.append(xg.grad.detach().clone())
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    res = {"rate": n / dt, "n": n, "seconds": dt, "out": torch.stack(outs), "grad": torch.stack(grads)}
    _CPU_CACHE[name] = res
    return res


def cpu_baseline_entry(cpu, B, note=""):
    return {"value": cpu["rate"], "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{cpu['n']} of {B} parameter sets, python loop fwd+bwd (oracle/sv_ref.py = the reference's "
                      f"pytorch path), {cpu['seconds']:.1f} s; host has {os.cpu_count()} logical cores{note}"}


def check_against_cpu(what, cpu, out_gpu, grad_gpu, cdt):
    """max |gpu - cpu| over the sampled sets, against tol * max(1, |ref|) (values) and 4x that (gradients)."""
    tol = 1e-5 if cdt == "c64" else 1e-11
    n = min(cpu["n"], out_gpu.shape[0])
    ref = cpu["out"][:n].reshape(n, -1)
    got = out_gpu[:n].reshape(n, -1).cpu()
    if ref.is_complex():
        ref, got = torch.view_as_real(ref.contiguous()), torch.view_as_real(got.contiguous().to(ref.dtype))
    err_o = float(((got.double() - ref.double()).abs() / ref.double().abs().clamp(min=1.0)).max())
    assert_parity(f"{what} outputs", err_o, tol)
    err_g = None
    if grad_gpu is not None:
        rg = cpu["grad"][:n].double()
        err_g = float(((grad_gpu[:n].cpu().double() - rg).abs() / rg.abs().clamp(min=1.0)).max())
        assert_parity(f"{what} gradients", err_g, 4 * tol)
    return {"sets": n, "max_err_out": err_o, "max_err_grad": err_g, "tolerance": tol, "asserted": True}


# Pre-searched plans of the BASELINE networks (the planner's on-disk cache, hyper_opt["plan_cache"]): searched once
# by scripts/make_bench_plans.py with exactly the options below and committed, so that a bench run spends its
# time on the device, not in the path search.  A missing or foreign file only means the search runs again.
PLAN_CACHE = os.path.join(ROOT, "ted-q_b200", "plans")
C5_HYPER = {"max_repeats": int(os.environ.get("TQ_C5_REPEATS", "128")),
            "max_repeats_greedy": 64,                                    # the plain greedy plan reported beside it (c5g)
            "reconf_sweeps": int(os.environ.get("TQ_C5_RECONF", "10")),  # subtree reconfiguration of the greedy tree
            "reconf_leaves": int(os.environ.get("TQ_C5_LEAVES", "9")),
            # independent searches (seeds 0 .. restarts-1); the calibrated model below picks the plan to run
            "restarts": int(os.environ.get("TQ_C5_RESTARTS", "8")),
            # objective of the reconfiguration: estimated step time on this engine (planner.step_time_model with the
            # calibrated five-value model: tensor-core steps max(flops / 200 TFLOP/s, bytes / 2.5 TB/s), steps the
            # dispatch rule keeps off the tensor cores at the FP32 GEMM / per-element rates, 1.5 us per launched step
            # = a 32-slice launch sequence's share of a launch) instead of the bare flop count
            "time_model": None if os.environ.get("TQ_C5_TIME_MODEL", "1") == "0" else (2.0e14, 2.5e12, 1.5e-6, 2.5e13, 1.5e12),
            "slicing_opts": {"target_size": 2 ** 27, "target_num_slices": 64}}


def c5_cpu_slices(n_slices_timed, slice_ids=None, bits=None, greedy_plan=False, dtype=torch.complex64):
    """Reference CPU arm of config 5: the same circuit, the same path and sliced indices, contracted slice by slice
    with torch.tensordot on the host cores (what tree.contract(arrays, backend='torch') does,
    pytorch_backend.py:339).  -> (seconds per slice, [slice amplitudes], n_slices, flops per slice)"""
    import tedq_b200 as qb
    from oracle import tn_ref
    from tedq_b200 import planner, workloads as W

    spec = W.lattice_rcs(5, 8, 12, seed=0, measure="state")
    circ = W.build_circuit(spec, qb)
    inputs, output = tn_ref.index_maps(circ)[0]
    arrays = tn_ref.operands(circ, torch.zeros(0), dtype)[0]
    caps = [np.array([1, 0], dtype=np.complex128), np.array([0, 1], dtype=np.complex128)]
    bits = [0] * 40 if bits is None else [int(b) for b in bits]
    inputs = [list(t) for t in inputs] + [[ix] for ix in output]
    arrays = list(arrays) + [caps[b] for b in bits]

    reconf = 0 if greedy_plan else C5_HYPER["reconf_sweeps"]
    leaves = 8 if greedy_plan else C5_HYPER["reconf_leaves"]
    tmodel = None if greedy_plan else C5_HYPER["time_model"]
    repeats = C5_HYPER["max_repeats_greedy"] if greedy_plan else C5_HYPER["max_repeats"]

    # same key and search as TNExecutor._plan_key() / _search(): the engine's amplitude plan and this one share a cache file
    key = dict(max_repeats=repeats, seed=0, minimize="flops", reconf_sweeps=reconf, reconf_leaves=leaves, time_model=tmodel,
               target_size=2 ** 27, target_num_slices=64)
    if not greedy_plan and C5_HYPER["restarts"] > 1:
        key["restarts"] = C5_HYPER["restarts"]
    info = planner.cached_plan(PLAN_CACHE, inputs, [], lambda: planner.search_plan(inputs, [], **key), **key)
    ids = list(slice_ids) if slice_ids is not None else list(range(n_slices_timed))
    tn_ref.contract_slice_torch(arrays, inputs, [], info.path, info.sliced, 0, dtype)   # warm-up (threads, allocator)
    amps = []
    t0 = time.perf_counter()
    for sid in ids:
        amps.append(complex(tn_ref.contract_slice_torch(arrays, inputs, [], info.path, info.sliced, sid, dtype)))
    dt = (time.perf_counter() - t0) / max(1, len(ids))
    return dt, amps, info.n_slices, 2.0 ** info.flops_log2


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c5":
        n_timed = int(os.environ.get("TQ_C5_CPU_SLICES", "4"))
        for _ in range(min(1, args.warmup)):
            c5_cpu_slices(1)
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
            "config": config_of("c5", C5_DESC),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"each step: {n_timed} of {n_slices} slices of ONE amplitude contracted with "
                                       f"torch.tensordot complex64 on the host ({fl_slice / dt / 1e9:.0f} GFLOP/s), "
                                       f"extrapolated x{n_slices} slices; host has {os.cpu_count()} logical cores"},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    spec, flat, cdt, desc, scaling = workload(args.workload, 1)
    per_step_budget = float(os.environ.get("TQ_REF_STEP_SECONDS", "8"))
    rates, sets = [], 0
    t_all = time.perf_counter()
    for i in range(args.steps):
        _CPU_CACHE.clear()
        r = time_cpu_reference(args.workload, spec, flat, cdt, per_step_budget, len(flat))
        rates.append(r["rate"])
        sets = r["n"]
    total = time.perf_counter() - t_all
    value = float(np.mean(rates))
    print(json.dumps({
        "impl": "reference", "metric": "circuit_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": cdt, "data": "synthetic",
        "config": config_of(args.workload, desc),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sets} parameter sets per step, python loop fwd+bwd (the reference has no batch "
                                   f"entry); host has {os.cpu_count()} logical cores"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------------------------
# state-vector configs (C1, C2, C3, C4, C4 at depth 20)
# ----------------------------------------------------------------------------------------------------------
def measure_sv(name, steps, warmup, device, dist, do_cpu, cpu_budget=10.0, with_clocks=False):
    import tedq_b200 as qb
    from tedq_b200 import workloads as W

    pk = load_peaks()
    spec, flat_np, cdt, desc, scaling = workload(name, dist.world)
    rd = torch.float32 if cdt == "c64" else torch.float64
    cd = torch.complex64 if cdt == "c64" else torch.complex128
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    cc = circ.compilecircuit(backend="pytorch_b200", dtype=cd)
    plan = cc.plan(device)
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

    step()
    torch.cuda.synchronize(device)
    res = {"dtype": cdt, "workload": desc, "scaling": scaling, "sets_per_gpu": B}
    # ---- parity gate (rank 0, N = 1: the CPU oracle on a bounded sample) BEFORE any timing
    cpu = None
    if do_cpu:
        cpu = time_cpu_reference(name, spec, flat_np, cdt, cpu_budget, B)
        res["parity"] = check_against_cpu(name, cpu, out.reshape(B, -1), grad, cdt)
        res["cpu_baseline"] = cpu_baseline_entry(cpu, B)
    for _ in range(warmup):
        step()
    dist.barrier()
    sampler = ClockSampler(device.index or 0) if with_clocks else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    evf = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        plan.forward(flat.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws_bytes, True, stream)
        evf[i].record()
        plan.backward(flat.data_ptr(), B, dy.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        ev[i][1].record()
    dist.barrier()
    if sampler:
        res["clocks"] = sampler.finish()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    fwd_ms = [a.elapsed_time(f) for (a, _), f in zip(ev, evf)]
    total_ms = dist.max(float(sum(step_ms)))
    sets_all = dist.sum(B)
    res["value"] = sets_all * steps / (total_ms * 1e-3)
    res["unit"] = "evals/s"
    res["ms_per_step"] = total_ms / steps
    res["steps"] = steps

    # ---- roofline: the adjoint sweep (dominant kernel) against HBM (algorithmic bytes of the tiled schedule) and
    # against the FP32 / FP64 pipe (algorithmic flops of the fused-block schedule): the slower bound is the roofline
    bwd_ms = float(np.mean([s - f for s, f in zip(step_ms, fwd_ms)]))
    fwd_avg = float(np.mean(fwd_ms))
    bytes_b, bytes_f = plan.hbm_bytes(True) * B, plan.hbm_bytes(False) * B
    fl_b, fl_f = plan.flops(True) * B, plan.flops(False) * B
    pipe_peak = pk["fp32"] if cdt == "c64" else pk["fp64"]
    t_hbm = bytes_b / (pk["hbm_gbs"] * 1e9)
    t_pipe = fl_b / (pipe_peak * 1e12)
    resident = plan.num_sweeps(True) == 1 and spec["num_qubits"] <= 13
    bound = "hbm" if t_hbm >= t_pipe else ("fp32-pipe" if cdt == "c64" else "fp64-pipe")
    ach_gbs = bytes_b / (bwd_ms * 1e-3) / 1e9
    ach_tf = fl_b / (bwd_ms * 1e-3) / 1e12
    res["roofline"] = {
        "kernel": "k_rg_bwd" if plan.num_register_groups(True) else "k_sweep_bwd", "bound": bound,
        "achieved": ach_gbs if bound == "hbm" else ach_tf,
        "peak": pk["hbm_gbs"] if bound == "hbm" else pipe_peak,
        "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
        "frac": max(t_hbm, t_pipe) / (bwd_ms * 1e-3),
        "traffic": None,
        "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": pk["hbm_gbs"], "frac": ach_gbs / pk["hbm_gbs"],
                "algorithmic_bytes_per_launch_sequence": bytes_b, "peak_kind": pk["hbm_kind"]},
        "pipe": {"achieved_tflops": ach_tf, "peak_tflops": pipe_peak, "frac": ach_tf / pipe_peak,
                 "algorithmic_flops": fl_b, "peak_kind": "measured: cuBLAS SIMT sgemm / dgemm 8192^3 (profiles/r02_peaks.json)"},
        "fwd": {"ms": fwd_avg, "achieved_gbs": bytes_f / (fwd_avg * 1e-3) / 1e9,
                "achieved_tflops": fl_f / (fwd_avg * 1e-3) / 1e12},
        "bwd_ms": bwd_ms, "sweeps_bwd": plan.num_sweeps(True), "sweeps_fwd": plan.num_sweeps(False),
        "launch_ms": bwd_ms / max(1, plan.num_sweeps(True)),
        "register_groups": {"fwd": plan.num_register_groups(False), "bwd": plan.num_register_groups(True)},
        "note": ("state resident in shared memory for the whole circuit: HBM sees parameters, outputs and gradients only"
                 if resident else "tiled sweeps: each sweep reads+writes psi (and lambda) once"),
    }
    res["gpu_launches"] = int((plan.launches(False) + plan.launches(True)) * steps)

    # ---- e2e (1): HOST buffers through the C-ABI host entry point (H2D, kernels, D2H inside the call)
    hp = np.ascontiguousarray(flat_np, dtype=np.float32 if cdt == "c64" else np.float64)
    hdy = np.ones((B, plan.out_reals), dtype=hp.dtype)
    n_e2e = max(2, min(steps, 10))
    for _ in range(2):
        plan.execute_host(hp, hdy)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        plan.execute_host(hp, hdy)
    e2e_s = dist.max(time.perf_counter() - t0)
    item = hp.dtype.itemsize
    res["e2e"] = {"value": sets_all * n_e2e / e2e_s, "unit": "evals/s", "steps": n_e2e,
                  "h2d_bytes_per_step": int(B * (P + plan.out_reals) * item),
                  "d2h_bytes_per_step": int(B * (P + plan.out_reals) * item),
                  "api": "tq_execute_host (C ABI, host buffers)"}
    # ---- e2e (2): the call a TeD-Q user makes: cc.batched(pinned host -> cuda) ... .backward(), results to the host
    hpin = torch.from_numpy(hp).pin_memory()

    def py_step():
        x = hpin.to(device, non_blocking=True).requires_grad_(True)
        y = cc.batched(x)
        (torch.view_as_real(y).sum() if y.is_complex() else y.sum()).backward()
        return y.detach().cpu(), x.grad.cpu()

    for _ in range(2):
        py_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        yh, gh = py_step()
    py_s = dist.max(time.perf_counter() - t0)
    res["e2e_python"] = {"value": sets_all * n_e2e / py_s, "unit": "evals/s", "steps": n_e2e,
                         "h2d_bytes_per_step": int(B * P * item), "d2h_bytes_per_step": int(B * (P + plan.out_reals) * item),
                         "api": "cc.batched(x_cuda).sum().backward() (autograd through B200Execute), pinned host in, host out"}
    if cpu is not None:     # the Python API's numbers are the same numbers
        check_against_cpu(name + " (python API)", cpu, yh.reshape(B, -1), gh, cdt)
    del ws, flush
    return res


# ----------------------------------------------------------------------------------------------------------
# tensor-network mode of C2 / C4 (the mode BASELINE.json names) and measurement-sharded C3
# ----------------------------------------------------------------------------------------------------------
def measure_tn_mode(name, steps, warmup, device, dist, do_cpu, cpu_budget=10.0, hyper=None, measurement_parallel=False,
                    batch=None):
    """Values from the contraction plans (one network per measurement, batched gate operands), gradient from the
    adjoint sweeps (backend.B200Execute) or the contraction tree (hyper_opt tn_backward), through the public API."""
    import tedq_b200 as qb
    from tedq_b200 import workloads as W

    pk = load_peaks()
    spec, flat_np, cdt, desc, scaling = workload(name, 1 if measurement_parallel else dist.world)
    if batch:
        flat_np = flat_np[:batch]
    rd = torch.float32 if cdt == "c64" else torch.float64
    cd = torch.complex64 if cdt == "c64" else torch.complex128
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    ho = dict(hyper or {"max_repeats": 8})
    if measurement_parallel:
        ho["measurement_parallel"] = dist.on
    t_plan = time.perf_counter()
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, dtype=cd, hyper_opt=ho)
    plan_s = time.perf_counter() - t_plan
    x = torch.tensor(flat_np, dtype=rd, device=device)
    nb, P = x.shape
    L = __import__("tedq_b200").capi.lib()

    def step():
        xx = x.clone().requires_grad_(True)
        y = cc.batched(xx)
        y.sum().backward()
        return y, xx.grad

    y, g = step()
    torch.cuda.synchronize(device)
    res = {"dtype": cdt, "scaling": "strong" if measurement_parallel else scaling, "sets_per_gpu": nb,
           "planner_search_s": round(plan_s, 2)}
    cpu = None
    if do_cpu:
        cpu = time_cpu_reference(name, spec, flat_np, cdt, cpu_budget, nb)
        res["parity"] = check_against_cpu(name + " tensor-network mode", cpu, y.detach().reshape(nb, -1), g, cdt)
        res["cpu_baseline"] = cpu_baseline_entry(
            cpu, nb, "; the reference's own tensor-network branch needs cotengra / jdtensorpath (absent offline): its "
                     "state-vector branch on the same circuit is the CPU arm")
    for _ in range(warmup):
        step()
    dist.barrier()
    l0 = int(L.tq_tn_launch_count())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    dist.barrier()
    l1 = int(L.tq_tn_launch_count())
    ms = dist.max(ev0.elapsed_time(ev1)) / steps
    with torch.no_grad():
        cc.batched(x)
        dist.barrier()
        ev0.record()
        for _ in range(steps):
            cc.batched(x)
        ev1.record()
        dist.barrier()
    fwd_ms = dist.max(ev0.elapsed_time(ev1)) / steps
    sets_all = nb if measurement_parallel else dist.sum(nb)
    plans = [cc._tn._plan(i, device) for i in range(len(cc._tn.networks))]
    kinds = [p.step_kernel(s) for p in plans for s in range(p.n_steps)]
    names = {0: "k_tn_step", 1: "k_tn_gemm" if cdt == "c64" else "k_tn_gemm_dmma", 2: "k_tc_pack+k_tc_gemm",
             3: "k_tn_dot", 4: "k_tn_fused", 5: "k_tn_apply", 6: "k_tn_seed", 7: "k_tn_chain"}
    flops = sum(p.flops for p in plans) * nb
    pipe_peak = pk["fp32"] if cdt == "c64" else pk["fp64"]
    # bytes: every step reads both operands and writes its result once
    byts = 0.0
    esz = 8 if cdt == "c64" else 16
    for p in plans:
        for s in range(p.n_steps):
            st = p.step(s)
            k, m, n, b = st[2:6]
            sets = nb if p.step_flags(s) & 2 else 1
            byts += sets * esz * (2.0 ** (k + m + b) + 2.0 ** (k + n + b) + 2.0 ** (m + n + b))
    t_hbm, t_pipe = byts / (pk["hbm_gbs"] * 1e9), flops / (pipe_peak * 1e12)
    res.update({
        "value": sets_all / (ms * 1e-3), "unit": "evals/s", "ms_per_step": ms, "steps": steps, "fwd_only_ms": fwd_ms,
        "fwd_only_evals_per_s": sets_all / (fwd_ms * 1e-3),
        "steps_by_kernel": {names.get(k, str(k)): kinds.count(k) for k in sorted(set(kinds))},
        "workload": f"{desc.split(',')[0]} in tensor-network mode, batch {nb}: fwd = contraction plan(s) "
                    f"({len(plans)} network(s), {sum(p.n_steps for p in plans)} pairwise steps, width "
                    f"{max(p.width for p in plans)}, {flops / nb:.3e} flop per set, in-repo planner) + bwd = "
                    f"{'contraction tree' if ho.get('tn_backward') == 'tree' else 'adjoint sweeps'}"
                    + (f"; the {len(plans)} measurement networks are dealt round-robin to the ranks, one all-reduce"
                       if measurement_parallel else ""),
        "roofline": {"kernel": "contraction plan (forward)", "bound": "hbm" if t_hbm > t_pipe else "pipe",
                     "achieved": byts / (fwd_ms * 1e-3) / 1e9 if t_hbm > t_pipe else flops / (fwd_ms * 1e-3) / 1e12,
                     "peak": pk["hbm_gbs"] if t_hbm > t_pipe else pipe_peak,
                     "unit": "GB/s" if t_hbm > t_pipe else "TFLOP/s",
                     "frac": max(t_hbm, t_pipe) / (fwd_ms * 1e-3), "traffic": None,
                     "algorithmic_bytes": byts, "algorithmic_flops": flops,
                     "note": "per-step algorithmic bytes (both operands read, result written once) and 8MNK flops of "
                             "the whole plan against the forward time; intermediates of the fused runs stay in L1/L2"},
        "gpu_launches": l1 - l0})
    # e2e through the Python API with host buffers
    hpin = torch.from_numpy(np.ascontiguousarray(flat_np)).pin_memory()

    def py_step():
        xx = hpin.to(device, non_blocking=True).requires_grad_(True)
        yy = cc.batched(xx)
        yy.sum().backward()
        return yy.detach().cpu(), xx.grad.cpu()

    py_step()
    dist.barrier()
    n_e2e = max(2, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        py_step()
    py_s = dist.max(time.perf_counter() - t0)
    item = hpin.element_size()
    res["e2e"] = {"value": sets_all * n_e2e / py_s, "unit": "evals/s", "steps": n_e2e,
                  "h2d_bytes_per_step": int(nb * P * item), "d2h_bytes_per_step": int((y.numel() + nb * P) * item),
                  "api": "cc.batched(x_cuda).sum().backward() on a tn_mode=True backend, pinned host in, host out"}
    return res


# ----------------------------------------------------------------------------------------------------------
# C5: sliced single-amplitude contraction (headline)
# ----------------------------------------------------------------------------------------------------------
def c5_hyper(greedy_plan, contract_parallel):
    extra = {"slice_batch": int(os.environ["TQ_C5_SLICE_BATCH"])} if "TQ_C5_SLICE_BATCH" in os.environ else {}
    if "TQ_C5_AMP_BATCH" in os.environ:      # amplitudes per launch sequence (default: fill 2^(28 - width) sets)
        extra["amplitude_batch"] = int(os.environ["TQ_C5_AMP_BATCH"])
    return {**extra, "max_repeats": C5_HYPER["max_repeats_greedy"] if greedy_plan else C5_HYPER["max_repeats"],
            "reconf_sweeps": 0 if greedy_plan else C5_HYPER["reconf_sweeps"],
            "reconf_leaves": 8 if greedy_plan else C5_HYPER["reconf_leaves"],
            "restarts": 1 if greedy_plan else C5_HYPER["restarts"],
            "time_model": None if greedy_plan else C5_HYPER["time_model"],
            "slicing_opts": dict(C5_HYPER["slicing_opts"], contract_parallel=contract_parallel),
            "plan_cache": PLAN_CACHE}


def measure_c5(steps, warmup, device, dist, do_cpu=False, greedy_plan=False, n_amps=AMPS, with_clocks=True,
               parity_groups=1):
    """BASELINE config 5.  Slices are sharded over ranks and combined with one all-reduce per step (strong scaling)."""
    import tedq_b200 as qb
    from tedq_b200 import capi, workloads as W

    pk = load_peaks()
    spec = W.lattice_rcs(5, 8, 12, seed=0)
    circ = W.build_circuit(spec, qb)
    t_plan = time.perf_counter()
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                             hyper_opt=c5_hyper(greedy_plan, dist.on))
    bits = c5_bitstrings(n_amps)
    amp = cc.amplitudes(bits[:1])
    torch.cuda.synchronize(device)
    plan_s = time.perf_counter() - t_plan
    info = cc._tn._amplitude_plan()[1]
    plan = cc._tn._amplitude_plan()[2]
    group = len(cc._tn.slice_members(0))          # slices that share one launch sequence (hyper_opt["slice_batch"])
    L = capi.lib()
    res = {"dtype": "c64", "scaling": "strong", "unit": "evals/s"}

    # ---- parity gate BEFORE timing (rank 0): slice-group amplitudes of two bitstrings against the CPU oracle's
    # torch.tensordot contraction of the same slices in complex128 (the complex64 host contraction has the same
    # rounding error as the device and cannot referee 1e-5), relative to the largest slice amplitude
    tag = "c5g" if greedy_plan else "c5"
    if dist.rank == 0 and parity_groups > 0:
        errs = []
        for which in ((0, 1) if n_amps > 1 else (0,)):     # <0...0| and one random bitstring
            members = [cc._tn.slice_members(i) for i in range(parity_groups)]
            _, amps, _, _ = c5_cpu_slices(0, slice_ids=[sid for m in members for sid in m], bits=bits[which],
                                          greedy_plan=greedy_plan, dtype=torch.complex128)
            got = [complex(cc.amplitude(bits[which].tolist(), slice_range=(i, i + 1)).cpu()) for i in range(parity_groups)]
            want = [sum(amps[i * group:(i + 1) * group]) for i in range(parity_groups)]
            scale = max(abs(a) for a in amps) or 1.0   # a slice can be exactly zero (a sliced wire next to a |0> cap)
            errs.append(max(abs(g - a) / scale for g, a in zip(got, want)))
        assert_parity(f"{tag} slice amplitudes vs host torch.tensordot (complex128)", max(errs), 1e-5)
        res["parity"] = {"slices_checked": parity_groups * group * len(errs), "max_rel_err": max(errs),
                         "tolerance": 1e-5, "asserted": True,
                         "reference": "oracle/tn_ref.contract_slice_torch, complex128, same path and slices"}
        if do_cpu:
            n_cpu = int(os.environ.get("TQ_C5_CPU_SLICES", "4"))
            cpu_dt, _, _, cpu_fl = c5_cpu_slices(n_cpu, greedy_plan=greedy_plan)
            res["cpu_baseline"] = {
                "value": 1.0 / (cpu_dt * 64), "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{n_cpu} of 64 slices of one amplitude contracted with torch.tensordot complex64 on the "
                          f"host ({cpu_dt:.2f} s per slice, {cpu_fl / cpu_dt / 1e9:.0f} GFLOP/s), extrapolated x64; "
                          f"host has {os.cpu_count()} logical cores"}
    for _ in range(max(1, warmup)):
        amp = cc.amplitudes(bits)
    dist.barrier()
    sampler = ClockSampler(device.index or 0) if with_clocks else None
    if sampler:
        sampler.start()
    l0 = int(L.tq_tn_launch_count())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        amp = cc.amplitudes(bits)          # every rank: its slice range of every amplitude; ONE all-reduce per step
    ev1.record()
    dist.barrier()
    l1 = int(L.tq_tn_launch_count())
    if sampler:
        res["clocks"] = sampler.finish()
    total_ms = dist.max(ev0.elapsed_time(ev1))
    res["value"] = n_amps * steps / (total_ms * 1e-3)
    res["ms_per_step"] = total_ms / steps
    res["ms_per_amplitude"] = total_ms / steps / n_amps
    res["steps"] = steps
    res["gpu_launches"] = (l1 - l0) + steps          # + tq_tn_operands (k_gate_tensors) once per step
    res["amplitude_0"] = [float(amp[0].real), float(amp[0].imag)]
    res["planner"] = {"search_s": info.search_s, "from_plan_cache": bool(info.from_cache),
                      "compile_and_first_call_s": round(plan_s, 2)}

    # ---- e2e: public API, HOST bitstrings in, HOST amplitudes out, wall clock
    hbits = torch.from_numpy(bits)
    cc.amplitudes(hbits).cpu()
    dist.barrier()
    n_e2e = max(2, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        host_amp = cc.amplitudes(hbits).cpu()
    e2e_s = dist.max(time.perf_counter() - t0)
    res["e2e"] = {"value": n_amps * n_e2e / e2e_s, "unit": "evals/s", "steps": n_e2e,
                  # per amplitude the operand pointer table (cap pointers chosen by the bitstring) goes to the device
                  "h2d_bytes_per_step": int(n_amps * plan.n_inputs * 16), "d2h_bytes_per_step": int(n_amps * 8),
                  "api": "cc.amplitudes(host bitstrings [A, 40]) -> .cpu(): the bitstrings select the closing caps "
                         "on the host (pointer tables are the H2D traffic), 8 bytes per amplitude return"}
    assert_parity("c5 e2e amplitudes equal the device-timed ones",
                  float((host_amp - amp.cpu()).abs().max() / amp.abs().max().cpu()), 1e-6)

    # ---- per-step table of one slice group (tq_tn_profile: CUDA events around every step) and the roofline
    rows, amps_per_seq, sets_per_seq, pin_ms = cc._tn.amplitudes_profile(torch.zeros((1, 0), device=device), bits, 0)
    per_slice = [r for r in rows if r["per_slice"]]
    slice_ms = sum(r["ms"] for r in per_slice)
    once_ms = sum(r["ms"] for r in rows if not r["per_slice"])
    table = []
    tc_flops = tc_ms = tc_gemm_ms = 0.0
    kernel_names = ["k_tn_step", "k_tn_gemm", "k_tc_pack+k_tc_gemm", "k_tn_dot", "k_tn_fused", "k_tn_apply", "k_tn_seed",
                    "k_tn_chain"]
    peak_cgemm = pk["tf32"] / 3.0
    for r in sorted(per_slice, key=lambda r: -r["ms"]):
        if r["ms"] < 0.01 * slice_ms:
            break
        if r["kernel"] == 4:   # a fused run of small steps is timed as one launch (on its first member)
            table.append({"kernel": "k_tn_fused (run of small steps)", "ms": round(r["ms"], 4)})
            continue
        sets = r["sets"] if r.get("per_set") else 1   # a grouped step runs `sets` slices' GEMMs in one launch
        fl = sets * 8.0 * 2.0 ** (r["k"] + r["m"] + r["n"] + r["b"])
        byts = sets * 8.0 * (2.0 ** (r["k"] + r["m"] + r["b"]) + 2.0 ** (r["k"] + r["n"] + r["b"]) + 2.0 ** (r["m"] + r["n"] + r["b"]))
        t_fl = fl / (peak_cgemm * 1e12)
        t_by = byts / (pk["hbm_gbs"] * 1e9)
        roof_ms = max(t_fl, t_by) * 1e3
        table.append({"M": 2 ** r["m"], "N": 2 ** r["n"], "K": 2 ** r["k"], "batch": 2 ** r["b"] * sets,
                      "kernel": kernel_names[r["kernel"]], "ms": round(r["ms"], 4), "pack_ms": round(r["pack_ms"], 4),
                      "algorithmic_tflops": round(fl / (r["ms"] * 1e-3) / 1e12, 2),
                      "bound": "tensor" if t_fl > t_by else "hbm", "roofline_ms": round(roof_ms, 4),
                      "frac": round(roof_ms / r["ms"], 3)})
        if r["kernel"] == 2 and t_fl > t_by:
            tc_flops += fl
            tc_ms += r["ms"]
            tc_gemm_ms += r["ms"] - r["pack_ms"]
    flops_amp = 2.0 ** info.flops_log2 * info.n_slices
    ach = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else 0.0
    traffic = traffic_detail = None      # dram__bytes_read + write of the longest GEMM launch, from one ncu --set full
    try:
        with open(os.path.join(ROOT, "profiles", "r02_c5_traffic.json")) as fh:
            traffic_detail = json.load(fh)
        traffic = traffic_detail["dram_bytes_per_launch"]
    except Exception:
        pass
    res["roofline"] = {
        "bound": "tensor", "kernel": "k_tc_gemm (+ k_tc_pack) on the tensor-bound steps of a slice group",
        "unit": "TFLOP/s", "achieved": ach, "peak": peak_cgemm, "frac": ach / peak_cgemm if tc_ms else None,
        "traffic": traffic, "traffic_detail": traffic_detail,
        "peak_kind": f"complex-GEMM tensor-core roofline = TF32 peak / 3; TF32 peak {pk['tf32']:.1f} TFLOP/s "
                     f"{pk['tc_kind']}; sustained {pk['tf32_sustained']:.1f}",
        "frac_of_sustained_peak": ach / (pk["tf32_sustained"] / 3.0) if tc_ms else None,
        "tensor_bound_steps_share_of_slice_time": tc_ms / slice_ms if slice_ms else None,
        "gemm_kernel_only_algorithmic_tflops": tc_flops / (tc_gemm_ms * 1e-3) / 1e12 if tc_gemm_ms else None,
        "gemm_kernel_only_frac": tc_flops / (tc_gemm_ms * 1e-3) / 1e12 / peak_cgemm if tc_gemm_ms else None,
        "whole_amplitude_algorithmic_tflops": flops_amp * n_amps * steps / (total_ms * 1e-3) / 1e12,
        "whole_amplitude_frac": flops_amp * n_amps * steps / (total_ms * 1e-3) / 1e12 / peak_cgemm,
        "note": "achieved = ALGORITHMIC flops (8*M*N*K per complex GEMM) of the steps whose roofline is the tensor "
                "pipe, divided by their measured time INCLUDING operand packing (tq_tn_profile events, live in this "
                "run).  A complex64 GEMM at fp32 accuracy executes 24*M*N*K TF32 flops (4M real decomposition x "
                "3-term error-compensated split), so its roofline is TF32 peak / 3.  whole_amplitude_* divides the "
                "flops of all pairwise steps x 64 slices by the timed region (launch gaps, HBM-bound steps, the "
                "all-reduce included).",
    }
    res["workload"] = (f"{C5_DESC}; {sets_per_seq} slices per launch sequence ({group} slices of each of {amps_per_seq} "
                       f"amplitudes), width {plan.width}, {plan.n_steps} pairwise steps, {flops_amp:.3e} flop per amplitude"
                       + (" (plain greedy tree, no subtree reconfiguration)" if greedy_plan else ""))
    res["per_slice_ms_profiled"] = slice_ms / sets_per_seq
    res["once_per_launch_sequence_ms_profiled"] = once_ms + pin_ms
    res["slices_per_launch_sequence"] = sets_per_seq
    res["amplitudes_per_launch_sequence"] = amps_per_seq
    res["steps_per_slice"] = len(per_slice)
    res["steps_once_per_call"] = len(rows) - len(per_slice)
    res["step_table"] = table
    return res


def measure_c5_simplified(steps, warmup, device, dist):
    """Config 5 again with tn_simplify=True (the reference's default flag; its own simplifier does not work):
    CNOT controls and RZ gates stay on shared wire indices, the plan needs no slicing (every rank contracts every
    amplitude: replicas)."""
    import tedq_b200 as qb
    from tedq_b200 import workloads as W

    spec = W.lattice_rcs(5, 8, 12, seed=0)
    cc = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=True,
                                                  hyper_opt={"max_repeats": 64, "plan_cache": PLAN_CACHE})
    bits = [0] * 40
    for _ in range(max(1, warmup)):
        amp = cc.amplitude(bits)
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        amp = cc.amplitude(bits)
    ev1.record()
    dist.barrier()
    ms = dist.max(ev0.elapsed_time(ev1)) / steps
    plan = cc._tn._amplitude_plan()[2]
    kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
    return {"value": 1e3 / ms, "unit": "evals/s", "ms_per_step": ms, "dtype": "c64", "scaling": "replicas",
            "amplitude": [float(amp.real), float(amp.imag)],
            "steps_by_kernel": {k: kinds.count(k) for k in sorted(set(kinds))},
            "workload": f"c5 with tn_simplify=True: same circuit and amplitude <0|U|0>, diagonal / controlled gates on "
                        f"shared wire indices: {plan.n_slices} slice(s), width {plan.width}, "
                        f"{plan.flops * plan.n_slices:.3e} flop per amplitude (dense network: 2.7e12); per-GPU rate"}


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--extras", default=os.environ.get("TQ_BENCH_EXTRAS", "c1,c2,c2tn,c3,c3tn,c4,c4d20,c4tn,c5g,c5s"),
                    help="other BASELINE configs measured briefly and reported inside the same JSON line ('none')")
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
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: ONE JSON line
        import datetime
        # a rank that dies inside a collective must take the job down in minutes, not after NCCL's default 10
        torch.distributed.init_process_group("nccl", device_id=device,
                                             timeout=datetime.timedelta(seconds=int(os.environ.get("TQ_NCCL_TIMEOUT", "300"))))
    dist = Dist(device)
    do_cpu = rank == 0 and world == 1 and os.environ.get("TQ_BENCH_CPU", "1") == "1"
    budget = float(os.environ.get("TQ_CPU_BASELINE_SECONDS", "8"))

    t_bench = time.perf_counter()
    if args.workload == "c5":
        main_res = measure_c5(args.steps, args.warmup, device, dist, do_cpu=do_cpu)
    else:
        main_res = measure_sv(args.workload, args.steps, args.warmup, device, dist, do_cpu, budget, with_clocks=True)

    extras = {}
    names = [] if args.extras in ("", "none") else [e for e in args.extras.split(",") if e and e != args.workload]
    for name in names:
        t0 = time.perf_counter()
        try:
            if name == "c5":
                r = measure_c5(3, 1, device, dist, do_cpu=do_cpu, n_amps=4, with_clocks=False)
            elif name == "c5g":
                r = measure_c5(3, 1, device, dist, do_cpu=False, greedy_plan=True, n_amps=4, with_clocks=False,
                               parity_groups=1)
                r["step_table"] = r["step_table"][:8]
            elif name == "c5s":
                r = measure_c5_simplified(20, 3, device, dist)
            elif name in ("c2tn", "c4tn"):
                r = measure_tn_mode(name[:2], 5, 3, device, dist, do_cpu, budget)
            elif name == "c3tn":
                # C3's 20 Z expectation values as 20 light-cone-pruned networks, dealt to the ranks (measurement
                # sharding), gradients by reverse mode through the contraction trees; 8 parameter sets
                r = measure_tn_mode("c3", 3, 2, device, dist, do_cpu, budget, measurement_parallel=True, batch=8,
                                    hyper={"max_repeats": 8, "light_cone": True, "tn_backward": "adjoint"})
            else:
                r = measure_sv(name, max(3, min(5, args.steps)), 3, device, dist, do_cpu, budget)
            r["bench_seconds"] = round(time.perf_counter() - t0, 1)
            extras[name] = r
        except ParityError:
            raise
        except Exception as exc:      # one broken extra must not lose the headline
            import traceback
            extras[name] = {"error": f"{type(exc).__name__}: {exc}", "trace": traceback.format_exc()[-600:]}
        torch.cuda.empty_cache()

    if rank == 0:
        wl = args.workload
        line = {
            "metric": "circuit_evals_per_sec", "value": main_res["value"], "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": main_res["scaling"], "vs_baseline": None,
            "dtype": main_res["dtype"], "data": "synthetic",
            "config": config_of(wl, C5_DESC if wl == "c5" else main_res["workload"]),
            "engine": {k: main_res[k] for k in ("workload", "ms_per_amplitude", "slices_per_launch_sequence",
                                                "per_slice_ms_profiled", "once_per_launch_sequence_ms_profiled", "amplitudes_per_launch_sequence", "steps_per_slice",
                                                "steps_once_per_call", "planner", "amplitude_0", "parity",
                                                "sets_per_gpu", "e2e_python") if k in main_res},
            "parallelism": (f"slices sharded over {world} GPU(s), one all-reduce per step" if wl == "c5" else
                            f"dp{world} (parameter sets sharded, no collective on the data path)"),
            "roofline": main_res["roofline"], "clocks": main_res.get("clocks"), "e2e": main_res.get("e2e"),
            "gpu_launches": main_res["gpu_launches"],
        }
        if "cpu_baseline" in main_res:
            line["cpu_baseline"] = main_res["cpu_baseline"]
        if "step_table" in main_res:
            line["step_table"] = main_res["step_table"]
        if extras:
            line["other_configs"] = extras
        line["bench_seconds"] = round(time.perf_counter() - t_bench, 1)
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
