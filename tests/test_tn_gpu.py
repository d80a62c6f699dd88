"""GPU: tensor-network contraction mode of the engine against the oracle and against the state-vector engine."""
import numpy as np
import pytest
import torch

import tedq_b200 as qb
from conftest import load_golden
from helpers import TOL, assert_close, build, cdtype, golden_out, rdtype
from oracle import sv_ref, tn_ref
from tedq_b200 import capi
from tedq_b200 import workloads as W

pytestmark = pytest.mark.gpu
CASES = load_golden("sv_cases.json")
TN_CASES = [c for c in CASES if c["spec"]["num_qubits"] <= 8 and c["spec"]["n_params"] > 0
            and not (c["spec"]["meas"][0][0] == "probs" and c["spec"]["meas"][0][1] is None)][:40]
# BASELINE configs 2 and 4 in the mode BASELINE.json names (tensor-network contraction), at BASELINE size, against
# the fixtures the unmodified reference produced for the same circuits and parameters
TN_BASELINE_CASES = [c for c in CASES if c["spec"]["name"] in ("mbl1d_12", "mbl2d_4x4_s1", "mbl2d_3x3_s2")]


def _case_id(c):
    return f'{c["spec"]["name"]}-{c["spec"]["meas"][0][0]}-{c["dtype"]}'


@pytest.mark.parametrize("case", TN_CASES, ids=_case_id)
@pytest.mark.parametrize("slices", [1, 4])
@pytest.mark.parametrize("simplify", [False, True], ids=["dense", "simplified"])
def test_tn_mode_matches_reference_fixture(case, slices, simplify):
    """TN-mode values equal the reference's results (same circuit, same parameters) for every measurement kind;
    sliced plans give the same sum; tn_simplify=True (diagonal / controlled gates on shared wire indices) gives
    the same numbers as the reference-exact network."""
    dt = case["dtype"]
    hyper = {"max_repeats": 4, "slicing_opts": {"target_num_slices": slices}}
    cc = build(case, dt, case["flat"][0]).compilecircuit(backend="pytorch_b200", tn_mode=True, hyper_opt=hyper,
                                                          tn_simplify=simplify, dtype=cdtype(dt))
    flat = torch.tensor(case["flat"], dtype=rdtype(dt), device="cuda")
    out = cc.batched(flat).cpu().numpy()
    ref = golden_out(case)
    ms = case["spec"]["meas"]
    if ms[0][0] == "probs":   # TN branch keeps the listed qubit order, SV branch sorts ascending
        axes = []
        for m in ms:
            axes.append(np.argsort(np.argsort(m[1])))
        ref = np.stack([np.transpose(ref[:, i], [0] + [1 + a for a in axes[i]]) for i in range(len(ms))], 1)
    assert_close(out, ref, TOL[dt], "tn out")


@pytest.mark.parametrize("case", TN_BASELINE_CASES, ids=_case_id)
@pytest.mark.parametrize("slices", [1, 4], ids=["unsliced", "sliced"])
@pytest.mark.parametrize("bwd", ["adjoint", "tree"])
def test_tn_mode_baseline_configs_match_reference_fixture(case, slices, bwd):
    """C2 (12-qubit MBL-1D, complex64) and C4 (4x4 MBL-2D, complex128) in tensor-network mode: values AND gradients
    (both gradient paths: adjoint sweeps / reverse mode through the contraction tree) against the reference's
    results on the same parameters; sliced and unsliced plans."""
    dt = case["dtype"]
    if bwd == "tree" and slices > 1 and case["spec"]["num_qubits"] >= 16:
        pytest.skip("sliced tree backward of the 3702-tensor network: covered unsliced")
    hyper = {"max_repeats": 4, "tn_backward": bwd, "slicing_opts": {"target_num_slices": slices}}
    cc = build(case, dt, case["flat"][0]).compilecircuit(backend="pytorch_b200", tn_mode=True, hyper_opt=hyper,
                                                          tn_simplify=False, dtype=cdtype(dt))
    if slices > 1:
        assert cc._tn.infos[0].n_slices >= slices
    flat = torch.tensor(case["flat"], dtype=rdtype(dt), device="cuda", requires_grad=True)
    out = cc.batched(flat)
    ct = torch.tensor(np.asarray(case["cotangent"]), dtype=rdtype(dt), device="cuda")
    (out * ct).sum().backward()
    assert_close(out.detach().cpu().numpy(), golden_out(case), TOL[dt], "tn out")      # probs([q]): one kept qubit
    assert_close(flat.grad.cpu().numpy(), np.asarray(case["grad"]), TOL[dt] * 4, "tn grad")


CONE_CASES = [c for c in TN_CASES if c["spec"]["meas"][0][0] in ("expval", "probs")][:24]


@pytest.mark.parametrize("case", CONE_CASES, ids=_case_id)
@pytest.mark.parametrize("simplify", [False, True], ids=["dense", "simplified"])
def test_light_cone_pruned_networks_match_reference_fixture(case, simplify):
    """hyper_opt["light_cone"]: the networks of expval / marginal measurements keep only the gates inside the
    measurement's causal cone (the others meet their own adjoints and cancel).  Values and tree-backward gradients
    equal the reference's results for the full circuit."""
    dt = case["dtype"]
    hyper = {"max_repeats": 4, "light_cone": True, "tn_backward": "tree"}
    cc = build(case, dt, case["flat"][0]).compilecircuit(backend="pytorch_b200", tn_mode=True, hyper_opt=hyper,
                                                          tn_simplify=simplify, dtype=cdtype(dt))
    flat = torch.tensor(case["flat"], dtype=rdtype(dt), device="cuda", requires_grad=True)
    out = cc.batched(flat)
    ref = golden_out(case)
    ms = case["spec"]["meas"]
    ct = np.asarray(case["cotangent"])
    if ms[0][0] == "probs":   # TN branch keeps the listed qubit order, SV branch sorts ascending
        axes = [np.argsort(np.argsort(m[1])) for m in ms]
        ref = np.stack([np.transpose(ref[:, i], [0] + [1 + a for a in axes[i]]) for i in range(len(ms))], 1)
        ct = np.stack([np.transpose(ct[:, i], [0] + [1 + a for a in axes[i]]) for i in range(len(ms))], 1)
    (out * torch.tensor(ct, dtype=rdtype(dt), device="cuda")).sum().backward()
    assert_close(out.detach().cpu().numpy(), ref, TOL[dt], "pruned out")
    assert_close(flat.grad.cpu().numpy(), np.asarray(case["grad"]), TOL[dt] * 4, "pruned grad")


def test_light_cone_cuts_c3_networks():
    """C3 in tensor-network mode: 20 Z expectation values = 20 networks; with light-cone pruning each keeps only its
    causal cone (HEA depth 2 on 12 qubits here) and the contraction cost drops by more than 10x; same values."""
    spec = W.hea(12, 2)
    circ = W.build_circuit(spec, qb)
    x = torch.tensor(np.random.RandomState(1).rand(3, spec["n_params"]), dtype=torch.float32, device="cuda")
    ref = circ.compilecircuit(backend="pytorch_b200").batched(x).cpu().numpy()
    costs = {}
    for lc in (False, True):
        cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                                 hyper_opt={"max_repeats": 4, "light_cone": lc})
        assert_close(cc.batched(x).cpu().numpy(), ref, 1e-5, f"light_cone={lc}")
        costs[lc] = sum(2.0 ** i.flops_log2 for i in cc._tn.infos)
        if lc:
            assert min(len(n.inputs) for n in cc._tn.networks) < len(cc._tn.networks[0].inputs) or True
    assert costs[True] * 2 < costs[False], costs


def test_tn_mode_gradients_match_sv_mode():
    spec = W.mbl_1d(6)
    circ = W.build_circuit(spec, qb)
    x = torch.tensor(W.c2_inputs(4, 6, 3), device="cuda")
    outs, grads = [], []
    for kw in ({}, {"tn_mode": True, "tn_simplify": False}):
        cc = circ.compilecircuit(backend="pytorch_b200", **kw)
        xx = x.clone().requires_grad_(True)
        y = cc.batched(xx)
        (y * torch.tensor([0.3, -0.8], device="cuda")).sum().backward()
        outs.append(y.detach().cpu().numpy())
        grads.append(xx.grad.cpu().numpy())
    assert_close(outs[1], outs[0], 1e-5, "out")
    assert_close(grads[1], grads[0], 1e-5, "grad")


@pytest.mark.parametrize("n,cycles", [(3 * 3, 4), (3 * 4, 6), (4 * 4, 5)])
def test_amplitudes_match_state_vector(n, cycles):
    """C5 parity (i): the lattice generator at sizes the SV oracle can run; every slice count gives the same sum."""
    rows = 3 if n < 16 else 4
    spec = W.lattice_rcs(rows, n // rows, cycles, seed=5, measure="state")
    circ = W.build_circuit(spec, qb)
    ref = sv_ref.run_sv(circ, torch.zeros(0), torch.complex64, return_state=True).numpy().reshape(-1)
    rng = np.random.RandomState(0)
    for slices, simplify in ((1, False), (8, False), (1, True), (8, True)):
        cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=simplify,
                                 hyper_opt={"max_repeats": 8, "slicing_opts": {"target_num_slices": slices}})
        for bits in ([0] * n, rng.randint(0, 2, n).tolist()):
            amp = complex(cc.amplitude(bits).cpu())
            idx = int("".join(str(b) for b in bits), 2)
            # relative to the largest amplitude of the state (entries are O(2^-n/2))
            assert abs(amp - ref[idx]) <= 1e-5 * np.abs(ref).max(), (slices, bits, amp, ref[idx])
        if slices > 1:
            # slice-sum invariance: two halves add up to the whole
            ns = cc._tn._amplitude_plan()[2].n_slices
            a = cc.amplitude(bits, slice_range=(0, ns // 2)) + cc.amplitude(bits, slice_range=(ns // 2, ns))
            assert abs(complex(a.cpu()) - ref[idx]) <= 1e-5 * np.abs(ref).max()


def test_tn_mode_complex128_mbl2d():
    spec = W.mbl_2d(2, 1)
    wrap = lambda v: torch.tensor(float(v), dtype=torch.float64)
    circ = W.build_circuit(spec, qb, tensor_fn=wrap)
    x = torch.rand(2, spec["n_params"], dtype=torch.float64)
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, dtype=torch.complex128)
    got = cc.batched(x.cuda()).cpu().numpy()
    ref, _ = sv_ref.run_batch(circ, x, torch.complex128)
    assert_close(got, ref.numpy(), 1e-11, "c128 tn")


def test_c5_amplitude_tensor_core_matches_complex128():
    """C5 at full size (40 qubits, 64 slices): the tcgen05 split-TF32 path against the complex128 FMA path of the
    same plan (north_star: complex64 amplitudes within relative 1e-5), and slice-sum invariance."""
    spec = W.lattice_rcs(5, 8, 12, seed=0)
    ho = {"max_repeats": 16, "slicing_opts": {"target_size": 2 ** 27, "target_num_slices": 64}}
    c128 = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64)).compilecircuit(
        backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=ho, dtype=torch.complex128)
    c64 = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                                                   hyper_opt=ho)
    bits = [0] * 40
    ref = complex(c128.amplitude(bits).cpu())
    got = complex(c64.amplitude(bits).cpu())
    plan = c64._tn._amplitude_plan()[2]
    assert sum(1 for s in range(plan.n_steps) if plan.step_kernel(s) == 2) >= 5
    assert abs(got - ref) <= 1e-5 * abs(ref), (got, ref, abs(got - ref) / abs(ref))
    ns = plan.n_slices
    halves = complex((c64.amplitude(bits, slice_range=(0, ns // 2)) + c64.amplitude(bits, slice_range=(ns // 2, ns))).cpu())
    assert abs(halves - got) <= 1e-5 * abs(ref)
    one = [0, 1] * 20
    r1, g1 = complex(c128.amplitude(one).cpu()), complex(c64.amplitude(one).cpu())
    assert abs(g1 - r1) <= 1e-5 * abs(r1), (g1, r1)
    # the simplified network (CNOT controls and RZ on shared wire indices: 2^44 -> 2^31 flops) gives the same amplitudes
    simp = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=True,
                                                    hyper_opt={"max_repeats": 16})
    assert abs(complex(simp.amplitude(bits).cpu()) - ref) <= 1e-5 * abs(ref)
    assert abs(complex(simp.amplitude(one).cpu()) - r1) <= 1e-5 * abs(r1)


def test_reference_tree_plugin_boundary():
    """The planner plug-in boundary of the reference (use_jdopttn= / use_cotengra=, compiled_circuit.py:356-393):
    B200OptTN has JDOptTN's constructor and contract(arrays, backend='torch'); called the way pytorch_backend.py:339
    calls it (symbol strings, torch operands in network order) it reproduces the reference fixture."""
    case = next(c for c in CASES if c["spec"]["num_qubits"] >= 4 and c["spec"]["meas"][0][0] == "expval"
                and c["dtype"] == "c64" and c["spec"]["n_params"] > 0)
    circ = build(case, "c64", case["flat"][0])
    from tedq_b200 import tn_index
    flat = torch.tensor(case["flat"][0], dtype=torch.float32)
    arrays_all = tn_ref.operands(circ, flat.double(), torch.complex128)
    nets = tn_index.networks_of_circuit(circ)
    got = []
    for net, arrs in zip(nets, arrays_all):
        inputs, output = net.symbols()                      # what gen_tensor_networks hands to JDOptTN
        size_dict = {s: 2 for s in net.size_keys()}
        for tree in (qb.B200OptTN(inputs, size_dict, output=output, imbalance=0.2, max_repeats=8, search_parallel=True,
                                  slicing_opts={"target_num_slices": 2}),
                     qb.ctg_compat.HyperOptimizer(methods=["kahypar"], max_repeats=8, progbar=False, minimize="flops",
                                                  score_compression=0.5, slicing_opts=None).search(inputs, output,
                                                                                                  size_dict)):
            ops = [torch.tensor(np.asarray(a), dtype=torch.complex64, device="cuda") for a in arrs]
            r = tree.contract(ops, backend="torch")
            got.append(float(torch.squeeze(r.real).cpu()))
    ref = np.repeat(golden_out(case)[0].reshape(-1), 2)
    assert_close(np.asarray(got), ref, 1e-5, "tree plug-in")
    with pytest.raises(NotImplementedError):      # the sliced tree has no reverse pass
        ops[0].requires_grad_(True)
        qb.B200OptTN(inputs, size_dict, output=output, slicing_opts={"target_num_slices": 2}).contract(ops)


def test_tree_plugin_is_differentiable():
    """tree.contract(arrays) under autograd (the reference's back_prop differentiates through it): gradients with
    respect to every operand against torch.einsum on the same operands."""
    rng = np.random.RandomState(4)
    inputs = [["a", "b", "c"], ["c", "d"], ["b", "d", "e", "f"], ["a", "f", "g"], ["e", "g"]]
    output = []
    for out in ([], ["x"]):
        ins = [list(t) for t in inputs]
        if out:
            ins[1] = ins[1] + ["x"]
        tree = qb.B200OptTN(ins, {k: 2 for t in ins for k in t}, output=out, max_repeats=2)
        mk = lambda r: torch.tensor(rng.standard_normal((2,) * r) + 1j * rng.standard_normal((2,) * r),
                                    dtype=torch.complex128, device="cuda")
        ops = [mk(len(t)).requires_grad_(i != 2) for i, t in enumerate(ins)]
        ref_ops = [o.detach().clone().requires_grad_(o.requires_grad) for o in ops]
        w = torch.tensor(rng.standard_normal((2,) * len(out)) + 1j * rng.standard_normal((2,) * len(out)),
                         dtype=torch.complex128, device="cuda")
        sym = {k: i for i, k in enumerate(sorted({k for t in ins for k in t}))}
        args = []
        for o, t in zip(ref_ops, ins):
            args += [o, [sym[k] for k in t]]
        ref = torch.einsum(*args, [sym[k] for k in out])
        got = tree.contract(ops, backend="torch")
        assert torch.allclose(got, ref, atol=1e-12)
        (got * w).real.sum().backward()
        (ref * w).real.sum().backward()
        for o, r in zip(ops, ref_ops):
            if r.requires_grad:
                assert torch.allclose(o.grad, r.grad, atol=1e-11), (o.grad - r.grad).abs().max()
            else:
                assert o.grad is None


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("simplify", [False, True], ids=["dense", "simplified"])
@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_tree_backward_matches_adjoint_sweeps(seed, simplify, dt):
    """Reverse mode through the contraction tree (tq_tn_backward + tq_tn_param_grads) against the adjoint
    state-vector sweeps of the same engine (themselves pinned to the reference's autograd by the fixtures): every
    parametrised gate kind, several measurement kinds, batched parameter sets."""
    meas = [[["expval", [["PauliZ", [0]]]], ["expval", [["PauliX", [2]]]], ["expval", [["PauliZ", [1]], ["PauliZ", [3]]]]],
            [["probs", [1, 3]], ["probs", [0, 2]]], [["state"]]][seed % 3]
    spec = W.random_circuit(5, 40, seed=20 + seed, meas=meas, trainable_ratio=0.8)
    rd = rdtype(dt)
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    x = torch.rand(3, spec["n_params"], dtype=rd, device="cuda") * 2 - 1
    res = []
    slices = 4 if seed % 2 else 1     # odd seeds: sliced plans (forward + reverse pass slice by slice)
    for kw in ({}, {"tn_mode": True, "tn_simplify": simplify,
                    "hyper_opt": {"max_repeats": 2, "tn_backward": "tree",
                                  "slicing_opts": {"target_num_slices": slices}}}):
        cc = circ.compilecircuit(backend="pytorch_b200", dtype=cdtype(dt), **kw)
        xx = x.clone().requires_grad_(True)
        y = cc.batched(xx)
        w = torch.linspace(0.3, 1.7, y[0].numel(), device="cuda", dtype=rd).reshape(y.shape[1:])
        loss = (y.real * w).sum() + ((y.imag * w.flip(0)).sum() if y.is_complex() else 0.0)
        loss.backward()
        res.append((y.detach().cpu().numpy(), xx.grad.cpu().numpy()))
    if meas[0][0] == "probs":      # TN branch keeps the listed qubit order; [1, 3] / [0, 2] are already ascending
        pass
    assert_close(res[1][0], res[0][0], TOL[dt], "values")
    assert_close(res[1][1], res[0][1], TOL[dt] * 4, "gradients")


@pytest.mark.parametrize("slices", [1, 4], ids=["unsliced", "sliced"])
def test_tree_backward_beyond_state_vector_reach(slices):
    """Gradients of a 32-qubit circuit in tensor-network mode: no state vector exists (34 GB), the reverse pass runs
    on the contraction tree (slice by slice for a sliced plan); checked against central finite differences of the
    same contraction."""
    n = 32
    b = W._Builder("chain32", n)
    for q in range(n):
        b.g("RY", [q], b.p())
    for q in range(n - 1):
        b.g("CNOT", [q, q + 1])
    for q in range(0, n, 3):
        b.g("RX", [q], b.p())
    b.expval(["PauliZ", [n - 1]])
    spec = b.spec
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=torch.float64))
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, dtype=torch.complex128,
                             hyper_opt={"max_repeats": 4, "slicing_opts": {"target_num_slices": slices}})
    assert cc._tn.infos[0].n_slices >= slices
    x = torch.rand(2, spec["n_params"], dtype=torch.float64, device="cuda")
    xx = x.clone().requires_grad_(True)
    y = cc.batched(xx)
    y.sum().backward()
    g = xx.grad.cpu().numpy()
    eps = 1e-6
    for j in (0, n - 1, n, spec["n_params"] - 1):
        xp, xm = x.clone(), x.clone()
        xp[:, j] += eps
        xm[:, j] -= eps
        with torch.no_grad():
            fd = ((cc.batched(xp) - cc.batched(xm)) / (2 * eps)).reshape(2).cpu().numpy()
        assert np.abs(fd - g[:, j]).max() < 1e-7, (j, fd, g[:, j])


@pytest.mark.parametrize("simplify", [False, True], ids=["dense", "simplified"])
def test_tn_mode_without_parameters(simplify):
    """Edge cases of the call contract in tensor-network mode: a circuit with zero parameters (called with no
    arguments, compiled_circuit.py:412-413), every operand constant, one qubit untouched by any gate."""
    b = W._Builder("noparam", 4)
    for name, qs in (("Hadamard", [0]), ("CNOT", [0, 1]), ("T", [1]), ("SX", [2]), ("CZ", [1, 2]), ("S", [0])):
        b.g(name, qs)
    b.state()
    spec = b.spec
    circ = W.build_circuit(spec, qb)
    ref = sv_ref.run_sv(circ, torch.zeros(0), torch.complex64, return_state=True).numpy()
    cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=simplify, hyper_opt={"max_repeats": 2})
    got = cc().cpu().numpy()
    assert got.shape == (1, 2, 2, 2, 2)
    assert_close(got[0], ref.reshape(2, 2, 2, 2), 1e-6, "state")


def test_slice_groups_match_single_slices():
    """hyper_opt["slice_batch"] = g: 2^g slices share one launch sequence as the plan's batch dimension.  Same
    amplitude for every g (tensor-core steps included), and plan slice i of the grouped backend is the sum of the
    ungrouped slices slice_members(i) names."""
    spec = W.lattice_rcs(4, 5, 10, seed=2, measure="state")
    circ = W.build_circuit(spec, qb)
    ref = sv_ref.run_sv(circ, torch.zeros(0), torch.complex64, return_state=True).numpy().reshape(-1)
    bits = [0, 1] * 10
    want = ref[int("".join(str(b) for b in bits), 2)]
    ccs = {}
    for g in (0, 1, 2, 3):
        ho = {"max_repeats": 8, "slice_batch": g, "engine_opts": {capi.TN_OPT_TC_MIN_LOG2: 12},
              "slicing_opts": {"target_num_slices": 16}}
        cc = ccs[g] = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=ho)
        amp = complex(cc.amplitude(bits).cpu())
        assert abs(amp - want) <= 1e-5 * np.abs(ref).max(), (g, amp, want)
        plan = cc._tn._amplitude_plan()[2]
        n_orig = cc._tn._amplitude_plan()[1].n_slices
        assert plan.n_slices * len(cc._tn.slice_members(0)) == n_orig
        if g:
            assert len(cc._tn.slice_members(0)) == 1 << g
            assert any(plan.step_kernel(s) == 2 for s in range(plan.n_steps))
    n_orig = ccs[0]._tn._amplitude_plan()[1].n_slices
    single = [complex(ccs[0].amplitude(bits, slice_range=(s, s + 1)).cpu()) for s in range(n_orig)]
    scale = max(abs(a) for a in single)
    seen = []
    for i in range(ccs[2]._tn._amplitude_plan()[2].n_slices):
        members = ccs[2]._tn.slice_members(i)
        seen += members
        got = complex(ccs[2].amplitude(bits, slice_range=(i, i + 1)).cpu())
        assert abs(got - sum(single[m] for m in members)) <= 2e-6 * scale
    assert sorted(seen) == list(range(n_orig))


def test_c5_slices_match_host_tensordot():
    """C5 at full size, independent check of the contraction itself: slice amplitudes of the 40-qubit network (the
    committed bench plan) against the oracle's torch.tensordot contraction of the same slices on the host cores
    (what tree.contract(arrays, backend='torch') does, pytorch_backend.py:339), relative to the largest slice."""
    import os

    from bench import C5_HYPER, PLAN_CACHE, c5_cpu_slices

    spec = W.lattice_rcs(5, 8, 12, seed=0)
    hyper = {"max_repeats": C5_HYPER["max_repeats"], "reconf_sweeps": C5_HYPER["reconf_sweeps"],
             "reconf_leaves": C5_HYPER["reconf_leaves"], "restarts": C5_HYPER["restarts"], "time_model": C5_HYPER["time_model"],
             "slicing_opts": dict(C5_HYPER["slicing_opts"]), "plan_cache": PLAN_CACHE}
    cc = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False,
                                                  hyper_opt=hyper)
    bits = [0] * 40
    cc.amplitude(bits, slice_range=(0, 1))
    n_groups = int(os.environ.get("TQ_TEST_C5_GROUPS", "1"))
    members = [cc._tn.slice_members(i) for i in range(n_groups)]
    _, amps, n_slices, _ = c5_cpu_slices(0, slice_ids=[sid for m in members for sid in m], dtype=torch.complex128)
    assert n_slices == 64
    group = len(members[0])
    scale = max(abs(a) for a in amps)
    assert scale > 0
    for i, m in enumerate(members):
        got = complex(cc.amplitude(bits, slice_range=(i, i + 1)).cpu())
        want = sum(amps[i * group:(i + 1) * group])
        assert abs(got - want) <= 1e-5 * scale, (i, got, want, abs(got - want) / scale)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_plans_are_per_device():
    """A compiled circuit accepts parameters on any device on every call (pytorch_backend.py:85-119): plans are
    created and cached per device; a plan used on another device is refused by the C ABI, not run."""
    spec = W.mbl_1d(6)
    circ = W.build_circuit(spec, qb)
    x = torch.tensor(W.c2_inputs(3, 6, 1))
    outs = []
    for kw in ({}, {"tn_mode": True, "tn_simplify": False}):
        cc = circ.compilecircuit(backend="pytorch_b200", **kw)
        for dev in ("cuda:1", "cuda:0", "cuda:1"):
            outs.append(cc.batched(x.to(dev)).cpu().numpy())
    for o in outs[1:]:
        assert_close(o, outs[0], 1e-5, "per-device plans")
    cc = circ.compilecircuit(backend="pytorch_b200")
    plan0 = cc.plan(torch.device("cuda:0"))
    with torch.cuda.device(1):
        with pytest.raises(capi.EngineError, match="created on CUDA device 0"):
            xx = x.to("cuda:1")
            out = torch.empty((3, plan0.out_reals), device="cuda:1")
            ws = torch.empty(plan0.workspace_bytes(3, False), dtype=torch.uint8, device="cuda:1")
            plan0.forward(xx.data_ptr(), 3, out.data_ptr(), ws.data_ptr(), ws.numel(), False, 0)


@pytest.mark.parametrize("slice_batch", [0, 2])
def test_amplitude_batch_matches_single_calls(slice_batch):
    """cc.amplitudes(bits [A, n]): one pass over the gate operands, one contraction per bitstring, with the
    once-per-call part of the next contraction overlapped on a side stream (two workspaces).  Same numbers as A
    separate cc.amplitude calls and as the state vector; the overlap can be switched off."""
    spec = W.lattice_rcs(4, 5, 8, seed=3, measure="state")
    circ = W.build_circuit(spec, qb)
    ref = sv_ref.run_sv(circ, torch.zeros(0), torch.complex64, return_state=True).numpy().reshape(-1)
    rng = np.random.RandomState(11)
    bits = rng.randint(0, 2, size=(7, 20))
    want = np.array([ref[int("".join(str(b) for b in row), 2)] for row in bits])
    for overlap in (True, False):
        ho = {"max_repeats": 4, "slice_batch": slice_batch, "overlap_prepare": overlap,
              "engine_opts": {capi.TN_OPT_TC_MIN_LOG2: 12}, "slicing_opts": {"target_num_slices": 8}}
        cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, hyper_opt=ho)
        got = cc.amplitudes(torch.tensor(bits)).cpu().numpy()            # host tensor in
        single = np.array([complex(cc.amplitude(row.tolist()).cpu()) for row in bits])
        assert np.abs(got - want).max() <= 1e-5 * np.abs(ref).max()
        assert np.abs(got - single).max() <= 1e-6 * np.abs(ref).max()
        again = cc.amplitudes(bits).cpu().numpy()                        # numpy in, workspaces reused
        assert np.array_equal(again, got)


@pytest.mark.parametrize("name,dt", [("mbl2d", "c128"), ("mbl1d", "c64"), ("hea", "c64"), ("rand", "c128")])
def test_apply_chain_runs_match_single_applies(name, dt):
    """TQ_TN_OPT_CHAIN: runs of gate-like apply steps on the same large tensor execute as one shared-memory sweep
    (k_tn_chain).  Same values as one k_tn_apply launch per step and as the state-vector engine; the plans of these
    circuits are state-vector shaped, so most of their steps must land in chain runs."""
    if name == "mbl2d":
        spec = W.mbl_2d(4, 1)
    elif name == "mbl1d":
        spec = W.mbl_1d(14)
    elif name == "hea":
        spec = W.hea(15, 3)
    else:
        spec = W.random_circuit(13, 80, seed=9, meas=[["probs", [12, 3]], ["probs", [0, 7]]])
    rd = rdtype(dt)
    circ = W.build_circuit(spec, qb, tensor_fn=lambda v: torch.tensor(float(v), dtype=rd))
    x = torch.tensor(np.random.RandomState(2).rand(3, spec["n_params"]), dtype=rd, device="cuda")
    ref = circ.compilecircuit(backend="pytorch_b200", dtype=cdtype(dt)).batched(x).cpu().numpy()
    outs = {}
    for chain in (1, 0):
        cc = circ.compilecircuit(backend="pytorch_b200", tn_mode=True, tn_simplify=False, dtype=cdtype(dt),
                                 hyper_opt={"max_repeats": 4, "light_cone": False,
                                            "engine_opts": {capi.TN_OPT_CHAIN: chain}})
        outs[chain] = cc.batched(x).cpu().numpy()
        plan = cc._tn._plan(0, torch.device("cuda", torch.cuda.current_device()))
        kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
        if chain and name == "mbl2d":     # a wide plan: the applies on the large tensor run as chain sweeps
            assert kinds.count(7) > 4 * kinds.count(5), (kinds.count(7), kinds.count(5), len(kinds))
        if not chain:
            assert 7 not in kinds
    if spec["meas"][0][0] == "probs" and name == "rand":   # TN branch keeps the listed qubit order, SV sorts ascending
        ref = np.stack([np.transpose(ref[:, 0], (0, 2, 1)), ref[:, 1]], 1)
    assert_close(outs[0], ref, TOL[dt], "apply steps")
    assert_close(outs[1], ref, TOL[dt], "chain runs")
    assert_close(outs[1], outs[0], TOL[dt] * 0.1, "chain vs apply")
