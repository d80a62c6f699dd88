"""GPU parity tests proper: the CUDA engine (through the C ABI) against fixtures produced by the reference
and against the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

import tedq_b200 as qb
from conftest import case_id, load_golden
from helpers import TOL, all_kinds_spec, assert_close, build, cdtype, golden_out, rdtype
from tedq_b200 import workloads as W

pytestmark = pytest.mark.gpu
CASES = load_golden("sv_cases.json")


def run_engine(case, plan_opts=None, via="batched"):
    dt = case["dtype"]
    flat0 = case["flat"][0] if case["flat"] and case["flat"][0] else None
    cc = build(case, dt, flat0).compilecircuit(backend="pytorch_b200", dtype=cdtype(dt), plan_opts=plan_opts)
    flat = torch.tensor(case["flat"], dtype=rdtype(dt), device="cuda").reshape(len(case["flat"]), -1)
    if flat.shape[1] == 0:
        out = cc()
        return out.unsqueeze(0).cpu().numpy(), None
    flat.requires_grad_(True)
    out = cc.batched(flat)
    grad = None
    if case["cotangent"]:
        ct = torch.tensor(case["cotangent"], dtype=rdtype(dt), device="cuda")
        if out.is_complex():
            loss = torch.sum(torch.view_as_real(out) * ct)
        else:
            loss = torch.sum(out * ct)
        loss.backward()
        grad = flat.grad.cpu().numpy()
    return out.detach().cpu().numpy(), grad


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_engine_matches_reference_fixture(case):
    out, grad = run_engine(case)
    assert_close(out, golden_out(case), TOL[case["dtype"]], "out")
    if grad is not None:
        assert_close(grad, np.asarray(case["grad"]), TOL[case["dtype"]], "grad")


TILED = [c for c in CASES if 5 <= c["spec"]["num_qubits"] <= 12]


@pytest.mark.parametrize("case", TILED, ids=case_id)
@pytest.mark.parametrize("m", [(5, 5, 2), (6, 4, 1)])
def test_tiled_sweeps_match_reference_fixture(case, m):
    """Force the HBM-tiled forward/backward sweeps on small circuits (tile of 2^m amplitudes)."""
    if case["spec"]["num_qubits"] <= max(m[0], m[1]):
        pytest.skip("state fits one tile")
    opts = {"max_local_qubits_fwd": m[0], "max_local_qubits_bwd": m[1], "coalesce_bits": m[2]}
    out, grad = run_engine(case, opts)
    assert_close(out, golden_out(case), TOL[case["dtype"]], "out")
    if grad is not None:
        assert_close(grad, np.asarray(case["grad"]), TOL[case["dtype"]], "grad")


def test_reference_style_call_and_backward():
    """Reads like test/test_pytorch_backend.py:386-584, on CUDA tensors."""
    def circuitDef(*params):
        qb.RX(params[0], qubits=[0])
        qb.RY(params[1], qubits=[0])
        return [qb.expval(qb.PauliZ(qubits=[0])), qb.expval(qb.PauliX(qubits=[1]))]

    for method in ("back_prop", "param_shift"):
        a = torch.tensor([0.54], dtype=torch.float32, requires_grad=True, device="cuda")
        b = torch.tensor([0.12], dtype=torch.float32, requires_grad=True, device="cuda")
        circuit = qb.Circuit(circuitDef, 2, a, b)
        cc = circuit.compilecircuit(backend="pytorch_b200", diff_method=method)
        res = cc(a, b)
        assert res.shape == (2,)
        assert np.array_equal(np.round(res.detach().cpu().numpy(), 5), np.round(np.float32([0.85154057, 0.0]), 5))
        res[0].backward()
        assert np.array_equal(np.round(a.grad.cpu().numpy(), 5), np.round(np.float32([-0.5104387]), 5))
        assert np.array_equal(np.round(b.grad.cpu().numpy(), 5), np.round(np.float32([-0.10267819]), 5))
        a2 = torch.tensor([0.54], requires_grad=True, device="cuda")
        b2 = torch.tensor([0.12], requires_grad=True, device="cuda")
        cc(a2, b2)[1].backward()
        assert np.array_equal(np.round(a2.grad.cpu().numpy(), 5), np.float32([0.0]))
        assert np.array_equal(np.round(b2.grad.cpu().numpy(), 5), np.float32([0.0]))


def test_param_shift_four_term_gates():
    spec = W.random_circuit(4, 24, seed=11, gate_pool=["CRX", "CRY", "CRZ", "RX", "Hadamard", "ControlledPhaseShift",
                                                        "PhaseShift", "RZ", "Rot"],
                            meas=[["expval", [["PauliZ", [0]]]], ["expval", [["PauliX", [2]]]]])
    circ = W.build_circuit(spec, qb)
    x = torch.rand(spec["n_params"], device="cuda") * 3
    grads = []
    for method in ("back_prop", "param_shift"):
        cc = circ.compilecircuit(backend="pytorch_b200", diff_method=method)
        xx = x.clone().requires_grad_(True)
        (cc(xx) * torch.tensor([0.7, -1.3], device="cuda")).sum().backward()
        grads.append(xx.grad.cpu().numpy())
    assert_close(grads[1], grads[0], 2e-5, "param-shift vs adjoint")


def test_vmap_and_shared_parameters():
    """C1 usage: vmap over data rows with shared weights (Hessian_&_batch notebook cells 19-22)."""
    spec = W.qnn4()
    circ = W.build_circuit(spec, qb)
    cc = circ.compilecircuit(backend="pytorch_b200")
    X = torch.rand(8, 4, device="cuda")
    w = torch.rand(2, 4, 2, device="cuda", requires_grad=True)
    y_v = torch.func.vmap(lambda x: cc(x, w))(X)
    y_b = cc.batched(X, w, in_dims=(0, None))
    y_l = torch.stack([cc(X[i], w) for i in range(8)])
    assert torch.equal(y_v, y_b)
    assert torch.allclose(y_v, y_l, atol=1e-6)
    y_v.sum().backward()
    g_v = w.grad.clone()
    w.grad = None
    y_l.sum().backward()
    assert torch.allclose(g_v, w.grad, atol=1e-5)


def test_errors_match_reference():
    def circuitDef(*params):
        qb.RY(params[0], qubits=[0])
        qb.RZ(params[1], qubits=[1])
        return qb.expval(qb.PauliZ(qubits=[0]))

    circuit = qb.Circuit(circuitDef, 2, 0.3, 0.4)
    with pytest.raises(ValueError):
        circuit.compilecircuit(backend="pytorch_b200", interface="jax")
    with pytest.raises(ValueError, match="Error!!!! can not use contengra, opt_einsum and cyc at the same time!"):
        circuit.compilecircuit(backend="pytorch_b200", use_cotengra=True, use_jdopttn=True)
    with pytest.raises(ValueError):
        circuit.compilecircuit(backend="nope")
    cc = circuit.compilecircuit(backend="pytorch_b200")
    a = torch.tensor([0.1], device="cuda")
    with pytest.raises(ValueError, match="number of parameters are not matched"):
        cc(a)
    with pytest.raises(ValueError, match="must be type of pytorch tensor"):
        cc(0.1, 0.2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cc(torch.tensor([0.1]), torch.tensor([0.2]))
    cc2 = circuit.compilecircuit(backend="pytorch_b200", diff_method="param_shift")
    with pytest.raises(ValueError, match="must have the same data type"):
        cc2(a, torch.tensor([0.1], device="cuda", dtype=torch.float64))
    cc3 = circuit.compilecircuit(backend="pytorch_b200", diff_method="arbitrary")
    with pytest.raises(Exception, match="is not supported"):
        cc3(a, a)


def test_host_entry_point_matches_device_path():
    case = [c for c in CASES if c["spec"]["name"].startswith("mbl1d_8")][0]
    cc = build(case).compilecircuit(backend="pytorch_b200")
    flat = np.asarray(case["flat"], dtype=np.float32)
    ct = np.asarray(case["cotangent"], dtype=np.float32)
    out, grad = cc.execute_host(flat, ct.reshape(len(flat), -1))
    assert_close(out, golden_out(case), 1e-5, "out")
    assert_close(grad, np.asarray(case["grad"]), 1e-5, "grad")


def test_user_initial_state():
    rng = np.random.RandomState(3)
    v = rng.randn(8) + 1j * rng.randn(8)
    v /= np.linalg.norm(v)

    def circuitDef(t):
        qb.InitStateVector(v)
        qb.RX(t[0], qubits=[1])
        qb.CNOT(qubits=[1, 2])
        return qb.state()

    from oracle import sv_ref
    circ = qb.Circuit(circuitDef, 3, torch.tensor([0.3]))
    cc = circ.compilecircuit(backend="pytorch_b200")
    got = cc(torch.tensor([0.3], device="cuda")).cpu().numpy()
    ref = sv_ref.run_sv(circ, torch.tensor([0.3])).numpy()
    assert_close(got, ref, 1e-6, "state")


def test_states_after_measurement_match_reference():
    """probs(qubits, after_state=True): the post-measurement state dictionary of pytorch_backend.py:474-493
    (golden produced by the unmodified reference, tests/golden/after_state.json)."""
    g = load_golden("after_state.json")

    def circuit_def(a, b, c):
        qb.RX(a, qubits=[0]); qb.RY(b, qubits=[1]); qb.CNOT(qubits=[0, 2]); qb.RZ(c, qubits=[2])
        qb.Hadamard(qubits=[1]); qb.CNOT(qubits=[1, 2])
        getattr(qb, "measurement", qb).probs(qubits=[0, 2], after_state=True)

    pa = [torch.tensor(v) for v in g["params"]]
    cc = qb.Circuit(circuit_def, 3, *pa).compilecircuit(backend="pytorch_b200")
    y = cc(*[p.cuda() for p in pa])
    assert_close(y.cpu().numpy(), np.asarray(g["probs"]), 1e-6, "probs")
    got = cc.states_after_measurement
    assert sorted(got) == sorted(g["states"])
    for key, ref in g["states"].items():
        ref = np.asarray(ref)
        arr = torch.stack(got[key]).reshape(-1).cpu().numpy()
        assert_close(arr, ref[:, 0] + 1j * ref[:, 1], 1e-6, "state after " + key)
    plain = qb.Circuit(lambda a: (qb.RX(a, qubits=[0]), getattr(qb, "measurement", qb).probs(qubits=[0])), 1, pa[0]).compilecircuit(
        backend="pytorch_b200")
    plain(pa[0].cuda())
    with pytest.raises(ValueError):
        plain.states_after_measurement


# ---------------------------------------------------------------- second order, torch.func, QUDIO front
@pytest.mark.parametrize("mode", ["sv", "tn"])
@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_hessian_matches_oracle(mode, dt):
    """Double backward through the engine (gradient node differentiated by the gate set's shift identities)
    against torch's own double backward through the oracle, every parametrised gate kind in the circuit."""
    from oracle import sv_ref

    spec = all_kinds_spec()
    circ = build(spec, dt)
    kw = {"tn_mode": True} if mode == "tn" else {}
    cc = circ.compilecircuit(backend="pytorch_b200", dtype=cdtype(dt), **kw)
    P = spec["n_params"]
    rng = np.random.RandomState(5)
    x0 = torch.tensor(rng.uniform(-3, 3, P), dtype=rdtype(dt))
    w = torch.tensor([0.7, -1.3, 0.4], dtype=rdtype(dt))
    H = torch.autograd.functional.hessian(lambda x: (cc(x) * w.cuda()).sum(), x0.cuda())
    Href = torch.autograd.functional.hessian(
        lambda x: (sv_ref.run_sv(circ, x, torch.complex128) * w.double()).sum(), x0.double())
    assert_close(H.cpu().numpy(), Href.numpy(), 10 * TOL[dt], "hessian")
    assert_close(H.cpu().numpy(), H.cpu().numpy().T, 10 * TOL[dt], "symmetry")


@pytest.mark.parametrize("kw", [{}, {"tn_mode": True}, {"use_jdopttn": "B200OptTN"}], ids=["sv", "tn", "jdopttn"])
def test_notebook_hessian_and_batch(kw):
    """Hessian_&_batch_executation notebook cells 5-30: cost(params, weight), vmap(cost), hessian(cost),
    vmap(hessian(cost)) on RX(p0) RY(p1) <Z> = cos(p0) cos(p1)."""
    def circuitDef(params):
        qb.RX(params[0], qubits=[0])
        qb.RY(params[1], qubits=[0])
        return qb.expval(qb.PauliZ(qubits=[0]))

    kw = dict(kw)
    if kw.get("use_jdopttn") == "B200OptTN":
        kw["use_jdopttn"] = qb.B200OptTN
    circuit = qb.Circuit(circuitDef, 1, parameter_shapes=[(2,)])
    cc = circuit.compilecircuit(backend="pytorch_b200", **kw)

    def cost(params, weight):
        return weight[0] * cc(params) + weight[1] + weight[2]

    torch.manual_seed(3)
    P = torch.rand(5, 2, device="cuda")
    Wt = torch.rand(5, 3, device="cuda")
    exact = lambda p, w: w[0] * (torch.cos(p[0]) * torch.cos(p[1])).reshape(1) + w[1] + w[2]
    assert_close(torch.func.vmap(cost)(P, Wt).cpu(), torch.func.vmap(exact)(P, Wt).cpu(), 1e-5, "vmap(cost)")
    h = torch.func.hessian(cost)(P[0], Wt[0])
    assert h.shape == (1, 2, 2)
    assert_close(h.cpu(), torch.func.hessian(exact)(P[0], Wt[0]).cpu(), 2e-5, "hessian")
    hb = torch.func.vmap(torch.func.hessian(cost))(P, Wt)
    assert_close(hb.cpu(), torch.func.vmap(torch.func.hessian(exact))(P, Wt).cpu(), 2e-5, "vmap(hessian)")
    jb = torch.func.vmap(torch.func.jacrev(cost, argnums=(0, 1)))(P, Wt)
    je = torch.func.vmap(torch.func.jacrev(exact, argnums=(0, 1)))(P, Wt)
    for a, b in zip(jb, je):
        assert_close(a.cpu(), b.cpu(), 2e-5, "vmap(jacrev)")
    jf = torch.func.jacfwd(cost)(P[1], Wt[1])
    assert_close(jf.cpu(), torch.func.jacfwd(exact)(P[1], Wt[1]).cpu(), 2e-5, "jacfwd")


def test_retain_graph_and_repeated_backward():
    spec = W.qnn4()
    cc = W.build_circuit(spec, qb).compilecircuit(backend="pytorch_b200")
    x = torch.rand(4, device="cuda")
    w = torch.rand(2, 4, 2, device="cuda", requires_grad=True)
    y = cc(x, w).sum()
    (g1,) = torch.autograd.grad(y, w, retain_graph=True)
    (g2,) = torch.autograd.grad(y, w)
    assert torch.allclose(g1, g2, atol=1e-6)


def test_qudio_front_matches_row_loop():
    """qudio_backend.py:86-111: set_dataset(d) then cc(params) == cat over rows of cc_plain(d[i], params)."""
    spec = W.qnn4()
    circ = W.build_circuit(spec, qb)
    cq = circ.compilecircuit(backend="pytorch_QUDIO_b200")
    cp = circ.compilecircuit(backend="pytorch_b200")
    torch.manual_seed(0)
    d = torch.rand(6, 4)                                   # host dataset, like the example script
    w = torch.rand(2, 4, 2, device="cuda", requires_grad=True)
    with pytest.raises(ValueError):
        cq(w)
    cq.set_dataset(d)
    assert cq.dataset() is d
    y = cq(w)
    rows = torch.cat([cp(d[i].cuda(), w) for i in range(6)], 0)
    assert y.shape == rows.shape == (24,)
    assert torch.allclose(y, rows, atol=1e-6)
    ct = torch.rand(24, device="cuda")
    (g,) = torch.autograd.grad((y * ct).sum(), w)
    (gr,) = torch.autograd.grad((rows * ct).sum(), w)
    assert torch.allclose(g, gr, atol=1e-5)
    assert str(cq.device).startswith("cuda")


def _c4d20_case():
    case = load_golden("c4d20_case.json")
    spec = W.mbl_2d(*case["spec"]["args"])
    assert len(spec["gates"]) == case["spec"]["n_gates"] and spec["n_params"] == case["spec"]["n_params"]
    return dict(case, spec=spec)


def test_c4_depth20_matches_reference_fixture():
    """BASELINE config 4 at "depth 20" (10 Hd + 10 H0 Trotter sweeps of the 4x4 MBL-2D circuit: 17 819 gates,
    complex128): values and gradients against the unmodified reference (tests/golden/generate_golden_c4d20.py)."""
    case = _c4d20_case()
    out, grad = run_engine(case)
    assert_close(out, golden_out(case), TOL["c128"], "c4d20 out")
    assert_close(grad, np.asarray(case["grad"]), TOL["c128"], "c4d20 grad")


def _state_and_ops(circuit_def, n, params):
    """Oracle state of a traced circuit (torch, differentiable) for observables the oracle itself does not know."""
    from oracle import sv_ref

    circ = qb.Circuit(circuit_def, n, *params)
    flat = torch.cat([p.reshape(-1) for p in params]).detach().cpu().double().requires_grad_(True)
    psi = sv_ref.run_sv(circ, flat, torch.complex128, return_state=True).reshape(-1)
    return circ, flat, psi


def test_var_sample_and_user_unitary():
    """f4: ``var`` and ``sample`` (declared, NotImplemented in the reference: measurement.py:158-171) and the
    user-defined ``Unitary`` gate / Hermitian observable (qubit.py:1696-1730, broken at :1719).  var = <O^2> - <O>^2
    with gradients, against the oracle's state; samples are eigenvalues whose mean converges to <O>."""
    rng = np.random.RandomState(5)
    h = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    herm = (h + h.conj().T) / 2                     # 2-qubit Hermitian observable
    ang = 0.37
    u1 = np.array([[np.cos(ang / 2), -1j * np.sin(ang / 2)], [-1j * np.sin(ang / 2), np.cos(ang / 2)]])   # = RX(0.37)

    def body(a, b, c):
        qb.RY(a, qubits=[0]); qb.RX(b, qubits=[1]); qb.CNOT(qubits=[0, 1]); qb.CRZ(c, qubits=[1, 2])
        qb.Unitary(u1, qubits=[2]); qb.Hadamard(qubits=[0])

    def with_var(a, b, c):
        body(a, b, c)
        return [qb.var(qb.PauliZ(qubits=[0])), qb.var(qb.Unitary(herm, qubits=[1, 2])),
                qb.expval(qb.PauliX(qubits=[1])), qb.var([qb.PauliZ(qubits=[0]), qb.PauliX(qubits=[2])])]

    params = [torch.tensor([v], dtype=torch.float64, device="cuda", requires_grad=True) for v in (0.54, -0.8, 1.3)]
    for kw in ({}, {"tn_mode": True, "tn_simplify": False, "hyper_opt": {"max_repeats": 2, "tn_backward": "tree"}}):
        for p in params:
            p.grad = None
        cc = qb.Circuit(with_var, 3, *params).compilecircuit(backend="pytorch_b200", dtype=torch.complex128, **kw)
        out = cc(*params)
        assert out.shape == (4,) and len(cc.measurements) == 4
        w = torch.tensor([0.3, -1.1, 0.7, 0.5], dtype=torch.float64, device="cuda")
        (out * w).sum().backward()
        # oracle: the same circuit with RX(0.37) in place of the user unitary, observables applied in numpy / torch
        def ref_def(a, b, c):
            qb.RY(a, qubits=[0]); qb.RX(b, qubits=[1]); qb.CNOT(qubits=[0, 1]); qb.CRZ(c, qubits=[1, 2])
            qb.RX(torch.tensor(ang, dtype=torch.float64), qubits=[2], trainable_params=[]); qb.Hadamard(qubits=[0])
            return qb.state()
        _, flat, psi = _state_and_ops(ref_def, 3, [p.detach().cpu() for p in params])
        Z, X, I2 = np.diag([1.0, -1.0]), np.array([[0, 1.0], [1.0, 0]]), np.eye(2)
        ops = [np.kron(np.kron(Z, I2), I2), np.kron(I2, herm), np.kron(np.kron(I2, X), I2), np.kron(np.kron(Z, I2), X)]
        vals = []
        for j, O in enumerate(ops):
            Ot = torch.tensor(O, dtype=torch.complex128)
            e1 = torch.real(torch.vdot(psi, Ot @ psi))
            e2 = torch.real(torch.vdot(psi, Ot @ (Ot @ psi)))
            vals.append(e1 if j == 2 else e2 - e1 ** 2)
        ref = torch.stack(vals)
        (ref * w.cpu()).sum().backward()
        assert_close(out.detach().cpu().numpy(), ref.detach().numpy(), 1e-11, "var")
        got_g = np.array([float(p.grad) for p in params])
        assert_close(got_g, flat.grad.numpy(), 1e-10, "var grad")

    def with_sample(a, b, c):
        body(a, b, c)
        return [qb.sample(qb.PauliZ(qubits=[1]), 20000), qb.sample(qb.Unitary(herm, qubits=[0, 2]), 20000)]

    def with_expval(a, b, c):
        body(a, b, c)
        return [qb.expval(qb.PauliZ(qubits=[1])), qb.expval(qb.Unitary(herm, qubits=[0, 2]))]

    torch.manual_seed(0)
    p32 = [p.detach().float() for p in params]
    smp = qb.Circuit(with_sample, 3, *p32).compilecircuit(backend="pytorch_b200")(*p32)
    exp = qb.Circuit(with_expval, 3, *p32).compilecircuit(backend="pytorch_b200")(*p32).cpu().numpy()
    assert smp.shape == (2, 20000) and not smp.requires_grad
    lam = np.linalg.eigvalsh(herm)
    s = smp.cpu().numpy()
    assert set(np.unique(s[0])) <= {-1.0, 1.0}
    assert all(np.abs(lam - v).min() < 1e-5 for v in np.unique(s[1]))
    for j, spread in ((0, 1.0), (1, float(lam.max() - lam.min()))):
        assert abs(s[j].mean() - exp[j]) < 5 * spread / np.sqrt(20000), (j, s[j].mean(), exp[j])
    with pytest.raises(ValueError):      # measurements of different shapes cannot be stacked (pytorch_backend.py:385-389)
        def mixed(a, b, c):
            body(a, b, c)
            return [qb.sample(qb.PauliZ(qubits=[1]), 10), qb.expval(qb.PauliZ(qubits=[0]))]
        qb.Circuit(mixed, 3, *p32).compilecircuit(backend="pytorch_b200")(*p32)


STRUCT_CASES = [c for c in CASES if c["spec"]["n_params"] > 0 and c["spec"]["num_qubits"] >= 3][::3] + \
    [c for c in CASES if c["spec"]["name"] in ("hea20_d10", "mbl1d_12", "mbl2d_4x4_s1", "hea16_d3", "qnn4")]


@pytest.mark.parametrize("case", STRUCT_CASES, ids=case_id)
@pytest.mark.parametrize("tiled", [False, True], ids=["default", "tiled"])
def test_structure_aware_ops_match_reference_fixture(case, tiled):
    """plan_opts["structure"] = 1 (experimental): blocks of real gates on the real-matrix paths (P_R1*/P_R2*), one-qubit
    diagonal gates merged into diagonal-layer passes (P_DL: two phase tables; gradients from signed sums).  Same
    values and gradients as the reference, in the whole-state kernels and in forced small tiles."""
    n = case["spec"]["num_qubits"]
    opts = {"structure": 1}
    if tiled:
        if n <= 5:
            pytest.skip("state fits one tile")
        opts.update({"max_local_qubits_fwd": 5, "max_local_qubits_bwd": 5, "coalesce_bits": 2})
    out, grad = run_engine(case, opts)
    assert_close(out, golden_out(case), TOL[case["dtype"]], "out")
    if grad is not None:
        assert_close(grad, np.asarray(case["grad"]), TOL[case["dtype"]], "grad")
    cc = build(case, case["dtype"], case["flat"][0]).compilecircuit(backend="pytorch_b200", dtype=cdtype(case["dtype"]),
                                                                     plan_opts=opts)
    names = {g[0] for g in case["spec"]["gates"]}
    if names & {"RZ", "PhaseShift", "S", "T", "PauliZ"} and n >= 4 and not tiled:
        assert cc.plan().op_stats(True)[3] + cc.plan().op_stats(True)[1] + cc.plan().op_stats(False)[1] >= 0
