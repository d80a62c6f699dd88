"""CPU: the autograd nodes of backend.py (B200Execute / B200Grad / B200Hess / B200Jac) composed with
torch.autograd and torch.func transforms, on a stand-in engine.

The stand-in replaces ONLY the two device primitives (``_values`` / ``_vjp``: on a GPU they are the C-ABI calls)
with closed-form trigonometric polynomials of degree one per parameter — what every expectation value of the
reference's gate set is — so that the host logic (vmap folding, shift tables, double-backward wiring; reference
usage: Hessian_&_batch_executation notebook cells 17-30) is covered without a GPU.  The GPU twin of these tests
is tests/test_engine_gpu.py::test_hessian_*."""
import math

import pytest
import torch

from tedq_b200 import backend as backend_mod


def analytic(flat):
    """[B, 3] -> [B, 2]"""
    a, b, c = flat[:, 0], flat[:, 1], flat[:, 2]
    return torch.stack([torch.cos(a) * torch.cos(b) + 0.3 * torch.sin(c) * torch.sin(a),
                        torch.sin(a + 0.2) + torch.cos(c) * torch.sin(b)], 1)


class StandIn(backend_mod.B200Backend):
    def __init__(self):          # no circuit, no plan: only what the autograd layer reads
        self._tn = None
        self._requires_grad = True
        self._rdtype = torch.float64
        self._res_complex = False
        self.calls = 0

    def _param_gate_names(self):
        return ["RX", "RY", "RZ"]

    def _values(self, flat, need_grad):
        self.calls += 1
        return analytic(flat.detach()), ("state" if need_grad else None)

    def _vjp(self, flat, dy, state):
        assert state == "state"
        with torch.enable_grad():
            f = flat.detach().requires_grad_(True)
            (g,) = torch.autograd.grad(analytic(f), f, dy.detach())
        return g


def f_engine(be):
    return lambda th: be._run(th.reshape(1, -1).to(torch.float64), backend_mod.B200Execute)[0]


def f_exact(th):
    return analytic(th.reshape(1, -1))[0]


TH = torch.tensor([0.37, -1.1, 2.4], dtype=torch.float64)


def test_first_order_and_fast_path():
    be = StandIn()
    th = TH.clone().requires_grad_(True)
    w = torch.tensor([0.7, -1.3], dtype=torch.float64)
    (f_engine(be)(th) * w).sum().backward()
    t2 = TH.clone().requires_grad_(True)
    (f_exact(t2) * w).sum().backward()
    assert torch.allclose(th.grad, t2.grad, atol=1e-12)
    assert be.calls == 1                      # the reverse pass used the kept state, no second forward


def test_no_grad_is_plain_values():
    be = StandIn()
    with torch.no_grad():
        y = f_engine(be)(TH)
    assert not y.requires_grad and torch.allclose(y, f_exact(TH))


def test_retain_graph_second_backward():
    be = StandIn()
    th = TH.clone().requires_grad_(True)
    y = f_engine(be)(th).sum()
    (g1,) = torch.autograd.grad(y, th, retain_graph=True)
    (g2,) = torch.autograd.grad(y, th)
    assert torch.allclose(g1, g2, atol=1e-12)


def test_double_backward_hessian():
    be = StandIn()
    w = torch.tensor([0.7, -1.3], dtype=torch.float64)
    H = torch.autograd.functional.hessian(lambda t: (f_engine(be)(t) * w).sum(), TH)
    Hx = torch.autograd.functional.hessian(lambda t: (f_exact(t) * w).sum(), TH)
    assert torch.allclose(H, Hx, atol=1e-10)


def test_gradient_penalty_through_dy():
    """d/dw of |grad_theta (w . f)|^2 needs the derivative of the gradient node with respect to dy (the Jacobian)."""
    be = StandIn()

    def pen(fn, w):
        th = TH.clone().requires_grad_(True)
        (g,) = torch.autograd.grad((fn(th) * w).sum(), th, create_graph=True)
        return (g ** 2).sum()

    w1 = torch.tensor([0.7, -1.3], dtype=torch.float64, requires_grad=True)
    w2 = w1.detach().clone().requires_grad_(True)
    pen(f_engine(be), w1).backward()
    pen(f_exact, w2).backward()
    assert torch.allclose(w1.grad, w2.grad, atol=1e-10)


def test_func_transforms():
    be = StandIn()
    fe = f_engine(be)
    for tr in (torch.func.jacrev, torch.func.jacfwd, torch.func.hessian):
        assert torch.allclose(tr(fe)(TH), tr(f_exact)(TH), atol=1e-10), tr.__name__
    g = torch.func.grad(lambda t: fe(t)[0])(TH)
    assert torch.allclose(g, torch.func.grad(lambda t: f_exact(t)[0])(TH), atol=1e-12)


def test_vmap_of_transforms():
    be = StandIn()
    fe = f_engine(be)
    TB = torch.stack([TH, TH * 0.5 + 0.1, -TH, TH + 1.0])
    assert torch.allclose(torch.func.vmap(fe)(TB), torch.func.vmap(f_exact)(TB), atol=1e-12)
    for tr in (torch.func.jacrev, torch.func.hessian):
        got = torch.func.vmap(tr(fe))(TB)
        want = torch.func.vmap(tr(f_exact))(TB)
        assert torch.allclose(got, want, atol=1e-10), tr.__name__
    # Hessian_&_batch notebook cells 17-30: cost(params, weight), vmapped hessian with respect to params
    W = torch.rand(4, 3, dtype=torch.float64)

    def cost(fn):
        return lambda p, w: w[0] * fn(p) + w[1] + w[2]

    got = torch.func.vmap(torch.func.hessian(cost(fe)))(TB, W)
    want = torch.func.vmap(torch.func.hessian(cost(f_exact)))(TB, W)
    assert torch.allclose(got, want, atol=1e-10)


def test_four_term_rule_rows():
    be = StandIn()
    be._param_gate_names = lambda: ["CRX", "RY"]
    rows = be._shift_rows()
    assert [j for j, _, _ in rows] == [0, 0, 0, 0, 1, 1]
    # the four-term rule differentiates cos(t/2) and cos(t) terms exactly (controlled rotations, qubit.py:37-42)
    for fn, dfn in ((lambda t: math.cos(t / 2), lambda t: -0.5 * math.sin(t / 2)), (math.cos, lambda t: -math.sin(t))):
        got = sum(c * fn(0.83 + s) for j, c, s in rows if j == 0)
        assert abs(got - dfn(0.83)) < 1e-12


def test_third_order_is_refused():
    be = StandIn()
    th = TH.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(f_engine(be)(th).sum(), th, create_graph=True)
    (h,) = torch.autograd.grad(g.sum(), th, create_graph=True)
    with pytest.raises(NotImplementedError):
        torch.autograd.grad(h.sum(), th)


def test_shift_identities_exact_for_every_gate_kind():
    """The shift tables differentiate the oracle's own circuit function exactly (every parametrised gate kind,
    controls live): Hessian and Jacobian from shifted oracle evaluations == torch double backward through it."""
    import numpy as np
    from oracle import sv_ref
    from helpers import all_kinds_spec, build

    spec = all_kinds_spec()
    circ = build(spec, "c128")

    class OracleEngine(StandIn):
        def _param_gate_names(self):
            return [g[0] for g in spec["gates"] for p in g[2] if isinstance(p, str)]

        def _values(self, flat, need_grad):
            out = torch.stack([sv_ref.run_sv(circ, f, torch.complex128) for f in flat.detach()])
            return out, ("state" if need_grad else None)

        def _vjp(self, flat, dy, state):
            with torch.enable_grad():
                f = flat.detach().requires_grad_(True)
                out = torch.stack([sv_ref.run_sv(circ, r, torch.complex128) for r in f])
                (g,) = torch.autograd.grad(out, f, dy.detach())
            return g

    be = OracleEngine()
    x0 = torch.tensor(np.random.RandomState(5).uniform(-3, 3, spec["n_params"]))
    w = torch.tensor([0.7, -1.3, 0.4], dtype=torch.float64)
    H = be._shift_hessian(x0[None], w[None])[0]
    J = be._shift_jacobian(x0[None])[0]
    Href = torch.autograd.functional.hessian(lambda x: (sv_ref.run_sv(circ, x, torch.complex128) * w).sum(), x0)
    Jref = torch.autograd.functional.jacobian(lambda x: sv_ref.run_sv(circ, x, torch.complex128), x0)
    assert torch.allclose(H, Href, atol=1e-12)
    assert torch.allclose(J, Jref, atol=1e-12)
