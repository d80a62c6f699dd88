"""``B200Backend`` — the ``backend="pytorch_b200"`` compiled circuit.

Host-side mirror of the reference's ``PyTorchBackend`` / ``CompiledCircuit`` interface
(tedq/backends/pytorch_backend.py:44-227, compiled_circuit.py:34-120, :549-647): same constructor
keywords, same positional parameter binding, same error types and messages, same stacked result
tensor, same autograd calling convention (``Function.apply(run_kwargs, tensors)``; ``backward``
returns ``(None, grads)``; pytorch_backend.py:1191-1223).  Everything numeric happens in
libtedq_b200.so (hand-written sm_100a CUDA reached through the C ABI in include/tedq_b200.h);
there is no CPU path: CPU tensors are rejected.

Additions the reference does not have (SURVEY.md 8b "Batch"):
  * ``batched(*params, in_dims=...)``: explicit batch of parameter sets in ONE launch sequence;
  * ``torch.func.vmap`` over ``__call__`` works (the Function carries a vmap rule);
  * ``dtype=torch.complex128`` keyword (the reference hard-codes ``tcomplex``, :38);
  * ``execute_host(params_np)``: HOST buffers in/out through ``tq_execute_host``.
"""
from __future__ import annotations

import dataclasses

import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import capi
from .ir import MEAS_PROBS, MEAS_STATE, CircuitIR, MeasRec, build_ir

BACKEND_NAME = "pytorch_b200"

_FOUR_TERM = {"CRX", "CRY", "CRZ"}  # qubit.py:37-42 CONTROL_GRAD_RECIPE


def _shift_recipe(name: str):
    """[(coefficient, multiplier, shift)] as ops_abc.py:261-280 / qubit.py:37-42."""
    if name in _FOUR_TERM:
        s2 = 1.0 / math.sqrt(2.0)
        m1 = s2 * (math.sqrt(2.0) + 1.0) / 4.0
        m2 = s2 * (math.sqrt(2.0) - 1.0) / 4.0
        return [(m1, 1.0, math.pi / 2), (-m1, 1.0, -math.pi / 2), (-m2, 1.0, 3 * math.pi / 2), (m2, 1.0, -3 * math.pi / 2)]
    return [(0.5, 1.0, math.pi / 2), (-0.5, 1.0, -math.pi / 2)]


def _is_functorch_wrapped(t) -> bool:
    try:
        return bool(torch._C._functorch.is_functorch_wrapped_tensor(t))
    except AttributeError:  # pragma: no cover - older torch
        return True


def _fold_vmapped(info, in_dims, *tensors):
    """vmap rule shared by every node: the ops are natively batched over dim 0, so the vmapped dimension is
    folded into it ([V, B, ...] -> [V*B, ...]; an un-vmapped operand is repeated V times)."""
    V = info.batch_size
    out = []
    for t, d in zip(tensors, in_dims):
        t = t.unsqueeze(0).expand((V,) + tuple(t.shape)) if d is None else t.movedim(d, 0)
        out.append(t.reshape((V * t.shape[1],) + tuple(t.shape[2:])))
    return out, V


def _unfold(t, V):
    return t.reshape((V, t.shape[0] // V) + tuple(t.shape[1:]))


class B200Execute(torch.autograd.Function):
    """Values of every measurement.  ``apply(run_kwargs, flat)`` with ``flat`` = [B, P] real.

    Calling convention follows TorchExecute (pytorch_backend.py:1191-1223): first argument is a dict
    carrying the backend, ``backward`` returns ``(None, grads)``.  The gradient is the engine's reverse pass
    (adjoint sweeps, or reverse mode through the contraction trees in tensor-network mode) on the state the
    forward call kept.  When the gradient itself has to be differentiable (``create_graph=True``,
    ``torch.func.hessian``/``jacrev``/``jacfwd``: Hessian_&_batch_executation notebook cells 24-30) ``backward``
    returns the ``B200Grad`` node instead, whose own derivatives come from the parameter-shift identities of the
    gate set (ops_abc.py:261-280, qubit.py:37-42) evaluated in batched launches."""

    @staticmethod
    def forward(run_kwargs, flat):
        backend = run_kwargs["backend"]
        out, ws = backend._values(flat, run_kwargs["need_grad"])
        run_kwargs["_ws"] = ws
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        run_kwargs, flat = inputs
        ctx.backend = run_kwargs["backend"]
        ctx.ws = run_kwargs.pop("_ws", None)
        ctx.save_for_backward(flat)
        ctx.save_for_forward(flat)

    @staticmethod
    def backward(ctx, dy):
        (flat,) = ctx.saved_tensors
        be = ctx.backend
        plain = not (_is_functorch_wrapped(flat) or _is_functorch_wrapped(dy))
        if plain and ctx.ws is not None and not (torch.is_grad_enabled() and (dy.requires_grad or flat.requires_grad)):
            grad = be._vjp(flat, dy, ctx.ws)
            ctx.ws = None
            return None, grad
        return None, B200Grad.apply({"backend": be}, flat, dy)

    @staticmethod
    def jvp(ctx, _, t_flat):
        (flat,) = ctx.saved_tensors
        jac = B200Jac.apply({"backend": ctx.backend}, flat)              # [B, n_meas, ..., P]
        return torch.einsum("b...p,bp->b...", jac, t_flat.to(jac.dtype))

    @staticmethod
    def vmap(info, in_dims, run_kwargs, flat):
        if in_dims[1] is None:
            return B200Execute.apply(run_kwargs, flat), None
        (flat,), V = _fold_vmapped(info, in_dims[1:], flat)
        return _unfold(B200Execute.apply(dict(run_kwargs), flat), V), 0


class B200Grad(torch.autograd.Function):
    """(flat [B, P], dy [B, n_meas, ...]) -> dy . J(flat)  [B, P], as a differentiable node: a forward pass and the
    engine's reverse pass.  Its derivatives: d/dflat = the dy-weighted Hessian (``B200Hess``), d/ddy = J (``B200Jac``)."""

    @staticmethod
    def forward(run_kwargs, flat, dy):
        return run_kwargs["backend"]._vjp_fresh(flat, dy)

    @staticmethod
    def setup_context(ctx, inputs, output):
        run_kwargs, flat, dy = inputs
        ctx.backend = run_kwargs["backend"]
        ctx.save_for_backward(flat, dy)
        ctx.save_for_forward(flat, dy)

    @staticmethod
    def backward(ctx, u):
        flat, dy = ctx.saved_tensors
        rk = {"backend": ctx.backend}
        g_flat = g_dy = None
        if ctx.needs_input_grad[1]:
            hess = B200Hess.apply(rk, flat, dy)                          # [B, P, P], symmetric
            g_flat = torch.einsum("bjk,bj->bk", hess, u.to(hess.dtype))
        if ctx.needs_input_grad[2]:
            jac = B200Jac.apply(rk, flat)
            g_dy = torch.einsum("b...p,bp->b...", jac, u.to(jac.dtype)).to(dy.dtype)
        return None, g_flat, g_dy

    @staticmethod
    def jvp(ctx, _, t_flat, t_dy):
        flat, dy = ctx.saved_tensors
        rk = {"backend": ctx.backend}
        out = 0
        if t_flat is not None:
            hess = B200Hess.apply(rk, flat, dy)
            out = out + torch.einsum("bjk,bk->bj", hess, t_flat.to(hess.dtype))
        if t_dy is not None:
            out = out + B200Grad.apply(rk, flat, t_dy)
        return out

    @staticmethod
    def vmap(info, in_dims, run_kwargs, flat, dy):
        (flat, dy), V = _fold_vmapped(info, in_dims[1:], flat, dy)
        return _unfold(B200Grad.apply(run_kwargs, flat, dy), V), 0


class B200Hess(torch.autograd.Function):
    """(flat, dy) -> H[b, j, k] = d(dy . J)_j / dflat_k from shifted gradient evaluations (one batched launch
    sequence per chunk of shifts).  Third derivatives are not provided."""

    @staticmethod
    def forward(run_kwargs, flat, dy):
        return run_kwargs["backend"]._shift_hessian(flat, dy)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("pytorch_b200: derivatives beyond second order are not implemented")

    @staticmethod
    def vmap(info, in_dims, run_kwargs, flat, dy):
        (flat, dy), V = _fold_vmapped(info, in_dims[1:], flat, dy)
        return _unfold(B200Hess.apply(run_kwargs, flat, dy), V), 0


class B200Jac(torch.autograd.Function):
    """flat -> J[b, m, ..., p] = d out[b, m, ...] / dflat_p from shifted forward evaluations in one batched launch
    sequence (the rule of jacobian_param_shift, pytorch_backend.py:180-212)."""

    @staticmethod
    def forward(run_kwargs, flat):
        return run_kwargs["backend"]._shift_jacobian(flat)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("pytorch_b200: derivatives beyond second order are not implemented")

    @staticmethod
    def vmap(info, in_dims, run_kwargs, flat):
        if in_dims[1] is None:
            return B200Jac.apply(run_kwargs, flat), None
        (flat,), V = _fold_vmapped(info, in_dims[1:], flat)
        return _unfold(B200Jac.apply(run_kwargs, flat), V), 0


class B200ParamShift(torch.autograd.Function):
    """``diff_method="param_shift"``: same rule as jacobian_param_shift (pytorch_backend.py:180-212) but all
    shifted circuits go through the device in ONE batched launch sequence."""

    @staticmethod
    def forward(run_kwargs, flat):
        out, _ = run_kwargs["backend"]._values(flat, False)
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        run_kwargs, flat = inputs
        ctx.backend = run_kwargs["backend"]
        ctx.save_for_backward(flat)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        (flat,) = ctx.saved_tensors
        return None, ctx.backend._param_shift_vjp(flat, dy)

    @staticmethod
    def vmap(info, in_dims, run_kwargs, flat):
        d = in_dims[1]
        if d is None:
            return B200ParamShift.apply(run_kwargs, flat), None
        flat = flat.movedim(d, 0)
        lead = flat.shape[:2]
        out = B200ParamShift.apply(dict(run_kwargs), flat.reshape(lead[0] * lead[1], flat.shape[2]))
        return out.reshape(lead + out.shape[1:]), 0


class _ObsShim:
    """A fixed Hermitian observable derived from the user's (O^2 of ``var``, an eigenprojector of ``sample``)."""

    is_observable = True
    name = "Hermitian"

    def __init__(self, qubits, matrix):
        self.qubits = [int(q) for q in qubits]
        self.matrix = np.asarray(matrix, dtype=np.complex128)
        self.num_qubits = len(self.qubits)
        self.parameters, self.trainable_params = [], []


class _MeasShim:
    def __init__(self, obs):
        from .frontend import Expectation

        self.return_type, self.obs, self.qubits, self.after_state = Expectation, obs, None, False


class _CircuitShim:
    """The user's circuit with its ``var`` / ``sample`` measurements expanded into plain expectation values."""

    def __init__(self, circuit, measurements):
        self.num_qubits, self.operators, self.init_state = circuit.num_qubits, circuit.operators, circuit.init_state
        self.measurements = measurements


def _obs_matrix(obs):
    """(qubits, dense matrix) of an observable or a product of single-qubit observables."""
    if isinstance(obs, list):
        from .ir import _kron_obs

        return _kron_obs(obs, None)
    k = len(obs.qubits)
    return tuple(int(q) for q in obs.qubits), np.asarray(obs.matrix, dtype=np.complex128).reshape(2 ** k, 2 ** k)


def _expand_measurements(circuit):
    """-> (inner measurement list, recipe) or (None, None) when the circuit has no ``var`` / ``sample``.
    recipe[j] = ("copy", i) | ("var", i_O, i_O2) | ("sample", [i_projectors], eigenvalues, shots)."""
    kinds = [getattr(ms.return_type, "value", ms.return_type) for ms in circuit.measurements]
    if not any(k in ("var", "sample") for k in kinds):
        return None, None
    inner, recipe = [], []
    for ms, kind in zip(circuit.measurements, kinds):
        if kind == "var":
            qs, mat = _obs_matrix(ms.obs)
            inner += [_MeasShim(_ObsShim(qs, mat)), _MeasShim(_ObsShim(qs, mat @ mat))]
            recipe.append(("var", len(inner) - 2, len(inner) - 1))
        elif kind == "sample":
            qs, mat = _obs_matrix(ms.obs)
            if len(qs) > 4:
                raise ValueError("sample: observables on more than 4 qubits are not supported")
            if not np.allclose(mat, mat.conj().T, atol=1e-10):
                raise ValueError("sample: the observable's matrix is not Hermitian")
            lam, vec = np.linalg.eigh(mat)
            first = len(inner)
            for i in range(len(lam)):
                inner.append(_MeasShim(_ObsShim(qs, np.outer(vec[:, i], vec[:, i].conj()))))
            recipe.append(("sample", list(range(first, len(inner))), lam, int(getattr(ms, "num_shots", 1))))
        else:
            inner.append(ms)
            recipe.append(("copy", len(inner) - 1))
    return inner, recipe


class B200Backend:
    """Drop-in for ``PyTorchBackend`` on B200 (constructor: pytorch_backend.py:67-82, compiled_circuit.py:48-69)."""

    def __init__(self, backend, circuit, use_cotengra=False, use_jdopttn=False, tn_mode=False, hyper_opt=None,
                 tn_simplify=True, **kwargs):
        self._calculation_mode = False
        self._use_cotengra = use_cotengra
        self._use_jdopttn = use_jdopttn
        if use_cotengra and use_jdopttn:
            raise ValueError("Error!!!! can not use contengra, opt_einsum and cyc at the same time!")
        if (use_cotengra or use_jdopttn) and tn_mode:
            raise ValueError("Error!!!! can not use contengra, opt_einsum and cyc at the same time!")
        self._tn_mode = tn_mode
        # the reference's simplifier is broken (tensor_network.py:94): accepted, never applied
        self._tn_simplify = tn_simplify
        self._hyper_opt = hyper_opt if isinstance(hyper_opt, dict) else {}
        self._backend = backend
        # var / sample (declared but NotImplemented in the reference, measurement.py:158-171) are served by plain
        # expectation values of derived observables (<O>, <O^2>; eigenprojectors) and combined after the engine ran
        self._user_measurements = circuit.measurements
        inner, self._recipe = _expand_measurements(circuit)
        if inner is not None:
            circuit = _CircuitShim(circuit, inner)
        self._circuit = circuit
        self._num_qubits = circuit.num_qubits
        self._operators = list(circuit.operators)
        self._measurements = circuit.measurements
        self._init_state = circuit.init_state
        self._requires_grad = kwargs.get("requires_grad", True)
        self._interface = kwargs.get("interface", "pytorch")
        self._diff_method = kwargs.get("diff_method", "back_prop")
        self._plan_opts = kwargs.get("plan_opts", None)
        cdtype = kwargs.get("dtype", torch.complex64)
        if cdtype not in (torch.complex64, torch.complex128):
            raise ValueError("dtype must be torch.complex64 or torch.complex128")
        self._cdtype = cdtype
        self._rdtype = torch.float32 if cdtype == torch.complex64 else torch.float64
        if self._interface != "pytorch":
            raise ValueError(f"{self._interface}: pytroch_backend only supports pytorch interface!")
        self._tn_engine = bool(use_cotengra or use_jdopttn or tn_mode)
        if self._tn_engine and self._init_state:
            raise ValueError(
                "Error!!!! tensor network contraction mode do not support user-defined initial quantum state!")

        self._ir: CircuitIR = build_ir(circuit)
        self._axeslist = self._ir.axeslist
        self._permutationlist = self._ir.permutationlist
        shapes = {m.shape for m in self._ir.meas}
        kinds = {m.is_complex for m in self._ir.meas}
        self._shapes_ok = len(shapes) == 1 and len(kinds) == 1
        self._res_shape = next(iter(shapes)) if self._shapes_ok else None
        self._res_complex = next(iter(kinds)) if self._shapes_ok else False
        self._plans = {}          # CUDA device index -> capi.Plan (a plan's tables live on ONE device)
        self._device = None
        self._last_flat = None
        self._tn = None
        if self._tn_engine:
            from .tn_backend import TNExecutor  # contraction-plan path (tensor-network mode)

            self._tn = TNExecutor(self, self._hyper_opt)
            if self._hyper_opt.get("slicing_opts"):
                self._calculation_mode = self._hyper_opt["slicing_opts"].get("contract_parallel", False)

    # ------------------------------------------------------------------ plan
    def plan(self, device=None) -> capi.Plan:
        """The engine plan of ``device`` (default: the current CUDA device).  Plans are created and cached per
        device, inside that device's context: their descriptor tables are device memory (the reference's backend
        accepts any device on every call, pytorch_backend.py:85-119)."""
        idx = capi.device_index(device)
        plan = self._plans.get(idx)
        if plan is None:
            dt = capi.TQ_C64 if self._cdtype == torch.complex64 else capi.TQ_C128
            with capi.on_device(idx):
                plan = self._plans[idx] = capi.Plan(self._ir, dt, self._plan_opts)
        return plan

    # ------------------------------------------------------------ device side
    def _forward_device(self, flat: torch.Tensor, need_grad: bool):
        plan = self.plan(flat.device)
        B = flat.shape[0]
        flat = flat.contiguous()
        out = torch.empty((B, plan.out_reals), dtype=self._rdtype, device=flat.device)
        ws_bytes = plan.workspace_bytes(B, need_grad)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=flat.device)
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        with torch.cuda.device(flat.device):
            plan.forward(flat.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws_bytes, need_grad, stream)
        return self._shape_result(out), (ws if need_grad else None)

    def _backward_device(self, flat, dy, ws):
        plan = self.plan(flat.device)
        B = flat.shape[0]
        flat = flat.contiguous()
        if self._res_complex:
            dy = torch.view_as_real(dy.contiguous().to(self._cdtype))
        dy = dy.to(self._rdtype).reshape(B, plan.out_reals).contiguous()
        grad = torch.empty((B, plan.n_params), dtype=self._rdtype, device=flat.device)
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        with torch.cuda.device(flat.device):
            plan.backward(flat.data_ptr(), B, dy.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(), stream)
        return grad

    # ---------------------------------------------------- mode-independent value / gradient primitives
    def _values(self, flat, need_grad):
        """[B, P] -> (stacked results [B, n_meas, ...], state kept for ``_vjp`` or None)."""
        if self._tn is None:
            return self._forward_device(flat, need_grad)
        kept = [] if need_grad and self._tn.tree_backward_available() else None
        out = self._tn._forward_values(flat, kept)
        return out, (("tn", kept) if need_grad else None)

    def _vjp(self, flat, dy, state):
        """dy . J(flat) -> [B, P] from the state ``_values(flat, True)`` kept."""
        if self._tn is None:
            return self._backward_device(flat, dy, state)
        kept = state[1]
        if kept is not None:
            return self._tn.tree_backward(flat.contiguous(), dy, kept)
        if self._num_qubits > 26:
            raise NotImplementedError("gradients of sliced networks beyond 26 qubits are not implemented")
        _, ws = self._forward_device(flat, True)       # adjoint state-vector sweeps of the same engine
        return self._backward_device(flat, dy, ws)

    def _vjp_fresh(self, flat, dy):
        flat = flat.detach()
        _, state = self._values(flat, True)
        return self._vjp(flat, dy.detach(), state)

    def _shift_rows(self):
        """[(parameter, coefficient, shift)] over every trainable slot (jacobian_param_shift, :180-212)."""
        return [(j, c, s) for j, name in enumerate(self._param_gate_names()) for c, _a, s in _shift_recipe(name)]

    _SHIFT_SETS_PER_LAUNCH = 4096

    def _shifted_chunks(self, flat):
        """Yields (rows, [B * T, P] shifted parameter sets) with B * T bounded per launch sequence."""
        B, P = flat.shape
        rows = self._shift_rows()
        per = max(1, self._SHIFT_SETS_PER_LAUNCH // max(B, 1))
        for lo in range(0, len(rows), per):
            part = rows[lo:lo + per]
            T = len(part)
            shifted = flat.detach().unsqueeze(1).repeat(1, T, 1)          # [B, T, P]
            idx = torch.tensor([j for j, _, _ in part], device=flat.device)
            add = torch.tensor([s for _, _, s in part], dtype=flat.dtype, device=flat.device)
            shifted[:, torch.arange(T, device=flat.device), idx] += add
            yield part, shifted.reshape(B * T, P)

    def _shift_jacobian(self, flat):
        """J [B, n_meas, ..., P] from shifted forward evaluations."""
        if self._res_complex:
            raise NotImplementedError("forward-mode / second-order derivatives need real-valued measurements")
        B, P = flat.shape
        jac = None
        for part, shifted in self._shifted_chunks(flat):
            ev, _ = self._values(shifted, False)
            ev = ev.reshape((B, len(part)) + tuple(ev.shape[1:]))
            if jac is None:
                jac = torch.zeros(tuple(ev.shape[2:]) + (B, P), dtype=ev.dtype, device=ev.device)
            coef = torch.tensor([c for _, c, _ in part], dtype=ev.dtype, device=ev.device)
            idx = torch.tensor([j for j, _, _ in part], device=ev.device)
            jac.index_add_(-1, idx, ev.movedim(0, -1).movedim(0, -1) * coef)   # [..., B, T]
        if jac is None:      # parameter-free circuit
            ev, _ = self._values(flat.detach(), False)
            return torch.zeros(tuple(ev.shape) + (0,), dtype=ev.dtype, device=ev.device)
        return jac.movedim(-2, 0)

    def _shift_hessian(self, flat, dy):
        """H[b, j, k] = d(dy . J)_j / dflat_k from shifted gradient evaluations."""
        B, P = flat.shape
        hess = torch.zeros((B, P, P), dtype=self._rdtype, device=flat.device)
        dy = dy.detach()
        for part, shifted in self._shifted_chunks(flat):
            T = len(part)
            g = self._vjp_fresh(shifted, dy.repeat_interleave(T, dim=0)).reshape(B, T, P)
            coef = torch.tensor([c for _, c, _ in part], dtype=g.dtype, device=g.device)
            idx = torch.tensor([j for j, _, _ in part], device=g.device)
            hess.index_add_(2, idx, (g * coef[None, :, None]).transpose(1, 2))
        return hess

    def _param_shift_vjp(self, flat, dy):
        """vjp = dy @ J with J from shifted evaluations (pytorch_backend.py:180-212, :1211-1223)."""
        B, P = flat.shape
        if P == 0:
            return torch.zeros((B, 0), dtype=self._rdtype, device=flat.device)
        jac = self._shift_jacobian(flat)
        if jac.dim() != 3:
            raise ValueError("parameter shift needs scalar measurements (1-D output), as in the reference")
        return torch.einsum("bmp,bm->bp", jac, dy.to(jac.dtype).reshape(B, -1))

    def _param_gate_names(self) -> List[str]:
        names = [None] * self._ir.n_params
        for g in self._ir.gates:
            for i in g.param_idx:
                if i >= 0:
                    names[i] = g.name
        return names

    def _shape_result(self, out: torch.Tensor) -> torch.Tensor:
        if not self._shapes_ok:
            raise ValueError("You can not have multiple measurements with different shapes!!")
        B = out.shape[0]
        nm = len(self._ir.meas)
        if self._res_complex:
            return torch.view_as_complex(out.reshape((B, nm) + self._res_shape + (2,)))
        return out.reshape((B, nm) + self._res_shape)

    # --------------------------------------------------------------- user API
    def check_parameters_torch_device(self, params):
        """pytorch_backend.py:85-119 — all parameters torch tensors on ONE device; here that device must be CUDA."""
        the_same = None
        for par in params:
            try:
                index = par.device.index if par.is_cuda else -1
            except AttributeError as error:
                raise ValueError("input parameters must be type of pytorch tensor!!") from error
            if the_same is None:
                the_same = index
            if the_same != index:
                raise ValueError("input parameters are not in the same device!!")
        self._device = params[0].device if params else None

    def _flatten(self, params) -> torch.Tensor:
        """Argument order, then C order (pytorch_backend.py:233-236) -> [1, P]."""
        if params:
            flat = torch.cat([p.reshape(-1).to(self._rdtype) for p in params])
        else:
            dev = self._device or torch.device("cuda", torch.cuda.current_device())
            flat = torch.zeros(0, dtype=self._rdtype, device=dev)
        if flat.numel() != self._ir.n_params:
            raise ValueError(
                f"Error!!!! number of parameters are not matched!! required {self._ir.n_params} but {flat.numel()} are given")
        return flat.unsqueeze(0)

    def _require_cuda(self, dev):
        if dev is None:
            if not torch.cuda.is_available():
                raise RuntimeError("pytorch_b200 needs a CUDA device: there is no CPU fallback")
            return
        if dev.type != "cuda":
            raise RuntimeError("pytorch_b200 runs on CUDA tensors only: there is no CPU fallback "
                               "(move the parameters to the GPU, or use backend='pytorch')")

    def __call__(self, *params):
        self.check_parameters_torch_device(params)
        self._require_cuda(self._device)
        if self._requires_grad is False:
            with torch.no_grad():
                return self.execute(*params)
        if self._diff_method == "back_prop":
            return self.execute(*params)
        if self._diff_method == "param_shift":
            dts = [p.dtype for p in params]
            if any(a != b for a, b in zip(dts, dts[1:])):
                raise ValueError("input parameters for parameter shift method must have the same data type!")
            return self.param_shift_execute(*params)
        if self._diff_method == "finite_diff":
            return self.finite_diff_execute(*params)
        raise Exception(  # same (bare) exception type as pytorch_backend.py:162-165
            f"Differentiation method {self._diff_method} is not supported. "
            f"Supported methods include {{back_prop, param_shift, finite_diff }}")

    def execute(self, *params):
        flat = self._flatten(params)
        self._last_flat = flat
        return self._run(flat, B200Execute)[0]

    def param_shift_execute(self, *params):
        flat = self._flatten(params)
        self._last_flat = flat
        return self._run(flat, B200ParamShift)[0]

    def finite_diff_execute(self, *params):
        raise NotImplementedError

    def _run(self, flat, fn):
        wrapped = _is_functorch_wrapped(flat)
        track = torch.is_grad_enabled() and bool(self._requires_grad) and (flat.requires_grad or wrapped)
        if not track and not wrapped:
            with torch.no_grad():
                return self._combine(self._values(flat, False)[0])
        # under torch.func transforms the reverse pass re-runs the forward (B200Grad): keep no state
        return self._combine(fn.apply({"backend": self, "need_grad": track and not wrapped}, flat))

    def _combine(self, inner: torch.Tensor) -> torch.Tensor:
        """Engine results [B, n_inner] -> the user's measurements [B, n_meas, ...] (identity without var / sample):
        var = <O^2> - <O>^2 (differentiable); sample = eigenvalues drawn from the exact outcome distribution."""
        if getattr(self, "_recipe", None) is None:
            return inner
        cols = []
        for item in self._recipe:
            if item[0] == "copy":
                cols.append(inner[:, item[1]])
            elif item[0] == "var":
                cols.append(inner[:, item[2]] - inner[:, item[1]] ** 2)
            else:
                _, idx, lam, shots = item
                p = inner[:, idx].detach().clamp(min=0).to(torch.float64)
                draws = torch.multinomial(p / p.sum(1, keepdim=True), shots, replacement=True)
                cols.append(torch.as_tensor(lam, dtype=inner.dtype, device=inner.device)[draws])
        shapes = {tuple(c.shape[1:]) for c in cols}
        if len(shapes) != 1:
            raise ValueError("You can not have multiple measurements with different shapes!!")
        return torch.stack(cols, 1)

    def batched(self, *params, in_dims=None):
        """Evaluate a batch of parameter sets at once -> [B, n_meas, ...].

        ``in_dims[i]`` is 0 (leading batch dimension) or None (shared by every set), like torch.func.vmap."""
        self.check_parameters_torch_device(params)
        self._require_cuda(self._device)
        if in_dims is None:
            in_dims = (0,) * len(params)
        sizes = {p.shape[0] for p, d in zip(params, in_dims) if d is not None}
        if len(sizes) != 1:
            raise ValueError("batched(): batched parameters must share one leading size")
        B = sizes.pop()
        cols = []
        for p, d in zip(params, in_dims):
            if d is None:
                cols.append(p.reshape(1, -1).to(self._rdtype).expand(B, -1))
            elif d == 0:
                cols.append(p.reshape(B, -1).to(self._rdtype))
            else:
                cols.append(p.movedim(d, 0).reshape(B, -1).to(self._rdtype))
        flat = torch.cat(cols, dim=1)
        if flat.shape[1] != self._ir.n_params:
            raise ValueError(
                f"Error!!!! number of parameters are not matched!! required {self._ir.n_params} but {flat.shape[1]} are given")
        self._last_flat = flat[:1]
        fn = B200ParamShift if self._diff_method == "param_shift" else B200Execute
        if self._requires_grad is False:
            with torch.no_grad():
                return self._run(flat, fn)
        return self._run(flat, fn)

    def amplitude(self, bits, *params, batched=False, slice_range=None):
        """<bits| U(params) |0...0> by sliced tensor-network contraction (BASELINE config 5).  Needs a
        tensor-network mode backend (``tn_mode=True`` or a planner); ``bits[q]`` is the value of qubit q."""
        if self._tn is None:
            raise ValueError("amplitude() needs tensor-network mode: compile with tn_mode=True (or use_jdopttn=...)")
        self.check_parameters_torch_device(params)
        if params:
            self._require_cuda(self._device)
        if batched:
            flat = torch.cat([p.reshape(p.shape[0], -1).to(self._rdtype) for p in params], dim=1)
        else:
            flat = self._flatten(params)
        with torch.no_grad():
            amp = self._tn.amplitude(flat.contiguous(), bits, slice_range)
        return amp if batched else amp[0]

    def amplitudes(self, bits_batch, *params, batched=False, slice_range=None):
        """Batch of amplitudes <b_a| U(params) |0...0>, ``bits_batch`` an [A, n] array / tensor of 0/1 (host or
        device) -> complex [A] (``batched=True``: [A, B]).  One pass over the gate operands, one contraction per
        bitstring, and under ``contract_parallel`` ONE all-reduce of the whole batch's partial sums."""
        if self._tn is None:
            raise ValueError("amplitudes() needs tensor-network mode: compile with tn_mode=True (or use_jdopttn=...)")
        self.check_parameters_torch_device(params)
        if params:
            self._require_cuda(self._device)
        if batched:
            flat = torch.cat([p.reshape(p.shape[0], -1).to(self._rdtype) for p in params], dim=1)
        else:
            flat = self._flatten(params)
        with torch.no_grad():
            return self._tn.amplitudes(flat.contiguous(), bits_batch, slice_range)

    def execute_host(self, params: np.ndarray, grad_out: Optional[np.ndarray] = None):
        """HOST numpy [B, P] -> (out [B, n_meas, ...], grad [B, P] or None); copies are inside the call."""
        if getattr(self, "_recipe", None) is not None:
            raise NotImplementedError("execute_host serves expval / probs / state measurements; call the backend "
                                      "with CUDA tensors for var / sample")
        out, grad = self.plan().execute_host(params, grad_out)
        B = out.shape[0]
        nm = len(self._ir.meas)
        if self._res_complex:
            out = out.reshape((B, nm) + self._res_shape + (2,))
            out = out[..., 0] + 1j * out[..., 1]
        else:
            out = out.reshape((B, nm) + self._res_shape)
        return out, grad

    # ------------------------------------------------------ reference properties
    @property
    def operators(self):
        return list(self._operators)

    @property
    def gates_names(self):
        return [op.name for op in self._operators]

    def _bound_parameters(self):
        """Per gate, the parameter list after the last positional re-binding (compiled_circuit.py:522-547)."""
        res = []
        flat = None if self._last_flat is None else self._last_flat.reshape(-1)
        for op, g in zip(self._operators, self._ir.gates):
            pars = list(op.parameters)
            if flat is not None:
                for pos, idx in enumerate(g.param_idx):
                    if idx >= 0:
                        pars[pos] = flat[idx:idx + 1]
            res.append(pars)
        return res

    @property
    def gate_parameters(self):
        return self._bound_parameters()

    @property
    def parameters(self):
        return [p for pars in self._bound_parameters() for p in pars]

    @property
    def trainable_parameters(self):
        out = []
        for op, pars in zip(self._operators, self._bound_parameters()):
            out.extend(pars[i] for i in op.trainable_params)
        return out

    @property
    def qubits(self):
        return [op.qubits for op in self._operators]

    interface = property(lambda self: self._interface)
    diff_method = property(lambda self: self._diff_method)
    backend = property(lambda self: self._backend)
    measurements = property(lambda self: self._user_measurements)
    calculation_mode = property(lambda self: self._calculation_mode)
    device = property(lambda self: self._device)

    @property
    def states_after_measurement(self):
        """Post-measurement states of the last call for ``probs(qubits, after_state=True)``
        (pytorch_backend.py:474-493, :555-560): {"01..": rescaled slice of the final state with the measured qubits
        fixed to that outcome}.  The reference stores them as a side effect of every call; here the final state is
        produced on demand by an auxiliary plan with a ``state()`` measurement, from the last call's parameters,
        so that calls which never read the dictionary pay nothing."""
        target = None
        for ms in self._ir.meas:      # the reference keeps the dictionary of the LAST such measurement
            if ms.kind == MEAS_PROBS and ms.after_state and ms.qubits:
                target = ms
        flat = getattr(self, "_last_flat", None)
        if target is None or flat is None:
            raise ValueError("no states after measurement found! please measure probs with qubits.")
        state_plans = self.__dict__.setdefault("_state_plans", {})
        didx = capi.device_index(flat.device)
        if didx not in state_plans:
            n = self._ir.num_qubits
            ir_state = dataclasses.replace(self._ir, meas=[MeasRec(MEAS_STATE, 0, (), shape=(2,) * n, is_complex=True)])
            dt = capi.TQ_C64 if self._cdtype == torch.complex64 else capi.TQ_C128
            with capi.on_device(didx):
                state_plans[didx] = capi.Plan(ir_state, dt, self._plan_opts)
        plan = state_plans[didx]
        flat = flat[:1].detach().contiguous()
        out = torch.empty((1, plan.out_reals), dtype=self._rdtype, device=flat.device)
        ws_bytes = plan.workspace_bytes(1, False)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=flat.device)
        with torch.cuda.device(flat.device):
            plan.forward(flat.data_ptr(), 1, out.data_ptr(), ws.data_ptr(), ws_bytes, False,
                         torch.cuda.current_stream(flat.device).cuda_stream)
        n = self._ir.num_qubits
        state = torch.view_as_complex(out.reshape((2,) * n + (2,)))
        qs = list(target.qubits)
        res = {}
        for i in range(2 ** len(qs)):
            outcome = [int(c) for c in bin(i)[2:].zfill(len(qs))]            # dec_to_bin, tools/helpers.py:22-40
            loc = dict(zip(qs, outcome))
            part = state[tuple(loc.get(ix, slice(None)) for ix in range(n))]
            scale = torch.sqrt(torch.sum(torch.abs(part) ** 2))             # rescale_state, tools/helpers.py:43-51
            res["".join(str(b) for b in outcome)] = [s / scale for s in part]
        return res


QUDIO_BACKEND_NAME = "pytorch_QUDIO_b200"


class B200QUDIOBackend(B200Backend):
    """Drop-in for ``QUDIOBackend`` (tedq/backends/qudio_backend.py:46-120): ``set_dataset(d)`` once, then
    ``__call__(*params)`` evaluates the circuit on every data row with the shared ``params``; row i binds
    ``(d[i], *params)`` positionally, as ``TorchModel.forward`` does (:149-150).  The reference feeds one row per
    visible GPU through ``torch.nn.DataParallel`` and concatenates the per-row results along dim 0 (:86-102);
    here all rows of this rank go through the device in ONE batched launch sequence and the result has the same
    layout: ``[n_rows * n_meas, ...]``.  Under ``torch.distributed`` the rows are sharded over ranks
    (dist.sharded_batched): every rank gets the full result, gradients flow through this rank's rows only and
    shared-parameter gradients are combined by the caller with one all-reduce (dist.allreduce_sum_)."""

    def __init__(self, backend, circuit, **kwargs):
        super().__init__(backend, circuit, **kwargs)
        self._dataset = None
        self._data_dev = None

    def set_dataset(self, dataset):
        if not isinstance(dataset, torch.Tensor):
            dataset = torch.as_tensor(np.asarray(dataset))
        if dataset.dim() < 2:
            dataset = dataset.reshape(-1, 1)
        self._dataset = dataset
        self._data_dev = None

    def dataset(self):
        return self._dataset

    def __call__(self, *params):
        if self._dataset is None:
            raise ValueError("set_dataset() must be called before the circuit is evaluated")
        self.check_parameters_torch_device(params)
        dev = self._device if params else torch.device("cuda", torch.cuda.current_device())
        self._require_cuda(dev)
        if self._data_dev is None or self._data_dev.device != dev:
            src = self._dataset
            if not src.is_cuda and torch.cuda.is_available():
                src = src.pin_memory()
            self._data_dev = src.to(dev, non_blocking=True)
        from . import dist as tqd

        in_dims = (0,) + (None,) * len(params)
        if tqd.world()[1] > 1:
            local = tqd.sharded_batched(self, self._data_dev, *params, in_dims=in_dims, gather=False)
            out = tqd.gather_rows(local, self._data_dev.shape[0], dev, keep_local_graph=True)
        else:
            out = self.batched(self._data_dev, *params, in_dims=in_dims)
        return out.reshape((out.shape[0] * out.shape[1],) + tuple(out.shape[2:]))
