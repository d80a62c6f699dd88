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


class B200Execute(torch.autograd.Function):
    """Adjoint-method autograd node.  ``apply(run_kwargs, flat)`` with ``flat`` = [B, P] real.

    Calling convention follows TorchExecute (pytorch_backend.py:1191-1223): first argument is a dict
    carrying the backend, ``backward`` returns ``(None, grads)``.
    """

    @staticmethod
    def forward(run_kwargs, flat):
        backend = run_kwargs["backend"]
        out, ws = backend._forward_device(flat, run_kwargs["need_grad"])
        run_kwargs["_ws"] = ws
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        run_kwargs, flat = inputs
        ctx.backend = run_kwargs["backend"]
        ctx.ws = run_kwargs.pop("_ws", None)
        ctx.save_for_backward(flat)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        (flat,) = ctx.saved_tensors
        if ctx.ws is None:
            raise RuntimeError("backward through a circuit evaluated with requires_grad=False")
        grad = ctx.backend._backward_device(flat, dy, ctx.ws)
        ctx.ws = None
        return None, grad

    @staticmethod
    def vmap(info, in_dims, run_kwargs, flat):
        # the op is natively batched over dim 0: fold the vmapped dim into it
        d = in_dims[1]
        if d is None:
            return B200Execute.apply(run_kwargs, flat), None
        flat = flat.movedim(d, 0)
        lead = flat.shape[:2]
        out = B200Execute.apply(dict(run_kwargs), flat.reshape(lead[0] * lead[1], flat.shape[2]))
        return out.reshape(lead + out.shape[1:]), 0


class B200ParamShift(torch.autograd.Function):
    """``diff_method="param_shift"``: same rule as jacobian_param_shift (pytorch_backend.py:180-212) but all
    shifted circuits go through the device in ONE batched launch sequence."""

    @staticmethod
    def forward(run_kwargs, flat):
        out, _ = run_kwargs["backend"]._forward_device(flat, False)
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
        self._plan: Optional[capi.Plan] = None
        self._device = None
        self._last_flat = None
        self._tn = None
        if self._tn_engine:
            from .tn_backend import TNExecutor  # contraction-plan path (tensor-network mode)

            self._tn = TNExecutor(self, self._hyper_opt)
            if self._hyper_opt.get("slicing_opts"):
                self._calculation_mode = self._hyper_opt["slicing_opts"].get("contract_parallel", False)

    # ------------------------------------------------------------------ plan
    def plan(self) -> capi.Plan:
        if self._plan is None:
            dt = capi.TQ_C64 if self._cdtype == torch.complex64 else capi.TQ_C128
            self._plan = capi.Plan(self._ir, dt, self._plan_opts)
        return self._plan

    # ------------------------------------------------------------ device side
    def _forward_device(self, flat: torch.Tensor, need_grad: bool):
        plan = self.plan()
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
        plan = self.plan()
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

    def _param_shift_vjp(self, flat, dy):
        """vjp = dy @ J with J from shifted evaluations (pytorch_backend.py:180-212, :1211-1223)."""
        B, P = flat.shape
        rows, coefs, cols = [], [], []
        for j, name in enumerate(self._param_gate_names()):
            for c, a, s in _shift_recipe(name):
                rows.append((j, a, s))
                coefs.append(c)
                cols.append(j)
        T = len(rows)
        shifted = flat.unsqueeze(1).repeat(1, T, 1)  # [B, T, P]
        for t, (j, a, s) in enumerate(rows):
            shifted[:, t, :] *= a
            shifted[:, t, j] += s
        ev, _ = self._forward_device(shifted.reshape(B * T, P), False)  # [B*T, n_meas]
        if ev.dim() != 2:
            raise ValueError("parameter shift needs scalar measurements (1-D output), as in the reference")
        ev = ev.reshape(B, T, -1)
        coef = torch.tensor(coefs, dtype=ev.dtype, device=ev.device)
        idx = torch.tensor(cols, dtype=torch.long, device=ev.device)
        contrib = torch.einsum("btm,bm,t->bt", ev, dy.to(ev.dtype).reshape(B, -1), coef)
        grad = torch.zeros((B, P), dtype=ev.dtype, device=ev.device)
        grad.index_add_(1, idx, contrib)
        return grad

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
        if self._tn is not None:
            return self._tn.run(flat)
        need_grad = torch.is_grad_enabled() and bool(self._requires_grad) and (
            flat.requires_grad or _is_functorch_wrapped(flat))
        return fn.apply({"backend": self, "need_grad": need_grad}, flat)

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

    def execute_host(self, params: np.ndarray, grad_out: Optional[np.ndarray] = None):
        """HOST numpy [B, P] -> (out [B, n_meas, ...], grad [B, P] or None); copies are inside the call."""
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
    measurements = property(lambda self: self._measurements)
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
        if getattr(self, "_state_plan", None) is None:
            n = self._ir.num_qubits
            ir_state = dataclasses.replace(self._ir, meas=[MeasRec(MEAS_STATE, 0, (), shape=(2,) * n, is_complex=True)])
            dt = capi.TQ_C64 if self._cdtype == torch.complex64 else capi.TQ_C128
            self._state_plan = capi.Plan(ir_state, dt, self._plan_opts)
        plan = self._state_plan
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
