"""Second drop-in boundary: the planner / contraction-tree plug-in the reference already has.

``PyTorchBackend`` hands every network to a third-party tree object and calls
``tree.contract(arrays, backend='torch')`` (pytorch_backend.py:276, :339); the tree classes are injected through
``use_jdopttn=`` / ``use_cotengra=`` (compiled_circuit.py:356-393).  The two classes here have those constructors
and that method, and run the contraction on the B200 engine, so the REFERENCE'S OWN backend can use it unchanged:

    circuit.compilecircuit(backend="pytorch", use_jdopttn=tedq_b200.B200OptTN, requires_grad=False,
                           hyper_opt={"max_repeats": 64, "slicing_opts": {"target_num_slices": 8}})
    circuit.compilecircuit(backend="pytorch", use_cotengra=tedq_b200.ctg_compat, requires_grad=False)

``contract`` is differentiable with respect to every operand that requires grad (the reference's back_prop runs
torch's tape through tree.contract): the backward is the engine's reverse pass over the same tree
(tq_tn_plan_enable_backward / tq_tn_backward), unsliced plans only — a sliced tree refuses operands that require
grad instead of silently detaching them.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import capi, planner


class B200OptTN:
    """``JDOptTN(input_indices, size_dict, output=..., imbalance=..., max_repeats=..., search_parallel=...,
    slicing_opts=...)`` -> object with ``contract(arrays, backend='torch')`` (compiled_circuit.py:388-393)."""

    def __init__(self, input_indices: Sequence[Sequence], size_dict: Optional[Dict] = None, output: Sequence = (),
                 imbalance: float = 0.2, max_repeats: int = 128, search_parallel: bool = True,
                 slicing_opts: Optional[dict] = None, **_ignored):
        ids: Dict[object, int] = {}
        for t in input_indices:
            for ix in t:
                ids.setdefault(ix, len(ids))
        for ix in output:
            ids.setdefault(ix, len(ids))
        if size_dict:
            bad = [k for k, v in size_dict.items() if int(v) != 2]
            if bad:
                raise ValueError(f"B200OptTN contracts qubit networks: every index has size 2 (got {bad[:3]})")
        self.inputs: List[List[int]] = [[ids[ix] for ix in t] for t in input_indices]
        self.output: List[int] = [ids[ix] for ix in output]
        info = planner.find_path(self.inputs, self.output, repeats=min(int(max_repeats), 64), seed=0)
        so = slicing_opts or {}
        tsize = so.get("target_size")
        tnum = int(so.get("target_num_slices", 1) or 1)
        if tsize or tnum > 1:
            info = planner.slice_path(self.inputs, self.output, info,
                                      target_size_log2=int(np.log2(tsize)) if tsize else None, target_num_slices=tnum)
        self.info = info
        self._plans: Dict[int, capi.TnPlan] = {}

    def _plan(self, dtype, device=None) -> capi.TnPlan:
        """Plans are cached per (dtype, CUDA device): their tables are device memory."""
        dt = capi.TQ_C64 if dtype == torch.complex64 else capi.TQ_C128
        key = (dt, capi.device_index(device))
        if key not in self._plans:
            with capi.on_device(key[1]):
                self._plans[key] = capi.TnPlan(self.inputs, self.output, self.info.path, self.info.sliced,
                                               [False] * len(self.inputs), dt)
        return self._plans[key]

    def _plan_bwd(self, dtype, needs, device=None) -> capi.TnPlan:
        dt = capi.TQ_C64 if dtype == torch.complex64 else capi.TQ_C128
        key = (dt, capi.device_index(device), tuple(bool(b) for b in needs))
        if key not in self._plans:
            with capi.on_device(key[1]):
                plan = capi.TnPlan(self.inputs, self.output, self.info.path, self.info.sliced,
                                   [False] * len(self.inputs), dt)
                plan.enable_backward(key[2])
            self._plans[key] = plan
        return self._plans[key]

    def contract(self, arrays, backend: str = "torch", prefer_einsum: bool = True, **_ignored) -> torch.Tensor:
        if backend != "torch":
            raise ValueError("B200OptTN.contract: only backend='torch' (device tensors) is supported")
        if len(arrays) != len(self.inputs):
            raise ValueError(f"expected {len(self.inputs)} arrays, got {len(arrays)}")
        if any(getattr(a, "requires_grad", False) for a in arrays) and torch.is_grad_enabled():
            if self.info.sliced:
                raise NotImplementedError(
                    "B200OptTN.contract: a sliced tree has no reverse pass; drop slicing_opts, compile the reference "
                    "backend with requires_grad=False, or use backend='pytorch_b200'")
            return _TreeContract.apply(self, *arrays)
        return self._contract_values(arrays)

    def _contract_values(self, arrays) -> torch.Tensor:
        dtype = arrays[0].dtype
        if dtype not in (torch.complex64, torch.complex128):
            raise ValueError(f"complex64 / complex128 operands expected, got {dtype}")
        dev = arrays[0].device
        if dev.type != "cuda":
            raise RuntimeError("B200OptTN.contract needs CUDA tensors: the engine has no CPU fallback")
        keep = [a.detach().to(dtype).contiguous() for a in arrays]
        for a, ix in zip(keep, self.inputs):
            if a.numel() != 1 << len(ix):
                raise ValueError(f"operand with {a.numel()} entries for {len(ix)} indices of size 2")
        plan = self._plan(dtype, dev)
        out = torch.zeros((1, 1 << len(self.output)), dtype=dtype, device=dev)
        ws_bytes = plan.workspace_bytes(1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ptrs = np.array([a.data_ptr() for a in keep], dtype=np.int64)
        with torch.cuda.device(dev):
            plan.contract(ptrs, np.zeros(len(keep), dtype=np.int64), 1, 0, plan.n_slices, out.data_ptr(), ws.data_ptr(),
                          ws_bytes, torch.cuda.current_stream(dev).cuda_stream)
        return out.reshape((2,) * len(self.output)) if self.output else out.reshape(())


class _TreeContract(torch.autograd.Function):
    """tree.contract with a backward: the reverse pass of the engine over the same contraction tree."""

    @staticmethod
    def forward(ctx, tree, *arrays):
        dtype = arrays[0].dtype
        dev = arrays[0].device
        if dev.type != "cuda":
            raise RuntimeError("B200OptTN.contract needs CUDA tensors: the engine has no CPU fallback")
        needs = [bool(a.requires_grad) for a in arrays]
        plan = tree._plan_bwd(dtype, needs, dev)
        keep = [a.detach().to(dtype).contiguous() for a in arrays]
        out = torch.zeros((1, 1 << len(tree.output)), dtype=dtype, device=dev)
        ws_bytes = plan.workspace_bytes(1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ptrs = np.array([a.data_ptr() for a in keep], dtype=np.int64)
        strides = np.zeros(len(keep), dtype=np.int64)
        with torch.cuda.device(dev):
            plan.contract(ptrs, strides, 1, 0, 1, out.data_ptr(), ws.data_ptr(), ws_bytes,
                          torch.cuda.current_stream(dev).cuda_stream)
        ctx.tree, ctx.plan, ctx.keep, ctx.ws, ctx.needs = tree, plan, keep, ws, needs
        ctx.ptrs, ctx.strides, ctx.shapes = ptrs, strides, [tuple(a.shape) for a in arrays]
        return out.reshape((2,) * len(tree.output)) if tree.output else out.reshape(())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        tree, plan, ws = ctx.tree, ctx.plan, ctx.ws
        dev = ws.device
        dtype = ctx.keep[0].dtype
        gout = dout.to(dtype).contiguous().reshape(1, -1)
        with torch.cuda.device(dev):
            plan.backward(ctx.ptrs, ctx.strides, 1, gout.data_ptr(), ws.data_ptr(), ws.numel(),
                          torch.cuda.current_stream(dev).cuda_stream)
        shared_off, perset_off, _ = plan.workspace_layout()
        pad = (-ws.data_ptr()) % 256
        esz = ctx.keep[0].element_size()
        flat = ws[pad:pad + (ws.numel() - pad) // esz * esz].view(dtype)
        grads = []
        for t, need in enumerate(ctx.needs):
            if not need:
                grads.append(None)
                continue
            rank = len(tree.inputs[t])
            off, space, bits = plan.grad_info(t, rank)
            base = (shared_off if space == -1 else perset_off) // esz + off
            g = torch.as_strided(flat, (2,) * rank, tuple(1 << b for b in bits), base)
            grads.append(g.conj().resolve_conj().reshape(ctx.shapes[t]).clone())   # the arena holds conj(dL/dT)
        return (None, *grads)


class _HyperOptimizer:
    """``ctg.HyperOptimizer(methods, max_repeats, progbar, minimize, score_compression, slicing_opts)
    .search(inputs, output, size_dict)`` -> tree (compiled_circuit.py:359-368)."""

    def __init__(self, methods=None, max_repeats: int = 128, progbar: bool = False, minimize: str = "flops",
                 score_compression: float = 0.5, slicing_opts: Optional[dict] = None, **_ignored):
        self.max_repeats = max_repeats
        self.slicing_opts = slicing_opts

    def search(self, inputs, output, size_dict) -> B200OptTN:
        return B200OptTN(inputs, size_dict, output=output, max_repeats=self.max_repeats, slicing_opts=self.slicing_opts)


class _CtgCompat:
    """Module-shaped object for ``use_cotengra=``: the reference reads ``.HyperOptimizer`` from it."""
    HyperOptimizer = _HyperOptimizer


ctg_compat = _CtgCompat()
