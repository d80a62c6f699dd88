"""Tensor-network contraction mode of the B200 backend (``use_jdopttn=`` / ``use_cotengra=`` / ``tn_mode=True``).

Mirrors the reference's TN branch of ``PyTorchBackend.execute`` (tedq/backends/pytorch_backend.py:242-355):
one network per measurement (index maps = tn_index.py, bit-exact with gen_tensor_networks), operands in the
reference's order (caps, gates, observable(s), adjoint gates reversed, caps), contracted along a static
pairwise plan.  The plan comes from planner.py (the reference's planners — cotengra, jdtensorpath,
opt_einsum — are third-party and absent offline); a planner object exposing
``find_path(inputs, output, size_dict) -> ssa pairs`` can be passed through ``use_jdopttn`` / ``use_cotengra``.

Sliced indices are sharded over ranks when ``torch.distributed`` is initialised and
``hyper_opt['slicing_opts']['contract_parallel']`` is set: every rank contracts its slice range and the
partial results are combined with ONE all-reduce (the reference's analogue is jdtensorpath's RPC master /
worker slice sum, examples/qubit_rpc.py:110-126).

Gradients: values come from the contraction; ``backward`` runs the adjoint state-vector sweeps of the same
engine (mathematically the same derivative, no autograd tape through ~10^3 tiny contractions).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import capi, planner, tn_index, tn_simplify
from .tn_index import OPD_ADJ, OPD_CAP, OPD_GATE, OPD_OBS


class TNExecutor:
    def __init__(self, backend, hyper_opt: Optional[dict] = None, amplitude_bits=None):
        self.backend = backend
        self.ho = hyper_opt or {}
        circuit = backend._circuit
        self.n = circuit.num_qubits
        # tn_simplify=True (the reference's default; its own simplifier does not work, tensor_network.py:94):
        # diagonal / controlled gates enter with shared wire indices and reduced tensors (tn_simplify.py)
        self.simplify = bool(getattr(backend, "_tn_simplify", False))
        # hyper_opt["light_cone"] (default on): expval / marginal networks keep only the gates inside the
        # measurement's causal cone (tn_index.light_cone); the gates outside cancel against their own adjoints.
        # False = the reference's networks, every gate twice (tensor_network.py:1027-1072)
        self.networks = (tn_simplify if self.simplify else tn_index).networks_of_circuit(
            circuit, prune_light_cone=bool(self.ho.get("light_cone", True)))
        self.gate_structs = tn_simplify.gate_structures(circuit) if self.simplify else [None] * len(circuit.operators)
        self.gate_batched = [any(i >= 0 for i in g.param_idx) for g in backend._ir.gates]
        self.infos: List[planner.PathInfo] = []
        self.plans: List[Optional[capi.TnPlan]] = []
        so = self.ho.get("slicing_opts") or {}
        self.contract_parallel = bool(so.get("contract_parallel", False))
        # hyper_opt["measurement_parallel"]: networks (one per measurement) are dealt round-robin to the ranks
        self.measurement_parallel = bool(self.ho.get("measurement_parallel", False))
        ext = backend._use_jdopttn or backend._use_cotengra
        for net in self.networks:
            info = None
            if ext and hasattr(ext, "find_path"):
                path = [tuple(p) for p in ext.find_path(net.inputs, net.output, {})]
                w, fl, _, _ = planner.path_cost(net.inputs, net.output, path)
                info = planner.PathInfo(path, [], w, fl, len(path))
            else:
                info = planner.cached_plan(self.ho.get("plan_cache"), net.inputs, net.output,
                                           lambda net=net: self._search(net), **self._plan_key())
            self.infos.append(info)
        self.plans = {}           # (network, CUDA device index) -> capi.TnPlan: a plan's tables live on one device
        self._const = {}

    def _plan_key(self):
        so = self.ho.get("slicing_opts") or {}
        key = {"max_repeats": int(self.ho.get("max_repeats", 16)), "seed": int(self.ho.get("seed", 0)),
               "minimize": self.ho.get("minimize", "flops"), "reconf_sweeps": int(self.ho.get("reconf_sweeps", 0)),
               "reconf_leaves": int(self.ho.get("reconf_leaves", 8)), "time_model": self.ho.get("time_model"),
               "target_size": so.get("target_size"), "target_num_slices": int(so.get("target_num_slices", 1) or 1)}
        if int(self.ho.get("restarts", 1)) > 1:
            key["restarts"] = int(self.ho["restarts"])
        return key

    def _search(self, net) -> planner.PathInfo:
        """Path search + slicing for one network (hyper_opt: max_repeats, seed, minimize, reconf_sweeps,
        reconf_leaves, time_model, restarts, slicing_opts{target_size, target_num_slices})."""
        return planner.search_plan(net.inputs, net.output, **self._plan_key())

    # ------------------------------------------------------------------
    def _engine_opts(self, plan):
        """hyper_opt["engine_opts"] = {TQ_TN_OPT_*: value}, applied to every plan this executor builds."""
        for opt, val in (self.ho.get("engine_opts") or {}).items():
            plan.set_option(int(opt), int(val))
        return plan

    def _plan(self, i, device=None) -> capi.TnPlan:
        key = (i, capi.device_index(device))
        if key not in self.plans:
            net, info = self.networks[i], self.infos[i]
            batched = [kind in (OPD_GATE, OPD_ADJ) and self.gate_batched[ref] for kind, ref in net.operands]
            dt = capi.TQ_C64 if self.backend._cdtype == torch.complex64 else capi.TQ_C128
            with capi.on_device(key[1]):
                self.plans[key] = self._engine_opts(
                    capi.TnPlan(net.inputs, net.output, info.path, info.sliced, batched, dt))
        return self.plans[key]

    def _constants(self, device):
        key = str(device)
        if key not in self._const:
            cd = self.backend._cdtype
            cap0 = torch.tensor([1.0, 0.0], dtype=cd, device=device)
            cap1 = torch.tensor([0.0, 1.0], dtype=cd, device=device)
            obs = []
            for ms in self.backend._measurements:
                rt = getattr(ms.return_type, "value", ms.return_type)
                if rt == "expval":
                    lst = ms.obs if isinstance(ms.obs, list) else [ms.obs]
                    mats = []
                    for o in lst:
                        m = np.asarray(o.matrix, dtype=np.complex128).reshape(-1)
                        st = tn_simplify.verified_structure(o.name, len(o.qubits), o.matrix) if self.simplify else None
                        mats.append(torch.tensor(m[list(st[2])] if st else m, dtype=cd, device=device))
                    obs.append(mats)
                else:
                    obs.append([])
            self._const[key] = (cap0, cap1, obs)
        return self._const[key]

    def _slice_range(self, n_slices):
        if self.contract_parallel and torch.distributed.is_available() and torch.distributed.is_initialized():
            r, w = torch.distributed.get_rank(), torch.distributed.get_world_size()
            per = (n_slices + w - 1) // w
            return min(n_slices, r * per), min(n_slices, (r + 1) * per), True
        return 0, n_slices, False

    def _reduced(self, plan_sv, device):
        """Once: gather list of the reduced gate / adjoint tensors (simplified networks).  -> dict with the device
        index array for tq_tn_gather, the number of reduced entries per parameter set and, per gate, the offset of
        its reduced tensor (ket half, adjoint half) inside the reduced buffer."""
        cache = self.__dict__.setdefault("_red_cache", {})
        key = str(device)
        if key not in cache:
            L = capi.lib()
            idx, off_g, off_a = [], {}, {}
            for g, st in enumerate(self.gate_structs):
                if st is None:
                    continue
                base = int(L.tq_tn_gate_offset(plan_sv.handle, g))
                off_g[g] = len(idx)
                idx += [base + o for o in st[2]]
            for g, st in enumerate(self.gate_structs):
                if st is None:
                    continue
                base = int(L.tq_tn_gate_offset(plan_sv.handle, g))
                off_a[g] = len(idx)
                idx += [-(base + o) - 1 for o in st[2]]
            cache[key] = {"idx": torch.tensor(idx, dtype=torch.int32, device=device) if idx else None,
                          "n": len(idx), "off_g": off_g, "off_a": off_a}
        return cache[key]

    def _gather_reduced(self, plan_sv, gm, am, total, B, stream):
        """Reduced operand buffer [B, n_red] of this call (None when nothing is reduced)."""
        red = self._reduced(plan_sv, gm.device)
        if not red["n"]:
            return None
        buf = torch.empty((B, red["n"]), dtype=gm.dtype, device=gm.device)
        dt = capi.TQ_C64 if gm.dtype == torch.complex64 else capi.TQ_C128
        capi.check(capi.lib().tq_tn_gather(gm.data_ptr(), am.data_ptr(), total, red["idx"].data_ptr(), red["n"],
                                           buf.data_ptr(), B, dt, stream), "tq_tn_gather")
        return buf

    def _operand_tables(self, i, net, plan_sv, total):
        """Per network, once: for every operand the base buffer (0 cap, 1 gate matrices, 2 adjoint matrices,
        3 reduced tensors, 4 + j observable j), its element offset inside that buffer and its parameter-set
        stride."""
        cache = self.__dict__.setdefault("_opd_tables", {})
        if i not in cache:
            L = capi.lib()
            red = self._reduced(plan_sv, self._table_device) if self.simplify else None
            reductions = net.reductions or [None] * len(net.operands)
            base, off, stride = [], [], []
            for (kind, ref), rd in zip(net.operands, reductions):
                if kind == OPD_CAP:
                    base.append(0), off.append(0), stride.append(0)
                elif kind == OPD_OBS:
                    base.append(4 + ref), off.append(0), stride.append(0)
                elif rd is not None:
                    base.append(3)
                    off.append(red["off_g"][ref] if kind == OPD_GATE else red["off_a"][ref])
                    stride.append(red["n"] if self.gate_batched[ref] else 0)
                else:
                    base.append(1 if kind == OPD_GATE else 2)
                    off.append(int(L.tq_tn_gate_offset(plan_sv.handle, ref)))
                    stride.append(total if self.gate_batched[ref] else 0)
            cache[i] = {"base": np.asarray(base, dtype=np.int64), "off": np.asarray(off, dtype=np.int64),
                        "stride": np.asarray(stride, dtype=np.int64), "any_batched": any(st != 0 for st in stride)}
        return cache[i]

    def tree_backward_available(self) -> bool:
        """Reverse mode through the contraction tree.  hyper_opt["tn_backward"] = "tree" | "adjoint" forces one of the
        two gradient paths; default: the adjoint state-vector sweeps while a state vector fits comfortably (<= 26
        qubits: they are the cheaper gradient there), the tree beyond.  With ``contract_parallel`` the slices of the
        reverse pass are sharded over the ranks like those of the forward pass (every rank runs forward + reverse
        pass of its slice range, ONE all-reduce of the [B, P] gradients; the reference's RPC runner does backward
        across its slice workers the same way, examples/qubit_rpc.py:81-83)."""
        mode = self.ho.get("tn_backward")
        if mode == "adjoint":
            return False
        ok = not self.measurement_parallel and any(self.gate_batched)
        if mode == "tree" and not ok:
            raise ValueError("tn_backward='tree' needs trainable parameters and no measurement_parallel")
        return ok and (mode == "tree" or self.n > 26)

    def _plan_bwd(self, i, device=None) -> capi.TnPlan:
        """Plan of network i with the reverse pass appended (forward-only calls keep the leaner plan)."""
        cache = self.__dict__.setdefault("_plans_bwd", {})
        key = (i, capi.device_index(device))
        if key not in cache:
            net, info = self.networks[i], self.infos[i]
            batched = [kind in (OPD_GATE, OPD_ADJ) and self.gate_batched[ref] for kind, ref in net.operands]
            dt = capi.TQ_C64 if self.backend._cdtype == torch.complex64 else capi.TQ_C128
            with capi.on_device(key[1]):
                plan = self._engine_opts(capi.TnPlan(net.inputs, net.output, info.path, info.sliced, batched, dt))
                plan.enable_backward(batched)
            cache[key] = plan
        return cache[key]

    def _grad_tables(self, i, device, slice_id=0):
        """Per network and slice, once: device int32 [n_gates, 16] tables (ket half, bra half) of the per-set-arena
        element offset of every gate-tensor entry's gradient (-1: no such entry — constant operand, entry removed
        by tn_simplify, or an entry whose sliced index bits differ from this slice) for tq_tn_param_grads."""
        cache = self.__dict__.setdefault("_grad_tabs", {})
        key = (i, str(device), int(slice_id))
        if key not in cache:
            net, plan = self.networks[i], self._plan_bwd(i, device)
            sliced = list(self.infos[i].sliced)
            ng = len(self.backend._ir.gates)
            og = np.full((ng, 16), -1, dtype=np.int32)
            oa = np.full((ng, 16), -1, dtype=np.int32)
            reductions = net.reductions or [None] * len(net.operands)
            for t, ((kind, ref), rd) in enumerate(zip(net.operands, reductions)):
                if kind not in (OPD_GATE, OPD_ADJ) or not self.gate_batched[ref]:
                    continue
                rank = len(net.inputs[t])
                off, space, bits = plan.grad_info(t, rank)
                assert space == -2, "a batched operand's gradient lives in the per-set arena"
                tab = og if kind == OPD_GATE else oa
                ix = net.inputs[t]
                for j in range(1 << rank):          # entry j of the (possibly reduced) operand tensor, C order
                    e, present = off, True
                    for q in range(rank):
                        bit = (j >> (rank - 1 - q)) & 1
                        if bits[q] >= 0:
                            e += bit << bits[q]
                        elif bit != ((slice_id >> sliced.index(ix[q])) & 1):   # sliced index: fixed by the slice
                            present = False
                    if present:
                        tab[ref, rd[j] if rd is not None else j] = e
            cache[key] = (torch.tensor(og, device=device), torch.tensor(oa, device=device))
        return cache[key]

    def contract_values(self, flat: torch.Tensor, keep=None):
        """-> list of complex tensors [B or 1, 2^n_out] (one per measurement).  ``keep`` (a list): use the plans with
        a reverse pass and append (network id, plan, ptrs, strides, workspace, keep-alive) for tree_backward."""
        be = self.backend
        dev = flat.device
        plan_sv = be.plan(dev)
        L = capi.lib()
        B = flat.shape[0]
        cd = be._cdtype
        total = int(L.tq_tn_gate_offset(plan_sv.handle, len(be._ir.gates)))
        gm = torch.empty((B, max(1, total)), dtype=cd, device=dev)
        am = torch.empty((B, max(1, total)), dtype=cd, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            capi.check(L.tq_tn_operands(plan_sv.handle, flat.data_ptr(), B, gm.data_ptr(), am.data_ptr(), stream),
                       "tq_tn_operands")
        cap0, cap1, obs = self._constants(dev)
        esz = gm.element_size()
        with torch.cuda.device(dev):
            red_buf = self._gather_reduced(plan_sv, gm, am, total, B, stream) if self.simplify else None
        results = []
        mine = None
        if self.measurement_parallel and torch.distributed.is_available() and torch.distributed.is_initialized():
            from . import dist as tqd
            mine = set(tqd.my_measurements(len(self.networks)))
        for i, net in enumerate(self.networks):
            if mine is not None and i not in mine:
                results.append(None)
                continue
            # operand pointers = base[kind] + offset * element size: vectorised, tables built once per network
            self._table_device = dev
            tab = self._operand_tables(i, net, plan_sv, total)
            use_bwd = keep is not None and bool(tab["any_batched"])   # a constant network has no gradient
            # a sliced plan re-runs its forward slice by slice inside tree_backward: plain forward here
            plan = self._plan_bwd(i, dev) if use_bwd and not self.infos[i].sliced else self._plan(i, dev)
            bases = np.array([cap0.data_ptr(), gm.data_ptr(), am.data_ptr(),
                              red_buf.data_ptr() if red_buf is not None else 0] + [o.data_ptr() for o in obs[i]],
                             dtype=np.int64)
            ptrs = bases[tab["base"]] + tab["off"] * esz
            strides = tab["stride"]
            any_b = bool(tab["any_batched"])
            out = torch.zeros((B if any_b else 1, 1 << plan.n_out), dtype=cd, device=dev)
            ws_bytes = plan.workspace_bytes(B)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            s0, s1, dist_on = self._slice_range(plan.n_slices)
            with torch.cuda.device(dev):
                if s1 > s0:
                    plan.contract(ptrs, strides, B, s0, s1, out.data_ptr(), ws.data_ptr(), ws_bytes, stream)
            if dist_on:
                if plan.n_slices == 1 and torch.distributed.get_rank() != 0:
                    out.zero_()
                torch.distributed.all_reduce(torch.view_as_real(out))
            if use_bwd:
                keep.append((i, self._plan_bwd(i, dev), ptrs, strides, None if self.infos[i].sliced else ws,
                             (gm, am, red_buf)))
            if not any_b and B > 1:
                out = out.expand(B, -1)
            results.append(out)
        return results

    def tree_backward(self, flat: torch.Tensor, dy: torch.Tensor, kept) -> torch.Tensor:
        """dL/dparams [B, P] by reverse mode through every network's contraction tree (tq_tn_backward) and the
        chain rule of the gate tensors (tq_tn_param_grads).  dy: [B, n_meas, ...] cotangent of the stacked result."""
        be = self.backend
        B = flat.shape[0]
        dev = flat.device
        L = capi.lib()
        stream = torch.cuda.current_stream(dev).cuda_stream
        grad = torch.zeros((B, be._ir.n_params), dtype=be._rdtype, device=dev)
        plan_sv = be.plan(dev)
        sharded = False
        for i, plan, ptrs, strides, ws, _alive in kept:
            gi = dy[:, i].reshape(B, -1)
            gout = (gi if gi.is_complex() else gi.to(be._rdtype) + 0j).to(be._cdtype).contiguous()
            _, perset_off, set_stride = plan.workspace_layout()
            scratch = None
            if ws is None:      # sliced plan: forward + reverse pass slice by slice, gradients add up
                ws_bytes = plan.workspace_bytes(B)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                scratch = torch.zeros((B, 1 << plan.n_out), dtype=be._cdtype, device=dev)
            base = (ws.data_ptr() + 255) // 256 * 256 + perset_off
            s0, s1, dist_on = self._slice_range(plan.n_slices)
            sharded = sharded or dist_on
            if dist_on and plan.n_slices == 1:      # an unsliced network: rank 0's reverse pass is the whole gradient
                s0, s1 = (0, 1) if torch.distributed.get_rank() == 0 else (0, 0)
            with torch.cuda.device(dev):
                for sl in range(s0, s1):
                    if scratch is not None:
                        plan.contract(ptrs, strides, B, sl, sl + 1, scratch.data_ptr(), ws.data_ptr(), ws.numel(), stream)
                    og, oa = self._grad_tables(i, dev, sl)
                    plan.backward(ptrs, strides, B, gout.data_ptr(), ws.data_ptr(), ws.numel(), stream, sl)
                    capi.check(L.tq_tn_param_grads(plan_sv.handle, flat.data_ptr(), B, base, set_stride, og.data_ptr(),
                                                   oa.data_ptr(), grad.data_ptr(), stream), "tq_tn_param_grads")
        if sharded:     # the contraction is a sum over slices: so are its parameter gradients
            torch.distributed.all_reduce(grad)
        return grad

    # ------------------------------------------------------------------ amplitudes (C5)
    def _amplitude_plan(self):
        if getattr(self, "_amp", None) is None:
            circuit = self.backend._circuit
            gq = [list(op.qubits) for op in circuit.operators]
            if self.simplify:
                net0 = tn_simplify.index_maps(self.n, gq, self.gate_structs, [("state", None)])[0]
            else:
                net0 = tn_index.index_maps(self.n, gq, [("state", None)])[0]
            net = amplitude_network(net0, [0] * self.n)
            info = planner.cached_plan(self.ho.get("plan_cache"), net.inputs, net.output,
                                       lambda: self._search(net), **self._plan_key())
            self._amp = [net, info, None]
        return self._amp

    def _amplitude_operands(self, flat: torch.Tensor, bits):
        """-> (plan, input pointers, strides, output tensor, workspace, keep-alive list)."""
        be = self.backend
        amp = self._amplitude_plan()
        net, info = amp[0], amp[1]
        dev = flat.device
        plans = self.__dict__.setdefault("_amp_plans", {})      # CUDA device index -> plan
        didx = capi.device_index(dev)
        if didx not in plans:
            batched = [kind == OPD_GATE and self.gate_batched[ref] for kind, ref in net.operands]
            dt = capi.TQ_C64 if be._cdtype == torch.complex64 else capi.TQ_C128
            grp = self._slice_group(net, info, batched)
            self._amp_group = grp
            with capi.on_device(didx):
                if grp is None:
                    plan = capi.TnPlan(net.inputs, net.output, info.path, info.sliced, batched, dt)
                else:   # the grouped indices leave the network: their values become the plan's batch ("set") dimension
                    gset = set(grp["indices"])
                    inputs2 = [[ix for ix in t if ix not in gset] for t in net.inputs]
                    batched2 = [bool(grp["axes"].get(t)) for t in range(len(net.inputs))]
                    plan = capi.TnPlan(inputs2, net.output, info.path, grp["rest"], batched2, dt)
                plans[didx] = self._engine_opts(plan)
        plan = amp[2] = plans[didx]       # amp[2]: the plan of the device used last (introspection)
        grp = getattr(self, "_amp_group", None)
        plan_sv = be.plan(dev)
        L = capi.lib()
        B = flat.shape[0]
        cd = be._cdtype
        total = int(L.tq_tn_gate_offset(plan_sv.handle, len(be._ir.gates)))
        gm = torch.empty((B, max(1, total)), dtype=cd, device=dev)
        am = torch.empty((B, max(1, total)), dtype=cd, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            capi.check(L.tq_tn_operands(plan_sv.handle, flat.data_ptr(), B, gm.data_ptr(), am.data_ptr(), stream),
                       "tq_tn_operands")
        cap0, cap1, _ = self._constants(dev)
        esz = gm.element_size()
        red = self._reduced(plan_sv, dev) if self.simplify else None
        with torch.cuda.device(dev):
            red_buf = self._gather_reduced(plan_sv, gm, am, total, B, stream) if self.simplify else None
        # operand pointers, vectorised: table (base buffer, element offset, stride, closing-cap qubit) built once
        tab = amp[3] if len(amp) > 3 else None
        if tab is None:
            reductions = net.reductions or [None] * len(net.operands)
            base, off, stride, capq = [], [], [], []
            for (kind, ref), rd in zip(net.operands, reductions):
                if kind == OPD_CAP:
                    base.append(0), off.append(0), stride.append(0)
                    capq.append(ref[0] if isinstance(ref, tuple) else -1)
                elif rd is not None:
                    base.append(2), off.append(red["off_g"][ref]), capq.append(-1)
                    stride.append(red["n"] if self.gate_batched[ref] else 0)
                else:
                    base.append(1), off.append(int(L.tq_tn_gate_offset(plan_sv.handle, ref))), capq.append(-1)
                    stride.append(total if self.gate_batched[ref] else 0)
            tab = {"base": np.asarray(base, dtype=np.int64), "off": np.asarray(off, dtype=np.int64),
                   "stride": np.asarray(stride, dtype=np.int64), "capq": np.asarray(capq, dtype=np.int64)}
            amp.append(tab)
        bases = np.array([cap0.data_ptr(), gm.data_ptr(), red_buf.data_ptr() if red_buf is not None else 0],
                         dtype=np.int64)
        ptrs = bases[tab["base"]] + tab["off"] * esz
        ptrs = self._patch_caps(ptrs, bits, dev)
        strides = tab["stride"]
        any_b = bool((strides != 0).any())
        keep = [gm, am, red_buf]
        self._amp_stacked = {}      # input id -> [G, entries] re-laid-out operand of the last call (slice groups)
        if grp is not None:
            # slice group: every combination of the grouped indices is one "set" of the plan.  The few operands that
            # carry a grouped index are re-laid out as [G sets][tensor without those indices] (tiny gate tensors).
            assert not any_b      # groups are only planned when no operand is batched over parameter sets
            G = 1 << len(grp["indices"])
            ptrs = ptrs.copy()
            strides = strides.copy()
            for t, axes in grp["axes"].items():
                rank = len(net.inputs[t])
                kind = int(tab["base"][t])
                if tab["capq"][t] >= 0 or kind == 0:
                    src = cap1 if ptrs[t] == cap1.data_ptr() else cap0
                    full = src.reshape(-1)[: 1 << rank]
                else:
                    buf = gm if kind == 1 else red_buf
                    off = int(tab["off"][t])
                    full = buf.reshape(-1)[off: off + (1 << rank)]
                full = full.reshape((2,) * rank)
                rows = []
                for sset in range(G):
                    sel = [slice(None)] * rank
                    for ax, gi in axes:        # gi: position of that index in the group (bit of the set id)
                        sel[ax] = (sset >> gi) & 1
                    rows.append(full[tuple(sel)].reshape(-1))
                stacked = torch.stack(rows).contiguous()
                self._amp_stacked[t] = stacked
                keep.append(stacked)
                ptrs[t] = stacked.data_ptr()
                strides[t] = stacked.shape[1]
            B = G
            any_b = True
        out = torch.zeros((B if any_b else 1, 1), dtype=cd, device=dev)
        ws_bytes = plan.workspace_bytes(B)
        ws = getattr(self, "_amp_ws", None)
        if ws is None or ws.numel() < ws_bytes or ws.device != dev:
            ws = self._amp_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        return plan, ptrs, strides, out, ws, ws_bytes, any_b, tuple(keep)

    def _patch_caps(self, ptrs, bits, dev):
        """Operand pointer table with the closing caps <bits| of the amplitude network (cap q points at [1, 0] or
        [0, 1]); every other operand is independent of the bitstring."""
        tab = self._amplitude_plan()[3]
        cap0, cap1, _ = self._constants(dev)
        closing = tab["capq"] >= 0
        if closing.any():
            ptrs = ptrs.copy()
            bit_arr = np.asarray(bits, dtype=np.int64).reshape(-1)[tab["capq"][closing]]
            ptrs[closing] = np.where(bit_arr == 1, cap1.data_ptr(), cap0.data_ptr())
        return ptrs

    def _slice_group(self, net, info, batched):
        """hyper_opt["slice_batch"] = g: 2^g slices go through the device together as the batch dimension of ONE
        launch sequence (default 5 when no gate is batched over parameter sets, fewer when ranks would go idle; 0 = one
        slice at a time).  A slice
        of a 40-qubit amplitude is ~20 launches of 20-90 us with one tile per SM: grouping slices gives every launch
        several tiles per SM (prologue / epilogue overlap inside the persistent kernels) and divides the launch count.
        -> None or {"indices": grouped sliced indices, "rest": the others, "axes": {tensor: [(axis, group bit)]}}."""
        g = int(self.ho.get("slice_batch", 5))
        # memory: the per-set arena is a few times the largest intermediate (2^width entries); keep a group's
        # largest tensors at <= 2^28 entries together (2 GiB complex64) unless the caller asked for a size
        if "slice_batch" not in self.ho:
            g = min(g, max(0, 28 - int(info.width)))
        if self.contract_parallel and torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()     # keep at least one group per rank
            while g > 0 and (info.n_slices >> g) < world:
                g -= 1
        if g <= 0 or any(batched) or not info.sliced or self.ho.get("tn_backward") == "tree":
            return None
        chosen, axes = [], {}
        for ix in reversed(list(info.sliced)):
            if len(chosen) == g:
                break
            holders = [t for t, idxs in enumerate(net.inputs) if ix in idxs]
            # keep every operand at rank >= 1 after the grouped indices are taken out
            if any(len(net.inputs[t]) - len(axes.get(t, [])) - 1 < 1 for t in holders):
                continue
            for t in holders:
                axes.setdefault(t, []).append((net.inputs[t].index(ix), len(chosen)))
            chosen.append(ix)
        if not chosen:
            return None
        rest = [ix for ix in info.sliced if ix not in chosen]
        return {"indices": chosen, "rest": rest, "axes": axes}

    def amplitude(self, flat: torch.Tensor, bits, slice_range=None):
        """<bits| U(params) |0...0> for every parameter set -> complex [B].  Slices are sharded over ranks
        (contract_parallel) and combined with one all-reduce."""
        plan, ptrs, strides, out, ws, ws_bytes, any_b, _keep = self._amplitude_operands(flat, bits)
        B = flat.shape[0]
        grouped = getattr(self, "_amp_group", None) is not None
        Bp = out.shape[0] if grouped else B      # sets of the plan: slice-group members, else parameter sets
        dev = flat.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        if slice_range is None:
            s0, s1, dist_on = self._slice_range(plan.n_slices)
        else:
            (s0, s1), dist_on = slice_range, False
        with torch.cuda.device(dev):
            if s1 > s0:
                plan.contract(ptrs, strides, Bp, s0, s1, out.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        if grouped:
            out = out.sum(0, keepdim=True)       # the grouped indices are summed like every sliced index
            any_b = False
        if dist_on:
            if plan.n_slices == 1 and torch.distributed.get_rank() != 0:
                out.zero_()
            torch.distributed.all_reduce(torch.view_as_real(out))
        return out.reshape(-1).expand(B) if not any_b else out.reshape(-1)

    def _amp_batch(self, n_amps: int, group: int, width: int) -> int:
        """Amplitudes that share one launch sequence (hyper_opt["amplitude_batch"]; default: as many as keep
        amplitudes x slice-group members <= 2^(28 - width) sets, the memory rule of the slice groups)."""
        if "amplitude_batch" in self.ho:
            return max(1, min(n_amps, int(self.ho["amplitude_batch"])))
        cap = 1 << max(0, 28 - int(width))
        return max(1, min(n_amps, cap // max(1, group)))

    def _multi_plan(self, dev, net, info, grp, closing):
        """Plan of the amplitude network whose closing caps differ per set: sets = amplitudes x slice-group members."""
        plans = self.__dict__.setdefault("_amp_plans_multi", {})
        didx = capi.device_index(dev)
        if didx not in plans:
            dt = capi.TQ_C64 if self.backend._cdtype == torch.complex64 else capi.TQ_C128
            if grp is None:
                inputs2, rest = net.inputs, info.sliced
                batched2 = [bool(c) for c in closing]
            else:
                gset = set(grp["indices"])
                inputs2 = [[ix for ix in t if ix not in gset] for t in net.inputs]
                rest = grp["rest"]
                batched2 = [bool(grp["axes"].get(t)) or bool(closing[t]) for t in range(len(net.inputs))]
            with capi.on_device(didx):
                plans[didx] = self._engine_opts(capi.TnPlan(inputs2, net.output, info.path, rest, batched2, dt))
        return plans[didx]

    def _amplitude_items(self, flat: torch.Tensor, bits_batch):
        """Launch sequences of a batch of amplitudes: -> (plan, [(pointers, strides, sets, output rows)], output
        tensor, workspace, workspace bytes, A, multi, grouped, any_b, keep-alive list)."""
        if torch.is_tensor(bits_batch):
            bits_batch = bits_batch.detach().cpu().numpy()
        bits_batch = np.asarray(bits_batch, dtype=np.int64).reshape(-1, self.n)
        A = bits_batch.shape[0]
        plan1, ptrs1, strides1, out0, ws, ws_bytes, any_b, keep1 = self._amplitude_operands(flat, bits_batch[0])
        B = flat.shape[0]
        net, info = self._amplitude_plan()[0], self._amplitude_plan()[1]
        grp = getattr(self, "_amp_group", None)
        grouped = grp is not None
        G = out0.shape[0] if grouped else 1
        dev = flat.device
        cd = self.backend._cdtype
        params_batched = any_b and not grouped       # gates batched over parameter sets: one amplitude at a time
        Ab = 1 if params_batched else self._amp_batch(A, G, info.width)
        multi = Ab > 1
        tab = self._amplitude_plan()[3]
        closing = tab["capq"] >= 0
        plan = self._multi_plan(dev, net, info, grp, closing) if multi else plan1
        chunks = [(a0, min(A, a0 + Ab)) for a0 in range(0, A, Ab)]
        keep = list(keep1)
        if multi:
            sets_max = Ab * G
            ws_bytes = plan.workspace_bytes(sets_max)
            ws = getattr(self, "_amp_ws_multi", None)
            if ws is None or ws.numel() < ws_bytes or ws.device != dev:
                ws = self._amp_ws_multi = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            base_ptrs, base_strides = ptrs1.copy(), strides1.copy()
            for t, stacked in self._amp_stacked.items():       # [G, entries] -> [Ab * G, entries], set = a * G + m
                rep = stacked.repeat(Ab, 1).contiguous()
                keep.append(rep)
                base_ptrs[t], base_strides[t] = rep.data_ptr(), rep.shape[1]
            cap_ids = np.nonzero(closing)[0]
            cap_q = tab["capq"][cap_ids]
            out = torch.zeros((A, G), dtype=cd, device=dev)
            chunk_args = []
            for a0, a1 in chunks:
                nb = a1 - a0
                bits_dev = torch.as_tensor(bits_batch[a0:a1][:, cap_q], device=dev)              # [nb, n_caps]
                caps = torch.nn.functional.one_hot(bits_dev, 2).to(cd)                             # [nb, n_caps, 2]
                caps = caps.permute(1, 0, 2).unsqueeze(2).expand(-1, -1, G, -1).reshape(len(cap_ids), nb * G, 2)
                caps = caps.contiguous()
                keep.append(caps)
                p, st = base_ptrs.copy(), base_strides.copy()
                p[cap_ids] = caps.data_ptr() + np.arange(len(cap_ids), dtype=np.int64) * (nb * G * 2 * caps.element_size())
                st[cap_ids] = 2
                chunk_args.append((p, st, nb * G, out[a0:a1]))
        else:
            out = torch.zeros((A,) + tuple(out0.shape), dtype=out0.dtype, device=dev)
            Bp = out0.shape[0] if grouped else B
            chunk_args = [(self._patch_caps(ptrs1, bits_batch[a], dev), strides1, Bp, out[a]) for a in range(A)]
        return plan, chunk_args, out, ws, ws_bytes, A, multi, grouped, any_b, keep

    def amplitudes(self, flat: torch.Tensor, bits_batch, slice_range=None):
        """<b_a| U(params) |0...0> for a batch of bitstrings b_a ([A, n] array of 0/1) -> complex [A] (or [A, B] when
        the gates are batched over parameter sets).  The gate operands are built once.  Several amplitudes share one
        launch sequence: the closing caps <b_a| become per-set operands and the plan's batch dimension runs over
        (amplitude, slice-group member) — every launch then has the tiles of up to 32 slices, whatever the rank's
        share of one amplitude's slices is.  Every rank contracts its slice range of every amplitude; the partial sums
        of the whole batch are combined with ONE all-reduce (contract_parallel)."""
        plan, chunk_args, out, ws, ws_bytes, A, multi, grouped, any_b, keep = self._amplitude_items(flat, bits_batch)
        dev = flat.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        if slice_range is None:
            s0, s1, dist_on = self._slice_range(plan.n_slices)
        else:
            (s0, s1), dist_on = slice_range, False
        n_items = len(chunk_args)
        overlap = n_items > 1 and s1 > s0 and bool(self.ho.get("overlap_prepare", True))
        with torch.cuda.device(dev):
            if not overlap:
                for p, st, sets, o in chunk_args:
                    if s1 > s0:
                        plan.contract(p, st, sets, s0, s1, o.data_ptr(), ws.data_ptr(), ws_bytes, stream)
            else:
                # Two workspaces, two streams: the once-per-call part of item i + 1 (hundreds of tiny, latency-bound
                # steps and the pinned operand images) runs on a side stream while the slices of item i
                # (throughput-bound GEMMs) run on the caller's stream.
                ws2 = getattr(self, "_amp_ws2", None)
                if ws2 is None or ws2.numel() < ws_bytes or ws2.device != dev:
                    ws2 = self._amp_ws2 = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                side = getattr(self, "_amp_side_stream", None)
                if side is None or side.device != dev:
                    side = self._amp_side_stream = torch.cuda.Stream(device=dev)
                wss = (ws, ws2)
                main = torch.cuda.current_stream(dev)
                ev_prep = [torch.cuda.Event() for _ in range(n_items)]
                ev_done = [torch.cuda.Event() for _ in range(n_items)]
                side.wait_stream(main)            # gate operands / caps / zeroed output are ready
                p, st, sets, _o = chunk_args[0]
                plan.contract_prepare(p, st, sets, s0, wss[0].data_ptr(), ws_bytes, side.cuda_stream)
                ev_prep[0].record(side)
                for i, (p, st, sets, o) in enumerate(chunk_args):
                    main.wait_event(ev_prep[i])
                    plan.contract_slices(p, st, sets, s0, s1, o.data_ptr(), wss[i & 1].data_ptr(), ws_bytes, stream)
                    ev_done[i].record(main)
                    if i + 1 < n_items:
                        if i >= 1:
                            side.wait_event(ev_done[i - 1])     # workspace (i + 1) & 1 is free again
                        pn, stn, setsn, _on = chunk_args[i + 1]
                        plan.contract_prepare(pn, stn, setsn, s0, wss[(i + 1) & 1].data_ptr(), ws_bytes,
                                              side.cuda_stream)
                        ev_prep[i + 1].record(side)
        if multi:
            out = out.sum(1, keepdim=True)       # the grouped indices are summed like every sliced index
            any_b = False
        elif grouped:
            out = out.sum(1, keepdim=True)
            any_b = False
        if dist_on:
            if plan.n_slices == 1 and torch.distributed.get_rank() != 0:
                out.zero_()
            torch.distributed.all_reduce(torch.view_as_real(out))
        out = out.reshape(A, -1)
        return out if any_b else out[:, 0]

    def slice_members(self, plan_slice: int):
        """Slice ids of the path's own slicing (bit j <-> info.sliced[j]) that plan slice ``plan_slice`` covers: one
        without slice groups, 2^g with them."""
        info = self._amplitude_plan()[1]
        grp = getattr(self, "_amp_group", None)
        sliced = list(info.sliced)
        if grp is None:
            return [int(plan_slice)]
        base = 0
        for r, ix in enumerate(grp["rest"]):
            base |= ((plan_slice >> r) & 1) << sliced.index(ix)
        out = []
        for m in range(1 << len(grp["indices"])):
            sid = base
            for gi, ix in enumerate(grp["indices"]):
                sid |= ((m >> gi) & 1) << sliced.index(ix)
            out.append(sid)
        return out

    def amplitude_profile(self, flat: torch.Tensor, bits, slice_id=0):
        """Per-step timing of ONE slice of the amplitude contraction (tq_tn_profile): list of dicts with the step's
        log2 extents, the kernel that ran it, milliseconds (whole / operand packing) and whether it repeats per
        slice."""
        plan, ptrs, strides, out, ws, ws_bytes, _, _keep = self._amplitude_operands(flat, bits)
        dev = flat.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        Bp = out.shape[0] if getattr(self, "_amp_group", None) is not None else flat.shape[0]
        with torch.cuda.device(dev):
            ms = plan.profile(ptrs, strides, Bp, slice_id, out.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        rows = []
        for s in range(plan.n_steps):
            st = plan.step(s)
            rows.append({"step": s, "k": st[2], "m": st[3], "n": st[4], "b": st[5], "kernel": plan.step_kernel(s),
                         "per_slice": bool(plan.step_flags(s) & 1), "per_set": bool(plan.step_flags(s) & 2),
                         "sets": Bp, "ms": float(ms[s, 0]), "pack_ms": float(ms[s, 1])})
        return rows

    def amplitudes_profile(self, flat: torch.Tensor, bits_batch, slice_id=0):
        """Per-step timing of ONE launch sequence of ``amplitudes(bits_batch)`` (its first chunk of amplitudes, plan
        slice ``slice_id``): -> (rows as amplitude_profile, amplitudes in that launch sequence, sets)."""
        plan, chunk_args, out, ws, ws_bytes, A, multi, grouped, _, _keep = self._amplitude_items(flat, bits_batch)
        dev = flat.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        p, st, sets, o = chunk_args[0]
        scratch = torch.zeros_like(o)
        with torch.cuda.device(dev):
            ms = plan.profile(p, st, sets, slice_id, scratch.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        rows = []
        for s in range(plan.n_steps):
            stp = plan.step(s)
            rows.append({"step": s, "k": stp[2], "m": stp[3], "n": stp[4], "b": stp[5], "kernel": plan.step_kernel(s),
                         "per_slice": bool(plan.step_flags(s) & 1), "per_set": bool(plan.step_flags(s) & 2),
                         "sets": sets, "ms": float(ms[s, 0]), "pack_ms": float(ms[s, 1])})
        G = o.shape[1] if (multi and o.dim() > 1) else (sets if grouped else 1)
        return rows, max(1, sets // max(1, G)), sets, float(getattr(plan, "last_pinned_pack_ms", 0.0))

    def run(self, flat: torch.Tensor) -> torch.Tensor:
        """Values (and, through ``backend.B200Execute``, gradients: reverse mode through the same contraction trees
        where ``tree_backward_available()``, otherwise the adjoint state-vector sweeps of the same engine)."""
        from .backend import B200Execute
        return self.backend._run(flat, B200Execute)

    def _forward_values(self, flat, keep=None):
        be = self.backend
        vals = self.contract_values(flat.contiguous(), keep)
        res = []
        B = flat.shape[0]
        for ms, v, net in zip(be._ir.meas, vals, self.networks):
            if v is None:      # another rank's measurement (measurement_parallel)
                res.append(None)
                continue
            # the network's open legs define the result shape (probs() with qubits=None contracts to a scalar
            # in the reference's TN branch: tensor_network.py:1017-1019 leaves the output empty)
            v = v.reshape((B,) + (2,) * len(net.output))
            res.append(v if ms.is_complex else v.real)   # torch.squeeze(result.real), pytorch_backend.py:340,:348
        if not be._shapes_ok:
            raise ValueError("You can not have multiple measurements with different shapes!!")
        if any(r is None for r in res):
            from . import dist as tqd
            like = next((r for r in res if r is not None), None)
            if like is None:   # more ranks than measurements: this rank only joins the all-reduce
                ms0, net0 = be._ir.meas[0], self.networks[0]
                like = torch.zeros((B,) + (2,) * len(net0.output), device=flat.device,
                                   dtype=be._cdtype if ms0.is_complex else be._rdtype)
            return tqd.combine_measurements({i: r for i, r in enumerate(res) if r is not None}, len(res), like)
        return torch.stack(res, 1)


def amplitude_network(net: tn_index.Network, bits):
    """Close the open legs of a ``state()`` network with basis vectors <b|: the reference has
    no amplitude measurement; this is its state network plus one cap per wire."""
    inputs = [list(t) for t in net.inputs]
    ops = list(net.operands)
    red = list(net.reductions) if net.reductions else []
    for q, ix in enumerate(net.output):
        inputs.append([ix])
        ops.append((OPD_CAP, (q, int(bits[q]))))
        if red:
            red.append(None)
    return tn_index.Network(inputs, [], ops, red)
