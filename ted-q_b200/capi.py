"""ctypes binding of include/tedq_b200.h (the C-ABI is the product boundary; torch only lends
device memory and the current stream).  There is no CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build
from .ir import CircuitIR

TQ_C64, TQ_C128 = 0, 1
MAX_MEAS_QUBITS = 64


class GateDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("nq", C.c_int32), ("qubits", C.c_int32 * 4),
        ("param_idx", C.c_int32 * 3), ("_pad", C.c_int32),
        ("param_const", C.c_double * 3), ("matrix_off", C.c_int64),
    ]


class MeasDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("flags", C.c_int32), ("nq", C.c_int32), ("_pad", C.c_int32),
        ("qubits", C.c_int32 * MAX_MEAS_QUBITS), ("matrix_off", C.c_int64),
    ]


class PlanOpts(C.Structure):
    _fields_ = [
        ("max_local_qubits_fwd", C.c_int32), ("max_local_qubits_bwd", C.c_int32),
        ("coalesce_bits", C.c_int32), ("threads", C.c_int32), ("fuse", C.c_int32),
        ("structure", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


TN_MAX_RANK = 32
TN_OPT_TENSOR_CORE, TN_OPT_TC_MIN_LOG2, TN_OPT_TC_CHUNK, TN_OPT_FUSE_SMALL, TN_OPT_TC_SPLITK = 0, 1, 2, 3, 4
TN_OPT_TC_GATHER = 5
TN_OPT_TC_FUSE_PACK = 6
TN_OPT_CHAIN = 7


class TnStep(C.Structure):
    _fields_ = [
        ("lhs", C.c_int32), ("rhs", C.c_int32),
        ("n_k", C.c_int32), ("n_m", C.c_int32), ("n_n", C.c_int32), ("n_b", C.c_int32),
        ("lhs_bits", C.c_int8 * TN_MAX_RANK), ("rhs_bits", C.c_int8 * TN_MAX_RANK),
        ("out_idx", C.c_int32 * TN_MAX_RANK),
    ]


_lib: Optional[C.CDLL] = None


class EngineError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def lib() -> C.CDLL:
    """Load (building in-tree first if needed) libtedq_b200.so; fail loudly otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    # build_library() is a digest check when the library is up to date; edited sources are rebuilt, never run stale.
    # Without nvcc (a deployment box) an existing library whose stamp matches the sources is loaded as is.
    path = _build.build_library()
    L = C.CDLL(path)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    L.tq_abi_version.restype = i32
    L.tq_last_error.restype = C.c_char_p
    L.tq_plan_create.restype = i32
    L.tq_plan_create.argtypes = [C.POINTER(GateDesc), i32, C.POINTER(MeasDesc), i32, C.POINTER(C.c_double), i64, i32,
                                 i32, i32, C.POINTER(C.c_double), C.POINTER(PlanOpts), C.POINTER(vp)]
    L.tq_plan_destroy.argtypes = [vp]
    L.tq_plan_destroy.restype = None
    L.tq_tn_subtree_order.restype = i32
    L.tq_tn_subtree_order.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(i32)]
    L.tq_tn_greedy_path.restype = i32
    L.tq_tn_greedy_path.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, C.POINTER(i32),
                                    C.POINTER(C.c_double), i64, C.c_double, C.c_double, C.POINTER(i32),
                                    C.POINTER(i64)]
    for name in ("tq_plan_num_qubits", "tq_plan_num_params"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = i32
    L.tq_plan_num_blocks.argtypes = [vp]
    L.tq_plan_num_blocks.restype = i32
    L.tq_plan_op_stats.argtypes = [vp, i32, i32]
    L.tq_plan_op_stats.restype = i32
    L.tq_plan_num_sweeps.argtypes = [vp, i32]
    L.tq_plan_num_sweeps.restype = i32
    L.tq_plan_sweep_bits.argtypes = [vp, i32, i32, C.POINTER(i32), i32]
    L.tq_plan_sweep_bits.restype = i32
    L.tq_plan_sweep_num_gates.argtypes = [vp, i32, i32]
    L.tq_plan_sweep_num_gates.restype = i32
    L.tq_plan_out_reals.argtypes = [vp]
    L.tq_plan_out_reals.restype = i64
    L.tq_plan_hbm_bytes.argtypes = [vp, i32]
    L.tq_plan_hbm_bytes.restype = i64
    L.tq_plan_flops.argtypes = [vp, i32]
    L.tq_plan_flops.restype = C.c_double
    L.tq_plan_launches.argtypes = [vp, i32]
    L.tq_plan_launches.restype = i64
    L.tq_workspace_bytes.argtypes = [vp, i64, i32]
    L.tq_workspace_bytes.restype = sz
    L.tq_forward.argtypes = [vp, vp, i64, vp, vp, sz, i32, vp]
    L.tq_forward.restype = i32
    L.tq_backward.argtypes = [vp, vp, i64, vp, vp, vp, sz, vp]
    L.tq_backward.restype = i32
    L.tq_workspace_state.argtypes = [vp, vp, i64]
    L.tq_workspace_state.restype = vp
    L.tq_execute_host.argtypes = [vp, vp, i64, vp, vp, vp]
    L.tq_execute_host.restype = i32
    L.tq_sv_axes_perm.argtypes = [i32, C.POINTER(i32), i32, C.POINTER(i32), C.POINTER(i32)]
    L.tq_sv_axes_perm.restype = i32
    pi32 = C.POINTER(i32)
    L.tq_tn_symbol.argtypes = [i32]
    L.tq_tn_symbol.restype = i32
    L.tq_tn_index_map.argtypes = [i32, pi32, pi32, i32, i32, pi32, pi32, i32, pi32, i32, pi32, pi32, i32, i32, pi32, pi32]
    L.tq_tn_index_map.restype = i32
    L.tq_tn_lower.argtypes = [pi32, pi32, i32, pi32, i32, pi32, i32, pi32, i32, C.POINTER(TnStep), pi32, pi32, pi32,
                              i32, pi32, pi32]
    L.tq_tn_lower.restype = i32
    L.tq_tn_plan_create.argtypes = [pi32, pi32, i32, pi32, i32, pi32, i32, pi32, i32, pi32, i32, C.POINTER(vp)]
    L.tq_tn_plan_create.restype = i32
    L.tq_tn_plan_destroy.argtypes = [vp]
    L.tq_tn_plan_destroy.restype = None
    L.tq_tn_plan_num_steps.argtypes = [vp]
    L.tq_tn_plan_num_steps.restype = i32
    L.tq_tn_plan_num_slices.argtypes = [vp]
    L.tq_tn_plan_num_slices.restype = i64
    L.tq_tn_plan_flops.argtypes = [vp]
    L.tq_tn_plan_flops.restype = C.c_double
    L.tq_tn_plan_width.argtypes = [vp]
    L.tq_tn_plan_width.restype = i32
    L.tq_tn_plan_get_step.argtypes = [vp, i32, C.POINTER(TnStep)]
    L.tq_tn_plan_get_step.restype = i32
    L.tq_tn_workspace_bytes.argtypes = [vp, i64]
    L.tq_tn_workspace_bytes.restype = sz
    L.tq_tn_contract.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i64, i64, i64, vp, vp, sz, vp]
    L.tq_tn_contract.restype = i32
    L.tq_tn_contract_prepare.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i64, i64, vp, sz, vp]
    L.tq_tn_contract_prepare.restype = i32
    L.tq_tn_contract_slices.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i64, i64, i64, vp, vp, sz, vp]
    L.tq_tn_contract_slices.restype = i32
    L.tq_dist_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.tq_dist_create.restype = i32
    L.tq_dist_destroy.argtypes = [vp]
    L.tq_dist_destroy.restype = None
    L.tq_dist_slice_range.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    L.tq_dist_slice_range.restype = i32
    L.tq_dist_allreduce.argtypes = [vp, vp, i64, i32, vp]
    L.tq_dist_allreduce.restype = i32
    L.tq_tn_contract_sharded.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(i64), i64, vp, vp, sz, vp]
    L.tq_tn_contract_sharded.restype = i32
    L.tq_tn_plan_enable_backward.argtypes = [vp, pi32]
    L.tq_tn_plan_enable_backward.restype = i32
    L.tq_tn_backward.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i64, i64, vp, vp, sz, vp]
    L.tq_tn_backward.restype = i32
    L.tq_tn_grad_info.argtypes = [vp, i32, C.POINTER(i64), pi32, pi32]
    L.tq_tn_grad_info.restype = i32
    L.tq_tn_workspace_layout.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.tq_tn_workspace_layout.restype = i32
    L.tq_tn_param_grads.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp, vp]
    L.tq_tn_param_grads.restype = i32
    L.tq_tn_gather.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, vp]
    L.tq_tn_gather.restype = i32
    L.tq_tn_plan_set_option.argtypes = [vp, i32, i32]
    L.tq_tn_plan_set_option.restype = i32
    L.tq_tn_plan_step_kernel.argtypes = [vp, i32]
    L.tq_tn_plan_step_kernel.restype = i32
    L.tq_tn_plan_step_fuse_to.argtypes = [vp, i32]
    L.tq_tn_plan_step_fuse_to.restype = i32
    L.tq_tn_plan_step_fuse_mode.argtypes = [vp, i32]
    L.tq_tn_plan_step_fuse_mode.restype = i32
    L.tq_tn_plan_step_flags.argtypes = [vp, i32]
    L.tq_tn_plan_step_flags.restype = i32
    L.tq_tn_profile.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i64, i64, vp, vp, sz, vp, C.POINTER(C.c_float)]
    L.tq_tn_profile.restype = i32
    L.tq_tn_launch_count.argtypes = []
    L.tq_tn_launch_count.restype = i64
    L.tq_tn_gate_offset.argtypes = [vp, i32]
    L.tq_tn_gate_offset.restype = i64
    L.tq_tn_operands.argtypes = [vp, vp, i64, vp, vp, vp]
    L.tq_tn_operands.restype = i32
    if L.tq_abi_version() != 1:
        raise EngineError("libtedq_b200.so ABI version mismatch: rebuild with `python -m tedq_b200.build --force`")
    _lib = L
    return L


def slice_range(n_slices: int, rank: int, world: int):
    """[begin, end) of the slices rank ``rank`` of ``world`` contracts (tq_dist_slice_range; host only)."""
    b, e = C.c_int64(), C.c_int64()
    check(lib().tq_dist_slice_range(n_slices, rank, world, C.byref(b), C.byref(e)), "tq_dist_slice_range")
    return int(b.value), int(e.value)


class Dist:
    """Owns a tq_dist*: the library's handle on an NCCL communicator the host created (``comm``: ncclComm_t as int)."""

    def __init__(self, comm: int, rank: int, world: int):
        h = C.c_void_p()
        check(lib().tq_dist_create(C.c_void_p(comm), rank, world, C.byref(h)), "tq_dist_create")
        self.handle, self.rank, self.world = h, rank, world

    def allreduce(self, ptr, count, dtype, stream):
        check(lib().tq_dist_allreduce(self.handle, ptr, count, dtype, stream), "tq_dist_allreduce")

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and _lib is not None:
            _lib.tq_dist_destroy(h)
            self.handle = None


def device_index(device=None) -> int:
    """CUDA device index of ``device`` (a torch.device, an int, None = the current device); -1 without CUDA."""
    import torch

    if device is not None and not isinstance(device, int):
        device = torch.device(device).index
    if device is None:
        return torch.cuda.current_device() if torch.cuda.is_available() else -1
    return int(device)


def on_device(index: int):
    """Context manager: CUDA device ``index`` current (no-op for -1 = no CUDA)."""
    import contextlib

    import torch

    return torch.cuda.device(index) if index >= 0 else contextlib.nullcontext()


E_WORKSPACE = -4   # TQ_E_WORKSPACE


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().tq_last_error().decode(errors="replace")
        raise EngineError(f"{what} failed ({rc}): {msg}")


def sv_axes_perm(n_qubits, qubits):
    k = len(qubits)
    q = (C.c_int32 * k)(*qubits)
    gp = (C.c_int32 * k)()
    pm = (C.c_int32 * n_qubits)()
    check(lib().tq_sv_axes_perm(n_qubits, q, k, gp, pm), "tq_sv_axes_perm")
    return list(gp), list(pm)


class Plan:
    """Owns a tq_plan*."""

    def __init__(self, ir: CircuitIR, dtype: int = TQ_C64, opts: Optional[dict] = None):
        L = lib()
        self.ir = ir
        self.dtype = dtype
        pool = []
        off = 0
        gates = (GateDesc * max(1, len(ir.gates)))()
        for i, g in enumerate(ir.gates):
            d = gates[i]
            d.kind = g.kind
            d.nq = len(g.qubits)
            if d.nq > 3:
                raise NotImplementedError(f"{g.name}: gates on more than 3 qubits are not supported")
            for j, q in enumerate(g.qubits):
                d.qubits[j] = q
            for j in range(3):
                d.param_idx[j] = g.param_idx[j] if j < len(g.param_idx) else -1
                d.param_const[j] = g.param_const[j] if j < len(g.param_const) else 0.0
            d.matrix_off = -1
            if g.matrix is not None:
                d.matrix_off = off
                m = np.ascontiguousarray(g.matrix, dtype=np.complex128).reshape(-1)
                pool.append(m)
                off += m.size
        meas = (MeasDesc * len(ir.meas))()
        for i, ms in enumerate(ir.meas):
            d = meas[i]
            d.kind = ms.kind
            d.flags = ms.flags
            d.nq = len(ms.qubits)
            for j, q in enumerate(ms.qubits):
                d.qubits[j] = q
            d.matrix_off = -1
            if ms.matrix is not None:
                d.matrix_off = off
                m = np.ascontiguousarray(ms.matrix, dtype=np.complex128).reshape(-1)
                pool.append(m)
                off += m.size
        poolarr = np.concatenate(pool) if pool else np.zeros(1, dtype=np.complex128)
        poolarr = np.ascontiguousarray(poolarr.view(np.float64))
        init_ptr = None
        if ir.init_state is not None:
            self._init = np.ascontiguousarray(ir.init_state.astype(np.complex128).view(np.float64))
            init_ptr = self._init.ctypes.data_as(C.POINTER(C.c_double))
        o = PlanOpts()
        o.coalesce_bits = -1
        o.fuse = -1
        for k, v in (opts or {}).items():
            setattr(o, k, v)
        handle = C.c_void_p()
        check(L.tq_plan_create(gates, len(ir.gates), meas, len(ir.meas), poolarr.ctypes.data_as(C.POINTER(C.c_double)),
                               off, ir.num_qubits, ir.n_params, dtype, init_ptr, C.byref(o), C.byref(handle)),
              "tq_plan_create")
        self.handle = handle
        self.out_reals = int(L.tq_plan_out_reals(handle))
        self.n_params = ir.n_params

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and _lib is not None:
            _lib.tq_plan_destroy(h)
            self.handle = None

    # -- introspection -----------------------------------------------------
    def num_sweeps(self, backward=False) -> int:
        return int(lib().tq_plan_num_sweeps(self.handle, int(backward)))

    def op_stats(self, backward=False):
        """(ops, diagonal-layer ops, gates inside them, real-path ops) emitted into the sweeps of one direction."""
        return tuple(int(lib().tq_plan_op_stats(self.handle, int(backward), w)) for w in range(4))

    def num_register_groups(self, backward=False) -> int:
        """Register groups emitted into the sweeps (plan_opts["structure"] = 2); 0 = the plan runs the default sweeps."""
        return int(lib().tq_plan_op_stats(self.handle, int(backward), 4))

    def num_blocks(self) -> int:
        return int(lib().tq_plan_num_blocks(self.handle))

    def sweep_bits(self, s, backward=False):
        buf = (C.c_int32 * 32)()
        n = lib().tq_plan_sweep_bits(self.handle, int(backward), s, buf, 32)
        return list(buf[:n])

    def sweep_num_gates(self, s, backward=False) -> int:
        return int(lib().tq_plan_sweep_num_gates(self.handle, int(backward), s))

    def hbm_bytes(self, backward=False) -> int:
        return int(lib().tq_plan_hbm_bytes(self.handle, int(backward)))

    def flops(self, backward=False) -> float:
        return float(lib().tq_plan_flops(self.handle, int(backward)))

    def launches(self, backward=False) -> int:
        return int(lib().tq_plan_launches(self.handle, int(backward)))

    def workspace_bytes(self, batch: int, with_backward: bool) -> int:
        return int(lib().tq_workspace_bytes(self.handle, batch, int(with_backward)))

    # -- execution on raw device pointers ------------------------------------
    def forward(self, params_ptr, batch, out_ptr, ws_ptr, ws_bytes, with_backward, stream):
        check(lib().tq_forward(self.handle, params_ptr, batch, out_ptr, ws_ptr, ws_bytes, int(with_backward), stream),
              "tq_forward")

    def backward(self, params_ptr, batch, grad_out_ptr, grad_params_ptr, ws_ptr, ws_bytes, stream):
        check(lib().tq_backward(self.handle, params_ptr, batch, grad_out_ptr, grad_params_ptr, ws_ptr, ws_bytes,
                                stream), "tq_backward")

    def execute_host(self, params: np.ndarray, grad_out: Optional[np.ndarray] = None):
        """HOST buffers in, HOST buffers out (H2D + kernels + D2H inside)."""
        rdt = np.float32 if self.dtype == TQ_C64 else np.float64
        params = np.ascontiguousarray(params, dtype=rdt)
        batch = params.shape[0] if params.ndim == 2 else 1
        out = np.empty((batch, self.out_reals), dtype=rdt)
        gp = None
        go_ptr = gp_ptr = None
        if grad_out is not None:
            grad_out = np.ascontiguousarray(grad_out, dtype=rdt)
            gp = np.empty((batch, self.n_params), dtype=rdt)
            go_ptr = grad_out.ctypes.data
            gp_ptr = gp.ctypes.data
        check(lib().tq_execute_host(self.handle, params.ctypes.data, batch, out.ctypes.data, go_ptr, gp_ptr),
              "tq_execute_host")
        return out, gp


# ---------------------------------------------------------------------------
# tensor-network side
# ---------------------------------------------------------------------------
def _i32(seq):
    seq = list(seq)
    return (C.c_int32 * max(1, len(seq)))(*seq)


def _flatten_inputs(inputs):
    off = [0]
    flat = []
    for t in inputs:
        flat.extend(int(i) for i in t)
        off.append(len(flat))
    return _i32(off), _i32(flat)


def tn_symbol(i: int) -> str:
    return chr(lib().tq_tn_symbol(i))


def tn_index_map(n_qubits, gate_qubits, kind, obs_qubits=(), kept=None):
    """C mirror of tn_index.index_maps for ONE measurement -> (inputs, output)."""
    kinds = {"expval": 0, "probs": 1, "state": 2}
    gnq = _i32([len(q) for q in gate_qubits])
    gq = _i32([(list(q) + [0, 0, 0, 0])[j] for q in gate_qubits for j in range(4)])
    onq = _i32([len(q) for q in obs_qubits])
    oq = _i32([(list(q) + [0, 0, 0, 0])[j] for q in obs_qubits for j in range(4)])
    kq = _i32(kept or [])
    n_t = 2 * n_qubits + 2 * len(gate_qubits) + len(obs_qubits) + 4
    cap_idx = 8 * n_t
    toff = (C.c_int32 * (n_t + 1))()
    tidx = (C.c_int32 * cap_idx)()
    oidx = (C.c_int32 * max(1, n_qubits))()
    nout = C.c_int32()
    nt = lib().tq_tn_index_map(n_qubits, gnq, gq, len(gate_qubits), kinds[kind], onq, oq, len(obs_qubits), kq,
                               -1 if kept is None else len(kept), toff, tidx, n_t, cap_idx, oidx, C.byref(nout))
    if nt < 0:
        check(nt, "tq_tn_index_map")
    inputs = [[tidx[j] for j in range(toff[t], toff[t + 1])] for t in range(nt)]
    return inputs, [oidx[j] for j in range(nout.value)]


def _step_tuple(st: TnStep):
    na = st.n_k + st.n_m + st.n_b
    nb = st.n_k + st.n_n + st.n_b
    no = st.n_m + st.n_n + st.n_b
    return (st.lhs, st.rhs, st.n_k, st.n_m, st.n_n, st.n_b, tuple(st.lhs_bits[:na]), tuple(st.rhs_bits[:nb]),
            tuple(st.out_idx[:no]))


def tn_lower(inputs, output, path, sliced=()):
    """C lowering (host only) -> (step tuples, slice entries [(tensor, ord, bit)], final_perm)."""
    toff, tidx = _flatten_inputs(inputs)
    n_steps = len(path)
    steps = (TnStep * max(1, n_steps))()
    cap = 64 * max(1, len(sliced)) + 8
    st_, so_, sb_ = (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_int32 * cap)()
    ne = C.c_int32()
    fperm = (C.c_int32 * max(1, len(output)))()
    rc = lib().tq_tn_lower(toff, tidx, len(inputs), _i32(output), len(output), _i32([x for p in path for x in p]),
                           n_steps, _i32(sliced), len(sliced), steps, st_, so_, sb_, cap, C.byref(ne), fperm)
    if rc < 0:
        check(rc, "tq_tn_lower")
    return ([_step_tuple(steps[i]) for i in range(n_steps)], [(st_[i], so_[i], sb_[i]) for i in range(ne.value)],
            [fperm[j] for j in range(len(output))])


class TnPlan:
    """Owns a tq_tn_plan*."""

    def __init__(self, inputs, output, path, sliced, input_batched, dtype=TQ_C64):
        toff, tidx = _flatten_inputs(inputs)
        handle = C.c_void_p()
        check(lib().tq_tn_plan_create(toff, tidx, len(inputs), _i32(output), len(output),
                                      _i32([x for p in path for x in p]), len(path), _i32(sliced), len(sliced),
                                      _i32([1 if b else 0 for b in input_batched]), dtype, C.byref(handle)),
              "tq_tn_plan_create")
        self.handle = handle
        self.n_inputs = len(inputs)
        self.n_out = len(output)
        self.dtype = dtype
        self.n_slices = int(lib().tq_tn_plan_num_slices(handle))
        self.flops = float(lib().tq_tn_plan_flops(handle))
        self.width = int(lib().tq_tn_plan_width(handle))
        self.n_steps = int(lib().tq_tn_plan_num_steps(handle))

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and _lib is not None:
            _lib.tq_tn_plan_destroy(h)
            self.handle = None

    def step(self, s):
        st = TnStep()
        check(lib().tq_tn_plan_get_step(self.handle, s, C.byref(st)), "tq_tn_plan_get_step")
        return _step_tuple(st)

    def workspace_bytes(self, batch):
        return int(lib().tq_tn_workspace_bytes(self.handle, batch))

    def enable_backward(self, input_needs_grad):
        """Append the reverse pass (unsliced plans): see tq_tn_plan_enable_backward."""
        check(lib().tq_tn_plan_enable_backward(self.handle, _i32([1 if b else 0 for b in input_needs_grad])),
              "tq_tn_plan_enable_backward")
        self.has_backward = True

    def backward(self, input_ptrs, input_strides, batch, grad_out_ptr, ws_ptr, ws_bytes, stream, slice_id=0):
        ptrs, strides, _keep = self._ptr_arrays(input_ptrs, input_strides)
        check(lib().tq_tn_backward(self.handle, ptrs, strides, batch, slice_id, grad_out_ptr, ws_ptr, ws_bytes, stream),
              "tq_tn_backward")

    def grad_info(self, t, rank):
        """-> (element offset, space (-1 shared / -2 per set), bit of every index of input t inside its gradient)."""
        off, space = C.c_int64(), C.c_int32()
        bits = (C.c_int32 * max(1, rank))()
        check(lib().tq_tn_grad_info(self.handle, t, C.byref(off), C.byref(space), bits), "tq_tn_grad_info")
        return int(off.value), int(space.value), [int(bits[i]) for i in range(rank)]

    def workspace_layout(self):
        """-> (byte offset of the shared arena, byte offset of the per-set arenas, per-set stride in entries)."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().tq_tn_workspace_layout(self.handle, C.byref(a), C.byref(b), C.byref(c)), "tq_tn_workspace_layout")
        return int(a.value), int(b.value), int(c.value)

    def set_option(self, option, value):
        check(lib().tq_tn_plan_set_option(self.handle, option, value), "tq_tn_plan_set_option")

    def step_kernel(self, s):
        """0: per-element kernel, 1: tiled fp32 FMA GEMM, 2: tcgen05 split-TF32 GEMM."""
        return int(lib().tq_tn_plan_step_kernel(self.handle, s))

    def step_fuse_to(self, s):
        """Step whose operand image step s writes from its epilogue (fused pack), -1: plain result."""
        return int(lib().tq_tn_plan_step_fuse_to(self.handle, s))

    def step_fuse_mode(self, s):
        return int(lib().tq_tn_plan_step_fuse_mode(self.handle, s))

    def step_flags(self, s):
        """bit 0: repeats per slice, bit 1: batched over parameter sets."""
        return int(lib().tq_tn_plan_step_flags(self.handle, s))

    def _ptr_arrays(self, input_ptrs, input_strides):
        """ctypes views of the operand pointer / stride tables; int64 numpy arrays go straight through."""
        n = self.n_inputs
        if isinstance(input_ptrs, np.ndarray):
            ptrs_np = np.ascontiguousarray(input_ptrs, dtype=np.int64)
            strides_np = np.ascontiguousarray(input_strides, dtype=np.int64)
            assert ptrs_np.shape == (n,) and strides_np.shape == (n,)
            return (ptrs_np.ctypes.data_as(C.POINTER(C.c_void_p)), strides_np.ctypes.data_as(C.POINTER(C.c_int64)),
                    (ptrs_np, strides_np))
        return (C.c_void_p * n)(*input_ptrs), (C.c_int64 * n)(*input_strides), None

    def profile(self, input_ptrs, input_strides, batch, slice_id, out_ptr, ws_ptr, ws_bytes, stream):
        """Per-step milliseconds of one slice: array [n_steps, 2] = (whole step, operand packing part)."""
        ptrs, strides, _keep = self._ptr_arrays(input_ptrs, input_strides)
        ms = (C.c_float * (2 * self.n_steps + 2))()
        check(lib().tq_tn_profile(self.handle, ptrs, strides, batch, slice_id, out_ptr, ws_ptr, ws_bytes, stream, ms),
              "tq_tn_profile")
        arr = np.asarray(ms, dtype=np.float32)
        self.last_pinned_pack_ms = float(arr[2 * self.n_steps])
        return arr[:2 * self.n_steps].reshape(self.n_steps, 2)

    def contract_prepare(self, input_ptrs, input_strides, batch, slice_begin, ws_ptr, ws_bytes, stream):
        """Once-per-call part of a contraction (slice-invariant steps, pinned operand images) into the workspace."""
        ptrs, strides, _keep = self._ptr_arrays(input_ptrs, input_strides)
        check(lib().tq_tn_contract_prepare(self.handle, ptrs, strides, batch, slice_begin, ws_ptr, ws_bytes, stream),
              "tq_tn_contract_prepare")

    def contract_slices(self, input_ptrs, input_strides, batch, slice_begin, slice_end, out_ptr, ws_ptr, ws_bytes,
                        stream):
        """Slice loop of a contraction on a workspace ``contract_prepare`` filled with the same inputs."""
        ptrs, strides, _keep = self._ptr_arrays(input_ptrs, input_strides)
        check(lib().tq_tn_contract_slices(self.handle, ptrs, strides, batch, slice_begin, slice_end, out_ptr, ws_ptr,
                                          ws_bytes, stream), "tq_tn_contract_slices")

    def contract_sharded(self, dist_handle, input_ptrs, input_strides, batch, out_ptr, ws_ptr, ws_bytes, stream):
        """This rank's slice range + one ncclAllReduce inside the library (tq_tn_contract_sharded)."""
        ptrs, strides, _keep = self._ptr_arrays(input_ptrs, input_strides)
        check(lib().tq_tn_contract_sharded(self.handle, dist_handle, ptrs, strides, batch, out_ptr, ws_ptr, ws_bytes,
                                           stream), "tq_tn_contract_sharded")

    def contract(self, input_ptrs, input_strides, batch, slice_begin, slice_end, out_ptr, ws_ptr, ws_bytes, stream):
        ptrs, strides, _keep = self._ptr_arrays(input_ptrs, input_strides)
        check(lib().tq_tn_contract(self.handle, ptrs, strides, batch, slice_begin, slice_end, out_ptr, ws_ptr, ws_bytes,
                                   stream), "tq_tn_contract")
