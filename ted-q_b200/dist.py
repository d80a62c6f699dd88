"""Multi-GPU helpers: one process per GPU (torchrun), ``torch.distributed`` for the plumbing.

The hot path shards in two natural ways:
  * batches of parameter sets / data rows — independent evaluations, no collective on the data path;
    outputs are all-gathered only when the caller wants the full batch on every rank
    (reference analogue: QUDIOBackend's DataParallel scatter/cat, qudio_backend.py:86-102);
  * sliced contraction indices — every rank contracts a contiguous range of slices and the partial sums are
    combined with ONE all-reduce (reference analogue: jdtensorpath RPC slice workers, examples/qubit_rpc.py:110-126);
  * measurements in tensor-network mode — one network per measurement (tensor_network.py:940-1097): rank r
    contracts networks r, r + W, r + 2W, ... and the rows are combined with ONE all-reduce of the (zero-filled)
    result tensor.
A single state vector is never split across GPUs.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n units for ``rank``: ceil(n / world) per rank, last ranks may be short/empty."""
    per = (n + world_size - 1) // world_size
    return min(n, rank * per), min(n, (rank + 1) * per)


def sharded_batched(backend, *params, in_dims=None, gather=True):
    """Evaluate rows [lo, hi) of a batch on this rank; optionally all-gather the results (same order as input)."""
    rank, ws = world()
    if in_dims is None:
        in_dims = (0,) * len(params)
    B = next(p.shape[0] for p, d in zip(params, in_dims) if d is not None)
    lo, hi = shard_range(B, rank, ws)
    local = [p if d is None else p.narrow(0, lo, hi - lo) for p, d in zip(params, in_dims)]
    out = backend.batched(*local, in_dims=in_dims) if hi > lo else None
    if not gather or ws == 1:
        return out
    return gather_rows(out, B, params[0].device)


def gather_rows(out, B: int, dev, keep_local_graph: bool = False):
    """All-gather the per-rank row blocks of ``shard_range`` into [B, ...] on every rank (values only).  With
    ``keep_local_graph`` this rank's rows stay attached to its autograd graph."""
    rank, ws = world()
    lo, hi = shard_range(B, rank, ws)
    per = (B + ws - 1) // ws
    shape = None if out is None else tuple(out.shape[1:])
    meta = [None] * ws
    dist.all_gather_object(meta, (shape, None if out is None else out.dtype))
    shape, dtype = next(m for m in meta if m[0] is not None)
    buf = torch.zeros((per,) + shape, dtype=dtype, device=dev)
    if out is not None:
        buf[: hi - lo] = out.detach()
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf)
    full = torch.cat(parts, 0)[:B]
    if keep_local_graph and out is not None and out.requires_grad:
        full = torch.cat([full[:lo], out, full[hi:]], 0)
    return full


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (complex tensors go through their real view)."""
    if world()[1] > 1:
        dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t)
    return t


def my_measurements(n_meas: int):
    """Measurement (network) ids this rank contracts: round-robin over ranks."""
    rank, ws = world()
    return list(range(rank, n_meas, ws))


def combine_measurements(rows: dict, n_meas: int, like: torch.Tensor) -> torch.Tensor:
    """rows: {measurement id: tensor [B, ...]} computed on this rank -> [B, n_meas, ...] on every rank.
    Missing rows are zero on this rank; ONE all-reduce(sum) fills them in."""
    out = torch.zeros((like.shape[0], n_meas) + tuple(like.shape[1:]), dtype=like.dtype, device=like.device)
    for i, r in rows.items():
        out[:, i] = r
    return allreduce_sum_(out)
