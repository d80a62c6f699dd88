"""GPU: fused runs of small contraction steps (k_tn_fused) against the one-launch-per-step kernels and against a
complex128 pairwise einsum of the same path (oracle/tn_ref.contract_path), for random small networks: shared and
batched operands, sliced indices, both dtypes."""
import numpy as np
import pytest
import torch

from oracle import tn_ref
from tedq_b200 import capi, planner

pytestmark = pytest.mark.gpu


def _random_network(rng, n_tensors, n_indices, max_rank, n_open):
    """Every index appears in exactly two tensors (or one tensor + the output)."""
    inputs = [[] for _ in range(n_tensors)]
    output = []
    for ix in range(n_indices):
        if len(output) < n_open:
            cand = [t for t in range(n_tensors) if len(inputs[t]) < max_rank]
            inputs[rng.choice(cand)].append(ix)
            output.append(ix)
            continue
        cand = [t for t in range(n_tensors) if len(inputs[t]) < max_rank]
        if len(cand) < 2:
            break
        a, b = rng.choice(cand, size=2, replace=False)
        inputs[a].append(ix)
        inputs[b].append(ix)
    for t in range(n_tensors):          # no empty tensors: hang a private pair of indices between neighbours
        if not inputs[t]:
            ix = max([i for l in inputs for i in l] + output + [-1]) + 1
            inputs[t].append(ix)
            inputs[(t + 1) % n_tensors].append(ix)
    for l in inputs:
        rng.shuffle(l)
    return inputs, output


def _run(inputs, output, path, sliced, arrays, batched, B, c128, fuse):
    plan = capi.TnPlan(inputs, output, path, sliced, batched, capi.TQ_C128 if c128 else capi.TQ_C64)
    plan.set_option(capi.TN_OPT_FUSE_SMALL, 1 if fuse else 0)
    kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
    cd = torch.complex128 if c128 else torch.complex64
    ts = [torch.tensor(a, dtype=cd, device="cuda").contiguous() for a in arrays]
    strides = [int(np.prod(a.shape[1:])) if b else 0 for a, b in zip(arrays, batched)]
    out = torch.zeros((B if any(batched) else 1, 1 << len(output)), dtype=cd, device="cuda")
    ws_bytes = plan.workspace_bytes(B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    plan.contract([t.data_ptr() for t in ts], strides, B, 0, plan.n_slices, out.data_ptr(), ws.data_ptr(), ws_bytes,
                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy(), kinds


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("c128", [False, True], ids=["c64", "c128"])
def test_fused_runs_match_einsum(seed, c128):
    rng = np.random.RandomState(100 + seed)
    n_t = int(rng.randint(4, 40))
    inputs, output = _random_network(rng, n_t, int(rng.randint(n_t, 3 * n_t)), 6, int(rng.randint(0, 4)))
    info = planner.find_path(inputs, output, repeats=2, seed=seed)
    sliced = []
    if seed % 3 == 0:
        info = planner.slice_path(inputs, output, info, target_num_slices=4)
        sliced = info.sliced
    B = 1 if seed % 2 else 5
    batched = [bool(rng.randint(2)) and B > 1 for _ in inputs]
    arrays = []
    for ix, b in zip(inputs, batched):
        shape = ((B,) if b else ()) + (2,) * len(ix)
        arrays.append((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex128) / 1.5)
    ref = []
    for s in range(B if any(batched) else 1):
        ref.append(np.asarray(tn_ref.contract_path([a[s] if b else a for a, b in zip(arrays, batched)], inputs, output,
                                                   info.path)).reshape(-1))
    ref = np.stack(ref)
    fused, kinds_f = _run(inputs, output, info.path, sliced, arrays, batched, B, c128, True)
    plain, kinds_p = _run(inputs, output, info.path, sliced, arrays, batched, B, c128, False)
    assert 4 in kinds_f and 4 not in kinds_p
    tol = (1e-11 if c128 else 1e-5) * np.abs(ref).max()
    assert np.abs(fused - ref).max() <= tol, np.abs(fused - ref).max()
    assert np.abs(plain - ref).max() <= tol, np.abs(plain - ref).max()


def test_fused_small_dot_product():
    """A 2^13-term dot product inside a fused run: K is split over 32 lanes and reduced with shuffles."""
    rng = np.random.RandomState(9)
    idx = list(range(13))
    a_idx, b_idx = list(idx), list(idx)
    rng.shuffle(b_idx)
    A = (rng.standard_normal((2,) * 13) + 1j * rng.standard_normal((2,) * 13))
    Bm = (rng.standard_normal((2,) * 13) + 1j * rng.standard_normal((2,) * 13))
    ref = np.einsum(A, a_idx, Bm, b_idx, [])
    got, kinds = _run([a_idx, b_idx], [], [(0, 1)], [], [A, Bm], [False, False], 1, False, True)
    assert kinds == [4]
    assert abs(got.reshape(-1)[0] - ref) <= 1e-5 * abs(ref) + 1e-4
