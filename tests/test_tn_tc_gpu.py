"""GPU: the tcgen05 split-TF32 contraction kernel against a complex128 einsum of the same step, and against the
fp32 FMA kernels it replaces (north_star: complex64 amplitudes within relative 1e-5).

Every case is a two-tensor network handed to the C ABI (tq_tn_plan_create / tq_tn_contract) with the tensor-core
threshold forced to zero, so that shapes the planner would normally leave on the FMA path run on tensor cores too:
ragged K (padding), either operand providing the accumulator rows, permuted index orders (the pack kernel's bit
permutation), kept-shared indices and batched parameter sets."""
import numpy as np
import pytest
import torch

from tedq_b200 import capi

pytestmark = pytest.mark.gpu


def _contract(inputs, output, arrays, batched, tensor_core, B=1, min_log2=0, chunk=None, c128=False, splitk=True,
              gather=False):
    plan = capi.TnPlan(inputs, output, [(0, 1)], [], batched, capi.TQ_C128 if c128 else capi.TQ_C64)
    plan.set_option(capi.TN_OPT_TENSOR_CORE, 1 if tensor_core else 0)
    plan.set_option(capi.TN_OPT_TC_MIN_LOG2, min_log2)
    if chunk is not None:
        plan.set_option(capi.TN_OPT_TC_CHUNK, chunk)
    if not splitk:
        plan.set_option(capi.TN_OPT_TC_SPLITK, 0)
    plan.set_option(capi.TN_OPT_TC_GATHER, 1 if gather else 0)
    kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
    dev = "cuda"
    cd = torch.complex128 if c128 else torch.complex64
    ts = [torch.tensor(a, dtype=cd, device=dev).contiguous() for a in arrays]
    ptrs = [t.data_ptr() for t in ts]
    strides = [int(np.prod(a.shape[1:])) if b else 0 for a, b in zip(arrays, batched)]
    out = torch.zeros((B if any(batched) else 1, 1 << len(output)), dtype=cd, device=dev)
    ws_bytes = plan.workspace_bytes(B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    plan.contract(ptrs, strides, B, 0, 1, out.data_ptr(), ws.data_ptr(), ws_bytes,
                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy(), kinds


def _rand(rng, shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def _case(rng, n_m, n_n, n_k, n_b, shuffle=True):
    ids = list(range(n_m + n_n + n_k + n_b))
    M, N, K, Bt = ids[:n_m], ids[n_m:n_m + n_n], ids[n_m + n_n:n_m + n_n + n_k], ids[n_m + n_n + n_k:]
    a_idx, b_idx, o_idx = M + K + Bt, K + N + Bt, M + N + Bt
    if shuffle:
        for l in (a_idx, b_idx, o_idx):
            rng.shuffle(l)
    return a_idx, b_idx, o_idx


def _einsum(a_idx, b_idx, o_idx, A, B, batch_a=False, batch_b=False):
    sub = lambda idx, batched: ([51] if batched else []) + list(idx)
    out_b = batch_a or batch_b
    return np.einsum(A.astype(np.complex128), sub(a_idx, batch_a), B.astype(np.complex128), sub(b_idx, batch_b),
                     sub(o_idx, out_b))


SHAPES = [  # (n_m, n_n, n_k, n_b)
    (7, 4, 0, 0), (7, 4, 3, 0), (7, 7, 4, 0), (8, 7, 5, 0), (4, 9, 6, 0), (10, 10, 8, 0), (7, 5, 1, 0),
    (9, 8, 7, 0), (7, 6, 4, 2), (5, 8, 5, 1), (11, 7, 6, 0), (7, 11, 9, 0), (8, 8, 12, 0),
    (7, 7, 12, 0), (8, 7, 11, 0), (7, 4, 11, 1),     # few tiles, long K: split-K partial sums
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "m%d_n%d_k%d_b%d" % s)
@pytest.mark.parametrize("shuffle", [False, True], ids=["canonical", "permuted"])
@pytest.mark.parametrize("gather", [True, False], ids=["gatherA", "images"])
def test_tc_step_matches_einsum(shape, shuffle, gather):
    """gatherA: the row operand is gathered by the GEMM kernel itself where it is streamed (<= 2 column tiles);
    images: both operands go through packed HBM images."""
    n_m, n_n, n_k, n_b = shape
    rng = np.random.RandomState(hash(shape) % 10000 + int(shuffle))
    a_idx, b_idx, o_idx = _case(rng, n_m, n_n, n_k, n_b, shuffle)
    A = _rand(rng, (2,) * len(a_idx))
    B = _rand(rng, (2,) * len(b_idx))
    ref = _einsum(a_idx, b_idx, o_idx, A, B).reshape(-1)
    got, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], True, gather=gather)
    assert kinds == [2], kinds                      # the tensor-core kernel really ran
    scale = np.abs(ref).max()
    assert np.abs(got.reshape(-1) - ref).max() <= 1e-5 * scale, np.abs(got.reshape(-1) - ref).max() / scale
    if gather:
        return
    fma, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], False)
    assert 2 not in kinds
    assert np.abs(fma.reshape(-1) - ref).max() <= 1e-5 * scale


def test_tc_error_is_fp32_class():
    """Split-TF32 with round-to-nearest drains keeps fp32-class accuracy at K = 4096 (a plain TF32 GEMM sits near
    1e-3; accumulating the whole K inside the tensor core shows its round-toward-zero bias, ~6e-5)."""
    rng = np.random.RandomState(3)
    a_idx, b_idx, o_idx = _case(rng, 8, 8, 12, 0)
    A, B = _rand(rng, (2,) * 20), _rand(rng, (2,) * 20)
    ref = _einsum(a_idx, b_idx, o_idx, A, B).reshape(-1)
    got, _ = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], True)
    rel = np.abs(got.reshape(-1) - ref).max() / np.abs(ref).max()
    assert rel < 2e-6, rel
    got, _ = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], True, splitk=False)
    assert np.abs(got.reshape(-1) - ref).max() / np.abs(ref).max() < 2e-6
    biased, _ = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], True, chunk=1 << 20, splitk=False)
    rel_biased = np.abs(biased.reshape(-1) - ref).max() / np.abs(ref).max()
    assert rel_biased > 4 * rel, (rel, rel_biased)


@pytest.mark.parametrize("which", ["a", "b", "both"])
@pytest.mark.parametrize("shape", [(8, 5, 6, 1), (7, 6, 11, 0), (4, 7, 2, 0), (4, 7, 1, 0), (7, 7, 4, 0), (6, 7, 4, 0),
                                   (8, 6, 6, 0), (8, 8, 6, 0), (7, 5, 6, 0), (8, 7, 5, 0)],
                         ids=lambda s: "m%d_n%d_k%d_b%d" % s)
@pytest.mark.parametrize("gather", [False, True], ids=["images", "gatherA"])
def test_tc_step_batched_parameter_sets(which, shape, gather):
    rng = np.random.RandomState(11)
    n_sets = 3
    a_idx, b_idx, o_idx = _case(rng, *shape)
    ba, bb = which in ("a", "both"), which in ("b", "both")
    A = _rand(rng, ((n_sets,) if ba else ()) + (2,) * len(a_idx))
    B = _rand(rng, ((n_sets,) if bb else ()) + (2,) * len(b_idx))
    ref = _einsum(a_idx, b_idx, o_idx, A, B, ba, bb).reshape(n_sets, -1)
    got, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [ba, bb], True, B=n_sets, gather=gather)
    assert kinds == [2]
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


def test_default_threshold_keeps_small_steps_off_tensor_cores():
    rng = np.random.RandomState(5)
    a_idx, b_idx, o_idx = _case(rng, 7, 4, 3, 0)
    A, B = _rand(rng, (2,) * len(a_idx)), _rand(rng, (2,) * len(b_idx))
    _, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], True, min_log2=20)
    assert kinds == [4]     # 2^14 MACs: a fused small step, not a tensor-core launch


@pytest.mark.parametrize("shape", [(0, 0, 15, 0), (2, 1, 14, 0), (0, 3, 12, 1), (0, 0, 21, 0)],
                         ids=lambda s: "m%d_n%d_k%d_b%d" % s)
@pytest.mark.parametrize("c128", [False, True], ids=["c64", "c128"])
def test_split_k_reduction(shape, c128):
    """The closing steps of an amplitude network: a few outputs, a long contracted extent."""
    n_m, n_n, n_k, n_b = shape
    rng = np.random.RandomState(n_k)
    a_idx, b_idx, o_idx = _case(rng, n_m, n_n, n_k, n_b)
    A, B = _rand(rng, (2,) * len(a_idx)), _rand(rng, (2,) * len(b_idx))
    if c128:
        A, B = A.astype(np.complex128) + 1e-9, B.astype(np.complex128) - 1e-9j
    ref = _einsum(a_idx, b_idx, o_idx, A, B).reshape(-1)
    got, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], True, min_log2=20, c128=c128)
    assert kinds == [3], kinds
    tol = 1e-11 if c128 else 1e-5
    assert np.abs(got.reshape(-1) - ref).max() <= tol * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(6, 6, 4, 0), (7, 8, 5, 0), (8, 6, 6, 1), (6, 9, 7, 0), (10, 10, 8, 0)],
                         ids=lambda s: "m%d_n%d_k%d_b%d" % s)
@pytest.mark.parametrize("shuffle", [False, True], ids=["canonical", "permuted"])
def test_complex128_gemm_steps_on_fp64_tensor_cores(shape, shuffle):
    """complex128 GEMM-shaped steps run on k_tn_gemm_dmma (mma.sync m8n8k4 f64, 4M real products): 1e-11."""
    n_m, n_n, n_k, n_b = shape
    rng = np.random.RandomState(sum(shape) * 7 + int(shuffle))
    a_idx, b_idx, o_idx = _case(rng, n_m, n_n, n_k, n_b, shuffle)
    A = (rng.standard_normal((2,) * len(a_idx)) + 1j * rng.standard_normal((2,) * len(a_idx)))
    B = (rng.standard_normal((2,) * len(b_idx)) + 1j * rng.standard_normal((2,) * len(b_idx)))
    ref = _einsum(a_idx, b_idx, o_idx, A, B).reshape(-1)
    got, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], True, c128=True)
    assert kinds == [1], kinds
    assert np.abs(got.reshape(-1) - ref).max() <= 1e-11 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(2, 12, 2, 0), (13, 1, 1, 0), (3, 10, 4, 1), (1, 14, 3, 3), (4, 11, 4, 0),
                                   (13, 4, 2, 0), (2, 9, 1, 4), (0, 12, 3, 0), (3, 12, 0, 0)],
                         ids=lambda s: "m%d_n%d_k%d_b%d" % s)
@pytest.mark.parametrize("c128", [False, True], ids=["c64", "c128"])
def test_apply_kernel_small_times_large(shape, c128):
    """Gate-sized operand times a large operand (k_tn_apply): either side small, kept-shared indices, both dtypes."""
    n_m, n_n, n_k, n_b = shape
    rng = np.random.RandomState(sum(shape) * 13 + n_m)
    a_idx, b_idx, o_idx = _case(rng, n_m, n_n, n_k, n_b, True)
    A = (rng.standard_normal((2,) * len(a_idx)) + 1j * rng.standard_normal((2,) * len(a_idx)))
    B = (rng.standard_normal((2,) * len(b_idx)) + 1j * rng.standard_normal((2,) * len(b_idx)))
    ref = _einsum(a_idx, b_idx, o_idx, A, B).reshape(-1)
    got, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [False, False], False, min_log2=20, c128=c128)
    assert kinds == [5], kinds
    tol = 1e-11 if c128 else 1e-5
    assert np.abs(got.reshape(-1) - ref).max() <= tol * np.abs(ref).max()


def test_apply_kernel_batched_sets():
    rng = np.random.RandomState(77)
    a_idx, b_idx, o_idx = _case(rng, 2, 13, 2, 1, True)
    n_sets = 4
    A = (rng.standard_normal((n_sets,) + (2,) * len(a_idx)) + 1j * rng.standard_normal((n_sets,) + (2,) * len(a_idx)))
    B = (rng.standard_normal((n_sets,) + (2,) * len(b_idx)) + 1j * rng.standard_normal((n_sets,) + (2,) * len(b_idx)))
    ref = _einsum(a_idx, b_idx, o_idx, A, B, True, True).reshape(n_sets, -1)
    got, kinds = _contract([a_idx, b_idx], o_idx, [A, B], [True, True], False, B=n_sets, min_log2=20)
    assert kinds == [5]
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()


def _contract_net(inputs, output, path, arrays, batched, B=1, opts=()):
    plan = capi.TnPlan(inputs, output, path, [], batched, capi.TQ_C64)
    plan.set_option(capi.TN_OPT_TC_MIN_LOG2, 0)
    plan.set_option(capi.TN_OPT_FUSE_SMALL, 0)
    for k, v in opts:
        plan.set_option(k, v)
    cd = torch.complex64
    ts = [torch.tensor(a, dtype=cd, device="cuda").contiguous() for a in arrays]
    ptrs = [t.data_ptr() for t in ts]
    strides = [int(np.prod(a.shape[1:])) if b else 0 for a, b in zip(arrays, batched)]
    out = torch.zeros((B if any(batched) else 1, 1 << len(output)), dtype=cd, device="cuda")
    ws_bytes = plan.workspace_bytes(B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    plan.contract(ptrs, strides, B, 0, 1, out.data_ptr(), ws.data_ptr(), ws_bytes,
                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    kinds = [plan.step_kernel(s) for s in range(plan.n_steps)]
    fused = [plan.step_fuse_to(s) for s in range(plan.n_steps)]
    return out.cpu().numpy(), kinds, fused


# (free of T0, K0, free of T1, K1 taken from T0's free, K1 taken from T1's free, free of T2, kept-shared, batch B)
CHAINS = [
    (9, 4, 5, 3, 0, 5, 0, 1),     # consumer contracts indices that were rows of the producer
    (9, 4, 6, 0, 3, 5, 0, 1),     # ... that were columns of the producer
    (10, 5, 6, 2, 3, 6, 0, 1),    # both
    (4, 5, 9, 0, 4, 7, 0, 1),     # the producer's rows come from its rhs (swap)
    (8, 3, 4, 4, 0, 9, 0, 1),     # the fused tensor becomes the COLUMN operand (B image) of the consumer
    (9, 4, 5, 3, 0, 5, 1, 1),     # a kept-shared index runs through both steps
    (9, 6, 5, 2, 2, 4, 0, 3),     # batched parameter sets
    (7, 11, 4, 3, 0, 5, 0, 1),    # producer with few tiles and a long K: split-K, falls back to plain + pack
]


@pytest.mark.parametrize("shape", CHAINS, ids=lambda s: "m%d_k%d_n%d_kr%d_kc%d_f%d_b%d_B%d" % s)
@pytest.mark.parametrize("shuffle", [False, True], ids=["canonical", "permuted"])
def test_fused_pack_chain_matches_einsum_and_unfused(shape, shuffle):
    """Fused pack (TQ_TN_OPT_TC_FUSE_PACK): the first tensor-core step writes the second step's operand image from
    its epilogue.  Same numbers as the pack-kernel path (to fp32 rounding), both within 1e-5 of a complex128 einsum."""
    f0, k0, f1, kr, kc, f2, nb, B = shape
    rng = np.random.RandomState(sum(shape) * 7 + int(shuffle))
    ids = iter(range(64))
    M0 = [next(ids) for _ in range(f0)]
    K0 = [next(ids) for _ in range(k0)]
    N0 = [next(ids) for _ in range(f1)]
    F2 = [next(ids) for _ in range(f2)]
    Bt = [next(ids) for _ in range(nb)]
    K1 = M0[:kr] + N0[:kc]
    t0, t1, t2 = M0 + K0 + Bt, K0 + N0 + Bt, K1 + F2 + Bt
    out = [i for i in M0 + N0 if i not in K1] + F2 + Bt
    if shuffle:
        for l in (t0, t1, t2, out):
            rng.shuffle(l)
    batched = [B > 1, False, False]
    arrays = [_rand(rng, ((B,) if b else ()) + (2,) * len(t)) / 2 for t, b in zip((t0, t1, t2), batched)]
    sub = lambda idx, b: ([51] if b else []) + list(idx)
    ref = np.einsum(arrays[0].astype(np.complex128), sub(t0, batched[0]), arrays[1].astype(np.complex128), sub(t1, False),
                    arrays[2].astype(np.complex128), sub(t2, False), sub(out, B > 1)).reshape(B, -1)
    path = [(0, 1), (3, 2)]
    on, kinds, fused = _contract_net([t0, t1, t2], out, path, arrays, batched, B)
    off, kinds_off, fused_off = _contract_net([t0, t1, t2], out, path, arrays, batched, B,
                                              opts=[(capi.TN_OPT_TC_FUSE_PACK, 0)])
    assert kinds == [2, 2] and kinds_off == [2, 2], kinds
    assert fused == [1, -1] and fused_off == [-1, -1], (fused, fused_off)
    assert np.abs(off - ref).max() <= 1e-5 * np.abs(ref).max()
    assert np.abs(on - ref).max() <= 1e-5 * np.abs(ref).max()
    # same hi / lo split of the same fp32 values; only the order in which k is accumulated may differ (the consumer
    # moves the three k bits its producer can store contiguously to the front): fp32 rounding noise
    assert np.abs(on - off).max() <= 2e-6 * np.abs(ref).max(), np.abs(on - off).max()
