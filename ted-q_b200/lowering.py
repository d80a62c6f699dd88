"""Lower (index maps, ssa path, sliced indices) into a static list of pairwise-contraction steps.

Python mirror of the C++ lowering behind ``tq_tn_lower`` (csrc/tq_tn.cu); the two are compared
bit-exactly in tests/test_tn_lowering.py ("same path in => same plan out").

All extents are 2, so a tensor of rank r is addressed by an r-bit string and a permutation of its
modes is a permutation of address bits.  Tensors are stored C-order: the LAST listed index is bit 0.
Input tensors keep their full layout even when some of their indices are sliced — the sliced bits
become a per-slice base offset and the remaining bits keep their physical positions.  Intermediates
are dense with layout  [kept-shared | lhs-only (M) | rhs-only (N)]  slow -> fast.

One step:   C[b, m, n] = sum_k A[b, m, k] * B[b, k, n]
  K  = indices shared by A and B that appear nowhere else (and are not open outputs)  -> contracted
  Bt = shared indices that must survive (another tensor or the output still carries them)
  M  = indices only in A,  N = indices only in B
``lhs_bits`` lists physical bit positions inside A of  [K..., M..., Bt...]  each group fast -> slow;
``rhs_bits`` the same for B with N instead of M.  K bits are listed in A's order for both.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple


@dataclass
class Step:
    lhs: int
    rhs: int
    n_k: int
    n_m: int
    n_n: int
    n_b: int
    lhs_bits: List[int]
    rhs_bits: List[int]
    out_idx: List[int]          # index id of every output bit, fast -> slow


@dataclass
class Lowered:
    steps: List[Step]
    n_inputs: int
    sliced: List[int]
    in_slice_bits: List[List[Tuple[int, int]]]   # per input: (slice ordinal, physical bit)
    final_perm: List[int]       # output bit j (fast->slow of the requested order) <- bit of the last tensor
    out_rank: int

    def as_tuples(self):
        return [(s.lhs, s.rhs, s.n_k, s.n_m, s.n_n, s.n_b, tuple(s.lhs_bits), tuple(s.rhs_bits), tuple(s.out_idx))
                for s in self.steps]


def lower(inputs: Sequence[Sequence[int]], output: Sequence[int], path: Sequence[Tuple[int, int]],
          sliced: Sequence[int] = ()) -> Lowered:
    sl_ord = {ix: o for o, ix in enumerate(sliced)}
    # live tensors: list of (index id, physical bit), fast -> slow
    tensors: Dict[int, List[Tuple[int, int]]] = {}
    in_slice_bits: List[List[Tuple[int, int]]] = []
    count: Dict[int, int] = {}
    for t, idxs in enumerate(inputs):
        r = len(idxs)
        cur, sb = [], []
        for pos, ix in enumerate(idxs):          # pos 0 = slowest
            bit = r - 1 - pos
            if ix in sl_ord:
                sb.append((sl_ord[ix], bit))
            else:
                cur.append((ix, bit))
                count[ix] = count.get(ix, 0) + 1
        cur.reverse()                            # fast -> slow
        tensors[t] = cur
        in_slice_bits.append(sb)
    for ix in output:
        if ix in sl_ord:
            raise ValueError("an open output index cannot be sliced")
        count[ix] = count.get(ix, 0) + 1
    steps: List[Step] = []
    nxt = len(inputs)
    for a, b in path:
        A, B = tensors.pop(a), tensors.pop(b)
        in_b = {ix for ix, _ in B}
        in_a = {ix for ix, _ in A}
        pos_b = {ix: bit for ix, bit in B}
        K = [(ix, bit) for ix, bit in A if ix in in_b and count[ix] == 2]
        Bt = [(ix, bit) for ix, bit in A if ix in in_b and count[ix] > 2]
        M = [(ix, bit) for ix, bit in A if ix not in in_b]
        N = [(ix, bit) for ix, bit in B if ix not in in_a]
        for ix, _ in M + N:
            if count[ix] < 2:
                raise ValueError(f"index {ix} is dangling (appears once and is not an output)")
        lhs_bits = [bit for _, bit in K] + [bit for _, bit in M] + [bit for _, bit in Bt]
        rhs_bits = [pos_b[ix] for ix, _ in K] + [bit for _, bit in N] + [pos_b[ix] for ix, _ in Bt]
        out = [ix for ix, _ in N] + [ix for ix, _ in M] + [ix for ix, _ in Bt]   # fast -> slow
        for ix, _ in K:
            count[ix] = 0
        for ix, _ in Bt:
            count[ix] -= 1
        steps.append(Step(a, b, len(K), len(M), len(N), len(Bt), lhs_bits, rhs_bits, out))
        tensors[nxt] = [(ix, j) for j, ix in enumerate(out)]
        nxt += 1
    if len(tensors) != 1:
        raise ValueError("path does not contract the network to a single tensor")
    (last,) = tensors.values()
    have = {ix: bit for ix, bit in last}
    if sorted(have) != sorted(output):
        raise ValueError("final tensor indices differ from the requested output")
    # requested order: output listed slow -> fast; bit j (fast -> slow) is output[-1-j]
    final_perm = [have[output[len(output) - 1 - j]] for j in range(len(output))]
    return Lowered(steps, len(inputs), list(sliced), in_slice_bits, final_perm, len(output))


def step_flops(step: Step) -> float:
    """8 * M * N * K * batch real flops (complex multiply-add = 8)."""
    return 8.0 * float(1 << (step.n_k + step.n_m + step.n_n + step.n_b))
