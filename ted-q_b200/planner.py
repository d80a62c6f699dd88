"""Contraction-path search and slicing for all-extent-2 tensor networks.

Stands in for the third-party planners the reference plugs in at compiled_circuit.py:340-393
(cotengra ``HyperOptimizer.search`` / ``JDOptTN``; neither is vendored, pinned or installable
offline — SURVEY.md 8c).  It exports what every one of them exports: an *ssa path* (pairs of tensor
ids; inputs are 0..n-1, the k-th contraction creates id n+k) plus the set of sliced indices.  The
*choice* of path is ours and is not comparable with the reference; what is pinned is "same path in
=> same lowered plan out" (lowering.py vs the C++ lowering, bit-exact).

Algorithm: randomised greedy (cost = |out| - alpha (|a| + |b|), Gumbel noise, several repeats,
keep the best by flops or width), then greedy index slicing until the largest intermediate fits
``target_size`` and at least ``target_num_slices`` slices exist (PathOptimizer.rst:38-63 semantics:
fix an index, contract each slice, sum).
"""
from __future__ import annotations

import heapq
import math
import random
from typing import Dict, List, Optional, Sequence, Tuple


class PathInfo:
    def __init__(self, path, sliced, width, flops_log2, n_steps):
        self.path = path            # list of (i, j) ssa ids
        self.sliced = sliced        # list of index ids
        self.width = width          # log2 of the largest intermediate of ONE slice
        self.flops_log2 = flops_log2  # log2(8 * sum_steps 2^|a u b|) of ONE slice (complex MAC = 8 flops)
        self.n_steps = n_steps

    @property
    def n_slices(self):
        return 1 << len(self.sliced)

    def __repr__(self):
        return (f"PathInfo(steps={self.n_steps}, width={self.width}, log2flops/slice={self.flops_log2:.2f}, "
                f"slices={self.n_slices})")


def _greedy_once(inputs: Sequence[Sequence[int]], output: Sequence[int], rng: random.Random, alpha: float,
                 temperature: float):
    n = len(inputs)
    sets: Dict[int, frozenset] = {i: frozenset(t) for i, t in enumerate(inputs)}
    where: Dict[int, set] = {}
    for i, s in sets.items():
        for ix in s:
            where.setdefault(ix, set()).add(i)
    keep = set(output)
    next_id = n
    path: List[Tuple[int, int]] = []
    heap: list = []

    def result_of(a, b):
        sa, sb = sets[a], sets[b]
        out = []
        for ix in sa | sb:
            if ix in keep or not where[ix] <= {a, b}:
                out.append(ix)
        return frozenset(out)

    def push(a, b):
        sa, sb = sets[a], sets[b]
        so = result_of(a, b)
        cost = float(1 << len(so)) - alpha * (float(1 << len(sa)) + float(1 << len(sb)))
        score = math.copysign(math.log2(abs(cost) + 1.0), cost)
        if temperature > 0:
            u = rng.random()
            score -= temperature * (-math.log(-math.log(u + 1e-300) + 1e-300))
        heapq.heappush(heap, (score, a, b))

    for ix, ts in where.items():
        ts = sorted(ts)
        for x in range(len(ts)):
            for y in range(x + 1, len(ts)):
                push(ts[x], ts[y])

    while heap:
        _, a, b = heapq.heappop(heap)
        if a not in sets or b not in sets:
            continue
        so = result_of(a, b)
        sa, sb = sets.pop(a), sets.pop(b)
        for ix in sa:
            where[ix].discard(a)
        for ix in sb:
            where[ix].discard(b)
        c = next_id
        next_id += 1
        path.append((a, b))
        sets[c] = so
        nbrs = set()
        for ix in so:
            nbrs |= where[ix]
            where[ix].add(c)
        for t in sorted(nbrs):
            push(t, c)
    # disconnected leftovers: outer products, smallest first
    rest = sorted(sets, key=lambda t: (len(sets[t]), t))
    while len(rest) > 1:
        a, b = rest[0], rest[1]
        so = frozenset((sets[a] | sets[b]))
        c = next_id
        next_id += 1
        path.append((a, b))
        del sets[a], sets[b]
        sets[c] = so
        rest = sorted(sets, key=lambda t: (len(sets[t]), t))
    return path


def path_cost(inputs, output, path, sliced=()):
    """(width, log2 flops of one slice, list of per-step union masks, list of result masks)."""
    sl = set(sliced)
    ids = {}
    for t in inputs:
        for ix in t:
            ids.setdefault(ix, len(ids))
    for ix in output:
        ids.setdefault(ix, len(ids))
    count: Dict[int, int] = {}
    for t in inputs:
        for ix in t:
            count[ix] = count.get(ix, 0) + 1
    for ix in output:
        count[ix] = count.get(ix, 0) + 1
    sets = [frozenset(ix for ix in t if ix not in sl) for t in inputs]
    width = max([len(s) for s in sets] + [0])
    total = 0.0
    unions, results = [], []
    for a, b in path:
        sa, sb = sets[a], sets[b]
        union = sa | sb
        res = []
        for ix in union:
            c = count[ix] - (ix in sa) - (ix in sb)
            if c > 0:
                res.append(ix)
                count[ix] = c + 1
            else:
                count[ix] = 0
        res = frozenset(res)
        sets.append(res)
        unions.append(union)
        results.append(res)
        width = max(width, len(res))
        total += float(1 << len(union))
    return width, (math.log2(total * 8.0) if total > 0 else 0.0), unions, results


def sequential_path(n_inputs: int):
    """Contract the inputs left to right: ((t0 t1) t2) t3 ...  For a circuit network whose operands are listed
    in time order (caps, gates, observable, adjoint gates, caps: tensor_network.py:850-1099) this is state-vector
    evolution written as a contraction path: width = number of qubits (+1), 2^(n+k) MACs per k-qubit gate."""
    path = []
    cur = 0
    for t in range(1, n_inputs):
        path.append((cur, t))
        cur = n_inputs + t - 1
    return path


def find_path(inputs, output, repeats: int = 16, seed: int = 0, minimize: str = "flops",
              alphas=(1.0, 0.5, 0.0), temperatures=(0.0, 0.3, 1.0)) -> PathInfo:
    """Best of ``repeats`` randomised greedy runs (first run is the deterministic alpha=1, T=0 greedy) and the
    time-ordered sequential path (deep circuits on few qubits: greedy merges wide, the sweep stays at width n)."""
    rng = random.Random(seed)
    best = None
    if len(inputs) > 1:
        path = sequential_path(len(inputs))
        width, fl, _, _ = path_cost(inputs, output, path)
        best = ((fl, width) if minimize == "flops" else (width, fl), path, width, fl)
    for r in range(max(1, repeats)):
        alpha = alphas[0] if r == 0 else rng.choice(alphas)
        temp = temperatures[0] if r == 0 else rng.choice(temperatures[1:])
        path = _greedy_once(inputs, output, rng, alpha, temp)
        width, fl, _, _ = path_cost(inputs, output, path)
        key = (fl, width) if minimize == "flops" else (width, fl)
        if best is None or key < best[0]:
            best = (key, path, width, fl)
    _, path, width, fl = best
    return PathInfo(path, [], width, fl, len(path))


def slice_path(inputs, output, info: PathInfo, target_size_log2: Optional[int] = None,
               target_num_slices: int = 1, max_sliced: int = 24) -> PathInfo:
    """Greedy slicing: repeatedly fix the index that leaves the cheapest total work, until the largest
    intermediate of a slice has at most 2^target_size_log2 elements and there are >= target_num_slices slices."""
    sliced: List[int] = list(info.sliced)
    out = set(output)
    need_slices = max(1, int(target_num_slices))
    while len(sliced) < max_sliced:
        width, fl, unions, results = path_cost(inputs, output, info.path, sliced)
        ok_size = target_size_log2 is None or width <= target_size_log2
        ok_num = (1 << len(sliced)) >= need_slices
        if ok_size and ok_num:
            break
        # candidates: indices of the widest intermediates (never an open output index)
        cand: Dict[int, int] = {}
        for res in results:
            if len(res) >= width - 1:
                for ix in res:
                    if ix not in out:
                        cand[ix] = cand.get(ix, 0) + 1
        if not cand:
            for u in unions:
                for ix in u:
                    if ix not in out:
                        cand[ix] = cand.get(ix, 0) + 1
        if not cand:
            break
        best = None
        for ix in sorted(cand, key=lambda i: (-cand[i], i))[:48]:
            w2, f2, _, _ = path_cost(inputs, output, info.path, sliced + [ix])
            key = (w2, f2) if not ok_size else (f2, w2)
            if best is None or key < best[0]:
                best = (key, ix)
        sliced.append(best[1])
    width, fl, _, _ = path_cost(inputs, output, info.path, sliced)
    return PathInfo(info.path, sliced, width, fl, len(info.path))
