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
import os
import random
import time
from typing import Dict, List, Optional, Sequence, Tuple


class PathInfo:
    def __init__(self, path, sliced, width, flops_log2, n_steps):
        self.path = path            # list of (i, j) ssa ids
        self.sliced = sliced        # list of index ids
        self.width = width          # log2 of the largest intermediate of ONE slice
        self.flops_log2 = flops_log2  # log2(8 * sum_steps 2^|a u b|) of ONE slice (complex MAC = 8 flops)
        self.n_steps = n_steps
        self.search_s = None        # seconds the path search took (None: unknown, e.g. a plan given by the caller)
        self.from_cache = False     # loaded from the on-disk plan cache

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


def _greedy_once_native(inputs: Sequence[Sequence[int]], output: Sequence[int], rng: random.Random, alpha: float,
                        temperature: float):
    """``_greedy_once`` through the compiled library (tq_tn_greedy_path, csrc/tq_planner.cu): the same path and the
    same state of ``rng`` afterwards (tests/test_planner_cpu.py).  The pairs are scored in the order the mirror
    scores them, which is decided here (dict / frozenset iteration order); the heap loop — the part that costs
    seconds on a 40-qubit network — runs compiled."""
    import ctypes as C

    import numpy as np

    from . import capi

    n = len(inputs)
    sets = {i: frozenset(t) for i, t in enumerate(inputs)}
    where: Dict[int, set] = {}
    for i, s in sets.items():
        for ix in s:
            where.setdefault(ix, set()).add(i)
    dense = {ix: k for k, ix in enumerate(where)}
    for ix in output:
        dense.setdefault(ix, len(dense))
    init = []
    for ix, ts in where.items():
        ts = sorted(ts)
        for x in range(len(ts)):
            for y in range(x + 1, len(ts)):
                init.append((ts[x], ts[y]))
    off = np.zeros(n + 1, dtype=np.int32)
    flat = []
    for i in range(n):
        flat.extend(dense[ix] for ix in sets[i])
        off[i + 1] = len(flat)
    idx = np.asarray(flat if flat else [0], dtype=np.int32)
    keep = np.zeros(max(1, len(dense)), dtype=np.int32)
    for ix in output:
        keep[dense[ix]] = 1
    pairs = np.asarray(init if init else [(0, 0)], dtype=np.int32).reshape(-1)
    path = np.zeros(2 * max(1, n - 1), dtype=np.int32)
    i32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    n_u = 4 * len(init) + 64 * n if temperature > 0 else 0
    state = rng.getstate()
    while True:
        rng.setstate(state)
        u = np.asarray([rng.random() for _ in range(n_u)] if n_u else [0.0], dtype=np.float64)
        used = C.c_int64(0)
        rc = capi.lib().tq_tn_greedy_path(n, len(dense), off.ctypes.data_as(i32p), idx.ctypes.data_as(i32p),
                                          keep.ctypes.data_as(i32p), len(init), pairs.ctypes.data_as(i32p),
                                          u.ctypes.data_as(f64p), n_u, float(alpha), float(temperature),
                                          path.ctypes.data_as(i32p), C.byref(used))
        if rc == capi.E_WORKSPACE and n_u:
            n_u *= 2
            continue
        capi.check(rc, "tq_tn_greedy_path")
        break
    rng.setstate(state)      # leave rng where the mirror would: exactly the numbers the pass consumed are gone
    for _ in range(used.value):
        rng.random()
    return [(int(path[2 * k]), int(path[2 * k + 1])) for k in range(n - 1)]


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


def sequential_blocked_path(inputs, output, max_block_rank: int = 4):
    """Time-ordered sweep with GATE BLOCKING: like ``sequential_path`` the operands are taken in their (time) order,
    but a small operand is first merged with the pending small tensors it is wired to while the merged tensor keeps
    rank <= ``max_block_rank`` (4 = a two-qubit block [out, out, in, in]); only full blocks are contracted with the
    growing state tensor.  Pending blocks sit on disjoint wires, so they commute and can be flushed in any order.
    The state tensor then sees one gate-like apply step per BLOCK instead of one per gate — what the state-vector
    engine's gate fusion does expressed as a contraction path; the engine runs the applies as
    shared-memory chain sweeps (k_tn_chain)."""
    n = len(inputs)
    count: Dict[int, int] = {}
    for t in inputs:
        for ix in t:
            count[ix] = count.get(ix, 0) + 1
    for ix in output:
        count[ix] = count.get(ix, 0) + 1
    sets: Dict[int, frozenset] = {t: frozenset(inputs[t]) for t in range(n)}
    path: List[Tuple[int, int]] = []
    nxt = [n]

    def result_of(a, b):
        union = sets[a] | sets[b]
        return frozenset(ix for ix in union if count[ix] - (ix in sets[a]) - (ix in sets[b]) > 0)

    def contract(a, b):
        res = result_of(a, b)
        for ix in sets[a] | sets[b]:
            c = count[ix] - (ix in sets[a]) - (ix in sets[b])
            count[ix] = c + 1 if c > 0 else 0
        path.append((a, b))
        new = nxt[0]
        nxt[0] += 1
        sets[new] = res
        del sets[a], sets[b]
        return new

    owner: Dict[int, int] = {}      # open index -> pending block that carries it
    pending: List[int] = []         # pending blocks (tensor ids), oldest first
    state = [None]

    def claim(t):
        for ix in sets[t]:
            owner[ix] = t

    def release(t):
        for ix in sets[t]:
            if owner.get(ix) == t:
                del owner[ix]

    def flush(blk):
        release(blk)
        pending.remove(blk)
        state[0] = blk if state[0] is None else contract(state[0], blk)

    for t in range(n):
        touching = []
        for ix in inputs[t]:
            b = owner.get(ix)
            if b is not None and b not in touching:
                touching.append(b)
        # merged index set if t joins every pending block it is wired to
        merged = set(sets[t])
        for b in touching:
            merged |= sets[b]
        group = touching + [t]
        inner = {ix for ix in merged if count[ix] - sum(ix in sets[g] for g in group) <= 0}
        if touching and len(merged - inner) <= max_block_rank:
            cur = touching[0]
            release(cur)
            pending.remove(cur)
            for b in touching[1:]:
                release(b)
                pending.remove(b)
                cur = contract(cur, b)
            cur = contract(cur, t)
            pending.append(cur)
            claim(cur)
        else:
            for b in touching:
                flush(b)
            pending.append(t)
            claim(t)
    for b in list(pending):
        flush(b)
    return path


# Calibrated on per-step profiles of three config-5 plans on a B200 (scripts/c5_plan_profile.py,
# profiles/r02_plan_profile_*.json; scripts/plan_model_rank.py prints model against measurement).  Tensor-core steps
# follow max(flops / 200 TFLOP/s, operand bytes / 2.5 TB/s) to within 20 % (the byte rate already contains the 2x / 4x
# operand images).  What the 3-value model got wrong is every step the engine's dispatch rule (csrc/tq_tn.cu:
# build_schedule) keeps OFF the tensor cores: it charged them the tensor-core rate.  The tiled FP32 GEMM runs at
# ~25 TFLOP/s, and a step that fits no kernel class — neither >= 128 x 16 free extents, nor >= 64 x 64, nor <= 64
# outputs, nor a gate-sized operand — falls to one thread per output element at ~1.7 TFLOP/s: one such step (K = 2^15,
# 64 x 32 outputs, 10.0 of 12.5 ms per 32 slices) is why the 9.7e11-flop plan of the large search measured 26 ms against
# the 13 ms the old model promised.
CALIBRATED_TIME_MODEL = (2.0e14, 2.5e12, 1.2e-5, 2.5e13, 1.5e12)


def step_time_model(n_a: int, n_b: int, n_out: int, n_union: int, model) -> float:
    """Estimated seconds of one pairwise step on the engine: max(compute time, HBM time) + launch overhead.
    model = (tensor-core algorithmic flop/s, bytes/s, seconds per step[, FP32-GEMM flop/s, per-element-kernel flop/s]);
    ranks are log2 element counts (complex64).  With five values the step is classified the way the engine dispatches
    it ("Dispatch"): tensor cores (>= 128 x 16 free extents, k + m + n + b >= 20), split-K reduction
    (<= 64 outputs, K >= 4096) and gate-sized applies are rate / bandwidth bound as before; >= 64 x 64 x 16 steps run the
    FP32 GEMM; everything else one thread per output element."""
    f, bw, t0 = model[0], model[1], model[2]
    if n_union <= 14:        # small steps ride in a fused run: no launch of their own
        return 1e-7
    if len(model) > 4 and model[3] > 0:
        k = n_union - n_out
        b = n_a + n_b - n_union - k
        m, n = n_a - k - b, n_b - k - b
        if n_union >= 20 and max(m, n) >= 7 and min(m, n) >= 4:
            pass                                           # tensor cores
        elif n_out <= 6 and k >= 12:
            pass                                           # split-K reduction: bandwidth bound
        elif m >= 6 and n >= 6 and k >= 4:
            f = model[3]                                   # tiled FP32 GEMM
        elif (min(m, n) <= 4 and k <= 4) and n_out >= 10 and b <= 8:
            pass                                           # gate-sized operand applied to a large one: bandwidth bound
        else:
            f = model[4]                                   # one thread per output element
    return max(8.0 * 2.0 ** n_union / f, 8.0 * (2.0 ** n_a + 2.0 ** n_b + 2.0 ** n_out) / bw) + t0


def _subtree_dp_py(leaf_sets, leaf_inside, count, time_model):
    """Cheapest pairwise contraction order of a subtree's leaves by dynamic programming over subsets.
    leaf_sets[i]: open indices of leaf i; leaf_inside[i]: {index: number of input tensors below leaf i that carry
    it}; count: {index: number of tensors of the network (+ output) that carry it}.  -> (cost of the whole
    subtree, split) with split[S] = the part of subset S (bit mask over leaves) that is contracted first and
    contains S's lowest leaf.  Python mirror of tq_tn_subtree_order (csrc/tq_planner.cu)."""
    L = len(leaf_sets)
    full = (1 << L) - 1

    def pair_cost(sa, sb, so):
        if time_model is None:
            return 2.0 ** len(sa | sb)
        return step_time_model(len(sa), len(sb), len(so), len(sa | sb), time_model)

    insm = {1 << i: leaf_inside[i] for i in range(L)}
    sidx = {1 << i: leaf_sets[i] for i in range(L)}
    best = {1 << i: 0.0 for i in range(L)}
    split = {}
    for S in range(1, full + 1):
        if S & (S - 1) == 0:
            continue
        low = S & -S
        m = dict(insm[S ^ low])
        for ix, c in insm[low].items():
            m[ix] = m.get(ix, 0) + c
        insm[S] = m
        sidx[S] = frozenset(ix for ix, c in m.items() if c < count[ix])
        b, bs = None, 0
        sub = (S - 1) & S
        while sub:
            if sub & low:   # canonical split: the lowest member stays in the first part
                o = S ^ sub
                c = best[sub] + best[o] + pair_cost(sidx[sub], sidx[o], sidx[S])
                if b is None or c < b:
                    b, bs = c, sub
            sub = (sub - 1) & S
        best[S], split[S] = b, bs
    return best[full], split


def _subtree_dp_native(leaf_sets, leaf_inside, count, time_model):
    """The same through the C ABI (tq_tn_subtree_order): dense arrays over the indices the leaves touch."""
    import ctypes as C

    import numpy as np

    from . import capi

    L = len(leaf_sets)
    idx = {}
    for m in leaf_inside:
        for ix in m:
            idx.setdefault(ix, len(idx))
    n_idx = len(idx)
    open_a = np.zeros((L, max(1, n_idx)), dtype=np.int32)
    inside_a = np.zeros((L, max(1, n_idx)), dtype=np.int32)
    for i in range(L):
        for ix, c in leaf_inside[i].items():
            inside_a[i, idx[ix]] = c
        for ix in leaf_sets[i]:
            open_a[i, idx[ix]] = 1
    count_a = np.zeros(max(1, n_idx), dtype=np.int32)
    for ix, x in idx.items():
        count_a[x] = count[ix]
    split = np.zeros(1 << L, dtype=np.int32)
    best = C.c_double(0.0)
    model = None
    if time_model is not None:
        model = (C.c_double * 5)(*([float(v) for v in time_model] + [0.0, 0.0])[:5])
    i32p = C.POINTER(C.c_int32)
    capi.check(capi.lib().tq_tn_subtree_order(L, n_idx, open_a.ctypes.data_as(i32p), inside_a.ctypes.data_as(i32p),
                                              count_a.ctypes.data_as(i32p), model, C.byref(best),
                                              split.ctypes.data_as(i32p)), "tq_tn_subtree_order")
    return best.value, split


def reconfigure(inputs, output, path, max_leaves: int = 8, sweeps: int = 3, sliced=(), time_model=None,
                native: bool = True):
    """Subtree reconfiguration: for every node of the contraction tree, cut out the subtree spanned by its
    ``max_leaves`` largest descendants, find the cheapest order of contracting those leaves by dynamic programming
    over subsets (cost = sum of 2^|indices of the pair|), and splice it in when it beats the current order.
    Repeats for ``sweeps`` passes or until nothing improves.  ``sliced`` indices are treated as fixed (dropped), so
    the same routine re-optimises the per-slice tree after slicing.  ``time_model`` (see step_time_model) replaces
    the flop count by an estimate of the step's run time, which keeps the tree away from long chains of skinny,
    bandwidth-bound steps that a pure flop count likes.  ``native``: the dynamic programme runs in the compiled
    library (tq_tn_subtree_order); False = its Python mirror, bit-identical.  -> ssa path over the same inputs."""
    import sys

    n_in = len(inputs)
    if n_in < 3:
        return list(path)
    sl = set(sliced)
    children = {n_in + k: (a, b) for k, (a, b) in enumerate(path)}
    root = n_in + len(path) - 1
    count: Dict[int, int] = {}
    for t in inputs:
        for ix in t:
            if ix not in sl:
                count[ix] = count.get(ix, 0) + 1
    for ix in output:
        count[ix] = count.get(ix, 0) + 1
    sys.setrecursionlimit(max(10000, 4 * n_in))

    def compute_sets():
        """index set of every node and, per node, how many leaves below it carry each index"""
        sets, inside = {}, {}
        stack = [(root, False)]
        while stack:
            v, done = stack.pop()
            if v < n_in:
                s = frozenset(ix for ix in inputs[v] if ix not in sl)
                sets[v] = s
                inside[v] = {ix: 1 for ix in s}
            elif not done:
                stack.append((v, True))
                stack.append((children[v][0], False))
                stack.append((children[v][1], False))
            else:
                a, b = children[v]
                m = dict(inside[a])
                for ix, c in inside[b].items():
                    m[ix] = m.get(ix, 0) + c
                inside[v] = m
                sets[v] = frozenset(ix for ix, c in m.items() if c < count[ix])
        return sets, inside

    def pair_cost(sa, sb, so):
        if time_model is None:
            return 2.0 ** len(sa | sb)
        return step_time_model(len(sa), len(sb), len(so), len(sa | sb), time_model)

    next_id = max(children) + 1
    for _ in range(max(1, sweeps)):
        sets, inside = compute_sets()
        improved = False
        for v in sorted(children, key=lambda u: -len(sets[u])):
            if v not in children:
                continue
            frontier, internal = [v], []
            while len(frontier) < max_leaves:
                cand = [u for u in frontier if u in children]
                if not cand:
                    break
                u = max(cand, key=lambda w: len(sets[w]))
                frontier.remove(u)
                internal.append(u)
                frontier += list(children[u])
            L = len(frontier)
            if L < 3:
                continue
            cur = sum(pair_cost(sets[children[u][0]], sets[children[u][1]], sets[u]) for u in internal)
            full = (1 << L) - 1
            dp = _subtree_dp_native if native and L <= 12 else _subtree_dp_py
            best_full, split = dp([sets[u] for u in frontier], [inside[u] for u in frontier], count, time_model)
            if best_full >= cur * 0.999:
                continue
            improved = True
            for u in internal:
                del children[u]
            stack = [(full, v)]      # rebuild the subtree under the same root id
            while stack:
                S, nid = stack.pop()
                parts = []
                first = int(split[S])
                for part in (first, S ^ first):
                    if part & (part - 1) == 0:
                        parts.append(frontier[part.bit_length() - 1])
                    else:
                        parts.append(next_id)
                        stack.append((part, next_id))
                        next_id += 1
                children[nid] = (parts[0], parts[1])
            sets, inside = compute_sets()
        if not improved:
            break
    new_path: List[Tuple[int, int]] = []
    ids: Dict[int, int] = {}
    stack = [(root, False)]
    while stack:             # post-order: children before parents, ssa numbering
        v, done = stack.pop()
        if v < n_in:
            ids[v] = v
        elif not done:
            stack.append((v, True))
            stack.append((children[v][1], False))
            stack.append((children[v][0], False))
        else:
            ids[v] = n_in + len(new_path)
            new_path.append((ids[children[v][0]], ids[children[v][1]]))
    return new_path


def path_time(inputs, output, path, sliced, time_model) -> float:
    """Estimated seconds of ONE slice of a path under ``time_model``."""
    sl = set(sliced)
    sets = [frozenset(ix for ix in t if ix not in sl) for t in inputs]
    _, _, unions, results = path_cost(inputs, output, path, sliced)
    total = 0.0
    for (a, b), u, r in zip(path, unions, results):
        total += step_time_model(len(sets[a]), len(sets[b]), len(r), len(u), time_model)
        sets.append(r)
    return total


def find_path(inputs, output, repeats: int = 16, seed: int = 0, minimize: str = "flops",
              alphas=(1.0, 0.5, 0.0), temperatures=(0.0, 0.3, 1.0), reconf_sweeps: int = 0,
              reconf_leaves: int = 8, time_model=None, native_greedy: bool = True) -> PathInfo:
    """Best of ``repeats`` randomised greedy runs (first run is the deterministic alpha=1, T=0 greedy) and the
    time-ordered sequential path (deep circuits on few qubits: greedy merges wide, the sweep stays at width n).
    ``native_greedy``: the greedy passes run in the compiled library (same paths, same random stream)."""
    rng = random.Random(seed)
    best = None
    if len(inputs) > 1:
        for path in (sequential_path(len(inputs)), sequential_blocked_path(inputs, output)):
            width, fl, _, _ = path_cost(inputs, output, path)
            key = (fl, width) if minimize == "flops" else (width, fl)
            if best is None or key < best[0]:
                best = (key, path, width, fl)
    for r in range(max(1, repeats)):
        alpha = alphas[0] if r == 0 else rng.choice(alphas)
        temp = temperatures[0] if r == 0 else rng.choice(temperatures[1:])
        path = (_greedy_once_native if native_greedy else _greedy_once)(inputs, output, rng, alpha, temp)
        width, fl, _, _ = path_cost(inputs, output, path)
        key = (fl, width) if minimize == "flops" else (width, fl)
        if best is None or key < best[0]:
            best = (key, path, width, fl)
    _, path, width, fl = best
    if reconf_sweeps > 0 and len(inputs) > 2:
        path = reconfigure(inputs, output, path, max_leaves=reconf_leaves, sweeps=reconf_sweeps, time_model=time_model)
        width, fl, _, _ = path_cost(inputs, output, path)
    return PathInfo(path, [], width, fl, len(path))


def slice_path(inputs, output, info: PathInfo, target_size_log2: Optional[int] = None,
               target_num_slices: int = 1, max_sliced: int = 24, reconf_sweeps: int = 0,
               reconf_leaves: int = 8, time_model=None) -> PathInfo:
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
    path = info.path
    if reconf_sweeps > 0 and sliced and len(inputs) > 2:   # re-optimise the per-slice tree (sliced indices fixed)
        cand = reconfigure(inputs, output, path, max_leaves=reconf_leaves, sweeps=reconf_sweeps, sliced=sliced,
                           time_model=time_model)
        w0, f0, _, _ = path_cost(inputs, output, path, sliced)
        w1, f1, _, _ = path_cost(inputs, output, cand, sliced)
        better = f1 < f0 if time_model is None else \
            path_time(inputs, output, cand, sliced, time_model) < path_time(inputs, output, path, sliced, time_model)
        if better and (target_size_log2 is None or w1 <= max(w0, target_size_log2)):
            path = cand
    width, fl, _, _ = path_cost(inputs, output, path, sliced)
    return PathInfo(path, sliced, width, fl, len(path))


def _search_once(args):
    """One search (find_path + slice_path) for one seed -> (score, PathInfo fields); module-level so that worker
    processes can run it."""
    (inputs, output, max_repeats, seed, minimize, reconf_sweeps, reconf_leaves, time_model, target_size,
     target_num_slices) = args
    t0 = time.time()
    info = find_path(inputs, output, repeats=int(max_repeats), seed=int(seed), minimize=minimize,
                     reconf_sweeps=int(reconf_sweeps), reconf_leaves=int(reconf_leaves), time_model=time_model)
    tnum = int(target_num_slices or 1)
    if target_size or tnum > 1:
        info = slice_path(inputs, output, info, target_size_log2=int(math.log2(target_size)) if target_size else None,
                          target_num_slices=tnum, reconf_sweeps=min(3, int(reconf_sweeps)),
                          reconf_leaves=int(reconf_leaves), time_model=time_model)
    if time_model is not None:
        score = path_time(inputs, output, info.path, info.sliced, time_model) * info.n_slices
    else:
        score = info.flops_log2 + len(info.sliced)
    return score, [list(p) for p in info.path], list(info.sliced), info.width, info.flops_log2, time.time() - t0


def _search_in_subprocesses(jobs, workers: int):
    """Each job in its own interpreter (``python -c "... planner._worker_main()"``, job and result as JSON on stdin /
    stdout), ``workers`` at a time.  Plain subprocesses rather than multiprocessing: nothing of the caller's main module
    is imported or re-run in the children, and they never see a GPU."""
    import json
    import subprocess
    import sys
    from concurrent.futures import ThreadPoolExecutor

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    env["CUDA_VISIBLE_DEVICES"] = ""
    code = "from tedq_b200 import planner; planner._worker_main()"

    def run(job):
        res = subprocess.run([sys.executable, "-c", code], input=json.dumps(job), capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError(res.stderr[-2000:])
        return tuple(json.loads(res.stdout.strip().splitlines()[-1]))

    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(run, jobs))


def _worker_main():
    import json
    import sys

    job = json.loads(sys.stdin.read())
    job[7] = None if job[7] is None else tuple(job[7])       # time_model
    print(json.dumps(list(_search_once(tuple(job)))))


def search_plan(inputs, output, max_repeats: int = 16, seed: int = 0, minimize: str = "flops", reconf_sweeps: int = 0,
                reconf_leaves: int = 8, time_model=None, target_size=None, target_num_slices: int = 1,
                restarts: int = 1, workers: Optional[int] = None) -> PathInfo:
    """The whole search the executors and the bench share: ``find_path`` + ``slice_path``, ``restarts`` times with
    seeds seed, seed + 1, ...; the plan with the smallest estimated run time (``path_time`` x slices under
    ``time_model``; flops without a model) is kept, ties to the lower seed.  The restarts are independent searches:
    what ranks them is the calibrated step-time model, which is the point of having one ("Planner").
    They run in ``workers`` child interpreters (default: one per core, at most one per restart; 0 / 1 = in this
    process) — the result does not depend on it."""
    restarts = max(1, int(restarts))
    inputs = [[int(i) for i in t] for t in inputs]
    output = [int(i) for i in output]
    jobs = [(inputs, output, max_repeats, int(seed) + r, minimize, reconf_sweeps, reconf_leaves,
             None if time_model is None else tuple(time_model), target_size, target_num_slices) for r in range(restarts)]
    if workers is None:
        workers = min(restarts, os.cpu_count() or 1)
    results = None
    if restarts > 1 and workers > 1:
        try:
            results = _search_in_subprocesses(jobs, workers)
        except Exception:            # no child processes in this environment: search in-process
            results = None
    if results is None:
        results = [_search_once(j) for j in jobs]
    best = min(range(restarts), key=lambda r: (results[r][0], r))
    _, path, sliced, width, fl, _ = results[best]
    info = PathInfo([tuple(p) for p in path], sliced, width, fl, len(path))
    return info


PLANNER_VERSION = 2   # bump when the search changes: stored plans of another version are searched again


def valid_plan(inputs, output, ssa, sliced) -> bool:
    """A stored path fits the network: n-1 steps, every tensor id consumed exactly once and only after it was
    produced, sliced indices are indices of the network and none of them is an open output."""
    n = len(inputs)
    if len(ssa) != n - 1:
        return False
    used = set()
    for s, pair in enumerate(ssa):
        if len(pair) != 2 or pair[0] == pair[1]:
            return False
        for t in pair:
            if not isinstance(t, int) or t < 0 or t >= n + s or t in used:
                return False
            used.add(t)
    if len(used) != 2 * (n - 1) or (n + len(ssa) - 1) in used:
        return False
    all_idx = {int(i) for t in inputs for i in t}
    out = {int(i) for i in output}
    return len(set(sliced)) == len(sliced) and all(i in all_idx and i not in out for i in sliced)


def cached_plan(cache_dir, inputs, output, build, **key_args) -> PathInfo:
    """On-disk plan cache ("plan cache keyed by the index-map hash"): ``build()`` -> PathInfo runs only
    when no valid file for sha1(index maps, output, key_args) exists under ``cache_dir``.  A stored plan is
    validated (``valid_plan``) and re-costed on load; a stale, foreign or other-version file is searched again
    and overwritten."""
    import hashlib
    import json
    import os

    if not cache_dir:
        import time

        t0 = time.perf_counter()
        info = build()
        info.search_s = time.perf_counter() - t0
        return info
    blob = json.dumps({"inputs": [list(map(int, t)) for t in inputs], "output": list(map(int, output)),
                       "args": {k: (list(v) if isinstance(v, tuple) else v) for k, v in sorted(key_args.items())}},
                      sort_keys=True).encode()
    path = os.path.join(cache_dir, hashlib.sha1(blob).hexdigest() + ".json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                d = json.load(fh)
            ssa = [tuple(int(x) for x in p) for p in d["path"]]
            sliced = [int(i) for i in d["sliced"]]
            if int(d.get("version", 1)) == PLANNER_VERSION and valid_plan(inputs, output, ssa, sliced):
                width, fl, _, _ = path_cost(inputs, output, ssa, sliced)
                info = PathInfo(ssa, sliced, width, fl, len(ssa))
                info.search_s = d.get("search_s")
                info.from_cache = True
                return info
        except (OSError, ValueError, KeyError, IndexError, TypeError):
            pass
    import time

    t0 = time.perf_counter()
    info = build()
    info.search_s = time.perf_counter() - t0
    os.makedirs(cache_dir, exist_ok=True)
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "w") as fh:
        json.dump({"version": PLANNER_VERSION, "search_s": round(info.search_s, 2),
                   "path": [list(p) for p in info.path], "sliced": list(info.sliced)}, fh)
    os.replace(tmp, path)
    return info
