"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol the header declares; host-only entry
points give the reference's integers.  No kernel is launched here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, load_golden
from tedq_b200 import build, capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tedq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/tedq_b200.h but not exported: {missing}"
    assert capi.lib().tq_abi_version() == 1


def test_library_contains_sm100a_code():
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("case", load_golden("sv_cases.json")[:30], ids=lambda c: c["spec"]["name"])
def test_c_sv_plan_integers_match_reference(case):
    """tq_sv_axes_perm reproduces the reference's per-gate (axes, permutation) (compiled_circuit.py:126-202)."""
    n = case["spec"]["num_qubits"]
    gates = case["spec"]["gates"]
    axes = list(reversed(case["axeslist"]))
    perms = list(reversed(case["permutationlist"]))
    for (name, qubits, _), a, p in zip(gates, axes, perms):
        gp, pm = capi.sv_axes_perm(n, qubits)
        assert gp == a[0] and list(qubits) == a[1] and pm == p


def test_no_cpu_fallback_without_cuda():
    import torch

    import tedq_b200 as qb

    if torch.cuda.is_available():
        pytest.skip("CUDA present")

    def circuit_def(t):
        qb.RX(t[0], qubits=[0])
        return qb.expval(qb.PauliZ(qubits=[0]))

    cc = qb.Circuit(circuit_def, 1, torch.tensor([0.1])).compilecircuit(backend="pytorch_b200")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cc(torch.tensor([0.1]))


def test_qudio_front_host_contract_without_cuda():
    """qudio_backend.py:86-111 twin: dataset bookkeeping and error behaviour of ``pytorch_QUDIO_b200`` that need no
    device — no dataset -> ValueError; CPU parameters -> the same loud no-fallback error as the plain backend."""
    import torch

    import tedq_b200 as qb

    def circuit_def(x, w):
        qb.RX(x[0], qubits=[0])
        qb.RY(w[0], qubits=[0])
        return qb.expval(qb.PauliZ(qubits=[0]))

    circ = qb.Circuit(circuit_def, 1, torch.tensor([0.1]), torch.tensor([0.2]))
    cq = circ.compilecircuit(backend="pytorch_QUDIO_b200")
    assert isinstance(cq, qb.B200QUDIOBackend) and cq.backend == "pytorch_QUDIO_b200"
    with pytest.raises(ValueError, match="set_dataset"):
        cq(torch.tensor([0.2]))
    cq.set_dataset([0.1, 0.2, 0.3])                 # 1-D data: one value per row
    assert tuple(cq.dataset().shape) == (3, 1)
    d = torch.rand(5, 1)
    cq.set_dataset(d)
    assert cq.dataset() is d
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            cq(torch.tensor([0.2]))
    with pytest.raises(ValueError, match="unknown backend input"):
        circ.compilecircuit(backend="pytorch_QUDIO")


def test_product_sources_never_touch_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under the package (Python or CUDA) may import, open or name it,
    nor read /root/reference at run time; bench.py may use it only in its CPU legs and __graft_entry__ only in
    smoke()."""
    import ast
    import os

    from conftest import ROOT

    pkg = os.path.join(ROOT, "ted-q_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(dirpath, f), encoding="utf-8").read()
            assert "/root/reference" not in text, f
            if f.endswith(".py"):
                tree = ast.parse(text)
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (f, names)
    # bench.py: oracle imports live inside functions of the CPU legs only (never at module level)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read())
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            names = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names)


def test_dist_slice_range_matches_python_sharding():
    """tq_dist_slice_range (the C ABI's multi-GPU entry) deals slices exactly like dist.shard_range does for the
    torch.distributed path: contiguous blocks of ceil(n / world), every slice exactly once."""
    from tedq_b200 import capi, dist

    for n in (0, 1, 5, 8, 64, 65, 1000):
        for world in (1, 2, 3, 4, 8):
            got = [capi.slice_range(n, r, world) for r in range(world)]
            assert got == [dist.shard_range(n, r, world) for r in range(world)]
            covered = [s for b, e in got for s in range(b, e)]
            assert covered == list(range(n))
    d = capi.Dist(0, 0, 1)      # world 1 needs no communicator; all-reduce is a no-op
    d.allreduce(0, 0, capi.TQ_C64, 0)
