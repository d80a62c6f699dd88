"""Fixture of BASELINE config 4 at "depth 20" (10 Hd + 10 H0 Trotter sweeps of the 4x4 MBL-2D circuit,
17 819 gates, complex128) from the UNMODIFIED reference, CPU only:

    python tests/golden/generate_golden_c4d20.py        # needs /root/reference; ~1 min

Writes tests/golden/c4d20_case.json.  The circuit is rebuilt from ``workloads.mbl_2d(4, 10)`` by the test (the
gate list itself would be 1 MB of JSON), so the file only holds parameters, outputs, cotangent and gradients.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import generate_golden as G  # noqa: E402  (stubs the reference's optional imports, imports tedq)
from tedq_b200 import workloads as W  # noqa: E402


def main():
    rng = np.random.RandomState(20)
    spec = W.mbl_2d(4, 10)
    flat = rng.uniform(0, 1, size=(1, spec["n_params"]))
    case = G.run_case(spec, flat.tolist(), "c128", seed=20)
    n_gates = len(spec["gates"])
    case["spec"] = {"workload": "mbl_2d", "args": [4, 10], "name": spec["name"], "num_qubits": spec["num_qubits"],
                    "n_params": spec["n_params"], "n_gates": n_gates, "meas": spec["meas"]}
    del case["axeslist"], case["permutationlist"]
    with open(os.path.join(HERE, "c4d20_case.json"), "w") as fh:
        json.dump(case, fh)
    print("c4d20:", n_gates, "gates, out", case["out"])


if __name__ == "__main__":
    main()
