"""Hook ``backend="pytorch_b200"`` into the reference's own ``Circuit.compilecircuit``.

The reference dispatches on a string (tedq/QInterpreter/circuits/circuit.py:268-285) and its tree is
read-only here, so the one extra branch a maintainer would add is installed at
import time by wrapping the method.  Every other backend string falls through to the original."""
from __future__ import annotations

from .backend import BACKEND_NAME, QUDIO_BACKEND_NAME, B200Backend, B200QUDIOBackend


def register_backend(tedq_module=None):
    """Idempotent.  Returns the patched ``Circuit`` class."""
    if tedq_module is None:
        import tedq as tedq_module  # noqa: F401  (must be importable: PYTHONPATH=/path/to/TeD-Q)
    from tedq.QInterpreter.circuits.circuit import Circuit

    if getattr(Circuit.compilecircuit, "_tedq_b200", False):
        return Circuit
    original = Circuit.compilecircuit

    def compilecircuit(self, backend=None, **kwargs):
        if backend == BACKEND_NAME:
            return B200Backend(backend, self, **kwargs)
        if backend == QUDIO_BACKEND_NAME:      # the "pytorch_QUDIO" branch's twin (circuit.py:276-277)
            return B200QUDIOBackend(backend, self, **kwargs)
        return original(self, backend=backend, **kwargs)

    compilecircuit._tedq_b200 = True
    compilecircuit.__doc__ = original.__doc__
    Circuit.compilecircuit = compilecircuit
    return Circuit
