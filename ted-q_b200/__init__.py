"""tedq_b200 — B200-native execution engine behind TeD-Q's ``compilecircuit(backend="pytorch_b200")``.

Import name is ``tedq_b200`` (the directory is ``ted-q_b200/``; ``tedq_b200/__init__.py`` at the repo
root aliases it because a hyphen is not importable)."""
from .backend import (BACKEND_NAME, QUDIO_BACKEND_NAME, B200Backend, B200Execute, B200Grad,  # noqa: F401
                      B200ParamShift, B200QUDIOBackend)
from .frontend import *  # noqa: F401,F403
from .frontend import Circuit  # noqa: F401
from .register import register_backend  # noqa: F401
from .tree_compat import B200OptTN, ctg_compat  # noqa: F401

__version__ = "0.1.0"
