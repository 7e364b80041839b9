"""helmholtz.jl_b200 -- B200-native shifted-Laplacian multigrid Helmholtz solve path.

Layout: `csrc/` CUDA kernels + C ABI (built into `lib/libhelmholtz_b200.so`), `api.py` the host-side
mirror of the reference's Julia interface, `julia/` the Julia shim, `workloads.py` synthetic models of
the benchmark configurations, `sharding.py` the RHS-to-rank partition.

The directory name contains a dot, so import it through `__graft_entry__.load_package()`.
"""
from . import _lib  # noqa: F401
from .api import *  # noqa: F401,F403
from . import api, sharding, workloads  # noqa: F401
