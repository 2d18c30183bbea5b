"""
xitorch_b200 -- B200-native (sm_100a) implementation of xitorch's iterative Krylov hot path behind
xitorch's own API surface: `LinearOperator`, `linalg.symeig/lsymeig/usymeig/svd`, `linalg.solve` and the
`method=` plug-in point (see DESIGN.md / INTEGRATION.md).

    import xitorch_b200 as xitorch
    A = xitorch.LinearOperator.m(mat, is_hermitian=True)            # mat: CUDA tensor
    evals, evecs = xitorch.linalg.symeig(A, neig=8, method="davidson", min_eps=1e-4)
    x = xitorch.linalg.solve(A, B, method="cg")

The Krylov methods run hand-written CUDA kernels through a C ABI (include/xitorch_b200.h); there is no
CPU fallback for them.
"""
from xitorch_b200.editable_module import EditableModule                # noqa: F401
from xitorch_b200.pure_function import get_pure_function, make_sibling  # noqa: F401
from xitorch_b200.linop import LinearOperator, MatrixLinearOperator   # noqa: F401
from xitorch_b200._utils import ConvergenceWarning, MathWarning        # noqa: F401
from xitorch_b200.debug import is_debug_enabled, set_debug_mode, enable_debug, disable_debug  # noqa: F401
from xitorch_b200 import linalg                                        # noqa: F401
from xitorch_b200 import optimize, grad                                # noqa: F401

from xitorch_b200.compat import install_as_xitorch                     # noqa: F401

__version__ = "0.1.0"
