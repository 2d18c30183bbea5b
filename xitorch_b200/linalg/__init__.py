from xitorch_b200.linalg.solve import solve                                   # noqa: F401
from xitorch_b200.linalg.symeig import symeig, lsymeig, usymeig, svd          # noqa: F401

__all__ = ["solve", "symeig", "lsymeig", "usymeig", "svd"]
