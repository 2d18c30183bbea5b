"""
oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (torch-CPU / ATen, exactly the arithmetic backend the
reference itself uses) of the xitorch Krylov hot path:

  * davidson                      (/root/reference/xitorch/_impls/linalg/symeig.py:100-227)
  * tallqr / to_fortran_order     (/root/reference/xitorch/_utils/tensor.py:8-32)
  * cg / bicgstab / gmres         (/root/reference/xitorch/_impls/linalg/solve.py:69-433)
  * _setup_linear_problem & co.   (/root/reference/xitorch/_impls/linalg/solve.py:437-445,540-663)
  * dense LinearOperator mm/rmm   (/root/reference/xitorch/_core/linop.py:676-708)
  * broyden1 rootfinder + the implicit gradient of its backward (oracle/rootfinder.py;
    /root/reference/xitorch/_impls/optimize/root/*.py, /root/reference/xitorch/optimize/rootfinder.py:331-366)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and only as the checker or
the CPU baseline -- never as the thing measured or shipped.  The product
package `xitorch_b200` must not import it (tests/test_no_oracle_in_product.py
enforces that).

Parity pinning: the reference publishes no golden vectors for this path
(SURVEY.md 8c).  The oracle is therefore pinned against outputs of the
reference itself, generated in the build container by
`oracle/gen_golden.py` (which imports /root/reference) and committed under
`tests/golden/`; `tests/test_oracle_golden.py` replays them without the
reference being present.
"""
from oracle.krylov import (  # noqa: F401
    DenseOp, tallqr, to_fortran_order, davidson, cg, bicgstab, gmres,
    setup_linear_problem, largest_eival_power, exacteig, exactsolve,
)
from oracle.rootfinder import broyden1_root, implicit_grad_dense  # noqa: F401
from oracle.problems import (  # noqa: F401
    make_herm, make_slow_herm, make_spd_c1, make_nonsym_c3, make_rootfinder_c4, make_herm_row_block,
)
