"""
Generate tests/golden/generalized_golden.pt by running the UNMODIFIED reference (imported from /root/reference, present
only in the build container) on generalized eigenproblems A x = lambda M x and on start blocks wider than neig
(`nguess`, reference symeig.py:137-138), and check the oracle restatement against it while doing so.

    python oracle/gen_golden_generalized.py        # rewrites tests/golden/generalized_golden.pt

TEST INFRASTRUCTURE (see oracle/__init__.py).  The fixture (inputs + reference outputs) is committed;
tests/test_oracle_golden.py and tests/test_davidson_generalized_host.py replay it without the reference being present.
"""
import os
import sys
import warnings

import torch

REF = os.environ.get("XITORCH_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from xitorch import LinearOperator as RefLinOp           # noqa: E402  (the reference)
from xitorch.linalg import symeig as ref_symeig          # noqa: E402

import oracle                                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
warnings.simplefilter("ignore")


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def spd(n, seed, weight=0.05):
    m = torch.randn(n, n, generator=gen(seed), dtype=torch.float64) * weight
    return m @ m.t() + torch.eye(n, dtype=torch.float64)


cases = []
for tag, n, neig, nguess, mode, with_m, batch, min_eps in (
        ("generalized-lowest", 96, 4, None, "lowest", True, (), 1e-9),
        ("generalized-uppest", 96, 3, None, "uppest", True, (), 1e-9),
        ("wide-start", 120, 4, 7, "lowest", False, (), 1e-9),
        ("wide-start-generalized", 120, 3, 6, "lowest", True, (), 1e-9),
        ("generalized-batched", 48, 2, None, "lowest", True, (2,), 1e-9),
        ("generalized-full-space", 9, 2, None, "lowest", True, (), 1e-9)):
    seed = 100 + len(cases)
    if batch:
        A = torch.stack([oracle.make_herm(n, neig, torch.float64, seed=seed + 10 * i) for i in range(batch[0])])
    else:
        A = oracle.make_herm(n, neig, torch.float64, seed=seed)
    Mm = spd(n, seed + 1) if with_m else None
    kw = {} if nguess is None else {"nguess": nguess}
    ev_ref, vec_ref = ref_symeig(RefLinOp.m(A, True), neig=neig, mode=mode,
                                 M=RefLinOp.m(Mm, True) if with_m else None, method="davidson", min_eps=min_eps, **kw)
    ev_o, vec_o, info = oracle.davidson(A, neig, mode, M=oracle.DenseOp(Mm, True) if with_m else None,
                                        min_eps=min_eps, return_info=True, **kw)
    err = ((ev_o - ev_ref).abs() / ev_ref.abs()).max().item()
    print("  %-26s n=%3d neig=%d nguess=%s %s  max rel |oracle - ref| = %.2e  oracle niter %d"
          % (tag, n, neig, nguess, mode, err, info["niter"]))
    assert err <= 1e-12, tag
    Md = Mm if with_m else torch.eye(n, dtype=torch.float64)
    resid = (A @ vec_ref - Md @ vec_ref * ev_ref.unsqueeze(-2)).abs().max().item()
    cases.append({"tag": tag, "A": A, "M": Mm, "neig": neig, "nguess": nguess, "mode": mode, "min_eps": min_eps,
                  "evals": ev_ref, "evecs_abs": vec_ref.abs(), "ref_resid": resid, "oracle_niter": info["niter"]})

torch.save({"davidson_generalized": cases}, os.path.join(OUT, "generalized_golden.pt"))
print("wrote", os.path.join(OUT, "generalized_golden.pt"),
      "%.1f KiB" % (os.path.getsize(os.path.join(OUT, "generalized_golden.pt")) / 1024))
