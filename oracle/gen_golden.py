"""
Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from
/root/reference, present only in the build container) on small seeded inputs,
and check the oracle restatement against it while doing so.

    python oracle/gen_golden.py            # rewrites tests/golden/

TEST INFRASTRUCTURE (see oracle/__init__.py).  The fixtures (inputs + reference
outputs) are committed; tests/test_oracle_golden.py replays them without the
reference being present (it does not exist on the GPU box).
"""
import os
import sys
import warnings

import torch

REF = os.environ.get("XITORCH_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import xitorch                                           # noqa: E402  (the reference)
from xitorch import LinearOperator as RefLinOp           # noqa: E402
from xitorch.linalg import symeig as ref_symeig, solve as ref_solve   # noqa: E402
from xitorch._utils.tensor import tallqr as ref_tallqr   # noqa: E402

import oracle                                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
warnings.simplefilter("ignore")


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def check(name, a, b, rtol, atol=0.0):
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    ok = err <= atol + rtol * scale
    print("  %-34s max|oracle-ref| = %.3e (scale %.3e) %s" % (name, err, scale, "ok" if ok else "MISMATCH"))
    assert ok, name


cases = {}

# ---------------------------------------------------------------- tallqr
V = torch.randn(64, 6, generator=gen(1), dtype=torch.float64)
q_ref, r_ref = ref_tallqr(V)
q_o, r_o = oracle.tallqr(V)
check("tallqr Q", q_o, q_ref, 0.0)
cases["tallqr"] = {"V": V, "Q": q_ref, "R": r_ref}

# ---------------------------------------------------------------- davidson
print("davidson")
dav = []
for (n, neig, mode, dtype, min_eps, batch) in [
        (128, 4, "lowest", torch.float64, 1e-8, ()),
        (128, 3, "uppest", torch.float64, 1e-8, ()),
        (96, 2, "lowest", torch.float64, 1e-8, (2,)),
        (512, 4, "lowest", torch.float32, 1e-4, ()),
]:
    if batch:
        A = torch.stack([oracle.make_herm(n, neig, torch.float64, seed=10 + b).to(dtype) for b in range(batch[0])])
    else:
        A = oracle.make_herm(n, neig, torch.float64, seed=7).to(dtype)
    Aop = RefLinOp.m(A, is_hermitian=True)
    ev_ref, vec_ref = ref_symeig(Aop, neig=neig, mode=mode, method="davidson", min_eps=min_eps)
    ev_o, vec_o, info = oracle.davidson(A, neig, mode, min_eps=min_eps, return_info=True)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    check("davidson evals n=%d %s" % (n, mode), ev_o, ev_ref, tol)
    check("davidson |evecs| n=%d %s" % (n, mode), vec_o.abs(), vec_ref.abs(), tol * 100, atol=tol * 100)
    ev_exact = torch.linalg.eigvalsh(A.double())
    ev_exact = ev_exact[..., :neig] if mode == "lowest" else ev_exact[..., -neig:]
    dav.append({"n": n, "neig": neig, "mode": mode, "dtype": dtype, "min_eps": min_eps,
                "A": A if n <= 128 else None, "seed": 7, "batch": batch,
                "evals": ev_ref, "evecs_abs": vec_ref.abs() if n <= 128 else None,
                "evals_exact_f64": ev_exact, "oracle_niter": info["niter"]})
cases["davidson"] = dav

# generalized problem A x = lambda M x
n, neig = 64, 3
A = oracle.make_herm(n, neig, torch.float64, seed=21)
Mm = torch.randn(n, n, generator=gen(22), dtype=torch.float64) * 0.05
Mm = Mm @ Mm.t() + torch.eye(n, dtype=torch.float64)
ev_ref, vec_ref = ref_symeig(RefLinOp.m(A, True), neig=neig, M=RefLinOp.m(Mm, True),
                             method="davidson", min_eps=1e-9)
ev_o, vec_o = oracle.davidson(A, neig, "lowest", M=oracle.DenseOp(Mm, True), min_eps=1e-9)
check("davidson generalized evals", ev_o, ev_ref, 1e-12)
cases["davidson_M"] = {"A": A, "M": Mm, "neig": neig, "min_eps": 1e-9, "evals": ev_ref,
                       "evecs_abs": vec_ref.abs()}

# ---------------------------------------------------------------- solve
print("solve")
sol = []


def run_solve(tag, method, A, B, E=None, M=None, herm=None, seed_call=4321, **opts):
    Aop = RefLinOp.m(A, is_hermitian=herm)
    Mop = None if M is None else RefLinOp.m(M, is_hermitian=True)
    torch.manual_seed(seed_call)          # the posdef probe draws from the global RNG
    x_ref = ref_solve(Aop, B, E, Mop, method=method, **opts)
    torch.manual_seed(seed_call)
    fn = getattr(oracle, method)
    x_o, info = fn(oracle.DenseOp(A, herm), B, E, None if M is None else oracle.DenseOp(M, True),
                   return_info=True, **opts)
    check("%s %s" % (method, tag), x_o, x_ref, 1e-11, atol=1e-13)
    x_exact = oracle.exactsolve(A, B, E, M)
    sol.append({"tag": tag, "method": method, "A": A, "B": B, "E": E, "M": M, "herm": herm,
                "opts": opts, "seed_call": seed_call, "x": x_ref, "x_exact": x_exact,
                "oracle_niter": info["niter"], "oracle_napply": info["napply"]})


# C1: cg on the asv SPD generator, N=256 fp64 (BASELINE configs[0])
A1 = oracle.make_spd_c1(256)
torch.manual_seed(123)
X1 = torch.randn(256, 3, dtype=torch.float64)
B1 = A1 @ X1
run_solve("C1 default", "cg", A1, B1, herm=True)
run_solve("C1 tight posdef", "cg", A1, B1, herm=True, posdef=True, rtol=1e-10, atol=1e-12)

# the reference test matrices (test_linop_fcns.py:474-524): A = 0.1*rand + I symmetrised
n = 100
A2 = torch.rand(n, n, generator=gen(31), dtype=torch.float64) * 0.1 + torch.eye(n, dtype=torch.float64)
A2s = (A2 + A2.t()) * 0.5
B2 = torch.rand(2, n, 5, generator=gen(32), dtype=torch.float64) + 0.1
run_solve("sym batchedB", "cg", A2s, B2, herm=True, rtol=1e-8)
run_solve("sym batchedB", "bicgstab", A2s, B2, herm=True, rtol=1e-8)
run_solve("nonsym batchedB", "bicgstab", A2, B2, herm=False, rtol=1e-8, posdef=True)
run_solve("nonsym probe", "bicgstab", A2, B2, herm=False, rtol=1e-8)
run_solve("nonsym->normal eq", "cg", A2, B2, herm=False, rtol=1e-9)
run_solve("sym", "gmres", A2s, B2[0, :, :2].contiguous(), herm=True, posdef=True)
run_solve("nonsym", "gmres", A2, B2[0, :, :2].contiguous(), herm=False, posdef=True)

# with E and M (test_linop_fcns.py:631-676)
n = 50
A3 = torch.rand(n, n, generator=gen(41), dtype=torch.float64) * 0.1 + torch.eye(n, dtype=torch.float64)
A3 = (A3 + A3.t()) * 0.5
M3 = torch.rand(n, n, generator=gen(42), dtype=torch.float64) * 0.05
M3 = (M3 + M3.t()) * 0.5 + 0.5 * torch.eye(n, dtype=torch.float64)
B3 = torch.rand(n, 4, generator=gen(43), dtype=torch.float64)
E3 = torch.rand(4, generator=gen(44), dtype=torch.float64) * 0.1
run_solve("AE", "cg", A3, B3, E=E3, herm=True, rtol=1e-9, posdef=True)
run_solve("AEM", "cg", A3, B3, E=E3, M=M3, herm=True, rtol=1e-9, posdef=True)
run_solve("AEM", "bicgstab", A3, B3, E=E3, M=M3, herm=True, rtol=1e-9, posdef=True)
run_solve("AE probe", "bicgstab", A3, B3, E=E3, herm=True, rtol=1e-9)
cases["solve"] = sol

# preconditioned cg / bicgstab (solve.py:122,136,171 and :247-248,276-287): Jacobi-like and approximate-inverse
# preconditioners on an ill-scaled system
print("preconditioned solve")
pre = []
n = 80
dg = torch.logspace(0, 3, n, dtype=torch.float64)
R4 = torch.rand(n, n, generator=gen(51), dtype=torch.float64) * 0.2
A4s = torch.diag(dg) + (R4 + R4.t()) * 0.5
A4n = torch.diag(dg) + R4
B4 = torch.rand(n, 3, generator=gen(52), dtype=torch.float64)
Pj = torch.diag(1.0 / dg)                                        # Jacobi
Pa = torch.linalg.inv(A4n + 0.05 * torch.diag(dg))               # approximate inverse
for tag, method, A, herm, kw in [
        ("jacobi", "cg", A4s, True, {"precond": Pj}),
        ("jacobi right", "bicgstab", A4n, False, {"precond_r": Pj}),
        ("approx-inverse left+right", "bicgstab", A4n, False, {"precond_l": Pa, "precond_r": Pj}),
        ("jacobi left", "bicgstab", A4n, False, {"precond_l": Pj})]:
    ref_kw = {k: RefLinOp.m(v, is_hermitian=bool(torch.allclose(v, v.t()))) for k, v in kw.items()}
    or_kw = {k: oracle.DenseOp(v, bool(torch.allclose(v, v.t()))) for k, v in kw.items()}
    opts = {"rtol": 1e-10, "atol": 1e-14, "posdef": True}
    x_ref = ref_solve(RefLinOp.m(A, is_hermitian=herm), B4, method=method, **opts, **ref_kw)
    x_o, info = getattr(oracle, method)(oracle.DenseOp(A, herm), B4, return_info=True, **opts, **or_kw)
    x_plain, info0 = getattr(oracle, method)(oracle.DenseOp(A, herm), B4, return_info=True, **opts)
    check("%s %s" % (method, tag), x_o, x_ref, 1e-11, atol=1e-13)
    print("     iterations with / without the preconditioner: %d / %d" % (info["niter"], info0["niter"]))
    pre.append({"tag": tag, "method": method, "A": A, "B": B4, "herm": herm, "precond": kw, "opts": opts,
                "x": x_ref, "x_exact": torch.linalg.solve(A, B4), "oracle_niter": info["niter"],
                "plain_niter": info0["niter"]})
cases["solve_precond"] = pre

torch.save(cases, os.path.join(OUT, "krylov_golden.pt"))
print("wrote", os.path.join(OUT, "krylov_golden.pt"),
      "%.1f KiB" % (os.path.getsize(os.path.join(OUT, "krylov_golden.pt")) / 1024),
      "reference xitorch", xitorch.__version__, "torch", torch.__version__)
