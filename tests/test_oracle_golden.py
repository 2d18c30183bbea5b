"""The CPU oracle replayed against the committed reference outputs (tests/golden/krylov_golden.pt, produced by
oracle/gen_golden.py from the unmodified reference).  Runs without the reference being present."""
import warnings

import torch

import oracle


def test_tallqr(golden):
    c = golden["tallqr"]
    Q, R = oracle.tallqr(c["V"])
    assert torch.equal(Q, c["Q"]) or torch.allclose(Q, c["Q"], rtol=0, atol=1e-14)
    assert torch.allclose(R, c["R"], rtol=0, atol=1e-14)
    assert torch.allclose(Q.t() @ Q, torch.eye(Q.shape[-1], dtype=Q.dtype), atol=1e-12)


def test_davidson_matches_reference(golden):
    for case in golden["davidson"]:
        n, neig, mode, dtype = case["n"], case["neig"], case["mode"], case["dtype"]
        if case["A"] is not None:
            A = case["A"]
        else:
            A = oracle.make_herm(n, neig, torch.float64, seed=case["seed"]).to(dtype)
        ev, vec, info = oracle.davidson(A, neig, mode, min_eps=case["min_eps"], return_info=True)
        tol = 1e-11 if dtype == torch.float64 else 2e-5
        assert torch.allclose(ev, case["evals"], rtol=tol, atol=tol), (n, mode)
        assert info["niter"] == case["oracle_niter"]
        if case["evecs_abs"] is not None:
            assert torch.allclose(vec.abs(), case["evecs_abs"], rtol=0, atol=1e-8)
        # and against the exact fp64 spectrum at the north_star tolerance
        rel = ((ev.double() - case["evals_exact_f64"]).abs() / case["evals_exact_f64"].abs()).max().item()
        assert rel <= 1e-5


def test_davidson_generalized_matches_reference(golden):
    c = golden["davidson_M"]
    ev, vec = oracle.davidson(c["A"], c["neig"], "lowest", M=oracle.DenseOp(c["M"], True), min_eps=c["min_eps"])
    assert torch.allclose(ev, c["evals"], rtol=1e-11, atol=1e-12)
    assert torch.allclose(vec.abs(), c["evecs_abs"], rtol=0, atol=1e-8)


def test_davidson_generalized_and_wide_start_match_reference(golden_generalized):
    """generalized problems (both modes, batched, the full-space exit) and nguess > neig: the oracle reproduces the
    reference's eigenvalues to the last bits and its iteration count"""
    for c in golden_generalized:
        Mop = None if c["M"] is None else oracle.DenseOp(c["M"], True)
        kw = {} if c["nguess"] is None else {"nguess": c["nguess"]}
        ev, vec, info = oracle.davidson(c["A"], c["neig"], c["mode"], M=Mop, min_eps=c["min_eps"], return_info=True, **kw)
        assert torch.allclose(ev, c["evals"], rtol=1e-12, atol=1e-13), c["tag"]
        assert torch.allclose(vec.abs(), c["evecs_abs"], rtol=0, atol=1e-8), c["tag"]
        assert info["niter"] == c["oracle_niter"], c["tag"]
        assert c["ref_resid"] <= 20 * c["min_eps"], c["tag"]          # the reference's answer solves the pencil


def test_solvers_match_reference(golden):
    for case in golden["solve"]:
        fn = getattr(oracle, case["method"])
        torch.manual_seed(case["seed_call"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, info = fn(oracle.DenseOp(case["A"], case["herm"]), case["B"], case["E"],
                         None if case["M"] is None else oracle.DenseOp(case["M"], True),
                         return_info=True, **case["opts"])
        assert torch.allclose(x, case["x"], rtol=1e-10, atol=1e-12), (case["tag"], case["method"])
        assert info["niter"] == case["oracle_niter"], (case["tag"], case["method"])
        # the reference's answer itself is a solution of the system to its tolerance
        err = ((case["x"] - case["x_exact"]).norm() / case["x_exact"].norm()).item()
        assert err <= 1e-4


def test_preconditioned_solvers_match_reference(golden):
    for case in golden["solve_precond"]:
        kw = {k: oracle.DenseOp(v, bool(torch.allclose(v, v.t()))) for k, v in case["precond"].items()}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, info = getattr(oracle, case["method"])(oracle.DenseOp(case["A"], case["herm"]), case["B"],
                                                      return_info=True, **case["opts"], **kw)
        assert torch.allclose(x, case["x"], rtol=1e-10, atol=1e-12), case["tag"]
        assert info["niter"] == case["oracle_niter"], case["tag"]


def test_exact_helpers():
    A = oracle.make_spd_c1(32)
    B = torch.ones(32, 2, dtype=torch.float64)
    E = torch.tensor([0.01, 0.02], dtype=torch.float64)
    x = oracle.exactsolve(A, B, E)
    assert torch.allclose(A @ x - x * E, B, atol=1e-10)
    ev, vec = oracle.exacteig(A, 3, "lowest")
    assert torch.allclose(A @ vec, vec * ev, atol=1e-10)


def test_problem_generators_are_deterministic():
    a1 = oracle.make_herm(64, 4)
    a2 = oracle.make_herm(64, 4)
    assert torch.equal(a1, a2) and torch.equal(a1, a1.t())
    ev = torch.linalg.eigvalsh(a1.double())
    assert (ev[:8] - (1 + torch.arange(8))).abs().max() < 0.05
    A, B = oracle.make_nonsym_c3(2, n=64)
    assert A.dtype == torch.bfloat16 and tuple(A.shape) == (2, 64, 64) and tuple(B.shape) == (2, 64, 1)
