"""GPU replay of tests/golden/generalized_golden.pt (outputs of the unmodified reference on generalized problems and on
start blocks wider than neig) through the CUDA path.  Collected last on purpose: it was added after the round's GPU
minutes were spent, the same code path is covered on hardware by tests/test_gpu_symeig.py
(`test_generalized_problem*`, `test_wide_start_block_and_preconditioner`) and on the CPU by
tests/test_davidson_generalized_host.py."""
import pytest
import torch

import xitorch_b200 as xt

pytestmark = pytest.mark.gpu
DEV = "cuda"
EIG_RTOL = 1e-5


def test_generalized_and_wide_start_fixture(golden_generalized):
    """outputs of the unmodified reference on generalized problems (both modes) and start blocks wider than neig
    (tests/golden/generalized_golden.pt) through the CUDA path: block-matvec kernel for A and M, one-CTA eigensolver"""
    for c in golden_generalized:
        if c["A"].dim() != 2 or c["A"].shape[-1] < 32:
            continue                                   # batched / full-space cases: covered by the host tests
        A, Mm = c["A"], c["M"]
        info = {}
        evals, evecs = xt.linalg.symeig(xt.LinearOperator.m(A.to(DEV), True), neig=c["neig"], mode=c["mode"],
                                        M=None if Mm is None else xt.LinearOperator.m(Mm.to(DEV), True),
                                        method="davidson", nguess=c["nguess"], min_eps=c["min_eps"], info=info)
        assert info["converged"] and info["engine"] == "host-composed", (c["tag"], info)
        assert ((evals.cpu() - c["evals"]).abs() / c["evals"].abs()).max().item() <= EIG_RTOL, c["tag"]
        X = evecs.cpu()
        Md = torch.eye(A.shape[-1], dtype=torch.float64) if Mm is None else Mm
        assert (A @ X - Md @ X * evals.cpu().unsqueeze(-2)).abs().max().item() <= 1e-6, c["tag"]
