"""The oracle is test infrastructure: nothing under xitorch_b200/ (or include/) may import or reference it,
and the product has no CPU fallback for the Krylov methods."""
import os
import re

import pytest
import torch

import xitorch_b200 as xt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_oracle_or_reference():
    bad = []
    for base in ("xitorch_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h")):
                    continue
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M) or "sys.path.insert(0, \"/root/reference" in src:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


@pytest.mark.parametrize("method", ["cg", "bicgstab", "gmres"])
def test_krylov_solve_has_no_cpu_path(method):
    A = torch.eye(8, dtype=torch.float64) * 2
    with pytest.raises(RuntimeError, match="CUDA"):
        xt.linalg.solve(xt.LinearOperator.m(A, True), torch.ones(8, 1, dtype=torch.float64), method=method)


@pytest.mark.parametrize("method", ["davidson", "lanczos"])
def test_krylov_symeig_has_no_cpu_path(method):
    A = torch.diag(torch.arange(1, 33, dtype=torch.float64))
    with pytest.raises(RuntimeError, match="CUDA"):
        xt.linalg.symeig(xt.LinearOperator.m(A, True), neig=2, method=method)
