"""CPU tests of the parameter plumbing (layer L1): `EditableModule`, `get_pure_function` / `make_sibling`, and the
`install_as_xitorch()` alias.  Contract = /root/reference/xitorch/_core/{editable_module,pure_function}.py and their
tests; the cases here are our own (the reference's test files themselves also pass against this package through the
alias)."""
import subprocess
import sys
import os
import warnings

import pytest
import torch

import xitorch_b200 as xt
from xitorch_b200 import EditableModule, get_pure_function, make_sibling
from xitorch_b200._utils import GetSetParamsError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = torch.float64


class Lens(EditableModule):
    """tensors at attribute paths of every supported shape: plain, list item, dict item, nested object"""

    def __init__(self):
        self.focal = torch.tensor([2.0, 3.0], dtype=DT)
        self.coeffs = [torch.tensor([0.5], dtype=DT), torch.tensor([0.25], dtype=DT)]
        self.table = {"shift": torch.tensor([1.0], dtype=DT)}
        self.alias = self.focal                     # the same tensor under a second name
        self.count = torch.tensor([3])              # integer tensor: never a parameter

    def power(self, x):
        return x / self.focal + self.coeffs[1] * x ** 2 + self.table["shift"] + self.alias

    def leaky(self, x):
        self.table["shift"] = self.table["shift"] + 1.0     # changes the state: not pure
        return x * self.focal

    def getparamnames(self, methodname, prefix=""):
        if methodname == "power":
            return [prefix + "focal", prefix + "coeffs[1]", prefix + "table[\"shift\"]", prefix + "alias"]
        if methodname == "leaky":
            return [prefix + "focal", prefix + "table[\"shift\"]"]
        if methodname == "short":
            return [prefix + "focal"]
        if methodname == "long":
            return [prefix + "focal", prefix + "coeffs[0]", prefix + "coeffs[1]", prefix + "table[\"shift\"]",
                    prefix + "alias"]
        raise KeyError(methodname)

    def short(self, x):
        return self.power(x)

    def long(self, x):
        return self.power(x)


def test_get_and_set_params_by_path():
    m = Lens()
    ps = m.getparams("power")
    assert [p is q for p, q in zip(ps, (m.focal, m.coeffs[1], m.table["shift"], m.alias))] == [True] * 4
    new = [p + 1 for p in ps]
    assert m.setparams("power", *new) == 4
    assert m.focal is new[0] and m.coeffs[1] is new[1] and m.table["shift"] is new[2] and m.alias is new[3]
    with pytest.raises(KeyError):
        m.getparams("nothing")


def test_unique_params_deduplicate_aliases():
    m = Lens()
    uniq = m.getuniqueparams("power")
    assert len(uniq) == 3 and uniq[0] is m.focal
    fresh = [torch.zeros_like(u) for u in uniq]
    m.setuniqueparams("power", *fresh)
    assert m.focal is fresh[0] and m.alias is fresh[0]          # both names receive the one new tensor
    assert m.coeffs[1] is fresh[1] and m.table["shift"] is fresh[2]
    # leaves only
    m2 = Lens()
    m2.focal = m2.alias = (torch.ones(2, dtype=DT).requires_grad_() * 2)      # non-leaf
    m2.cached_getparamnames("power", refresh=True)
    assert all(t.is_leaf for t in m2.getuniqueparams("power", onlyleaves=True))
    assert len(m2.getuniqueparams("power", onlyleaves=True)) == 2


def test_assertparams_diagnoses_name_lists(capsys):
    m = Lens()
    x = torch.tensor([1.0, 2.0], dtype=DT)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        m.assertparams(m.power, x)                               # correct list: silent
    assert '"power" method check done' in capsys.readouterr().out
    with pytest.warns(UserWarning, match="does not include"):
        m.assertparams(m.short, x)
    with pytest.warns(UserWarning, match="excess"):
        m.assertparams(m.long, x)
    with pytest.raises(GetSetParamsError):
        m.assertparams(m.leaky, x)
    with pytest.raises(TypeError):
        m.assertparams(lambda x: x, x)
    with pytest.raises(RuntimeError):
        m.assertparams(Lens().power, x)


# ---------------------------------------------------------------------------------------------- pure functions
def test_pure_function_of_plain_function_has_no_state():
    def f(a, b):
        return a * b
    pf = get_pure_function(f)
    assert pf(2, 3) == 6 and pf.objparams() == []
    assert get_pure_function(pf) is pf
    with pf.useobjparams([]):
        assert pf(2, 5) == 10
    with pytest.raises(RuntimeError):
        get_pure_function(3)


def test_method_state_substitution_nests_and_restores():
    m = Lens()
    x = torch.tensor([1.0, 2.0], dtype=DT)
    pf = get_pure_function(m.power)
    base = pf(x)
    orig = list(pf.objparams())
    assert len(orig) == 3                                        # de-duplicated
    one = [p + 1 for p in orig]
    two = [p + 2 for p in orig]
    with pf.useobjparams(one):
        assert m.focal is one[0] and m.alias is one[0]
        v1 = pf(x)
        with pf.useobjparams(two):
            assert m.focal is two[0]
        assert m.focal is one[0]
    assert m.focal is orig[0] and m.coeffs[1] is orig[1]
    assert torch.equal(pf(x), base) and not torch.equal(v1, base)
    with pytest.raises(RuntimeError):
        with pf.useobjparams(one[:2]):
            pass


def test_nn_module_parameters_are_state_and_stay_registered():
    lin = torch.nn.Linear(3, 2).to(DT)
    pf = get_pure_function(lin.forward)
    x = torch.ones(1, 3, dtype=DT)
    names = [nm for nm, _ in lin.named_parameters()]
    w2 = [torch.zeros_like(p) for p in pf.objparams()]
    with pf.useobjparams(w2):
        assert torch.equal(pf(x), torch.zeros(1, 2, dtype=DT))
    assert [nm for nm, _ in lin.named_parameters()] == names     # re-registered as parameters
    assert isinstance(lin.weight, torch.nn.Parameter)
    # a callable object works like its __call__ / forward
    assert len(get_pure_function(lin).objparams()) == 2


def test_sibling_shares_state_of_all_parents():
    m1, m2 = Lens(), Lens()
    x = torch.tensor([1.0, 2.0], dtype=DT)

    @make_sibling(m1.power, m2.power)
    def both(x):
        return m1.power(x) - 2 * m2.power(x)

    assert len(both.objparams()) == 6
    zeros = [torch.ones_like(p) for p in both.objparams()]
    with both.useobjparams(zeros):
        assert m1.focal is zeros[0] and m2.focal is zeros[3]
        inside = both(x)
    assert torch.allclose(inside, -(x + x ** 2 + 2))
    assert not torch.allclose(both(x), inside)
    with pytest.raises(TypeError):
        make_sibling()


def test_gradients_reach_object_state_through_jac():
    # the autograd boundary use: objparams travel as explicit inputs of the operator
    m = Lens()
    m.focal = m.alias = torch.tensor([2.0, 3.0], dtype=DT, requires_grad=True)
    x = torch.tensor([1.0, 2.0], dtype=DT, requires_grad=True)
    J = xt.grad.jac(m.power, (x,), idxs=0)
    dense = J.fullmatrix()
    expect = torch.diag(1 / m.focal + 2 * m.coeffs[1] * x)
    assert torch.allclose(dense, expect)
    (g,) = torch.autograd.grad(dense.sum(), m.focal)
    assert torch.allclose(g, -1 / m.focal.detach() ** 2)


# ---------------------------------------------------------------------------------------------- the alias
def test_install_as_xitorch_in_a_clean_interpreter():
    code = r"""
import sys
sys.path.insert(0, %r)
import xitorch_b200
xitorch_b200.install_as_xitorch()
import xitorch
from xitorch import LinearOperator, EditableModule
from xitorch.linalg import symeig, solve, svd, lsymeig, usymeig
from xitorch.linalg.symeig import symeig as s2
from xitorch.optimize import rootfinder, equilibrium, minimize
from xitorch.grad.jachess import jac, hess
from xitorch.grad import jac as j2
from xitorch._core.editable_module import EditableModule as E2
from xitorch._core.pure_function import get_pure_function, make_sibling
from xitorch._utils.exceptions import ConvergenceWarning, MathWarning, GetSetParamsError
from xitorch._utils.bcast import get_bcasted_dims, normalize_bcast_dims
from xitorch._utils.misc import get_method, set_default_option
from xitorch.debug.modes import is_debug_enabled, enable_debug
assert xitorch.__xitorch_b200__ and s2 is symeig and E2 is EditableModule and j2 is jac
assert xitorch.linalg.symeig is symeig or callable(xitorch.linalg.symeig)
import torch
A = torch.tensor([[2.0, 1.0], [1.0, 3.0]], dtype=torch.float64)
ev, _ = symeig(LinearOperator.m(A, is_hermitian=True), method="exacteig")
assert torch.allclose(ev, torch.linalg.eigvalsh(A))
xitorch_b200.install_as_xitorch()          # idempotent
print("alias-ok")
""" % ROOT
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "alias-ok" in out.stdout


def test_install_refuses_to_shadow_another_xitorch():
    code = r"""
import sys, types
sys.modules["xitorch"] = types.ModuleType("xitorch")
sys.path.insert(0, %r)
import xitorch_b200
try:
    xitorch_b200.install_as_xitorch()
except RuntimeError:
    xitorch_b200.install_as_xitorch(force=True)
    import xitorch
    assert xitorch.__xitorch_b200__
    print("refused-then-forced")
""" % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "refused-then-forced" in out.stdout
