"""
Case definitions shared by `oracle/gen_golden_optimize.py` (run against the reference) and
`tests/test_optimize_cpu.py` (run against this package): every builder takes the package (`xt`) it should use, so the
very same problem is posed to both.  All cases are tiny (CPU, fp64, exact linear algebra in backward), they pin the
plumbing around the hot path -- object methods as `fcn`, second derivatives, the non-Broyden methods.
"""
import torch

DT = torch.float64


# ------------------------------------------------------------------------------------------------ pure functions
def rf_fcn(y, A):
    return torch.tanh(A @ y + 0.1) + y / 2.0


def min_fcn(y, A):                       # the doctest function of optimize/rootfinder.py:232-239
    return torch.sum((A @ y) ** 2 + y / 2.0)


def make_inputs(kind):
    g = torch.Generator().manual_seed(7)
    n = 6
    A = 0.3 * torch.randn(n, n, generator=g, dtype=DT) / n ** 0.5
    if kind == "minimize":
        A = A + torch.eye(n, dtype=DT)
    if kind == "equilibrium":            # contraction: y = tanh(A y + 0.1) + y / 2
        A = A - 0.2 * torch.eye(n, dtype=DT)
    return A, torch.zeros(n, 1, dtype=DT)


METHOD_CASES = [
    ("rootfinder", "newton", {}),
    ("rootfinder", "broyden2", {}),
    ("equilibrium", "anderson_acc", {"f_tol": 1e-10, "x_tol": 1e-10, "feat_ndims": 2}),
    ("equilibrium", "broyden1", {}),
    ("equilibrium", "linearmixing", {"alpha": -0.7}),
    ("minimize", "broyden1", {}),
    ("minimize", "newton", {}),
    ("minimize", "gd", {"step": 0.05, "maxiter": 4000, "f_rtol": 1e-14, "x_rtol": 1e-12}),
    ("minimize", "adam", {"step": 0.02, "maxiter": 4000, "f_rtol": 1e-14, "x_rtol": 1e-12}),
]


# ------------------------------------------------------------------------------------------------ object methods
def _activation(kind, z):
    if kind == "minimize":
        return (z - 0.1) ** 2
    return torch.cos(z) if kind == "equilibrium" else torch.sigmoid(z)


def make_module(xt, base, kind):
    """an object whose `forward(x)` depends on tensors held as attributes; returns (factory, leaf tensors)"""
    g = torch.Generator().manual_seed(11)
    nb, nr = 2, 3
    # nn.Parameter leaves: assigning them registers them with nn.Module and keeps them plain attributes elsewhere
    A = torch.nn.Parameter(0.5 * torch.randn(nr, nr, generator=g, dtype=DT))
    diag = torch.nn.Parameter(torch.randn(nb, nr, generator=g, dtype=DT))
    bias = torch.nn.Parameter(0.1 * torch.randn(nb, nr, generator=g, dtype=DT))

    def forward_impl(self, x):
        M = self.A.unsqueeze(0) + torch.diag_embed(self.diag)
        y = torch.bmm(M.expand(x.shape[0], -1, -1), x.unsqueeze(-1)).squeeze(-1)
        out = _activation(kind, 2 * y) + 2 * self.bias
        if kind == "rootfinder":
            out = out + x
        return out.sum() if kind == "minimize" else out

    class Editable(xt.EditableModule):
        def __init__(self, A, diag, bias):
            self.A, self.diag, self.bias = A, diag, bias

        forward = forward_impl

        def getparamnames(self, methodname, prefix=""):
            return [prefix + "A", prefix + "diag", prefix + "bias"]

    class NN(torch.nn.Module):
        def __init__(self, A, diag, bias):
            super().__init__()
            self.A, self.diag, self.bias = A, diag, bias

        forward = forward_impl

    class NNInEditable(xt.EditableModule):
        def __init__(self, A, diag, bias):
            self.module = NN(A, diag, bias)

        def forward(self, x):
            return self.module.forward(x)

        def getparamnames(self, methodname, prefix=""):
            return [nm for nm, _ in self.module.named_parameters(prefix=prefix + "module")]

    factory = {"editable": Editable, "nn": NN, "nn_in_editable": NNInEditable}[base]
    return factory, (A, diag, bias)


def module_loss(xt, solver, base, kind, A, diag, bias):
    factory, _ = make_module(xt, base, kind)
    model = factory(A, diag, bias)
    g = torch.Generator().manual_seed(13)
    y0 = torch.randn(2, 3, generator=g, dtype=DT)
    opts = {"rootfinder": dict(f_tol=1e-12, alpha=-0.5), "equilibrium": dict(f_tol=1e-12, alpha=-0.5),
            "minimize": dict(f_tol=1e-12, alpha=-0.5)}[kind]
    y = solver(model.forward, y0, method="broyden1", bck_options={"method": "exactsolve"}, **opts)
    w = torch.linspace(0.5, 1.5, y.numel(), dtype=DT).reshape(y.shape)
    return (w * y ** 2).sum()


def first_and_second(loss, tensors):
    """gradients of the loss and of |gradient|^2 (a scalar second-derivative probe); unused tensors give zeros"""
    def dense(gs):
        return [torch.zeros_like(t) if g is None else g for g, t in zip(gs, tensors)]
    grads = dense(torch.autograd.grad(loss, tensors, create_graph=True, allow_unused=True))
    gsum = sum((g ** 2).sum() for g in grads)
    grads2 = dense(torch.autograd.grad(gsum, tensors, allow_unused=True))
    return [g.detach() for g in grads], grads2


# ------------------------------------------------------------------------------------------------ degenerate spectrum
def degenerate_loss(xt, symeig, a, mat, P2):
    """loss that does not depend on the basis chosen inside the degenerate eigenspaces (the requirement under which
    the eigenvector derivative exists; reference test_symeig_A_degenerate, _tests/test_linop_fcns.py:178-235)"""
    P, _ = torch.linalg.qr(mat)
    vals = torch.cat((a[:2], a[1:2], a[2:], a[2:]))              # spectrum a0, a1, a1, a2, a2
    A = (P * vals.unsqueeze(0)) @ P.T
    A = (A + A.T) * 0.5
    evals, evecs = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=3, method="exacteig",
                          bck_options={"method": "exactsolve"})
    sub = evecs[:, 1:3]                                          # the degenerate pair: only its projector is used
    proj = sub @ sub.T
    return (evals ** 2).sum() + ((P2 @ proj) * P2).sum() + (evecs[:, 0] @ P2 @ evecs[:, 0])


def _degenerate_inputs(offset):
    g = torch.Generator().manual_seed(17)
    n = 5
    mat = torch.randn(n, n, generator=g, dtype=DT).requires_grad_()
    P2 = torch.randn(n, n, generator=g, dtype=DT).requires_grad_()
    a = (torch.tensor([1.0, 2.0, 3.0], dtype=DT) + offset).requires_grad_()
    return a, mat, P2


degenerate_loss.inputs = _degenerate_inputs
