"""Batched projected L-BFGS driver (host logic of the fit; reference: scipy L-BFGS-B via botorch,
scamlgp/utils.py:175,190)."""
import numpy as np
import scipy.optimize
import torch

from scamlgp_b200.lbfgs import lbfgs_minimize


def _rosen(x):
    a, b = x[:, :-1], x[:, 1:]
    f = (100.0 * (b - a * a) ** 2 + (1 - a) ** 2).sum(1)
    g = torch.zeros_like(x)
    g[:, :-1] += -400.0 * a * (b - a * a) - 2 * (1 - a)
    g[:, 1:] += 200.0 * (b - a * a)
    return f, g


def test_rosenbrock_rows_converge_independently():
    gen = torch.Generator().manual_seed(0)
    x0 = torch.randn(16, 5, dtype=torch.float64, generator=gen)
    calls = []

    def fun(x, active):
        calls.append(int(active.sum()))
        return _rosen(x)

    res = lbfgs_minimize(fun, x0, maxiter=500, gtol=1e-8, ftol=0.0)
    assert bool(res.converged.all()) and not bool(res.failed.any())
    assert float(_rosen(res.x)[1].abs().max()) < 1e-7  # stationary (5-D Rosenbrock has a second local minimum)
    assert int(((res.x - 1.0).abs().amax(1) < 1e-5).sum()) >= 14
    assert calls[-1] < 16  # rows drop out as they converge
    # a row's trajectory does not depend on its batch mates
    res1 = lbfgs_minimize(lambda x, a: _rosen(x), x0[3:4], maxiter=500, gtol=1e-8, ftol=0.0)
    assert torch.equal(res1.x[0], res.x[3]) and int(res1.iterations[0]) == int(res.iterations[3])


def test_lower_bounds_match_scipy_lbfgsb():
    gen = torch.Generator().manual_seed(1)
    A = torch.randn(4, 6, 6, dtype=torch.float64, generator=gen)
    Q = A @ A.transpose(1, 2) + 0.5 * torch.eye(6, dtype=torch.float64)
    b = torch.randn(4, 6, dtype=torch.float64, generator=gen)

    def fun(x, active):
        Qx = torch.einsum("eij,ej->ei", Q, x)
        return 0.5 * (x * Qx).sum(1) - (b * x).sum(1), Qx - b

    lower = torch.full((6,), 1e-10, dtype=torch.float64)
    res = lbfgs_minimize(fun, torch.full((4, 6), 0.3, dtype=torch.float64), lower=lower, maxiter=300, gtol=1e-9, ftol=0.0)
    assert bool(res.converged.all())
    for e in range(4):
        sp = scipy.optimize.minimize(lambda v: (0.5 * v @ Q[e].numpy() @ v - b[e].numpy() @ v, Q[e].numpy() @ v - b[e].numpy()),
                                     np.full(6, 0.3), jac=True, method="L-BFGS-B", bounds=[(1e-10, None)] * 6,
                                     options=dict(ftol=0, gtol=1e-10))
        assert np.abs(res.x[e].numpy() - sp.x).max() < 1e-6
        assert (res.x[e] >= 1e-10).all()


def test_nan_start_fails_row_and_nan_trial_backtracks():
    def fun(x, active):
        f = (x * x).sum(1)
        f = torch.where(x[:, 0] > 2.0, torch.full_like(f, float("nan")), f)  # "non-PSD" region
        return f, 2 * x

    x0 = torch.tensor([[3.0, 0.0], [1.5, 1.0], [-1.0, 4.0]], dtype=torch.float64)
    res = lbfgs_minimize(fun, x0, maxiter=100, gtol=1e-10, ftol=0.0)
    assert bool(res.failed[0]) and bool(torch.isnan(res.f[0]))
    assert bool(res.converged[1]) and bool(res.converged[2])
    assert float(res.x[1:].abs().max()) < 1e-8
