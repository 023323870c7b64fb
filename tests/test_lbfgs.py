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


# ---- device-side update kernel (csrc/scaml_lbfgs.cuh) vs the torch statement above ------------------------ #
import pytest  # noqa: E402

from scamlgp_b200.lbfgs import lbfgs_minimize_device  # noqa: E402


def _device_vs_torch(eng):
    dev = eng.device
    gen = torch.Generator().manual_seed(0)
    # (1) Rosenbrock, unconstrained, 16 rows x 5 dims
    x0 = torch.randn(16, 5, dtype=torch.float64, generator=gen)
    ref = lbfgs_minimize(lambda x, a: _rosen(x), x0, maxiter=500, gtol=1e-8, ftol=0.0)
    res = lbfgs_minimize_device(eng, lambda x, a: _rosen(x), x0.to(dev), maxiter=500, gtol=1e-8, ftol=0.0)
    assert bool(res.converged.all()) and not bool(res.failed.any())
    # same algorithm, reductions associated differently (warp tree vs torch.sum): same minimiser, similar effort
    assert float((res.x.cpu() - ref.x).abs().max()) < 1e-6
    assert abs(int(res.iterations.sum()) - int(ref.iterations.sum())) <= 0.2 * int(ref.iterations.sum())
    # a row's trajectory does not depend on its batch mates: bit-identical alone
    one = lbfgs_minimize_device(eng, lambda x, a: _rosen(x), x0[3:4].to(dev), maxiter=500, gtol=1e-8, ftol=0.0)
    assert torch.equal(one.x[0], res.x[3]) and int(one.iterations[0]) == int(res.iterations[3])
    # (2) bound-constrained convex quadratics, D = 70 (> one warp's width), vs scipy L-BFGS-B
    E, D = 3, 70
    A = torch.randn(E, D, D, dtype=torch.float64, generator=gen)
    Q = A @ A.transpose(1, 2) / D + 0.5 * torch.eye(D, dtype=torch.float64)
    b = torch.randn(E, D, dtype=torch.float64, generator=gen)
    Qd, bd = Q.to(dev), b.to(dev)

    def fun(x, active):
        Qx = torch.einsum("eij,ej->ei", Qd, x)
        return 0.5 * (x * Qx).sum(1) - (bd * x).sum(1), Qx - bd

    lower = torch.cat([torch.full((D - 8,), 1e-10), torch.full((8,), float("-inf"))]).to(torch.float64)
    res = lbfgs_minimize_device(eng, fun, torch.full((E, D), 0.3, dtype=torch.float64, device=dev), lower=lower.to(dev),
                                maxiter=400, gtol=1e-9, ftol=0.0)
    assert bool(res.converged.all())
    for e in range(E):
        Qe, be = Q[e].numpy(), b[e].numpy()
        sp = scipy.optimize.minimize(lambda v: (0.5 * v @ Qe @ v - be @ v, Qe @ v - be), np.full(D, 0.3), jac=True,
                                     method="L-BFGS-B", bounds=[(1e-10, None)] * (D - 8) + [(None, None)] * 8,
                                     options=dict(ftol=0, gtol=1e-10, maxiter=2000))
        assert np.abs(res.x[e].cpu().numpy() - sp.x).max() < 1e-6
        assert bool((res.x[e, : D - 8] >= 1e-10).all())
    # (2b) long rows (D >= 512): the 8-warp CTA-per-row variant of the kernel (target fit: D = M + d + 2); diagonal
    #      quadratic + bounds, against the torch statement of the algorithm and the closed-form minimiser
    E, D = 2, 700
    qd = (0.5 + torch.rand(E, D, dtype=torch.float64, generator=gen)).to(dev)
    cd = torch.randn(E, D, dtype=torch.float64, generator=gen).to(dev)
    lower = torch.cat([torch.full((D - 20,), 1e-10), torch.full((20,), float("-inf"))]).to(torch.float64).to(dev)

    def fun_w(x, active):
        return 0.5 * (qd * x * x).sum(1) - (cd * x).sum(1), qd * x - cd

    x0w = torch.full((E, D), 0.3, dtype=torch.float64, device=dev)
    res = lbfgs_minimize_device(eng, fun_w, x0w, lower=lower, maxiter=300, gtol=1e-9, ftol=0.0)
    exact = torch.maximum(cd / qd, lower)
    assert bool(res.converged.all())
    assert float((res.x - exact).abs().max()) < 1e-6
    ref = lbfgs_minimize(fun_w, x0w, lower=lower, maxiter=300, gtol=1e-9, ftol=0.0)  # torch statement, same device
    assert float((res.x - ref.x).abs().max()) < 1e-6
    assert abs(int(res.iterations.sum()) - int(ref.iterations.sum())) <= 0.25 * int(ref.iterations.sum()) + 2
    again = lbfgs_minimize_device(eng, fun_w, x0w, lower=lower, maxiter=300, gtol=1e-9, ftol=0.0)
    assert torch.equal(again.x, res.x)  # fixed-order reductions: bit-reproducible
    # (3) NaN at the start fails the row; NaN at a trial point backtracks
    def fun_nan(x, active):
        f = (x * x).sum(1)
        f = torch.where(x[:, 0] > 2.0, torch.full_like(f, float("nan")), f)
        return f, 2 * x

    x0 = torch.tensor([[3.0, 0.0], [1.5, 1.0], [-1.0, 4.0]], dtype=torch.float64, device=dev)
    res = lbfgs_minimize_device(eng, fun_nan, x0, maxiter=100, gtol=1e-10, ftol=0.0)
    assert bool(res.failed[0]) and bool(torch.isnan(res.f[0]))
    assert bool(res.converged[1]) and bool(res.converged[2])
    assert float(res.x[1:].abs().max()) < 1e-8


def test_device_kernel_matches_torch_statement_emulated(emu_lib):
    from tests.emu_engine import EmuEngine

    _device_vs_torch(EmuEngine(emu_lib))


@pytest.mark.gpu
def test_device_kernel_matches_torch_statement_gpu(engine):
    _device_vs_torch(engine)
