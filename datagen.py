"""Synthetic workload generator shared by bench.py, the tests and the examples.

Neutral ground: neither product code (the package never imports it) nor oracle (it evaluates no GP arithmetic) --
it only draws inputs: Hartmann-6-family meta-tasks (alpha ranges of the reference's benchmark,
scamlgp/benchmarking/benchmarks/hartmann_3d.py:31-34, closed form benchmarking/functions/hartmann.py:170-185), a
separable smooth family for d != 6, and raw hyper-parameter rows the way `optimize_marginal_likelihood` visits them
(row 0 = initial values, rows 1.. = prior draws, scamlgp/utils.py:173-203; priors / bounds scamlgp/model.py:25-70).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

DT = torch.float64
PRIOR_GAMMA, PRIOR_LOGNORMAL = 1, 2

_H6_A = [[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]]
_H6_P = [[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
         [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]]


def hartmann6(X: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    A = torch.tensor(_H6_A, dtype=DT)
    P = 1e-4 * torch.tensor(_H6_P, dtype=DT)
    e = torch.exp(-(A[None] * (X[:, None, :] - P[None]) ** 2).sum(-1))  # n x 4
    return -(e * alpha[None]).sum(-1)


def synthetic_tasks(M: int, n: int, d: int, seed: int = 0, noise_sd: float = 0.1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Hartmann-6 family for d == 6, separable smooth family otherwise.  Returns X[M,n,d], Y[M,n] (host, fp64)."""
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(M, n, d, dtype=DT, generator=g)
    if d == 6:
        lo = torch.tensor([1.0, 1.18, 2.8, 3.2], dtype=DT)
        hi = torch.tensor([1.02, 1.2, 3.0, 3.4], dtype=DT)
        al = lo + (hi - lo) * torch.rand(M, 4, dtype=DT, generator=g)
        Y = torch.stack([hartmann6(X[i], al[i]) for i in range(M)])
    else:
        a = 1.0 + torch.rand(M, 1, d, dtype=DT, generator=g)
        ph = torch.rand(M, 1, d, dtype=DT, generator=g)
        Y = (torch.sin(3.0 * a * X + 6.28 * ph) + (X - ph) ** 2).sum(-1) / math.sqrt(d)
    Y = Y + noise_sd * torch.randn(M, n, dtype=DT, generator=g)
    return X, Y


def _logit(v: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    u = (v - lo) / (hi - lo)
    return torch.log(u) - torch.log1p(-u)


def initial_theta_raw(d: int, spec) -> torch.Tensor:
    """Raw (pre-sigmoid) image of the reference's initial values (model.py:31,55,67); `spec` is any object with the
    HyperSpec fields (ls_init / os_init / noise_init and the *_bounds)."""
    ls = _logit(torch.full((d,), float(spec.ls_init), dtype=DT), *spec.ls_bounds)
    os_ = _logit(torch.tensor([float(spec.os_init)], dtype=DT), *spec.os_bounds)
    nz = _logit(torch.tensor([float(spec.noise_init)], dtype=DT), *spec.noise_bounds)
    return torch.cat([ls, os_, nz])


def sample_theta_raw(M: int, R: int, d: int, spec, seed: int = 0) -> torch.Tensor:
    """[M, R, d+2] raw hyper-parameter rows: row 0 = initial values; rows 1.. = prior samples clipped into the
    Interval (1 warm start + (R-1) prior restarts)."""
    g = torch.Generator().manual_seed(seed + 12345)
    out = torch.empty(M, R, d + 2, dtype=DT)
    out[:, 0] = initial_theta_raw(d, spec)

    def draw(prior, shape, bounds):
        kind, p1, p2 = prior
        if kind == PRIOR_GAMMA:
            # Gamma(k, rate) as a sum of k exponentials (all reference Gamma priors have integer k): generator-exact
            k = int(p1)
            assert float(k) == p1
            u = torch.rand(*shape, k, dtype=DT, generator=g)
            v = -torch.log(u).sum(-1) / p2
        else:
            v = torch.exp(p1 + p2 * torch.randn(*shape, dtype=DT, generator=g))
        lo, hi = bounds
        eps = 1e-6 * (hi - lo)
        return v.clamp(lo + eps, hi - eps)

    if R > 1:
        ls = draw(spec.ls_prior, (M, R - 1, d), spec.ls_bounds)
        os_ = draw(spec.os_prior, (M, R - 1), spec.os_bounds)
        nz = draw(spec.noise_prior, (M, R - 1), spec.noise_bounds)
        out[:, 1:, :d] = _logit(ls, *spec.ls_bounds)
        out[:, 1:, d] = _logit(os_, *spec.os_bounds)
        out[:, 1:, d + 1] = _logit(nz, *spec.noise_bounds)
    return out
