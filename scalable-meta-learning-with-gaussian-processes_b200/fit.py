"""Batched multi-restart marginal-likelihood fits (rows a1, a2, a4 of SURVEY 8a).

The reference fits M source GPs one after another, each with 1 warm start + `num_restarts`
prior-sampled restarts of scipy L-BFGS-B (scamlgp/model.py:176-188, utils.py:139-212).  Here
all M x (1 + num_restarts) optimisations advance in lock-step on the GPU: one fused
`scaml_lml_grad` launch plus one `scaml_lbfgs_step` launch (device-side update, one warp per row) per L-BFGS
round over the rows that are still active.

Semantics kept (scamlgp/utils.py:139-212):
  * row 0 of every task is the warm start (the parameters the caller's modules hold), rows 1..R-1
    are drawn from the priors, a draw being rejected (<= `num_retries` times) when the
    constraint's inverse transform is not finite (utils.py:47-69);
  * a row whose objective is not finite (Cholesky failed even after the jitter ladder) is
    skipped -- the reference catches the RuntimeError and continues (utils.py:180-198);
  * the best row by final (LML + log priors)/n wins, earlier rows winning ties (utils.py:200-203);
  * all rows failed -> ModelFittingError (utils.py:207-212).
Dropped on purpose: the reference warm-starts task k from task k-1's optimum because it
deep-copies the previously fitted modules (model.py:177-178) -- a sequential dependency that is
an artefact of the loop, not part of the model.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from ._capi import PRIOR_GAMMA, PRIOR_LOGNORMAL, PRIOR_NONE, HyperSpec
from .engine import Engine, SourceBatch
from .lbfgs import LbfgsResult, lbfgs_minimize_device
from .modules import ModelFittingError

DT = torch.float64
DEFAULT_FIT_OPTIONS = dict(maxiter=200, gtol=1e-5, ftol=2.2e-9, history=10)


def _draw(prior: Tuple[int, float, float], shape, generator) -> Optional[torch.Tensor]:
    kind, p1, p2 = prior
    if kind == PRIOR_GAMMA:
        conc = torch.full(tuple(shape), float(p1), dtype=DT)
        return torch._standard_gamma(conc, generator=generator) / float(p2)
    if kind == PRIOR_LOGNORMAL:
        return torch.exp(float(p1) + float(p2) * torch.randn(tuple(shape), dtype=DT, generator=generator))
    return None


def _logit(v: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    u = (v - lo) / (hi - lo)
    return torch.log(u) - torch.log1p(-u)


def sample_prior_column(prior, bounds, shape, fallback_raw: torch.Tensor, generator, num_retries: int = 5,
                        name: str = "prior") -> torch.Tensor:
    """Raw (pre-sigmoid) samples of one parameter block, `sample_all_priors` semantics.

    Entries whose inverse transform is not finite are redrawn up to `num_retries` times; if any is
    still invalid the reference raises RuntimeError (utils.py:63-67) and so does this."""
    v = _draw(prior, shape, generator)
    if v is None:  # no prior registered: the parameter keeps its current value (named_priors skips it)
        return fallback_raw.expand(tuple(shape)).clone()
    lo, hi = bounds
    raw = _logit(v, lo, hi)
    for _ in range(num_retries):
        bad = ~torch.isfinite(raw)
        if not bool(bad.any()):
            break
        redraw = _draw(prior, shape, generator)
        raw = torch.where(bad, _logit(redraw, lo, hi), raw)
    if not bool(torch.isfinite(raw).all()):
        raise RuntimeError(f"Sampling of {name} failed {num_retries} times. Please check the compatibility between "
                           "prior support and the constraint.")
    return raw


def sample_theta_raw(spec: HyperSpec, theta0: torch.Tensor, M: int, S: int, generator=None) -> torch.Tensor:
    """[M, S, P] prior draws of (lengthscales, outputscale, noise) in raw space (host tensors)."""
    d = theta0.numel() - 2
    ls = sample_prior_column(spec.ls_prior, spec.ls_bounds, (M, S, d), theta0[:d], generator, name="lengthscale_prior")
    os_ = sample_prior_column(spec.os_prior, spec.os_bounds, (M, S, 1), theta0[d:d + 1], generator,
                              name="outputscale_prior")
    nz = sample_prior_column(spec.noise_prior, spec.noise_bounds, (M, S, 1), theta0[d + 1:], generator,
                             name="noise_prior")
    return torch.cat([ls, os_, nz], dim=-1)


@dataclass
class SourceFit:
    theta_raw: torch.Tensor  # [M, P] best row per task (device)
    lml: torch.Tensor  # [M] (LML + log priors)/n at the optimum
    best_row: torch.Tensor  # [M] index of the winning restart
    all_theta_raw: torch.Tensor  # [M, R, P]
    all_lml: torch.Tensor  # [M, R] (-inf for failed rows)
    result: LbfgsResult


def fit_sources(engine: Engine, batch: SourceBatch, spec: HyperSpec, theta_init: torch.Tensor,
                fit_options: Optional[dict] = None) -> SourceFit:
    """Maximise (LML + log priors)/n of every (task, row) in lock-step.  theta_init [M, R, P] on any device."""
    opts = dict(DEFAULT_FIT_OPTIONS)
    opts.update(fit_options or {})
    M, R, P = theta_init.shape
    dev = batch.X.device
    x0 = theta_init.to(dev, DT).reshape(M * R, P).contiguous()

    def fun(x, active):
        skip = (~active).to(torch.int32).reshape(M, R).contiguous()
        lml, grad, _ = engine.lml_grad(batch, x.reshape(M, R, P).contiguous(), spec, skip=skip)
        return -lml.reshape(-1), -grad.reshape(M * R, P)

    res = lbfgs_minimize_device(engine, fun, x0, **opts)
    lml = torch.where(res.failed | ~torch.isfinite(res.f), torch.full_like(res.f, float("-inf")), -res.f).reshape(M, R)
    best = torch.argmax(lml, dim=1)  # first maximum wins ties: the warm start is row 0
    best_lml = lml.gather(1, best.unsqueeze(1)).squeeze(1)
    if bool(torch.isinf(best_lml).any()):
        bad = torch.nonzero(torch.isinf(best_lml)).flatten().tolist()
        raise ModelFittingError("Hyperparameter optimization failed for all attempts. Usually this indicates a "
                                f"problem with model's input data or hyperparameter priors definitions. (tasks {bad})")
    xs = res.x.reshape(M, R, P)
    theta = xs.gather(1, best.reshape(M, 1, 1).expand(M, 1, P)).squeeze(1).contiguous()
    return SourceFit(theta, best_lml, best, xs, lml, res)


@dataclass
class TargetFit:
    weights: torch.Tensor  # [M]
    theta_raw: torch.Tensor  # [P]
    lml: float
    best_row: int
    all_lml: torch.Tensor  # [R]
    result: LbfgsResult


def fit_target(engine: Engine, source_means: torch.Tensor, source_covs: torch.Tensor, Xt: torch.Tensor,
               yt: torch.Tensor, mu_all: float, s_all: float, spec: HyperSpec, w_init: torch.Tensor,
               theta_init: torch.Tensor, w_prior=(PRIOR_GAMMA, 1.0, 1.0), w_lower: float = 1e-10,
               fit_options: Optional[dict] = None) -> TargetFit:
    """Maximise the ScaML-GP target objective over [weights (M) | raw kernel params (P)] for R rows.

    w_init [R, M], theta_init [R, P].  The weights are box-bounded below (GreaterThan(1e-10, transform=None),
    reference model.py:333-337); the kernel parameters are free in raw space."""
    opts = dict(DEFAULT_FIT_OPTIONS)
    opts.update(fit_options or {})
    dev = Xt.device
    R, M = w_init.shape
    P = theta_init.shape[1]
    x0 = torch.cat([w_init.to(dev, DT), theta_init.to(dev, DT)], dim=1).contiguous()
    lower = torch.cat([torch.full((M,), w_lower, dtype=DT), torch.full((P,), float("-inf"), dtype=DT)]).to(dev)

    def fun(x, active):
        # all R rows are evaluated every round (R is small): no data-dependent shapes, no host sync; the update
        # kernel ignores the rows that are no longer active
        lml, gw, gt, _ = engine.target_lml_grad_safe(source_means, source_covs, Xt, yt, x[:, :M].contiguous(),
                                                     x[:, M:].contiguous(), mu_all, s_all, spec, w_prior)
        return -lml, -torch.cat([gw, gt], dim=1)

    res = lbfgs_minimize_device(engine, fun, x0, lower=lower, **opts)
    lml = torch.where(res.failed | ~torch.isfinite(res.f), torch.full_like(res.f, float("-inf")), -res.f)
    best = int(torch.argmax(lml))
    if not bool(torch.isfinite(lml[best])):
        raise ModelFittingError("Hyperparameter optimization failed for all attempts. Usually this indicates a "
                                "problem with model's input data or hyperparameter priors definitions.")
    return TargetFit(res.x[best, :M].contiguous(), res.x[best, M:].contiguous(), float(lml[best]), best, lml, res)
