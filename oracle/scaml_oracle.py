"""CPU fp64 ORACLE for the ScaML-GP hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain torch-fp64 *restatement* of the arithmetic the reference
(`/root/reference/scamlgp/model.py`, `scamlgp/utils.py`) delegates to
botorch 0.7.3 / gpytorch 1.9.0 / linear-operator 0.2.0 (poetry.lock:135,703,1020).
Those packages are not installable in this image (no network, no wheels), so the
restatement follows their published algorithms and is anchored on the reference's
own call sites.

PARITY STATUS: **parity unpinned at the third-party boundary** -- the reference's
tests hold no numerical golden values for the GP path (tests/optimizer_test.py and
scamlgp/testing.py assert behaviour/determinism only).  The oracle is instead
pinned against independent implementations available here (tests/test_oracle.py):
  * sklearn GaussianProcessRegressor (LML, d LML / d log-theta, mean, std),
  * torch autograd vs the analytic trace-formula gradient,
  * mpmath 50-digit arithmetic at small n.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module.  The product path
(`scamlgp_b200`) never does; it fails loudly when the CUDA library is missing.

Conventions
  theta_raw = [raw_lengthscale_1..d, raw_outputscale, raw_noise]  (P = d + 2)
  theta     = lo + (hi - lo) * sigmoid(raw)            gpytorch `Interval`
  K_y       = s * kappa(r^2) + noise * I ,  r^2_ab = sum_j ((x_aj - x_bj)/l_j)^2
  objective = ( log N(y~ | 0, K_y) + sum log p(theta) ) / n     (gpytorch
              ExactMarginalLogLikelihood: priors added, then divided by n)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

DT = torch.float64

KERNEL_RBF = 0
KERNEL_MATERN12 = 1
KERNEL_MATERN32 = 2
KERNEL_MATERN52 = 3

PRIOR_NONE = 0
PRIOR_GAMMA = 1  # (concentration, rate)
PRIOR_LOGNORMAL = 2  # (mu, sigma)

JITTER_LADDER = (0.0, 1e-8, 1e-7, 1e-6)  # linear_operator psd_safe_cholesky, fp64


@dataclass
class HyperSpec:
    """Kernel family, Interval constraints and priors of one GP.

    Defaults = the source-GP settings, reference model.py:25-70.
    """

    kernel: int = KERNEL_RBF
    ls_bounds: Tuple[float, float] = (1e-4, 1e2)  # model.py:52-56
    os_bounds: Tuple[float, float] = (1e-4, 1e2)  # model.py:64-68
    noise_bounds: Tuple[float, float] = (1e-8, 1e-2)  # model.py:31
    ls_prior: Tuple[int, float, float] = (PRIOR_GAMMA, 3.0, 6.0)  # model.py:41
    os_prior: Tuple[int, float, float] = (PRIOR_GAMMA, 2.0, 0.15)  # model.py:42
    noise_prior: Tuple[int, float, float] = (PRIOR_LOGNORMAL, -8.0, 2.0)  # model.py:28
    ls_init: float = 0.5
    os_init: float = 1.0
    noise_init: float = 1e-3

    @staticmethod
    def source(kernel: int = KERNEL_RBF) -> "HyperSpec":
        return HyperSpec(kernel=kernel)

    @staticmethod
    def target(kernel: int = KERNEL_RBF) -> "HyperSpec":
        """Target-GP defaults, reference model.py:73-105."""
        return HyperSpec(
            kernel=kernel,
            ls_prior=(PRIOR_LOGNORMAL, 0.5, 1.5),
            os_prior=(PRIOR_LOGNORMAL, -2.0, 3.0),
            ls_init=1.0,
            os_init=0.1,
        )


# --------------------------------------------------------------------------- #
# constraints / priors
# --------------------------------------------------------------------------- #
def constrain(raw: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    """gpytorch Interval.transform: lo + (hi-lo)*sigmoid(raw)."""
    return lo + (hi - lo) * torch.sigmoid(raw)


def unconstrain(val: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    """gpytorch Interval.inverse_transform: logit((v-lo)/(hi-lo))."""
    p = (torch.as_tensor(val, dtype=DT) - lo) / (hi - lo)
    return torch.log(p) - torch.log1p(-p)


def split_theta(theta_raw: torch.Tensor, spec: HyperSpec):
    d = theta_raw.shape[-1] - 2
    ls = constrain(theta_raw[..., :d], *spec.ls_bounds)
    os_ = constrain(theta_raw[..., d], *spec.os_bounds)
    noise = constrain(theta_raw[..., d + 1], *spec.noise_bounds)
    return ls, os_, noise


def pack_theta(ls, os_, noise, spec: HyperSpec) -> torch.Tensor:
    ls = torch.as_tensor(ls, dtype=DT).reshape(-1)
    return torch.cat(
        [
            unconstrain(ls, *spec.ls_bounds),
            unconstrain(torch.as_tensor(os_, dtype=DT).reshape(1), *spec.os_bounds),
            unconstrain(torch.as_tensor(noise, dtype=DT).reshape(1), *spec.noise_bounds),
        ]
    )


def initial_theta_raw(d: int, spec: HyperSpec) -> torch.Tensor:
    return pack_theta(torch.full((d,), spec.ls_init, dtype=DT), spec.os_init, spec.noise_init, spec)


def log_prior(x: torch.Tensor, prior: Tuple[int, float, float]) -> torch.Tensor:
    """Summed log density of `prior` at (constrained) `x` (torch.distributions forms)."""
    kind, p1, p2 = prior
    if kind == PRIOR_NONE:
        return torch.zeros((), dtype=DT)
    if kind == PRIOR_GAMMA:
        return (p1 * math.log(p2) + (p1 - 1.0) * torch.log(x) - p2 * x - math.lgamma(p1)).sum()
    if kind == PRIOR_LOGNORMAL:
        lx = torch.log(x)
        return (-lx - math.log(p2) - 0.5 * math.log(2 * math.pi) - (lx - p1) ** 2 / (2 * p2 * p2)).sum()
    raise ValueError(kind)


# --------------------------------------------------------------------------- #
# standardisation (botorch Standardize(m=1); reference model.py:185)
# --------------------------------------------------------------------------- #
def standardize(Y: torch.Tensor) -> Tuple[torch.Tensor, float, float]:
    Y = Y.reshape(-1).to(DT)
    ybar = Y.mean()
    s = Y.std(unbiased=True) if Y.numel() > 1 else torch.tensor(float("nan"), dtype=DT)
    if not bool(s >= 1e-8):
        s = torch.ones((), dtype=DT)
    return (Y - ybar) / s, float(ybar), float(s)


# --------------------------------------------------------------------------- #
# kernel
# --------------------------------------------------------------------------- #
def sq_dist(x1: torch.Tensor, x2: torch.Tensor, mode: str = "direct", same: bool = False) -> torch.Tensor:
    """Squared Euclidean distance of already length-scaled inputs.

    mode="expansion" follows gpytorch.kernels.kernel.sq_dist (quadratic expansion on
    inputs centred by mean(x1), diagonal zeroed when x1 is x2, clamp_min 0);
    mode="direct" is sum_j (a_j - b_j)^2, what the CUDA kernels evaluate.
    """
    if mode == "direct":
        diff = x1.unsqueeze(-2) - x2.unsqueeze(-3)
        return (diff * diff).sum(-1)
    adj = x1.mean(-2, keepdim=True)
    a = x1 - adj
    b = x2 - adj
    an = (a * a).sum(-1, keepdim=True)
    bn = (b * b).sum(-1, keepdim=True)
    a_ = torch.cat([-2.0 * a, an, torch.ones_like(an)], dim=-1)
    b_ = torch.cat([b, torch.ones_like(bn), bn], dim=-1)
    res = a_ @ b_.transpose(-1, -2)
    if same:
        res = res - torch.diag_embed(torch.diagonal(res, dim1=-2, dim2=-1))
    return res.clamp_min(0.0)


def kappa(r2: torch.Tensor, kernel: int) -> torch.Tensor:
    """Unit-variance stationary kernel as a function of squared scaled distance."""
    if kernel == KERNEL_RBF:
        return torch.exp(-0.5 * r2)
    r = torch.sqrt(r2.clamp_min(1e-30))
    if kernel == KERNEL_MATERN12:
        return torch.exp(-r)
    if kernel == KERNEL_MATERN32:
        return (1.0 + math.sqrt(3.0) * r) * torch.exp(-math.sqrt(3.0) * r)
    if kernel == KERNEL_MATERN52:
        return (1.0 + math.sqrt(5.0) * r + (5.0 / 3.0) * r * r) * torch.exp(-math.sqrt(5.0) * r)
    raise ValueError(kernel)


def kernel_matrix(
    X1: torch.Tensor,
    X2: torch.Tensor,
    ls: torch.Tensor,
    os_: torch.Tensor,
    kernel: int = KERNEL_RBF,
    mode: str = "direct",
    same: bool = False,
) -> torch.Tensor:
    """ScaleKernel(base(ard))(X1, X2) = s * kappa(r^2) (no noise)."""
    if kernel != KERNEL_RBF and mode == "expansion":
        # gpytorch MaternKernel centres by the mean over x1 before scaling
        mu = X1.reshape(-1, X1.shape[-1]).mean(0)
        X1 = X1 - mu
        X2 = X2 - mu
    r2 = sq_dist(X1 / ls, X2 / ls, mode=mode, same=same)
    return os_ * kappa(r2, kernel)


# --------------------------------------------------------------------------- #
# objective: (LML + log priors) / n       reference utils.py:171-177 (+ gpytorch MLL)
# --------------------------------------------------------------------------- #
def lml_objective(
    X: torch.Tensor,
    ytil: torch.Tensor,
    theta_raw: torch.Tensor,
    spec: HyperSpec,
    mode: str = "direct",
    jitter: float = 0.0,
) -> torch.Tensor:
    n = X.shape[-2]
    ls, os_, noise = split_theta(theta_raw, spec)
    K = kernel_matrix(X, X, ls, os_, spec.kernel, mode=mode, same=True)
    Ky = K + (noise + jitter) * torch.eye(n, dtype=DT)
    L = torch.linalg.cholesky(Ky)
    z = torch.linalg.solve_triangular(L, ytil.reshape(-1, 1), upper=False)
    quad = (z * z).sum()
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum()
    res = -0.5 * (quad + logdet + n * math.log(2 * math.pi))
    res = res + log_prior(ls, spec.ls_prior) + log_prior(os_, spec.os_prior) + log_prior(noise, spec.noise_prior)
    return res / n


def lml_and_grad_autograd(X, ytil, theta_raw, spec, mode="direct", jitter=0.0):
    """Value + gradient wrt RAW parameters by autograd (how the reference gets it)."""
    t = theta_raw.detach().clone().requires_grad_(True)
    val = lml_objective(X, ytil, t, spec, mode=mode, jitter=jitter)
    (g,) = torch.autograd.grad(val, t)
    return val.detach(), g


def _neg2_dkappa(r2: torch.Tensor, kernel: int) -> torch.Tensor:
    """-2 * d kappa / d r^2."""
    if kernel == KERNEL_RBF:
        return torch.exp(-0.5 * r2)
    r = torch.sqrt(r2.clamp_min(1e-30))
    if kernel == KERNEL_MATERN12:
        return torch.where(r2 > 0, torch.exp(-r) / r, torch.zeros_like(r))
    if kernel == KERNEL_MATERN32:
        return 3.0 * torch.exp(-math.sqrt(3.0) * r)
    if kernel == KERNEL_MATERN52:
        return (5.0 / 3.0) * (1.0 + math.sqrt(5.0) * r) * torch.exp(-math.sqrt(5.0) * r)
    raise ValueError(kernel)


def _dlog_prior(x: torch.Tensor, prior) -> torch.Tensor:
    kind, p1, p2 = prior
    if kind == PRIOR_NONE:
        return torch.zeros_like(x)
    if kind == PRIOR_GAMMA:
        return (p1 - 1.0) / x - p2
    lx = torch.log(x)
    return -1.0 / x - (lx - p1) / (p2 * p2 * x)


def lml_and_grad_analytic(X, ytil, theta_raw, spec, jitter=0.0):
    """Same quantity through the trace formula (what the CUDA kernel computes)."""
    n, d = X.shape
    ls, os_, noise = split_theta(theta_raw, spec)
    Xs = X / ls
    diff2 = (Xs.unsqueeze(1) - Xs.unsqueeze(0)) ** 2  # n,n,d  (= u_abj)
    r2 = diff2.sum(-1)
    kap = kappa(r2, spec.kernel)
    K = os_ * kap
    Ky = K + (noise + jitter) * torch.eye(n, dtype=DT)
    L = torch.linalg.cholesky(Ky)
    Linv = torch.linalg.solve_triangular(L, torch.eye(n, dtype=DT), upper=False)
    Kinv = Linv.T @ Linv
    z = Linv @ ytil
    alpha = Linv.T @ z
    quad = (z * z).sum()
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum()
    val = -0.5 * (quad + logdet + n * math.log(2 * math.pi))
    val = val + log_prior(ls, spec.ls_prior) + log_prior(os_, spec.os_prior) + log_prior(noise, spec.noise_prior)
    W = torch.outer(alpha, alpha) - Kinv
    G = os_ * _neg2_dkappa(r2, spec.kernel)
    g_ls = 0.5 * torch.einsum("ab,ab,abj->j", W, G, diff2) / ls + _dlog_prior(ls, spec.ls_prior)
    g_os = 0.5 * (W * kap).sum() + _dlog_prior(os_, spec.os_prior)
    g_noise = 0.5 * torch.trace(W) + _dlog_prior(noise, spec.noise_prior)
    g = torch.cat([g_ls, g_os.reshape(1), g_noise.reshape(1)])

    def dsig(raw, lo, hi):
        sg = torch.sigmoid(raw)
        return (hi - lo) * sg * (1 - sg)

    chain = torch.cat(
        [
            dsig(theta_raw[:d], *spec.ls_bounds),
            dsig(theta_raw[d : d + 1], *spec.os_bounds),
            dsig(theta_raw[d + 1 : d + 2], *spec.noise_bounds),
        ]
    )
    return val / n, g * chain / n


def lml_with_jitter_ladder(X, ytil, theta_raw, spec, mode="direct"):
    """psd_safe_cholesky semantics: retry with 1e-8, 1e-7, 1e-6 on the diagonal."""
    for level, jit in enumerate(JITTER_LADDER):
        try:
            v, g = lml_and_grad_autograd(X, ytil, theta_raw, spec, mode=mode, jitter=jit)
            if torch.isfinite(v):
                return v, g, level
        except Exception:  # torch.linalg.LinAlgError
            continue
    nan = torch.full((), float("nan"), dtype=DT)
    return nan, torch.full_like(theta_raw, float("nan")), -1


# --------------------------------------------------------------------------- #
# fitted state + posterior          reference model.py:128-134, SURVEY A.7
# --------------------------------------------------------------------------- #
@dataclass
class TaskState:
    X: torch.Tensor  # n x d
    ytil: torch.Tensor  # n  (standardised)
    ybar: float
    ystd: float
    ls: torch.Tensor
    os: torch.Tensor
    noise: torch.Tensor
    L: torch.Tensor
    alpha: torch.Tensor
    kernel: int = KERNEL_RBF


def factorize(X, Y, theta_raw, spec, jitter=0.0) -> TaskState:
    ytil, ybar, ystd = standardize(Y)
    ls, os_, noise = split_theta(theta_raw, spec)
    n = X.shape[0]
    Ky = kernel_matrix(X, X, ls, os_, spec.kernel) + (noise + jitter) * torch.eye(n, dtype=DT)
    L = torch.linalg.cholesky(Ky)
    alpha = torch.cholesky_solve(ytil.reshape(-1, 1), L).reshape(-1)
    return TaskState(X, ytil, ybar, ystd, ls, os_, noise, L, alpha, spec.kernel)


def posterior(st: TaskState, Xs: torch.Tensor, full_cov: bool = False):
    """Un-standardised posterior of one source GP at Xs (no observation noise)."""
    Ks = kernel_matrix(Xs, st.X, st.ls, st.os, st.kernel)  # B x n
    mu = Ks @ st.alpha
    V = torch.linalg.solve_triangular(st.L, Ks.T, upper=False)  # n x B
    if full_cov:
        Kss = kernel_matrix(Xs, Xs, st.ls, st.os, st.kernel)
        cov = Kss - V.T @ V
        return st.ybar + st.ystd * mu, st.ystd**2 * cov
    var = st.os - (V * V).sum(0)
    return st.ybar + st.ystd * mu, st.ystd**2 * var


def significant_weights_mask(weights, std_y, threshold):
    """reference model.py:192-215."""
    ws = weights * std_y
    return ws * len(weights) / ws.sum() >= threshold


def scaml_prior_predict(states: Sequence[TaskState], weights: torch.Tensor, Xs: torch.Tensor,
                        prune_threshold: Optional[float] = None):
    """n_t = 0 ScaML-GP prediction: sum_i w_i mu_i, sum_i w_i^2 var_i (model.py:108-135).

    The target kernel's own variance k_t(x,x) is NOT added here (the host adds it).
    """
    mean = torch.zeros(Xs.shape[0], dtype=DT)
    var = torch.zeros(Xs.shape[0], dtype=DT)
    mask = torch.ones(len(states), dtype=torch.bool)
    if prune_threshold is not None:
        std_y = torch.tensor([s.ystd for s in states], dtype=DT)
        mask = significant_weights_mask(weights, std_y, prune_threshold)
    for st, w, keep in zip(states, weights, mask):
        if not bool(keep):
            continue
        m, v = posterior(st, Xs)
        mean = mean + w * m
        var = var + w * w * v
    return mean, var


# --------------------------------------------------------------------------- #
# target GP (ScaMLGP train / eval forward)      reference model.py:219-384, A.8
# --------------------------------------------------------------------------- #
@dataclass
class TargetCache:
    Xt: torch.Tensor  # n_t x d
    yt_std: torch.Tensor  # n_t, standardised with the all-data transform
    mu_all: float
    s_all: float
    source_means: torch.Tensor  # n_t x M   (raw-Y units)
    source_covs: torch.Tensor  # n_t x n_t x M


def all_data_standardizer(states: Sequence[TaskState], Yt: torch.Tensor):
    """model.py:264-276: mean/std over the concatenation of all raw source Y and target Y."""
    ys = [st.ybar + st.ystd * st.ytil for st in states] + [Yt.reshape(-1).to(DT)]
    Yall = torch.cat(ys)
    mu = Yall.mean()
    s = Yall.std(unbiased=True) if Yall.numel() > 1 else torch.tensor(float("nan"), dtype=DT)
    if not bool(s >= 1e-8):
        s = torch.ones((), dtype=DT)
    return float(mu), float(s)


def build_target_cache(states, Xt, Yt) -> TargetCache:
    mu_all, s_all = all_data_standardizer(states, Yt)
    means, covs = [], []
    for st in states:
        m, c = posterior(st, Xt, full_cov=True)
        means.append(m)
        covs.append(c)
    return TargetCache(Xt, (Yt.reshape(-1).to(DT) - mu_all) / s_all, mu_all, s_all,
                       torch.stack(means, -1), torch.stack(covs, -1))


def target_objective(cache: TargetCache, weights: torch.Tensor, theta_raw: torch.Tensor,
                     spec: HyperSpec, weights_prior=(PRIOR_GAMMA, 1.0, 1.0)) -> torch.Tensor:
    """(LML + priors)/n_t of the ScaML-GP target model in training mode (model.py:360-383)."""
    nt = cache.Xt.shape[0]
    ls, os_, noise = split_theta(theta_raw, spec)
    mean = (cache.source_means @ weights - cache.mu_all) / cache.s_all
    cov = (cache.source_covs @ weights**2) / cache.s_all**2
    cov = cov + kernel_matrix(cache.Xt, cache.Xt, ls, os_, spec.kernel) + noise * torch.eye(nt, dtype=DT)
    L = torch.linalg.cholesky(cov)
    z = torch.linalg.solve_triangular(L, (cache.yt_std - mean).reshape(-1, 1), upper=False)
    res = -0.5 * ((z * z).sum() + 2.0 * torch.log(torch.diagonal(L)).sum() + nt * math.log(2 * math.pi))
    res = res + log_prior(ls, spec.ls_prior) + log_prior(os_, spec.os_prior) + log_prior(noise, spec.noise_prior)
    res = res + log_prior(weights, weights_prior)
    return res / nt


def scaml_posterior(states, weights, cache: Optional[TargetCache], theta_raw, spec: HyperSpec,
                    Xs: torch.Tensor, prune_threshold: Optional[float] = 1e-3, full_cov: bool = False):
    """Full ScaML-GP posterior mean/variance at Xs (q=1 per candidate), un-standardised; full_cov=True returns the
    JOINT covariance [q, q] of the points Xs instead of their variances (one q-batch of model.py:364-375).

    n_t = 0  -> prior: sum_i w_i mu_i(x), sum_i w_i^2 var_i(x) + s_t   (A.8 last bullets)
    n_t > 0  -> exact conditioning on the target data in the all-data standardised space.
    """
    ls, os_, noise = split_theta(theta_raw, spec)
    M = len(states)
    mask = torch.ones(M, dtype=torch.bool)
    if prune_threshold is not None:
        std_y = torch.tensor([s.ystd for s in states], dtype=DT)
        mask = significant_weights_mask(weights, std_y, prune_threshold)
    if cache is None or cache.Xt.shape[0] == 0:
        if full_cov:
            mean = torch.zeros(Xs.shape[0], dtype=DT)
            cov = kernel_matrix(Xs, Xs, ls, os_, spec.kernel)
            for st, w, keep in zip(states, weights, mask):
                if bool(keep):
                    m, c = posterior(st, Xs, full_cov=True)
                    mean, cov = mean + w * m, cov + w * w * c
            return mean, cov
        mean, var = scaml_prior_predict(states, weights, Xs, prune_threshold)
        return mean, var + os_
    nt = cache.Xt.shape[0]
    Xj = torch.cat([cache.Xt, Xs], 0)
    mean_j = torch.zeros(Xj.shape[0], dtype=DT)
    cov_j = torch.zeros(Xj.shape[0], Xj.shape[0], dtype=DT)
    for st, w, keep in zip(states, weights, mask):
        if not bool(keep):
            continue
        m, c = posterior(st, Xj, full_cov=True)
        mean_j = mean_j + w * m
        cov_j = cov_j + w * w * c
    mean_j = (mean_j - cache.mu_all) / cache.s_all
    cov_j = cov_j / cache.s_all**2 + kernel_matrix(Xj, Xj, ls, os_, spec.kernel)
    Ktt = cov_j[:nt, :nt] + noise * torch.eye(nt, dtype=DT)
    Kst = cov_j[nt:, :nt]
    L = torch.linalg.cholesky(Ktt)
    a = torch.cholesky_solve((cache.yt_std - mean_j[:nt]).reshape(-1, 1), L).reshape(-1)
    mu = mean_j[nt:] + Kst @ a
    V = torch.linalg.solve_triangular(L, Kst.T, upper=False)
    if full_cov:
        return cache.mu_all + cache.s_all * mu, cache.s_all**2 * (cov_j[nt:, nt:] - V.T @ V)
    var = torch.diagonal(cov_j[nt:, nt:]) - (V * V).sum(0)
    return cache.mu_all + cache.s_all * mu, cache.s_all**2 * var


def scaml_posterior_grad(states, weights, cache: Optional[TargetCache], theta_raw, spec: HyperSpec,
                         Xs: torch.Tensor, prune_threshold: Optional[float] = 1e-3):
    """d mean / d x and d var / d x [B, d] of `scaml_posterior` by autograd -- what botorch's optimize_acqf
    differentiates when it runs L-BFGS-B on the acquisition function over `ScaMLGP.forward` (eval branch,
    model.py:364-375; consumer utils.py:215-224).  Every candidate's mean / variance depends on its own row of Xs
    only (q = 1), so the gradient of the sum over candidates is the per-candidate gradient."""
    Xg = Xs.clone().requires_grad_(True)
    mean, var = scaml_posterior(states, weights, cache, theta_raw, spec, Xg, prune_threshold)
    (dmean,) = torch.autograd.grad(mean.sum(), Xg, retain_graph=True)
    (dvar,) = torch.autograd.grad(var.sum(), Xg)
    return mean.detach(), var.detach(), dmean, dvar


def ucb(mean: torch.Tensor, var: torch.Tensor, beta: float = 9.0) -> torch.Tensor:
    """reference utils.py:215-224: botorch UCB with maximize=False -> -mu + sqrt(beta var)."""
    return -mean + torch.sqrt(beta * var.clamp_min(1e-9))


# --------------------------------------------------------------------------- #
# synthetic meta-data: the generators live in the neutral `datagen.py` at the repo root (bench.py's product arm
# must not import the oracle); re-exported here for the tests that take their inputs from the oracle namespace
# --------------------------------------------------------------------------- #
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)
from datagen import hartmann6, sample_theta_raw, synthetic_tasks  # noqa: E402,F401
