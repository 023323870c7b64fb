"""CPU restatement of the reference's BO loop on the oracle -- TEST INFRASTRUCTURE (the checker of row f4, never the
product path; imported only by tests/, scripts/bo_parity.py's oracle arm and examples' comparison tables).

What the reference does per study (scamlgp/optimizer.py:28-185, utils.py:139-224; third-party pieces restated from
SURVEY A.6 / A.9):
  * meta-fit: every source GP, (LML + priors)/n maximised by scipy L-BFGS-B from the default parameters plus 5 prior
    draws, best final value wins (model.py:138-189, utils.py:139-212);
  * report: a new ScaMLGP on all valid evaluations (weights reset to 1/M, kernel / likelihood carried over), refitted from
    the current parameters plus `num_restarts` prior draws (optimizer.py:176-185);
  * suggestion: UCB (beta = 9, minimisation: -mu + 3 sigma) maximised over [0,1]^d by raw-sample screening followed by
    L-BFGS-B from the best raw samples (botorch optimize_acqf, A.9), autograd gradients.
All arithmetic goes through oracle/scaml_oracle.py (torch fp64 on the CPU).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from scipy.optimize import minimize

from oracle import scaml_oracle as O

DT = torch.float64


def _draw_prior(prior, shape, rng: np.random.Generator) -> np.ndarray:
    kind, p1, p2 = prior
    if kind == O.PRIOR_GAMMA:
        return rng.gamma(p1, 1.0 / p2, shape)
    if kind == O.PRIOR_LOGNORMAL:
        return np.exp(p1 + p2 * rng.standard_normal(shape))
    raise ValueError(kind)


def sample_theta_raw(spec: O.HyperSpec, d: int, rng: np.random.Generator) -> torch.Tensor:
    """One draw of every prior, mapped to raw space; rejected (<= 5 retries) while not finite (utils.py:31-69)."""
    for _ in range(6):
        ls = torch.tensor(_draw_prior(spec.ls_prior, d, rng), dtype=DT)
        os_ = torch.tensor(_draw_prior(spec.os_prior, 1, rng), dtype=DT)
        nz = torch.tensor(_draw_prior(spec.noise_prior, 1, rng), dtype=DT)
        raw = torch.cat([O.unconstrain(ls, *spec.ls_bounds), O.unconstrain(os_, *spec.os_bounds),
                         O.unconstrain(nz, *spec.noise_bounds)])
        if bool(torch.isfinite(raw).all()):
            return raw
    raise RuntimeError("prior sampling failed")


def fit_source(X: torch.Tensor, Y: torch.Tensor, spec: O.HyperSpec, num_restarts: int, rng) -> torch.Tensor:
    d = X.shape[1]
    yt = O.standardize(Y)[0]

    def f(x):
        try:
            v, g = O.lml_and_grad_autograd(X, yt, torch.tensor(x, dtype=DT), spec)
        except Exception:  # NotPSD -> NaN loss, scipy backs off (utils.py:180-198)
            return float("nan"), np.zeros_like(x)
        return -float(v), -g.numpy()

    starts = [O.initial_theta_raw(d, spec)] + [sample_theta_raw(spec, d, rng) for _ in range(num_restarts)]
    best, best_v = None, -np.inf
    for s in starts:
        res = minimize(f, s.numpy(), jac=True, method="L-BFGS-B")
        th = torch.tensor(res.x, dtype=DT)
        try:
            v = float(O.lml_objective(X, yt, th, spec))
        except Exception:
            continue
        if np.isfinite(v) and v > best_v:
            best, best_v = th, v
    if best is None:
        raise RuntimeError("Hyperparameter optimization failed for all attempts.")
    return best


def fit_target(cache: O.TargetCache, M: int, th0: torch.Tensor, spec: O.HyperSpec, num_restarts: int, rng
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    d = cache.Xt.shape[1]

    def f(x):
        w = torch.tensor(x[:M], dtype=DT, requires_grad=True)
        th = torch.tensor(x[M:], dtype=DT, requires_grad=True)
        try:
            v = O.target_objective(cache, w, th, spec)
            gw, gt = torch.autograd.grad(v, [w, th])
        except Exception:
            return float("nan"), np.zeros_like(x)
        return -float(v), -torch.cat([gw, gt]).numpy()

    starts = [(torch.full((M,), 1.0 / M, dtype=DT), th0)]
    for _ in range(num_restarts):
        starts.append((torch.tensor(rng.gamma(1.0, 1.0, M), dtype=DT).clamp_min(1e-10), sample_theta_raw(spec, d, rng)))
    bounds = [(1e-10, None)] * M + [(None, None)] * (d + 2)
    best, best_v = None, -np.inf
    for w0, t0 in starts:
        res = minimize(f, torch.cat([w0, t0]).numpy(), jac=True, method="L-BFGS-B", bounds=bounds)
        w, th = torch.tensor(res.x[:M], dtype=DT), torch.tensor(res.x[M:], dtype=DT)
        try:
            v = float(O.target_objective(cache, w, th, spec))
        except Exception:
            continue
        if np.isfinite(v) and v > best_v:
            best, best_v = (w, th), v
    if best is None:
        raise RuntimeError("Hyperparameter optimization failed for all attempts.")
    return best


def suggest(states, w, cache: Optional[O.TargetCache], th, spec: O.HyperSpec, d: int, gen: torch.Generator,
            raw_samples: int, num_restarts: int, maxiter: int) -> np.ndarray:
    def af(X):
        m, v = O.scaml_posterior(states, w, cache, th, spec, X)
        return O.ucb(m, v)

    X0 = torch.rand(raw_samples, d, dtype=DT, generator=gen)
    with torch.no_grad():
        v0 = af(X0)
    starts = X0[torch.topk(v0, min(num_restarts, raw_samples)).indices]
    S = starts.shape[0]

    def f(x):
        X = torch.tensor(x.reshape(S, d), dtype=DT, requires_grad=True)
        val = af(X).sum()
        (g,) = torch.autograd.grad(val, X)
        return -float(val.detach()), -g.numpy().reshape(-1)

    res = minimize(f, starts.numpy().reshape(-1), jac=True, method="L-BFGS-B", bounds=[(0.0, 1.0)] * (S * d),
                   options=dict(maxiter=maxiter))
    cand = torch.cat([torch.tensor(res.x.reshape(S, d), dtype=DT).clamp(0, 1), starts])
    with torch.no_grad():
        vals = af(cand)
    return cand[int(torch.argmax(vals))].numpy()


def run_study(meta_X: Sequence[np.ndarray], meta_y: Sequence[np.ndarray], bounds: np.ndarray, objective, noise: float,
              noise_rng: np.random.Generator, evals: int, seed: int, num_restarts: int = 5, raw_samples: int = 1024,
              af_restarts: int = 32, af_maxiter: int = 60) -> List[float]:
    """Noise-free objective values of one study (inputs in the original space; the GP sees [0,1]^d like
    `to_numerical` maps them, utils.py:98-106)."""
    lo, hi = bounds[:, 0], bounds[:, 1]
    d = len(lo)
    rng = np.random.default_rng(seed + 7919)  # restart draws of the fits
    gen = torch.Generator().manual_seed(seed)
    sspec, tspec = O.HyperSpec.source(), O.HyperSpec.target()
    states = []
    for X, y in zip(meta_X, meta_y):
        Xu = torch.tensor((X - lo) / (hi - lo), dtype=DT)
        order = np.lexsort(np.asarray(Xu).T[::-1])  # sorted evaluations: order independence (utils.py:99)
        Xu, yt = Xu[order], torch.tensor(y, dtype=DT)[order]
        states.append(O.factorize(Xu, yt, fit_source(Xu, yt, sspec, num_restarts, rng), sspec))
    M = len(states)
    w, th = torch.full((M,), 1.0 / M, dtype=DT), O.initial_theta_raw(d, tspec)
    cache = None
    Xs, ys, values = [], [], []
    for _ in range(evals):
        x = suggest(states, w, cache, th, tspec, d, gen, raw_samples, af_restarts, af_maxiter)
        f = objective(lo + x * (hi - lo))
        values.append(f)
        Xs.append(x)
        ys.append(f + noise_rng.normal(0, noise))
        cache = O.build_target_cache(states, torch.tensor(np.array(Xs), dtype=DT), torch.tensor(ys, dtype=DT))
        w, th = fit_target(cache, M, th, tspec, num_restarts, rng)
    return values
