"""Batched projected L-BFGS: E independent minimisations advanced in lock-step.

The reference fits every GP with scipy's L-BFGS-B, one optimisation at a time, one
objective evaluation per Python closure call (botorch `fit_gpytorch_mll` from
scamlgp/utils.py:175,190; defaults maxiter 15000, ftol 2.2e-9, gtol 1e-5, m = 10).  Here
all E = tasks x restarts rows (source GPs) or R restart rows (target GP) move together:
every round is ONE batched objective call -- one `scaml_lml_grad` /
`scaml_target_lml_grad` launch over the rows that are still active -- and the two-loop
recursion, the Armijo bookkeeping and the convergence tests are elementwise torch ops on
[E, D] tensors living next to the kernel outputs (no per-row host round trip).

Rows are independent: row e's iterates depend only on row e's data, whichever other rows
share the batch, so results are deterministic and invariant to batch composition.

Semantics kept from L-BFGS-B: history m = 10, convergence on the projected-gradient
inf-norm (gtol) or the relative decrease (ftol), simple lower bounds (the ScaML-GP weights
are bounded below by 1e-10 with no transform, scamlgp/model.py:333-337).  A NaN objective
(non-PSD Cholesky, scamlgp/utils.py:180-198) at a trial point shortens the step; at the
starting point it fails the row.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch


@dataclass
class LbfgsResult:
    x: torch.Tensor  # [E, D] final iterates
    f: torch.Tensor  # [E] objective at x (NaN for failed rows)
    converged: torch.Tensor  # [E] bool: gtol or ftol reached
    failed: torch.Tensor  # [E] bool: objective not finite at the start / line search collapsed at once
    iterations: torch.Tensor  # [E] int64 accepted steps
    evaluations: int  # batched objective calls issued


def _projected_grad_inf(x, g, lower):
    if lower is None:
        return g.abs().amax(1)
    return (x - torch.maximum(x - g, lower)).abs().amax(1)


def _two_loop(g, S, Y, rho, count, head, m):
    """d = -H g with per-row circular history.  S, Y: [E, m, D]; rho: [E, m]; count: [E]."""
    E, D = g.shape
    q = g.clone()
    alphas = []
    idxs = []
    ar = torch.arange(E, device=g.device)
    for j in range(m):  # newest -> oldest
        idx = (head - 1 - j) % m
        valid = (count > j).to(g.dtype)
        s = S[ar, idx]
        y = Y[ar, idx]
        a = rho[ar, idx] * (s * q).sum(1) * valid
        q = q - a.unsqueeze(1) * y
        alphas.append(a)
        idxs.append((idx, valid, s, y))
    # initial scaling gamma = s.y / y.y of the newest pair
    idx0 = (head - 1) % m
    s0, y0 = S[ar, idx0], Y[ar, idx0]
    yy = (y0 * y0).sum(1)
    gamma = torch.where((count > 0) & (yy > 0), (s0 * y0).sum(1) / yy.clamp_min(1e-300), torch.ones_like(yy))
    r = q * gamma.unsqueeze(1)
    for j in reversed(range(m)):  # oldest -> newest
        idx, valid, s, y = idxs[j]
        b = rho[ar, idx] * (y * r).sum(1) * valid
        r = r + (alphas[j] - b).unsqueeze(1) * s
    return -r


def lbfgs_minimize(fun: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                   x0: torch.Tensor, lower: Optional[torch.Tensor] = None, maxiter: int = 200,
                   gtol: float = 1e-5, ftol: float = 2.2e-9, history: int = 10, max_ls: int = 20,
                   max_evals: Optional[int] = None) -> LbfgsResult:
    """Minimise E objectives f_e(x_e) jointly.

    fun(x [E, D], active [E] bool) -> (f [E], g [E, D]); entries of inactive rows are ignored
    (the kernels skip them).  lower: None, or a tensor broadcastable to [E, D] (-inf = free).
    """
    E, D = x0.shape
    dev, dt = x0.device, x0.dtype
    m = history
    x = x0.clone()
    if lower is not None:
        lower = lower.to(device=dev, dtype=dt).expand(E, D)
        x = torch.maximum(x, lower)
    all_rows = torch.ones(E, dtype=torch.bool, device=dev)
    f, g = fun(x, all_rows)
    f, g = f.clone(), g.clone()
    evals = 1
    failed = ~(torch.isfinite(f) & torch.isfinite(g).all(1))
    g = torch.where(failed.unsqueeze(1), torch.zeros_like(g), g)
    converged = (~failed) & (_projected_grad_inf(x, g, lower) <= gtol)
    active = ~(failed | converged)
    iters = torch.zeros(E, dtype=torch.int64, device=dev)

    S = torch.zeros(E, m, D, dtype=dt, device=dev)
    Y = torch.zeros(E, m, D, dtype=dt, device=dev)
    rho = torch.zeros(E, m, dtype=dt, device=dev)
    count = torch.zeros(E, dtype=torch.int64, device=dev)
    head = torch.zeros(E, dtype=torch.int64, device=dev)
    ar = torch.arange(E, device=dev)

    def direction(xc, gc):
        gg = gc
        if lower is not None:
            # active set: variables sitting on their bound with the gradient pushing outward stay fixed
            fixed = (xc <= lower) & (gc > 0)
            gg = torch.where(fixed, torch.zeros_like(gc), gc)
        d = _two_loop(gg, S, Y, rho, count, head, m)
        if lower is not None:
            d = torch.where(fixed, torch.zeros_like(d), d)
        slope = (gc * d).sum(1)
        bad = ~(slope < 0)  # not a descent direction (stale curvature): steepest descent
        d = torch.where(bad.unsqueeze(1), -gg, d)
        slope = torch.where(bad, -(gg * gg).sum(1), slope)
        return d, slope, bad

    d, slope, _ = direction(x, g)
    dn = d.norm(dim=1).clamp_min(1e-300)
    t = torch.minimum(torch.ones_like(dn), 1.0 / dn)  # L-BFGS-B first step: min(1, 1/||d||)
    ls_count = torch.zeros(E, dtype=torch.int64, device=dev)
    budget = max_evals if max_evals is not None else maxiter * 3 + max_ls

    while evals < budget and bool(active.any()):
        xt = x + t.unsqueeze(1) * d
        if lower is not None:
            xt = torch.maximum(xt, lower)
        ft, gt = fun(xt, active)
        evals += 1
        step = xt - x
        # Armijo on the projected step
        dec = (g * step).sum(1)
        finite = torch.isfinite(ft) & torch.isfinite(gt).all(1)
        ok = active & finite & (ft <= f + 1e-4 * dec)
        # ---- accepted rows: curvature pair, convergence tests, next direction ------------ #
        s_new = step
        y_new = torch.where(ok.unsqueeze(1), gt - g, torch.zeros_like(g))
        sy = (s_new * y_new).sum(1)
        upd = ok & (sy > 1e-10 * s_new.norm(dim=1) * y_new.norm(dim=1)) & (sy > 0)
        hi = head[upd]
        ui = ar[upd]
        S[ui, hi] = s_new[upd]
        Y[ui, hi] = y_new[upd]
        rho[ui, hi] = 1.0 / sy[upd]
        head = torch.where(upd, (head + 1) % m, head)
        count = torch.where(upd, torch.clamp(count + 1, max=m), count)
        rel = (f - ft) / torch.maximum(torch.maximum(f.abs(), ft.abs()), torch.ones_like(f))
        f_prev = f
        x = torch.where(ok.unsqueeze(1), xt, x)
        g = torch.where(ok.unsqueeze(1), gt, g)
        f = torch.where(ok, ft, f)
        iters = iters + ok.to(torch.int64)
        conv_now = ok & ((_projected_grad_inf(x, g, lower) <= gtol) | (rel <= ftol))
        converged = converged | conv_now
        # ---- rejected rows: shrink (quadratic interpolation, safeguarded) ----------------- #
        rej = active & ~ok
        ls_count = torch.where(rej, ls_count + 1, torch.zeros_like(ls_count))
        denom = 2.0 * (ft - f_prev - dec)
        tq = torch.where(finite & (denom > 0), -dec / denom.clamp_min(1e-300), torch.full_like(t, 0.5))
        shrink = torch.clamp(tq, 0.1, 0.5)
        t_rej = t * shrink
        stalled = rej & (ls_count >= max_ls)
        # a row whose line search collapses has reached the resolution of the objective: stop it
        converged = converged | (stalled & (iters > 0))
        failed = failed | (stalled & (iters == 0))
        hit_max = ok & (iters >= maxiter)
        active = active & ~(conv_now | stalled | hit_max)
        # ---- next trial points --------------------------------------------------------------- #
        d_new, slope_new, _ = direction(x, g)
        d = torch.where(ok.unsqueeze(1), d_new, d)
        t = torch.where(ok, torch.ones_like(t), t_rej)
        # first accepted step of a row without curvature yet keeps the cautious 1/||d|| scaling
        nocurv = ok & (count == 0)
        t = torch.where(nocurv, torch.minimum(torch.ones_like(t), 1.0 / d.norm(dim=1).clamp_min(1e-300)), t)

    f_out = torch.where(failed, torch.full_like(f, float("nan")), f)
    return LbfgsResult(x=x, f=f_out, converged=converged, failed=failed, iterations=iters, evaluations=evals)


# ------------------------------------------------------------------------------------------------ #
# device-side driver: the same algorithm with the whole update in ONE kernel (csrc/scaml_lbfgs.cuh)
# ------------------------------------------------------------------------------------------------ #
FLAG_ACTIVE, FLAG_CONVERGED, FLAG_FAILED = 1, 2, 4


def lbfgs_minimize_device(engine, fun: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                          x0: torch.Tensor, lower: Optional[torch.Tensor] = None, maxiter: int = 200,
                          gtol: float = 1e-5, ftol: float = 2.2e-9, history: int = 10, max_ls: int = 20,
                          max_evals: Optional[int] = None, poll_every: int = 4) -> LbfgsResult:
    """`lbfgs_minimize` with the per-round update done by `scaml_lbfgs_step` (one warp per row).

    Every round is the objective launch(es) + one update launch; the host reads back only the number of
    active rows, every `poll_every` rounds.  `lower` must be a [D] tensor (or None)."""
    from ._capi import CLbfgsState

    E, D = x0.shape
    dev, dt = x0.device, torch.float64
    m = int(history)
    xt = x0.to(dt).clone().contiguous()
    low = None
    if lower is not None:
        low = lower.to(device=dev, dtype=dt).reshape(-1).contiguous()
        assert low.numel() == D, "device L-BFGS takes one lower bound per variable"
        xt = torch.maximum(xt, low)

    def buf(*shape, dtype=dt):
        return torch.zeros(*shape, dtype=dtype, device=dev)

    st = dict(x=buf(E, D), f=buf(E), g=buf(E, D), d=buf(E, D), t=buf(E), S=buf(E, m, D), Y=buf(E, m, D), rho=buf(E, m),
              count=buf(E, dtype=torch.int32), head=buf(E, dtype=torch.int32), iters=buf(E, dtype=torch.int32),
              ls_count=buf(E, dtype=torch.int32), flags=buf(E, dtype=torch.int32))
    cst = CLbfgsState(**{k: v.data_ptr() for k, v in st.items()})
    flags = st["flags"]
    active = torch.ones(E, dtype=torch.bool, device=dev)
    budget = max_evals if max_evals is not None else maxiter * 3 + max_ls
    evals = 0
    init = True
    while evals < budget:
        ft, gt = fun(xt, active)
        ft, gt = ft.contiguous(), gt.contiguous()
        evals += 1
        engine.lib.lbfgs_step(cst, xt.data_ptr(), ft.data_ptr(), gt.data_ptr(), None if low is None else low.data_ptr(),
                              E, D, m, init, gtol, ftol, maxiter, max_ls, engine._stream())
        engine.launches += 1
        init = False
        active = (flags & FLAG_ACTIVE) != 0
        if evals % poll_every == 0 and not bool(active.any()):
            break
    failed = (flags & FLAG_FAILED) != 0
    converged = (flags & FLAG_CONVERGED) != 0
    f_out = torch.where(failed, torch.full_like(st["f"], float("nan")), st["f"])
    return LbfgsResult(x=st["x"], f=f_out, converged=converged, failed=failed, iterations=st["iters"].to(torch.int64),
                       evaluations=evals)
