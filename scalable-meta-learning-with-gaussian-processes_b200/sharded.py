"""Task sharding over the GPUs of one box: one process per GPU, torch.distributed for plumbing.

The source GPs factor independently (reference scamlgp/model.py:176-188 fits them one by one and
model.py:128 queries them one by one), so the M tasks are cut into contiguous blocks, one per rank:

  * fit       -- no collective on the data path: every rank optimises its own tasks x restarts; afterwards one
                 `all_gather` of the fitted rows [M/G, P+1] so that every rank knows every task's parameters;
  * predict   -- every rank reduces its own tasks inside the prediction kernel (weighted mean / variance /
                 cross-covariance partials for the replicated candidates), then ONE `all_reduce(sum)` over the
                 stacked partials ([2, B] for q = 1);
  * target    -- the per-task caches `source_means` [n_t, M] / `source_covs` [n_t, n_t, M] (model.py:278-289) are
                 computed for the local tasks and `all_gather`ed along the task axis; the small target-GP fit
                 (n_t <= 116) then runs replicated and bit-identically on every rank.

Traffic is tiny next to the compute (16 MB per all-reduce at B = 1 Mi against ~1 s of FP64 work), so plain
NCCL collectives are used; there is nothing to fuse.  The same code runs on the `gloo` backend with CPU
tensors -- that is how tests/test_sharded.py covers it at world_size 2 without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ._capi import HyperSpec
from .engine import Engine, FittedSources, SourceBatch, TargetState
from .fit import SourceFit, fit_sources

DT = torch.float64


def task_partition(M: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of ceil(M / world) tasks; trailing ranks may own fewer (or zero) tasks."""
    per = (M + world - 1) // world
    return [(min(r * per, M), min((r + 1) * per, M)) for r in range(world)]


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _all_gather_rows(local: torch.Tensor, counts: Sequence[int], group) -> torch.Tensor:
    """Concatenate per-rank row blocks of different heights along dim 0 (padded all_gather)."""
    rank, world = _world(group)
    if world == 1:
        return local
    per = max(max(counts), 1)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


@dataclass
class ShardedFit:
    local: SourceFit  # this rank's tasks
    theta_raw: torch.Tensor  # [M, P] all tasks (gathered)
    lml: torch.Tensor  # [M]


class ShardedSources:
    """This rank's block of the source GPs plus the collectives that make it look like all M of them."""

    def __init__(self, engine: Engine, tasks: Sequence[Tuple[torch.Tensor, torch.Tensor]], group=None):
        """tasks: ALL M tasks as (X_i [n_i, d], Y_i [n_i]) host tensors, identical on every rank (meta-data are
        a few MB); only this rank's block is moved to its GPU."""
        self.engine, self.group = engine, group
        self.rank, self.world = _world(group)
        self.M = len(tasks)
        self.parts = task_partition(self.M, self.world)
        self.counts = [hi - lo for lo, hi in self.parts]
        self.lo, self.hi = self.parts[self.rank]
        if self.hi <= self.lo:
            raise ValueError(f"rank {self.rank} owns no task: use at most M = {self.M} ranks")
        n_max = max(int(t[0].shape[-2]) for t in tasks)  # common padding: identical kernels on every rank
        local = list(tasks[self.lo:self.hi])
        self.batch = SourceBatch.from_ragged(local, engine.device, n_max=n_max)
        # the all-data Standardize of the target model needs every raw Y (model.py:264-276): cheap, host side
        self.all_Y = torch.cat([t[1].reshape(-1).to(DT) for t in tasks])
        self.ystd_all: Optional[torch.Tensor] = None  # [M] per-task Standardize state of ALL tasks (gathered)
        self.ybar_all: Optional[torch.Tensor] = None
        self.fitted: Optional[FittedSources] = None
        self._condA: Optional[torch.Tensor] = None  # A_m = K_m^-1 K_m(X_m, X_t) of the local tasks
        self._cond_Xt: Optional[torch.Tensor] = None  # the target inputs `_condA` was prepared for (a private copy)
        self._cond_gen = -1  # fit generation `_condA` belongs to
        self._gen = 0  # bumped by every fit() / set_parameters(): A_m depends on the hyper-parameters

    # ---- fit: no collective on the data path ----------------------------------------------------------- #
    def fit(self, spec: HyperSpec, theta_init: torch.Tensor, fit_options: Optional[dict] = None) -> ShardedFit:
        """theta_init [M, R, P] for ALL tasks (identical on every rank: restart draws are positional)."""
        local = fit_sources(self.engine, self.batch, spec, theta_init[self.lo:self.hi].contiguous(), fit_options)
        self.fitted = self.engine.factorize(self.batch, local.theta_raw, spec)
        self._invalidate()
        rows = torch.cat([local.theta_raw, local.lml.unsqueeze(1), self.batch.ystd.unsqueeze(1),
                          self.batch.ybar.unsqueeze(1)], dim=1)
        rows = _all_gather_rows(rows, self.counts, self.group)
        P = local.theta_raw.shape[1]
        self.ystd_all = rows[:, P + 1].contiguous()
        self.ybar_all = rows[:, P + 2].contiguous()
        return ShardedFit(local, rows[:, :P].contiguous(), rows[:, P].contiguous())

    def set_parameters(self, spec: HyperSpec, theta_raw_all: torch.Tensor) -> None:
        """Factorise this rank's block at given parameters [M, P] (e.g. restored from a previous fit)."""
        th = theta_raw_all[self.lo:self.hi].to(self.engine.device, DT).contiguous()
        self.fitted = self.engine.factorize(self.batch, th, spec)
        self._invalidate()
        rows = _all_gather_rows(torch.stack([self.batch.ystd, self.batch.ybar], dim=1), self.counts, self.group)
        self.ystd_all, self.ybar_all = rows[:, 0].contiguous(), rows[:, 1].contiguous()

    def _invalidate(self) -> None:
        self._gen += 1
        self._condA = self._cond_Xt = None

    def cond_A(self, Xt: torch.Tensor) -> torch.Tensor:
        """A_m = K_m^-1 K_m(X_m, X_t) of the local tasks, cached per (fit generation, CONTENTS of X_t): a refit or
        a different set of target inputs -- even one that re-uses the old tensor's address -- recomputes it."""
        Xt = Xt.to(self.engine.device, DT).contiguous()
        hit = (self._condA is not None and self._cond_gen == self._gen and self._cond_Xt is not None
               and self._cond_Xt.shape == Xt.shape and bool(torch.equal(self._cond_Xt, Xt)))
        if not hit:
            self._condA = self.engine.cond_prepare(self.fitted, Xt)
            self._cond_Xt, self._cond_gen = Xt.clone(), self._gen
        return self._condA

    def _local_w(self, w: torch.Tensor) -> torch.Tensor:
        return w.to(self.engine.device, DT)[self.lo:self.hi].contiguous()

    # ---- predict: partials reduced in-kernel over local tasks, one all_reduce over ranks ---------------- #
    def predict_weighted(self, w: torch.Tensor, Xc: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """sum_m w_m mu_m(x), sum_m w_m^2 var_m(x) over ALL tasks; w [M], Xc [B, d] replicated."""
        B = Xc.shape[0]
        part = torch.empty(2, B, dtype=DT, device=self.engine.device)
        self.engine.predict_weighted(self.fitted, self._local_w(w), Xc.to(self.engine.device, DT),
                                     out=(part[0], part[1]))
        if self.world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return part[0], part[1]

    def predict_cross(self, w: torch.Tensor, XA: torch.Tensor, XB: Optional[torch.Tensor] = None):
        """Weighted joint prior blocks over ALL tasks: mean [nA], cov [nA, nB]."""
        mean, cov = self.engine.predict_cross(self.fitted, XA, XB, w=self._local_w(w))
        if self.world > 1:
            flat = torch.cat([mean.reshape(-1), cov.reshape(-1)])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            mean, cov = flat[: mean.numel()].reshape(mean.shape), flat[mean.numel():].reshape(cov.shape)
        return mean, cov

    def target_caches(self, Xt: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """`source_means` [n_t, M], `source_covs` [n_t, n_t, M] for ALL tasks, gathered along the task axis."""
        eng = self.engine
        if eng.cond_supported(self.fitted, Xt.shape[0]):  # from A_m (any n the prediction kernel accepts)
            sm, sc = eng.cond_caches(self.fitted, Xt, self.cond_A(Xt))
        else:
            sm, sc = eng.predict_cross(self.fitted, Xt)
        if self.world == 1:
            return sm, sc
        nt = Xt.shape[0]
        rows = torch.cat([sm.t().reshape(-1, nt), sc.permute(2, 0, 1).reshape(-1, nt * nt)], dim=1)  # [M_loc, ...]
        rows = _all_gather_rows(rows.contiguous(), self.counts, self.group)
        sm_all = rows[:, :nt].t().contiguous()
        sc_all = rows[:, nt:].reshape(self.M, nt, nt).permute(1, 2, 0).contiguous()
        return sm_all, sc_all

    def posterior(self, w: torch.Tensor, Xc: torch.Tensor, tstate: Optional[TargetState] = None,
                  prior_outputscale: float = 0.0):
        """ScaML-GP posterior mean / variance at Xc [B, d] (q = 1).  `w` are the (already pruned) weights of
        all M tasks.  tstate None -> prior-only model (n_t = 0, optimizer.py:135-141): var + s_t."""
        eng = self.engine
        Xc = Xc.to(eng.device, DT).contiguous()
        if tstate is None:
            pm, pv = self.predict_weighted(w, Xc)
            return pm, pv + prior_outputscale
        n_t = tstate.Xt.shape[0]
        if eng.cond_supported(self.fitted, n_t):
            # fused path: local prior mean / variance / cross-covariance in one prediction launch, then ONE
            # all_reduce over the stacked partials [B, 2 + n_t]
            pm, pv, cross = eng.predict_conditioned(self.fitted, self._local_w(w), Xc, tstate.Xt,
                                                    self.cond_A(tstate.Xt))
            if self.world > 1:
                flat = torch.cat([pm.unsqueeze(1), pv.unsqueeze(1), cross], dim=1).contiguous()
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                pm, pv, cross = flat[:, 0].contiguous(), flat[:, 1].contiguous(), flat[:, 2:].contiguous()
        else:
            pm, pv = self.predict_weighted(w, Xc)
            _, cross = self.predict_cross(w, Xc, tstate.Xt)
        return eng.target_posterior(tstate, pm.contiguous(), pv.contiguous(), cross.contiguous(), Xc)

    def posterior_with_grad(self, w: torch.Tensor, Xc: torch.Tensor, tstate: Optional[TargetState] = None,
                            prior_outputscale: float = 0.0):
        """`posterior` plus the analytic gradients d mean / dx, d var / dx [B, d] (B <= 128 candidates per call).

        Every rank contracts its own block of tasks (csrc/scaml_grad.cuh); the target-kernel terms are added on rank
        0 only, so ONE all_reduce(sum) over [B, 2 d] gives the full gradient (the values need the [B, 2 + n_t]
        all_reduce of `posterior` first: beta depends on the cross-covariance over ALL tasks)."""
        eng = self.engine
        Xc = Xc.to(eng.device, DT).contiguous()
        wl = self._local_w(w)
        U = eng.cond_prepare(self.fitted, Xc, wl)
        d = Xc.shape[1]
        if tstate is None:
            pm, pv, _ = eng.values_from_u(self.fitted, wl, Xc, U)
            if self.world > 1:
                flat = torch.stack([pm, pv]).contiguous()
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                pm, pv = flat[0].contiguous(), flat[1].contiguous()
            mean, var = pm, pv + prior_outputscale
            dm, dv = eng.posterior_grad(self.fitted, wl, Xc, U)
        else:
            n_t = tstate.Xt.shape[0]
            if not eng.cond_supported(self.fitted, n_t):
                raise NotImplementedError("candidate gradients need n <= 512 points per task, d <= 16, n_t <= 128")
            condA = self.cond_A(tstate.Xt)
            pm, pv, cross = eng.prior_values(self.fitted, wl, Xc, U, tstate.Xt, condA)
            if self.world > 1:
                flat = torch.cat([pm.unsqueeze(1), pv.unsqueeze(1), cross], dim=1).contiguous()
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                pm, pv, cross = flat[:, 0].contiguous(), flat[:, 1].contiguous(), flat[:, 2:].contiguous()
            mean, var, beta = eng.target_posterior_beta(tstate, pm, pv, cross, Xc)
            dm, dv = eng.posterior_grad(self.fitted, wl, Xc, U, tstate, condA, beta,
                                        target_terms=(self.rank == 0))
        if self.world > 1:
            g = torch.cat([dm, dv], dim=1).contiguous()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            dm, dv = g[:, :d].contiguous(), g[:, d:].contiguous()
        return mean, var, dm, dv
