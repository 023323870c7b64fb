"""Host layer (model.py / utils.py / fit.py mirrors of the reference API) driven end to end.

The same checks run twice:
  * not-gpu: through `EmuEngine` (CPU logic emulation of the kernel sources, tests only) at tiny sizes,
  * gpu:     through the product `Engine` (libscaml_b200.so) at the reference's experiment sizes.
The oracle is the checker: at the hyper-parameters the fit returns, every quantity the model exposes
(source posteriors, caches, target objective, conditioned posterior, UCB) must match the oracle's
restatement of scamlgp/model.py within the north-star tolerances.
"""
import numpy as np
import pytest
import torch

from oracle import scaml_oracle as O
from tests.helpers import TOL_GRAD, TOL_LML, TOL_MEAN_VAR, rel_err

DT = torch.float64


@pytest.fixture(scope="module")
def emu_engine(emu_lib):
    from tests.emu_engine import EmuEngine

    return EmuEngine(emu_lib)


def _meta_data(M, n, d, seed, ragged=False):
    from scamlgp_b200.modules import SupervisedDataset

    X, Y = O.synthetic_tasks(M, n, d, seed=seed)
    out = {}
    for i in range(M):
        ni = n if not ragged else max(1, n - 3 * i)
        out[f"task{i}"] = SupervisedDataset(X[i, :ni], Y[i, :ni].reshape(-1, 1))
    return out


def _oracle_states(gps):
    spec = O.HyperSpec.source()
    states = []
    for gp in gps.values():
        from scamlgp_b200.modules import theta_raw_of

        X = gp.train_inputs[0]
        th = theta_raw_of(gp.likelihood, gp.covar_module, X.shape[-1])
        states.append(O.factorize(X, gp._raw_Y.reshape(-1), th, spec))
    return states


def _record(line: str) -> None:
    import os

    print(line)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out) and torch.cuda.is_available():
        with open(os.path.join(out, "parity_r2.txt"), "a") as f:
            f.write(line + "\n")


def _pipeline(eng, M, n, d, n_t, B, restarts, fit_options, ragged=False):
    from scamlgp_b200.model import ScaMLGP, meta_fit_scamlgp
    from scamlgp_b200.modules import theta_raw_of
    from scamlgp_b200.utils import UpperConfidenceBound, optimize_marginal_likelihood

    md = _meta_data(M, n, d, seed=11, ragged=ragged)
    gps = meta_fit_scamlgp(md, num_restarts_log_likelihood=restarts, seed=0, engine=eng, fit_options=fit_options)
    assert list(gps.keys()) == list(md.keys())
    fit = gps.fit
    # ---- (1) the fit only ever improves on the warm start, and reports the oracle's objective ---------- #
    ospec = O.HyperSpec.source()
    th0 = O.initial_theta_raw(d, ospec)
    for i, (tid, gp) in enumerate(gps.items()):
        X, Y = md[tid].X(), md[tid].Y().reshape(-1)
        yt, ybar, ystd = O.standardize(Y)
        v0, _ = O.lml_and_grad_autograd(X, yt, th0, ospec)
        th = theta_raw_of(gp.likelihood, gp.covar_module, d)
        v1, g1 = O.lml_and_grad_autograd(X, yt, th, ospec)
        assert float(v1) >= float(v0) - 1e-12
        assert abs(float(fit.lml[i]) - float(v1)) <= TOL_LML * abs(float(v1))
        assert abs(float(gp.outcome_transform.means) - ybar) < 1e-12 * max(1.0, abs(ybar))
        assert abs(float(gp.outcome_transform.stdvs) - ystd) < 1e-12 * ystd
    states = _oracle_states(gps)
    g = torch.Generator().manual_seed(5)
    # ---- (2) a single source GP's posterior (model.py:128-134) ------------------------------------------- #
    Xq = torch.rand(9, d, dtype=DT, generator=g)
    first = next(iter(gps.values()))
    p = first.posterior(Xq)
    om, oc = O.posterior(states[0], Xq, full_cov=True)
    scale = float(states[0].os) * states[0].ystd ** 2
    assert rel_err(p.mean.reshape(-1).numpy(), om.numpy()) < TOL_MEAN_VAR
    assert float((p.mvn.covariance_matrix - oc).abs().max()) < TOL_MEAN_VAR * scale
    # ---- (3) prior-only model (n_t = 0, optimizer.py:135-141) ---------------------------------------------- #
    Xc = torch.rand(B, d, dtype=DT, generator=g)
    prior_model = ScaMLGP(torch.empty(0, d, dtype=DT), torch.empty(0, 1, dtype=DT), gps, engine=eng).eval()
    w0 = torch.full((M,), 1.0 / M, dtype=DT)
    tspec = O.HyperSpec.target()
    pp = prior_model.posterior(Xc.unsqueeze(1))
    om, ov = O.scaml_posterior(states, w0, None, O.initial_theta_raw(d, tspec), tspec, Xc)
    assert pp.mean.shape == (B, 1) and pp.variance.shape == (B, 1)
    assert rel_err(pp.mean.reshape(-1).numpy(), om.numpy()) < TOL_MEAN_VAR
    assert rel_err(pp.variance.reshape(-1).numpy(), ov.numpy()) < TOL_MEAN_VAR
    # ---- (4) target model: caches, training-branch prior, fit, conditioned posterior ----------------------- #
    Xt = torch.rand(n_t, d, dtype=DT, generator=g)
    Yt = (torch.sin(3.0 * Xt).sum(1, keepdim=True) + 0.05 * torch.randn(n_t, 1, dtype=DT, generator=g))
    model = ScaMLGP(Xt, Yt, gps, engine=eng)
    cache = O.build_target_cache(states, Xt, Yt)
    cscale = float(max(float(s.os) * s.ystd ** 2 for s in states))
    assert rel_err(model.source_means.cpu().numpy(), cache.source_means.numpy()) < TOL_MEAN_VAR
    assert float((model.source_covs.cpu() - cache.source_covs).abs().max()) < TOL_MEAN_VAR * cscale
    assert abs(float(model.outcome_transform.means) - cache.mu_all) < 1e-12 * max(1.0, abs(cache.mu_all))
    assert abs(float(model.outcome_transform.stdvs) - cache.s_all) < 1e-12 * cache.s_all
    assert rel_err(model.train_targets.numpy(), cache.yt_std.numpy()) < 1e-12
    mvn = model.forward(Xt)  # training branch (model.py:360-363,376-383)
    ls, os_, noise = O.split_theta(O.initial_theta_raw(d, tspec), tspec)
    o_mean = (cache.source_means @ w0 - cache.mu_all) / cache.s_all
    o_cov = (cache.source_covs @ w0 ** 2) / cache.s_all ** 2 + O.kernel_matrix(Xt, Xt, ls, os_, tspec.kernel)
    assert rel_err(mvn.mean.cpu().numpy(), o_mean.numpy()) < TOL_MEAN_VAR
    assert float((mvn.covariance_matrix.cpu() - o_cov).abs().max()) < TOL_MEAN_VAR * float(o_cov.abs().max())
    v_start = float(O.target_objective(cache, w0, O.initial_theta_raw(d, tspec), tspec))
    tfit = optimize_marginal_likelihood(model, restarts, generator=torch.Generator().manual_seed(1),
                                        **(fit_options or {}))
    w = model.weights
    th = model.theta_raw()
    assert bool((w >= 1e-10).all())
    v_end = float(O.target_objective(cache, w, th, tspec))
    assert v_end >= v_start - 1e-12
    e_fit = abs(tfit.lml - v_end) / abs(v_end)
    model.eval()
    post = model.posterior(Xc)
    om, ov = O.scaml_posterior(states, w, cache, th, tspec, Xc)
    vscale = float(ov.abs().max())
    e_mean = rel_err(post.mean.reshape(-1).numpy(), om.numpy())
    e_var = float((post.variance.reshape(-1) - ov).abs().max()) / vscale
    e_var_self = float(((post.variance.reshape(-1) - ov).abs() / ov).max())
    # conditioning of the n_t x n_t target system the posterior solves with (oracle side)
    ls_t, os_t, nz_t = O.split_theta(th, tspec)
    Ktt = (cache.source_covs @ w ** 2) / cache.s_all ** 2 + O.kernel_matrix(Xt, Xt, ls_t, os_t, tspec.kernel) \
        + nz_t * torch.eye(n_t, dtype=DT)
    _record(f"public API (M={M}, n={n}, d={d}, n_t={n_t}): fitted target objective rel {e_fit:.2e}; conditioned posterior "
            f"mean rel {e_mean:.2e}, variance rel-to-max {e_var:.2e}, rel-to-itself {e_var_self:.2e}; "
            f"cond(K_tt) {float(torch.linalg.cond(Ktt)):.2e}")
    assert e_fit <= TOL_LML and e_mean < TOL_MEAN_VAR and e_var < TOL_MEAN_VAR
    # ---- (4b) q > 1 / batch-shaped inputs: joint posterior over the q points of every batch element --------- #
    nb, q = 3, 4
    Xq = torch.rand(nb, q, d, dtype=DT, generator=g)
    th0 = O.initial_theta_raw(d, tspec)
    for mdl, args in ((prior_model, (states, w0, None, th0, tspec)), (model, (states, w, cache, th, tspec))):
        pj = mdl.posterior(Xq)
        assert pj.mean.shape == (nb, q, 1) and pj.variance.shape == (nb, q, 1)
        assert pj.mvn.covariance_matrix.shape == (nb, q, q)
        for i in range(nb):
            om_j, oc_j = O.scaml_posterior(*args, Xq[i], full_cov=True)
            assert rel_err(pj.mean[i].reshape(-1).numpy(), om_j.numpy()) < TOL_MEAN_VAR
            assert float((pj.mvn.covariance_matrix[i] - oc_j).abs().max()) < TOL_MEAN_VAR * float(oc_j.abs().max())
        # the diagonal of the joint covariance is the q = 1 variance of the same points
        p1 = mdl.posterior(Xq.reshape(-1, 1, d))
        assert float((p1.variance.reshape(nb, q) - pj.variance.reshape(nb, q)).abs().max()) < 1e-9 * vscale
        mv = mdl.forward(Xq)  # batch_shape x n x d -> one joint prior per batch element (model.py:114-121)
        assert mv.mean.shape == (nb, q) and mv.covariance_matrix.shape == (nb, q, q)
    # ---- (5) acquisition value (utils.py:215-224) ----------------------------------------------------------- #
    af = UpperConfidenceBound(model)
    assert rel_err(af(Xc.unsqueeze(1)).numpy(), O.ucb(om, ov).numpy()) < 1e-7
    # ---- (5b) analytic candidate gradients vs autograd through the oracle (what optimize_acqf differentiates) --- #
    Xg = Xc[: min(B, 40)]
    for mdl, args in ((prior_model, (states, w0, None, O.initial_theta_raw(d, tspec), tspec)),
                      (model, (states, w, cache, th, tspec))):
        gm, gv, gdm, gdv = mdl.posterior_with_grad(Xg.unsqueeze(1))
        rm, rv, rdm, rdv = O.scaml_posterior_grad(*args, Xg)
        assert gdm.shape == Xg.shape and gdv.shape == Xg.shape
        assert rel_err(gm.numpy(), rm.numpy()) < TOL_MEAN_VAR
        assert float((gv - rv).abs().max()) < TOL_MEAN_VAR * float(rv.abs().max())
        assert rel_err(gdm.numpy(), rdm.numpy()) < TOL_GRAD, rel_err(gdm.numpy(), rdm.numpy())
        assert rel_err(gdv.numpy(), rdv.numpy()) < TOL_GRAD, rel_err(gdv.numpy(), rdv.numpy())
    val, grad = af.value_and_grad(Xg)
    Xa = Xg.clone().requires_grad_(True)
    om_a, ov_a = O.scaml_posterior(states, w, cache, th, tspec, Xa)
    oa = O.ucb(om_a, ov_a)
    (og,) = torch.autograd.grad(oa.sum(), Xa)
    assert rel_err(val.numpy(), oa.detach().numpy()) < 1e-7 and rel_err(grad.numpy(), og.numpy()) < 1e-6
    # ---- (6) state_dict round trip (utils.py:169,205) ------------------------------------------------------ #
    sd = model.state_dict()
    model.weights = torch.full((M,), 0.5, dtype=DT)
    model.load_state_dict(sd)
    assert torch.equal(model.weights, w)
    # ---- (7) optimize_marginal_likelihood on ONE source GP (the reference calls it per task, model.py:187) --- #
    from scamlgp_b200.modules import set_theta_raw

    gp_last = list(gps.values())[-1]
    Xl, Yl = gp_last.train_inputs[0], gp_last._raw_Y.reshape(-1)
    ytl = O.standardize(Yl)[0]
    set_theta_raw(gp_last.likelihood, gp_last.covar_module, th0)  # back to the defaults, then refit this task alone
    sfit = optimize_marginal_likelihood(gp_last, 1, generator=torch.Generator().manual_seed(2), **(fit_options or {}))
    th_new = theta_raw_of(gp_last.likelihood, gp_last.covar_module, d)
    v_new = float(O.lml_objective(Xl, ytl, th_new, ospec))
    assert v_new >= float(O.lml_objective(Xl, ytl, th0, ospec)) - 1e-12
    assert abs(float(sfit.lml[0]) - v_new) <= TOL_LML * abs(v_new)
    st_new = O.factorize(Xl, Yl, th_new, ospec)  # the owner's packed device state was refreshed for this task
    pl = gp_last.posterior(Xc[:7])
    oml, ocl = O.posterior(st_new, Xc[:7], full_cov=True)
    assert rel_err(pl.mean.reshape(-1).numpy(), oml.numpy()) < TOL_MEAN_VAR
    assert float((pl.mvn.covariance_matrix - ocl).abs().max()) < TOL_MEAN_VAR * float(st_new.os) * st_new.ystd ** 2
    return gps, model


def test_pipeline_on_emulator(emu_engine):
    _pipeline(emu_engine, M=2, n=14, d=2, n_t=4, B=5, restarts=1, fit_options=dict(maxiter=6), ragged=True)


@pytest.mark.gpu
@pytest.mark.parametrize("M,n,d,n_t,ragged", [(8, 32, 2, 10, False), (16, 64, 6, 24, True)])
def test_pipeline_on_gpu(engine, M, n, d, n_t, ragged):
    _pipeline(engine, M=M, n=n, d=d, n_t=n_t, B=300, restarts=5, fit_options=None, ragged=ragged)


@pytest.mark.gpu
@pytest.mark.parametrize("M,n,d,seed", [(12, 64, 6, 4), (8, 32, 2, 7), (4, 128, 4, 9), (6, 48, 10, 11)])
def test_batched_lbfgs_reaches_scipy_lbfgsb_optimum(engine, M, n, d, seed):
    """The reference optimises with scipy L-BFGS-B (botorch fit_gpytorch_mll, utils.py:175); from the same
    start our lock-step L-BFGS must end at an objective at least as good (up to the ftol both use).  Shapes: the
    Hartmann-6 family of config 2, the Branin shape of config 1 (d = 2, n = 32), a longer task and a d = 10 one."""
    from scipy.optimize import minimize

    from scamlgp_b200._capi import HyperSpec
    from scamlgp_b200.engine import SourceBatch
    from scamlgp_b200.fit import fit_sources

    X, Y = O.synthetic_tasks(M, n, d, seed=seed)
    ospec = O.HyperSpec.source()
    th0 = O.initial_theta_raw(d, ospec)
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    fit = fit_sources(engine, batch, HyperSpec.source(), th0.reshape(1, 1, -1).expand(M, 1, -1).contiguous())
    for m in range(M):
        yt, _, _ = O.standardize(Y[m])

        def f(x):
            v, g = O.lml_and_grad_autograd(X[m], yt, torch.tensor(x, dtype=DT), ospec)
            return -float(v), -g.numpy()

        res = minimize(f, th0.numpy(), jac=True, method="L-BFGS-B")
        assert float(fit.lml[m]) >= -res.fun - 5e-6 * max(1.0, abs(res.fun)), (m, float(fit.lml[m]), -res.fun)


@pytest.mark.gpu
def test_meta_fit_is_deterministic_and_order_invariant(engine):
    """scamlgp/testing.py:50-100: shuffling the meta-data must not change what is learnt per task."""
    from scamlgp_b200.model import meta_fit_scamlgp

    md = _meta_data(6, 32, 3, seed=2)
    a = meta_fit_scamlgp(md, seed=3, engine=engine)
    b = meta_fit_scamlgp(md, seed=3, engine=engine)
    assert torch.equal(a.fit.theta_raw, b.fit.theta_raw)
    # warm-start row only (restart draws are positional): same optimum whatever the neighbours in the batch
    rev = dict(reversed(list(md.items())))
    c = meta_fit_scamlgp(md, num_restarts_log_likelihood=0, seed=3, engine=engine)
    e = meta_fit_scamlgp(rev, num_restarts_log_likelihood=0, seed=3, engine=engine)
    assert torch.equal(c.fit.theta_raw, e.fit.theta_raw.flip(0))
