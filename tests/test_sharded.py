"""Task sharding over ranks (scamlgp_b200/sharded.py): world_size-2 `gloo` processes on CPU, each driving the
emulation engine (tests only), must reproduce the single-process result over all tasks -- fitted rows,
weighted prior prediction, per-task caches and the conditioned posterior."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = torch.float64
M, N, D, NT, B = 3, 12, 2, 3, 4  # 3 tasks over 2 ranks: blocks of 2 and 1 (uneven on purpose)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tasks():
    from oracle import scaml_oracle as O

    X, Y = O.synthetic_tasks(M, N, D, seed=21)
    return [(X[i, : N - 2 * i], Y[i, : N - 2 * i]) for i in range(M)]  # ragged


def _run(engine, group):
    from oracle import scaml_oracle as O
    from scamlgp_b200._capi import HyperSpec
    from scamlgp_b200.sharded import ShardedSources

    g = torch.Generator().manual_seed(9)
    spec, tspec = HyperSpec.source(), HyperSpec.target()
    th0 = O.sample_theta_raw(M, 2, D, O.HyperSpec.source(), seed=2)
    src = ShardedSources(engine, _tasks(), group=group)
    fit = src.fit(spec, th0, fit_options=dict(maxiter=4))
    w = torch.tensor([0.5, 0.3, 0.2], dtype=DT)
    Xc = torch.rand(B, D, dtype=DT, generator=g)
    Xt = torch.rand(NT, D, dtype=DT, generator=g)
    yt = torch.randn(NT, dtype=DT, generator=g)
    pm, pv = src.predict_weighted(w, Xc)
    sm, sc = src.target_caches(Xt)
    ts = engine.target_factorize(sm, sc, Xt, yt, w, O.initial_theta_raw(D, O.HyperSpec.target()), 0.1, 1.3, tspec)
    mean, var = src.posterior(w, Xc, ts)
    # analytic candidate gradients: local contraction per rank, target-kernel terms on rank 0, one all_reduce
    gmean, gvar, dm, dv = src.posterior_with_grad(w, Xc, ts)
    _, _, dm0, dv0 = src.posterior_with_grad(w, Xc, None, prior_outputscale=0.1)
    return dict(theta=fit.theta_raw, lml=fit.lml, ystd=src.ystd_all, pm=pm.clone(), pv=pv.clone(), sm=sm, sc=sc,
                mean=mean, var=var, gmean=gmean, gvar=gvar, dm=dm, dv=dv, dm0=dm0, dv0=dv0)


def _bo_run(engine, group):
    """The public drop-in on sharded meta-tasks: `ScaMLGPBO(..., group=)` -- meta-fit (optimizer.py:128-133), report
    (:176-185: caches + target refit) and suggestions (eval branch, model.py:364-375) with 3 tasks over the ranks."""
    from scamlgp_b200.optimizer import ScaMLGPBO
    from scamlgp_b200.space import ContinuousParameter, Evaluation, Objective, ParameterSpace

    space = ParameterSpace()
    space.add(ContinuousParameter("x0", (0.0, 1.0)))
    space.add(ContinuousParameter("x1", (0.0, 1.0)))
    loss = Objective("loss", greater_is_better=False)
    md = {f"t{i}": [Evaluation(configuration={"x0": float(x[0]), "x1": float(x[1])}, objectives={"loss": float(y)})
                    for x, y in zip(X, Y)] for i, (X, Y) in enumerate(_tasks())}
    opt = ScaMLGPBO(space, loss, md, seed=5, num_initial_random_samples=0, max_pending_evaluations=3, engine=engine,
                    num_restarts_log_likelihood=1, fit_options=dict(maxiter=4), group=group,
                    af_optimizer_kwargs=dict(raw_samples=16, num_restarts=2, maxiter=4))
    out = []
    for step in range(3):
        spec = opt.generate_evaluation_specification()  # step 0: prior-only model (n_t = 0)
        c = spec.configuration
        out.append([c["x0"], c["x1"]])
        opt.report(spec.create_evaluation(objectives={"loss": (c["x0"] - 0.3) ** 2 + (c["x1"] - 0.6) ** 2}))
    post = opt.model.posterior(torch.tensor([[0.2, 0.4], [0.7, 0.1]], dtype=DT))
    joint = opt.model.posterior(torch.tensor([[[0.2, 0.4], [0.7, 0.1], [0.5, 0.5]]], dtype=DT))  # q = 3
    theta = torch.stack([torch.cat([g.covar_module.base_kernel.raw_lengthscale.reshape(-1),
                                    g.covar_module.raw_outputscale.reshape(-1)]) for g in opt.source_gps.values()])
    return dict(bo_x=torch.tensor(out, dtype=DT), bo_w=opt.model.weights.clone(), bo_mean=post.mean.reshape(-1),
                bo_var=post.variance.reshape(-1), bo_jmean=joint.mean.reshape(-1),
                bo_jcov=joint.mvn.covariance_matrix.reshape(-1), bo_theta=theta)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from scamlgp_b200._capi import ScamlLib
    from tests.emu_build import build_emu
    from tests.emu_engine import EmuEngine

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = EmuEngine(ScamlLib(build_emu()))
        res = _run(eng, None)
        res.update(_bo_run(eng, True))
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_task_partition():
    from scamlgp_b200.sharded import task_partition

    assert task_partition(4096, 8) == [(i * 512, (i + 1) * 512) for i in range(8)]
    assert task_partition(3, 2) == [(0, 2), (2, 3)]
    assert task_partition(5, 4) == [(0, 2), (2, 4), (4, 5), (5, 5)]


def test_two_rank_gloo_matches_single_process(emu_lib, tmp_path):
    from tests.emu_engine import EmuEngine

    single = _run(EmuEngine(emu_lib), None)
    single.update(_bo_run(EmuEngine(emu_lib), None))  # group=None: the un-sharded product path
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    # the value half of posterior_with_grad comes from U (no second pass over the factors): same numbers to rounding
    assert float((single["gmean"] - single["mean"]).abs().max()) <= 1e-11 * float(single["mean"].abs().max())
    assert float((single["gvar"] - single["var"]).abs().max()) <= 1e-10 * float(single["var"].abs().max())
    for k in single:
        # every rank ends with the same replicated result ...
        assert torch.equal(r0[k], r1[k]), k
        # ... equal to the single-process one: bit-identical where nothing is re-associated across ranks
        if k in ("theta", "lml", "ystd", "sm", "sc", "bo_theta"):
            assert torch.equal(r0[k], single[k]), k
        elif k.startswith("bo_"):
            # a 2-rank ScaMLGPBO proposes the same configurations as one rank: the sums over tasks are
            # re-associated across ranks (1e-13), the optimisers on top of them amplify that only mildly
            scale = float(single[k].abs().max())
            assert float((r0[k] - single[k]).abs().max()) <= 1e-6 * scale, (k, r0[k], single[k])
        else:
            scale = float(single[k].abs().max())
            assert float((r0[k] - single[k]).abs().max()) <= 1e-13 * scale, k
