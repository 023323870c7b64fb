"""Round-2 parity rows (VERDICT r1 "What's weak" 1 and "Next round" 2), all through the C ABI on the B200:

  * the CPU arm of bench.py evaluates the oracle in gpytorch's `expansion` distance mode at prior-sampled
    hyper-parameters -- the kernels (direct differences) are compared with exactly that here, worst gap recorded;
  * posterior variances: error relative to the variance ITSELF next to the error relative to the prior scale;
  * config 4 at its full shape (n = 512, d = 10, blocked DMMA Cholesky kernel) on 64 tasks x 2 rows;
  * config 5 at its full candidate count (1 Mi) through size-independent properties + sampled oracle rows.

Measured values are appended to gpurun_out/parity_r2.txt (when that directory exists) so that DESIGN.md section 6 can
quote them.  Reference anchors: scamlgp/utils.py:171-177 (objective), model.py:128-134,264-289,359-384 (posteriors).
"""
import os

import numpy as np
import pytest
import torch

import datagen
from oracle import scaml_oracle as O
from scamlgp_b200._capi import HyperSpec
from tests.helpers import TOL_GRAD, TOL_LML, TOL_MEAN_VAR, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(line: str) -> None:
    print(line)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_r2.txt"), "a") as f:
            f.write(line + "\n")


def test_expansion_mode_parity_at_bench_rows(engine):
    """The first 24 tasks x 6 rows of bench.py's config-3 generator (seed 0): kernels vs the oracle in BOTH distance
    modes.  `expansion` (|a|^2 - 2ab + |b|^2 on centred, length-scaled inputs, clamp_min 0: gpytorch's sq_dist,
    SURVEY A.4) is what `bench.py --impl reference` evaluates."""
    from scamlgp_b200.engine import SourceBatch

    M, R, n, d = 24, 6, 256, 6
    X, Y = datagen.synthetic_tasks(4096, n, d, seed=0)
    X, Y = X[:M], Y[:M]
    ospec, cspec = O.HyperSpec.source(), HyperSpec.source()
    th = datagen.sample_theta_raw(4096, R, d, ospec, seed=0)[:M]
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    lml, grad, info = engine.lml_grad(batch, th.cuda().contiguous(), cspec)
    assert int(info.abs().max()) == 0
    lml, grad = lml.cpu(), grad.cpu()
    worst = {"direct": [0.0, 0.0], "expansion": [0.0, 0.0]}
    for m in range(M):
        yt = O.standardize(Y[m])[0]
        for r in range(R):
            for mode in worst:
                v, g = O.lml_and_grad_autograd(X[m], yt, th[m, r], ospec, mode=mode)
                worst[mode][0] = max(worst[mode][0], abs(float(lml[m, r]) - float(v)) / abs(float(v)))
                worst[mode][1] = max(worst[mode][1], float((grad[m, r] - g).abs().max() / g.abs().max()))
    for mode, (ev, eg) in worst.items():
        record(f"bench rows (24 tasks x 6 prior-sampled rows, n=256, d=6) vs oracle[{mode}]: LML rel {ev:.2e}, grad rel {eg:.2e}")
        assert ev < TOL_LML and eg < TOL_GRAD, (mode, ev, eg)


def test_variance_error_relative_to_the_variance_itself(engine):
    """Posterior variances are differences s - |v|^2; the tests bound their error relative to the prior scale.  Here
    the error relative to the variance ITSELF is measured as well: asserted below 1e-9 wherever the variance has not
    cancelled below 1e-3 of the prior scale, and below 1e-9 * prior/variance (the same absolute bound) everywhere.
    The figure for variances down to 1e-4 of the prior scale is RECORDED, not asserted: it is the absolute error
    (4e-14 .. 8e-14 of the prior scale, both sides' rounding) times up to 1e4, measured between 6e-11 and 6e-10
    depending on the build and on the box's CPU (the oracle's own LAPACK rounding is half of it)."""
    from scamlgp_b200.engine import SourceBatch

    M, n, d, B = 16, 256, 6, 2048
    X, Y = datagen.synthetic_tasks(M, n, d, seed=31)
    ospec, cspec = O.HyperSpec.source(), HyperSpec.source()
    th = datagen.sample_theta_raw(M, 2, d, ospec, seed=31)[:, 1].contiguous()
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    fs = engine.factorize(batch, th.cuda(), cspec)
    states = [O.factorize(X[m], Y[m], th[m], ospec) for m in range(M)]
    g = torch.Generator().manual_seed(4)
    worst_rel = worst_scaled = worst_prior = 0.0
    for m in range(M):
        # half of the candidates anywhere in the cube, half right next to this task's training inputs (cancellation)
        Xc = torch.rand(B, d, dtype=torch.float64, generator=g)
        near = X[m][torch.randint(0, n, (B // 2,), generator=g)]
        Xc[B // 2:] = (near + 1e-3 * torch.randn(B // 2, d, dtype=torch.float64, generator=g)).clamp(0, 1)
        w = torch.zeros(M, dtype=torch.float64)
        w[m] = 1.0
        _, var = engine.predict_weighted(fs, w.cuda(), Xc.cuda())
        _, ov = O.posterior(states[m], Xc)
        var = var.cpu()
        prior = float(states[m].os) * states[m].ystd ** 2
        rel = (var - ov).abs() / ov
        big = ov > 1e-4 * prior
        worst_rel = max(worst_rel, float(rel[big].max()))
        worst_scaled = max(worst_scaled, float((rel * ov / prior).max()))
        worst_prior = max(worst_prior, float(((var - ov).abs() / prior).max()))
        assert float(rel[ov > 1e-3 * prior].max()) < TOL_MEAN_VAR
        assert float((rel * ov / prior).max()) < TOL_MEAN_VAR
    record(f"posterior variance (16 tasks x 2048 candidates, half of them 1e-3 from a training input): error relative "
           f"to the variance itself {worst_rel:.2e} (where var > 1e-4 prior), relative to the prior scale {worst_prior:.2e}")


def test_config4_full_shape(engine):
    """config 4 shape: 64 tasks x R = 2 x n = 512 x d = 10, blocked DMMA Cholesky (the 4-warp kernel by shape since the end of round 2; both
    variants are compared with the oracle at this n in test_gpu_parity.py::test_both_fit_kernel_variants_...)."""
    from scamlgp_b200.engine import SourceBatch

    M, R, n, d = 64, 2, 512, 10
    X, Y = datagen.synthetic_tasks(M, n, d, seed=1000)  # block 0 of bench.py's config-4 generator
    ospec, cspec = O.HyperSpec.source(), HyperSpec.source()
    th = datagen.sample_theta_raw(M, R, d, ospec, seed=1000)
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    lml, grad, info = engine.lml_grad(batch, th.cuda().contiguous(), cspec)
    assert int(info.abs().max()) == 0
    lml, grad = lml.cpu(), grad.cpu()
    ev = eg = 0.0
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    for m in range(M):
        yt = O.standardize(Y[m])[0]
        for r in range(R):
            v, g = O.lml_and_grad_autograd(X[m], yt, th[m, r], ospec)
            ev = max(ev, abs(float(lml[m, r]) - float(v)) / abs(float(v)))
            eg = max(eg, float((grad[m, r] - g).abs().max() / g.abs().max()))
    record(f"config-4 shape (64 tasks x R2 x n=512 x d=10) vs oracle: LML rel {ev:.2e}, grad rel {eg:.2e}")
    assert ev < TOL_LML and eg < TOL_GRAD, (ev, eg)
    # schedule independence: the same rows in another batch composition are bit-identical
    perm = torch.randperm(M, generator=torch.Generator().manual_seed(3))
    bp = SourceBatch.from_padded(X[perm].cuda(), Y[perm].cuda())
    l2, g2, _ = engine.lml_grad(bp, th[perm].cuda().contiguous(), cspec)
    assert torch.equal(l2.cpu(), lml[perm]) and torch.equal(g2.cpu(), grad[perm])


def test_one_mi_candidates_properties(engine):
    """config 5 at its full candidate count (B = 1 048 576, d = 6) on 32 base GPs (n = 256): candidates are
    independent (any slice predicted alone is bit-identical), the weighted sum is linear in disjoint task sets,
    variances are positive, and 256 sampled candidates match the oracle."""
    from scamlgp_b200.engine import SourceBatch

    M, n, d, B = 32, 256, 6, 1 << 20
    X, Y = datagen.synthetic_tasks(M, n, d, seed=5)
    ospec, cspec = O.HyperSpec.source(), HyperSpec.source()
    th = datagen.sample_theta_raw(M, 2, d, ospec, seed=5)[:, 1].contiguous()
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    fs = engine.factorize(batch, th.cuda(), cspec)
    g = torch.Generator().manual_seed(100)
    Xc = torch.rand(B, d, dtype=torch.float64, generator=g).cuda()
    w = torch.full((M,), 1.0 / M, dtype=torch.float64).cuda()
    mean, var = engine.predict_weighted(fs, w, Xc)
    assert bool(torch.isfinite(mean).all()) and bool((var > 0).all())
    # (1) candidate independence: a slice in the middle, predicted alone
    lo, hi = 500_000, 500_000 + 70_001
    m2, v2 = engine.predict_weighted(fs, w, Xc[lo:hi].contiguous())
    assert torch.equal(m2, mean[lo:hi]) and torch.equal(v2, var[lo:hi])
    # (2) linearity over disjoint task sets (what the [2, B] all-reduce of a task-sharded prediction relies on)
    wa, wb = w.clone(), w.clone()
    wa[M // 2:] = 0.0
    wb[: M // 2] = 0.0
    ma, va = engine.predict_weighted(fs, wa, Xc)
    mb, vb = engine.predict_weighted(fs, wb, Xc)
    assert float((ma + mb - mean).abs().max()) < 1e-13 * float(mean.abs().max())
    assert float((va + vb - var).abs().max()) < 1e-13 * float(var.abs().max())
    # (3) sampled candidates against the oracle
    idx = torch.randint(0, B, (256,), generator=torch.Generator().manual_seed(1))
    states = [O.factorize(X[m], Y[m], th[m], ospec) for m in range(M)]
    om, ov = O.scaml_prior_predict(states, w.cpu(), Xc[idx.cuda()].cpu())
    assert rel_err(mean[idx.cuda()].cpu().numpy(), om.numpy()) < TOL_MEAN_VAR
    assert rel_err(var[idx.cuda()].cpu().numpy(), ov.numpy()) < TOL_MEAN_VAR


@pytest.mark.parametrize("M,R,n,d", [(1332, 2, 256, 6), (444, 2, 512, 10)])
def test_lml_grad_is_bit_reproducible_at_full_occupancy(engine, M, R, n, d):
    """Several waves of the persistent grid with dynamic work distribution: which CTA evaluates which row, with which
    warp -> role rotation and next to which neighbours differs from launch to launch; every output bit must not
    (scripts/fit_determinism.py is the long form: 180 launches at the bench's shapes)."""
    from scamlgp_b200.engine import SourceBatch

    X, Y = datagen.synthetic_tasks(M, n, d, seed=21)
    spec = HyperSpec.source()
    th = datagen.sample_theta_raw(M, R, d, spec, seed=21).cuda().contiguous()
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    l0, g0, i0 = (t.clone() for t in engine.lml_grad_raw(batch, th, spec))
    assert int(i0.abs().max()) == 0
    for _ in range(6):
        l, g, i = engine.lml_grad_raw(batch, th, spec)
        assert torch.equal(l, l0) and torch.equal(g, g0) and torch.equal(i, i0)
