"""Generates tests/golden/scaml_golden_target_v1.npz from the ORACLE: the ScaML-GP target path
(reference scamlgp/model.py:219-384) on a Forrester-family scenario in the spirit of the reference's own
test data (tests/meta_data_examples.py:141-175: y = a f(x) + b (x - 0.5) - c with f the Forrester function):
3 source tasks x 12 points, 5 target points, 9 candidates.  Stored: inputs, the source hyper-parameters used,
the per-task caches, the training-branch objective + gradients at two (w, theta) rows, and the conditioned
posterior mean / variance with weight pruning (tau = 1e-3), and their gradients wrt the candidates.
Run:  python tests/golden/make_golden_target.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # noqa: E402

DT = torch.float64


def forrester(x, a, b, c):
    return a * (6 * x - 2) ** 2 * np.sin(12 * x - 4) + b * (x - 0.5) - c


def main():
    g = np.random.default_rng(42)
    desc = [(0.95, 0.02, 1.0), (0.7, -0.3, 0.2), (1.2, 0.5, -0.5)]
    M, n, nt, B, d = len(desc), 12, 5, 9, 1
    Xs = np.sort(g.random((M, n, d)), axis=1)
    Ys = np.stack([forrester(Xs[m, :, 0], *desc[m]) for m in range(M)])
    Xt = g.random((nt, d))
    Yt = forrester(Xt[:, 0], 1.0, 0.1, 0.3)
    Xc = np.linspace(0.03, 0.97, B).reshape(-1, 1)
    sspec, tspec = O.HyperSpec.source(), O.HyperSpec.target()
    th_src = O.sample_theta_raw(M, 2, d, sspec, seed=7)[:, 1]  # one prior draw per task
    states = [O.factorize(torch.tensor(Xs[m]), torch.tensor(Ys[m]), th_src[m], sspec) for m in range(M)]
    cache = O.build_target_cache(states, torch.tensor(Xt), torch.tensor(Yt))
    W = torch.tensor([[1.0 / M] * M, [0.8, 1e-7, 0.35]], dtype=DT)  # row 1: task 1 is pruned at tau = 1e-3
    TH = torch.stack([O.initial_theta_raw(d, tspec), O.sample_theta_raw(1, 2, d, tspec, seed=3)[0, 1]])
    obj, gw, gt, pm, pv = [], [], [], [], []
    for r in range(2):
        w = W[r].clone().requires_grad_(True)
        th = TH[r].clone().requires_grad_(True)
        v = O.target_objective(cache, w, th, tspec)
        v.backward()
        obj.append(float(v))
        gw.append(w.grad.numpy().copy())
        gt.append(th.grad.numpy().copy())
        m_, v_ = O.scaml_posterior(states, W[r], cache, TH[r], tspec, torch.tensor(Xc), prune_threshold=1e-3)
        pm.append(m_.numpy())
        pv.append(v_.numpy())
    # candidate gradients d mean / dx, d var / dx (autograd through the oracle: what botorch's optimize_acqf differentiates)
    dpm, dpv = [], []
    for r in range(2):
        _, _, dm_, dv_ = O.scaml_posterior_grad(states, W[r], cache, TH[r], tspec, torch.tensor(Xc), prune_threshold=1e-3)
        dpm.append(dm_.numpy())
        dpv.append(dv_.numpy())
    _, _, dprior_m, dprior_v = O.scaml_posterior_grad(states, W[0], None, TH[0], tspec, torch.tensor(Xc),
                                                      prune_threshold=1e-3)
    prior_m, prior_v = O.scaml_posterior(states, W[0], None, TH[0], tspec, torch.tensor(Xc), prune_threshold=1e-3)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scaml_golden_target_v1.npz")
    np.savez_compressed(
        out, Xs=Xs, Ys=Ys, Xt=Xt, Yt=Yt, Xc=Xc, theta_src=th_src.numpy(), W=W.numpy(), TH=TH.numpy(),
        source_means=cache.source_means.numpy(), source_covs=cache.source_covs.numpy(), mu_all=cache.mu_all,
        s_all=cache.s_all, yt_std=cache.yt_std.numpy(), objective=np.array(obj), grad_w=np.stack(gw),
        grad_theta=np.stack(gt), post_mean=np.stack(pm), post_var=np.stack(pv), prior_mean=prior_m.numpy(),
        prior_var=prior_v.numpy(), dpost_mean=np.stack(dpm), dpost_var=np.stack(dpv), dprior_mean=dprior_m.numpy(),
        dprior_var=dprior_v.numpy())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
