"""Shared helpers of the parity tests (oracle = checker)."""
import numpy as np
import torch

from oracle import scaml_oracle as O
from scamlgp_b200._capi import HyperSpec

# tolerances stated by BASELINE.json north_star
TOL_LML = 1e-9
TOL_MEAN_VAR = 1e-9
TOL_GRAD = 1e-7


def specs(kernel=0, target=False):
    if target:
        return O.HyperSpec.target(kernel), HyperSpec.target(kernel)
    return O.HyperSpec.source(kernel), HyperSpec.source(kernel)


def make_problem(M, R, n, d, seed=0, n_valid=None, kernel=0):
    ospec, cspec = specs(kernel)
    X, Y = O.synthetic_tasks(M, n, d, seed=seed)
    th = O.sample_theta_raw(M, R, d, ospec, seed=seed)
    nv = np.full(M, n, dtype=np.int32) if n_valid is None else np.asarray(n_valid, dtype=np.int32)
    yt = torch.zeros(M, n, dtype=torch.float64)
    ybar = np.zeros(M)
    ystd = np.zeros(M)
    for m in range(M):
        yy, yb, ys = O.standardize(Y[m, : nv[m]])
        yt[m, : nv[m]] = yy
        ybar[m], ystd[m] = yb, ys
    return dict(X=X, Y=Y, yt=yt, th=th, nv=nv, ybar=ybar, ystd=ystd, ospec=ospec, cspec=cspec, M=M, R=R, n=n, d=d)


def oracle_lml_grad(pb, mode="direct"):
    M, R, d = pb["M"], pb["R"], pb["d"]
    v = np.zeros((M, R))
    g = np.zeros((M, R, d + 2))
    for m in range(M):
        nv = int(pb["nv"][m])
        for r in range(R):
            vv, gg = O.lml_and_grad_autograd(pb["X"][m, :nv], pb["yt"][m, :nv], pb["th"][m, r], pb["ospec"], mode=mode)
            v[m, r] = float(vv)
            g[m, r] = gg.numpy()
    return v, g


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def grad_rel_err(g, g_ref):
    """max over evaluations of ||g - g_ref||_inf / ||g_ref||_inf."""
    g = np.asarray(g).reshape(-1, g.shape[-1])
    r = np.asarray(g_ref).reshape(-1, g.shape[-1])
    return float(np.max(np.abs(g - r).max(1) / np.maximum(np.abs(r).max(1), 1e-300)))


def lml_rel_err(v, v_ref):
    v = np.asarray(v).ravel()
    r = np.asarray(v_ref).ravel()
    return float(np.max(np.abs(v - r) / np.abs(r)))
