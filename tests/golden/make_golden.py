"""Generates tests/golden/scaml_golden_v1.npz from the ORACLE (oracle/scaml_oracle.py).

The reference itself cannot be imported in this image (botorch/gpytorch/blackboxopt are
absent), and its tests pin no numerical GP outputs, so these vectors are
produced by the oracle after it has been pinned against sklearn / mpmath / autograd
(tests/test_oracle.py).  Inputs reuse the reference's own fixtures where they exist:
  * META_DATA_1D               scamlgp/testing.py:18-28   (x0 in [0.5, 3] mapped to [0,1])
  * META_DATA_2D_SPACE         tests/meta_data_examples.py:50-85 (2 points per task)
  * Forrester family, 32 pts   tests/meta_data_examples.py:141-175 (a=0.95,b=0.02,c=1)
  * Hartmann-6 family          scamlgp/benchmarking/functions/hartmann.py:170-185
  * Branin family              scamlgp/benchmarking/functions/branin.py:9-42
Run:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # noqa: E402

DT = torch.float64


def branin(x1, x2, a, b, c, r, s, t):
    return a * (x2 - b * x1**2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s


def cases():
    out = {}
    # 1) META_DATA_1D (7 points, d=1)
    x0 = np.array([0.8, 1.49, 1.56, 2.5, 3.0, 1.2, 2.7])
    y0 = np.array([-6.07, -18.6, -19.9, -33.2, -29.2, -31.1, -30.2])
    order = np.argsort(x0)  # the reference sorts evaluations before conversion (utils.py:99)
    out["meta1d"] = [((x0[order] - 0.5) / 2.5).reshape(-1, 1), y0[order]]
    # 2) META_DATA_2D_SPACE: two tasks x two points (x0 in [-3,3]?, the reference tests use
    #    blackboxopt's spaces; we map with bounds x0:[-3,3], x1:[-1,0])
    t1 = np.array([[2.5, -0.1], [-2.5, -0.2]])
    t2 = np.array([[1.0, -0.3], [2.0, -0.2]])
    sc = lambda t: np.stack([(t[:, 0] + 3) / 6, (t[:, 1] + 1)], 1)
    out["meta2d_t1"] = [sc(t1), np.array([1.0, 0.0])]
    out["meta2d_t2"] = [sc(t2), np.array([1.0, 1.0])]
    # 3) single-point task (n = 1): Standardize falls back to std = 1
    out["single"] = [np.array([[0.02]]), np.array([-4.07])]
    # 4) Forrester 32 points
    g = np.random.default_rng(0)
    xf = np.sort(g.random(32))
    yf = 0.95 * ((6 * xf - 2) ** 2 * np.sin(12 * xf - 4)) + 0.02 * xf + 1
    out["forrester32"] = [xf.reshape(-1, 1), yf]
    # 5) Hartmann-6, n = 64 and n = 100 (ragged vs the 64 grid)
    X, Y = O.synthetic_tasks(2, 100, 6, seed=11)
    out["hartmann6_n64"] = [X[0, :64].numpy(), Y[0, :64].numpy()]
    out["hartmann6_n100"] = [X[1].numpy(), Y[1].numpy()]
    # 6) Branin n = 32, d = 2, noise sd 1
    xb = g.random((32, 2))
    yb = branin(-5 + 15 * xb[:, 0], 15 * xb[:, 1], 1.1, 0.12, 1.5, 6.0, 10.0, 0.04) + g.normal(0, 1.0, 32)
    out["branin32"] = [xb, yb]
    return out


def main():
    data = {}
    names = []
    for name, (X, Y) in cases().items():
        X = torch.tensor(np.asarray(X), dtype=DT)
        Y = torch.tensor(np.asarray(Y), dtype=DT)
        n, d = X.shape
        for kern in (O.KERNEL_RBF, O.KERNEL_MATERN52):
            spec = O.HyperSpec.source(kern)
            th = O.sample_theta_raw(1, 3, d, spec, seed=len(names))[0]
            yt, ybar, ystd = O.standardize(Y)
            vals, grads = [], []
            for r in range(3):
                v, gr = O.lml_and_grad_autograd(X, yt, th[r], spec)
                vals.append(float(v))
                grads.append(gr.numpy())
            st = O.factorize(X, Y, th[1], spec)
            Xs = torch.rand(9, d, dtype=DT, generator=torch.Generator().manual_seed(3))
            mu, var = O.posterior(st, Xs)
            key = f"{name}__k{kern}"
            names.append(key)
            data[key + "__X"] = X.numpy()
            data[key + "__Y"] = Y.numpy()
            data[key + "__theta_raw"] = th.numpy()
            data[key + "__lml"] = np.array(vals)
            data[key + "__grad"] = np.stack(grads)
            data[key + "__Xs"] = Xs.numpy()
            data[key + "__post_mean"] = mu.numpy()
            data[key + "__post_var"] = var.numpy()
            data[key + "__alpha"] = st.alpha.numpy()
            data[key + "__ybar_ystd"] = np.array([ybar, ystd])
    data["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scaml_golden_v1.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, len(names), "cases")


if __name__ == "__main__":
    main()
