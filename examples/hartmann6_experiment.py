"""Config 2 of BASELINE.json end to end: synthetic Hartmann-6 meta-data (64 tasks x 64 points, d = 6, noise 0.1) through
`ScaMLGPBO` on the B200 engine.  Task family as in the reference (scamlgp/benchmarking/benchmarks/hartmann_3d.py:31-34
ranges for the alpha coefficients, A / P matrices of benchmarking/functions/hartmann.py:170-185; closed form restated here).
Reports the simple regret (best noise-free value found minus the task's minimum, located by multi-start L-BFGS-B on the
closed form) after 10 / 20 / 40 evaluations, averaged over the studies, next to uniform random search, for both
acquisition optimisers (L-BFGS-B with the analytic posterior gradients / zeroth-order batched search).

  python examples/hartmann6_experiment.py [--tasks 64] [--points 64] [--studies 4] [--evals 40]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scamlgp_b200.engine import Engine  # noqa: E402
from scamlgp_b200.optimizer import ScaMLGPBO  # noqa: E402
from scamlgp_b200.space import ContinuousParameter, Evaluation, Objective, ParameterSpace  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from harness import compute_regrets  # noqa: E402  (restates scamlgp/benchmarking/plotting.py:21-53)

A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                     [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])


def hartmann6(X, alpha):
    X = np.atleast_2d(X)
    e = np.exp(-(A[None] * (X[:, None, :] - P[None]) ** 2).sum(-1))
    return -(e * alpha[None]).sum(-1)


def draw_alpha(rng):
    return np.array([rng.uniform(1.0, 1.02), rng.uniform(1.18, 1.2), rng.uniform(2.8, 3.0), rng.uniform(3.2, 3.4)])


def global_min(alpha, rng):
    best = np.inf
    for x0 in rng.random((24, 6)):
        r = minimize(lambda v: float(hartmann6(v, alpha)[0]), x0, method="L-BFGS-B", bounds=[(0, 1)] * 6)
        best = min(best, r.fun)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=64)
    ap.add_argument("--points", type=int, default=64)
    ap.add_argument("--studies", type=int, default=4)
    ap.add_argument("--evals", type=int, default=40)
    ap.add_argument("--noise", type=float, default=0.1)
    ap.add_argument("--methods", default="lbfgsb,batched", help="acquisition optimisers to run (comma-separated)")
    args = ap.parse_args()
    eng = Engine(torch.device("cuda:0"))
    space = ParameterSpace()
    for k in range(6):
        space.add(ContinuousParameter(f"x{k}", (0.0, 1.0)))
    obj = Objective("loss", False)
    marks = [m for m in (10, 20, 40, 80) if m <= args.evals]
    print(f"Hartmann-6, {args.tasks} meta-tasks x {args.points} points, noise {args.noise}, {args.studies} studies")
    for method in args.methods.split(","):
        reg, reg_rs, t_fit, t_step = [], [], [], []
        for study in range(args.studies):
            rng = np.random.default_rng(study)
            md = {}
            for k in range(args.tasks):
                al = draw_alpha(rng)
                X = rng.random((args.points, 6))
                y = hartmann6(X, al) + rng.normal(0, args.noise, args.points)
                md[k] = [Evaluation(configuration={f"x{j}": float(X[i, j]) for j in range(6)},
                                    objectives={"loss": float(y[i])}) for i in range(args.points)]
            target = draw_alpha(rng)
            fmin = global_min(target, rng)
            t0 = time.perf_counter()
            opt = ScaMLGPBO(space, obj, md, seed=study, engine=eng, af_optimizer_kwargs={"method": method})
            t_fit.append(time.perf_counter() - t0)
            values = []
            t0 = time.perf_counter()
            for _ in range(args.evals):
                spec = opt.generate_evaluation_specification()
                x = np.array([spec.configuration[f"x{j}"] for j in range(6)])
                f = float(hartmann6(x, target)[0])
                opt.report(spec.create_evaluation(objectives={"loss": f + rng.normal(0, args.noise)}))
                values.append({"loss": f})
            curve = compute_regrets(False, "loss", fmin, values)
            t_step.append((time.perf_counter() - t0) / args.evals)
            reg.append([curve[m - 1] for m in marks])
            fr = np.minimum.accumulate(hartmann6(rng.random((args.evals, 6)), target)) - fmin
            reg_rs.append([fr[m - 1] for m in marks])
        reg, reg_rs = np.array(reg), np.array(reg_rs)
        print(f" acquisition optimiser: {method}")
        for j, m in enumerate(marks):
            print(f"  simple regret after {m:3d} evaluations: ScaML-GP mean {reg[:, j].mean():.3f} (median "
                  f"{np.median(reg[:, j]):.3f})   random search mean {reg_rs[:, j].mean():.3f}")
        print(f"  meta-fit {np.mean(t_fit):.2f} s per study, {np.mean(t_step) * 1e3:.0f} ms per BO step (suggest + refit)")


if __name__ == "__main__":
    main()
