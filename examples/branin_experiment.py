"""Config 1 of BASELINE.json in spirit: the reference's Branin ScaML-GP experiment
(scamlgp/benchmarking/configurations/branin.py:47-75) end to end through `ScaMLGPBO` on the B200 engine.

  M in {8, 32} meta-tasks x 32 points, d = 2, observation noise sigma = 1.0, 40 evaluations per study.
  Task descriptors a in [0.5,1.5], b in [0.1,0.15], c in [1,2], r in [5,7], s in [8,12], t in [0.03,0.05]
  (scamlgp/benchmarking/benchmarks/branin.py:34-47); x1 in [-5,10], x2 in [0,15].
The reference samples tasks with `parameterspace` (not installable here), so the same DISTRIBUTIONS are drawn
with numpy.random.default_rng(seed); meta-data are therefore not bit-identical to the reference's.  Reports
the simple regret (best noise-free value found minus the task's global minimum on a dense grid) after
10 / 20 / 40 evaluations, averaged over the studies, next to uniform random search on the same tasks.

  python examples/branin_experiment.py [--tasks 8] [--studies 8] [--evals 40]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scamlgp_b200.engine import Engine  # noqa: E402
from scamlgp_b200.optimizer import ScaMLGPBO  # noqa: E402
from scamlgp_b200.space import ContinuousParameter, Evaluation, Objective, ParameterSpace  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from harness import compute_regrets  # noqa: E402  (restates scamlgp/benchmarking/plotting.py:21-53)


def branin(x1, x2, a, b, c, r, s, t):
    return a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s


def draw_task(rng):
    return dict(a=rng.uniform(0.5, 1.5), b=rng.uniform(0.1, 0.15), c=rng.uniform(1.0, 2.0), r=rng.uniform(5.0, 7.0),
                s=rng.uniform(8.0, 12.0), t=rng.uniform(0.03, 0.05))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=8)
    ap.add_argument("--points", type=int, default=32)
    ap.add_argument("--studies", type=int, default=8)
    ap.add_argument("--evals", type=int, default=40)
    ap.add_argument("--noise", type=float, default=1.0)
    ap.add_argument("--af-method", choices=["lbfgsb", "batched"], default="lbfgsb",
                    help="acquisition optimiser: L-BFGS-B with analytic gradients (botorch's recipe) or zeroth order")
    args = ap.parse_args()
    eng = Engine(torch.device("cuda:0"))
    space = ParameterSpace()
    space.add(ContinuousParameter("x1", (-5.0, 10.0)))
    space.add(ContinuousParameter("x2", (0.0, 15.0)))
    obj = Objective("loss", False)
    g1, g2 = np.meshgrid(np.linspace(-5, 10, 600), np.linspace(0, 15, 600))
    marks = [m for m in (10, 20, 40, 80) if m <= args.evals]
    reg, reg_rs, t_fit, t_step = [], [], [], []
    for study in range(args.studies):
        rng = np.random.default_rng(study)
        md = {}
        for k in range(args.tasks):
            p = draw_task(rng)
            x1, x2 = rng.uniform(-5, 10, args.points), rng.uniform(0, 15, args.points)
            y = branin(x1, x2, **p) + rng.normal(0, args.noise, args.points)
            md[k] = [Evaluation(configuration={"x1": float(a), "x2": float(b)}, objectives={"loss": float(v)})
                     for a, b, v in zip(x1, x2, y)]
        target = draw_task(rng)
        fmin = float(branin(g1, g2, **target).min())
        t0 = time.perf_counter()
        opt = ScaMLGPBO(space, obj, md, seed=study, engine=eng, af_optimizer_kwargs={"method": args.af_method})
        t_fit.append(time.perf_counter() - t0)
        values = []
        t0 = time.perf_counter()
        for _ in range(args.evals):
            spec = opt.generate_evaluation_specification()
            f = float(branin(spec.configuration["x1"], spec.configuration["x2"], **target))
            opt.report(spec.create_evaluation(objectives={"loss": f + rng.normal(0, args.noise)}))
            values.append({"loss": f})
        curve = compute_regrets(False, "loss", fmin, values)
        t_step.append((time.perf_counter() - t0) / args.evals)
        reg.append([curve[m - 1] for m in marks])
        xr1, xr2 = rng.uniform(-5, 10, args.evals), rng.uniform(0, 15, args.evals)
        fr = np.minimum.accumulate(branin(xr1, xr2, **target)) - fmin
        reg_rs.append([fr[m - 1] for m in marks])
        print(f"study {study}: regret " + ", ".join(f"@{m} {r:.3f}" for m, r in zip(marks, reg[-1])), flush=True)
    reg, reg_rs = np.array(reg), np.array(reg_rs)
    print(f"Branin, {args.tasks} meta-tasks x {args.points} points, noise {args.noise}, {args.studies} studies, "
          f"acquisition optimiser: {args.af_method}")
    for j, m in enumerate(marks):
        print(f"  simple regret after {m:3d} evaluations: ScaML-GP mean {reg[:, j].mean():.3f} (median "
              f"{np.median(reg[:, j]):.3f})   random search mean {reg_rs[:, j].mean():.3f}")
    print(f"  meta-fit {np.mean(t_fit):.2f} s per study, {np.mean(t_step)*1e3:.0f} ms per BO step (suggest + refit)")


if __name__ == "__main__":
    main()
