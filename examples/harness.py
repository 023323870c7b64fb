"""Shared pieces of the end-to-end experiments (row f4 of SURVEY 8: harness parity) -- used by the product arm
(`ScaMLGPBO` on the B200 engine, examples/*_experiment.py) and by the CPU oracle arm (oracle/bo_loop.py) so that both
see the same meta-data, the same target task and the same noise draws for a given study seed.

  compute_regrets     restates scamlgp/benchmarking/plotting.py:21-53 (running minimum of loss - optimum)
  branin_study        reference experiment scamlgp/benchmarking/configurations/branin.py:47-75, task family
                      benchmarks/branin.py:34-47 (a, b, c, r, s, t ranges; x1 in [-5,10], x2 in [0,15])
  hartmann6_study     synthetic Hartmann-6 family (benchmarks/hartmann_3d.py:31-34 alpha ranges, A / P matrices of
                      benchmarking/functions/hartmann.py:170-185)
The reference samples tasks with `parameterspace` (not installable here): the same DISTRIBUTIONS are drawn with
numpy.random.default_rng(seed), so meta-data are not bit-identical to the reference's.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Callable, Dict, List

import numpy as np


def compute_regrets(greater_is_better: bool, name: str, optimum: float, objective_values: List[dict]) -> List[float]:
    """Regret at every step of a study: running minimum of (signed objective - signed optimum); (very small) negative
    regrets are possible when the optimum was located numerically and raise a warning below -1e-6."""
    sign = -1.0 if greater_is_better else 1.0
    regrets: List[float] = []
    for ovs in objective_values:
        regret = sign * ovs[name] - sign * optimum
        if regret < -1e-6:
            warnings.warn(f"A negative regret was detected. The regret value was {regret}.", Warning)
        regrets.append(regret if not regrets else min(regret, regrets[-1]))
    return regrets


@dataclass
class Study:
    names: List[str]  # parameter names
    bounds: np.ndarray  # [d, 2]
    meta_X: List[np.ndarray]  # per task [n, d] in the ORIGINAL space
    meta_y: List[np.ndarray]  # per task [n] noisy observations
    objective: Callable[[np.ndarray], float]  # noise-free target task, original space
    optimum: float
    noise: float
    rng: np.random.Generator  # draws the observation noise of the target evaluations (after the meta-data draws)


def _branin(x1, x2, a, b, c, r, s, t):
    return a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s


def branin_study(seed: int, tasks: int = 8, points: int = 32, noise: float = 1.0) -> Study:
    rng = np.random.default_rng(seed)
    draw = lambda: dict(a=rng.uniform(0.5, 1.5), b=rng.uniform(0.1, 0.15), c=rng.uniform(1.0, 2.0),  # noqa: E731
                        r=rng.uniform(5.0, 7.0), s=rng.uniform(8.0, 12.0), t=rng.uniform(0.03, 0.05))
    mx, my = [], []
    for _ in range(tasks):
        p = draw()
        x1, x2 = rng.uniform(-5, 10, points), rng.uniform(0, 15, points)
        mx.append(np.stack([x1, x2], 1))
        my.append(_branin(x1, x2, **p) + rng.normal(0, noise, points))
    target = draw()
    g1, g2 = np.meshgrid(np.linspace(-5, 10, 600), np.linspace(0, 15, 600))
    fmin = float(_branin(g1, g2, **target).min())
    return Study(["x1", "x2"], np.array([[-5.0, 10.0], [0.0, 15.0]]), mx, my,
                 lambda x: float(_branin(x[0], x[1], **target)), fmin, noise, rng)


_A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
_P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                      [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])


def hartmann6(X, alpha):
    X = np.atleast_2d(X)
    e = np.exp(-(_A[None] * (X[:, None, :] - _P[None]) ** 2).sum(-1))
    return -(e * alpha[None]).sum(-1)


def hartmann6_study(seed: int, tasks: int = 64, points: int = 64, noise: float = 0.1) -> Study:
    from scipy.optimize import minimize

    rng = np.random.default_rng(seed)
    draw = lambda: np.array([rng.uniform(1.0, 1.02), rng.uniform(1.18, 1.2), rng.uniform(2.8, 3.0),  # noqa: E731
                             rng.uniform(3.2, 3.4)])
    mx, my = [], []
    for _ in range(tasks):
        al = draw()
        X = rng.random((points, 6))
        mx.append(X)
        my.append(hartmann6(X, al) + rng.normal(0, noise, points))
    target = draw()
    best = np.inf
    for x0 in rng.random((24, 6)):
        r = minimize(lambda v: float(hartmann6(v, target)[0]), x0, method="L-BFGS-B", bounds=[(0, 1)] * 6)
        best = min(best, r.fun)
    return Study([f"x{k}" for k in range(6)], np.array([[0.0, 1.0]] * 6), mx, my,
                 lambda x: float(hartmann6(np.asarray(x), target)[0]), float(best), noise, rng)


def run_product_study(study: Study, engine, evals: int, seed: int, af_optimizer_kwargs: Dict = None,
                      fit_options: Dict = None) -> List[float]:
    """One study through the public drop-in (`ScaMLGPBO`); returns the regret curve (compute_regrets on the noise-free
    objective values, like the reference's "(noise free)" objective, plotting.py:56-70)."""
    from scamlgp_b200.optimizer import ScaMLGPBO
    from scamlgp_b200.space import ContinuousParameter, Evaluation, Objective, ParameterSpace

    space = ParameterSpace()
    for nm, (lo, hi) in zip(study.names, study.bounds):
        space.add(ContinuousParameter(nm, (float(lo), float(hi))))
    obj = Objective("loss", False)
    md = {k: [Evaluation(configuration={nm: float(v) for nm, v in zip(study.names, x)}, objectives={"loss": float(y)})
              for x, y in zip(X, Y)] for k, (X, Y) in enumerate(zip(study.meta_X, study.meta_y))}
    opt = ScaMLGPBO(space, obj, md, seed=seed, engine=engine, af_optimizer_kwargs=af_optimizer_kwargs,
                    fit_options=fit_options)
    values = []
    for _ in range(evals):
        spec = opt.generate_evaluation_specification()
        f = study.objective(np.array([spec.configuration[nm] for nm in study.names]))
        opt.report(spec.create_evaluation(objectives={"loss": f + study.rng.normal(0, study.noise)}))
        values.append({"loss": f})
    return compute_regrets(False, "loss", study.optimum, values)
