"""Wall time of the public BO loop (ScaMLGPBO) with many meta-tasks (GPU box):
python scripts/bo_step_bench.py [M] [n] [d] [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200.engine import Engine
from scamlgp_b200.optimizer import ScaMLGPBO
from scamlgp_b200.space import ContinuousParameter, Evaluation, Objective, ParameterSpace

M, n, d, steps = (int(a) for a in (sys.argv[1:5] + ["4096", "256", "6", "6"][len(sys.argv) - 1:]))
eng = Engine(torch.device("cuda:0"))
X, Y = O.synthetic_tasks(M, n, d, seed=5)
space = ParameterSpace()
for k in range(d):
    space.add(ContinuousParameter(f"x{k}", (0.0, 1.0)))
obj = Objective("loss", False)
md = {m: [Evaluation(configuration={f"x{k}": float(X[m, i, k]) for k in range(d)}, objectives={"loss": float(Y[m, i])})
          for i in range(n)] for m in range(M)}
al = torch.tensor([1.01, 1.19, 2.9, 3.3], dtype=torch.float64)
f = lambda c: float(O.hartmann6(torch.tensor([[c[f"x{k}"] for k in range(6)]], dtype=torch.float64), al)) if d == 6 else \
    float(sum((c[f"x{k}"] - 0.3) ** 2 for k in range(d)))
torch.cuda.synchronize()
t0 = time.perf_counter()
opt = ScaMLGPBO(space, obj, md, seed=0, engine=eng)
torch.cuda.synchronize()
print(f"ScaMLGPBO.__init__ (meta-data conversion + meta-fit of {M} x {n} x {d}): {time.perf_counter()-t0:.2f} s", flush=True)
best = float("inf")
for s in range(steps):
    t0 = time.perf_counter()
    spec = opt.generate_evaluation_specification()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    y = f(spec.configuration)
    opt.report(spec.create_evaluation(objectives={"loss": y}))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    best = min(best, y)
    print(f"step {s}: suggest {t1-t0:.3f} s, report (caches + target fit) {t2-t1:.3f} s, loss {y:.4f}, best {best:.4f}",
          flush=True)
