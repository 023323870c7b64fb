"""cProfile of ScaMLGPBO.report / generate_evaluation_specification with many meta-tasks (GPU box):
python scripts/bo_step_profile.py [M] [n] [d] [steps]"""
import cProfile
import io
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200.engine import Engine
from scamlgp_b200.optimizer import ScaMLGPBO
from scamlgp_b200.space import ContinuousParameter, Evaluation, Objective, ParameterSpace

M, n, d, steps = (int(a) for a in (sys.argv[1:5] + ["4096", "256", "6", "4"][len(sys.argv) - 1:]))
eng = Engine(torch.device("cuda:0"))
X, Y = O.synthetic_tasks(M, n, d, seed=5)
space = ParameterSpace()
for k in range(d):
    space.add(ContinuousParameter(f"x{k}", (0.0, 1.0)))
obj = Objective("loss", False)
md = {m: [Evaluation(configuration={f"x{k}": float(X[m, i, k]) for k in range(d)}, objectives={"loss": float(Y[m, i])})
          for i in range(n)] for m in range(M)}
opt = ScaMLGPBO(space, obj, md, seed=0, engine=eng)
f = lambda c: float(sum((c[f"x{k}"] - 0.3) ** 2 for k in range(d)))
for s in range(2):  # warm-up
    spec = opt.generate_evaluation_specification()
    opt.report(spec.create_evaluation(objectives={"loss": f(spec.configuration)}))
for name in ("report", "suggest"):
    pr = cProfile.Profile()
    for s in range(steps):
        if name == "suggest":
            pr.enable()
        spec = opt.generate_evaluation_specification()
        torch.cuda.synchronize()
        if name == "suggest":
            pr.disable()
        ev = spec.create_evaluation(objectives={"loss": f(spec.configuration)})
        if name == "report":
            pr.enable()
        opt.report(ev)
        torch.cuda.synchronize()
        if name == "report":
            pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).strip_dirs().sort_stats("tottime").print_stats(22)
    print(f"===== {name} x {steps} =====")
    print("\n".join(l[:150] for l in buf.getvalue().splitlines() if l.strip())[:6000])
