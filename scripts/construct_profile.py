import cProfile, io, os, pstats, sys, time
import torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import datagen as D
from scamlgp_b200.engine import Engine
from scamlgp_b200.optimizer import ScaMLGPBO
from scamlgp_b200.space import ContinuousParameter, Evaluation, Objective, ParameterSpace
M, n, d = 4096, 256, 6
eng = Engine(torch.device("cuda:0"))
X, Y = D.synthetic_tasks(M, n, d, seed=5)
space = ParameterSpace()
for k in range(d):
    space.add(ContinuousParameter(f"x{k}", (0.0, 1.0)))
obj = Objective("loss", False)
t0 = time.time()
md = {m: [Evaluation(configuration={f"x{k}": float(X[m, i, k]) for k in range(d)}, objectives={"loss": float(Y[m, i])})
          for i in range(n)] for m in range(M)}
print("building meta_data dict (user side):", round(time.time() - t0, 2), "s")
for rep in range(2):
    pr = cProfile.Profile()
    torch.cuda.synchronize(); t0 = time.time()
    pr.enable()
    opt = ScaMLGPBO(space, obj, md, seed=0, engine=eng)
    torch.cuda.synchronize()
    pr.disable()
    print("construct:", round(time.time() - t0, 3), "s")
buf = io.StringIO()
pstats.Stats(pr, stream=buf).strip_dirs().sort_stats("tottime").print_stats(18)
print("\n".join(l[:140] for l in buf.getvalue().splitlines() if l.strip())[:5000])
