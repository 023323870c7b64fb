"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python scripts/sanitize_driver.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch
from scamlgp_b200.fit import fit_sources, fit_target

eng = Engine(torch.device("cuda:0"))
dev = eng.device
for impl, (M, R, n, d, nvs) in (("4", (3, 2, 192, 3, [192, 70, 1])), ("8", (2, 1, 128, 3, [128, 65]))):
    os.environ["SCAML_FIT_IMPL"] = impl
    X, Y = O.synthetic_tasks(M, n, d, seed=1)
    batch = SourceBatch.from_padded(X.to(dev), Y.to(dev), torch.tensor(nvs, dtype=torch.int32).to(dev))
    th = O.sample_theta_raw(M, R, d, O.HyperSpec.source(), seed=1).to(dev).contiguous()
    for kern in (0, 3):
        spec = HyperSpec.source(kern)
        lml, grad, info = eng.lml_grad(batch, th, spec)
        fs = eng.factorize(batch, th[:, 0].contiguous(), spec)
        torch.cuda.synchronize()
        assert int(info.abs().max()) == 0 and int(fs.info.abs().max()) == 0
del os.environ["SCAML_FIT_IMPL"]
spec = HyperSpec.source()
fit = fit_sources(eng, batch, spec, th, dict(maxiter=5))
fs = eng.factorize(batch, fit.theta_raw, spec)
g = torch.Generator().manual_seed(0)
Xc = torch.rand(150, d, dtype=torch.float64, generator=g).to(dev)
Xt = torch.rand(7, d, dtype=torch.float64, generator=g).to(dev)
yt = torch.randn(7, dtype=torch.float64, generator=g).to(dev)
w = torch.tensor([0.6, 0.4], dtype=torch.float64, device=dev)
eng.predict_weighted(fs, w, Xc)
eng.predict_cross(fs, Xc[:40], Xt, w=w)
A = eng.cond_prepare(fs, Xt)
sm, sc = eng.cond_caches(fs, Xt, A)
pm, pv, cross = eng.predict_conditioned(fs, w, Xc, Xt, A)
K = eng.kernel_matrix(batch.X, fs.theta, 0, batch.n_valid)
tspec = HyperSpec.target()
th_t = O.initial_theta_raw(d, O.HyperSpec.target()).to(dev)
tf = fit_target(eng, sm, sc, Xt, yt, 0.1, 1.2, tspec, torch.stack([w, w * 0.5]), torch.stack([th_t, th_t + 0.1]),
                fit_options=dict(maxiter=5))
ts = eng.target_factorize(sm, sc, Xt, yt, tf.weights, tf.theta_raw, 0.1, 1.2, tspec)
mean, var = eng.target_posterior(ts, pm, pv, cross, Xc)
torch.cuda.synchronize()
assert bool(torch.isfinite(mean).all() and torch.isfinite(var).all())
print("sanitize driver ok", eng.launches, "launches")
