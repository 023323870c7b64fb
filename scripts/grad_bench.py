"""Time split of the candidate-gradient path (ScaMLGP.posterior_with_grad) at many meta-tasks (GPU box).

usage: python scripts/grad_bench.py [M n d n_t B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch

M, n, d, nt, B = (int(a) for a in (sys.argv[1:6] + ["4096", "256", "6", "32", "64"][len(sys.argv) - 1:]))
from scamlgp_b200._capi import ScamlLib
lib = ScamlLib(os.environ["SCAML_LIB"]) if os.environ.get("SCAML_LIB") else None
eng = Engine(torch.device("cuda:0"), lib=lib)
dev = eng.device
X, Y = O.synthetic_tasks(M, n, d, seed=0)
batch = SourceBatch.from_padded(X.to(dev), Y.to(dev))
th = O.sample_theta_raw(M, 1, d, O.HyperSpec.source(), seed=0)[:, 0].to(dev).contiguous()
fs = eng.factorize(batch, th, HyperSpec.source())
g = torch.Generator().manual_seed(0)
Xt = torch.rand(nt, d, dtype=torch.float64, generator=g).to(dev)
Yt = torch.sin(3.0 * Xt).sum(1)
Xc = torch.rand(B, d, dtype=torch.float64, generator=g).to(dev)
w = torch.full((M,), 1.0 / M, dtype=torch.float64, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


print(f"M={M} n={n} d={d} n_t={nt} B={B}")
A = eng.cond_prepare(fs, Xt)
sm, sc = eng.cond_caches(fs, Xt, A)
Yall = torch.cat([batch.Y_raw.reshape(-1), Yt])
mu, s = float(Yall.mean()), float(Yall.std())
tht = O.initial_theta_raw(d, O.HyperSpec.target()).to(dev)
ts = eng.target_factorize(sm, sc, Xt, ((Yt - mu) / s).contiguous(), w, tht, mu, s, HyperSpec.target())
U = eng.cond_prepare(fs, Xc)
pm, pv, cross = eng.predict_conditioned(fs, w, Xc, Xt, A)
mean, var, beta = eng.target_posterior_beta(ts, pm, pv, cross, Xc)
t_u = timed(lambda: eng.cond_prepare(fs, Xc))
t_p = timed(lambda: eng.predict_conditioned(fs, w, Xc, Xt, A))
t_v = timed(lambda: eng.values_from_u(fs, w, Xc, U, Xt, A))
vm, vv, vc = eng.values_from_u(fs, w, Xc, U, Xt, A)
print(f"  values_from_u vs predict_conditioned: mean {float((vm - pm).abs().max() / pm.abs().max()):.1e} "
      f"var {float((vv - pv).abs().max() / pv.abs().max()):.1e} cross {float((vc - cross).abs().max() / cross.abs().max()):.1e}")
t_b = timed(lambda: eng.target_posterior_beta(ts, pm, pv, cross, Xc))
t_g0 = timed(lambda: eng.posterior_grad(fs, w, Xc, U))
t_g = timed(lambda: eng.posterior_grad(fs, w, Xc, U, ts, A, beta))  # consumes U (timing only: U is re-mixed every call)
print(f"  cond_prepare(Xc)  U = K^-1 k*  (DMMA)    {t_u:8.3f} ms")
print(f"  predict_conditioned (values + cross)     {t_p:8.3f} ms   (not on the gradient path any more)")
print(f"  values_from_u (mean, var, cross from U)  {t_v:8.3f} ms")
print(f"  target_posterior_beta                    {t_b:8.3f} ms")
print(f"  posterior_grad (contract + finish)       {t_g:8.3f} ms   (prior only: {t_g0:.3f} ms)")
tot = t_u + t_v + t_b + t_g
print(f"  value + gradient at {B} candidates        {tot:8.3f} ms  = {M * B / tot * 1e3 / 1e6:.1f} M (task, candidate) gradients/s")
# 2 d + 1 central finite-difference evaluations through the value path would need
t_fd = timed(lambda: eng.predict_conditioned(fs, w, Xc.repeat(2 * d + 1, 1).contiguous(), Xt, A))
print(f"  same by central differences (2d+1 value evaluations): {t_fd:8.3f} ms")
