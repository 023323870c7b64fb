"""Time split of ScaMLGP.posterior (n_t > 0) at many meta-tasks (GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch

M, n, d, nt, B = (int(a) for a in (sys.argv[1:6] + ["4096", "256", "6", "20", "4096"][len(sys.argv) - 1:]))
eng = Engine(torch.device("cuda:0"))
dev = eng.device
X, Y = O.synthetic_tasks(M, n, d, seed=0)
batch = SourceBatch.from_padded(X.to(dev), Y.to(dev))
th = O.sample_theta_raw(M, 1, d, O.HyperSpec.source(), seed=0)[:, 0].to(dev).contiguous()
fs = eng.factorize(batch, th, HyperSpec.source())
g = torch.Generator().manual_seed(0)
Xt = torch.rand(nt, d, dtype=torch.float64, generator=g).to(dev)
Xc = torch.rand(B, d, dtype=torch.float64, generator=g).to(dev)
w = torch.full((M,), 1.0 / M, dtype=torch.float64, device=dev)


def timed(fn, reps=3):
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
print(f"  predict_weighted(B)              {timed(lambda: eng.predict_weighted(fs, w, Xc)):8.2f} ms")
print(f"  predict_cross(Xc, Xt, w) [B,n_t] {timed(lambda: eng.predict_cross(fs, Xc, Xt, w=w)):8.2f} ms")
A = eng.cond_prepare(fs, Xt)
print(f"  cond_prepare(Xt) (A_m, once/report){timed(lambda: eng.cond_prepare(fs, Xt)):7.2f} ms")
print(f"  predict_conditioned (fused)      {timed(lambda: eng.predict_conditioned(fs, w, Xc, Xt, A)):8.2f} ms")
_, c0 = eng.predict_cross(fs, Xc, Xt, w=w)
_, _, c1 = eng.predict_conditioned(fs, w, Xc, Xt, A)
print(f"  max |cross_fused - cross_kernel| / max|cross| = {float((c1 - c0).abs().max() / c0.abs().max()):.2e}")
print(f"  predict_cross(Xt) caches         {timed(lambda: eng.predict_cross(fs, Xt)):8.2f} ms")
