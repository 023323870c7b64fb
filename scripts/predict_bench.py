"""Prediction kernel alone (GPU box): weighted prior prediction and the conditioned (fused cross-covariance) variant.
usage: python scripts/predict_bench.py [M n d B n_t kernel]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datagen as D
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch

M, n, d, B, nt, kern = (int(a) for a in (sys.argv[1:7] + ["4096", "256", "6", "18944", "32", "0"][len(sys.argv) - 1:]))
lib = None
if os.environ.get("SCAML_LIB"):  # A/B experiment builds (build.build_variant)
    from scamlgp_b200._capi import ScamlLib

    lib = ScamlLib(os.environ["SCAML_LIB"])
eng = Engine(torch.device("cuda:0"), lib=lib)
dev = eng.device
X, Y = D.synthetic_tasks(M, n, d, seed=0)
batch = SourceBatch.from_padded(X.to(dev), Y.to(dev))
spec = HyperSpec.source()
spec.kernel = kern
th = D.sample_theta_raw(M, 1, d, spec, seed=0)[:, 0].to(dev).contiguous()
fs = eng.factorize(batch, th, spec)
g = torch.Generator().manual_seed(0)
Xt = torch.rand(nt, d, dtype=torch.float64, generator=g).to(dev)
Xc = torch.rand(B, d, dtype=torch.float64, generator=g).to(dev)
w = torch.full((M,), 1.0 / M, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


flop_pt = n * n + n * (3 * d + 12)
ms = timed(lambda: eng.predict_weighted(fs, w, Xc))
print(f"M={M} n={n} d={d} B={B} kernel={kern}: predict_weighted {ms:8.2f} ms  {M * B / ms / 1e3:7.1f} M points/s  "
      f"{M * B * flop_pt / ms / 1e9:6.2f} TFLOP/s")
if nt > 0:
    A = eng.cond_prepare(fs, Xt)
    ms = timed(lambda: eng.predict_conditioned(fs, w, Xc, Xt, A))
    print(f"   n_t={nt}: predict_conditioned {ms:8.2f} ms  {M * B / ms / 1e3:7.1f} M points/s")
