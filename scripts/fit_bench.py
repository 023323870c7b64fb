"""LML+grad throughput at an arbitrary shape (GPU box): python scripts/fit_bench.py M R n d [kernel]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch

M, R, n, d = (int(a) for a in sys.argv[1:5])
kernel = int(sys.argv[5]) if len(sys.argv) > 5 else 0
lib = None
if os.environ.get("SCAML_LIB"):  # A/B experiment builds (build.build_variant)
    from scamlgp_b200._capi import ScamlLib

    lib = ScamlLib(os.environ["SCAML_LIB"])
eng = Engine(torch.device("cuda:0"), lib=lib)
X, Y = O.synthetic_tasks(M, n, d, seed=0)
th = O.sample_theta_raw(M, R, d, O.HyperSpec.source(kernel), seed=0).cuda().contiguous()
batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
spec = HyperSpec.source(kernel)
for _ in range(2):
    out = eng.lml_grad_raw(batch, th, spec)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = eng.lml_grad_raw(batch, th, spec)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[1]
F = float(n) ** 3 + float(n) ** 2 * (2.5 * d + 10.0)
print(f"M={M} R={R} n={n} d={d} kernel={kernel}: {ms:.3f} ms  {M*R/ms*1e3:.0f} evals/s  {M*R*F/ms/1e9:.2f} TFLOP/s algorithmic  "
      f"info!=0: {int((out[2] != 0).sum())}  sum(lml)={float(out[0].sum()):.12e} sum|grad|={float(out[1].abs().sum()):.12e}", flush=True)
