"""Timing-only ablations of the fit kernel (GPU box; results are WRONG by design).
bits: 1 skip inverse sweeps, 32 skip cholesky sweeps, 2 cheap exp (assemble), 4 cheap exp (grad),
      8 skip DMMA in the streamed products, 16 skip cp.async staging."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O
from scamlgp_b200 import HyperSpec, build
from scamlgp_b200._capi import ScamlLib
from scamlgp_b200.engine import Engine, SourceBatch

lib = ScamlLib(build.build_ablate())
lib.lib.scaml_debug_set_ablate.argtypes = [C.c_int]
eng = Engine(torch.device("cuda:0"), lib=lib)
M, R, n, d = 4096, 6, 256, 6
X, Y = O.synthetic_tasks(M, n, d, seed=0)
th = O.sample_theta_raw(M, R, d, O.HyperSpec.source(), seed=0).cuda().contiguous()
batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
spec = HyperSpec.source()
cases = [int(a) for a in sys.argv[1:]] or [0, 1, 33, 2, 4, 6, 8, 16, 24, 39, 63]
for bits in cases:
    lib.lib.scaml_debug_set_ablate(bits)
    for _ in range(2):
        eng.lml_grad_raw(batch, th, spec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.lml_grad_raw(batch, th, spec)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"ablate bits={bits:3d}: {min(ts):7.3f} ms", flush=True)
