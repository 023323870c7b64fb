"""Config-3 LML+grad launches (GPU box): median time of 5 launches, bit-reproducibility across launches.
usage: python scripts/fit_sweep.py          (SCAML_FIT_IMPL=4|8 selects the kernel variant)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch


def main():
    M, R, n, d = 4096, 6, 256, 6
    eng = Engine(torch.device("cuda:0"))
    X, Y = O.synthetic_tasks(M, n, d, seed=0)
    th = O.sample_theta_raw(M, R, d, O.HyperSpec.source(), seed=0).cuda().contiguous()
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    spec = HyperSpec.source()
    for _ in range(2):
        out = eng.lml_grad_raw(batch, th, spec)
    torch.cuda.synchronize()
    ref = out[0].clone()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.lml_grad_raw(batch, th, spec)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"impl={os.environ.get('SCAML_FIT_IMPL', 'auto')}: median {ms:7.3f} ms  {M*R/ms*1e3:9.0f} evals/s  "
          f"(min {min(ts):.3f})  bitwise_same={bool(torch.equal(ref, out[0]))}", flush=True)


if __name__ == "__main__":
    main()
