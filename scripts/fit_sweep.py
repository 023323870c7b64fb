"""Experiment driver (GPU box): config-3 LML+grad launches under different SCAML_FIT_STAGGER_NS values.
usage: python scripts/fit_sweep.py [ns ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch


def main():
    vals = [float(a) for a in sys.argv[1:]] or [0, 50e3, 100e3, 223e3, 330e3, 450e3]
    M, R, n, d = 4096, 6, 256, 6
    eng = Engine(torch.device("cuda:0"))
    X, Y = O.synthetic_tasks(M, n, d, seed=0)
    th = O.sample_theta_raw(M, R, d, O.HyperSpec.source(), seed=0).cuda().contiguous()
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    spec = HyperSpec.source()
    ref = None
    for ns in vals:
        os.environ["SCAML_FIT_STAGGER_NS"] = str(ns)
        for _ in range(2):
            out = eng.lml_grad_raw(batch, th, spec)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.lml_grad_raw(batch, th, spec)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        if ref is None:
            ref = out[0].clone()
        same = bool(torch.equal(ref, out[0]))
        ms = sorted(ts)[len(ts) // 2]
        print(f"stagger_ns={ns:9.0f}: median {ms:7.3f} ms  {M*R/ms*1e3:9.0f} evals/s  (min {min(ts):.3f})  bitwise_same={same}",
              flush=True)


if __name__ == "__main__":
    main()
