"""Per-round anatomy of `meta_fit_scamlgp` at config 3 (diagnostics; GPU box): active rows and device time of every
batched objective launch.  The instrumentation synchronises every round, so the wall clock printed here is NOT the
meta-fit time (scripts/meta_fit_bench.py measures that).

  python scripts/meta_fit_rounds.py [--tasks 4096]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datagen
from scamlgp_b200.engine import Engine
from scamlgp_b200.model import meta_fit_scamlgp
from scamlgp_b200.modules import SupervisedDataset


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=4096)
    args = ap.parse_args()
    eng = Engine(torch.device("cuda:0"))
    M, n, d = args.tasks, 256, 6
    X, Y = datagen.synthetic_tasks(M, n, d, seed=0)
    md = {i: SupervisedDataset(X[i], Y[i].reshape(-1, 1)) for i in range(M)}
    meta_fit_scamlgp({i: md[i] for i in range(64)}, seed=0, engine=eng)
    torch.cuda.synchronize()
    rounds = []
    orig = eng.lml_grad

    def wrapped(batch, theta, spec, skip=None, **kw):
        act = int((skip == 0).sum()) if skip is not None else theta.shape[0] * theta.shape[1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(batch, theta, spec, skip=skip, **kw)
        e1.record()
        torch.cuda.synchronize()
        rounds.append((act, e0.elapsed_time(e1)))
        return out

    eng.lml_grad = wrapped
    meta_fit_scamlgp(md, seed=0, engine=eng)
    tot_rows = sum(a for a, _ in rounds)
    tot_ms = sum(t for _, t in rounds)
    print(f"{len(rounds)} objective launches, {tot_rows} evaluations, {tot_ms:.1f} ms of objective kernels "
          f"({tot_rows / tot_ms:.0f} evals/ms overall)")
    print("round  active_rows  ms      evals/ms")
    for i, (a, t) in enumerate(rounds):
        print(f"{i:4d}  {a:10d}  {t:7.3f}  {a / max(t, 1e-9):8.0f}")


if __name__ == "__main__":
    main()
