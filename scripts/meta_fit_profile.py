"""Host-side profile of `meta_fit_scamlgp` at config 3 (GPU box): cProfile, top entries by cumulative time, plus the
wall clock split into before / inside / after the lock-step optimisation.

  python scripts/meta_fit_profile.py [--tasks 4096]
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datagen
from scamlgp_b200 import fit as fitmod
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

    marks = {}
    orig = fitmod.lbfgs_minimize_device

    def wrapped(*a, **k):
        torch.cuda.synchronize()
        marks["t_in"] = time.perf_counter()
        out = orig(*a, **k)
        torch.cuda.synchronize()
        marks["t_out"] = time.perf_counter()
        return out

    fitmod.lbfgs_minimize_device = wrapped
    t0 = time.perf_counter()
    meta_fit_scamlgp(md, seed=0, engine=eng)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"wall {t1 - t0:.3f} s = before the optimisation {marks['t_in'] - t0:.3f} + lock-step L-BFGS "
          f"{marks['t_out'] - marks['t_in']:.3f} + after {t1 - marks['t_out']:.3f}")
    fitmod.lbfgs_minimize_device = orig
    pr = cProfile.Profile()
    pr.enable()
    meta_fit_scamlgp(md, seed=0, engine=eng)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())


if __name__ == "__main__":
    main()
