"""Run-to-run bit reproducibility of the fused LML+grad launch at full occupancy (GPU box).

  python scripts/fit_determinism.py [reps]

Evaluates the same batch `reps` times (dynamic work distribution: which CTA runs which evaluation, and with which
warp -> role rotation, differs from launch to launch) and compares every output bit for bit with the first launch.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datagen
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lib = None
if os.environ.get("SCAML_LIB"):
    from scamlgp_b200._capi import ScamlLib

    lib = ScamlLib(os.environ["SCAML_LIB"])
eng = Engine(torch.device("cuda:0"), lib=lib)
spec = HyperSpec.source()
bad_total = 0
for (M, R, n, d) in [(2048, 2, 512, 10), (4096, 6, 256, 6), (1776, 2, 384, 6)]:
    X, Y = datagen.synthetic_tasks(M, n, d, seed=0)
    th = datagen.sample_theta_raw(M, R, d, spec, seed=0).cuda().contiguous()
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    l0, g0, i0 = (t.clone() for t in eng.lml_grad_raw(batch, th, spec))
    nbad = 0
    for r in range(reps):
        l, g, i = eng.lml_grad_raw(batch, th, spec)
        dl = (l != l0) & ~(torch.isnan(l) & torch.isnan(l0))
        dg = (g != g0) & ~(torch.isnan(g) & torch.isnan(g0))
        if bool(dl.any()) or bool(dg.any()) or not torch.equal(i, i0):
            nbad += 1
            rows = torch.nonzero(dl | dg.any(-1))
            print(f"  n={n}: launch {r}: {rows.shape[0]} evaluations differ, e.g.")
            for m_, r_ in rows[:6].tolist():
                cols = torch.nonzero(dg[m_, r_]).flatten().tolist()
                print(f"    task {m_} row {r_}: lml {float(l0[m_, r_]):.17g} -> {float(l[m_, r_]):.17g}; grad entries {cols}: "
                      f"max |dg| {float((g[m_, r_] - g0[m_, r_]).abs().max()):.3e} (|g| max {float(g0[m_, r_].abs().max()):.3e})")
    print(f"n={n} d={d} M={M} R={R}: {reps} launches, {nbad} differ from the first", flush=True)
    bad_total += nbad
print("DETERMINISTIC" if bad_total == 0 else f"NON-DETERMINISTIC: {bad_total} launches", flush=True)
