"""One process = one evaluation of a fixed batch; prints hashes of the inputs (host generator, device standardisation)
and of the outputs, so that runs of separate processes can be compared (scripts/fit_determinism.py compares launches
inside one process).   for i in $(seq 12); do python scripts/fit_determinism_proc.py; done | sort | uniq -c"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datagen
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch


def h(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:10]


M, R, n, d = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (2048, 2, 512, 10)
eng = Engine(torch.device("cuda:0"))
spec = HyperSpec.source()
X, Y = datagen.synthetic_tasks(M, n, d, seed=0)
th = datagen.sample_theta_raw(M, R, d, spec, seed=0)
batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
thd = th.cuda().contiguous()
outs = []
for _ in range(4):
    l, g, i = eng.lml_grad_raw(batch, thd, spec)
    outs.append((h(l), h(g)))
print(f"n={n} X {h(X)} Y {h(Y)} theta {h(th)} y_std {h(batch.y)} ybar {h(batch.ybar)} | lml/grad of 4 launches:",
      " ".join(f"{a}/{b}" for a, b in outs), f"| sum(lml)={float(l.sum()):.12e}")
