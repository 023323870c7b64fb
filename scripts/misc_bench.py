"""Secondary measurements (GPU box): standalone kernel-matrix assembly (HBM-store bound), cross-covariance
caches, target objective, and wall time of the public meta-fit API at config 2 / config 3 shapes."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch
from scamlgp_b200.model import meta_fit_scamlgp
from scamlgp_b200.modules import SupervisedDataset

eng = Engine(torch.device("cuda:0"))
dev = eng.device


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
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


what = sys.argv[1:] or ["kmat", "cross", "metafit2", "metafit3"]
if "kmat" in what:
    for M, n, d in [(4096, 256, 6), (1024, 512, 10)]:
        X, _ = O.synthetic_tasks(M, n, d, seed=0)
        X = X.to(dev)
        theta = torch.cat([torch.full((M, d), 0.5), torch.ones(M, 1), torch.full((M, 1), 1e-3)], 1).to(dev, torch.float64)
        K = torch.empty(M, n, n, dtype=torch.float64, device=dev)
        ms = timed(lambda: eng.kernel_matrix(X, theta, 0, None, out=K))
        gb = (M * n * n * 8 + M * n * d * 8) / 1e9
        print(f"kmat M={M} n={n} d={d}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s algorithmic (stores {M*n*n*8/1e9:.2f} GB)", flush=True)
if "cross" in what:
    M, n, d, nt = 4096, 256, 6, 80
    X, Y = O.synthetic_tasks(M, n, d, seed=0)
    batch = SourceBatch.from_padded(X.to(dev), Y.to(dev))
    th = O.sample_theta_raw(M, 1, d, O.HyperSpec.source(), seed=0)[:, 0].to(dev).contiguous()
    fs = eng.factorize(batch, th, HyperSpec.source())
    Xt = torch.rand(nt, d, dtype=torch.float64, device=dev)
    ms = timed(lambda: eng.predict_cross(fs, Xt), reps=3, warm=1)
    print(f"source caches (predict_cross reduce=0) M={M} n={n} n_t={nt}: {ms:.2f} ms "
          f"({M*(n*n*nt + nt*nt*n)*2/ms/1e9:.2f} TFLOP/s algorithmic)", flush=True)
    sm, sc = eng.predict_cross(fs, Xt)
    R = 6
    w = torch.full((R, M), 1.0 / M, dtype=torch.float64, device=dev)
    tth = O.initial_theta_raw(d, O.HyperSpec.target()).to(dev).repeat(R, 1).contiguous()
    yt = torch.randn(nt, dtype=torch.float64, device=dev)
    ms = timed(lambda: eng.target_lml_grad(sm, sc, Xt, yt, w, tth, 0.0, 1.0, HyperSpec.target()))
    bytes_ = 2 * R * nt * nt * M * 8  # reduce pass + weight-gradient pass stream the covariance cache once per row
    print(f"target objective R={R} M={M} n_t={nt}: {ms:.3f} ms  {bytes_/ms/1e6:.0f} GB/s over the [n_t,n_t,M] cache", flush=True)
    del fs, sm, sc
for tag, (M, n, d) in (("metafit2", (64, 64, 6)), ("metafit3", (4096, 256, 6))):
    if tag not in what:
        continue
    X, Y = O.synthetic_tasks(M, n, d, seed=1)
    md = {i: SupervisedDataset(X[i], Y[i].reshape(-1, 1)) for i in range(M)}
    meta_fit_scamlgp({0: md[0], 1: md[1]}, seed=0, engine=eng)  # warm-up (allocations, module load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gps = meta_fit_scamlgp(md, seed=0, engine=eng)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = gps.fit.result
    evals = int(res.evaluations)
    print(f"meta_fit_scamlgp M={M} n={n} d={d} (1+5 restarts): {dt:.2f} s wall, {evals} batched objective calls, "
          f"{int(res.converged.sum())}/{res.converged.numel()} rows converged, "
          f"median iterations {int(res.iterations.median())}, launches {eng.launches}", flush=True)
