"""Multi-GPU check of scamlgp_b200.sharded (run under torchrun, one rank per GPU, NCCL):
every rank fits its block of the tasks, the gathered rows and the all-reduced weighted prediction must equal
the single-GPU result computed on rank 0 over all tasks.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine, SourceBatch
from scamlgp_b200.fit import fit_sources
from scamlgp_b200.sharded import ShardedSources


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    eng = Engine(dev)
    M, n, d, R, B, nt = 37, 96, 6, 3, 5000, 12
    X, Y = O.synthetic_tasks(M, n, d, seed=3)
    tasks = [(X[i, : n - (i % 5)], Y[i, : n - (i % 5)]) for i in range(M)]
    th0 = O.sample_theta_raw(M, R, d, O.HyperSpec.source(), seed=1)
    spec, tspec = HyperSpec.source(), HyperSpec.target()
    src = ShardedSources(eng, tasks)
    fit = src.fit(spec, th0)
    g = torch.Generator().manual_seed(0)
    w = torch.rand(M, dtype=torch.float64, generator=g)
    Xc = torch.rand(B, d, dtype=torch.float64, generator=g)
    Xt = torch.rand(nt, d, dtype=torch.float64, generator=g).to(dev)
    yt = torch.randn(nt, dtype=torch.float64, generator=g).to(dev)
    pm, pv = src.predict_weighted(w, Xc)
    sm, sc = src.target_caches(Xt)
    ts = eng.target_factorize(sm, sc, Xt, yt, w.to(dev), O.initial_theta_raw(d, O.HyperSpec.target()).to(dev), 0.2, 1.1,
                              tspec)
    mean, var = src.posterior(w, Xc, ts)
    # analytic candidate gradients over the task shards: local contraction, target-kernel terms on rank 0, one all_reduce
    Xg = Xc[:100]
    gmean, gvar, gdm, gdv = src.posterior_with_grad(w, Xg, ts)
    ok = True
    if rank == 0:
        batch = SourceBatch.from_ragged(tasks, dev)
        ref = fit_sources(eng, batch, spec, th0)
        fs = eng.factorize(batch, ref.theta_raw, spec)
        rm, rv = eng.predict_weighted(fs, w.to(dev), Xc.to(dev))
        rsm, rsc = eng.cond_caches(fs, Xt, eng.cond_prepare(fs, Xt))
        wd, Xgd = w.to(dev), Xg.to(dev).contiguous()
        A1 = eng.cond_prepare(fs, Xt)
        U1 = eng.cond_prepare(fs, Xgd, wd)
        a1, b1, c1 = eng.prior_values(fs, wd, Xgd, U1, Xt, A1)
        m1, v1, beta1 = eng.target_posterior_beta(ts, a1, b1, c1, Xgd)
        dm1, dv1 = eng.posterior_grad(fs, wd, Xgd, U1, ts, A1, beta1)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        checks = {
            "sharded grad: values 1e-11": rel(gmean, m1) < 1e-11 and rel(gvar, v1) < 1e-10,
            "sharded grad: dmean 1e-11": rel(gdm, dm1) < 1e-11,
            "sharded grad: dvar 1e-10": rel(gdv, dv1) < 1e-10,
            "theta bitwise": torch.equal(ref.theta_raw, fit.theta_raw),
            "lml bitwise": torch.equal(ref.lml, fit.lml),
            "caches bitwise": torch.equal(rsm, sm) and torch.equal(rsc, sc),
            "prior mean 1e-13": float((rm - pm).abs().max() / rm.abs().max()) < 1e-13,
            "prior var 1e-13": float((rv - pv).abs().max() / rv.abs().max()) < 1e-13,
            "posterior finite": bool(torch.isfinite(mean).all() and torch.isfinite(var).all() and (var > 0).all()),
        }
        ok = all(checks.values())
        print(f"world={world}", checks, "OK" if ok else "FAILED", flush=True)
    flag = torch.tensor([int(ok)], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
