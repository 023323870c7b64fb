"""Product-level parity of the BO loop (row f4): the public drop-in (`ScaMLGPBO` on the B200 engine) against the CPU
oracle loop (oracle/bo_loop.py: scipy L-BFGS-B fits and acquisition optimisation over the oracle's torch-fp64
arithmetic, i.e. the reference's algorithm without its third-party stack) on IDENTICAL study seeds -- the same
meta-data, target task and observation-noise draws (examples/harness.py).  Regret = compute_regrets
(scamlgp/benchmarking/plotting.py:21-53).  The two arms use different optimisers for the fits, so trajectories differ;
what is compared is the regret statistics over the studies.

  python scripts/bo_parity.py --arm oracle  [--studies 64] [--procs 8]     (CPU; writes profiles/r2_bo_oracle_<bench>.json)
  python scripts/bo_parity.py --arm product [--studies 64]                 (GPU box; prints the comparison table)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import harness  # noqa: E402

AF = dict(raw_samples=512, num_restarts=16, maxiter=40)


def make_study(bench, seed):
    return harness.branin_study(seed) if bench == "branin" else harness.hartmann6_study(seed, tasks=16, points=64)


def _oracle_study(args):
    bench, seed, evals = args
    import torch

    from oracle import bo_loop

    torch.set_num_threads(1)
    st = make_study(bench, seed)
    t0 = time.perf_counter()
    vals = bo_loop.run_study(st.meta_X, st.meta_y, st.bounds, st.objective, st.noise, st.rng, evals, seed,
                             raw_samples=AF["raw_samples"], af_restarts=AF["num_restarts"], af_maxiter=AF["maxiter"])
    reg = harness.compute_regrets(False, "loss", st.optimum, [{"loss": v} for v in vals])
    return seed, reg, time.perf_counter() - t0


def table(name, R, marks):
    R = np.array(R)
    return f"  {name:28s}" + "  ".join(f"@{m}: mean {R[:, m - 1].mean():.4f} median {np.median(R[:, m - 1]):.4f}" for m in marks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["oracle", "product"], required=True)
    ap.add_argument("--bench", choices=["branin", "hartmann6"], default="branin")
    ap.add_argument("--studies", type=int, default=64)
    ap.add_argument("--evals", type=int, default=40)
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--first", type=int, default=0, help="oracle arm: first study seed (results are merged into the JSON)")
    ap.add_argument("--save", default=None, help="product arm: also write the per-study regrets (product, random search) "
                    "to this JSON; --compare PATH re-prints the table from it against the oracle JSON without a GPU")
    ap.add_argument("--compare", default=None)
    args = ap.parse_args()
    path = os.path.join(ROOT, "profiles", f"r2_bo_oracle_{args.bench}.json")
    marks = [m for m in (5, 10, 20, 40) if m <= args.evals]
    if args.arm == "oracle":
        import multiprocessing as mp

        with mp.get_context("fork").Pool(args.procs) as pool:
            res = pool.map(_oracle_study, [(args.bench, s, args.evals) for s in range(args.first, args.first + args.studies)])
        out = {"bench": args.bench, "evals": args.evals, "af": AF, "regrets": {str(s): r for s, r, _ in res},
               "seconds_per_study": float(np.mean([t for _, _, t in res]))}
        if os.path.exists(path):  # merge with the studies of earlier runs
            old = json.load(open(path))
            if old.get("evals") == args.evals and old.get("af") == AF:
                old["regrets"].update(out["regrets"])
                out["regrets"] = old["regrets"]
        with open(path, "w") as f:
            json.dump(out, f)
        print(table(f"oracle loop (CPU, {len(out['regrets'])} studies)", list(out["regrets"].values()), marks))
        print(f"  {out['seconds_per_study']:.1f} s per study (1 thread)")
        return
    ref = json.load(open(path)) if os.path.exists(path) else None
    prod, rs, secs = [], [], []
    if args.compare:
        saved = json.load(open(args.compare))
        prod, rs, secs = saved["product"], saved["random_search"], saved["seconds"]
        args.studies, args.evals = len(prod), saved["evals"]
        marks = [m for m in (5, 10, 20, 40) if m <= args.evals]
    else:
        import torch

        from scamlgp_b200.engine import Engine

        eng = Engine(torch.device("cuda:0"))
    for s in range(0 if args.compare else args.studies):
        st = make_study(args.bench, s)
        t0 = time.perf_counter()
        prod.append(harness.run_product_study(st, eng, args.evals, s, af_optimizer_kwargs=dict(AF)))
        secs.append(time.perf_counter() - t0)
        st2 = make_study(args.bench, s)  # random search on the same target task
        lo, hi = st2.bounds[:, 0], st2.bounds[:, 1]
        xs = lo + st2.rng.random((args.evals, len(lo))) * (hi - lo)
        rs.append(harness.compute_regrets(False, "loss", st2.optimum, [{"loss": st2.objective(x)} for x in xs]))
    if args.save:
        with open(args.save, "w") as f:
            json.dump({"bench": args.bench, "evals": args.evals, "af": AF, "product": [list(map(float, r)) for r in prod],
                       "random_search": [list(map(float, r)) for r in rs], "seconds": secs}, f)
    print(f"{args.bench}: {args.studies} studies x {args.evals} evaluations, identical seeds in every arm; simple regret "
          "(compute_regrets)")
    print(table("ScaMLGPBO on the B200", prod, marks))
    if ref is not None:
        common = [s for s in range(args.studies) if str(s) in ref["regrets"]]
        print(table(f"oracle loop, CPU ({len(common)} studies)", [ref["regrets"][str(s)][: args.evals] for s in common], marks))
        a = np.array([prod[s][args.evals - 1] for s in common])
        b = np.array([ref["regrets"][str(s)][args.evals - 1] for s in common])
        print(f"  final regret, product vs oracle per study: product <= oracle in {int((a <= b + 1e-12).sum())}/{len(common)}; "
              f"mean log10 ratio {np.mean(np.log10((a + 1e-6) / (b + 1e-6))):+.3f}")
        print(f"  seconds per study: product {np.mean(secs):.2f} (GPU, incl. meta-fit) vs oracle {ref['seconds_per_study']:.1f} (1 CPU thread)")
    print(table("random search", rs, marks))


if __name__ == "__main__":
    main()
