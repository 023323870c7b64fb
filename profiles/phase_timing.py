"""Per-phase cycle breakdown of scaml_fit_kernel (diagnostics build -DSCAML_PROF, clock64 per CTA).
usage (GPU box): python profiles/phase_timing.py [M] [R] [n] [d]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scaml_oracle as O  # input generator only
from scamlgp_b200 import HyperSpec, build
from scamlgp_b200._capi import ScamlLib
from scamlgp_b200.engine import Engine, SourceBatch

NAMES = ["setup", "B gemm(update)", "B assemble(exp)+store", "B diag_factor", "B trsm+store", "C gemm1", "C gemm2+store",
         "C z solve", "D gemm", "D grad epilogue", "final", " diag: chol#1 (warp0)", " diag: L10,D11,Tm", " diag: chol#2 rest (outputs)", " chain warp: 32 chol steps (x8)", " follower warp: 32 inverse steps incl. waits (x8)"]


def main():
    M, R, n, d = (int(a) for a in (sys.argv[1:5] + ["1184", "2", "256", "6"][len(sys.argv) - 1:]))
    lib = ScamlLib(os.environ["SCAML_LIB"] if os.environ.get("SCAML_LIB") else build.build_prof())
    import ctypes as C

    eng = Engine(torch.device("cuda:0"), lib=lib)
    prof = torch.zeros(148 * 4 * 16, dtype=torch.int64, device="cuda")
    lib.lib.scaml_debug_set_prof.argtypes = [C.c_void_p]
    lib.lib.scaml_debug_set_prof(prof.data_ptr())
    X, Y = O.synthetic_tasks(M, n, d, seed=0)
    th = O.sample_theta_raw(M, R, d, O.HyperSpec.source(), seed=0).cuda().contiguous()
    batch = SourceBatch.from_padded(X.cuda(), Y.cuda())
    for _ in range(2):
        eng.lml_grad_raw(batch, th, HyperSpec.source())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.lml_grad_raw(batch, th, HyperSpec.source())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    p = prof.view(-1, 16).cpu().double()
    grid = min(M * R, 148 * int(os.environ.get('SCAML_FIT_CTAS_PER_SM', '3')))
    p = p[:grid]
    evals_per_cta = M * R / grid
    tot = p[:, :14].sum(1).mean().item()  # entries 14, 15 are sub-intervals measured on the chain warp
    print(f"M={M} R={R} n={n} d={d}: {ms:.3f} ms, {M*R/ms*1e3:.0f} evals/s, grid={grid}, "
          f"cycles/CTA={tot:.0f} ({tot/evals_per_cta:.0f} per eval per CTA)")
    for i, nm in enumerate(NAMES[:16]):
        c = p[:, i].mean().item()
        print(f"  {nm:26s} {c/evals_per_cta:12.0f} cyc/eval  {100*c/tot:6.2f}%")


if __name__ == "__main__":
    main()
