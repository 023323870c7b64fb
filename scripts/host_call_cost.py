"""Host-side cost of the C-ABI calls on the target-fit path (GPU box): is a call asynchronous (a few us) or does it
block?  python scripts/host_call_cost.py [M] [n_t]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scamlgp_b200 import HyperSpec
from scamlgp_b200.engine import Engine

M, nt = (int(a) for a in (sys.argv[1:3] + ["4096", "6"][len(sys.argv) - 1:]))
d, R = 6, 6
eng = Engine(torch.device("cuda:0"))
dev = eng.device
g = torch.Generator().manual_seed(0)
sm = torch.randn(nt, M, dtype=torch.float64, generator=g).to(dev)
Q = torch.randn(M, nt, nt, dtype=torch.float64, generator=g)
sc = (Q @ Q.transpose(1, 2) + 0.1 * torch.eye(nt, dtype=torch.float64)).permute(1, 2, 0).contiguous().to(dev)
Xt = torch.rand(nt, d, dtype=torch.float64, generator=g).to(dev)
yt = torch.randn(nt, dtype=torch.float64, generator=g).to(dev)
w = torch.full((R, M), 1.0 / M, dtype=torch.float64, device=dev)
th = torch.zeros(R, d + 2, dtype=torch.float64, device=dev)
spec = HyperSpec.target()


def cost(name, fn, reps=200):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"  {name:42s} host {1e6 * (t1 - t0) / reps:8.1f} us/call   + drain {1e6 * (t2 - t1) / reps:8.1f} us/call")


print(f"M={M} n_t={nt} R={R}")
cost("target_lml_grad (3 launches)", lambda: eng.target_lml_grad(sm, sc, Xt, yt, w, th, 0.0, 1.0, spec))
lml = torch.empty(R, dtype=torch.float64, device=dev)
gw = torch.empty(R, M, dtype=torch.float64, device=dev)
gt = torch.empty(R, d + 2, dtype=torch.float64, device=dev)
info = torch.empty(R, dtype=torch.int32, device=dev)
need = eng.lib.target_workspace_bytes(nt, R)
ws = torch.empty(need // 8 + 8, dtype=torch.float64, device=dev)
p = lambda t: t.data_ptr()
st = torch.cuda.current_stream().cuda_stream
cost("  raw C call only (no torch.empty)", lambda: eng.lib.target_lml_grad(p(sm), p(sc), p(Xt), p(yt), p(w), p(th), None, 0.0, 1.0,
                                                                          p(lml), p(gw), p(gt), p(info), p(ws), need, M, nt, d, R, spec,
                                                                          (1, 1.0, 1.0), st))
cost("torch.empty x 4", lambda: (torch.empty(R, dtype=torch.float64, device=dev), torch.empty(R, M, dtype=torch.float64, device=dev),
                                 torch.empty(R, d + 2, dtype=torch.float64, device=dev), torch.empty(R, dtype=torch.int32, device=dev)))
X = torch.rand(64, 64, d, dtype=torch.float64, generator=g).to(dev)
thk = torch.ones(64, d + 2, dtype=torch.float64, device=dev)
cost("kernel_matrix (1 launch, reference point)", lambda: eng.kernel_matrix(X, thk, 0))
