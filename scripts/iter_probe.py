#!/usr/bin/env python
"""DnCNN PSGLA iteration time at B chains of H x W (CUDA events, 256 MiB L2 flush between iterations not applied: back to back).
   python scripts/iter_probe.py [B H W [n]]   -- A/B runs: set the PSGLA_* switch in the environment of the process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgla_b200 as P
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 256, 256)
n = int(sys.argv[4]) if len(sys.argv) >= 5 else 40
torch.manual_seed(0)
im = torch.rand(1, 3, H, W, device="cuda")
den = P.DnCNN(pretrained=P.lipschitz_dncnn_state_dict(0))
dg, init, y, mask = P.make_inpainting(im)
s = 2 / 255
r = P.psgla_run(init, dg, den, n_iter=n + 20, n_chains=B, alpha=1.0, lambd=5.0, sig_float=s, delta=s * s, n_inter=10, n_inter_mmse=10, seed=0)
for i in range(10):
    r.step(i)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        r.step(10 + i)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / n)
sw = {k: v for k, v in os.environ.items() if k.startswith("PSGLA_")}
print("B=%d %dx%d %s: %.4f ms per iteration (best of 3 x %d), checksum %.6f" % (B, H, W, sw, best, n, r.X.double().abs().sum().item()), flush=True)
