#!/usr/bin/env python
"""Single-chain DnCNN PSGLA iteration time against network depth: the slope is the period of one fused layer-pair launch
(conv_fused2.cu), the intercept what the first layer, the last layer (+ Langevin post / next pre) and the launches cost.
   python scripts/depth_probe.py [B H W]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgla_b200 as P
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1, 256, 256)
torch.manual_seed(0)
im = torch.rand(1, 3, H, W, device="cuda")
dg, init, y, mask = P.make_inpainting(im)
s = 2 / 255
res = {}
for rep in range(2):
    for depth in (2, 4, 8, 12, 20, 36):
        den = P.DnCNN(depth=depth, pretrained=P.lipschitz_dncnn_state_dict(0, depth=depth))
        r = P.psgla_run(init, dg, den, n_iter=1400, n_chains=B, alpha=1.0, lambd=5.0, sig_float=s, delta=s * s, n_inter=10, n_inter_mmse=10, seed=0)
        for i in range(600):
            r.step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(600, 1200):
            r.step(i)
        e1.record()
        torch.cuda.synchronize()
        res[depth] = e0.elapsed_time(e1) / 600 * 1e3
        del r, den
    ds = sorted(res)
    print("B=%d %dx%d " % (B, H, W) + "  ".join("depth %d: %.1f us" % (d, res[d]) for d in ds), flush=True)
    print("   per fused pair (depth 20 -> 36): %.2f us; (4 -> 20): %.2f us; depth 2 (first + last only): %.1f us" % (
        (res[36] - res[20]) / 8, (res[20] - res[4]) / 8, res[2]), flush=True)
