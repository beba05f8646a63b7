#!/usr/bin/env python
"""Few-chain DnCNN PSGLA iteration latency: per-layer launches (PSGLA_CONV_FUSE2=0) vs two hidden layers per launch
(conv_fused2.cu), and bit-equality of the iterates.   python scripts/fuse2_probe.py [H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import psgla_b200 as P  # noqa: E402

H, W = (int(v) for v in sys.argv[1:3]) if len(sys.argv) >= 3 else (256, 256)
torch.manual_seed(0)
im = torch.rand(1, 3, H, W, device="cuda")
den = P.DnCNN(pretrained=P.lipschitz_dncnn_state_dict(0))
dg, init, y, mask = P.make_inpainting(im)
s = 2 / 255
kw = dict(alpha=1.0, lambd=5.0, sig_float=s, delta=s * s, n_inter=10, n_inter_mmse=10, seed=0)


def run(B, n=200):
    r = P.psgla_run(init, dg, den, n_iter=n + 20, n_chains=B, **kw)
    for i in range(20):
        r.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20, 20 + n):
        r.step(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n, r.X.clone()


for B in (1, 2):
    rows = []
    ref = None
    for label, env in (("per-layer launches", {"PSGLA_CONV_FUSE2": "0"}), ("layer pairs fused", {"PSGLA_CONV_FUSE2": "1"})):
        os.environ.update(env)
        us, X = run(B)
        if ref is None:
            ref = X
        rows.append("%s %.1f us%s" % (label, us, " (bit-identical)" if torch.equal(X, ref) else " (max diff %.2g)" % (X - ref).abs().max().item()))
    print("B=%-3d %dx%d  " % (B, H, W) + " | ".join(rows), flush=True)
