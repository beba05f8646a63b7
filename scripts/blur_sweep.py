#!/usr/bin/env python
"""Deblurring "pre" kernel timings (development aid): the four-pass kernel against the row-streaming A^T A kernel for several
rows-per-block settings (PSGLA_ATA_RH).   python scripts/blur_sweep.py [B H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import psgla_b200 as P  # noqa: E402

B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 256, 256)
torch.manual_seed(0)
im = torch.rand(1, 3, H, W, device="cuda")
den = P.DnCNN(pretrained=P.lipschitz_dncnn_state_dict(0))
prm = P.sampler_params("pnp_ula", den="DnCNN")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def time_pre(label):
    dg, init, y = P.make_deblurring(im, l=4, blur_type="uniform")
    run = P.pnpula_run(init, dg, P.PriorGrad(den, prm["alpha"], prm["s1"], prm["s2"]), prm["delta"], prm["lambd"], n_iter=4, n_inter=1,
                       n_inter_mmse=1, seed=0, n_chains=B)
    run.pre(0, run.pre_params)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run.pre(1, run.pre_params)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    px = B * H * W
    print("%-28s %7.1f us  (min %7.1f)  %.0f GB/s algorithmic (68 B/px)" % (label, sorted(ts)[2], min(ts), 68 * px / (sorted(ts)[2] * 1e-6) / 1e9), flush=True)


os.environ["PSGLA_BLUR_4PASS"] = "1"
time_pre("four-pass")
os.environ["PSGLA_BLUR_4PASS"] = "0"
for rh in (0, 8, 16, 32, 64, 128):
    if rh:
        os.environ["PSGLA_ATA_RH"] = str(rh)
    time_pre("A^T A rows/block %s" % (rh or "auto"))
