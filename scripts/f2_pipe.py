#!/usr/bin/env python
"""Wall-clock (globaltimer) timeline of 36 consecutive fused layer-pair launches in the single-chain PSGLA loop, nothing
synchronised in between (conv_fused2.cu, PSGLA_F2_TRACE=2): where the period of a launch goes.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgla_b200 as P
im = torch.rand(1, 3, 256, 256, device="cuda")
den = P.DnCNN(pretrained=P.lipschitz_dncnn_state_dict(0))
dg, init, y, mask = P.make_inpainting(im)
s = 2 / 255
r = P.psgla_run(init, dg, den, n_iter=2000, n_chains=1, alpha=1.0, lambd=5.0, sig_float=s, delta=s * s, n_inter=10, n_inter_mmse=10, seed=0)
for i in range(600):
    r.step(i)
os.environ["PSGLA_F2_TRACE"] = "2"
for i in range(600, 640):
    r.step(i)
torch.cuda.synchronize()
