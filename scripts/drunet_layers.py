"""DRUNet PSGLA iterations at the bench's shape (DIAG_B chains of DIAG_HW) for a per-launch ncu list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python scripts/drunet_layers.py
Without ncu it prints the iteration time from CUDA events."""
import os, sys, contextlib
sys.path.insert(0, os.getcwd())
import torch
import psgla_b200 as P

B = int(os.environ.get("DIAG_B", "64"))
H, W = [int(v) for v in os.environ.get("DIAG_HW", "320x480").split("x")]
N = int(os.environ.get("DIAG_N", "3"))
dev = torch.device("cuda")
den = P.DRUNet(pretrained=P.random_drunet_state_dict(0), device=dev)
im = torch.rand(1, 3, H, W, device=dev)
dg, init, y, mask = P.make_inpainting(im, 0.5, 1.0, 0)
s = 5 / 255
with contextlib.redirect_stdout(sys.stderr):
    run = P.psgla_run(init, dg, den, 1.0, 25.0, s, s * s, n_iter=1000, n_inter=10, n_inter_mmse=10, seed=0, n_chains=B)
for i in range(2):
    run.step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(2, 2 + N):
    run.step(i)
e1.record()
torch.cuda.synchronize()
print("DRUNet B=%d %dx%d: %.3f ms per PSGLA iteration" % (B, H, W, e0.elapsed_time(e1) / N))
