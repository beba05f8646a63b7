import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, psgla_b200 as P
im = torch.rand(1, 3, 256, 256, device="cuda")
den = P.DnCNN(pretrained=P.lipschitz_dncnn_state_dict(0))
dg, init, y, mask = P.make_inpainting(im)
s = 2 / 255
r = P.psgla_run(init, dg, den, n_iter=2000, n_chains=1, alpha=1.0, lambd=5.0, sig_float=s, delta=s*s, n_inter=10, n_inter_mmse=10, seed=0)
for i in range(20): r.step(i)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(300): r.step(20 + i)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("enqueue %.1f us/step, gpu %.1f us/step, wall incl sync %.1f us/step" % ((t1 - t0) / 300 * 1e6, e0.elapsed_time(e1) / 300 * 1e3, (t2 - t0) / 300 * 1e6), flush=True)
