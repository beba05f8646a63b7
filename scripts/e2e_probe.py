import sys, time, os, contextlib
sys.path.insert(0, os.getcwd())
import torch, psgla_b200 as P
import bench
dev = torch.device("cuda", 0)
im = bench.synthetic_image(torch, 256, 256, 0, dev)
den = P.DnCNN(pretrained=P.lipschitz_dncnn_state_dict(0), device=dev)
dg, init, y, mask = P.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
s = 2 / 255
kw = dict(alpha=1.0, lambd=5.0, sig_float=s, delta=s * s, n_inter=10, n_inter_mmse=10, seed=0)
def call(n):
    with contextlib.redirect_stdout(sys.stderr):
        t0 = time.perf_counter()
        run = P.psgla_run(init, dg, den, n_iter=n, n_chains=32, **kw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        for i in range(n): run.step(i)
        t2 = time.perf_counter()
        torch.cuda.synchronize(); t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t1) * 1e3
for n in (100, 100, 300, 300):
    a, b, c = call(n)
    print("n=%d setup %.1f ms, host loop %.1f ms (%.1f us/it), loop+sync %.1f ms (%.3f ms/it)" % (n, a, b, b * 1e3 / n, c, c / n))
