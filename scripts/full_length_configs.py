#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] at their STATED length on the reference's own test images (one-off evidence runs; the
test-suite holds configs[2] at N = 10^4, tests/test_set3c_full_gpu.py).   python scripts/full_length_configs.py [--quick]

configs[3]: set3c uniform-blur deblurring (9 x 9, l = 4), PnP-ULA with DnCNN, the script's default table (N = 100 000, n_inter =
  n_inter_mmse = 10, delta ~ 1e-10; sampling_images.py:105-106,147-168,358), one chain as the reference runs it, against the fp32
  oracle (the reference's loop on the same device) on the same CUDA noise stream: PSNR / SSIM of the MMSE estimate, std map.
configs[4]: a CBSD-sized image (set1c/castle.png, 481 x 321) inpainting PSGLA with the DRUNet-architecture denoiser, 64 chains,
  N = 10 000, statistics-only mode through run_image_set (per-call replication padding to 488 x 328)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import psgla_b200 as P  # noqa: E402
from oracle import image_oracle as io_  # noqa: E402

quick = "--quick" in sys.argv
out = {}
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# ---------------------------------------------------------------- configs[3]
im = P.load_image(os.path.join(ROOT, "tests", "golden", "set3c", "starfish.png"), "cuda")
sd = P.lipschitz_dncnn_state_dict(0)
den = P.DnCNN(pretrained=sd)
net = io_.DnCNN().cuda()
net.load_state_dict(sd)
net.eval()
dg, init, y = P.make_deblurring(im, l=4, blur_type="uniform", sigma=1.0, seed_ip=0)
prm = P.sampler_params("pnp_ula", den="DnCNN")
if quick:
    prm = dict(prm, N=2000)
kw = P.as_pnpula_kwargs(prm, seed=4)
delta = torch.tensor(kw.pop("delta"), device="cuda", dtype=torch.float32)
lambd = torch.tensor(kw.pop("lambd"), device="cuda", dtype=torch.float32)
t0 = time.perf_counter()
Xg, Mg, M2g = P.pnpula(init, dg, P.PriorGrad(den, prm["alpha"], prm["s1"], prm["s2"]), delta, lambd, **kw)  # rng: torch's CUDA stream, in-kernel
torch.cuda.synchronize()
t_cuda = time.perf_counter() - t0
b = P.posterior_summary(im[0], Xg, Mg, M2g)
n_samples, n_windows = len(Xg), len(Mg)
last_g = Xg[-1].clone()
del Xg, Mg, M2g
torch.cuda.empty_cache()
ref_dg = lambda x: io_.deblur_data_grad(x, dg.h1d, 4, y, dg.sigma2)  # noqa: E731  the reference's conv2d formulation
t0 = time.perf_counter()
with torch.no_grad():
    Xr, Mr, M2r = io_.pnpula(init, ref_dg, io_.make_prior_grad(net, prm["alpha"], prm["s1"], prm["s2"], device="cuda"), delta, lambd,
                             device="cuda", **kw)
torch.cuda.synchronize()
t_ref = time.perf_counter() - t0
a = P.posterior_summary(im[0], Xr, Mr, M2r)
out["configs3_pnpula_deblur_dncnn"] = {
    "image": "set3c/starfish.png", "N": prm["N"], "n_inter": prm["n_inter"], "delta": prm["delta"], "lambd": prm["lambd"],
    "samples": n_samples, "windows": n_windows, "samples_oracle": len(Xr), "windows_oracle": len(Mr),
    "seconds_cuda": t_cuda, "seconds_oracle_fp32_same_gpu": t_ref,
    "psnr_mmse_cuda": b["psnr_mmse"].item(), "psnr_mmse_oracle": a["psnr_mmse"].item(),
    "ssim_mmse_cuda": b["ssim_mmse"].item(), "ssim_mmse_oracle": a["ssim_mmse"].item(),
    "psnr_observation": P.psnr_ssim(y[0], im[0])[0].item(),
    "max_abs_diff_std_map": (a["std"] - b["std"]).abs().max().item(), "max_abs_diff_xmmse": (a["xmmse"] - b["xmmse"]).abs().max().item(),
    "max_abs_diff_last_iterate": (Xr[-1] - last_g).abs().max().item(),
    "max_abs_dpsnr_samples": (a["psnr_samples"] - b["psnr_samples"]).abs().max().item()}
print(json.dumps(out["configs3_pnpula_deblur_dncnn"]), flush=True)
del Xr, Mr, M2r, a, b
torch.cuda.empty_cache()

# ---------------------------------------------------------------- configs[4]
castle = P.load_image(os.path.join(ROOT, "tests", "golden", "set1c", "castle.png"))[0]
dru = P.DRUNet(pretrained=P.random_drunet_state_dict(0))
torch.cuda.reset_peak_memory_stats()
n4 = 300 if quick else 10000
prm4 = dict(P.sampler_params("psgla", den="DRUNet", lambd=25.0, N=n4))
prm4["n_inter"] = prm4["n_inter_mmse"] = max(prm4["n_inter"], 1)  # the script's int(N / 1000) is 0 below N = 1000 (quick mode only)
t0 = time.perf_counter()
rows = P.run_image_set([castle], dru, problem="inpainting", alg="psgla", n_chains=64, params=prm4, seed=0)
torch.cuda.synchronize()
t4 = time.perf_counter() - t0
out["configs4_psgla_inpainting_drunet_64_chains"] = {
    "image": "set1c/castle.png (481 x 321, padded per call to 488 x 328)", "N": n4, "n_inter": prm4["n_inter"], "chains": 64,
    "seconds": t4, "image_iterations_per_s": 64 * n4 / t4, "peak_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "row": rows[0]}
print(json.dumps(out["configs4_psgla_inpainting_drunet_64_chains"]), flush=True)
with open(os.path.join(ROOT, "gpurun_out", "full_length_configs%s.json" % ("_quick" if quick else "")), "w") as fh:
    json.dump(out, fh, indent=1)
