"""-m gpu (slow, ~3 min): BASELINE.json configs[2] at its stated length on a real reference test image.

`datasets/set3c/butterfly.png` (256 x 256, shipped under tests/golden/set3c/ -- a reference-held test input), random
inpainting 50 %, sigma = 1/255, PSGLA with the script's DnCNN table (s = 2/255, lambda = 5, delta = s^2, alpha = 1),
N = 10 000 iterations, n_inter = n_inter_mmse = 10 (sampling_images.py:105-106,170-198,351).

The product runs as a drop-in: `psgla(..., seed=k)` alone, noise generated inside the fused kernels as torch's CUDA Philox
stream.  The fp32 oracle (the reference's loop, restoration_algorithms.py:163-285, cuDNN fp32 convolutions with TF32 off) draws
`torch.randn(generator=Generator("cuda").manual_seed(k))` per iteration like the reference.  Compared on the RESULT, as the
north star asks: PSNR / SSIM of the MMSE estimate, the posterior std map, the per-sample PSNR curve and the bookkeeping counts
(1 000 samples, 909 window means).  Two denoisers:
  * "lipschitz": the seeded random-init, Lipschitz-0.9 DnCNN BASELINE.json prescribes for the unavailable checkpoint.  It
    restores nothing -- unobserved pixels random-walk and drift with the network's biases, |X| reaches ~1e2 -- so PSNR is
    meaningless as a quality figure but a sharp parity probe: both implementations must drift identically.  Bounds relative to
    the state's scale.
  * "smoothing": psgla_b200.smoothing_dncnn_state_dict, a hand-written DnCNN-architecture network whose residual is
    eps (G x - x), i.e. an actual (weak) denoiser: the chain is stable, the MMSE estimate must beat the initialisation by
    several dB in BOTH implementations, and the two must agree to 0.05 dB / 1e-3 SSIM."""
import os

import pytest
import torch

import psgla_b200 as P
from conftest import observed
from oracle import image_oracle as io_

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("which", ["lipschitz", "smoothing"])
def test_set3c_psgla_n10000_final_psnr_ssim_parity(which):
    N = int(os.environ.get("PSGLA_FULL_N", "10000"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    im = P.load_image(os.path.join(HERE, "golden", "set3c", "butterfly.png"), "cuda")
    assert tuple(im.shape) == (1, 3, 256, 256)
    sd = P.lipschitz_dncnn_state_dict(0) if which == "lipschitz" else P.smoothing_dncnn_state_dict()
    den = P.DnCNN(pretrained=sd)
    net = io_.DnCNN().cuda()
    net.load_state_dict(sd)
    net.eval()
    dg, init, y, mask = P.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    prm = P.sampler_params("psgla", den="DnCNN", N=N)
    kw = P.as_psgla_kwargs(prm, seed=3)
    assert kw["n_iter"] == N and kw["n_inter"] == N // 1000 and abs(kw["sig_float"] - 2 / 255) < 1e-12 and kw["lambd"] == 5.0
    Xg, Mg, M2g = P.psgla(init, dg, den, **kw)  # drop-in call: the seed alone, rng = torch's CUDA stream in-kernel
    okw = dict(kw, alpha=torch.tensor(kw["alpha"], device="cuda"), lambd=torch.tensor(kw["lambd"], device="cuda"))
    with torch.no_grad():
        Xr, Mr, M2r = io_.psgla(init, dg, net, device="cuda", **okw)
    n_inter = kw["n_inter"]
    assert len(Xg) == len(Xr) == (N + n_inter - 1) // n_inter and len(Mg) == len(Mr) == N // (n_inter + 1) == len(M2g) == len(M2r)
    a, b = P.posterior_summary(im[0], Xr, Mr, M2r), P.posterior_summary(im[0], Xg, Mg, M2g)
    p_init = P.psnr_ssim(init[0], im[0])[0].item()
    scale = max(1.0, torch.stack(Xr[-10:]).abs().max().item())  # |X| of the oracle chain at the end of the run
    print("%s: oracle PSNR(mmse) %.3f dB SSIM %.4f | cuda PSNR(mmse) %.3f dB SSIM %.4f | init %.2f dB | state scale %.3g"
          % (which, a["psnr_mmse"].item(), a["ssim_mmse"].item(), b["psnr_mmse"].item(), b["ssim_mmse"].item(), p_init, scale))
    observed(which + " |dPSNR(mmse)| dB", abs(a["psnr_mmse"].item() - b["psnr_mmse"].item()), 0.05)
    observed(which + " |dSSIM(mmse)|", abs(a["ssim_mmse"].item() - b["ssim_mmse"].item()), 1e-3)
    observed(which + " max |d std map| / scale", (a["std"] - b["std"]).abs().max().item() / scale, 5e-3)
    observed(which + " max |d xmmse| / scale", (a["xmmse"] - b["xmmse"]).abs().max().item() / scale, 5e-3)
    observed(which + " max |dPSNR(sample)| dB", (a["psnr_samples"] - b["psnr_samples"]).abs().max().item(), 0.1)
    observed(which + " max |dPSNR(running mmse)| dB", (a["psnr_running"] - b["psnr_running"]).abs().max().item(), 0.05)
    observed(which + " last iterate max abs error / scale", (Xr[-1] - Xg[-1]).abs().max().item() / scale, 1e-2)
    if which == "smoothing":
        assert scale < 2.0  # a stable chain in (about) [0, 1]
        assert a["psnr_mmse"].item() > p_init + 5.0 and b["psnr_mmse"].item() > p_init + 5.0  # and it restores: both sides
