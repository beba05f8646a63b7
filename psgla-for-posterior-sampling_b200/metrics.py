"""Image-quality metrics of the reference's post-processing (sampling_images.py:371-442) on the GPU: PSNR / SSIM of every
stored sample, of the cumulative posterior mean, and the posterior standard-deviation map -- without the reference's
per-sample device-to-host copies (1 000 samples x 786 KB at 256 x 256) and without skimage.

``psnr`` / ``ssim`` follow skimage 0.24's ``peak_signal_noise_ratio`` / ``structural_similarity(channel_axis=...)`` defaults as
the reference calls them (7 x 7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03, data_range = 1).
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["psnr_ssim", "posterior_summary"]


def _stack(x):
    if isinstance(x, (list, tuple)):
        x = torch.stack([t.reshape(t.shape[-3:]) for t in x])
    if x.dim() == 3:
        x = x[None]
    if x.dim() != 4:
        raise ValueError("expected [n, C, H, W] (or a list of [C, H, W] tensors)")
    if not x.is_cuda:
        raise RuntimeError("metrics run on the GPU: pass CUDA tensors (there is no CPU path)")
    return x.to(torch.float32).contiguous()


def psnr_ssim(images, ref, data_range=1.0):
    """PSNR and SSIM of ``images`` [n, C, H, W] (or a list of [C, H, W]) against ``ref`` [C, H, W]; two CUDA tensors [n]."""
    x = _stack(images)
    r = ref.reshape(ref.shape[-3:]).to(x.device, torch.float32).contiguous()
    if tuple(r.shape) != tuple(x.shape[1:]):
        raise ValueError("reference shape %s does not match the images %s" % (tuple(r.shape), tuple(x.shape[1:])))
    n = int(x.shape[0])
    lib = _lib.lib()
    psnr = torch.empty(n, dtype=torch.float32, device=x.device)
    ssim = torch.empty(n, dtype=torch.float32, device=x.device)
    ws = torch.empty(lib.psgla_img_metrics_workspace_bytes(n), dtype=torch.uint8, device=x.device)
    shape = _lib.ImgShape(n, int(x.shape[1]), int(x.shape[2]), int(x.shape[3]))
    with torch.cuda.device(x.device):
        _lib.check(lib.psgla_img_psnr_ssim(shape, _lib.ptr(x), _lib.ptr(r), float(data_range), _lib.ptr(ws), ws.numel(),
                                           _lib.ptr(psnr), _lib.ptr(ssim), _lib.stream_ptr(x.device)), "psgla_img_psnr_ssim")
    return psnr, ssim


def posterior_summary(im, Xlist, Xlist_mmse, Xlist_mmse2, data_range=1.0):
    """The numbers sampling_images.py:373-439 derives from a sampler's three lists, computed on the GPU.

    Returns a dict: ``psnr_samples`` / ``ssim_samples`` (per stored sample, :373-384), ``psnr_running`` / ``ssim_running`` of
    the cumulative mean of the window means for i = 1..n-1 (:411-424), ``xmmse`` (mean of the window means, :427),
    ``psnr_mmse`` / ``ssim_mmse`` (:428-433), ``std`` = sqrt(max(E[X^2] - E[X]^2, 0)) (:436-439)."""
    out = {}
    if len(Xlist):
        out["psnr_samples"], out["ssim_samples"] = psnr_ssim(Xlist, im, data_range)
    if len(Xlist_mmse):
        M = _stack(Xlist_mmse)
        counts = torch.arange(1, M.shape[0] + 1, device=M.device, dtype=torch.float32)[:, None, None, None]
        running = torch.cumsum(M, 0) / counts
        if M.shape[0] > 1:
            out["psnr_running"], out["ssim_running"] = psnr_ssim(running[1:], im, data_range)
        xmmse = M.mean(0)
        out["xmmse"] = xmmse
        p, s = psnr_ssim(xmmse, im, data_range)
        out["psnr_mmse"], out["ssim_mmse"] = p[0], s[0]
        if len(Xlist_mmse2):
            var = _stack(Xlist_mmse2).mean(0) - xmmse ** 2
            out["std"] = torch.sqrt(torch.clamp(var, min=0.0))
    return out
