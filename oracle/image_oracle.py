"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch fp32 restatement of the reference's image path.

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  The product package never does.

Parity status
  * ``psgla`` / ``pnpula`` loops, inpainting / deblurring operators, hyper-parameter table:
    PINNED -- ``tests/test_oracle_image.py`` runs the unmodified reference
    (``oracle/ref_loader.py``) on the same inputs in the build container and requires
    bit-identical outputs; committed fixtures in ``tests/golden/image_golden.npz`` pin
    them where the reference is absent.
  * ``DnCNN`` arithmetic lives in third-party ``deepinv==0.2.1`` (environment.yml:311),
    which is neither vendored nor installed: the architecture is restated from its
    published definition (20 conv3x3, 64 features, bias, ReLU, residual, no BN) and is
    PARITY UNPINNED against deepinv itself.
  * PSNR / SSIM (skimage 0.24.0 absent): restated, PARITY UNPINNED.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------- denoiser


class DnCNN(nn.Module):
    """deepinv.models.DnCNN(in_channels=3, out_channels=3, depth=20, bias=True, nf=64) as constructed at
    sampling_images.py:130: in_conv -> ReLU -> 18 x (conv -> ReLU) -> out_conv, plus the input (residual).
    ``sigma`` is accepted and ignored, as in deepinv 0.2.1.  State-dict keys match deepinv's
    (``in_conv``, ``conv_list.{0..17}``, ``out_conv``)."""

    def __init__(self, in_channels=3, out_channels=3, depth=20, bias=True, nf=64):
        super().__init__()
        self.depth = depth
        self.in_conv = nn.Conv2d(in_channels, nf, 3, 1, 1, bias=bias)
        self.conv_list = nn.ModuleList([nn.Conv2d(nf, nf, 3, 1, 1, bias=bias) for _ in range(depth - 2)])
        self.out_conv = nn.Conv2d(nf, out_channels, 3, 1, 1, bias=bias)

    def forward(self, x, sigma=None):
        h = F.relu(self.in_conv(x))
        for conv in self.conv_list:
            h = F.relu(conv(h))
        return self.out_conv(h) + x


def make_dncnn_weights(seed=0, lipschitz=0.9, n_power_iter=30, spatial=32):
    """Seeded random-init DnCNN state dict with a spectrally scaled residual branch.

    The pretrained ``dncnn_sigma2_lipschitz_color.pth`` (README.md:28-29) cannot be fetched
    offline, so BASELINE.json prescribes random-init "Lipschitz-controlled" weights:
    default ``nn.Conv2d`` init under ``torch.manual_seed(seed)``, then every conv is divided by
    its operator norm (power iteration on a ``spatial``-sized periodic grid) and the product is
    scaled so the residual branch has Lipschitz bound ``lipschitz`` < 1.  Biases are scaled with
    their layer so that activations stay O(1).  Deterministic; both the oracle and the CUDA path
    consume the returned fp32 tensors (the CUDA path packs bf16 *from* them).
    """
    g = torch.Generator().manual_seed(seed)
    net = DnCNN()
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 4:
                fan_in = p.shape[1] * 9
                bound = 1.0 / math.sqrt(fan_in)
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.01)
        convs = [net.in_conv, *net.conv_list, net.out_conv]
        per_layer = lipschitz ** (1.0 / len(convs))
        for conv in convs:
            w = conv.weight
            w_adj = w.transpose(0, 1).flip(2, 3)  # adjoint of the circular convolution
            v = torch.randn(1, w.shape[1], spatial, spatial, generator=g)
            v = v / v.norm()
            s = torch.tensor(1.0)
            for _ in range(n_power_iter):
                u = F.conv2d(F.pad(v, [1, 1, 1, 1], mode="circular"), w)
                v = F.conv2d(F.pad(u, [1, 1, 1, 1], mode="circular"), w_adj)
                s = v.norm()  # -> largest eigenvalue of K^T K
                v = v / s
            op_norm = math.sqrt(float(s))
            conv.weight.mul_(per_layer / op_norm)
    return {k: v.detach().clone().float() for k, v in net.state_dict().items()}


class _ResBlock(nn.Module):
    """KAIR ``ResBlock(mode='CRC', bias=False)``: x + conv(relu(conv(x))); parameters live under ``res.0`` / ``res.2``."""

    def __init__(self, c):
        super().__init__()
        self.res = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=False), nn.ReLU(inplace=False), nn.Conv2d(c, c, 3, 1, 1, bias=False))

    def forward(self, x):
        return x + self.res(x)


class DRUNet(nn.Module):
    """``deepinv.models.DRUNet(in_channels=3, out_channels=3)`` as constructed at sampling_images.py:136 -- the DPIR
    network of Zhang et al. (KAIR ``UNetRes``: nc=[64,128,256,512], nb=4, act 'R', strideconv down, convtranspose up,
    no biases), restated from its published definition because deepinv 0.2.1 is absent (PARITY UNPINNED against
    deepinv).  ``forward(x, sigma)`` concatenates a constant noise-level channel and returns the *denoised image*
    (no global residual).  State-dict keys follow the ``drunet_color.pth`` checkpoint: ``m_head``, ``m_down{1,2,3}.{0..3}
    .res.{0,2}``, ``m_down{k}.4`` (2x2 stride-2 conv), ``m_body.{0..3}``, ``m_up{k}.0`` (2x2 stride-2 transposed conv),
    ``m_up{k}.{1..4}``, ``m_tail``.  Inputs whose H or W is not a multiple of 8 are replication-padded at the bottom / right,
    denoised and cropped -- KAIR's ``utils_model.test_pad`` (``ReplicationPad2d((0, pw, 0, ph))``), the rule deepinv applies to
    small inputs [recalled; its rule for large inputs, a 4-way overlapping split, is not restated]."""

    def __init__(self, in_channels=3, out_channels=3, nc=(64, 128, 256, 512), nb=4):
        super().__init__()
        self.m_head = nn.Conv2d(in_channels + 1, nc[0], 3, 1, 1, bias=False)
        self.m_down1 = nn.Sequential(*[_ResBlock(nc[0]) for _ in range(nb)], nn.Conv2d(nc[0], nc[1], 2, 2, 0, bias=False))
        self.m_down2 = nn.Sequential(*[_ResBlock(nc[1]) for _ in range(nb)], nn.Conv2d(nc[1], nc[2], 2, 2, 0, bias=False))
        self.m_down3 = nn.Sequential(*[_ResBlock(nc[2]) for _ in range(nb)], nn.Conv2d(nc[2], nc[3], 2, 2, 0, bias=False))
        self.m_body = nn.Sequential(*[_ResBlock(nc[3]) for _ in range(nb)])
        self.m_up3 = nn.Sequential(nn.ConvTranspose2d(nc[3], nc[2], 2, 2, 0, bias=False), *[_ResBlock(nc[2]) for _ in range(nb)])
        self.m_up2 = nn.Sequential(nn.ConvTranspose2d(nc[2], nc[1], 2, 2, 0, bias=False), *[_ResBlock(nc[1]) for _ in range(nb)])
        self.m_up1 = nn.Sequential(nn.ConvTranspose2d(nc[1], nc[0], 2, 2, 0, bias=False), *[_ResBlock(nc[0]) for _ in range(nb)])
        self.m_tail = nn.Conv2d(nc[0], out_channels, 3, 1, 1, bias=False)

    def forward(self, x, sigma):
        if isinstance(sigma, torch.Tensor):
            sigma = float(sigma.reshape(-1)[0])
        H0, W0 = x.shape[2], x.shape[3]
        ph, pw = (-H0) % 8, (-W0) % 8
        if ph or pw:
            return self.forward(F.pad(x, (0, pw, 0, ph), mode="replicate"), sigma)[:, :, :H0, :W0]
        noise_map = torch.full((x.shape[0], 1, x.shape[2], x.shape[3]), float(sigma), dtype=x.dtype, device=x.device)
        x0 = torch.cat((x, noise_map), 1)
        x1 = self.m_head(x0)
        x2 = self.m_down1(x1)
        x3 = self.m_down2(x2)
        x4 = self.m_down3(x3)
        h = self.m_body(x4)
        h = self.m_up3(h + x4)
        h = self.m_up2(h + x3)
        h = self.m_up1(h + x2)
        return self.m_tail(h + x1)


def make_drunet_weights(seed=0, gain=0.5):
    """Seeded random-init DRUNet state dict (checkpoints are unreachable offline).  Uniform +-gain*sqrt(3/fan_in) keeps
    activations O(1) through the 4 x 4 residual blocks per scale; the tail is scaled down so that D(x) stays near the
    input range.  Deterministic; the oracle and the CUDA path consume the same fp32 tensors."""
    g = torch.Generator().manual_seed(seed)
    net = DRUNet()
    with torch.no_grad():
        for name, p in net.named_parameters():
            fan_in = p.shape[1] * p.shape[2] * p.shape[3] if not name.startswith("m_up") or ".res." in name else p.shape[0] * 4
            bound = gain * math.sqrt(3.0 / fan_in)
            if ".res.2." in name:
                bound *= 0.5  # second conv of a residual block: keep the branch a perturbation of the identity
            if name.startswith("m_tail"):
                bound *= 0.25
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
    return {k: v.detach().clone().float() for k, v in net.state_dict().items()}


# ----------------------------------------------------------------------------- operators


def make_inpainting(im_t, prop=0.5, sigma=1.0, seed_ip=0, device="cpu"):
    """sampling_images.py:283-302.  Returns dict(mask, y, sigma2, init, data_grad)."""
    sigma1 = sigma / 255.0
    sigma2 = sigma1 ** 2
    sigma2t = torch.tensor(sigma2, dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed_ip)
    mask = torch.rand((im_t.shape[2], im_t.shape[3]), generator=gen, device=device)  # :287
    mask_2d = 1 * (mask > prop)  # :289
    mask = (torch.ones(im_t.shape[1])[None, :, None, None].to(device)) * mask_2d[None, None, :, :]  # :291
    neg_mask = 1 - mask
    y_t = mask * im_t + torch.normal(torch.zeros(*im_t.size()).to(device),
                                     std=sigma1 * torch.ones(*im_t.size()).to(device), generator=gen)  # :294
    data_grad = lambda x: -mask * (x - y_t) / (sigma2t)  # :295
    init = mask * y_t + neg_mask * 0.5 * torch.ones(y_t.shape).to(device)  # :302
    return dict(mask=mask, y=y_t, sigma2=sigma2, sigma2t=sigma2t, init=init, data_grad=data_grad)


def blur_taps(l=4, blur_type="uniform", si=1.0):
    """1 x (2l+1) normalised taps, sampling_images.py:306-312 (float64)."""
    if blur_type == "uniform":
        h = np.ones((1, 2 * l + 1))
    elif blur_type == "gaussian":
        h = np.array([[np.exp(-i ** 2 / (2 * si ** 2)) for i in range(-l, l + 1)]])
    else:
        raise ValueError(blur_type)
    return h / np.sum(h)


def blur_operators(h, l, C=3, device="cpu"):
    """``A`` and ``AT`` exactly as sampling_images.py:313-330 builds them from the 1 x (2l+1) taps ``h``: circular pad l +
    depthwise conv2d with flip(h^T h) resp. h^T h."""
    h = np.asarray(h, dtype=np.float64).reshape(1, -1)
    h_ = np.dot(h.T, h)  # :313
    h_conv = np.copy(np.flip(h_))  # :314-315
    hconv = torch.from_numpy(h_conv).type(torch.FloatTensor).to(device)
    hcorr = torch.from_numpy(h_).type(torch.FloatTensor).to(device)
    ones = torch.ones(C, hconv.shape[0], hconv.shape[1]).to(device)
    hconv = hconv.unsqueeze(0)[None, :, :, :] * ones[:, None, :, :]  # :323-324
    hcorr = hcorr.unsqueeze(0)[None, :, :, :] * ones[:, None, :, :]
    A = lambda x: F.conv2d(F.pad(x, [l, l, l, l], mode="circular"), hconv, groups=x.size(1), padding=0)  # :329
    AT = lambda x: F.conv2d(F.pad(x, [l, l, l, l], mode="circular"), hcorr, groups=x.size(1), padding=0)  # :330
    return A, AT


def deblur_data_grad(x, h, l, y, sigma2):
    """``-AT(A(x) - y) / sigma2`` (sampling_images.py:338) in the reference's conv2d formulation, on x's device."""
    A, AT = blur_operators(h, l, x.shape[1], x.device)
    return -AT(A(x) - y) / torch.tensor(sigma2, dtype=torch.float32, device=x.device)


def make_deblurring(im_t, l=4, blur_type="uniform", si=1.0, sigma=1.0, seed_ip=0, device="cpu"):
    """sampling_images.py:304-341.  Returns dict(A, AT, y, sigma2, init, data_grad, h)."""
    sigma1 = sigma / 255.0
    sigma2 = sigma1 ** 2
    sigma2t = torch.tensor(sigma2, dtype=torch.float32, device=device)
    h = blur_taps(l, blur_type, si)
    A, AT = blur_operators(h, l, im_t.shape[1], device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed_ip)
    y_t = A(im_t) + torch.normal(torch.zeros(*im_t.size()).to(device),
                                 std=sigma1 * torch.ones(*im_t.size()).to(device), generator=gen)  # :335
    data_grad = lambda x: -AT(A(x) - y_t) / (sigma2t)  # :338
    return dict(A=A, AT=AT, y=y_t, sigma2=sigma2, sigma2t=sigma2t, init=y_t, data_grad=data_grad, h=h)


# ----------------------------------------------------------------------------- hyper-parameters


def resolve_params(alg, den="DnCNN", sigma=1.0, alpha=1.0, N=10000, s=None, lambd=None, N_given=False):
    """The numbers sampling_images.py:100-123,147-198 resolves, without the ``sys.argv`` sniffing:
    ``s``/``lambd``/``N_given`` = None/False mean "flag absent from the command line"."""
    sigma1 = sigma / 255.0
    sigma2 = sigma1 ** 2
    out = dict(sigma1=sigma1, sigma2=sigma2, alpha=alpha, n_inter=int(N / 1000))  # :105 uses the *parsed* N
    out["n_inter_mmse"] = out["n_inter"]
    if alg == "pnp_ula":
        s_ = (2.0 / 255.0) if (s is None and den == "DnCNN") else (5.0 if s is None else s)  # :149-152
        s1 = s_ / 255.0  # :153 (second division by 255 -- quirk kept)
        s2 = s1 ** 2
        N_ = 100000 if (not N_given and den == "DnCNN") else N  # :159-162
        lam = 0.5 / (2 / sigma2 + alpha / s2)  # :164
        delta = 1 / 3 / (1 / sigma2 + 1 / lam + alpha / s2)  # :167
        out.update(s=s_, s1=s1, s2=s2, N=N_, lambd=lam, delta=delta)
    elif alg == "psgla":
        if den == "DnCNN":
            s_ = 2.0 / 255.0 if s is None else s / 255.0  # :172-175
            lam = 5.0 if lambd is None else lambd  # :176-179
            N_ = N
        else:
            s_ = (5.0 if s is None else s) / 255.0  # :194
            lam = 1.0 if lambd is None else lambd
            N_ = N
        out.update(s=s_, N=N_, lambd=lam, delta=s_ ** 2)  # :198
    else:
        raise ValueError(alg)
    return out


# ----------------------------------------------------------------------------- samplers


def _stats_step(X, xmmse, xmmse2, iter_mmse, n_inter_mmse, lists, im_shape, dtype, device):
    """restoration_algorithms.py:128-144 / :255-271 (identical in both samplers)."""
    Xlist_mmse, Xlist_mmse2 = lists
    if iter_mmse <= n_inter_mmse - 1:
        xmmse = iter_mmse / (iter_mmse + 1) * xmmse + 1 / (iter_mmse + 1) * X
        xmmse2 = iter_mmse / (iter_mmse + 1) * xmmse2 + 1 / (iter_mmse + 1) * X ** 2
        iter_mmse += 1
    else:
        xmmse = iter_mmse / (iter_mmse + 1) * xmmse + 1 / (iter_mmse + 1) * X
        xmmse2 = iter_mmse / (iter_mmse + 1) * xmmse2 + 1 / (iter_mmse + 1) * X ** 2
        Xlist_mmse.append(torch.squeeze(xmmse))
        Xlist_mmse2.append(torch.squeeze(xmmse2))
        xmmse = torch.zeros(im_shape, dtype=dtype, device=device)
        xmmse2 = torch.zeros(im_shape, dtype=dtype, device=device)
        iter_mmse = 0
    return xmmse, xmmse2, iter_mmse


def psgla(init, data_grad, denoiser, alpha, lambd, sig_float=0.0055, delta=4e-5, n_iter=5000, n_inter=1000,
          n_inter_mmse=1000, seed=None, device="cpu", noise=None):
    """restoration_algorithms.py:163-285.  ``noise`` (n_iter, *init.shape) replaces ``torch.randn`` (:232)."""
    dtype = torch.float32
    im_shape = init.shape
    X = init.clone().detach()
    xmmse = torch.zeros(im_shape, dtype=dtype, device=device)
    xmmse2 = torch.zeros(im_shape, dtype=dtype, device=device)
    delta = torch.tensor(delta).to(device).to(torch.float32)  # :203
    sig = torch.tensor(sig_float).to(device).to(torch.float32)  # :205
    if seed is not None:
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
    if n_inter_mmse is None:
        n_inter_mmse = np.copy(n_inter)
    Xlist, Xlist_mmse, Xlist_mmse2 = [], [], []
    iter_mmse = 0
    noise_ratio = torch.tensor(np.sqrt(2)).to(device).to(torch.float32)  # :228
    with torch.no_grad():
        for i in range(n_iter):
            Z = noise[i] if noise is not None else torch.randn(im_shape, generator=gen, dtype=dtype, device=device)
            grad_log_data = data_grad(X)
            Y = X + (delta / lambd) * grad_log_data + noise_ratio * sig * Z  # :236
            X = (1 - alpha) * Y + alpha * denoiser.forward(Y, sig)  # :238
            if i % n_inter == 0:
                Xlist.append(torch.squeeze(X))
            xmmse, xmmse2, iter_mmse = _stats_step(X, xmmse, xmmse2, iter_mmse, n_inter_mmse,
                                                   (Xlist_mmse, Xlist_mmse2), im_shape, dtype, device)
    return Xlist, Xlist_mmse, Xlist_mmse2


def pnpula(init, data_grad, prior_grad, delta, lambd, n_iter=5000, n_inter=1000, n_inter_mmse=1000, seed=None,
           device="cpu", c_min=-1, c_max=2, noise=None):
    """restoration_algorithms.py:38-160.  ``delta`` and ``lambd`` are 0-dim tensors (:79)."""
    dtype = torch.float32
    im_shape = init.shape
    X = init.clone().detach()
    One = torch.ones(im_shape, dtype=dtype, device=device)
    xmmse = torch.zeros(im_shape, dtype=dtype, device=device)
    xmmse2 = torch.zeros(im_shape, dtype=dtype, device=device)
    brw = torch.sqrt(2 * delta).to(device)  # :79
    if seed is not None:
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
    if n_inter_mmse is None:
        n_inter_mmse = np.copy(n_inter)
    Xlist, Xlist_mmse, Xlist_mmse2 = [], [], []
    iter_mmse = 0
    with torch.no_grad():
        for i in range(n_iter):
            Z = noise[i] if noise is not None else torch.randn(im_shape, generator=gen, dtype=dtype, device=device)
            grad_log_prior = prior_grad(X)
            grad_log_data = data_grad(X)
            out = torch.where(X > c_min, X, c_min * One)  # :110
            proj = torch.where(out < c_max, out, c_max * One)  # :111
            gradPi = grad_log_prior - (X - proj) / lambd + grad_log_data  # :113
            X = X + delta * gradPi + brw * Z  # :115
            if i % n_inter == 0:
                Xlist.append(torch.squeeze(X))
            xmmse, xmmse2, iter_mmse = _stats_step(X, xmmse, xmmse2, iter_mmse, n_inter_mmse,
                                                   (Xlist_mmse, Xlist_mmse2), im_shape, dtype, device)
    return Xlist, Xlist_mmse, Xlist_mmse2


def pnp(init, data_grad, Pb, denoiser, alpha, lambd, sig_float=0.0055, delta=1e-5, n_iter=500, device="cpu"):
    """restoration_algorithms.py:386-463 (PnP forward-backward): PSGLA without the noise, every iterate stored, the
    denoiser level held at 40/255 for the first n_iter // 10 iterations of an inpainting problem (:444-447)."""
    X = init.clone().detach()
    delta = torch.tensor(delta).to(device).to(torch.float32)
    sig = torch.tensor(sig_float).to(device).to(torch.float32)
    Xlist = []
    with torch.no_grad():
        for i in range(n_iter):
            sig_den = 40.0 / 255.0 if (Pb == "inpainting" and i < n_iter // 10) else sig
            Y = X + (delta / lambd) * data_grad(X)  # :449
            X = (1 - alpha) * Y + alpha * denoiser.forward(Y, sig_den)  # :451
            Xlist.append(torch.squeeze(X))
    return Xlist, [torch.squeeze(X)], []


def red(init, data_grad, Pb, denoiser, lambd, sig_float=0.0055, delta=1e-5, n_iter=500, device="cpu"):
    """restoration_algorithms.py:465-529 (RED): X+ = X + delta grad - delta lambd (X - D(X; sig)); level 50/255 for the
    first 10 iterations of an inpainting problem (:512-515)."""
    X = init.clone().detach()
    delta = torch.tensor(delta).to(device).to(torch.float32)
    sig = torch.tensor(sig_float).to(device).to(torch.float32)
    Xlist = []
    with torch.no_grad():
        for i in range(n_iter):
            sig_den = 50.0 / 255.0 if (i < 10 and Pb == "inpainting") else sig
            X = X + delta * data_grad(X) - delta * lambd * (X - denoiser.forward(X, sig_den))  # :518
            Xlist.append(torch.squeeze(X))
    return Xlist, [torch.squeeze(X)], []


class SigmaBlendDenoiser:
    """Test double that makes the noise-level argument observable: D(x, sigma) = x + (1 + 4 sigma) R(x) with R the
    residual of a (sigma-blind) DnCNN.  Used to pin the sigma-annealing schedules of pnp / red."""

    def __init__(self, dncnn):
        self.net = dncnn

    def forward(self, x, sigma):
        s = float(sigma.reshape(-1)[0]) if isinstance(sigma, torch.Tensor) else float(sigma)
        return x + (1.0 + 4.0 * s) * (self.net(x) - x)


def make_prior_grad(denoiser, alpha, s1, s2, device="cpu"):
    """sampling_images.py:155-157."""
    alphat = torch.tensor(alpha, dtype=torch.float32, device=device)
    s2t = torch.tensor(s2, dtype=torch.float32, device=device)
    Ds = lambda x: denoiser.forward(x, s1)
    return lambda x: alphat * (Ds(x) - x) / s2t


# ----------------------------------------------------------------------------- metrics (restated, unpinned)


def psnr(ref, img, data_range=1.0):
    """skimage.metrics.peak_signal_noise_ratio (sampling_images.py:377,429): 10 log10(R^2 / MSE), float64."""
    ref = np.asarray(ref, dtype=np.float64)
    img = np.asarray(img, dtype=np.float64)
    return 10.0 * np.log10(data_range ** 2 / np.mean((ref - img) ** 2))


def ssim(ref, img, data_range=1.0):
    """skimage.metrics.structural_similarity defaults with channel_axis=2 (sampling_images.py:381,433):
    7x7 uniform window, sample covariance, K1=.01, K2=.03, mean over the cropped interior, averaged over channels."""
    from scipy.ndimage import uniform_filter
    ref = np.asarray(ref, dtype=np.float64)
    img = np.asarray(img, dtype=np.float64)
    win, K1, K2 = 7, 0.01, 0.03
    NP = win * win
    cov_norm = NP / (NP - 1)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    pad = (win - 1) // 2
    vals = []
    for ch in range(ref.shape[2]):
        a, b = ref[..., ch], img[..., ch]
        ux, uy = uniform_filter(a, win), uniform_filter(b, win)
        uxx, uyy, uxy = uniform_filter(a * a, win), uniform_filter(b * b, win), uniform_filter(a * b, win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        vals.append(S[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))
