"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle only finishes small cases):
sharding / segmentation / batch invariance, closed-form recursions, FFT diagonalisation of the circular blur,
adjointness, bookkeeping counts, and statistical agreement with the closed-form posterior."""
import numpy as np
import pytest
import torch

import psgla_b200 as P
from oracle import gmm2d_oracle as o
from oracle import image_oracle as io_

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------------------------------- configs[1]: 2D GMM
def test_gmm2d_full_size_sharding_and_segmentation_invariance():
    """10^6 chains x 10^4 steps: one launch == 8 shards x 3 segments, bit for bit (Philox keyed by global chain id / step)."""
    mu, Sig, pi = P.gaussian_mixt_example("cross")
    D = P.Theorical_MMSE(mu, Sig, pi)
    y = np.array([-6.0, 6.0])
    n, steps = 1000000, 10000
    kw = dict(y=y, delta=0.3, A=np.eye(2), sigma=1.0, denoiser=D, alpha=2 / 3)
    whole = P.GMMChains("psgla", n_chains=n, seed=3, **kw)
    whole.run(steps)
    parts = []
    for r in range(8):
        a, b = P.dist.shard_range(n, r, 8)
        ch = P.GMMChains("psgla", n_chains=b - a, seed=3, chain_id0=a, **kw)
        for seg in (1, 4999, 5000):
            ch.run(seg)
        parts.append(ch.state)
    assert torch.equal(whole.state, torch.cat(parts, 0))
    X = whole.state.double().cpu().numpy()
    assert np.isfinite(X).all()
    # stationary law against an oracle population (NumPy noise): first two moments
    Xo = o.run_chains("psgla", 600, np.tile(y, (20000, 1)), y, 0.3, np.eye(2), 1, mu, Sig, pi, 2 / 3, rng=np.random.default_rng(5))
    se = np.sqrt(Xo.var(0) / len(Xo) + X.var(0) / len(X))
    assert np.all(np.abs(X.mean(0) - Xo.mean(0)) < 6 * se + 1e-3)
    assert np.abs(np.cov(X.T) - np.cov(Xo.T)).max() < 0.1 * np.abs(np.cov(Xo.T)).max() + 0.02
    # and the W2^2 to exact posterior samples stays in the range the reference's figure reports for PSGLA (< 1)
    rng = np.random.default_rng(0)
    post = P.sample_posterior(np.eye(2), y, 1, 100000, mu, Sig, pi, rng=rng)
    assert np.mean([P.Wasserstein_distance(X, post, rng=rng) for _ in range(4)]) < 1.0


def test_gmm2d_full_size_replay_linearity_of_noise_free_pnpula():
    """With the noise replayed as zeros PnP-ULA is a deterministic map: 10^6 identical chains stay identical and equal the
    float64 oracle's single chain after 10^4 steps (fixed point of the drift), within the fp32 tolerance."""
    mu, Sig, pi = P.gaussian_mixt_example("disymmetric_gaussians")
    D = P.Theorical_MMSE(mu, Sig, pi)
    y = np.array([0.0, -2.0])
    ch = P.GMMChains("pnp_ula", y, 0.1, np.eye(2), 1.0, D, 1.5, 0.5, n_chains=1000000, seed=0)
    zeros = torch.zeros((1, 1000000, 2), device="cuda")
    for _ in range(50):  # 50 replayed steps reach the fixed point to fp32 accuracy
        ch.run(1, noise=zeros)
    X = ch.state
    assert torch.equal(X, X[:1].expand_as(X))
    want = o.pnp_ula(51, y, y, 0.1, np.eye(2), 1, o.theorical_mmse(mu, Sig, pi), 0.5, 1.5, noise=np.zeros((50, 2)))[-1]
    assert np.abs(X[0].double().cpu().numpy() - want).max() < 1e-4


# ----------------------------------------------------------------------------------------------------- configs[2]: 256^2 inpainting
@pytest.fixture(scope="module")
def dncnn():
    return P.DnCNN(pretrained=P.lipschitz_dncnn_state_dict(0))


def _inpainting_256(seed=0):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand((1, 3, 18, 18), generator=g)
    im = torch.nn.functional.interpolate(low, size=(256, 256), mode="bicubic", align_corners=False).clamp(0, 1).cuda()
    return im, P.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)


def test_psgla_256_batch_invariance_determinism_and_bookkeeping(dncnn):
    """set1c-sized problem (256 x 256 x 3, 50 % masked, sigma = 1/255, s = 2/255, lambda = 5, delta = s^2), 60 iterations:
    chains of a batch equal the same chains run alone (global chain ids), reruns are bit-identical, and the list lengths
    follow the reference's rules (restoration_algorithms.py:241-271)."""
    im, (dg, init, y, mask) = _inpainting_256()
    s = 2 / 255
    kw = dict(alpha=1.0, lambd=5.0, sig_float=s, delta=s * s, n_iter=60, n_inter=10, n_inter_mmse=10, seed=7)
    Xb, Mb, M2b = P.psgla(init, dg, dncnn, n_chains=3, **kw)
    assert len(Xb) == 6 and len(Mb) == len(M2b) == 60 // 11
    for c in (0, 2):
        Xs, Ms, _ = P.psgla(init, dg, dncnn, n_chains=1, chain_id0=c, **kw)
        assert all(torch.equal(a[c], b[0]) for a, b in zip(Xb, Xs))
        assert all(torch.equal(a[c], b[0]) for a, b in zip(Mb, Ms))
    Xb2, _, _ = P.psgla(init, dg, dncnn, n_chains=3, **kw)
    assert all(torch.equal(a, b) for a, b in zip(Xb, Xb2))
    X = Xb[-1]
    assert torch.isfinite(X).all()
    # observed pixels are pulled to the observation (gain 0.8 per iteration), masked ones are not
    m = mask.expand(3, -1, -1, -1)[:, 0] if mask.dim() == 4 else mask
    obs_err = ((X - y[0]).abs() * mask[0]).max().item()
    assert obs_err < 0.1
    assert not torch.equal(Xb[-1][0], Xb[-1][1])  # different chains, different noise


def test_psgla_256_zero_denoiser_closed_form_recursion():
    """With zero network weights D = identity and PSGLA is the affine recursion X+ = X - g mask (X - y) + sqrt(2) s Z;
    replaying the library's own Philox draws (psgla_img_noise) in torch reproduces 40 iterations at 256 x 256."""
    im, (dg, init, y, mask) = _inpainting_256(1)
    sd = {k: torch.zeros_like(v) for k, v in P.random_dncnn_state_dict(0).items()}
    den = P.DnCNN(pretrained=sd)
    s = 2 / 255
    n_iter, B = 40, 2
    Xg, Mg, _ = P.psgla(init, dg, den, alpha=1.0, lambd=5.0, sig_float=s, delta=s * s, n_iter=n_iter, n_inter=n_iter - 1,
                        n_inter_mmse=n_iter, seed=11, n_chains=B)
    lib = P._lib.lib()
    shape = P._lib.ImgShape(B, 3, 256, 256)
    X = init.expand(B, -1, -1, -1).clone()
    g = np.float32(np.float32(s * s) / np.float32(5.0)) / np.float32(dg.sigma2)
    c = np.float32(np.sqrt(np.float32(2))) * np.float32(s)
    z = torch.empty_like(X)
    for i in range(n_iter):
        P._lib.check(lib.psgla_img_noise(shape, 11, 0, i, z.data_ptr(), None), "psgla_img_noise")
        X = X - float(g) * (mask * (X - y)) + float(c) * z
    torch.cuda.synchronize()
    assert (Xg[-1] - X).abs().max().item() < 5e-5


# ----------------------------------------------------------------------------------------------------- configs[3]: 9x9 uniform blur
@pytest.mark.parametrize("blur_type,l", [("uniform", 4), ("gaussian", 4)])
def test_blur_256_fft_diagonalisation_and_adjointness(blur_type, l):
    """A is a circular convolution (sampling_images.py:313-330): the DFT diagonalises it, and it is self-adjoint."""
    torch.manual_seed(0)
    x = torch.rand(2, 3, 256, 256, device="cuda")
    z = torch.rand(2, 3, 256, 256, device="cuda")
    h = P.blur_taps(l, blur_type, 1.0).reshape(-1)
    op = P.DeblurDataGrad(h, l, torch.zeros_like(x), (1 / 255) ** 2)
    Ax = op.A(x)
    k = torch.zeros(256, 256, dtype=torch.float64, device="cuda")
    h2 = torch.from_numpy(np.outer(h, h)).cuda()
    for dy in range(-l, l + 1):
        for dx in range(-l, l + 1):
            k[dy % 256, dx % 256] = h2[dy + l, dx + l]
    ref = torch.fft.ifft2(torch.fft.fft2(x.double()) * torch.fft.fft2(k)).real
    assert (Ax.double() - ref).abs().max().item() < 2e-6
    lhs, rhs = (Ax.double() * z.double()).sum().item(), (x.double() * op.A(z).double()).sum().item()
    assert abs(lhs - rhs) < 1e-6 * abs(lhs)


def test_pnpula_256_deblur_runs_and_matches_single_chain(dncnn):
    """set3c-sized deblurring (uniform 9 x 9) with PnP-ULA and the script's parameters for --s 5: batched == single."""
    torch.manual_seed(3)
    im = torch.rand(1, 3, 256, 256, device="cuda")
    dg, init, y = P.make_deblurring(im, l=4, blur_type="uniform", sigma=1.0, seed_ip=0)
    prm = io_.resolve_params("pnp_ula", s=5.0)
    pg = P.PriorGrad(dncnn, 1.0, prm["s1"], prm["s2"])
    kw = dict(delta=torch.tensor(prm["delta"], device="cuda"), lambd=torch.tensor(prm["lambd"], device="cuda"), n_iter=30,
              n_inter=10, n_inter_mmse=10, seed=5)
    Xb, Mb, _ = P.pnpula(init, dg, pg, n_chains=2, **kw)
    Xs, Ms, _ = P.pnpula(init, dg, pg, n_chains=1, chain_id0=1, **kw)
    assert len(Xb) == 3 and len(Mb) == 30 // 11
    assert all(torch.equal(a[1], b[0]) for a, b in zip(Xb + Mb, Xs + Ms))
    assert torch.isfinite(Xb[-1]).all()


# ----------------------------------------------------------------------------------------------------- configs[4]: DRUNet, CBSD-sized
def test_drunet_320x480_forward_against_fp32_torch():
    sd = io_.make_drunet_weights(seed=0)
    den = P.DRUNet(pretrained=sd)
    net = io_.DRUNet().cuda()
    net.load_state_dict(sd)
    x = torch.rand(2, 3, 320, 480, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    with torch.no_grad():
        ref = net(x, 5 / 255)
    got = den.forward(x, 5 / 255)
    torch.cuda.synchronize()
    assert ((got - ref).norm() / ref.norm()).item() < 3e-2
    # batch invariance: image 1 alone gives the same bits
    assert torch.equal(den.forward(x[1:2], 5 / 255)[0], got[1])
