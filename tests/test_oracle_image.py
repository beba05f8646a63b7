"""Pins oracle/image_oracle.py (samplers, operators, parameter table) against the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import image_oracle as io_
from oracle import ref_loader

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_golden.npz"))


def _den():
    den = io_.DnCNN(depth=4, nf=8)
    den.load_state_dict({k[4:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("den.")})
    return den.eval()


def test_inpainting_operator_golden():
    im = torch.from_numpy(G["im"])
    inp = io_.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    assert np.array_equal(inp["mask"].numpy(), G["inp.mask"])
    assert np.array_equal(inp["y"].numpy(), G["inp.y"])
    assert np.array_equal(inp["init"].numpy(), G["inp.init"])


def test_deblurring_operator_golden():
    im = torch.from_numpy(G["im"])
    deb = io_.make_deblurring(im, l=2, blur_type="gaussian", si=1.0, sigma=1.0, seed_ip=0)
    assert np.array_equal(deb["h"], G["deb.h"])
    assert np.allclose(deb["A"](im).numpy(), G["deb.Ax"], atol=1e-7)
    assert np.allclose(deb["y"].numpy(), G["deb.y"], atol=1e-7)
    assert np.allclose(deb["data_grad"](im).numpy(), G["deb.grad_at_im"], rtol=1e-5, atol=1e-2)


def test_psgla_golden_replay_and_seeded():
    im = torch.from_numpy(G["im"])
    inp = io_.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    alpha, lambd, s, delta, n_iter, n_inter, n_mm = G["psgla.params"]
    kw = dict(init=inp["init"], data_grad=inp["data_grad"], denoiser=_den(), alpha=torch.tensor(float(alpha), dtype=torch.float32),
              lambd=torch.tensor(float(lambd)), sig_float=float(s), delta=float(delta), n_iter=int(n_iter),
              n_inter=int(n_inter), n_inter_mmse=int(n_mm))
    for extra in (dict(seed=0), dict(noise=torch.from_numpy(G["psgla.noise"]))):
        Xl, Xm, Xm2 = io_.psgla(**kw, **extra)
        assert len(Xl) == 8 and len(Xm) == 4  # thinning / window bookkeeping (restoration_algorithms.py:241,256-271)
        assert np.allclose(torch.stack(Xl).numpy(), G["psgla.X"], atol=1e-6)
        assert np.allclose(torch.stack(Xm).numpy(), G["psgla.M"], atol=1e-6)
        assert np.allclose(torch.stack(Xm2).numpy(), G["psgla.M2"], atol=1e-6)


def test_pnpula_golden_replay():
    im = torch.from_numpy(G["im"])
    deb = io_.make_deblurring(im, l=2, blur_type="gaussian", si=1.0, sigma=1.0, seed_ip=0)
    delta, lam, alpha, s1, s2, n_iter, n_inter, n_mm, _ = G["ula.params"]
    pg = io_.make_prior_grad(_den(), float(alpha), float(s1), float(s2))
    Xl, Xm, Xm2 = io_.pnpula(init=deb["init"], data_grad=deb["data_grad"], prior_grad=pg,
                             delta=torch.tensor(float(delta), dtype=torch.float32), lambd=torch.tensor(float(lam), dtype=torch.float32),
                             n_iter=int(n_iter), n_inter=int(n_inter), n_inter_mmse=int(n_mm), noise=torch.from_numpy(G["ula.noise"]))
    assert len(Xl) == 8 and len(Xm) == 4
    assert np.allclose(torch.stack(Xl).numpy(), G["ula.X"], atol=2e-6)
    assert np.allclose(torch.stack(Xm).numpy(), G["ula.M"], atol=2e-6)
    assert np.allclose(torch.stack(Xm2).numpy(), G["ula.M2"], atol=2e-6)


def test_resolve_params_table():
    p = io_.resolve_params("psgla")  # sampling_images.py:170-198 defaults
    assert p["s"] == 2.0 / 255 and p["lambd"] == 5.0 and p["delta"] == (2.0 / 255) ** 2 and p["n_inter"] == 10
    assert abs((p["delta"] / p["lambd"]) / p["sigma2"] - 0.8) < 1e-12
    q = io_.resolve_params("pnp_ula")  # double /255 quirk, N=100000, n_inter from the parsed N
    assert q["N"] == 100000 and q["n_inter"] == 10
    assert abs(q["s1"] - 2.0 / 255 / 255) < 1e-18
    assert abs(q["lambd"] - 4.73e-10) / 4.73e-10 < 2e-3 and abs(q["delta"] - 1.05e-10) / 1.05e-10 < 5e-3
    q5 = io_.resolve_params("pnp_ula", s=5.0)
    assert abs(q5["lambd"] - 3.77e-6) / 3.77e-6 < 2e-3 and abs(q5["delta"] - 1.00e-6) / 1.00e-6 < 5e-3


def test_window_count_rule():
    # a window closes after n_inter_mmse+1 iterations => floor(N/(n_inter_mmse+1)) entries
    x0 = torch.zeros(1, 3, 4, 4)
    Xl, Xm, _ = io_.psgla(x0, lambda x: -x, _IdDen(), torch.tensor(1.0), torch.tensor(1.0), 0.01, 1e-4, n_iter=100,
                          n_inter=10, n_inter_mmse=10, seed=0)
    assert len(Xl) == 10 and len(Xm) == 100 // 11


class _IdDen:
    def forward(self, x, s):
        return x


def test_psnr_ssim_sanity():
    rng = np.random.default_rng(0)
    a = rng.uniform(size=(32, 32, 3))
    assert io_.ssim(a, a) == pytest.approx(1.0)
    b = np.clip(a + 0.1, 0, 2)
    assert io_.psnr(a, b) == pytest.approx(20.0, abs=1e-6)


def test_dncnn_weights_deterministic_and_contractive():
    sd1 = io_.make_dncnn_weights(seed=0, n_power_iter=8, spatial=16)
    sd2 = io_.make_dncnn_weights(seed=0, n_power_iter=8, spatial=16)
    assert all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    assert sd1["in_conv.weight"].shape == (64, 3, 3, 3) and sd1["conv_list.17.weight"].shape == (64, 64, 3, 3)
    assert sum(v.numel() for v in sd1.values()) == 668227
    net = io_.DnCNN()
    net.load_state_dict(sd1)
    x1, x2 = torch.rand(1, 3, 24, 24), torch.rand(1, 3, 24, 24)
    with torch.no_grad():
        r1, r2 = net(x1) - x1, net(x2) - x2
    assert (r1 - r2).norm() <= 1.0 * (x1 - x2).norm()


def test_pnp_and_red_golden():
    """PnP forward-backward and RED restatements against fixtures made by the unmodified reference
    (restoration_algorithms.py:386-529), with a sigma-sensitive test denoiser so that the annealing schedules count."""
    im = torch.from_numpy(G["im"])
    sden = io_.SigmaBlendDenoiser(_den())
    inp = io_.make_inpainting(im)
    a, lam, s, delta, n = [float(v) for v in G["pnp.params"]]
    Xl, Xf, rest = io_.pnp(inp["init"], inp["data_grad"], "inpainting", sden, torch.tensor(a), torch.tensor(lam), sig_float=s,
                           delta=delta, n_iter=int(n))
    assert rest == [] and len(Xf) == 1 and torch.equal(Xf[0], Xl[-1])
    assert np.array_equal(torch.stack(Xl).numpy(), G["pnp.X"])
    lam, s, delta, n = [float(v) for v in G["red.params"]]
    Xl, Xf, rest = io_.red(inp["init"], inp["data_grad"], "inpainting", sden, torch.tensor(lam), sig_float=s, delta=delta,
                           n_iter=int(n))
    assert np.array_equal(torch.stack(Xl).numpy(), G["red.X"])
    # the schedule matters: a constant level gives a different trajectory
    Xc, _, _ = io_.red(inp["init"], inp["data_grad"], "deblurring", sden, torch.tensor(lam), sig_float=s, delta=delta, n_iter=int(n))
    assert not np.array_equal(torch.stack(Xc).numpy(), G["red.X"])


def test_drunet_restatement_shape_and_size():
    """deepinv's DRUNet is absent (parity unpinned): the restatement has the published parameter count and key set."""
    net = io_.DRUNet()
    assert sum(p.numel() for p in net.parameters()) == 32640960
    keys = list(net.state_dict())
    assert keys[0] == "m_head.weight" and keys[-1] == "m_tail.weight" and len(keys) == 64
    assert net.state_dict()["m_down1.4.weight"].shape == (128, 64, 2, 2)
    assert net.state_dict()["m_up3.0.weight"].shape == (512, 256, 2, 2)
    sd = io_.make_drunet_weights(seed=1)
    net.load_state_dict(sd)
    with torch.no_grad():
        y = net(torch.rand(1, 3, 16, 24), 0.02)
    assert y.shape == (1, 3, 16, 24) and torch.isfinite(y).all()


@pytest.mark.reference
def test_live_reference_pnp_red_bit_identical():
    ra = ref_loader.load_restoration_algorithms()
    im = torch.from_numpy(G["im"])
    sden = io_.SigmaBlendDenoiser(_den())
    deb = io_.make_deblurring(im, l=2, blur_type="gaussian")
    inp = io_.make_inpainting(im)
    for prob, Pb in ((inp, "inpainting"), (deb, "deblurring")):
        kw = dict(init=prob["init"], data_grad=prob["data_grad"], Pb=Pb, denoiser=sden, alpha=torch.tensor(0.7),
                  lambd=torch.tensor(5.0), sig_float=2 / 255, delta=(2 / 255) ** 2, n_iter=25, device="cpu")
        a, b = ra.pnp(**kw), io_.pnp(**kw)
        assert all(torch.equal(x, y) for x, y in zip(a[0] + a[1], b[0] + b[1])) and a[2] == b[2] == []
        kw = dict(init=prob["init"], data_grad=prob["data_grad"], Pb=Pb, denoiser=sden, lambd=torch.tensor(1500.0),
                  sig_float=2 / 255, delta=2e-5, n_iter=13, device="cpu")
        a, b = ra.red(**kw), io_.red(**kw)
        assert all(torch.equal(x, y) for x, y in zip(a[0] + a[1], b[0] + b[1]))


@pytest.mark.reference
def test_live_reference_bit_identical():
    ra = ref_loader.load_restoration_algorithms()
    im = torch.from_numpy(G["im"])
    den = _den()
    inp = io_.make_inpainting(im)
    kw = dict(init=inp["init"], data_grad=inp["data_grad"], denoiser=den, alpha=torch.tensor(1.0), lambd=torch.tensor(5.0),
              sig_float=2 / 255, delta=(2 / 255) ** 2, seed=11, device="cpu", n_iter=30, n_inter=4, n_inter_mmse=5)
    a = ra.psgla(**kw)
    b = io_.psgla(**kw)
    for la, lb in zip(a, b):
        assert len(la) == len(lb)
        assert all(torch.equal(x, y) for x, y in zip(la, lb))
    deb = io_.make_deblurring(im, l=4, blur_type="uniform")
    p = io_.resolve_params("pnp_ula", s=5.0)
    pg = io_.make_prior_grad(den, 1.0, p["s1"], p["s2"])
    kw = dict(init=deb["init"], data_grad=deb["data_grad"], prior_grad=pg, delta=torch.tensor(p["delta"], dtype=torch.float32),
              lambd=torch.tensor(p["lambd"], dtype=torch.float32), seed=2, device="cpu", n_iter=30, n_inter=4, n_inter_mmse=5)
    a = ra.pnpula(**kw)
    b = io_.pnpula(**kw)
    for la, lb in zip(a, b):
        assert len(la) == len(lb)
        assert all(torch.equal(x, y) for x, y in zip(la, lb))


def test_ssim_against_the_definition_window_by_window():
    """skimage is absent, so the SSIM restatement (uniform_filter form) is checked against the definition itself: for every
    interior pixel the 7x7 patches' means, UNBIASED variances and covariance (np.cov, ddof = 1 -- skimage's
    use_sample_covariance=True default), S = (2 mu_x mu_y + C1)(2 cov + C2) / ((mu_x^2 + mu_y^2 + C1)(var_x + var_y + C2)),
    averaged over the interior and then over channels (channel_axis = 2)."""
    rng = np.random.default_rng(3)
    a = rng.uniform(size=(19, 23, 3))
    b = np.clip(a + 0.08 * rng.normal(size=a.shape), 0, 1)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    per_channel = []
    for ch in range(3):
        vals = []
        for i in range(3, a.shape[0] - 3):
            for j in range(3, a.shape[1] - 3):
                pa = a[i - 3:i + 4, j - 3:j + 4, ch].reshape(-1)
                pb = b[i - 3:i + 4, j - 3:j + 4, ch].reshape(-1)
                cov = np.cov(pa, pb, ddof=1)
                mx, my = pa.mean(), pb.mean()
                vals.append((2 * mx * my + C1) * (2 * cov[0, 1] + C2) / ((mx * mx + my * my + C1) * (cov[0, 0] + cov[1, 1] + C2)))
        per_channel.append(np.mean(vals))
    assert io_.ssim(a, b) == pytest.approx(float(np.mean(per_channel)), abs=1e-12)
    assert io_.psnr(a, b) == pytest.approx(10 * np.log10(1.0 / np.mean((a - b) ** 2)), abs=1e-12)


def test_depth20_reference_fixture_pins_oracle_and_weight_regeneration():
    """tests/golden/image_golden_d20.npz holds the UNMODIFIED reference's psgla / pnpula outputs with the full DnCNN (depth 20,
    64 features; weights regenerated here from the seed).  The oracle must reproduce them on the CPU -- which pins, in one go, the
    oracle loops, the operators, the script's default PnP-ULA table (delta ~ 1e-10) and the weight recipe the gpu tests use."""
    G20 = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_golden_d20.npz"))
    seed, n_pi, spatial = (int(v) for v in G20["weights"])
    den = io_.DnCNN()
    den.load_state_dict(io_.make_dncnn_weights(seed=seed, n_power_iter=n_pi, spatial=spatial))
    den.eval()
    im = torch.from_numpy(G20["im"])
    inp = io_.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    assert torch.equal(inp["mask"], torch.from_numpy(G20["inp.mask"])) and torch.equal(inp["y"], torch.from_numpy(G20["inp.y"]))
    alpha, lambd, s, delta, n_iter, n_inter, n_mm = G20["psgla.params"]
    X, M, M2 = io_.psgla(inp["init"], inp["data_grad"], den, torch.tensor(float(alpha)), torch.tensor(float(lambd)), float(s),
                         float(delta), n_iter=int(n_iter), n_inter=int(n_inter), n_inter_mmse=int(n_mm), seed=7, device="cpu")
    # 20 fp32 conv layers: thread count / SIMD width of the host may reorder sums, so 1e-6 instead of bit equality
    for got, want in ((X, G20["psgla.X"]), (M, G20["psgla.M"]), (M2, G20["psgla.M2"])):
        assert len(got) == want.shape[0]
        assert np.abs(torch.stack(got).numpy() - want).max() < 1e-6
    deb = io_.make_deblurring(im, l=2, blur_type="gaussian", si=1.0, sigma=1.0, seed_ip=0)
    delta, lambd, alpha, s1, s2, n_iter, n_inter, n_mm, l = G20["ula.params"]
    pu = io_.resolve_params("pnp_ula")
    assert (pu["delta"], pu["lambd"], pu["s1"], pu["s2"]) == (float(delta), float(lambd), float(s1), float(s2))
    assert 1.0e-10 < delta < 1.1e-10  # the script's default: s = 2/255 divided by 255 once more (sampling_images.py:149-153)
    pg = io_.make_prior_grad(den, float(alpha), float(s1), float(s2))
    X, M, M2 = io_.pnpula(deb["init"], deb["data_grad"], pg, torch.tensor(float(delta), dtype=torch.float32),
                          torch.tensor(float(lambd), dtype=torch.float32), n_iter=int(n_iter), n_inter=int(n_inter),
                          n_inter_mmse=int(n_mm), seed=11, device="cpu")
    for got, want in ((X, G20["ula.X"]), (M, G20["ula.M"]), (M2, G20["ula.M2"])):
        assert len(got) == want.shape[0]
        assert np.abs(torch.stack(got).numpy() - want).max() < 1e-6
