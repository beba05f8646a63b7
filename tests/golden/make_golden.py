"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/gmm2d_golden.json and tests/golden/image_golden.npz.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import image_oracle as io_  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def gmm2d():
    u2d = ref_loader.load_utils_2D()
    s2d = ref_loader.load_sampling_2D()
    out = {"denoiser": [], "trajectories": [], "posterior": [], "noise_head": None}
    pts = [(1.0, 2.0), (-6.0, 6.0), (0.3, -4.0), (0.0, 0.0), (7.5, -7.5), (-0.2, 0.1)]
    for name in ("symetric_gaussians", "cross", "disymmetric_gaussians"):
        mu, Sig, pi = u2d.gaussian_mixt_example(name)
        D = u2d.Theorical_MMSE(mu, Sig, pi)
        for p in pts:
            for eps in (0.3, 0.5, 0.05):
                v = D(np.array(p), eps)
                out["denoiser"].append({"prior": name, "x": list(p), "eps": eps, "D": [float(v[0]), float(v[1])]})
        A = np.eye(2)
        for y in ([0, 0], [0, -2], [-6, 6]):
            y = np.array(y)
            mu_c, sig_c, p_l = u2d.constantes_conditionnal_prob(A, y, 1, mu, Sig, pi)
            out["posterior"].append({"prior": name, "y": y.tolist(),
                                     "mu": [np.asarray(m, dtype=float).tolist() for m in mu_c],
                                     "Sigma": [np.asarray(s, dtype=float).tolist() for s in sig_c],
                                     "p": np.asarray(p_l, dtype=float).tolist()})
            for alg in ("psgla", "pnp_ula"):
                N = 40
                np.random.seed(0)
                if alg == "psgla":
                    X = s2d.SnoPnP_ULA(N, y, y, 0.3, A, 1, D, 2 / 3)
                else:
                    X = s2d.PnP_ULA(N, y, y, 0.1, A, 1, D, 0.5, 1.5)
                np.random.seed(0)
                noise = np.random.randn(N - 1, 2)  # the same stream, drawn 2 at a time (sampling_2D.py:35,62)
                out["trajectories"].append({"prior": name, "y": y.tolist(), "alg": alg, "N": N,
                                            "X": np.asarray(X, dtype=float).tolist(), "noise": noise.tolist()})
    np.random.seed(0)
    out["noise_head"] = np.random.randn(4).tolist()
    # a non-identity A and sigma != 1 (the commented alternative at sampling_2D.py:83)
    mu, Sig, pi = u2d.gaussian_mixt_example("cross")
    D = u2d.Theorical_MMSE(mu, Sig, pi)
    A = np.array([[2.0, 0.0], [0.3, 1.0]])
    y = np.array([0.5, -1.0])
    for alg in ("psgla", "pnp_ula"):
        np.random.seed(1)
        X = s2d.SnoPnP_ULA(25, y, y, 0.05, A, 1.5, D, 0.5) if alg == "psgla" else s2d.PnP_ULA(25, y, y, 0.02, A, 1.5, D, 0.4, 1.2)
        np.random.seed(1)
        noise = np.random.randn(24, 2)
        out["trajectories"].append({"prior": "cross", "y": y.tolist(), "alg": alg, "N": 25, "A": A.tolist(), "sigma": 1.5,
                                    "params": ([0.05, 0.5] if alg == "psgla" else [0.02, 0.4, 1.2]),
                                    "X": np.asarray(X, dtype=float).tolist(), "noise": noise.tolist()})
    with open(os.path.join(HERE, "gmm2d_golden.json"), "w") as fh:
        json.dump(out, fh)
    print("gmm2d:", len(out["denoiser"]), "denoiser points,", len(out["trajectories"]), "trajectories")


def images():
    """Reference psgla / pnpula (restoration_algorithms.py) on a 16x16 colour image, CPU, with a small
    seeded DnCNN-shaped denoiser restated in torch (deepinv is absent).  The fixtures pin the sampler loops,
    the thinning / running-statistics bookkeeping and the operators -- not deepinv."""
    ra = ref_loader.load_restoration_algorithms()
    torch.manual_seed(0)
    H = W = 16
    im = torch.rand(1, 3, H, W)
    den = io_.DnCNN(depth=4, nf=8)
    torch.manual_seed(1)
    for p in den.parameters():
        torch.nn.init.normal_(p, std=0.05)
    den.eval()
    out = {"im": im.numpy(), **{"den." + k: v.numpy() for k, v in den.state_dict().items()}}

    inp = io_.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    prm = io_.resolve_params("psgla")
    n_iter, n_inter, n_mm = 24, 3, 4
    Xl, Xm, Xm2 = ra.psgla(init=inp["init"], data_grad=inp["data_grad"], denoiser=den,
                           alpha=torch.tensor(0.8), lambd=torch.tensor(prm["lambd"]), sig_float=prm["s"],
                           delta=prm["delta"], seed=0, device="cpu", n_iter=n_iter, n_inter=n_inter, n_inter_mmse=n_mm)
    gen = torch.Generator().manual_seed(0)
    noise = torch.stack([torch.randn(im.shape, generator=gen) for _ in range(n_iter)])
    out.update({"inp.mask": inp["mask"].numpy(), "inp.y": inp["y"].numpy(), "inp.init": inp["init"].numpy(),
                "psgla.noise": noise.numpy(), "psgla.X": torch.stack(Xl).numpy(), "psgla.M": torch.stack(Xm).numpy(),
                "psgla.M2": torch.stack(Xm2).numpy(),
                "psgla.params": np.array([0.8, prm["lambd"], prm["s"], prm["delta"], n_iter, n_inter, n_mm])})

    deb = io_.make_deblurring(im, l=2, blur_type="gaussian", si=1.0, sigma=1.0, seed_ip=0)
    s1, s2 = 5.0 / 255.0, (5.0 / 255.0) ** 2
    sigma2 = deb["sigma2"]
    alpha = 1.0
    lam = 0.5 / (2 / sigma2 + alpha / s2)
    delta = 1 / 3 / (1 / sigma2 + 1 / lam + alpha / s2)
    pg = io_.make_prior_grad(den, alpha, s1, s2)
    Xl, Xm, Xm2 = ra.pnpula(init=deb["init"], data_grad=deb["data_grad"], prior_grad=pg,
                            delta=torch.tensor(delta, dtype=torch.float32), lambd=torch.tensor(lam, dtype=torch.float32),
                            seed=3, device="cpu", n_iter=n_iter, n_inter=n_inter, n_inter_mmse=n_mm)
    gen = torch.Generator().manual_seed(3)
    noise = torch.stack([torch.randn(im.shape, generator=gen) for _ in range(n_iter)])
    out.update({"deb.y": deb["y"].numpy(), "deb.h": deb["h"], "deb.Ax": deb["A"](im).numpy(),
                "deb.grad_at_im": deb["data_grad"](im).numpy(),
                "ula.noise": noise.numpy(), "ula.X": torch.stack(Xl).numpy(), "ula.M": torch.stack(Xm).numpy(),
                "ula.M2": torch.stack(Xm2).numpy(),
                "ula.params": np.array([delta, lam, alpha, s1, s2, n_iter, n_inter, n_mm, 2])})
    # PnP forward-backward and RED (restoration_algorithms.py:386-529) with a sigma-sensitive test denoiser, so that the
    # sigma-annealing schedules are pinned too
    sden = io_.SigmaBlendDenoiser(den)
    Xl, Xf, _ = ra.pnp(init=inp["init"], data_grad=inp["data_grad"], Pb="inpainting", denoiser=sden, alpha=torch.tensor(0.9),
                       lambd=torch.tensor(5.0), sig_float=prm["s"], delta=prm["delta"], n_iter=30, device="cpu")
    out.update({"pnp.X": torch.stack(Xl).numpy(), "pnp.params": np.array([0.9, 5.0, prm["s"], prm["delta"], 30])})
    Xl, Xf, _ = ra.red(init=inp["init"], data_grad=inp["data_grad"], Pb="inpainting", denoiser=sden, lambd=torch.tensor(2000.0),
                       sig_float=prm["s"], delta=2e-5, n_iter=14, device="cpu")
    out.update({"red.X": torch.stack(Xl).numpy(), "red.params": np.array([2000.0, prm["s"], 2e-5, 14])})
    np.savez_compressed(os.path.join(HERE, "image_golden.npz"), **out)
    print("images: psgla", out["psgla.X"].shape, out["psgla.M"].shape, "ula", out["ula.X"].shape, out["ula.M"].shape)


def images_depth20():
    """The same two reference samplers with the FULL denoiser architecture the CUDA path is built for (DnCNN depth 20, 64
    features), so that the reference's outputs can be compared with the kernels directly (tests/test_image_gpu.py::
    test_reference_fixture_replayed_through_cuda).  The weights are not stored: the tests regenerate them from the seed with
    oracle.image_oracle.make_dncnn_weights(seed=0, n_power_iter=10, spatial=16), exactly as here."""
    ra = ref_loader.load_restoration_algorithms()
    sd = io_.make_dncnn_weights(seed=0, n_power_iter=10, spatial=16)
    den = io_.DnCNN()
    den.load_state_dict(sd)
    den.eval()
    torch.manual_seed(5)
    H = W = 16
    im = torch.rand(1, 3, H, W)
    out = {"im": im.numpy(), "weights": np.array([0, 10, 16])}  # seed, n_power_iter, spatial
    n_iter, n_inter, n_mm = 12, 2, 3

    inp = io_.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    prm = io_.resolve_params("psgla")
    Xl, Xm, Xm2 = ra.psgla(init=inp["init"], data_grad=inp["data_grad"], denoiser=den, alpha=torch.tensor(1.0),
                           lambd=torch.tensor(prm["lambd"]), sig_float=prm["s"], delta=prm["delta"], seed=7, device="cpu",
                           n_iter=n_iter, n_inter=n_inter, n_inter_mmse=n_mm)
    gen = torch.Generator().manual_seed(7)
    noise = torch.stack([torch.randn(im.shape, generator=gen) for _ in range(n_iter)])
    out.update({"inp.mask": inp["mask"].numpy(), "inp.y": inp["y"].numpy(), "inp.init": inp["init"].numpy(),
                "psgla.noise": noise.numpy(), "psgla.X": torch.stack(Xl).numpy(), "psgla.M": torch.stack(Xm).numpy(),
                "psgla.M2": torch.stack(Xm2).numpy(),
                "psgla.params": np.array([1.0, prm["lambd"], prm["s"], prm["delta"], n_iter, n_inter, n_mm])})

    # PnP-ULA with the script's DEFAULT table (sampling_images.py:147-168: s = 2/255 divided by 255 once more, delta ~ 1e-10)
    deb = io_.make_deblurring(im, l=2, blur_type="gaussian", si=1.0, sigma=1.0, seed_ip=0)
    pu = io_.resolve_params("pnp_ula")
    pg = io_.make_prior_grad(den, pu["alpha"], pu["s1"], pu["s2"])
    Xl, Xm, Xm2 = ra.pnpula(init=deb["init"], data_grad=deb["data_grad"], prior_grad=pg,
                            delta=torch.tensor(pu["delta"], dtype=torch.float32), lambd=torch.tensor(pu["lambd"], dtype=torch.float32),
                            seed=11, device="cpu", n_iter=n_iter, n_inter=n_inter, n_inter_mmse=n_mm)
    gen = torch.Generator().manual_seed(11)
    noise = torch.stack([torch.randn(im.shape, generator=gen) for _ in range(n_iter)])
    out.update({"deb.y": deb["y"].numpy(), "deb.h": deb["h"], "ula.noise": noise.numpy(), "ula.X": torch.stack(Xl).numpy(),
                "ula.M": torch.stack(Xm).numpy(), "ula.M2": torch.stack(Xm2).numpy(),
                "ula.params": np.array([pu["delta"], pu["lambd"], pu["alpha"], pu["s1"], pu["s2"], n_iter, n_inter, n_mm, 2])})
    np.savez_compressed(os.path.join(HERE, "image_golden_d20.npz"), **out)
    print("images depth 20: psgla", out["psgla.X"].shape, out["psgla.M"].shape, "ula", out["ula.X"].shape, out["ula.M"].shape,
          "ula delta %.3g lambd %.3g" % (pu["delta"], pu["lambd"]))


if __name__ == "__main__":
    assert ref_loader.reference_available(), "needs /root/reference"
    if "--only-d20" not in sys.argv:
        gmm2d()
        images()
    images_depth20()
