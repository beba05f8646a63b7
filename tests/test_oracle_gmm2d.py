"""Pins oracle/gmm2d_oracle.py against the reference: committed golden vectors + (in the container) the live reference."""
import json
import os

import numpy as np
import pytest

from oracle import gmm2d_oracle as o
from oracle import ref_loader

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gmm2d_golden.json")))


def test_noise_stream_head():
    np.random.seed(0)
    assert np.allclose(np.random.randn(4), GOLD["noise_head"], rtol=0, atol=0)


def test_denoiser_golden():
    for g in GOLD["denoiser"]:
        D = o.theorical_mmse(*o.gaussian_mixt_example(g["prior"]))
        v = D(np.array(g["x"]), g["eps"])
        assert np.allclose(v, g["D"], rtol=1e-13, atol=1e-13), g


def test_folded_denoiser_equals_reference_form():
    rng = np.random.default_rng(0)
    for name in o.PRIOR_NAMES:
        mu, Sig, pi = o.gaussian_mixt_example(name)
        D = o.theorical_mmse(mu, Sig, pi)
        for eps in (0.3, 0.5, 0.05):
            c = o.folded_constants(mu, Sig, pi, eps)
            x = rng.uniform(-7, 7, size=(64, 2))
            got = o.folded_denoiser(x, c)
            want = np.array([D(xi, eps) for xi in x])
            assert np.allclose(got, want, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("idx", range(len(GOLD["trajectories"])))
def test_trajectory_golden(idx):
    g = GOLD["trajectories"][idx]
    mu, Sig, pi = o.gaussian_mixt_example(g["prior"])
    D = o.theorical_mmse(mu, Sig, pi)
    y = np.array(g["y"], dtype=float)
    A = np.array(g.get("A", np.eye(2)))
    sigma = g.get("sigma", 1)
    noise = np.array(g["noise"])
    if g["alg"] == "psgla":
        delta, alpha = g.get("params", [o.PSGLA_DELTA, o.PSGLA_ALPHA])
        X = o.snopnp_ula(g["N"], y, y, delta, A, sigma, D, alpha, noise=noise)
        Xb = o.run_chains("psgla", g["N"] - 1, y[None], y, delta, A, sigma, mu, Sig, pi, alpha, noise=noise[:, None, :], thin=1)[1]
    else:
        delta, eps, alpha = g.get("params", [o.ULA_DELTA, o.ULA_EPSILON, o.ULA_ALPHA])
        X = o.pnp_ula(g["N"], y, y, delta, A, sigma, D, eps, alpha, noise=noise)
        Xb = o.run_chains("pnp_ula", g["N"] - 1, y[None], y, delta, A, sigma, mu, Sig, pi, alpha, epsilon=eps, noise=noise[:, None, :], thin=1)[1]
    want = np.array(g["X"])
    assert X.shape == want.shape == (g["N"], 2)
    assert np.allclose(X, want, rtol=1e-12, atol=1e-12)
    assert np.allclose(Xb[:, 0, :], want[1:], rtol=1e-10, atol=1e-10)  # folded-constant batched form


def test_posterior_constants_golden():
    for g in GOLD["posterior"]:
        mu, Sig, pi = o.gaussian_mixt_example(g["prior"])
        m, S, p = o.constantes_conditionnal_prob(np.eye(2), np.array(g["y"]), 1, mu, Sig, pi)
        assert np.allclose(np.array(m), np.array(g["mu"]), atol=1e-12)
        assert np.allclose(np.array(S), np.array(g["Sigma"]), atol=1e-12)
        assert np.allclose(p, g["p"], atol=1e-12)


def test_wasserstein_restatement_sanity():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((1000, 2))
    assert o.wasserstein_distance(a, a, rng=rng) < 1e-12
    shift = o.wasserstein_distance(a, a + np.array([3.0, 0.0]), rng=rng)
    assert abs(shift - 9.0) < 1e-9  # W2^2 of a pure translation


@pytest.mark.reference
def test_live_reference_trajectory_and_unseeded_stream():
    """Run the unmodified reference next to the oracle on the global NumPy stream (no replay)."""
    u2d = ref_loader.load_utils_2D()
    s2d = ref_loader.load_sampling_2D()
    for name in o.PRIOR_NAMES:
        mu, Sig, pi = u2d.gaussian_mixt_example(name)
        Dr = u2d.Theorical_MMSE(mu, Sig, pi)
        Do = o.theorical_mmse(*o.gaussian_mixt_example(name))
        y = np.array([0, -2])
        np.random.seed(5)
        Xr = s2d.SnoPnP_ULA(200, y, y, 0.3, np.eye(2), 1, Dr, 2 / 3)
        np.random.seed(5)
        Xo = o.snopnp_ula(200, y, y, 0.3, np.eye(2), 1, Do, 2 / 3)
        assert np.array_equal(Xr, Xo)
        np.random.seed(6)
        Xr = s2d.PnP_ULA(200, y, y, 0.1, np.eye(2), 1, Dr, 0.5, 1.5)
        np.random.seed(6)
        Xo = o.pnp_ula(200, y, y, 0.1, np.eye(2), 1, Do, 0.5, 1.5)
        assert np.array_equal(Xr, Xo)
        np.random.seed(7)
        Pr = u2d.sample_posterior(np.eye(2), y, 1, 500, mu, Sig, pi)
        np.random.seed(7)
        Po = o.sample_posterior(np.eye(2), y, 1, 500, *o.gaussian_mixt_example(name))
        assert np.allclose(Pr, Po, atol=1e-10)
