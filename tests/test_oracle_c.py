"""not-gpu: the plain-C restatement of the 2D path (oracle/gmm2d_oracle.c) against the NumPy oracle (itself pinned to the
unmodified reference) and against the golden vectors generated from the reference (tests/golden/gmm2d_golden.json).
Tolerance: both are float64 and follow the reference's operation order; 2x2 inverses are closed-form in C and LAPACK in
NumPy, so agreement is to ~1e-12 relative, not bit-exact."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as c
from oracle import gmm2d_oracle as o

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gmm2d_golden.json")))


def test_c_oracle_builds_and_exports():
    h = c.lib()
    for fn in ("gmm2d_denoise", "gmm2d_pnp_ula", "gmm2d_snopnp_ula", "gmm2d_run_chain"):
        assert hasattr(h, fn)


@pytest.mark.parametrize("name", o.PRIOR_NAMES)
def test_c_denoiser_matches_numpy_oracle(name):
    prior = o.gaussian_mixt_example(name)
    D = o.theorical_mmse(*prior)
    rng = np.random.default_rng(0)
    for eps in (0.1, 0.3, 0.5, 2.0):
        for _ in range(50):
            x = rng.normal(size=2) * 4
            want = D(x, eps)
            got = c.denoise(*prior, x, eps)
            assert np.allclose(got, want, rtol=1e-11, atol=1e-12), (name, eps, x, got, want)


def test_c_denoiser_golden_points_from_the_reference():
    # SURVEY.md section 8(c) golden vectors (generated from the reference code, float64)
    sym = o.gaussian_mixt_example("symetric_gaussians")
    dis = o.gaussian_mixt_example("disymmetric_gaussians")
    cro = o.gaussian_mixt_example("cross")
    cases = [(sym, (1, 2), 0.3, (2.4155574579, 3.0616680901)), (sym, (-6, 6), 0.3, (-3.8766637928, 3.8766637928)),
             (sym, (0.3, -4), 0.5, (-1.895331879, -4.4142135608)), (dis, (0, 0), 0.3, (0, 1.0616582701)),
             (dis, (1, 2), 0.5, (0.5857864376, 2.4142135624)), (cro, (1, 2), 0.3, (0.5378260051, 1.4934134538)),
             (cro, (0.3, -4), 0.5, (-0.6603325426, -2.7757174641))]
    for prior, x, eps, want in cases:
        got = c.denoise(*prior, np.array(x, dtype=float), eps)
        assert np.allclose(got, want, atol=2e-9), (x, eps, got, want)


@pytest.mark.parametrize("name", o.PRIOR_NAMES)
@pytest.mark.parametrize("yi", [0, 1, 2])
def test_c_samplers_match_numpy_oracle_under_replayed_noise(name, yi):
    prior = o.gaussian_mixt_example(name)
    D = o.theorical_mmse(*prior)
    y = np.asarray(o.OBSERVATIONS[yi], dtype=float)
    A = np.eye(2)
    N = 400
    noise = np.random.default_rng(10 * yi + len(name)).standard_normal((N - 1, 2))
    want = o.snopnp_ula(N, y, y, o.PSGLA_DELTA, A, 1, D, o.PSGLA_ALPHA, noise=noise)
    got = c.snopnp_ula(N, y, y, o.PSGLA_DELTA, A, 1, prior, o.PSGLA_ALPHA, noise)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9), np.abs(got - want).max()
    want = o.pnp_ula(N, y, y, o.ULA_DELTA, A, 1, D, o.ULA_EPSILON, o.ULA_ALPHA, noise=noise)
    got = c.pnp_ula(N, y, y, o.ULA_DELTA, A, 1, prior, o.ULA_EPSILON, o.ULA_ALPHA, noise)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9), np.abs(got - want).max()


def test_c_samplers_against_golden_trajectories_of_the_reference():
    """Every trajectory fixture tests/golden/make_golden.py made by running the unmodified reference (np.random.seed stream
    replayed as `noise`), including the non-identity A / sigma = 1.5 case."""
    n = 0
    for g in GOLD["trajectories"]:
        prior = o.gaussian_mixt_example(g["prior"])
        y = np.asarray(g["y"], dtype=float)
        noise = np.asarray(g["noise"], dtype=float)
        want = np.asarray(g["X"], dtype=float)
        A = np.asarray(g.get("A", np.eye(2)), dtype=float)
        sigma = g.get("sigma", 1)
        N = g["N"]
        if g["alg"] == "psgla":
            delta, alpha = g.get("params", [o.PSGLA_DELTA, o.PSGLA_ALPHA])
            got = c.snopnp_ula(N, y, y, delta, A, sigma, prior, alpha, noise)
        else:
            delta, eps, alpha = g.get("params", [o.ULA_DELTA, o.ULA_EPSILON, o.ULA_ALPHA])
            got = c.pnp_ula(N, y, y, delta, A, sigma, prior, eps, alpha, noise)
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9), (g["prior"], g["alg"], np.abs(got - want).max())
        n += 1
    assert n == 20


def test_c_denoiser_against_golden_points_of_the_reference():
    for g in GOLD["denoiser"]:
        prior = o.gaussian_mixt_example(g["prior"])
        got = c.denoise(*prior, np.asarray(g["x"], dtype=float), g["eps"])
        want = np.asarray(g["D"], dtype=float)
        if np.isfinite(want).all():
            assert np.allclose(got, want, rtol=1e-10, atol=1e-12), (g, got)
        else:  # far from every mode the reference's plain exp underflows to 0 / 0; the C restatement must do the same
            assert not np.isfinite(got).all()


def test_c_run_chain_population_matches_the_numpy_oracle_population():
    """The compiled baseline draws its own noise (only used for timing), so it is checked in distribution: the final states
    of many chains against the NumPy oracle's, same prior / observation / step count.  PSGLA with delta = 0.3 is biased, so
    the oracle population, not the exact posterior, is the yardstick: means within 4 standard errors, spreads within 15 %."""
    prior = o.gaussian_mixt_example("symetric_gaussians")
    D = o.theorical_mmse(*prior)
    y = np.array([0.0, -2.0])
    rng = np.random.default_rng(0)
    ref = np.array([o.snopnp_ula(150, y, y, o.PSGLA_DELTA, np.eye(2), 1, D, o.PSGLA_ALPHA,
                                 noise=rng.standard_normal((149, 2)))[-1] for _ in range(150)])
    got = np.array([c.run_chain("psgla", 150, y, y, o.PSGLA_DELTA, np.eye(2), 1, prior, 1.0, o.PSGLA_ALPHA, seed=s)
                    for s in range(4000)])
    assert np.isfinite(got).all()
    se = ref.std(0) / np.sqrt(len(ref))
    assert (np.abs(got.mean(0) - ref.mean(0)) < 4 * se).all(), (got.mean(0), ref.mean(0), se)
    assert (np.abs(got.std(0) / ref.std(0) - 1) < 0.15).all(), (got.std(0), ref.std(0))


@pytest.mark.reference
def test_c_oracle_against_the_live_reference():
    """In the build container: the UNMODIFIED reference (sampling_2D.py / utils_2D.py, loaded by oracle/ref_loader.py) run
    next to the C restatement on the global NumPy stream -- denoiser on random points and 300-step trajectories of both
    samplers for all three priors, plus a non-identity A with sigma != 1."""
    from oracle import ref_loader
    u2d = ref_loader.load_utils_2D()
    s2d = ref_loader.load_sampling_2D()
    rng = np.random.default_rng(3)
    for name in o.PRIOR_NAMES:
        mu, Sig, pi = u2d.gaussian_mixt_example(name)
        prior = o.gaussian_mixt_example(name)
        Dr = u2d.Theorical_MMSE(mu, Sig, pi)
        for eps in (0.05, 0.3, 0.5):
            for _ in range(40):
                x = rng.uniform(-7, 7, size=2)
                assert np.allclose(c.denoise(*prior, x, eps), Dr(x, eps), rtol=1e-11, atol=1e-12)
        for y, A, sigma in ((np.array([0.0, -2.0]), np.eye(2), 1), (np.array([0.5, -1.0]), np.array([[2.0, 0.0], [0.3, 1.0]]), 1.5)):
            N = 300
            np.random.seed(11)
            Xr = s2d.SnoPnP_ULA(N, y, y, 0.3, A, sigma, Dr, 2 / 3)
            np.random.seed(11)
            noise = np.random.randn(N - 1, 2)
            assert np.allclose(c.snopnp_ula(N, y, y, 0.3, A, sigma, prior, 2 / 3, noise), Xr, rtol=1e-9, atol=1e-9)
            np.random.seed(12)
            Xr = s2d.PnP_ULA(N, y, y, 0.1, A, sigma, Dr, 0.5, 1.5)
            np.random.seed(12)
            noise = np.random.randn(N - 1, 2)
            assert np.allclose(c.pnp_ula(N, y, y, 0.1, A, sigma, prior, 0.5, 1.5, noise), Xr, rtol=1e-9, atol=1e-9)
