"""not-gpu: host-side logic of the product package (posterior constants, metrics, sharding, operators' callable form)."""
import json
import os

import numpy as np
import pytest
import torch

import psgla_b200 as P
from oracle import gmm2d_oracle as o
from oracle import image_oracle as io_

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gmm2d_golden.json")))


def test_priors_match_reference_table():
    for name in o.PRIOR_NAMES:
        a, b = P.gaussian_mixt_example(name), o.gaussian_mixt_example(name)
        for x, y in zip(a, b):
            assert np.allclose(np.asarray(x, dtype=float), np.asarray(y, dtype=float))


def test_posterior_constants_golden():
    for g in GOLD["posterior"]:
        mu, Sig, pi = P.gaussian_mixt_example(g["prior"])
        m, S, p = P.constantes_conditionnal_prob(np.eye(2), np.array(g["y"]), 1, mu, Sig, pi)
        assert np.allclose(np.array(m), np.array(g["mu"]), atol=1e-10)
        assert np.allclose(np.array(S), np.array(g["Sigma"]), atol=1e-10)
        assert np.allclose(p, g["p"], atol=1e-10)


def test_sample_posterior_same_stream_as_oracle():
    mu, Sig, pi = P.gaussian_mixt_example("disymmetric_gaussians")
    y = np.array([0, -2])
    np.random.seed(3)
    a = P.sample_posterior(np.eye(2), y, 1, 400, mu, Sig, pi)
    np.random.seed(3)
    b = o.sample_posterior(np.eye(2), y, 1, 400, mu, Sig, pi)
    assert np.allclose(a, b, atol=1e-10)


def test_wasserstein_matches_oracle_and_translation():
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((300, 2)), rng.standard_normal((300, 2)) + 1.0
    w1 = P.Wasserstein_distance(a, b, rng=np.random.default_rng(5))
    w2 = o.wasserstein_distance(a, b, rng=np.random.default_rng(5))
    assert w1 == pytest.approx(w2, rel=1e-12)
    assert P.Wasserstein_distance(a, a + np.array([0.0, 2.0]), rng=np.random.default_rng(0)) == pytest.approx(4.0, abs=0.5)
    assert P.sliced_wasserstein_distance(a, a) == 0.0


def test_shard_range_tiles():
    for n in (0, 1, 7, 68, 10 ** 6):
        for ws in (1, 2, 3, 8):
            blocks = [P.dist.shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_structured_operators_are_callable_like_the_reference_closures():
    torch.manual_seed(0)
    im = torch.rand(1, 3, 12, 12)
    ref = io_.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    dg, init, y, mask = P.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    assert torch.equal(mask, ref["mask"]) and torch.equal(y, ref["y"]) and torch.equal(init, ref["init"])
    x = torch.rand(1, 3, 12, 12)
    assert torch.equal(dg(x), ref["data_grad"](x))
    # the blur is a CUDA stencil: the structured deblurring operator has no CPU path (parity with the reference's conv2d
    # formulation is a gpu test, tests/test_image_gpu.py::test_blur_against_reference_formulation)
    with pytest.raises(RuntimeError, match="no CPU path"):
        P.make_deblurring(im, l=2, blur_type="gaussian", si=1.0)
    # the reference's own psgla accepts the structured callable (CPU, tiny)
    den = io_.DnCNN(depth=3, nf=4)
    a = io_.psgla(init, dg, den, torch.tensor(1.0), torch.tensor(5.0), 2 / 255, (2 / 255) ** 2, n_iter=3, n_inter=1, n_inter_mmse=1, seed=0)
    b = io_.psgla(ref["init"], ref["data_grad"], den, torch.tensor(1.0), torch.tensor(5.0), 2 / 255, (2 / 255) ** 2, n_iter=3, n_inter=1, n_inter_mmse=1, seed=0)
    assert all(torch.equal(u, v) for u, v in zip(a[0], b[0]))


def test_opaque_callables_are_rejected():
    with pytest.raises(TypeError):
        P.pnpula(torch.zeros(1, 3, 8, 8), lambda x: x, lambda x: x, 1e-6, 1e-6, seed=0)
    with pytest.raises(ValueError):
        P.DeblurDataGrad(np.array([0.2, 0.3, 0.5]), 1, torch.zeros(1, 3, 4, 4), 1.0)  # non-symmetric taps


def test_sampler_params_reproduce_the_script_table():
    """psgla_b200.sampler_params against the oracle's restatement of sampling_images.py:100-123,147-198 (itself pinned by
    tests/test_oracle_image.py::test_resolve_params_table), including the quirks."""
    from oracle import image_oracle as io_
    cases = [("psgla", "DnCNN", {}), ("psgla", "DnCNN", dict(s=4.0, lambd=7.0)), ("psgla", "DRUNet", {}),
             ("psgla", "DRUNet", dict(s=3.0, lambd=25.0, N=2000)), ("pnp_ula", "DnCNN", {}), ("pnp_ula", "DnCNN", dict(s=5.0)),
             ("pnp_ula", "DRUNet", dict(N=30000)), ("pnp_ula", "DnCNN", dict(N=100000))]
    for alg, den, kw in cases:
        got = P.sampler_params(alg, den=den, **kw)
        okw = dict(kw)
        n_given = "N" in okw
        want = io_.resolve_params(alg, den=den, N=okw.pop("N", 10000), N_given=n_given, **okw)
        for k in ("s", "lambd", "delta", "N", "n_inter", "n_inter_mmse", "sigma2"):
            assert got[k] == pytest.approx(want[k], rel=1e-15), (alg, den, kw, k)
    p = P.sampler_params("pnp_ula")  # the double division by 255 and the thinning computed from the parsed N
    assert p["s1"] == pytest.approx(2 / 255 / 255) and p["N"] == 100000 and p["n_inter"] == 10 and (p["c_min"], p["c_max"]) == (-1, 2)
    assert P.sampler_params("psgla", den="DRUNet")["delta"] / P.sampler_params("psgla", den="DRUNet")["lambd"] / (1 / 255) ** 2 == pytest.approx(25.0)
    assert set(P.as_psgla_kwargs(P.sampler_params("psgla"))) == {"alpha", "lambd", "sig_float", "delta", "seed", "n_iter", "n_inter", "n_inter_mmse"}


def test_torch_cuda_randn_policy_arithmetic():
    """psgla_torch_cuda_randn_policy restates ATen's calc_execution_policy (block 256, unroll 4, grid capped at
    SMs x resident blocks); host code, checked here against hand-computed cases for a 148-SM / 2048-thread device."""
    import ctypes as C
    lib = P._lib.lib()
    t, s = C.c_uint32(), C.c_uint64()
    cases = {1: (256, 4), 256: (256, 4), 257: (512, 4), 3 * 256 * 256: (196608, 4), 148 * 8 * 256: (303104, 4),
             148 * 8 * 256 * 4: (303104, 4), 148 * 8 * 256 * 4 + 1: (303104, 8), 32 * 3 * 256 * 256: (303104, 24)}
    for numel, want in cases.items():
        assert lib.psgla_torch_cuda_randn_policy(numel, 148, 2048, C.byref(t), C.byref(s)) == 0
        assert (t.value, s.value) == want, numel
    assert lib.psgla_torch_cuda_randn_policy(0, 148, 2048, C.byref(t), C.byref(s)) == -1


def test_wasserstein_assignment_equals_the_transport_lp_on_small_clouds():
    """POT's ``ot.emd2`` (utils_2D.py:242-243) is absent; the product and the oracle replace it by the assignment problem.
    For equal-size uniform clouds the transport polytope's vertices are permutation matrices (Birkhoff), so the LP optimum is
    the best permutation: checked here by brute force over all n! permutations for n <= 6, and against scipy's LP solver."""
    import itertools
    from scipy.optimize import linprog
    rng = np.random.default_rng(0)
    for n in (2, 3, 5, 6):
        a, b = rng.normal(size=(n, 2)), rng.normal(size=(n, 2)) + 0.5
        M = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
        brute = min(sum(M[i, p[i]] for i in range(n)) for p in itertools.permutations(range(n))) / n
        # transport LP: min <M, T>, T 1 = 1/n, T^T 1 = 1/n, T >= 0
        A_eq = np.zeros((2 * n, n * n))
        for i in range(n):
            A_eq[i, i * n:(i + 1) * n] = 1
            A_eq[n + i, i::n] = 1
        lp = linprog(M.reshape(-1), A_eq=A_eq, b_eq=np.full(2 * n, 1.0 / n), bounds=(0, None), method="highs")
        assert lp.status == 0
        got_p = P.Wasserstein_distance(a, b, n_sub=n, rng=np.random.default_rng(1))
        got_o = o.wasserstein_distance(a, b, n_sub=n, rng=np.random.default_rng(1)) if hasattr(o, "wasserstein_distance") else got_p
        assert abs(got_p - brute) < 1e-12 and abs(lp.fun - brute) < 1e-9 and abs(got_o - brute) < 1e-12


def test_sliced_wasserstein_is_exact_for_translations_and_permutation_invariant():
    """1-D optimal transport between equal-size clouds pairs sorted projections: a translated copy has sliced W2 = |t| E|cos|-free
    closed form per direction (sqrt(mean (theta . t)^2)), and shuffling either cloud changes nothing."""
    rng = np.random.default_rng(2)
    X = rng.normal(size=(500, 2))
    t = np.array([0.7, -1.3])
    Y = X + t
    got = P.sliced_wasserstein_distance(X, Y, n_projections=64, seed=5)
    th = np.random.default_rng(5).standard_normal((2, 64))
    th /= np.linalg.norm(th, axis=0, keepdims=True)
    want = np.sqrt(np.mean((t @ th) ** 2))
    assert abs(got - want) < 1e-12
    assert abs(P.sliced_wasserstein_distance(X[rng.permutation(500)], Y[rng.permutation(500)], 64, seed=5) - got) < 1e-12


def test_wasserstein_unequal_sizes_solves_the_transport_lp():
    """ot.emd2 with uniform weights is the transport LP for any pair of sizes (utils_2D.py:242-243); sample_posterior returns
    sum_i int(pi_i N) points, so 99-against-100 clouds do occur.  Equal sizes: the LP equals the assignment problem; unequal:
    a hand-checked 2-against-3 instance and consistency with the equal-size value when one point is dropped."""
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(60, 2)), rng.normal(size=(60, 2)) + 0.5
    M = ((a[:, None] - b[None]) ** 2).sum(-1)
    eq = P.Wasserstein_distance(a, b, rng=np.random.default_rng(1))
    assert abs(P.utils_2D._uniform_transport_cost(M) - eq) < 1e-9
    x = np.array([[0.0, 0.0], [1.0, 0.0]])
    y = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    assert abs(P.Wasserstein_distance(x, y, rng=np.random.default_rng(0)) - 0.5) < 1e-9  # 1/6 and 1/3 of the mass move one unit
    un = P.Wasserstein_distance(a[:59], b, rng=np.random.default_rng(1))
    assert abs(un - eq) < 0.1 * eq


def test_compiled_reference_loader_and_reference_arm(tmp_path):
    """oracle/_ref/*.pycode (the reference's modules compiled by build(), what the GPU box loads) must behave like the sources:
    with /root/reference hidden, the golden PSGLA trajectory row still comes out, and `bench.py --impl reference` (the CPU arm
    the driver runs) prints one JSON line of kind "reference"."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from oracle import ref_loader
    if not ref_loader._compiled_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    env = dict(os.environ, PSGLA_REFERENCE_ROOT=str(tmp_path / "nowhere"))
    code = ("import numpy as np\n"
            "from oracle import ref_loader as r\n"
            "assert r.reference_kind() == 'compiled'\n"
            "ns, u = r.load_sampling_2D(), r.load_utils_2D()\n"
            "D = u.Theorical_MMSE(*u.gaussian_mixt_example('cross'))\n"
            "np.random.seed(0)\n"
            "X = ns.SnoPnP_ULA(6, np.array([0., -2.]), np.array([0., -2.]), 0.3, np.eye(2), 1, D, 2 / 3)\n"
            "assert np.allclose(X[1], [0.1366241292, -0.6247715824], atol=1e-9), X[1]\n"
            "assert callable(r.load_restoration_algorithms().psgla)\n")
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-chain-steps", "300",
                          "--skip-image"], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == "langevin_chain_steps_per_sec_2d_gmm"
