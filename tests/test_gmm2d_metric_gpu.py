"""-m gpu: the population sliced-W2 metric on the device (csrc/gmm2d_metric.cu) against the NumPy restatement of
``ot.sliced.sliced_wasserstein_distance`` (psgla_b200.sliced_wasserstein_distance, float64), and the radix sort inside it
against ``numpy.sort`` bit for bit.

Tolerance: the device projects in fp32 (the population is fp32), the checker in float64 -> 1e-5 relative on the distance;
the sorted fp32 projections themselves must equal numpy's sort of the same fp32 values exactly."""
import ctypes as C

import numpy as np
import pytest
import torch

import psgla_b200 as P
from oracle import gmm2d_oracle as o

pytestmark = pytest.mark.gpu


def _theta(n_proj, seed):
    rng = np.random.default_rng(seed)
    th = rng.standard_normal((2, n_proj))
    th /= np.linalg.norm(th, axis=0, keepdims=True)
    return np.ascontiguousarray(th.T, dtype=np.float32)


@pytest.mark.parametrize("n,n_proj", [(1, 1), (31, 3), (4096, 2), (4097, 5), (100003, 7), (1 << 20, 4)])
@pytest.mark.parametrize("precision", [0, 1])
def test_sorted_projections_equal_numpy_sort(n, n_proj, precision):
    lib = P._lib.lib()
    rng = np.random.default_rng(n + n_proj)
    x = rng.standard_normal((n, 2)) * 3.0
    x[::7] = 0.0  # ties and signed zeros
    x[1::11, 0] *= -1e-30
    xt = torch.from_numpy(x).cuda().to(torch.float64 if precision else torch.float32).contiguous()
    th = _theta(n_proj, 0)
    nbytes = lib.psgla_gmm2d_sw2_workspace_bytes(n, n_proj)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty((n_proj, n), dtype=torch.float32, device="cuda")
    P._lib.check(lib.psgla_gmm2d_sorted_projections(xt.data_ptr(), precision, n, th.ctypes.data_as(C.POINTER(C.c_float)),
                                                    n_proj, out.data_ptr(), ws.data_ptr(), nbytes, None), "sorted_proj")
    torch.cuda.synchronize()
    x32 = xt.to(torch.float32).cpu().numpy()
    # the kernel's projection: fma(t0, x0, t1 * x1) in fp32
    prod = (th[:, 1:2].astype(np.float32) * x32[None, :, 1]).astype(np.float32)
    want = (th[:, 0:1].astype(np.float64) * x32[None, :, 0].astype(np.float64) + prod.astype(np.float64)).astype(np.float32)
    want = np.sort(want, axis=1)
    got = out.cpu().numpy()
    assert np.array_equal(np.abs(got), np.abs(want)) and np.array_equal(got == 0, want == 0)  # +-0 order is free
    assert (np.diff(got, axis=1) >= 0).all()


@pytest.mark.parametrize("n,n_proj", [(1000, 50), (65536, 50), (300001, 16)])
def test_sliced_w2_against_numpy_restatement(n, n_proj):
    mu, Sig, pi = o.gaussian_mixt_example("disymmetric_gaussians")
    y = np.array([0.0, -2.0])
    D = P.Theorical_MMSE(mu, Sig, pi)
    ch = P.GMMChains("psgla", y, 0.3, np.eye(2), 1, D, 2 / 3, n_chains=n, seed=5)
    ch.run(50)
    post = P.sample_posterior(np.eye(2), y, 1, n, mu, Sig, pi, rng=np.random.default_rng(1))[:n]
    if len(post) < n:  # int(pi * N) truncation of sample_gaussian
        post = np.concatenate([post, post[:n - len(post)]])
    ch.set_reference_sample(post, n_projections=n_proj, seed=3)
    got = float(ch.sliced_w2().item())
    want = P.sliced_wasserstein_distance(ch.state.double().cpu().numpy(), post.astype(np.float32).astype(np.float64),
                                         n_projections=n_proj, seed=3)
    assert abs(got - want) <= 1e-5 * max(want, 1e-3), (got, want)


def test_population_metric_every_100_steps_through_the_drop_in():
    """SnoPnP_ULA(..., compute_metric_each_step=True, n_chains=...) returns the finals and the metric after steps
    1, 101, 201, ...; the values must equal a step-by-step replay with host-side NumPy evaluation, and decrease from the
    x_0 = y start towards the sampler's bias floor."""
    mu, Sig, pi = o.gaussian_mixt_example("symetric_gaussians")
    y = np.array([0.0, -2.0])
    D = P.Theorical_MMSE(mu, Sig, pi)
    n = 20000
    post = P.sample_posterior(np.eye(2), y, 1, n, mu, Sig, pi, rng=np.random.default_rng(2))
    post = np.concatenate([post, post[:n - len(post)]]) if len(post) < n else post[:n]
    fin, W = P.SnoPnP_ULA(302, y, y, 0.3, np.eye(2), 1, D, 2 / 3, Sample_posterior=post, compute_metric_each_step=True,
                          n_chains=n, seed=9)
    assert fin.shape == (n, 2) and len(W) == 4  # after steps 1, 101, 201, 301
    ch = P.GMMChains("psgla", y, 0.3, np.eye(2), 1, D, 2 / 3, n_chains=n, seed=9)
    want, done = [], 0
    for t in (1, 101, 201, 301):
        ch.run(t - done)
        done = t
        want.append(P.sliced_wasserstein_distance(ch.state.double().cpu().numpy(), post.astype(np.float32).astype(np.float64),
                                                  n_projections=50, seed=0))
    assert np.allclose(W, want, rtol=1e-5, atol=1e-6), (W, want)
    assert np.array_equal(fin, ch.state.double().cpu().numpy())
    assert W[0] > 2 * W[-1]  # PSGLA's delta = 0.3 discretisation bias keeps the floor near 0.28 here


def test_sliced_w2_argument_errors():
    lib = P._lib.lib()
    assert lib.psgla_gmm2d_sw2_workspace_bytes(10, 0) == 0 and lib.psgla_gmm2d_sw2_workspace_bytes(10, 129) == 0
    x = torch.zeros(10, 2, device="cuda")
    th = _theta(2, 0)
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    out = torch.empty(2, 10, device="cuda")
    rc = lib.psgla_gmm2d_sorted_projections(x.data_ptr(), 0, 10, th.ctypes.data_as(C.POINTER(C.c_float)), 2, out.data_ptr(),
                                            ws.data_ptr(), 256, None)
    assert rc == -4 and b"workspace" in lib.psgla_last_error()
