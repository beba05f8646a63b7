"""Host-side mirror of the reference's ``utils_2D.py`` for the hot path: the GMM priors, the closed-form
MMSE denoiser (as a *structured callable* the CUDA kernels can read) and the true-posterior helpers the
W2 metric needs.  Plot helpers of the reference are out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["gaussian_mixt_example", "Theorical_MMSE", "GMMDenoiser", "constantes_conditionnal_prob",
           "sample_gaussian", "sample_posterior", "Wasserstein_distance", "sliced_wasserstein_distance"]


def gaussian_mixt_example(name):
    """The three priors of the reference (utils_2D.py:23-33); returns mu_list, sigma_list, pi_list."""
    table = {
        "symetric_gaussians": ([[5, 5], [-5, -5]], [np.eye(2), np.eye(2)]),
        "cross": ([[0, 0], [0, 0]], [[[2, 0.5], [0.5, 0.15]], [[0.15, 0.5], [0.5, 2.0]]]),
        "disymmetric_gaussians": ([[0, 3], [0, -5]], [np.eye(2), np.eye(2) / 5]),
    }
    if name not in table:
        raise ValueError("unknown prior %r" % (name,))
    mus, sigs = table[name]
    return [np.array(m) for m in mus], sigs, [0.5, 0.5]


class GMMDenoiser:
    """Exact MMSE denoiser of a 2D Gaussian-mixture prior (utils_2D.py:209-233), evaluated on the GPU.

    Callable like the reference closure -- ``D(x, epsilon) -> ndarray(2,) float64`` -- and also accepts a
    batch ``(n, 2)``.  The samplers do not call it: they read ``mu`` / ``Sigma`` / ``pi`` and fuse the
    denoiser into the chain kernel.  Keeps the reference's quirk that ``sqrt(epsilon)`` plays the role of
    the noise variance (utils_2D.py:223-226); unlike the reference it normalises the mixture weights in
    log space, so it stays finite far away from every mode.
    """

    def __init__(self, mu_list, sigma_list, pi_list):
        self.mu = np.asarray([np.asarray(m, dtype=np.float64) for m in mu_list], dtype=np.float64).reshape(-1, 2)
        self.Sigma = np.asarray([np.asarray(s, dtype=np.float64) for s in sigma_list], dtype=np.float64).reshape(-1, 2, 2)
        self.pi = np.asarray(pi_list, dtype=np.float64).reshape(-1)
        r = self.mu.shape[0]
        if not (self.Sigma.shape[0] == r == self.pi.shape[0]):
            raise ValueError("mu_list, sigma_list and pi_list must have the same length")
        if r < 1 or r > _lib.GMM_MAX_COMPONENTS:
            raise ValueError("between 1 and %d mixture components are supported" % _lib.GMM_MAX_COMPONENTS)

    @property
    def n_components(self):
        return self.mu.shape[0]

    def fill(self, prob: "_lib.GmmProblem"):
        prob.n_components = self.n_components
        for i in range(self.n_components):
            for j in range(2):
                prob.mu[i][j] = float(self.mu[i, j])
            for j in range(4):
                prob.Sigma[i][j] = float(self.Sigma[i].reshape(-1)[j])
            prob.pi[i] = float(self.pi[i])

    def __call__(self, x, epsilon):
        torch = _lib.require_cuda()
        x_np = np.asarray(x, dtype=np.float64)
        flat = np.ascontiguousarray(x_np.reshape(-1, 2))
        prob = _lib.GmmProblem()
        self.fill(prob)
        dev = torch.device("cuda", torch.cuda.current_device())
        xd = torch.from_numpy(flat).to(dev)
        out = torch.empty_like(xd)
        _lib.check(_lib.lib().psgla_gmm2d_denoise(C.byref(prob), float(epsilon), 1, _lib.ptr(xd), _lib.ptr(out),
                                                  flat.shape[0], _lib.stream_ptr(dev)), "psgla_gmm2d_denoise")
        return out.cpu().numpy().reshape(x_np.shape)


def Theorical_MMSE(mu_list, sigma_list, pi_list):
    """Drop-in for utils_2D.py:209 -- returns the structured callable above."""
    return GMMDenoiser(mu_list, sigma_list, pi_list)


# ----------------------------------------------------------------------------- true posterior (metric side, host)


def _spd_sqrt(S):
    w, v = np.linalg.eigh(np.asarray(S, dtype=np.float64))
    return (v * np.sqrt(w)) @ v.T


def constantes_conditionnal_prob(A, y, sigma, mu_list, sigma_list, pi_list):
    """Closed-form posterior of x | y for y = A x + n, n ~ N(0, sigma I) with a GMM prior (utils_2D.py:139-162;
    ``sigma`` is the noise *variance* there).  Returns (mu_cond_list, sigma_cond_list, p_list)."""
    A = np.asarray(A, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    AtA, Aty = A.T @ A / sigma, A @ y / sigma
    mus, covs, logp = [], [], []
    for mu, Sig, pi in zip(mu_list, sigma_list, pi_list):
        mu = np.asarray(mu, dtype=np.float64)
        Sig = np.asarray(Sig, dtype=np.float64)
        prec0 = np.linalg.inv(Sig)
        prec = prec0 + AtA
        cov = np.linalg.inv(prec)
        m = cov @ (prec0 @ mu + Aty)
        root = _spd_sqrt(Sig)
        logdet = np.log(np.linalg.det(root @ A.T @ A @ root + sigma * np.eye(2)))
        logp.append(np.log(pi) + 0.5 * (m @ prec @ m - mu @ prec0 @ mu - y @ y / sigma) - 0.5 * logdet)
        mus.append(m)
        covs.append(cov)
    logp = np.asarray(logp)
    p = np.exp(logp - logp.max())
    return mus, covs, p / p.sum()


def sample_gaussian(mu_list, sigma_list, pi_list, N, rng=None):
    """N draws of a Gaussian mixture, component sizes int(pi_i N) as in utils_2D.py:85-101.  ``rng=None`` uses the
    global legacy NumPy stream in the same order as the reference."""
    legacy = rng is None
    parts = []
    for mu, Sig, pi in zip(mu_list, sigma_list, pi_list):
        n = int(pi * N)
        u = np.random.randn(2, n) if legacy else rng.standard_normal((2, n))
        parts.append(np.asarray(mu, dtype=np.float64)[:, None] + _spd_sqrt(Sig) @ u)
    X = np.concatenate(parts, axis=1).T
    return np.random.permutation(X) if legacy else rng.permutation(X)


def sample_posterior(A, y, sigma, N, mu_list, sigma_list, pi_list, rng=None):
    """utils_2D.py:164-169."""
    return sample_gaussian(*constantes_conditionnal_prob(A, y, sigma, mu_list, sigma_list, pi_list), N, rng=rng)


def _uniform_transport_cost(M):
    """min <T, M> over T >= 0 with row sums 1/n1 and column sums 1/n2: the transport LP ``ot.emd2(a=[], b=[], M)`` solves for
    clouds of UNEQUAL size (n1 n2 variables, n1 + n2 equalities; HiGHS through scipy)."""
    from scipy.optimize import linprog
    from scipy.sparse import coo_matrix
    n1, n2 = M.shape
    rows = np.concatenate([np.repeat(np.arange(n1), n2), n1 + np.tile(np.arange(n2), n1)])
    cols = np.concatenate([np.arange(n1 * n2), np.arange(n1 * n2)])
    A = coo_matrix((np.ones(2 * n1 * n2), (rows, cols)), shape=(n1 + n2, n1 * n2)).tocsr()
    b = np.concatenate([np.full(n1, 1.0 / n1), np.full(n2, 1.0 / n2)])
    res = linprog(M.reshape(-1), A_eq=A[:-1], b_eq=b[:-1], bounds=(0, None), method="highs")  # one equality is redundant
    if res.status != 0:
        raise RuntimeError("transport LP did not converge: %s" % res.message)
    return float(res.fun)


def Wasserstein_distance(sample1, sample2, n_sub=1000, rng=None):
    """Exact optimal transport between random ``n_sub``-point subsamples with squared-Euclidean cost, i.e. W2^2
    (utils_2D.py:235-244: ``ot.dist`` default metric + ``ot.emd2`` with uniform weights).  POT is replaced by the assignment
    problem -- the same linear programme for equal-size uniform clouds -- and, for clouds of unequal size (``sample_posterior``
    returns sum_i int(pi_i N) points, e.g. 99 against a 100-point chain), by the uniform-marginal transport LP itself."""
    from scipy.optimize import linear_sum_assignment
    perm = np.random.permutation if rng is None else rng.permutation
    s1 = perm(np.asarray(sample1))[:n_sub]
    s2 = perm(np.asarray(sample2))[:n_sub]
    M = ((s1[:, None, :] - s2[None, :, :]) ** 2).sum(-1)
    if len(s1) != len(s2):
        return _uniform_transport_cost(M)
    rows, cols = linear_sum_assignment(M)
    return float(M[rows, cols].sum() / len(s1))


def sliced_wasserstein_distance(X, Y, n_projections=50, seed=None):
    """Sliced W2 (``ot.sliced.sliced_wasserstein_distance(..., p=2)``, sampling_2D.py:168-170) for equal-size clouds."""
    rng = np.random.default_rng(seed)
    X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
    if X.shape != Y.shape:
        raise ValueError("equal-size clouds required")
    theta = rng.standard_normal((2, n_projections))
    theta /= np.linalg.norm(theta, axis=0, keepdims=True)
    px, py = np.sort(X @ theta, axis=0), np.sort(Y @ theta, axis=0)
    return float(np.sqrt(np.mean((px - py) ** 2)))
