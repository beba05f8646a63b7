"""Drop-in for the samplers of the reference's ``sampling_2D.py``: ``PnP_ULA`` (:21-45) and ``SnoPnP_ULA`` (:48-72,
the paper's PSGLA), running as one persistent CUDA kernel (csrc/gmm2d.cu) through the C ABI.

Same positional arguments and return values as the reference.  Extra keyword-only arguments:
  noise            (N-1, 2) or (N-1, n_chains, 2) standard normals to replay instead of drawing new ones.
  rng              "numpy" (default for a single chain): draw the N-1 Gaussian pairs from the *global* NumPy
                   stream, exactly the draws the reference makes, and replay them on the GPU -- so
                   ``np.random.seed(k)`` reproduces the reference trajectory;  "philox" (default when
                   ``n_chains`` is given): in-kernel Philox4x32-10, nothing crosses PCIe.
  n_chains         run that many independent chains from x_0 (or from a (n_chains, 2) x_0).
  return_trajectory  False -> only the final states, shape (n_chains, 2); True -> reference layout (N, 2) or,
                   with chains, (N // thin..., n_chains, 2).
  compute_metric_each_step with n_chains: returns (final states, [sliced W2 of the population vs
                   Sample_posterior[:n_chains] after steps 1, 101, 201, ...]), evaluated on the device.
  seed, chain_id0, philox_offset   Philox key, global id of the first chain, global index of the first step.
  dtype            "float64" (default for one chain: matches the float64 reference to ~1e-12 under replay) or
                   "float32" (throughput path; default when n_chains is given).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .utils_2D import GMMDenoiser, Wasserstein_distance

__all__ = ["PnP_ULA", "SnoPnP_ULA", "GMMChains", "run_chains"]


def _problem(alg, y, delta, A, sigma, denoiser, alpha, epsilon):
    if not isinstance(denoiser, GMMDenoiser):
        raise TypeError("MMSE_denoiser must be the structured callable returned by Theorical_MMSE (a GMMDenoiser); "
                        "an opaque Python callable cannot be fused into the CUDA chain kernel and there is no CPU fallback")
    prob = _lib.GmmProblem()
    prob.alg = alg
    prob.delta, prob.alpha, prob.epsilon, prob.sigma = float(delta), float(alpha), float(epsilon), float(sigma)
    A = np.asarray(A, dtype=np.float64).reshape(2, 2)
    y = np.asarray(y, dtype=np.float64).reshape(2)
    for j in range(4):
        prob.A[j] = float(A.reshape(-1)[j])
    prob.y[0], prob.y[1] = float(y[0]), float(y[1])
    denoiser.fill(prob)
    return prob


class GMMChains:
    """Device-resident population of chains for one (prior, observation, algorithm) cell.

    ``state`` is a CUDA tensor (n_chains, 2); ``run(n_steps)`` advances every chain in place with one kernel
    launch.  ``chain_id0`` is the global id of ``state[0]`` so that sharding chains over GPUs does not change
    any chain's noise stream.
    """

    def __init__(self, alg, y, delta, A, sigma, denoiser, alpha, epsilon=1.0, n_chains=1, x0=None, seed=0,
                 chain_id0=0, dtype="float32", device=None):
        if not isinstance(denoiser, GMMDenoiser):
            _problem(0, y, delta, A, sigma, denoiser, alpha, epsilon)  # raises the TypeError
        torch = _lib.require_cuda()
        self.torch = torch
        self.alg = {"psgla": _lib.ALG_PSGLA, "snopnp_ula": _lib.ALG_PSGLA, "pnp_ula": _lib.ALG_PNPULA}[alg.lower()] \
            if isinstance(alg, str) else int(alg)
        self.problem = _problem(self.alg, y, delta, A, sigma, denoiser, alpha, epsilon)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.tdtype = {"float32": torch.float32, "float64": torch.float64}[str(dtype).replace("torch.", "")]
        self.precision = 0 if self.tdtype == torch.float32 else 1
        self.n_chains = int(n_chains)
        if isinstance(x0, torch.Tensor):  # (n_chains, 2) host (pinned: asynchronous copy) or device tensor
            if tuple(x0.shape) != (self.n_chains, 2):
                raise ValueError("a tensor x0 must have shape (n_chains, 2)")
            self.state = x0.to(self.device, self.tdtype, non_blocking=True).contiguous()
            if self.state.data_ptr() == x0.data_ptr():
                self.state = self.state.clone()
        else:
            x0 = np.asarray(y if x0 is None else x0, dtype=np.float64)
            x0 = np.broadcast_to(x0.reshape(-1, 2), (self.n_chains, 2))
            self.state = torch.from_numpy(np.array(x0, dtype=np.float64, order="C")).to(self.device, self.tdtype).contiguous()
        self.seed, self.chain_id0, self.step = int(seed), int(chain_id0), 0

    def run(self, n_steps, noise=None, thin=0):
        """Advance by n_steps.  noise: CUDA/NumPy (n_steps, n_chains, 2) to replay.  thin>0: also return the states
        after every thin-th step as a CUDA tensor (n_steps // thin, n_chains, 2)."""
        torch = self.torch
        n_steps = int(n_steps)
        traj = None
        if thin:
            traj = torch.empty((n_steps // thin, self.n_chains, 2), dtype=self.tdtype, device=self.device)
        nz = None
        if noise is not None:
            nz = torch.as_tensor(noise).to(self.device, self.tdtype).reshape(n_steps, self.n_chains, 2).contiguous()
        with torch.cuda.device(self.device):
            rc = _lib.lib().psgla_gmm2d_run(C.byref(self.problem), self.precision, _lib.ptr(self.state), self.n_chains,
                                            self.chain_id0, n_steps, self.step, self.seed, _lib.ptr(nz),
                                            _lib.ptr(traj), int(thin) if thin else 1, _lib.stream_ptr(self.device))
        _lib.check(rc, "psgla_gmm2d_run")
        self.step += n_steps
        return traj

    # ---- population metric on the device (sampling_2D.py:38-39,65-66,168-170)
    def set_reference_sample(self, sample, n_projections=50, seed=0):
        """Fix the posterior sample ``(n_chains, 2)`` (e.g. ``sample_posterior(...)``) the population is compared with:
        draws the unit directions like ``utils_2D.sliced_wasserstein_distance`` and sorts the sample's projections once."""
        torch = self.torch
        rng = np.random.default_rng(seed)
        theta = rng.standard_normal((2, int(n_projections)))
        theta /= np.linalg.norm(theta, axis=0, keepdims=True)
        self._theta = np.ascontiguousarray(theta.T, dtype=np.float32)  # [P][2]
        self._theta_c = self._theta.ctypes.data_as(C.POINTER(C.c_float))
        P = self._theta.shape[0]
        lib = _lib.lib()
        nbytes = lib.psgla_gmm2d_sw2_workspace_bytes(self.n_chains, P)
        if nbytes == 0:
            raise ValueError("n_projections must be in 1..128")
        ref = torch.as_tensor(np.asarray(sample) if not isinstance(sample, torch.Tensor) else sample)
        if tuple(ref.shape) != (self.n_chains, 2):
            raise ValueError("the reference sample must have shape (n_chains, 2) = (%d, 2)" % self.n_chains)
        ref = ref.to(self.device, torch.float32).contiguous()
        self._sw_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._ref_sorted = torch.empty((P, self.n_chains), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.psgla_gmm2d_sorted_projections(_lib.ptr(ref), 0, self.n_chains, self._theta_c, P,
                                                    _lib.ptr(self._ref_sorted), _lib.ptr(self._sw_ws), nbytes,
                                                    _lib.stream_ptr(self.device))
        _lib.check(rc, "psgla_gmm2d_sorted_projections")

    def sliced_w2(self, out=None):
        """Sliced W2 between the current population and the reference sample; returns a 0-dim CUDA float64 tensor (or
        fills ``out``) without synchronising the host."""
        if getattr(self, "_ref_sorted", None) is None:
            raise RuntimeError("call set_reference_sample(sample) first")
        torch = self.torch
        if out is None:
            out = torch.empty((), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().psgla_gmm2d_sliced_w2(_lib.ptr(self.state), self.precision, self.n_chains, self._theta_c,
                                                  self._theta.shape[0], _lib.ptr(self._ref_sorted), _lib.ptr(self._sw_ws),
                                                  self._sw_ws.numel(), out.data_ptr(), _lib.stream_ptr(self.device))
        _lib.check(rc, "psgla_gmm2d_sliced_w2")
        return out

    def run_with_metric(self, n_steps, every=100):
        """Advance by n_steps, evaluating the sliced W2 of the population after global steps 1, every+1, 2*every+1, ...
        (the cadence of sampling_2D.py:38,65: ``i % 100 == 0`` after step i).  Returns a CUDA float64 tensor of the
        values; the host never waits inside the loop."""
        n_steps, every = int(n_steps), int(every)
        marks = [t for t in range(n_steps) if t % every == 0]
        vals = self.torch.empty(len(marks), dtype=self.torch.float64, device=self.device)
        done = 0
        for k, t in enumerate(marks):
            self.run(t + 1 - done)
            done = t + 1
            self.sliced_w2(out=vals[k])
        if done < n_steps:
            self.run(n_steps - done)
        return vals


def run_chains(alg, n_steps, y, delta, A, sigma, denoiser, alpha, epsilon=1.0, n_chains=1, x0=None, seed=0,
               chain_id0=0, dtype="float32", device=None, noise=None, thin=0, philox_offset=0):
    """One-shot helper: returns (final_state CUDA tensor (n_chains, 2), thinned trajectory or None)."""
    chains = GMMChains(alg, y, delta, A, sigma, denoiser, alpha, epsilon, n_chains, x0, seed, chain_id0, dtype, device)
    chains.step = int(philox_offset)
    traj = chains.run(n_steps, noise=noise, thin=thin)
    return chains.state, traj


def _sampler(alg, N, x_0, y, delta, A, sigma, denoiser, alpha, epsilon, Sample_posterior, compute_metric_each_step,
             noise, rng, n_chains, return_trajectory, seed, chain_id0, philox_offset, dtype):
    N = int(N)
    single = n_chains is None
    nc = 1 if single else int(n_chains)
    if rng is None:
        rng = "numpy" if (single and noise is None) else "philox"
    if dtype is None:
        dtype = "float64" if single else "float32"
    chains = GMMChains(alg, y, delta, A, sigma, denoiser, alpha, epsilon, nc, x_0, seed, chain_id0, dtype)
    chains.step = int(philox_offset)
    n_steps = max(N - 1, 0)
    if noise is not None:
        noise = np.asarray(noise, dtype=np.float64).reshape(n_steps, nc, 2)

    def draw(k):  # the reference draws np.random.randn(n) once per step from the global stream (sampling_2D.py:35,62)
        return np.random.randn(k, 2).reshape(k, 1, 2)

    x0_host = chains.state.cpu().numpy().astype(np.float64)
    if compute_metric_each_step and not single:
        # a population instead of one trajectory: sliced W2 between the current population and Sample_posterior[:n_chains]
        # at the reference's cadence, computed where the chains live (no per-evaluation D2H)
        if noise is not None or rng != "philox":
            raise ValueError("the population metric runs with rng='philox'")
        chains.set_reference_sample(np.asarray(Sample_posterior, dtype=np.float64)[:nc])
        vals = chains.run_with_metric(n_steps, every=100)
        return chains.state.cpu().numpy().astype(np.float64), [float(v) for v in vals.cpu().numpy()]
    if compute_metric_each_step:
        # Same interleaving of RNG use as the reference: the metric (which permutes with the global stream) runs
        # after steps 0, 100, 200, ... (sampling_2D.py:38-39,65-66).
        rows, W, done = [x0_host], [], 0
        while done < n_steps:
            seg = 1 if done == 0 else min(100, n_steps - done)
            z = noise[done:done + seg] if noise is not None else (draw(seg) if rng == "numpy" else None)
            rows.append(chains.run(seg, noise=z, thin=1).cpu().numpy().astype(np.float64).reshape(seg, 2))
            done += seg
            if (done - 1) % 100 == 0:
                X = np.concatenate(rows, axis=0)
                W.append(Wasserstein_distance(X, np.asarray(Sample_posterior)[:len(X), :]))
        return np.concatenate(rows, axis=0), W
    z = noise if noise is not None else (draw(n_steps) if (rng == "numpy" and n_steps) else None)
    if rng == "numpy" and not single and noise is None:
        raise ValueError("rng='numpy' replays the reference's single global stream; use rng='philox' with n_chains")
    if not return_trajectory:
        chains.run(n_steps, noise=z)
        return chains.state.cpu().numpy().astype(np.float64)
    traj = chains.run(n_steps, noise=z, thin=1)
    out = np.concatenate([x0_host[None], traj.cpu().numpy().astype(np.float64)], axis=0)
    return out[:, 0, :] if single else out


def PnP_ULA(N, x_0, y, delta, A, sigma, MMSE_denoiser, epsilon, alpha, Sample_posterior=[],
            compute_metric_each_step=False, *, noise=None, rng=None, n_chains=None, return_trajectory=True, seed=0,
            chain_id0=0, philox_offset=0, dtype=None):
    """PnP-ULA, sampling_2D.py:21-45: x+ = x + delta score(x) + alpha delta/eps (D(x, eps) - x) + sqrt(2 delta) z.
    Returns the trajectory ``(N, 2)`` float64 including x_0 (and the W2^2 list with compute_metric_each_step)."""
    return _sampler("pnp_ula", N, x_0, y, delta, A, sigma, MMSE_denoiser, alpha, epsilon, Sample_posterior,
                    compute_metric_each_step, noise, rng, n_chains, return_trajectory, seed, chain_id0, philox_offset, dtype)


def SnoPnP_ULA(N, x_0, y, delta, A, sigma, MMSE_denoiser, alpha, Sample_posterior=[], compute_metric_each_step=False,
               *, noise=None, rng=None, n_chains=None, return_trajectory=True, seed=0, chain_id0=0, philox_offset=0,
               dtype=None):
    """PSGLA, sampling_2D.py:48-72: x+ = D(x + (delta/alpha) score(x) + sqrt(2 delta) z, delta)."""
    return _sampler("psgla", N, x_0, y, delta, A, sigma, MMSE_denoiser, alpha, 1.0, Sample_posterior,
                    compute_metric_each_step, noise, rng, n_chains, return_trajectory, seed, chain_id0, philox_offset, dtype)
