"""The numbers the reference's driver script resolves before it calls a sampler (sampling_images.py:100-123,147-198),
without its ``sys.argv`` sniffing: a flag the user did not give is passed as ``None`` (the script tests
``'--s' in sys.argv`` etc., so "flag present with its default value" and "flag absent" differ).

``sampler_params("psgla", den="DnCNN")`` -> ``dict(s, lambd, delta, N, n_inter, n_inter_mmse, sigma1, sigma2, alpha, ...)``
ready to be splatted into ``psgla`` / ``pnpula`` (see ``as_psgla_kwargs`` / ``as_pnpula_kwargs``).  Quirks kept on purpose:

* ``n_inter = int(N / 1000)`` uses the *parsed* ``--N`` (default 10 000), before the per-algorithm override, so default
  PnP-ULA (N = 100 000) still thins every 10 iterations (:105 vs :159-162);
* PnP-ULA divides ``s`` by 255 a second time: ``s1 = s / 255`` where the DnCNN default ``s`` is already ``2 / 255``
  (:149-153);
* PSGLA with a non-DnCNN denoiser defaults to ``s = 5/255, lambd = 1``, i.e. a data-term gain (delta/lambd)/sigma^2 = 25.
"""
from __future__ import annotations

__all__ = ["sampler_params", "as_psgla_kwargs", "as_pnpula_kwargs"]

_DEFAULT_N, _DEFAULT_S, _DEFAULT_LAMBD = 10000, 5.0, 1.0  # argparse defaults, sampling_images.py:20-32


def sampler_params(alg, den="DnCNN", sigma=1.0, alpha=1.0, N=None, s=None, lambd=None):
    """``N`` / ``s`` / ``lambd`` = None mean "flag absent from the command line"."""
    n_parsed = _DEFAULT_N if N is None else int(N)
    sigma1 = sigma / 255.0
    sigma2 = sigma1 ** 2
    out = dict(alg=alg, den=den, sigma1=sigma1, sigma2=sigma2, alpha=alpha, n_inter=int(n_parsed / 1000))  # :105
    out["n_inter_mmse"] = out["n_inter"]  # :106
    if alg == "pnp_ula":
        s_ = 2.0 / 255.0 if (s is None and den == "DnCNN") else (_DEFAULT_S if s is None else s)  # :149-152
        s1 = s_ / 255.0  # :153
        s2 = s1 ** 2
        n_run = 100000 if (N is None and den == "DnCNN") else n_parsed  # :159-162
        lam = 0.5 / (2 / sigma2 + alpha / s2)  # :164
        delta = 1 / 3 / (1 / sigma2 + 1 / lam + alpha / s2)  # :167
        out.update(s=s_, s1=s1, s2=s2, N=n_run, lambd=lam, delta=delta, c_min=-1, c_max=2)  # the call omits c_min/c_max (:358)
    elif alg == "psgla":
        n_run = n_parsed
        if den == "DnCNN":
            s_ = 2.0 / 255.0 if s is None else s / 255.0  # :172-175
            lam = 5.0 if lambd is None else lambd  # :176-179
        elif den == "TV":
            s_ = 10.0 / 255.0 if s is None else s / 255.0  # :181-184
            lam = 10.0 if lambd is None else lambd
            n_run = 1000 if N is None else n_parsed
        else:
            s_ = (_DEFAULT_S if s is None else s) / 255.0  # :194
            lam = _DEFAULT_LAMBD if lambd is None else lambd
        out.update(s=s_, N=n_run, lambd=lam, delta=s_ ** 2)  # :198
    else:
        raise ValueError("alg must be 'psgla' or 'pnp_ula'")
    return out


def as_psgla_kwargs(p, seed=0):
    """Keyword arguments of ``psgla`` as the script passes them (sampling_images.py:351)."""
    return dict(alpha=p["alpha"], lambd=p["lambd"], sig_float=p["s"], delta=p["delta"], seed=seed, n_iter=p["N"],
                n_inter=p["n_inter"], n_inter_mmse=p["n_inter_mmse"])


def as_pnpula_kwargs(p, seed=0):
    """Keyword arguments of ``pnpula`` as the script passes them (sampling_images.py:358); build ``prior_grad`` with
    ``PriorGrad(denoiser, p["alpha"], p["s1"], p["s2"])``."""
    return dict(delta=p["delta"], lambd=p["lambd"], seed=seed, n_iter=p["N"], n_inter=p["n_inter"], n_inter_mmse=p["n_inter_mmse"])
