"""TEST INFRASTRUCTURE ONLY -- builds and binds oracle/gmm2d_oracle.c (the plain-C float64 restatement of
sampling_2D.py:21-72 and utils_2D.py:209-233).  Same audience as the other oracle modules: tests/, smoke(), and the
cpu_baseline / --impl reference legs of bench.py; the product never imports it.

    build()   gcc -O2 -shared -fPIC oracle/gmm2d_oracle.c -o oracle/_build/libgmm2d_oracle.so -lm   (git-ignored output)
    denoise / pnp_ula / snopnp_ula   NumPy-in, NumPy-out wrappers with the signatures of oracle/gmm2d_oracle.py
    run_chain                        compiled single-chain baseline (its own noise stream), returns the final state
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gmm2d_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libgmm2d_oracle.so")
_lib = None
_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        raise RuntimeError("no C compiler for oracle/gmm2d_oracle.c")
    os.makedirs(OUT_DIR, exist_ok=True)
    res = subprocess.run([gcc, "-O2", "-shared", "-fPIC", "-std=c99", SRC, "-o", LIB, "-lm"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed for %s:\n%s" % (SRC, res.stderr))
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        h = C.CDLL(build())
        h.gmm2d_denoise.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, _dp]
        h.gmm2d_pnp_ula.argtypes = [C.c_long, _dp, _dp, C.c_double, _dp, C.c_double, C.c_int, _dp, _dp, _dp, C.c_double,
                                    C.c_double, _dp, _dp]
        h.gmm2d_snopnp_ula.argtypes = [C.c_long, _dp, _dp, C.c_double, _dp, C.c_double, C.c_int, _dp, _dp, _dp, C.c_double, _dp,
                                       _dp]
        h.gmm2d_run_chain.argtypes = [C.c_int, C.c_long, _dp, _dp, C.c_double, _dp, C.c_double, C.c_int, _dp, _dp, _dp,
                                      C.c_double, C.c_double, C.c_ulonglong, _dp]
        for f in (h.gmm2d_denoise, h.gmm2d_pnp_ula, h.gmm2d_snopnp_ula, h.gmm2d_run_chain):
            f.restype = None
        _lib = h
    return _lib


def _arr(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a if shape is None else a.reshape(shape)


def _p(a):
    return a.ctypes.data_as(_dp)


def _prior(mu_list, sigma_list, pi_list):
    r = len(mu_list)
    return r, _arr(mu_list, (r, 2)), _arr([np.asarray(s, dtype=np.float64) for s in sigma_list], (r, 4)), _arr(pi_list, (r,))


def denoise(mu_list, sigma_list, pi_list, x, epsilon):
    r, mu, Sig, pi = _prior(mu_list, sigma_list, pi_list)
    x = _arr(x, (2,))
    out = np.empty(2)
    lib().gmm2d_denoise(r, _p(mu), _p(Sig), _p(pi), _p(x), float(epsilon), _p(out))
    return out


def pnp_ula(N, x_0, y, delta, A, sigma, prior, epsilon, alpha, noise):
    r, mu, Sig, pi = _prior(*prior)
    x0, yy, AA, nz = _arr(x_0, (2,)), _arr(y, (2,)), _arr(A, (4,)), _arr(noise, (max(int(N) - 1, 0), 2))
    traj = np.empty((int(N), 2))
    lib().gmm2d_pnp_ula(int(N), _p(x0), _p(yy), float(delta), _p(AA), float(sigma), r, _p(mu), _p(Sig), _p(pi), float(epsilon),
                        float(alpha), _p(nz), _p(traj))
    return traj


def snopnp_ula(N, x_0, y, delta, A, sigma, prior, alpha, noise):
    r, mu, Sig, pi = _prior(*prior)
    x0, yy, AA, nz = _arr(x_0, (2,)), _arr(y, (2,)), _arr(A, (4,)), _arr(noise, (max(int(N) - 1, 0), 2))
    traj = np.empty((int(N), 2))
    lib().gmm2d_snopnp_ula(int(N), _p(x0), _p(yy), float(delta), _p(AA), float(sigma), r, _p(mu), _p(Sig), _p(pi), float(alpha),
                           _p(nz), _p(traj))
    return traj


def run_chain(alg, n_steps, x_0, y, delta, A, sigma, prior, epsilon, alpha, seed=0):
    r, mu, Sig, pi = _prior(*prior)
    x0, yy, AA = _arr(x_0, (2,)), _arr(y, (2,)), _arr(A, (4,))
    out = np.empty(2)
    lib().gmm2d_run_chain(0 if alg in ("psgla", "snopnp_ula") else 1, int(n_steps), _p(x0), _p(yy), float(delta), _p(AA),
                          float(sigma), r, _p(mu), _p(Sig), _p(pi), float(epsilon), float(alpha), int(seed), _p(out))
    return out
