#!/usr/bin/env python
"""Launch-geometry sweep of the 2D chain kernel (development aid; numbers quoted in DESIGN.md come from here).

    python scripts/gmm2d_sweep.py [--chains 1000000,125000] [--steps 10000] [--out gpurun_out/gmm2d_sweep.json]

For every geometry "cpt,nb,block,pack,dynamic" (PSGLA_GMM_GEOM, see csrc/gmm2d.cu launch_run) it times one cell of
BASELINE.json configs[1] (symmetric prior, y = (0,-2), PSGLA and PnP-ULA) with CUDA events and checks that the final states
equal those of the round-1 geometry bit for bit (same Philox stream, same arithmetic).  Also runs the FP32 issue-rate probe.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GEOMS = ["4,1,128,0,0", "4,1,128,1,0", "2,2,128,1,0", "2,1,128,1,0", "2,2,128,0,0", "1,4,128,0,0",
         "1,4,32,0,1", "1,2,32,0,1", "2,2,32,0,1", "2,2,32,1,1", "2,1,32,1,1", "2,1,32,0,1", "4,1,32,0,1", "4,1,32,1,1",
         "1,4,64,0,1", "2,2,64,1,1", "4,1,64,1,1", "2,2,128,1,1", "4,1,128,0,1", "4,1,128,1,1", "1,4,128,0,1"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", default="1000000,125000")
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--geoms", default=",".join(g.replace(",", ":") for g in GEOMS))
    ap.add_argument("--priors", default="symetric_gaussians")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gmm2d_sweep.json"))
    args = ap.parse_args()
    import numpy as np
    import torch
    import psgla_b200 as P
    dev = torch.device("cuda", 0)
    lib = P._lib.lib()
    res = {"steps": args.steps, "rows": []}

    # FP32 issue peak
    scratch = torch.empty(148 * 8 * 256, dtype=torch.float32, device=dev)
    for mode in (0, 1):
        flop = C.c_double()
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            P._lib.check(lib.psgla_selftest_fp32_rate(mode, 200000, 8, scratch.data_ptr(), C.byref(flop), None), "fp32_rate")
            e1.record()
            torch.cuda.synchronize()
            best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        res["fp32_tflops_%s" % ("ffma" if mode == 0 else "ffma2")] = best
        print("fp32 rate mode %d: %.1f TFLOP/s" % (mode, best), flush=True)

    y = np.array([0.0, -2.0])
    cases = [(pr, alg, prm) for pr in args.priors.split(",")
             for alg, prm in (("psgla", dict(delta=0.3, alpha=2 / 3, epsilon=1.0)), ("pnp_ula", dict(delta=0.1, alpha=1.5, epsilon=0.5)))]
    for prior, alg, prm in cases:
        mu, Sig, pi = P.gaussian_mixt_example(prior)
        D = P.Theorical_MMSE(mu, Sig, pi)
        for n in [int(v) for v in args.chains.split(",")]:
            want = None
            for geom in [g.replace(":", ",") for g in args.geoms.split(",")]:
                geom, _, struct = geom.partition("/")  # "geom/s": cap the structure specialisation at s (PSGLA_GMM_STRUCT)
                os.environ["PSGLA_GMM_GEOM"] = "" if geom == "default" else geom
                os.environ.pop("PSGLA_GMM_STRUCT", None)
                if struct:
                    os.environ["PSGLA_GMM_STRUCT"] = struct
                    geom = geom + "/" + struct
                x0 = torch.tensor(y, dtype=torch.float32, device=dev).repeat(n, 1).contiguous()
                ch = P.GMMChains(alg, y, prm["delta"], np.eye(2), 1.0, D, prm["alpha"], prm["epsilon"], n_chains=n, x0=x0, seed=0,
                                 dtype="float32", device=dev)
                times = []
                try:
                    for rep in range(args.reps + 1):
                        ch.state.copy_(x0)
                        ch.step = 0
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        ch.run(args.steps)
                        e1.record()
                        torch.cuda.synchronize()
                        if rep:
                            times.append(e0.elapsed_time(e1))
                except RuntimeError as exc:
                    print(alg, n, geom, "FAILED", exc, flush=True)
                    continue
                fin = ch.state.clone()
                if want is None:
                    want = fin
                same = bool(torch.equal(fin, want))
                ms = min(times)
                row = dict(prior=prior, alg=alg, chains=n, geom=geom, ms=ms, steps_per_s=n * args.steps / (ms * 1e-3), bit_equal=same,
                           launches=int(lib.psgla_gmm2d_last_launches()))
                res["rows"].append(row)
                print("%-12.12s %-8s n=%-8d geom=%-14s %8.3f ms  %.3e steps/s  launches=%d  bit_equal=%s"
                      % (prior, alg, n, geom, ms, row["steps_per_s"], row["launches"], same), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
