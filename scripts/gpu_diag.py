"""Development aid: runs each GPU check in its own process (a trapping kernel poisons its CUDA context) and prints
compact results.  Usage on the GPU box:  python scripts/gpu_diag.py [check ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check_umma():
    import torch
    import psgla_b200 as P
    lib = P._lib.lib()
    torch.manual_seed(0)
    a = (torch.randn(136, 64, device="cuda")).to(torch.bfloat16).contiguous()
    b = (torch.randn(64, 64, device="cuda")).to(torch.bfloat16).contiguous()
    for mode in (0, 1):
        for shift in (0, 1, 2, 3, 8):
            d = torch.zeros(128, 64, device="cuda")
            rc = lib.psgla_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), shift, mode, None)
            torch.cuda.synchronize()
            ref = a[shift:shift + 128].float() @ b.float().t()
            print("umma mode=%d shift=%d rc=%d maxerr=%.4g (ref max %.3g)" % (mode, shift, rc, (d - ref).abs().max().item(),
                                                                            ref.abs().max().item()), flush=True)


def _conv_case(B, H, W, layer, depth=20, seed=0):
    import torch
    import torch.nn.functional as F
    import psgla_b200 as P
    lib = P._lib.lib()
    sd = P.random_dncnn_state_dict(seed, depth, scale=3.0)
    den = P.DnCNN(depth=depth, pretrained=sd)
    names = ["in_conv"] + ["conv_list.%d" % i for i in range(depth - 2)] + ["out_conv"]
    w = sd[names[layer] + ".weight"].cuda()
    bias = sd[names[layer] + ".bias"].cuda()
    cin = w.shape[1]
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    xb = x.to(torch.bfloat16)
    cpad = 16 if layer == 0 else 64
    xin = torch.zeros(B, H, W, cpad, device="cuda", dtype=torch.bfloat16)
    xin[..., :cin] = xb.permute(0, 2, 3, 1)
    xin = xin.contiguous()
    shape = P._lib.ImgShape(B, 3, H, W)
    ref = F.conv2d(xb.float(), w.to(torch.bfloat16).float(), bias, padding=1)
    if layer == depth - 1:
        out = torch.full((B, 3, H, W), float("nan"), device="cuda")
        rc = lib.psgla_conv3x3_layer(den.packed.data_ptr(), depth, layer, shape, xin.data_ptr(), out.data_ptr(), 0, None)
        torch.cuda.synchronize()
        got = out
    else:
        out = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        rc = lib.psgla_conv3x3_layer(den.packed.data_ptr(), depth, layer, shape, xin.data_ptr(), out.data_ptr(), 1, None)
        torch.cuda.synchronize()
        got = out.float().permute(0, 3, 1, 2)
        ref = ref.relu()
    err = (got - ref).abs()
    print("conv layer=%d B=%d H=%d W=%d rc=%d maxerr=%.4g meanerr=%.4g refmax=%.3g nan=%d" % (
        layer, B, H, W, rc, err.nan_to_num(1e9).max().item(), err.nan_to_num(0).mean().item(), ref.abs().max().item(),
        int(torch.isnan(got).sum().item())), flush=True)
    if err.nan_to_num(1e9).max().item() > 0.1:
        bad = (err.nan_to_num(1e9) > 0.1).nonzero()
        print("   first bad idx:", bad[:4].tolist(), " n_bad=", bad.shape[0], flush=True)


def check_conv_mid():
    _conv_case(1, 8, 128, 1)
    _conv_case(2, 40, 256, 5)
    _conv_case(1, 33, 200, 3)


def check_conv_first():
    _conv_case(1, 8, 128, 0)
    _conv_case(2, 37, 150, 0)


def check_conv_last():
    _conv_case(1, 8, 128, 19)
    _conv_case(2, 37, 150, 19)


def check_dncnn():
    import torch
    import psgla_b200 as P
    from oracle import image_oracle as io_
    sd = io_.make_dncnn_weights(seed=0, n_power_iter=10, spatial=16)
    den = P.DnCNN(pretrained=sd)
    net = io_.DnCNN().cuda()
    net.load_state_dict(sd)
    x = torch.rand(2, 3, 64, 96, device="cuda")
    with torch.no_grad():
        ref = net(x)
    got = den.forward(x)
    torch.cuda.synchronize()
    r_ref, r_got = ref - x, got - x
    print("dncnn full: max|D-Dref|=%.4g  |R|max=%.4g  rel residual err=%.4g" % (
        (got - ref).abs().max().item(), r_ref.abs().max().item(), ((r_got - r_ref).norm() / r_ref.norm()).item()), flush=True)


def check_gmm():
    import numpy as np
    import psgla_b200 as P
    from oracle import gmm2d_oracle as o
    for name in o.PRIOR_NAMES:
        mu, Sig, pi = o.gaussian_mixt_example(name)
        D = P.Theorical_MMSE(mu, Sig, pi)
        Do = o.theorical_mmse(mu, Sig, pi)
        y = np.array([0.0, -2.0])
        rng = np.random.default_rng(0)
        noise = rng.standard_normal((299, 2))
        Xo = o.snopnp_ula(300, y, y, 0.3, np.eye(2), 1, Do, 2 / 3, noise=noise)
        X64 = P.SnoPnP_ULA(300, y, y, 0.3, np.eye(2), 1, D, 2 / 3, noise=noise)
        X32 = P.SnoPnP_ULA(300, y, y, 0.3, np.eye(2), 1, D, 2 / 3, noise=noise, dtype="float32")
        print("gmm psgla %-22s fp64 err %.3g  fp32 err %.3g" % (name, np.abs(X64 - Xo).max(), np.abs(X32 - Xo).max()), flush=True)
        Xo = o.pnp_ula(300, y, y, 0.1, np.eye(2), 1, Do, 0.5, 1.5, noise=noise)
        X64 = P.PnP_ULA(300, y, y, 0.1, np.eye(2), 1, D, 0.5, 1.5, noise=noise)
        X32 = P.PnP_ULA(300, y, y, 0.1, np.eye(2), 1, D, 0.5, 1.5, noise=noise, dtype="float32")
        print("gmm ula   %-22s fp64 err %.3g  fp32 err %.3g" % (name, np.abs(X64 - Xo).max(), np.abs(X32 - Xo).max()), flush=True)


def check_gmm_speed():
    import time
    import numpy as np
    import torch
    import psgla_b200 as P
    mu, Sig, pi = P.gaussian_mixt_example("symetric_gaussians")
    D = P.Theorical_MMSE(mu, Sig, pi)
    for alg, args in (("psgla", dict(delta=0.3, alpha=2 / 3)), ("pnp_ula", dict(delta=0.1, alpha=1.5, epsilon=0.5))):
        for nc in (1 << 20, 1000000, 148 * 2048 * 4):
            ch = P.GMMChains(alg, np.array([0.0, -2.0]), A=np.eye(2), sigma=1, denoiser=D, n_chains=nc, **args)
            ch.run(100)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ch.run(2000)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("gmm speed %s chains=%d: %.3f ms for 2000 steps -> %.3e chain-steps/s" % (alg, nc, ms, nc * 2000 / ms * 1e3), flush=True)


CHECKS = {k[6:]: v for k, v in list(globals().items()) if k.startswith("check_")}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        CHECKS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CHECKS)
    for n in names:
        print("=== %s" % n, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], timeout=300, capture_output=True, text=True,
                               env=dict(os.environ))
            print(r.stdout[-4000:], end="")
            if r.returncode != 0:
                print("--- rc=%d stderr tail:\n%s" % (r.returncode, r.stderr[-2500:]))
        except subprocess.TimeoutExpired:
            print("--- TIMEOUT")
