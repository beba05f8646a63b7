"""GPU parity tests of the DRUNet path: the general tcgen05 conv layers against torch.nn.functional on bf16-rounded
operands (fp32 accumulate on both sides), then the whole denoiser and the samplers against the fp32 oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import psgla_b200 as P

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _run_layer(mode, x, w_taps, res1=None, res2=None, relu=False):
    """x: bf16 NCHW; w_taps: bf16 [taps][Cout][Cin]; returns fp32 NCHW."""
    lib = P._lib.lib()
    B, Cin, H, W = x.shape
    Cout = w_taps.shape[1]
    Ho, Wo = {0: (H, W), 1: (H // 2, W // 2), 2: (2 * H, 2 * W)}[mode]
    xin = _nhwc(x)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    r1 = _nhwc(res1) if res1 is not None else None
    r2 = _nhwc(res2) if res2 is not None else None
    P._lib.check(lib.psgla_convg_layer(mode, B, H, W, Cin, Cout, w_taps.data_ptr(), xin.data_ptr(),
                                       r1.data_ptr() if r1 is not None else None, r2.data_ptr() if r2 is not None else None,
                                       out.data_ptr(), int(relu), None), "psgla_convg_layer")
    torch.cuda.synchronize()
    return out.float().permute(0, 3, 1, 2)


def _tol(ref):
    return 2 ** -8 * max(1.0, ref.abs().max().item()) * 1.01  # one bf16 rounding of the output


@pytest.mark.parametrize("B,H,W,Cin,Cout,res,relu", [
    (2, 20, 24, 128, 128, 0, True),    # PX = 24, 5 rows per tile (120 of 128 tile pixels used)
    (1, 9, 240, 128, 128, 1, False),   # two 128-pixel strips, ragged second strip
    (1, 64, 64, 256, 256, 2, False),   # N tile 256, two residual inputs
    (2, 8, 8, 512, 512, 1, True),      # deepest scale, 8 k-blocks
    (1, 7, 130, 64, 192, 0, False),    # N tile 64 x 3, odd extents
    (1, 1, 1, 128, 128, 0, False),
])
def test_conv3x3_general(B, H, W, Cin, Cout, res, relu):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = _bf(torch.randn(B, Cin, H, W, device="cuda", generator=g))
    w = _bf(torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / np.sqrt(9 * Cin))
    rs = [_bf(torch.randn(B, Cout, H, W, device="cuda", generator=g)) for _ in range(res)]
    ref = F.conv2d(x.float(), w.float(), padding=1)
    for r in rs:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    w_taps = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    got = _run_layer(0, x, w_taps, *(rs + [None, None])[:2], relu=relu)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= _tol(ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 40, 64, 128), (1, 8, 8, 256, 512), (1, 2, 260, 128, 256)])
def test_strided_down_conv(B, H, W, Cin, Cout):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = _bf(torch.randn(B, Cin, H, W, device="cuda", generator=g))
    w = _bf(torch.randn(Cout, Cin, 2, 2, device="cuda", generator=g) / np.sqrt(4 * Cin))
    ref = F.conv2d(x.float(), w.float(), stride=2)
    got = _run_layer(1, x, w.permute(2, 3, 0, 1).reshape(4, Cout, Cin).contiguous())
    assert (got - ref).abs().max().item() <= _tol(ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout,res", [(2, 8, 20, 128, 64, 0), (1, 4, 4, 512, 256, 0), (1, 3, 130, 256, 128, 0)])
def test_transposed_up_conv(B, H, W, Cin, Cout, res):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = _bf(torch.randn(B, Cin, H, W, device="cuda", generator=g))
    w = _bf(torch.randn(Cin, Cout, 2, 2, device="cuda", generator=g) / np.sqrt(Cin))  # ConvTranspose2d layout
    ref = F.conv_transpose2d(x.float(), w.float(), stride=2)
    got = _run_layer(2, x, w.permute(2, 3, 1, 0).reshape(4, Cout, Cin).contiguous())
    assert (got - ref).abs().max().item() <= _tol(ref)


def test_convg_argument_errors():
    lib = P._lib.lib()
    t = torch.zeros(16, device="cuda", dtype=torch.bfloat16)
    assert lib.psgla_convg_layer(0, 1, 4, 4, 48, 64, t.data_ptr(), t.data_ptr(), None, None, t.data_ptr(), 0, None) == -1
    assert lib.psgla_convg_layer(1, 1, 5, 4, 64, 128, t.data_ptr(), t.data_ptr(), None, None, t.data_ptr(), 0, None) == -1
    assert lib.psgla_convg_layer(7, 1, 4, 4, 64, 64, t.data_ptr(), t.data_ptr(), None, None, t.data_ptr(), 0, None) == -1
