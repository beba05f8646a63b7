"""-m gpu: rng="torch_cuda" -- the reference's per-iteration ``torch.randn(shape, generator=Generator("cuda").manual_seed(seed))``
(restoration_algorithms.py:86-87,104 / :212-213,232) regenerated inside the fused kernels from the seed alone.

The checker here is torch itself, live on the same GPU: the stream must agree BIT FOR BIT (integer Philox, then curand's
Box-Muller with the same libdevice logf / sqrtf / __sincosf), so the tolerance is 0."""
import ctypes as C

import pytest
import torch

import psgla_b200 as P
from oracle import image_oracle as io_

pytestmark = pytest.mark.gpu


def _policy(numel):
    props = torch.cuda.get_device_properties(0)
    t, s = C.c_uint32(), C.c_uint64()
    P._lib.check(P._lib.lib().psgla_torch_cuda_randn_policy(numel, props.multi_processor_count,
                                                            props.max_threads_per_multi_processor, C.byref(t), C.byref(s)),
                 "policy")
    return t.value, s.value


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (1, 3, 256, 256), (1, 3, 321, 481), (5,), (1, 3, 7, 9),
                                   (8, 3, 256, 256), (32, 3, 256, 256), (3, 3, 321, 481)])
@pytest.mark.parametrize("seed", [0, 1234567891011])
def test_stream_equals_torch_randn(shape, seed):
    lib = P._lib.lib()
    numel = 1
    for d in shape:
        numel *= d
    threads, step = _policy(numel)
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = torch.empty(shape, device="cuda")
    for call in range(3):  # successive calls advance the generator's offset by `step`
        want = torch.randn(shape, generator=g, device="cuda")
        P._lib.check(lib.psgla_img_noise_torch_cuda(numel, seed, call * step, threads, out.data_ptr(), None), "noise")
        torch.cuda.synchronize()
        assert torch.equal(out, want), (shape, seed, call, (out - want).abs().max().item())
    # the policy's offset step is what torch's generator actually advanced by
    assert g.get_offset() == 3 * step


@pytest.fixture(scope="module")
def den():
    return P.DnCNN(pretrained=io_.make_dncnn_weights(seed=0, n_power_iter=5, spatial=16))


@pytest.mark.parametrize("problem,H,W,B", [("inpainting", 64, 64, None), ("inpainting", 40, 50, 3), ("inpainting", 33, 31, None),
                                           ("deblurring", 64, 64, None), ("deblurring", 48, 36, 2), ("deblurring", 33, 31, None)])
def test_samplers_with_in_kernel_torch_stream_equal_replayed_torch_randn(den, problem, H, W, B):
    """psgla / pnpula with rng="torch_cuda" (noise generated in the pre kernel) and with rng="torch" (torch.randn called per
    iteration like the reference, tensor replayed) must give identical samples and moments."""
    torch.manual_seed(1)
    im = torch.rand(1, 3, H, W, device="cuda")
    if problem == "inpainting":
        dg, init, _, _ = P.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    else:
        dg, init, _ = P.make_deblurring(im, l=4, sigma=1.0, seed_ip=0)
    prm = io_.resolve_params("psgla")
    kw = dict(alpha=torch.tensor(1.0, device="cuda"), lambd=torch.tensor(prm["lambd"], device="cuda"), sig_float=prm["s"],
              delta=prm["delta"], n_iter=7, n_inter=2, n_inter_mmse=2, seed=3, n_chains=B)
    a = P.psgla(init, dg, den, rng="torch", **kw)
    b = P.psgla(init, dg, den, rng="torch_cuda", **kw)
    torch.cuda.synchronize()
    for la, lb in zip(a, b):
        assert len(la) == len(lb)
        for ta, tb in zip(la, lb):
            assert torch.equal(ta, tb)
    pg = P.PriorGrad(den, 1.0, 5 / 255, (5 / 255) ** 2)
    kw = dict(delta=torch.tensor(1e-5, device="cuda"), lambd=torch.tensor(2e-5, device="cuda"), n_iter=5, n_inter=1,
              n_inter_mmse=2, seed=11, n_chains=B)
    a = P.pnpula(init, dg, pg, rng="torch", **kw)
    b = P.pnpula(init, dg, pg, rng="torch_cuda", **kw)
    torch.cuda.synchronize()
    for la, lb in zip(a, b):
        for ta, tb in zip(la, lb):
            assert torch.equal(ta, tb)


def test_torch_stream_argument_errors():
    lib = P._lib.lib()
    out = torch.empty(16, device="cuda")
    assert lib.psgla_img_noise_torch_cuda(16, 0, 2, 256, out.data_ptr(), None) == -1  # offset not a multiple of 4
    assert lib.psgla_img_noise_torch_cuda(16, 0, 0, 100, out.data_ptr(), None) == -1  # threads not a multiple of 256
    assert b"psgla_img_noise_torch_cuda" in lib.psgla_last_error()


@pytest.mark.parametrize("problem", ["inpainting", "deblurring"])
def test_seed_only_drop_in_against_the_reference_algorithm_on_cuda(problem):
    """The drop-in property the torch stream buys: the reference algorithm (oracle restatement, bit-identical to the
    unmodified reference on the CPU, run here on cuda with ITS OWN torch.Generator(seed)) and psgla(..., seed=seed,
    rng="torch_cuda") consume the same noise without any tensor crossing between them; what remains is the bf16 denoiser
    against the fp32 one.  Tolerance 1e-3 abs on iterates in [0, 1] over 6 iterations (observed ~1e-4), bookkeeping exact."""
    sd = io_.make_dncnn_weights(seed=0, n_power_iter=5, spatial=16)
    den = P.DnCNN(pretrained=sd)
    net = io_.DnCNN().cuda()
    net.load_state_dict(sd)
    torch.manual_seed(2)
    im = torch.rand(1, 3, 64, 64, device="cuda")
    if problem == "inpainting":
        dg, init, _, _ = P.make_inpainting(im, prop=0.5, sigma=1.0, seed_ip=0)
    else:
        dg, init, _ = P.make_deblurring(im, l=4, sigma=1.0, seed_ip=0)
    prm = io_.resolve_params("psgla")
    kw = dict(alpha=torch.tensor(1.0, device="cuda"), lambd=torch.tensor(prm["lambd"], device="cuda"), sig_float=prm["s"],
              delta=prm["delta"], n_iter=6, n_inter=2, n_inter_mmse=2, seed=5)
    Xr, Mr, M2r = io_.psgla(init, dg, net.eval(), device="cuda", **kw)
    Xg, Mg, M2g = P.psgla(init, dg, den, rng="torch_cuda", **kw)
    torch.cuda.synchronize()
    assert len(Xr) == len(Xg) == 3 and len(Mr) == len(Mg) == 2 and len(M2r) == len(M2g) == 2
    for a, b in zip(Xr + Mr + M2r, Xg + Mg + M2g):
        assert (a - b).abs().max().item() < 1e-3
    # and the noise really is the same: with a different seed the iterates differ at the noise scale
    Xo, _, _ = P.psgla(init, dg, den, rng="torch_cuda", **dict(kw, seed=6))
    assert (Xo[0] - Xg[0]).abs().max().item() > 1e-2
