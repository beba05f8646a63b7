#!/usr/bin/env python
"""tcgen05.cp self-tests: A operand copied shared memory -> tensor memory without registers (single CTA, row-shifted start
addresses; CTA pair).  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgla_b200 as P
lib = P._lib.lib()
torch.manual_seed(0)
a = torch.randn(136, 64, device="cuda").to(torch.bfloat16).contiguous()
b = torch.randn(64, 64, device="cuda").to(torch.bfloat16).contiguous()
for mode in (2, 3):
    for shift in (0, 1, 2, 5, 8):
        d = torch.zeros(128, 64, device="cuda")
        P._lib.check(lib.psgla_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), shift, mode, None), "selftest")
        torch.cuda.synchronize()
        ref = a[shift:shift + 128].float() @ b.float().t()
        print("mode %d shift %d: max err %.3g" % (mode, shift, (d - ref).abs().max().item()), flush=True)
a2 = torch.randn(256, 64, device="cuda").to(torch.bfloat16).contiguous()
for mode in (0, 1, 2):
    d = torch.zeros(256, 64, device="cuda")
    P._lib.check(lib.psgla_selftest_umma2(a2.data_ptr(), b.data_ptr(), d.data_ptr(), mode, None), "selftest2")
    torch.cuda.synchronize()
    ref = a2.float() @ b.float().t()
    print("pair mode %d: max err %.3g" % (mode, (d - ref).abs().max().item()), flush=True)
# rates (cycles per iteration on the pair's leader; one pair, and 74 pairs = the whole GPU)
names = {2: "4 cp (128x256b, 4 KB each per CTA)", 3: "conv row: 12 cp feeding this row + 36 MMAs", 4: "conv row: 12 cp ahead + 36 MMAs",
         5: "conv row: 36 MMAs alone"}
for pairs in (1, 74):
    for mode in (2, 5, 3, 4):
        iters = 2000
        out = torch.zeros(pairs, dtype=torch.int64, device="cuda")
        P._lib.check(lib.psgla_selftest_mma_rate2(mode, 64, iters, pairs, out.data_ptr(), None), "mma_rate2")
        torch.cuda.synchronize()
        c = out.cpu().numpy() / iters
        print("pairs=%2d  %-48s %8.1f cycles per iteration (max over pairs %.1f)" % (pairs, names[mode], c.mean(), c.max()), flush=True)
