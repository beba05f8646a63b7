#!/usr/bin/env python
"""Measured per-SM issue rates of the instruction classes the 2D chain kernel mixes (development aid).
    python scripts/pipe_rates.py > gpurun_out/pipe_rates.txt"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import psgla_b200 as P  # noqa: E402

NAMES = ["IMAD.WIDE.U32", "IMAD.HI.U32", "IMAD (lo)", "LOP3", "MUFU.EX2", "I2FP.F32.U32", "FFMA", "MUFU.SIN (+FMUL.RZ)",
         "IMAD.WIDE + FFMA 1:1", "MUFU.EX2 + IMAD.WIDE 1:1", "MUFU.EX2 + 4 FFMA", "chain-step mix (6M 9W 19F 10L 2C)"]
lib = P._lib.lib()
scratch = torch.empty(148 * 8 * 256, dtype=torch.float32, device="cuda")
sms = torch.cuda.get_device_properties(0).multi_processor_count
for mode, name in enumerate(NAMES):
    ops = C.c_double()
    best = None
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        P._lib.check(lib.psgla_selftest_pipe_rate(mode, 20000, 8, scratch.data_ptr(), C.byref(ops), None), "pipe_rate")
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        best = t if best is None else min(best, t)
    per_s = ops.value / best
    clk = torch.cuda.clock_rate() * 1e6 if hasattr(torch.cuda, "clock_rate") else 1.965e9
    print("%-36s %8.2f Gop/s/SM  = %6.2f thread-instr/clk/SM at %.2f GHz%s"
          % (name, per_s / sms / 1e9, per_s / sms / clk, clk / 1e9,
             "  -> %.1f SM-sub-partition cycles per warp-step" % (clk * 4 * 32 / (per_s / sms)) if mode == 11 else ""), flush=True)
