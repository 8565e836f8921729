#!/usr/bin/env python
"""tools/mma_rate.py: cycles per tcgen05.mma (M = 128, K = 8, tf32, A operand in TMEM) on an idle SM.
probe: loop-invariant operands (nothing but the MMAs in the loop); gemm: the self-test GEMM's issue loop (descriptor arithmetic per MMA)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tempme_b200 import _lib
L = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (16, 32, 64, 96, 128, 192, 256):
    for groups in (6, 24):
        for rep in range(2):
            _lib.check(L.tm_selftest_mma_rate(N, groups, _lib.ptr(out), st), "tm_selftest_mma_rate")
            torch.cuda.synchronize()
        o = out.tolist()
        print(f"probe N={N:3d} MMAs={groups * 8:3d}: issue {o[0]:6d} cyc ({o[0] / (groups * 8):6.1f} / MMA)  complete {o[1]:6d} cyc ({o[1] / (groups * 8):6.1f} / MMA; floor 128*N/256 = {N / 2:.0f})")
