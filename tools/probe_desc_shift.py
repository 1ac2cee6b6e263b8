#!/usr/bin/env python
"""Hardware probe (B200): can a tcgen05 SWIZZLE_128B K-major operand be read through a descriptor whose start address is
shifted by whole 128-byte rows inside the 1024-byte swizzle atom, and does the descriptor's 'matrix base offset' field have to
name the shift?  The GEMM kernel's debug flag 64 makes the loaders store every activation row one row lower in the stage and
the MMA issuer start its descriptors one row (128 B) later; flag 128 additionally sets base offset = 1.  A correct product
under one of the two variants means the row-shifted reuse of one staged tile by several conv taps is possible."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyrecurrentunet_b200 import _lib as L

fn = L.lib.tru_debug_pw
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 7 + [C.c_int] * 4 + [C.c_void_p]
L.lib.tru_debug_set_flags.argtypes = [C.c_int]


def run(M, K, N, flags):
    torch.manual_seed(1)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    L.lib.tru_debug_set_flags(flags)
    L.check(fn(x.data_ptr(), None, None, w.data_ptr(), b.data_ptr(), out.data_ptr(), None, M, K, N, 1, None), "debug_pw")
    torch.cuda.synchronize()
    L.lib.tru_debug_set_flags(0)
    ref = x.double() @ w.double().t() + b.double()
    return ((out.double() - ref).abs().max() / ref.abs().max()).item()


for shape in ((128 * 300 + 17, 128, 128), (5000, 192, 64)):
    for flags, name in ((0, "no shift"), (64, "shift 1 row, base offset 0"), (64 | 128, "shift 1 row, base offset 1")):
        print("M=%d K=%d N=%d  %-28s max rel err %.3e" % (shape + (name, run(*shape, flags))), flush=True)
