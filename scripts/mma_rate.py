"""Cycles per tcgen05.mma (M=128, K=16, bf16, one CTA, operands resident) for several N."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from news_recommendation_mind_b200 import _lib
lib = _lib.load()
lib.mr_debug_mma_rate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for layout in (0, 2):
    for N in (64, 80, 96, 128, 144, 160, 176, 192, 208, 224, 240, 256):
        it = 2000
        for _ in range(2):
            lib.mr_debug_mma_rate(N, it, layout, ctypes.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
        print("b_layout %d N %3d: %.1f cycles/MMA  (%.0f flop/cycle)" % (layout, N, out.item() / it, 128 * N * 16 * 2 / (out.item() / it)))
