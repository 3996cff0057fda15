"""Hardware probe for the tcgen05 descriptor conventions: runs every selftest variant (including the
LBO/SBO-swapped encoding) and prints the relative error of each, without stopping at failures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_tile, expected

for swap in (0,):  # swap=1 (LBO/SBO exchanged) faults on hardware: the tc05.cuh convention is the right one
    for a_mn, b_mn in [(0, 0), (1, 1), (0, 1), (1, 0)]:
        for N, K, shift in [(160, 304, 0), (160, 304, -4), (160, 304, 4), (144, 160, 1), (256, 64, 0), (16, 16, 0)]:
            g = torch.Generator(device="cuda").manual_seed(1)
            a = torch.randn((K, 128) if a_mn else (128, K), generator=g, device="cuda")
            b = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda")
            try:
                d = run_tile(a, b, a_mn, b_mn, N, K, shift, 8, swap)
                ref = expected(a, b, a_mn, b_mn, N, K, shift)
                err = float((d.double() - ref).norm() / ref.norm())
            except Exception as e:  # noqa
                err = repr(e)
            print("swap=%d a_mn=%d b_mn=%d N=%d K=%d shift=%d -> %s" % (swap, a_mn, b_mn, N, K, shift, err), flush=True)
