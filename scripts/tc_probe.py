"""Hardware probe for tcgen05 shared-memory descriptor conventions: runs selftest variants (operand majorness,
row shift, swizzle mode, base-offset handling) and prints the relative error of each without stopping."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_tile, expected

def one(a_mn, b_mn, N, K, shift, al, bl, bom):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn((K + 16, 128) if a_mn else (128, K), generator=g, device="cuda")
    b = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda")
    try:
        d = run_tile(a, b, a_mn, b_mn, N, K, shift, 8, 0, al, bl, bom)
        ref = expected(a, b, a_mn, b_mn, N, K, shift)
        err = "%.2e" % float((d.double() - ref).norm() / ref.norm())
    except Exception as e:  # noqa
        err = repr(e)[:80]
    print("a_mn=%d b_mn=%d N=%3d K=%3d shift=%2d a_layout=%d b_layout=%d base_off=%d -> %s" % (a_mn, b_mn, N, K, shift, al, bl, bom, err), flush=True)

for bom in (0, 1):
    for shift in (0, 8, 4, -4, 1):
        one(0, 0, 160, 304, shift, 2, 0, bom)      # K-major SW128 A (conv fwd / dgrad)
        one(1, 1, 160, 128, shift, 2, 0, bom)      # MN-major SW128 A (weight gradients)
for bl, N in ((2, 128), (2, 192), (4, 160), (6, 160), (4, 96)):
    one(1, 1, N, 128, 4, 2, bl, 0)                 # MN-major swizzled B
    one(0, 0, N, 128, 4, 2, bl, 0)                 # K-major swizzled B
