"""Static check of the compiled kernels (no GPU needed): no non-coherent global load may sit above griddepcontrol.wait.

Every kernel of the hot chain is launched with programmatic dependent launch and calls pdl_wait() (SASS: ACQBULK) before its
first global access.  ptxas does not order ld.global.nc (SASS: LDG...CONSTANT -- what loads through `const T* __restrict__`
compile to) against that wait; in round 2 it had hoisted the first W_hh loads of rnn_mma_fwd_kernel<*, 4|8> above it, i.e.
above the completion of the kernel that writes the buffer (intermittent wrong LSTM states at >= 4 sequences per CTA; fixed with
pdl_acquire, csrc/common.cuh).  This test disassembles every object of the library and fails if that pattern re-appears."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "news_recommendation_mind_b200", "build")


def test_no_noncoherent_load_above_griddepcontrol_wait():
    from news_recommendation_mind_b200 import build
    build.build()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    offenders, kernels_with_wait = [], 0
    for obj in sorted(f for f in os.listdir(BUILD) if f.endswith(".o")):
        sass = subprocess.run([exe, "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        fn, seen, early = None, False, 0
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn, seen, early = m.group(1), False, 0
                continue
            if "ACQBULK" in line:
                if not seen:
                    kernels_with_wait += 1
                    if early:
                        offenders.append((obj, fn, early))
                seen = True
            elif not seen and re.search(r"LDG[^;]*CONSTANT", line):
                early += 1
    assert kernels_with_wait > 20, kernels_with_wait          # the check looked at the kernels it is meant for
    assert not offenders, offenders
