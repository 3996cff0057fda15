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


def _sass_by_function(obj):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    out, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            out[fn] = []
        elif fn is not None:
            out[fn].append(line)
    return {k: "\n".join(v) for k, v in out.items()}


def test_hot_kernels_use_tcgen05_tmem_and_tma():
    """the tensor path is Blackwell's own (north_star items 2, 3): every hot kernel issues tcgen05.mma (SASS UTCHMMA), reads its
    accumulators from tensor memory (LDTM), and the operand movers are TMA (UTMALDG) / bulk copies (UBLKCP); the conv forward runs
    as CTA pairs (UTCHMMA.2CTA, multicast commit); the LSTM recurrences take W_hh from tensor memory (an MMA with a tmem A operand); none of them
    falls back to the legacy warp-level HMMA pipe."""
    from news_recommendation_mind_b200 import build
    build.build()
    want = {
        "tapgemm2.o": ("tapgemm2_kernel", ["UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "LDTM", "UBLKCP"]),     # A rows are GATHERED (token ids): cp.async
        "cnn_tail.o": ("cnn_tail_fwd_kernel", ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP"]),
        "tokred.o": ("tokred_kernel", ["UTCHMMA", "LDTM", "UTMALDG"]),
        "tapgemm.o": ("tapgemm_kernel", ["UTCHMMA", "LDTM"]),
        "rnn_tc.o": ("rnn_tc_fwd_kernel", ["UTCHMMA", "LDTM", "STTM"]),
    }
    for obj, (kernel, mnemonics) in want.items():
        fns = {k: v for k, v in _sass_by_function(obj).items() if kernel in k}
        assert fns, (obj, kernel)
        for name, text in fns.items():
            for mn in mnemonics:
                assert mn in text, (obj, name, mn)
            assert not re.search(r"\bHMMA\.", text), (obj, name, "legacy mma.sync in a tcgen05 kernel")
    for kernel in ("cnn_tail_bwd_kernel", "rnn_tc_bwd_kernel"):
        obj = "cnn_tail.o" if "tail" in kernel else "rnn_tc.o"
        fns = {k: v for k, v in _sass_by_function(obj).items() if kernel in k}
        assert fns and all("UTCHMMA" in t and "LDTM" in t for t in fns.values()), kernel


def test_hot_kernels_fit_their_register_and_stack_budget():
    """the persistent tcgen05 kernels are sized against the register file (threads x registers <= 64 K per SM, one CTA per SM) and
    must not spill: a spill shows up as a stack frame beyond the few bytes ptxas keeps for the 64-bit address arithmetic of the
    tensor-map / descriptor helpers, or as a LOCAL segment.  Budgets = threads per CTA of the launch (csrc/*.cu) -> registers."""
    from news_recommendation_mind_b200 import build
    build.build()
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    # kernel -> (object, threads per CTA as launched, stack bytes tolerated)
    budget = {
        "cnn_tail_fwd_kernel": ("cnn_tail.o", 14 * 32, 64),
        "cnn_tail_bwd_kernel": ("cnn_tail.o", 18 * 32, 64),
        "tapgemm2_kernel": ("tapgemm2.o", 448, 64),
        "tokred_kernel": ("tokred.o", 288, 64),
        "rnn_tc_fwd_kernel": ("rnn_tc.o", 512, 32),
        "rnn_tc_bwd_kernel": ("rnn_tc.o", 512, 32),
    }
    for kernel, (obj, threads, stack_max) in budget.items():
        txt = subprocess.run([exe, "--dump-resource-usage", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        rows = re.findall(r"Function (\S*%s\S*):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)" % kernel, txt)
        assert rows, (kernel, obj)
        for name, reg, stack, _shared, local in rows:
            assert int(local) == 0, (name, "local memory", local)
            assert int(stack) <= stack_max, (name, "stack frame (spill?)", stack)
            assert int(reg) * threads <= 65536, (name, reg, threads)
