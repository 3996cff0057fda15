"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol that
include/mindrec.h declares with matching arity, and refuses to compute without a B200."""
import os
import re

import pytest
import torch

import news_recommendation_mind_b200 as mr
from news_recommendation_mind_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "mindrec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"MR_API\s+[\w\s\*]+?\b(mr_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.mr_version() >= 100


def test_every_declared_symbol_is_exported_with_matching_arity():
    build.build()
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 30
    assert set(declared) == set(_lib.exported_symbols())
    for name, nargs in declared.items():
        fn = getattr(lib, name)                      # raises if not exported
        assert len(fn.argtypes) == nargs, (name, len(fn.argtypes), nargs)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.mr_device_check(0) != 0
    assert b"not present" in lib.mr_last_error()
    # product modules refuse CPU tensors loudly instead of silently computing on the host
    from helpers import manager_for, build_model
    man = manager_for("cnn", "lstm", 3, 4, 8, 16, 8, 2, device="cpu")
    model = build_model(man, 50)
    x = {"cdd_encoded_index": torch.ones(2, 3, 8, dtype=torch.long), "cdd_attn_mask": torch.ones(2, 3, 8, dtype=torch.long),
         "his_encoded_index": torch.ones(2, 4, 8, dtype=torch.long), "his_attn_mask": torch.ones(2, 4, 8, dtype=torch.long),
         "his_mask": torch.ones(2, 4, 1, dtype=torch.float64), "user_id": torch.ones(2, dtype=torch.long)}
    with pytest.raises(RuntimeError):
        model(x)


def test_state_dict_keys_match_reference(golden):
    from helpers import manager_for, build_model
    for name, encn, encu in [("tt_cnn_lstm", "cnn", "lstm"), ("tt_cnn_gru", "cnn", "gru"), ("tt_cnn_attn", "cnn", "attn"),
                             ("tt_mha_lstm", "mha", "lstm"), ("tt_cnn_lstur", "cnn", "lstur")]:
        g = golden(name)
        B, C, S, L, E, H, V, hn = [int(v) for v in g["meta"]]
        model = build_model(manager_for(encn, encu, C, S, L, E, H, hn, device="cpu"), V)
        ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        ref = {k: tuple(v.shape) for k, v in g["params"].items()}
        assert ours == ref, (name, set(ours) ^ set(ref))
        assert model.name == "twotower__%s__%s" % (encn, encu)


def test_workspace_planning_runs_without_a_gpu_and_is_bounded():
    """the *_workspace_bytes / *_plan_bytes entry points are pure host arithmetic (caller-owned buffers): callable here,
    positive and modest at the BASELINE shapes (HBM is for activations and tables, not scratch), -1 on impossible shapes"""
    import ctypes
    from news_recommendation_mind_b200._lib import CnnShape, RnnShape
    lib = _lib.load()
    MB = 1 << 20
    for N, L in ((256 * 55, 32), (256 * 110, 48), (32768, 32)):           # config 2/3 step, config 5 step, one evaluation chunk
        for precision in (0, 1):
            s = CnnShape(N, L, 300, 150, 30522, precision)
            fwd = lib.mr_news_cnn_workspace_bytes(ctypes.byref(s), 0)
            bwd = lib.mr_news_cnn_workspace_bytes(ctypes.byref(s), 1)
            assert 0 < fwd <= bwd < 8192 * MB, (N, L, precision, fwd, bwd)
        s = CnnShape(N, L, 300, 150, 30522, 1)
        assert 0 < lib.mr_news_cnn_bwd_table_workspace_bytes(ctypes.byref(s)) < 8192 * MB
        assert 0 < lib.mr_token_group_plan_bytes(N * L, 30522) < 64 * MB
        assert 0 < lib.mr_embed_grad_workspace_bytes(N * L, 300, 30522) < 1024 * MB
    assert lib.mr_token_group_plan_bytes(-1, 30522) == -1 and lib.mr_token_group_plan_bytes(1 << 31, 30522) == -1
    assert lib.mr_embed_grad_workspace_bytes(10, 0, 30522) == -1
    for kind in (0, 1):
        for precision in (0, 1):
            r = RnnShape(256, 50, 150, kind, 0, precision)
            fwd, bwd = lib.mr_rnn_workspace_bytes(ctypes.byref(r), 0), lib.mr_rnn_workspace_bytes(ctypes.byref(r), 1)
            assert 0 < fwd and 0 < bwd < 1024 * MB, (kind, precision, fwd, bwd)
    assert lib.mr_linear_workspace_bytes(1024, 150, 300) >= 0
    assert lib.mr_linear_tc_workspace_bytes(256 * 110 * 48, 300, 300, 30522, 1) > 0


def test_header_is_plain_c_and_warning_free():
    """include/mindrec.h is the binding surface of a C / cgo / JNI consumer: it must compile as C (not only as C++ inside nvcc)
    with -Wall -Wextra -Werror, and every prototype must be usable from a C translation unit."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "mindrec.h")
    src = '#include "%s"\nint (*probe_a)(void) = mr_version;\nconst char* (*probe_b)(void) = mr_last_error;\nint main(void) { return MR_OK; }\n' % hdr
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", "-"], input=src,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_a_c_program_links_and_calls_the_library(tmp_path):
    """the boundary from a consumer that is not Python: a C99 program including mindrec.h, linked against libmindrec.so, reads the
    version and -- on a machine without a B200 -- gets a negative status with a reason from a compute entry point instead of a crash
    or a silent host fallback"""
    import shutil
    import subprocess
    from news_recommendation_mind_b200 import build
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lib = build.build()
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include "mindrec.h"\n'
                   "int main(void) {\n"
                   "  float out[4] = {0};\n"
                   "  int rc = mr_embed_gather_f32(0, 1, 0, out, 1, 4, 8, 0);\n"
                   '  printf("%d %d %d %s\\n", mr_version(), mr_device_check(0), rc, mr_last_error());\n'
                   "  return 0;\n}\n")
    exe = tmp_path / "probe"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        lib, "-Wl,-rpath," + os.path.dirname(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    version, dev_rc, rc, reason = out.stdout.strip().split(" ", 3)
    assert int(version) >= 1
    if not torch.cuda.is_available():
        assert int(dev_rc) < 0 and int(rc) < 0 and reason
    else:
        assert int(rc) < 0 and reason                                  # null pointers are refused, nothing is launched


def test_only_the_checkers_import_the_oracle():
    """oracle/ is test infrastructure: the product package never imports it (a product path through the oracle would void every
    parity claim); outside tests/ only bench.py (its `cpu_baseline` / `--impl reference` legs) and __graft_entry__.py (smoke(), and
    build() staging the checker) may."""
    import ast
    allowed = {"bench.py", "__graft_entry__.py"}
    offenders = []
    for base, dirs, files in os.walk(ROOT):
        rel = os.path.relpath(base, ROOT)
        dirs[:] = [d for d in dirs if not d.startswith(".") and d not in ("__pycache__", "gpurun_out", "build", "_ref", "baseline")]
        if os.path.islink(base) or rel.split(os.sep)[0] in ("tests", "oracle"):
            continue
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(base, f)
            tree = ast.parse(open(path).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.module:
                    names = [node.module]
                if any(n == "oracle" or n.startswith("oracle.") for n in names) and os.path.relpath(path, ROOT) not in allowed:
                    offenders.append(os.path.relpath(path, ROOT))
    assert not offenders, offenders
    # bench.py: the oracle only inside the two CPU legs
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, ast.ImportFrom) and n.module == "oracle" for n in ast.walk(fn))
        if uses:
            assert fn.name in ("_reference_available", "cpu_train_baseline", "cpu_eval_baseline", "run_reference"), fn.name
