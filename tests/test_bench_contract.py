"""CPU checks of bench.py's contract: the reference arm (the reference's own TwoTower on the host cores) prints one JSON line with
this arm's metric / unit / config, and both arms build their `config` from the same function."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(out):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_one_line_on_this_arms_config():
    sys.path.insert(0, ROOT)
    import bench
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)
    assert d["impl"] == "reference" and d["metric"] == "train_impressions_per_sec" and d["unit"] == "impressions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    args = types.SimpleNamespace(config=2, gpus=1, ddp=False)
    assert d["config"] == bench.config_dict(args, bench.CONFIGS[2])          # what the GPU arm prints for the same flags
    assert d["e2e"] == {"value": d["value"], "unit": "impressions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and "256-impression" in cb["sample"]
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "models")):
        assert cb["kind"] == "reference"                                      # the staged, unmodified reference was what ran


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_algorithmic_work_figures_are_the_surveys():
    """the figures `roofline` / `step_roofline` are computed from are SURVEY.md 8(d)'s (padding and recompute not counted):
    1.720 GFLOP per training impression at config 2, 1.684 with the MHA user encoder, 5.08 at config 5; 270,000 FLOP per token in
    the conv forward; the peaks come from the driver-written MEASURED_PEAKS.json when present, else the profiling guide's fallback."""
    sys.path.insert(0, ROOT)
    import bench
    assert abs(bench.flop_per_impression(bench.CONFIGS[2]) / 1e9 - 1.7204) < 1e-3
    assert abs(bench.flop_per_impression(bench.CONFIGS[3]) / 1e9 - 1.684) < 2e-3
    assert abs(bench.flop_per_impression(bench.CONFIGS[5]) / 1e9 - 5.08) < 1e-2
    assert bench.FLOP_PER_TOKEN_CONV == 2 * 3 * 300 * 150
    pk = bench.peaks()
    assert set(pk) == {"hbm", "tf_burst", "tf_sustained", "source"} and pk["source"] in ("measured", "fallback")
    if pk["source"] == "measured":
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        assert (pk["hbm"], pk["tf_burst"], pk["tf_sustained"]) == (m["hbm_gbs"], m["bf16_tflops"], m["bf16_tflops_sustained"])
    assert 5000 < pk["hbm"] < 8000 and pk["tf_sustained"] <= pk["tf_burst"] < 2300
