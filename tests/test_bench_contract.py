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
