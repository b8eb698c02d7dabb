"""The reference arm of bench.py (the reference's own CPU path timed on the host cores: the unmodified modules staged in
baseline/_ref, or the oracle port where that copy is absent) must keep printing the JSON line the driver parses -- it is the
one place outside tests/ and smoke() that may execute oracle/."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "videos/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] == 0
    staged = os.path.exists(os.path.join(REPO, "baseline", "_ref", "baselines", "learned_models.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    sys.path.insert(0, REPO)
    import bench
    assert line["config"] == bench.config_for(1)      # the same config dictionary as the GPU arm prints


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=REPO, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
