"""bench.py contract pieces that run without a GPU: the reference arm (the reference's algorithm on host cores)
prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "3",
                        "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_s" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0
    # the unmodified reference when it is on this machine (/root/reference here, oracle/_ref on the GPU box), else the port
    staged = os.path.isdir("/root/reference/mini_gym") or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "mini_gym"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["same_config"] is True and d["config"]["envs_per_gpu"] == 32768
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "3"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_ppo_arm_json_line():
    """`ppo.reference` legs of bench.py: the unmodified reference learner (compute_returns + update) on host cores."""
    staged = os.path.isdir("/root/reference/mini_gym") or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "mini_gym"))
    if not staged:
        import pytest
        pytest.skip("reference not staged on this machine")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference-ppo", "--ppo-envs", "64",
                        "--ref-device", "cpu", "--ref-epochs", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert d["unit"] == "samples/s" and d["value"] > 0 and d["device"] == "cpu" and d["epochs_run"] == 1 and len(d["losses"]) == 3
